#!/usr/bin/env bash
# Multi-GPU job: data-parallel trainer check, inference + training bench at N GPUs (N = $1).
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run); N=$N run 29511 scripts/ddp_check.py" > gpurun_out/g${N}_ddp.log 2>&1; echo "rc=$?" >> gpurun_out/g${N}_ddp.log
timeout 900 bash -c "$(declare -f run); N=$N run 29512 bench.py --gpus $N --steps 20 --warmup 3" > gpurun_out/g${N}_bench.json 2> gpurun_out/g${N}_bench.err; echo "rc=$?" >> gpurun_out/g${N}_bench.err
timeout 900 bash -c "$(declare -f run); N=$N run 29513 bench.py --gpus $N --workload train --steps 10 --warmup 3 --reps 3" > gpurun_out/g${N}_bench_train.json 2> gpurun_out/g${N}_bench_train.err; echo "rc=$?" >> gpurun_out/g${N}_bench_train.err
grep -E "ddp_check|rc=" gpurun_out/g${N}_ddp.log | tail -3; tail -2 gpurun_out/g${N}_bench.err; head -c 600 gpurun_out/g${N}_bench.json; echo; head -c 400 gpurun_out/g${N}_bench_train.json
