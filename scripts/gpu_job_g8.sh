#!/usr/bin/env bash
mkdir -p gpurun_out
bash scripts/gpu_job_g.sh 8 > gpurun_out/g8_all.log 2>&1
run4() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 bash -c "$(declare -f run4); run4 29531 bench.py --gpus 4 --steps 20 --warmup 3 --no-residue-e2e" > gpurun_out/g4_bench.json 2> gpurun_out/g4_bench.err
timeout 600 bash -c "$(declare -f run4); run4 29532 bench.py --gpus 4 --workload train --steps 10 --warmup 3 --reps 3" > gpurun_out/g4_bench_train.json 2> gpurun_out/g4_bench_train.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-residue-e2e > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err
nvidia-smi topo -m > gpurun_out/g8_topo.txt 2>&1
grep -E "ddp_check" gpurun_out/g8_ddp.log; for f in g8_bench g4_bench g1_bench g8_bench_train g4_bench_train; do grep '^{' gpurun_out/$f.json | head -c 300; echo; done
