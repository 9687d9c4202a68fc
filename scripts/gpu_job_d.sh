#!/usr/bin/env bash
# Round-2 GPU job D: acceptance on all 125 pairs (GPU graph build -> forward -> RMSD table), backward / losses / trainer
# tests, train bench and phase profile after the EMD solver rewrite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_acceptance.py -q -s > gpurun_out/d_acceptance.log 2>&1; echo "rc=$?" >> gpurun_out/d_acceptance.log
timeout 1200 python -m pytest tests/test_gpu_backward.py -q > gpurun_out/d_backward.log 2>&1; echo "rc=$?" >> gpurun_out/d_backward.log
timeout 600 python scripts/train_profile.py > gpurun_out/d_train_phases.log 2>&1
timeout 600 python bench.py --workload train --steps 10 --reps 3 > gpurun_out/d_bench_train.json 2> gpurun_out/d_bench_train.err; echo rc=$? >> gpurun_out/d_bench_train.err
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo rc=$? >> gpurun_out/d_bench.err
tail -15 gpurun_out/d_acceptance.log; tail -5 gpurun_out/d_backward.log; tail -4 gpurun_out/d_train_phases.log
