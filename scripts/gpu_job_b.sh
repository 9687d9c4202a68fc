#!/usr/bin/env bash
# Round-2 GPU job B: backward / losses / optimiser kernels against the oracles, then the full GPU suite.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_backward.py -q -s > gpurun_out/b_backward.log 2>&1; echo "rc=$?" >> gpurun_out/b_backward.log
timeout 600 python bench.py --workload train --steps 5 --reps 3 > gpurun_out/b_bench_train.json 2> gpurun_out/b_bench_train.err; echo rc=$? >> gpurun_out/b_bench_train.err
tail -60 gpurun_out/b_backward.log
