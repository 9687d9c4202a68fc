#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -x > gpurun_out/u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/u_pytest.log
timeout 300 python scripts/emd_stats.py > gpurun_out/u_emd.log 2>&1; echo "rc=$?" >> gpurun_out/u_emd.log
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/u_train.log 2>&1; echo "rc=$?" >> gpurun_out/u_train.log
tail -5 gpurun_out/u_pytest.log; tail -4 gpurun_out/u_emd.log; tail -2 gpurun_out/u_train.log | cut -c1-400
