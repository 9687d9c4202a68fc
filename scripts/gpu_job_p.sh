#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python scripts/forward_ab.py variants/libeqd_prev.so equidock_public_b200/libeqd_iegmn.so > gpurun_out/p_ab.log 2>&1; echo "rc=$?" >> gpurun_out/p_ab.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.log 2>&1; echo "rc=$?" >> gpurun_out/p_bench.log
tail -4 gpurun_out/p_ab.log; tail -3 gpurun_out/p_pytest.log; tail -2 gpurun_out/p_bench.log | cut -c1-300
