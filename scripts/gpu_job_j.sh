#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/j_pytest.log
timeout 600 python scripts/edge_knobs.py > gpurun_out/j_knobs.log 2>&1; echo "rc=$?" >> gpurun_out/j_knobs.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.log 2>&1; echo "rc=$?" >> gpurun_out/j_bench.log
tail -6 gpurun_out/j_pytest.log; cat gpurun_out/j_knobs.log; tail -2 gpurun_out/j_bench.log
