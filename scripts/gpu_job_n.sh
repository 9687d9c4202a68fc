#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python scripts/edge_variants.py > gpurun_out/n_variants.log 2>&1; echo "rc=$?" >> gpurun_out/n_variants.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/n_bench.log 2>&1; echo "rc=$?" >> gpurun_out/n_bench.log
cat gpurun_out/n_variants.log; tail -4 gpurun_out/n_pytest.log; tail -2 gpurun_out/n_bench.log | cut -c1-400
