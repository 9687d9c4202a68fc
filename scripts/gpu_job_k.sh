#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python scripts/edge_variants.py > gpurun_out/k_variants.log 2>&1; echo "rc=$?" >> gpurun_out/k_variants.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest.log
cat gpurun_out/k_variants.log; tail -5 gpurun_out/k_pytest.log
