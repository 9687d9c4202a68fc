#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "losses or trainer" > gpurun_out/f_losses.log 2>&1; echo "rc=$?" >> gpurun_out/f_losses.log
timeout 900 python -m pytest tests/test_gpu_acceptance.py -q -s -k "graph" > gpurun_out/f_acceptance.log 2>&1; echo "rc=$?" >> gpurun_out/f_acceptance.log
timeout 600 python scripts/train_profile.py > gpurun_out/f_train_phases.log 2>&1
tail -5 gpurun_out/f_losses.log; grep -E "passed|failed|Error" gpurun_out/f_acceptance.log | head; tail -6 gpurun_out/f_train_phases.log | cut -c1-1500
