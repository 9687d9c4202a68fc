#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_acceptance.py -q -s > gpurun_out/f_acceptance.log 2>&1; echo "rc=$?" >> gpurun_out/f_acceptance.log
timeout 600 python scripts/train_profile.py > gpurun_out/f_train_phases.log 2>&1
grep -E "passed|failed|worst|C-RMSD|oracle nbrs|gpu nbrs|differing" gpurun_out/f_acceptance.log | head -20; tail -6 gpurun_out/f_train_phases.log | cut -c1-1500
