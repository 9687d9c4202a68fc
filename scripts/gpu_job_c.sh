#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python scripts/train_profile.py > gpurun_out/c_train_phases.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_train_launches.csv python scripts/train_profile.py --ncu > gpurun_out/c_ncu.log 2>&1
cat gpurun_out/c_train_phases.log | tail -8
