#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log
timeout 300 python scripts/emd_trajectory.py 14 > gpurun_out/z_traj.log 2>&1; echo "rc=$?" >> gpurun_out/z_traj.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/z_train_launches.csv python scripts/train_profile.py --ncu > gpurun_out/z_ncu.log 2>&1
tail -2 gpurun_out/z_pytest.log; grep "^step" gpurun_out/z_traj.log | awk 'NR%4==1' | cut -c1-150; grep "ot_emd" gpurun_out/z_train_launches.csv | tail -1 | cut -c1-200
