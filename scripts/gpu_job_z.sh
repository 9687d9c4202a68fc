#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -q -x > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log
EQD_LIB_PATH=$PWD/variants/libeqd_otprof.so timeout 300 python scripts/emd_stats.py > gpurun_out/z_prof.log 2>&1; echo "rc=$?" >> gpurun_out/z_prof.log
timeout 300 python scripts/emd_trajectory.py 30 > gpurun_out/z_traj.log 2>&1; echo "rc=$?" >> gpurun_out/z_traj.log
tail -3 gpurun_out/z_pytest.log; grep "OT_PROF\|rc=\|^step 2" gpurun_out/z_prof.log | tail -4 | cut -c1-260; grep "^step" gpurun_out/z_traj.log | awk 'NR%4==1' | cut -c1-150
