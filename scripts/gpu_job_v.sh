#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 400 python scripts/emd_trajectory.py 90 > gpurun_out/v_new.log 2>&1; echo "rc=$?" >> gpurun_out/v_new.log
EQD_LIB_PATH=$PWD/variants/libeqd_oldemd.so timeout 400 python scripts/emd_trajectory.py 90 > gpurun_out/v_old.log 2>&1; echo "rc=$?" >> gpurun_out/v_old.log
echo NEW; grep "^step" gpurun_out/v_new.log | awk '{print $2, $3}' | tr '\n' ' '; echo; echo OLD; grep "^step" gpurun_out/v_old.log | awk '{print $2, $3}' | tr '\n' ' '; echo; tail -3 gpurun_out/v_new.log | cut -c1-170; tail -3 gpurun_out/v_old.log | cut -c1-170
