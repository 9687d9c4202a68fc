#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python scripts/emd_stats.py > gpurun_out/t_emd.log 2>&1; echo "rc=$?" >> gpurun_out/t_emd.log
tail -8 gpurun_out/t_emd.log
