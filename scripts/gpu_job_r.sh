#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python scripts/graph_build_profile.py > gpurun_out/r_gb.log 2>&1; echo "rc=$?" >> gpurun_out/r_gb.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_gb_launches.csv python scripts/graph_build_profile.py 1 > gpurun_out/r_gb_ncu.log 2>&1
tail -3 gpurun_out/r_gb.log; python scripts/launch_summary.py gpurun_out/r_gb_launches.csv | tail -12
