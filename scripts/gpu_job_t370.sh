#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_stage_tc --launch-skip 9 --launch-count 1 -o gpurun_out/t370_edge -f python scripts/profile_step.py --iters 2 --pairs 370 > gpurun_out/t370_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/t370_ncu.log
tail -2 gpurun_out/t370_ncu.log
