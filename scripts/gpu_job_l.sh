#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python scripts/edge_variants.py > gpurun_out/l_variants.log 2>&1; echo "rc=$?" >> gpurun_out/l_variants.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:edge_stage_tc --launch-skip 3 -c 1 -f -o gpurun_out/l_edge31 python scripts/edge_one.py 31 > gpurun_out/l_ncu.log 2>&1; echo "rc=$?" >> gpurun_out/l_ncu.log
cat gpurun_out/l_variants.log; tail -3 gpurun_out/l_ncu.log
