#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ot_emd_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/h_emd -f python scripts/train_profile.py --ncu > gpurun_out/h_ncu.log 2>&1
tail -3 gpurun_out/h_ncu.log; ls -la gpurun_out/h_emd.ncu-rep
