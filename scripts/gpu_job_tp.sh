#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python scripts/train_profile.py > gpurun_out/tp.log 2>&1; echo "rc=$?" >> gpurun_out/tp.log
tail -12 gpurun_out/tp.log
