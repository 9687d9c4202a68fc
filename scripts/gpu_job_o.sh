#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python scripts/attn_variants.py > gpurun_out/o_attn.log 2>&1; echo "rc=$?" >> gpurun_out/o_attn.log
cat gpurun_out/o_attn.log
