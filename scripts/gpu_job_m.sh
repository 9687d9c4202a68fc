#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 150 python scripts/edge_variants.py > gpurun_out/m_variants.log 2>&1; echo "rc=$?" >> gpurun_out/m_variants.log
cat gpurun_out/m_variants.log
