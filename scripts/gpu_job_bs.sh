#!/usr/bin/env bash
mkdir -p gpurun_out
for b in 256 284 370 444; do
  timeout 300 python bench.py --pairs-per-gpu $b --steps 20 --warmup 5 --no-cpu-baseline --no-residue-e2e > gpurun_out/bs_$b.log 2>&1; echo "rc=$?" >> gpurun_out/bs_$b.log
done
python - <<'PY'
import json
for b in (256,284,370,444):
    for line in open(f'gpurun_out/bs_{b}.log'):
        if line.startswith('{'):
            d=json.loads(line); print(b, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'edge_ms', round(d['kernels_ms']['edge_stage'],4), 'node_ms', round(d['kernels_ms']['node_stage'],4), 'frac', round(d['roofline']['frac'],4))
PY
