#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/d370_bench.log 2>&1; echo "rc=$?" >> gpurun_out/d370_bench.log
python - <<'PY'
import json
for line in open('gpurun_out/d370_bench.log'):
    if line.startswith('{'):
        d=json.loads(line); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'from_res', round(d['e2e_from_residues']['value'],1), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:40], 'launches', d['gpu_launches'], d['config']['pairs_per_gpu'])
PY
tail -1 gpurun_out/d370_bench.log
