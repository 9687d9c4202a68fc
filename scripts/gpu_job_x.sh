#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/x_train.log 2>&1; echo "rc=$?" >> gpurun_out/x_train.log
python - <<'PY'
import json
for line in open('gpurun_out/x_train.log'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['rep_ms'], d['e2e']['value'], d['e2e'].get('rep_s'), d.get('cpu_baseline',{}).get('value'), d['clocks'])
PY
