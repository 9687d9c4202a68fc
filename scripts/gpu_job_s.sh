#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_acceptance.py -q -x > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
timeout 300 python scripts/graph_build_profile.py > gpurun_out/s_gb.log 2>&1; echo "rc=$?" >> gpurun_out/s_gb.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.log 2>&1; echo "rc=$?" >> gpurun_out/s_bench.log
tail -5 gpurun_out/s_pytest.log; tail -2 gpurun_out/s_gb.log; python - <<'PY'
import json
for line in open('gpurun_out/s_bench.log'):
    if line.startswith('{'):
        d=json.loads(line); print('value', d['value'], 'e2e', d['e2e']['value'], 'from_residues', d['e2e_from_residues']['value'])
PY
