#!/usr/bin/env bash
# Final check of the round: smoke(), the whole GPU suite, the default bench, the two other inference workloads, the reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/f_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log
timeout 600 python bench.py > gpurun_out/f_bench.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench.log
timeout 300 python bench.py --workload db5-testset --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_db5.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench_db5.log
timeout 300 python bench.py --workload large --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_large.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench_large.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/f_bench_ref.log
tail -3 gpurun_out/f_smoke.log; tail -3 gpurun_out/f_pytest.log
python - <<'PY'
import json
for f in ['f_bench','f_bench_db5','f_bench_large','f_bench_ref']:
    try:
        for line in open(f'gpurun_out/{f}.log'):
            if line.startswith('{'):
                d=json.loads(line); print(f, d.get('impl','b200'), 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', d.get('roofline',{}) and d['roofline'].get('frac'), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
