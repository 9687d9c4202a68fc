#!/usr/bin/env bash
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 400 bash -c "$(declare -f run); run 8 29701 bench.py --gpus 8 --steps 20 --warmup 5 --no-residue-e2e --no-cpu-baseline" > gpurun_out/y8b_bench.json 2> gpurun_out/y8b_bench.err
python - <<'PY'
import json
for line in open('gpurun_out/y8b_bench.json'):
    if line.startswith('{'):
        d=json.loads(line); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'per-rank', [round(x,3) for x in d.get('per_rank_ms_per_step',[])], d['config']['pairs_per_gpu'])
PY
