#!/usr/bin/env bash
# final kernels on 8 GPUs of one box: inference bench at N = 8, 4, 2, 1 back to back (train at N = 8)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 8 4 2; do
  timeout 400 bash -c "$(declare -f run); run $n $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-residue-e2e --no-cpu-baseline" > gpurun_out/y${n}_bench.json 2> gpurun_out/y${n}_bench.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-residue-e2e > gpurun_out/y1_bench.json 2> gpurun_out/y1_bench.err
timeout 500 bash -c "$(declare -f run); run 8 29650 bench.py --gpus 8 --workload train --steps 10 --warmup 3 --reps 3 --no-cpu-baseline" > gpurun_out/y8_bench_train.json 2> gpurun_out/y8_bench_train.err
python - <<'PY'
import json
for f in ['y1_bench','y2_bench','y4_bench','y8_bench','y8_bench_train']:
    try:
        for line in open(f'gpurun_out/{f}.json'):
            if line.startswith('{'):
                d=json.loads(line); print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'per-rank', [round(x,3) for x in d.get('per_rank_ms_per_step',[])][:8])
    except Exception as e: print(f, 'ERR', e)
PY
