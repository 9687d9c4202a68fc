#!/usr/bin/env bash
# Round-2 GPU job A: parity table, GPU tests, the three inference bench workloads, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
python scripts/parity_table.py > gpurun_out/a_parity.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "rc=$?" >> gpurun_out/a_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cuda-graph --no-cpu-baseline > gpurun_out/a_bench_nograph.json 2> gpurun_out/a_bench_nograph.err
timeout 300 python bench.py --workload db5-testset --steps 20 --no-cpu-baseline > gpurun_out/a_bench_db5.json 2> gpurun_out/a_bench_db5.err
timeout 300 python bench.py --workload large --steps 10 --no-cpu-baseline > gpurun_out/a_bench_large.json 2> gpurun_out/a_bench_large.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --reps 1 --no-cpu-baseline --no-cuda-graph > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_parity.log | tail -12; cat gpurun_out/a_bench.json | head -c 1500
