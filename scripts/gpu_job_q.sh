#!/usr/bin/env bash
# Round-2 GPU job Q: evidence for the final forward kernels -- launch list, --set full of the edge stage, sanitizers, launch stress.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches.csv python scripts/profile_step.py --iters 1 > gpurun_out/q_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_stage_tc --launch-skip 9 --launch-count 1 -o gpurun_out/q_edge_stage_tc -f python scripts/profile_step.py --iters 2 > gpurun_out/q_ncu2.log 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool python scripts/sanitize_small.py 3 > gpurun_out/q_sanitize_fwd_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/q_sanitize_fwd_$tool.log
done
timeout 300 python bench.py --steps 3000 --warmup 5 --reps 3 --no-residue-e2e > gpurun_out/q_stress_bench.log 2>&1; echo "rc=$?" >> gpurun_out/q_stress_bench.log
for f in gpurun_out/q_sanitize_*.log; do echo == $f; grep -E "ERROR SUMMARY|rc=|done" $f | tail -3; done
tail -2 gpurun_out/q_stress_bench.log | cut -c1-300
