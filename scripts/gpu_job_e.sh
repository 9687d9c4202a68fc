#!/usr/bin/env bash
# Round-2 GPU job E: evidence -- launch list, --set full of the edge stage, compute-sanitizer logs.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches.csv python scripts/profile_step.py --iters 1 > gpurun_out/e_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_stage_tc --launch-skip 9 --launch-count 1 -o gpurun_out/e_edge_stage_tc -f python scripts/profile_step.py --iters 2 > gpurun_out/e_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwd_edge_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/e_bwd_edge -f python scripts/train_profile.py --ncu > gpurun_out/e_ncu3.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py 4 > gpurun_out/e_sanitize_fwd_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/e_sanitize_fwd_$tool.log
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_train_small.py > gpurun_out/e_sanitize_train_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/e_sanitize_train_$tool.log
done
for f in gpurun_out/e_sanitize_*.log; do echo == $f; tail -3 $f; done
