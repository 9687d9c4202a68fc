#!/usr/bin/env bash
# Round-2 GPU job E: evidence -- launch list, --set full of the edge stage and of the edge backward, compute-sanitizer logs.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches.csv python scripts/profile_step.py --iters 1 > gpurun_out/e_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_stage_tc --launch-skip 9 --launch-count 1 -o gpurun_out/e_edge_stage_tc -f python scripts/profile_step.py --iters 2 > gpurun_out/e_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_edge_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/e_bwd_edge -f python scripts/train_profile.py --ncu > gpurun_out/e_ncu3.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/e_train_launches.csv python scripts/train_profile.py --ncu > gpurun_out/e_ncu4.log 2>&1
for tool in memcheck synccheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool python scripts/sanitize_small.py 3 > gpurun_out/e_sanitize_fwd_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/e_sanitize_fwd_$tool.log
done
for tool in memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool python scripts/sanitize_train_small.py > gpurun_out/e_sanitize_train_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/e_sanitize_train_$tool.log
done
for f in gpurun_out/e_sanitize_*.log; do echo == $f; grep -E "ERROR SUMMARY|rc=|done" $f | tail -3; done
