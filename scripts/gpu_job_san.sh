#!/usr/bin/env bash
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool python scripts/sanitize_train_small.py > gpurun_out/san_train_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/san_train_$tool.log
done
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool python scripts/sanitize_graph_small.py > gpurun_out/san_graph_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/san_graph_$tool.log
done
for f in gpurun_out/san_*.log; do echo == $f; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|done" $f | tail -3; done
