#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_ops.py 2>&1 | grep -vE "Warning|warn" | tail -25
done > gpurun_out/r2_sanitizer.txt 2>&1
cat gpurun_out/r2_sanitizer.txt | cut -c1-220
