#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json; d=json.loads(open('$1').read())
print('$2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],2))"; }
{
  lscpu | grep -E "Model name|^CPU\(s\)|MHz" | head -4
  for h in 1 0 1; do
    B2S_HOIST_COORDS=$h python bench.py --steps 16 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_h$h.json
    show gpurun_out/r2_bench_h$h.json hoist=$h
  done
} > gpurun_out/r2_call29.txt 2>&1
cat gpurun_out/r2_call29.txt
