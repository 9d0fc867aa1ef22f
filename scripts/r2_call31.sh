#!/bin/bash
mkdir -p gpurun_out
{
  for b in 1; do
    python bench.py --batch $b --steps 10 --warmup 4 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_minkunet34_b$b.json
    python -c "
import json,sys; d=json.loads(open('gpurun_out/r2_bench_minkunet34_b$b.json').read()); print('batch $b value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
  done
} > gpurun_out/r2_call31.txt 2>&1
cat gpurun_out/r2_call31.txt
