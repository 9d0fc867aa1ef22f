#!/bin/bash
mkdir -p gpurun_out
{
  echo "== all gpu tests"
  timeout 1800 python -m pytest tests -q -m gpu -p no:warnings 2>&1 | tail -8
  for b in 1 12 16; do
    echo "== bench minkunet34 native batch $b"
    timeout 400 python bench.py --batch $b --steps 10 --warmup 4 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" | tee gpurun_out/r2_bench_minkunet34_b$b.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'])"
  done
  echo "== bench minkunet34 native fp32 batch 16"
  timeout 600 python bench.py --dtype fp32 --steps 4 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" | tee gpurun_out/r2_bench_minkunet34_fp32.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
} > gpurun_out/r2_call12.txt 2>&1
cat gpurun_out/r2_call12.txt | grep -v "^{" | tail -30
