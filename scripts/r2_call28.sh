#!/bin/bash
mkdir -p gpurun_out
{
  echo "== tests"
  timeout 1500 python -m pytest tests -q -m gpu -p no:warnings 2>&1 | tail -6
  echo "== bench"
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>gpurun_out/r2_bench_hoist.err | grep -E "^\{" > gpurun_out/r2_bench_hoist.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_hoist.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], 'launches', d.get('gpu_launches'), 'loss', d.get('last_loss'))" || tail -5 gpurun_out/r2_bench_hoist.err
  echo "== timeline"
  timeout 400 python scripts/step_timeline.py 2>&1 | grep -v -i Warn | tail -30
} > gpurun_out/r2_call28.txt 2>&1
cat gpurun_out/r2_call28.txt | cut -c1-240
