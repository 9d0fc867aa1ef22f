#!/bin/bash
mkdir -p gpurun_out
{
  echo "== bn tests"
  timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_steps.py tests/test_gpu_syncbn.py tests/test_gpu_models.py -q -m gpu -p no:warnings -k "norm or bn or sums or minkunet or train" 2>&1 | tail -4
  echo "== membound (bn rows)"
  python scripts/membound_ops.py --batch 4 2>&1 | grep -E "^#|bn_|^op" 
  echo "== bench"
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_bn2.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_bn2.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'])"
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 24 2>&1 | grep -v Warn | grep -A26 "^# "
} > gpurun_out/r2_call24.txt 2>&1
cat gpurun_out/r2_call24.txt | cut -c1-220
