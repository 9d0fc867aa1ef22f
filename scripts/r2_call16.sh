#!/bin/bash
mkdir -p gpurun_out
{
  echo "== all gpu tests"
  timeout 1800 python -m pytest tests -q -m gpu -p no:warnings 2>&1 | tail -6
  bash scripts/r2_call15.sh
  echo "== bench native"
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})"
  timeout 300 python scripts/membound_ops.py --batch 4 2>&1 | grep -E "bn_|devox"
  bash scripts/r2_sanitize.sh
} > gpurun_out/r2_call16.txt 2>&1
cat gpurun_out/r2_call16.txt | cut -c1-200 | tail -120
