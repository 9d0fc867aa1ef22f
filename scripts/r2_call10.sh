#!/bin/bash
mkdir -p gpurun_out
{
  echo "== all gpu tests"
  timeout 1500 python -m pytest tests -q -m gpu -p no:warnings -x 2>&1 | tail -8
  echo "== bench native"
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>&1 | grep -E "^\{"
  echo "== profile native"
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 16 2>&1 | grep -v Warn
  echo "== membound ops"
  timeout 300 python scripts/membound_ops.py --batch 4 2>&1 | tail -60
} > gpurun_out/r2_call10.txt 2>&1
grep -E "passed|failed|error|Error" gpurun_out/r2_call10.txt | head; grep -A18 "^# minkunet34" gpurun_out/r2_call10.txt | cut -c1-130
grep -A60 "^# batch" gpurun_out/r2_call10.txt | cut -c1-140
python - <<'PY'
import json
for l in open('gpurun_out/r2_call10.txt'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})
PY
