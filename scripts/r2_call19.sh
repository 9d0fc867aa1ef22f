#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read())
rc=d.get('ref_cuda') or {}
print(sys.argv[1].split('/')[-1], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'frac', d.get('roofline',{}).get('frac') if d.get('roofline') else None, '| ref_cuda', rc.get('value'))
" $1; }
{
  echo "== all gpu tests"
  timeout 1800 python -m pytest tests -q -m gpu -p no:warnings -s 2>&1 | grep -E "^\[|passed|failed|Error" | cut -c1-400
  for cfg in minkunet34 spvcnn18 cylinder480 rpvnet34; do
    python bench.py --config $cfg --model-src reference --steps 8 --warmup 3 --no-config1 --no-cpu-baseline --no-ref-cuda 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_${cfg}_ref_v2.json; show gpurun_out/r2_bench_${cfg}_ref_v2.json
  done
  python bench.py --steps 12 --warmup 5 --no-config1 --no-cpu-baseline --no-ref-cuda 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_native_v2.json; show gpurun_out/r2_bench_native_v2.json
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src reference --top 16 2>&1 | grep -v Warn | grep -A18 "^# " | cut -c1-150
} > gpurun_out/r2_call19.txt 2>&1
cat gpurun_out/r2_call19.txt
