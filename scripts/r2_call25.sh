#!/bin/bash
# final bench lines of BASELINE configs 2-5 driving the reference's own classes (with ref_cuda + cpu_baseline)
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.loads(open('$1').read())
print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'ref_cuda', (d.get('ref_cuda') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'))"; }
{
  for cfg in minkunet34 spvcnn18 cylinder480 rpvnet34; do
    timeout 700 python bench.py --config $cfg --model-src reference --steps 8 --warmup 3 --no-config1 2>gpurun_out/r2_final_$cfg.err | grep -E "^\{" > gpurun_out/r2_final_$cfg.json
    show gpurun_out/r2_final_$cfg.json || tail -5 gpurun_out/r2_final_$cfg.err
  done
} > gpurun_out/r2_call25.txt 2>&1
cat gpurun_out/r2_call25.txt | cut -c1-260
