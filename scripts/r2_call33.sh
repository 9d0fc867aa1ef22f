#!/bin/bash
mkdir -p gpurun_out
{
  timeout 600 python -m pytest tests/test_gpu_steps.py tests/test_gpu_model.py -q -m gpu -p no:warnings 2>&1 | tail -4
  python bench.py --steps 12 --warmup 5 --no-cpu-baseline --no-ref-cuda --no-config1 2>gpurun_out/r2_bench_final.err | grep -E "^\{" > gpurun_out/r2_bench_final.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], 'launches', d.get('gpu_launches'))" || tail -5 gpurun_out/r2_bench_final.err
} > gpurun_out/r2_call33.txt 2>&1
cat gpurun_out/r2_call33.txt
