#!/bin/bash
mkdir -p gpurun_out
{
  echo "== all gpu tests"
  timeout 1800 python -m pytest tests -q -m gpu -p no:warnings 2>&1 | tail -4
  echo "== smoke"
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  echo "== default bench (full line)"
  python bench.py 2>gpurun_out/r2_bench_default.err | grep -E "^\{" > gpurun_out/r2_bench_minkunet34_native.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_minkunet34_native.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})
print('ref_cuda', (d.get('ref_cuda') or {}).get('value'), (d.get('ref_cuda') or {}).get('e2e')); print('cpu', (d.get('cpu_baseline') or {}).get('value')); print('cfg1', d.get('config1')); print('clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))" || tail -5 gpurun_out/r2_bench_default.err
  for b in 1 12; do
    python bench.py --batch $b --steps 10 --warmup 4 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_minkunet34_b$b.json
    python -c "
import json,sys; d=json.loads(open('gpurun_out/r2_bench_minkunet34_b$b.json').read()); print('batch $b value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
  done
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 24 2>&1 | grep -v Warn | grep -A26 "^# "
  timeout 300 python scripts/step_timeline.py 2>&1 | grep -v -i Warn | tail -28
} > gpurun_out/r2_call30.txt 2>&1
cat gpurun_out/r2_call30.txt | cut -c1-260
