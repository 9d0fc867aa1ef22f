#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read())
rc=d.get('ref_cuda') or {}
cb=d.get('cpu_baseline') or {}
print(sys.argv[1].split('/')[-1], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), 'frac', d.get('roofline',{}).get('frac') if d.get('roofline') else None, '| ref_cuda', rc.get('value'), '| cpu', cb.get('value'), '| cfg1', d.get('config1'))
" $1; }
{
  /usr/bin/time -f "default run wall %e s" python bench.py 2>gpurun_out/r2_bench_default.err | grep -E "^\{" > gpurun_out/r2_bench_minkunet34_native.json; tail -1 gpurun_out/r2_bench_default.err; show gpurun_out/r2_bench_minkunet34_native.json
  python bench.py --config minkunet34 --model-src reference --steps 8 --warmup 3 --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_minkunet34_refclass.json; show gpurun_out/r2_bench_minkunet34_refclass.json
  for cfg in spvcnn18 cylinder480 rpvnet34; do
    python bench.py --config $cfg --steps 8 --warmup 3 --no-config1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_$cfg.json; show gpurun_out/r2_bench_$cfg.json
  done
  python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | grep -E "^\{" > gpurun_out/r2_bench_reference_arm.json; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_reference_arm.json').read()); print('reference arm', d['value'], d['cpu_baseline']['sample'])"
  timeout 300 python scripts/profile_models.py --config rpvnet34 --top 22 2>&1 | grep -v Warn | grep -A24 "^# " | cut -c1-150 > gpurun_out/r2_prof_rpvnet34.txt; head -8 gpurun_out/r2_prof_rpvnet34.txt
} 2>&1 | tee gpurun_out/r2_call14.txt
