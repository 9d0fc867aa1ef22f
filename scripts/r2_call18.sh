#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L0_128x96,L1_96x96,L2_64x64,L2_128x128,L3_256x256,L4_256x256
{
  echo "== all gpu tests"
  timeout 1800 python -m pytest tests -q -m gpu -p no:warnings 2>&1 | tail -6
  for b in 4 16; do
    echo "== microbench batch $b"
    timeout 300 python scripts/conv_microbench.py --batch $b --iters 5 --hash-order --layers $L | grep -E " fwd | dgrad |^#"
  done
  echo "== default bench (full line)"
  python bench.py 2>gpurun_out/r2_bench_default.err | grep -E "^\{" > gpurun_out/r2_bench_minkunet34_native.json
  python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_minkunet34_native.json').read())
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', d['roofline']['frac'], d['roofline']['per_family'])
print('ref_cuda', d.get('ref_cuda')); print('cpu', d.get('cpu_baseline')); print('cfg1', d.get('config1')); print('clocks', d.get('clocks'), 'launches', d.get('gpu_launches'))"
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 30 2>&1 | grep -v Warn | grep -A32 "^# "
} > gpurun_out/r2_call18.txt 2>&1
cat gpurun_out/r2_call18.txt | cut -c1-200
