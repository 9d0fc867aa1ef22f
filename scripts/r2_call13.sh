#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L0_128x96,L1_96x96,L2_128x128,L3_256x256
{
  echo "== new tests"
  timeout 600 python -m pytest tests/test_gpu_steps.py tests/test_gpu_ops.py -q -m gpu -p no:warnings -x 2>&1 | tail -5
  for b in 4 16; do for ch in 0 1; do
    echo "== batch $b B2S_WGRAD_CHUNKED=$ch"
    B2S_WGRAD_CHUNKED=$ch timeout 300 python scripts/conv_microbench.py --batch $b --iters 4 --hash-order --layers $L | grep -E " wgrad "
  done; done
  echo "== bench native"
  timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})"
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 14 2>&1 | grep -v Warn
} > gpurun_out/r2_call13.txt 2>&1
cat gpurun_out/r2_call13.txt | cut -c1-150
