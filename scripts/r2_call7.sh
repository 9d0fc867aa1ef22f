#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L0_128x96,L1_96x96,L2_128x128,L3_256x256
{
  for bm in 0 1; do
    echo "== batch 16, B2S_TILE_BATCH_MAJOR=$bm"
    B2S_TILE_BATCH_MAJOR=$bm timeout 300 python scripts/conv_microbench.py --batch 16 --iters 4 --hash-order --layers $L
  done
  echo "== bench native (batch-major)"
  timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>&1 | grep -E "^\{"
} > gpurun_out/r2_call7.txt 2>&1
grep -E "^==|^L[0-4]|^#" gpurun_out/r2_call7.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2_call7.txt'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', d['value'], d['ms_per_step'], {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})
PY
