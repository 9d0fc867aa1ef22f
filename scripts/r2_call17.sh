#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L0_128x96,L1_96x96,L2_128x128,L3_256x256
{
  for cfg in "B2S_TILE_SPATIAL=0" "B2S_TILE_SPATIAL=1" "B2S_TILE_SPATIAL=1 B2S_TILE_SPATIAL_BITS=7" "B2S_TILE_SPATIAL=1 B2S_TILE_SPATIAL_BITS=3" "B2S_TC_T=1"; do
    echo "== batch 16 [$cfg]"
    env $cfg timeout 300 python scripts/conv_microbench.py --batch 16 --iters 4 --hash-order --layers $L | grep -E " fwd | dgrad "
  done
  for cfg in "B2S_TILE_SPATIAL=0" "B2S_TILE_SPATIAL=1"; do
    echo "== bench [$cfg]"
    env $cfg timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>/dev/null | grep -E "^\{" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})"
  done
} > gpurun_out/r2_call17.txt 2>&1
cat gpurun_out/r2_call17.txt | cut -c1-100
