#!/bin/bash
mkdir -p gpurun_out
L=L2_192x128,L3_128x128,L3_256x256,L3_384x256,L4_256x256
{
  for cfg in "B2S_WG_MIN_STAGES=3" "B2S_WG_MIN_STAGES=2"; do
    for b in 4 16; do
      echo "== batch $b [$cfg]"
      env $cfg timeout 300 python scripts/conv_microbench.py --batch $b --iters 5 --hash-order --layers $L | grep -E " wgrad "
    done
  done
} > gpurun_out/r2_call20.txt 2>&1
cat gpurun_out/r2_call20.txt | cut -c1-110
