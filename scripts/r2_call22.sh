#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L0_128x96,L1_96x96,L2_64x64,L2_128x128,L3_256x256,L4_256x256
{
  for cfg in "B2S_TC4_CTAS=2" "B2S_TC4_CTAS=3"; do
    for b in 4 16; do
      echo "== batch $b [$cfg]"
      env $cfg timeout 300 python scripts/conv_microbench.py --batch $b --iters 5 --hash-order --layers $L | grep -E " fwd | dgrad "
    done
  done
  echo "== conv tests with 3 CTAs"
  B2S_TC4_CTAS=3 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_steps.py tests/test_properties.py -q -m gpu -p no:warnings 2>&1 | tail -3
} > gpurun_out/r2_call22.txt 2>&1
cat gpurun_out/r2_call22.txt | cut -c1-110
