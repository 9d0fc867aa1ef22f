#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L2_128x128,L3_256x256
{
  for cfg in "" "B2S_TC_T=2" "B2S_TC4_DBG=15" "B2S_TC4_DBG=2" "B2S_TC4_DBG=1" "B2S_TC4_DBG=3" "B2S_TC4_DBG=4" "B2S_TC4_MERGE=0" "B2S_TC_STAGES=2" "B2S_TC_STAGES=4"; do
    echo "== batch 4 [$cfg]"
    env $cfg timeout 200 python scripts/conv_microbench.py --batch 4 --iters 5 --hash-order --layers $L | grep -E " fwd | dgrad " 
  done
} > gpurun_out/r2_call8.txt 2>&1
cat gpurun_out/r2_call8.txt | cut -c1-75
