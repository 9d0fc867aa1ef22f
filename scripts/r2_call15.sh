#!/bin/bash
mkdir -p gpurun_out
L=L0_32x32,L0_96x96,L2_128x128,L3_256x256
{
  for cfg in "" "B2S_TC4_SLEEP_NS=0" "B2S_TC4_SLEEP_NS=256" "B2S_TC4_DBG=16" "B2S_TC4_DBG=15" "B2S_TC4_DBG=31" "B2S_TC4_DBG=31 B2S_TC4_SLEEP_NS=0"; do
    echo "== batch 4 [$cfg]"
    env $cfg timeout 200 python scripts/conv_microbench.py --batch 4 --iters 5 --hash-order --layers $L | grep -E " fwd " 
  done
} > gpurun_out/r2_call15.txt 2>&1
cat gpurun_out/r2_call15.txt | cut -c1-75
