#!/bin/bash
# First gpurun call of round 2: validate and time the two TMA-gather variants written (but not run) in
# round 1.  Usage:  gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'   (results in gpurun_out/)
mkdir -p gpurun_out
{
  echo "== parity, default kernels"
  timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_properties.py -x -q -m gpu 2>&1 | tail -3
  echo "== parity, B2S_TC_GATHER4=1 (forward / input gradient through tile::gather4)"
  B2S_TC_GATHER4=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_properties.py -x -q -m gpu 2>&1 | tail -8
  echo "== parity, B2S_WG_GATHER4=1 (weight gradient through tile::gather4)"
  B2S_WG_GATHER4=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_properties.py -x -q -m gpu 2>&1 | tail -8
  echo "== microbench, default"
  timeout 120 python scripts/conv_microbench.py --batch 4 --iters 6 --hash-order
  echo "== microbench, both gather4 variants"
  B2S_TC_GATHER4=1 B2S_WG_GATHER4=1 timeout 120 python scripts/conv_microbench.py --batch 4 --iters 6 --hash-order
  echo "== bench, both gather4 variants (steps 6)"
  B2S_TC_GATHER4=1 B2S_WG_GATHER4=1 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline
} > gpurun_out/round2_first_call.txt 2>&1
tail -5 gpurun_out/round2_first_call.txt
