#!/bin/bash
mkdir -p gpurun_out
{
  echo "== step table + epilogue sums"
  timeout 300 python -m pytest tests/test_gpu_steps.py -x -q -p no:warnings 2>&1 | tail -15
  echo "== op parity"
  timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -p no:warnings 2>&1 | tail -15
  echo "== properties (full size)"
  timeout 900 python -m pytest tests/test_properties.py -x -q -m gpu -p no:warnings 2>&1 | tail -15
  echo "== microbench"
  timeout 300 python scripts/conv_microbench.py --batch 4 --iters 6 --hash-order
  echo "== bench native"
  timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-config1 2>&1 | grep -E "^\{"
} > gpurun_out/r2_call5.txt 2>&1
grep -E "passed|failed|error|Error" gpurun_out/r2_call5.txt | head; grep -E "^L[0-4]" gpurun_out/r2_call5.txt | head -60
