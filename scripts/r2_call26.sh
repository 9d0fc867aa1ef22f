#!/bin/bash
mkdir -p gpurun_out
{
  echo "== new tests"
  timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_ref_models.py -q -m gpu -p no:warnings -k "backend_conv_names or range_lib or range_ops" 2>&1 | tail -15
  echo "== timeline"
  timeout 400 python scripts/step_timeline.py 2>&1 | grep -v Warn | tail -40
} > gpurun_out/r2_call26.txt 2>&1
cat gpurun_out/r2_call26.txt | cut -c1-240
