#!/bin/bash
mkdir -p gpurun_out
{
  echo "== op parity"
  timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_steps.py -q -m gpu -p no:warnings 2>&1 | tail -8
  echo "== properties (full size)"
  timeout 900 python -m pytest tests/test_properties.py tests/test_gpu_model.py tests/test_gpu_frontend.py -q -m gpu -p no:warnings 2>&1 | tail -8
  echo "== profile native"
  timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 45 2>&1 | grep -v Warn
} > gpurun_out/r2_call6.txt 2>&1
grep -E "passed|failed|error|Error" gpurun_out/r2_call6.txt | head; grep -A46 "^# minkunet34" gpurun_out/r2_call6.txt | cut -c1-130
