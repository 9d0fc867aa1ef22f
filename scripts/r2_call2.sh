#!/bin/bash
# Round-2 call 2: the reference's own segmentors on this backend (goldens + reference CUDA build), then one
# short bench line per BASELINE config with the reference CUDA arm measured beside it.
mkdir -p gpurun_out
{
  echo "== drop-in tests"
  timeout 900 python -m pytest tests/test_gpu_ref_models.py -q -m gpu -s 2>&1 | grep -v Warning | tail -40
  for cfg in minkunet34 spvcnn18 cylinder480 rpvnet34; do
    echo "== bench $cfg (reference class on this backend)"
    timeout 600 python bench.py --config $cfg --model-src reference --steps 6 --warmup 3 --no-cpu-baseline --no-config1 2>&1 | grep -E "^\{|Error|error" | tail -3
  done
  echo "== bench minkunet34 native"
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda 2>&1 | grep -E "^\{|Error|error" | tail -3
} > gpurun_out/r2_call2.txt 2>&1
tail -5 gpurun_out/r2_call2.txt
