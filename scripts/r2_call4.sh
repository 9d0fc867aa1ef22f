#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ref_models.py -q -m gpu -s -p no:warnings > gpurun_out/r2_call4_tests.txt 2>&1
grep -E "^\[|passed|failed" gpurun_out/r2_call4_tests.txt | cut -c1-700
for cfg in rpvnet34 cylinder480 spvcnn18; do
  timeout 300 python scripts/profile_models.py --config $cfg --top 28 2>&1 | grep -v Warn > gpurun_out/r2_prof_$cfg.txt
done
timeout 300 python scripts/profile_models.py --config minkunet34 --model-src native --top 40 2>&1 | grep -v Warn > gpurun_out/r2_prof_minkunet34_native.txt
timeout 300 python scripts/profile_models.py --config minkunet34 --model-src reference --top 28 2>&1 | grep -v Warn > gpurun_out/r2_prof_minkunet34_ref.txt
timeout 300 python scripts/profile_models.py --config rpvnet34 --backend ref_cuda --top 20 2>&1 | grep -v Warn > gpurun_out/r2_prof_rpvnet34_b3.txt
head -12 gpurun_out/r2_prof_rpvnet34.txt
