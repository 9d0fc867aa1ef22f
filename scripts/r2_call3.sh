#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ref_models.py -q -m gpu -s -p no:warnings > gpurun_out/r2_call3_tests.txt 2>&1
grep -E "^\[|passed|failed" gpurun_out/r2_call3_tests.txt | cut -c1-900
timeout 600 python bench.py --config rpvnet34 --steps 4 --warmup 3 --no-cpu-baseline --no-config1 --no-ref-cuda 2>&1 | grep -E "^\{" > gpurun_out/r2_call3_rpv.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_call3_rpv.json').read()); print('rpvnet', d['value'], {k:(round(v['ms'],1),round(v['tflops'])) for k,v in d['roofline']['per_family'].items()})"
