#!/bin/bash
# ncu launch list of the default bench step, final round-2 state (per-launch times are cold-cache and serialised)
mkdir -p gpurun_out
timeout 270 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv \
  python bench.py --steps 1 --warmup 3 --pool 1 --batch 16 --no-cpu-baseline --no-ref-cuda --no-config1 > gpurun_out/r2_launches_final.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r2_launches_final.csv; tail -2 gpurun_out/r2_launches_final.log | cut -c1-300
