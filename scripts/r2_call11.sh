#!/bin/bash
# ncu evidence: full-set captures of the conv kernels, launch list of the bench step, DRAM bytes of the memory-bound ops
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"gather_gemm_tc4|wgrad_tc" -c 12 -o gpurun_out/r2_conv_full -f \
  python scripts/conv_microbench.py --batch 4 --once --hash-order --layers L0_96x96,L3_256x256 > gpurun_out/r2_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2_membound_ncu.csv python scripts/membound_ops.py --batch 4 --once > gpurun_out/r2_membound_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-ref-cuda --no-config1 > gpurun_out/r2_launches.log 2>&1
ls -la gpurun_out/r2_conv_full.ncu-rep gpurun_out/r2_membound_ncu.csv gpurun_out/r2_launches.csv
