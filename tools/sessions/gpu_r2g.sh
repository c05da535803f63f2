#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or end_to_end" 2>&1 | tail -5
timeout 1500 python tools/gpu_sweep.py T170L60 \
  "sw_solver_variant=4" \
  "sw_solver_variant=6,x0=16,x1=3" \
  "sw_solver_variant=6,x0=20,x1=3" \
  "sw_solver_variant=6,x0=20,x1=0" \
  "sw_solver_variant=6,x0=24,x1=3" \
  "sw_solver_variant=6,x0=28,x1=3" \
  "sw_solver_variant=6,x0=28,x1=0" \
  "sw_solver_variant=5,x0=20,x1=3" \
  2>&1 | tee gpurun_out/r2g_sweep.txt
RRTMG_TUNE="sw_solver_variant=6,x0=20,x1=3" timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_solver_sm -s 2 -c 1 -f -o gpurun_out/r2g_v6 \
     python bench.py --steps 1 --warmup 1 --workload T170L60 --no-cpu > gpurun_out/r2g_v6.log 2>&1
ncu -i gpurun_out/r2g_v6.ncu-rep --page raw --csv > gpurun_out/r2g_v6_raw.csv 2>/dev/null
ncu -i gpurun_out/r2g_v6.ncu-rep --page source --csv > gpurun_out/r2g_v6_source.csv 2>/dev/null
rm -f gpurun_out/r2g_v6.ncu-rep
