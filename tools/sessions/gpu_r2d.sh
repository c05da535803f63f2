#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
W=T170L60
timeout 1500 python tools/gpu_sweep.py $W \
  "taumol_bin=0" \
  "taumol_bin=1,taumol_run=1" \
  "taumol_bin=1,taumol_run=2" \
  "taumol_bin=1,taumol_run=4" \
  "taumol_bin=1,taumol_run=8" \
  "taumol_bin=1,taumol_run=16" \
  "taumol_bin=1,taumol_run=32" \
  "taumol_bin=1,taumol_run=64" \
  "taumol_bin=1,taumol_run=8,taumol_order=1" \
  "taumol_bin=1,taumol_run=8,taumol_sync=16" \
  "taumol_bin=1,taumol_run=8,taumol_sync=1" \
  2>&1 | tee gpurun_out/r2d_sweep.txt
RRTMG_TUNE="taumol_bin=1,taumol_run=8" timeout 600 ncu --set full --clock-control none --import-source on -k regex:lw_taumol -s 2 -c 1 -f -o gpurun_out/r2d_lwtm \
     python bench.py --steps 1 --warmup 1 --workload T170L60 --no-cpu > gpurun_out/r2d_lwtm.log 2>&1
ncu -i gpurun_out/r2d_lwtm.ncu-rep --page raw --csv > gpurun_out/r2d_lwtm_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_lwtm.ncu-rep --page source --csv > gpurun_out/r2d_lwtm_source.csv 2>/dev/null
rm -f gpurun_out/r2d_lwtm.ncu-rep
