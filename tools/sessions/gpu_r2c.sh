#!/bin/bash
# round 2, session c: per-instruction stall attribution of lw_taumol (binned / not binned)
set -u
mkdir -p gpurun_out
run_ncu() { # name, tune, kernel regex, skip
  RRTMG_TUNE="$2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c 1 -f -o gpurun_out/r2c_$1 \
     python bench.py --steps 1 --warmup 1 --workload T170L60 --no-cpu > gpurun_out/r2c_$1.log 2>&1
  ncu -i gpurun_out/r2c_$1.ncu-rep --page source --csv > gpurun_out/r2c_$1_source.csv 2>/dev/null
  rm -f gpurun_out/r2c_$1.ncu-rep
}
run_ncu lwtm_bin1 "taumol_bin=1" lw_taumol 2
run_ncu lwtm_bin0 "taumol_bin=0" lw_taumol 2
ls -la gpurun_out | tail -5
