#!/bin/bash
# round 2, session y: first run of the fused SW column kernel; LW column kernel with 20 / 24 warps per block
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2y_sweep.txt; }
: > gpurun_out/r2y_sweep.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
sweep "" "sw_fused=0"
python tools/gpu_sweep.py T42L40 "" "sw_fused=0" 2>&1 | tee -a gpurun_out/r2y_sweep.txt
bash tools/gpu_ncu_one.sh sw_column_kernel "" r2y_swcol
rm -f gpurun_out/*.ncu-rep
for W in 20 24; do
  RRTMG_B200_DEFS="-DLW_COL_WARPS=$W" python mima_b200/build.py --force | tail -1
  echo "--- lw_column: $W warps per block" | tee -a gpurun_out/r2y_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
