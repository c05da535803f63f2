#!/bin/bash
# round 2, session u: block geometry of the fused LW column kernel (instruction-cache pressure); library rebuilt on the box
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2u_sweep.txt; }
: > gpurun_out/r2u_sweep.txt
for G in "16 1" "8 2" "8 1" "12 1" "4 4"; do
  set -- $G
  RRTMG_B200_DEFS="-DLW_COL_WARPS=$1 -DLW_COL_BLOCKS=$2" python mima_b200/build.py --force | tail -1
  echo "--- lw_column: $1 warps per block, $2 blocks per SM (launch bounds)" | tee -a gpurun_out/r2u_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
bash tools/gpu_ncu_one.sh lw_column_kernel "" r2u_lwcol
rm -f gpurun_out/*.ncu-rep
