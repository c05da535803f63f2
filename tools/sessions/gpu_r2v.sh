#!/bin/bash
# round 2, session v: fused LW column kernel with the next layer's state prefetched into L1 and a pipelined up sweep
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2v_sweep.txt; }
: > gpurun_out/r2v_sweep.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "--- lw_column: 16 warps x 1 block" | tee -a gpurun_out/r2v_sweep.txt
sweep ""
bash tools/gpu_ncu_one.sh lw_column_kernel "" r2v_lwcol
rm -f gpurun_out/*.ncu-rep
for G in "8 2" "12 1"; do
  set -- $G
  RRTMG_B200_DEFS="-DLW_COL_WARPS=$1 -DLW_COL_BLOCKS=$2" python mima_b200/build.py --force | tail -1
  echo "--- lw_column: $1 warps per block, $2 blocks per SM (launch bounds)" | tee -a gpurun_out/r2v_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
