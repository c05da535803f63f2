#!/bin/bash
# round 2, session w: fused LW column kernel -- streaming scratch accesses (L1::no_allocate loads, .cs stores) and the L1 prefetch, A/B
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2w_sweep.txt; }
: > gpurun_out/r2w_sweep.txt
for G in "1 1" "1 0" "0 0" "0 1"; do
  set -- $G
  RRTMG_B200_DEFS="-DLW_COL_STREAM=$1 -DLW_COL_PREFETCH=$2" python mima_b200/build.py --force | tail -1
  echo "--- lw_column: stream=$1 prefetch=$2" | tee -a gpurun_out/r2w_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
