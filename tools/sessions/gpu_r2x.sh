#!/bin/bash
# round 2, session x: fused LW column kernel -- tile-major setcoef state, next layer's state loaded behind the band formula (LW_COL_PIPE)
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2x_sweep.txt; }
: > gpurun_out/r2x_sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for G in 1 0; do
  RRTMG_B200_DEFS="-DLW_COL_PIPE=$G" python mima_b200/build.py --force | tail -1
  echo "--- lw_column: pipe=$G" | tee -a gpurun_out/r2x_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
bash tools/gpu_ncu_one.sh lw_column_kernel "" r2x_lwcol
rm -f gpurun_out/*.ncu-rep
