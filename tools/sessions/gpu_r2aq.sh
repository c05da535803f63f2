#!/bin/bash
# round 2, session aq: 20 / 24 warps per block (96 / 80 registers) for the column kernels now that the table rows come from shared memory
set -u
mkdir -p gpurun_out
O=gpurun_out/r2aq_sweep.txt
: > $O
python tools/gpu_sweep.py T170L60 "" "col_warps=20" "col_warps=24" "" 2>&1 | tee -a $O
