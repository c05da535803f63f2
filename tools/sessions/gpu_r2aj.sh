#!/bin/bash
# round 2, session aj: the exp/tfn (LW) and exp/1/exp (SW) look-up tables of the column kernels staged into shared memory by
# TMA bulk copies (x2=1: LW, x3=1: SW; one 16-warp block per SM) against the L1 gathers of the default kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2aj_sweep.txt
: > $O
RRTMG_TUNE=x2=1,x3=1 timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee -a $O
python tools/gpu_sweep.py T170L60 "" "col_warps=16" "x2=1" "x3=1" "x2=1,x3=1" 2>&1 | tee -a $O
python tools/gpu_sweep.py T42L40 "" "x2=1,x3=1" 2>&1 | tee -a $O
