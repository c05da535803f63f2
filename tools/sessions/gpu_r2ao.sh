#!/bin/bash
# round 2, session ao: scratch of the lowest x4 layers of lw_column kept in L2 (plain write-back stores instead of evict-first ones,
# discard.global.L2 of every line the upward sweep has consumed)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2ao_sweep.txt
: > $O
RRTMG_TUNE=x4=8 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O
python tools/gpu_sweep.py T170L60 "" "x4=4" "x4=8" "x4=12" "x4=16" "x4=24" 2>&1 | tee -a $O
