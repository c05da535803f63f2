#!/bin/bash
# round 2, session al: 256-bit accesses in the column kernels -- k-table row slices read as LDG.256 groups (LW and SW), LW scratch
# written and read as two g-points per instruction (STG.256 / LDG.256).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2al_sweep.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3 | tee -a $O
python tools/gpu_sweep.py T170L60 "" "col_warps=16" 2>&1 | tee -a $O
python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a $O
