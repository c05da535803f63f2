#!/bin/bash
# round 2, session am: the k-distribution tables of the column kernels staged into shared memory -- per block one TMA bulk copy of the
# task's slice of its band table (rows x task g-points, odd stride in 16-byte units), read with LDS.128; one 16-warp block per SM.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2am_sweep.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3 | tee -a $O
python tools/gpu_sweep.py T170L60 "" "col_warps=8" 2>&1 | tee -a $O
python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a $O
python tools/gpu_sweep.py T85L40 "" 2>&1 | tee -a $O
