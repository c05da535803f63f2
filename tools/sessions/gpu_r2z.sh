#!/bin/bash
# round 2, session z: block shape of the fused kernels against the batch size (col_warps = 8: two 8-warp blocks per SM, 16: one 16-warp block)
set -u
mkdir -p gpurun_out
: > gpurun_out/r2z_sweep.txt
for W in T42L40 T85L40; do
  python tools/gpu_sweep.py $W "col_warps=8" "col_warps=16" 2>&1 | tee -a gpurun_out/r2z_sweep.txt
done
python tools/gpu_sweep.py T170L60 "col_warps=8,chunk=16384" "col_warps=16,chunk=16384" "col_warps=8,chunk=32768" "col_warps=16,chunk=32768" "col_warps=8" "col_warps=16" 2>&1 | tee -a gpurun_out/r2z_sweep.txt
