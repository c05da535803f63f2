#!/bin/bash
# round 2, session s: first run of the fused clear-sky LW column kernel -- GPU suite, then A/B against the staged kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python tools/gpu_sweep.py T170L60 "lw_fused=1" "lw_fused=0" 2>&1 | tee gpurun_out/r2s_sweep.txt
timeout 600 python tools/gpu_sweep.py T42L40 "lw_fused=1" "lw_fused=0" 2>&1 | tee -a gpurun_out/r2s_sweep.txt
