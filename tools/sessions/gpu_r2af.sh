#!/bin/bash
# round 2, session af: block order of the column kernels -- super-groups of 4096 columns, longest task first inside one
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3
python tools/gpu_sweep.py T170L60 "" "chunk=65536" "chunk=32768" "chunk=16384" 2>&1 | tee gpurun_out/r2af_sweep.txt
python tools/gpu_sweep.py T85L40 "" 2>&1 | tee -a gpurun_out/r2af_sweep.txt
python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a gpurun_out/r2af_sweep.txt
