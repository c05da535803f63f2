#!/bin/bash
# round 2, session ae: scratch and task sums laid out [tile][task][lay]... (a warp's successive layers contiguous)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3
python tools/gpu_sweep.py T170L60 "" "chunk=65536" "chunk=16384" 2>&1 | tee gpurun_out/r2ae_sweep.txt
python tools/gpu_sweep.py T85L40 "" 2>&1 | tee -a gpurun_out/r2ae_sweep.txt
python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a gpurun_out/r2ae_sweep.txt
