#!/bin/bash
# round 2, session an: setcoef state of the next layer prefetched into shared memory by cp.async (LW column kernel first)
set -u
mkdir -p gpurun_out
O=gpurun_out/r2an_sweep.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3 | tee -a $O
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a $O
python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a $O
