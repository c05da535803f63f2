#!/bin/bash
# round 2, session k: table gathers of the two solvers with an L1 evict_last hint (A/B in one session; nvcc is on the box)
set -u
mkdir -p gpurun_out
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee gpurun_out/r2k_sweep.txt
RRTMG_B200_DEFS=-DRRTMG_TBL_EL python mima_b200/build.py --force | tail -1
echo "--- tables with L1::evict_last" | tee -a gpurun_out/r2k_sweep.txt
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a gpurun_out/r2k_sweep.txt
python tools/gpu_sweep.py T85L40 "" 2>&1 | tee -a gpurun_out/r2k_sweep.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "end_to_end or repeated" 2>&1 | tail -2
python mima_b200/build.py --force | tail -1
echo "--- default again" | tee -a gpurun_out/r2k_sweep.txt
python tools/gpu_sweep.py T85L40 "" 2>&1 | tee -a gpurun_out/r2k_sweep.txt
