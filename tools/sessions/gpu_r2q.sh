#!/bin/bash
# round 2, session q: lw_rtrn ring geometries; lw_taumol writing a band-major staging layout (timing ablation)
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2q_sweep.txt; }
RRTMG_B200_DEFS="-DRRTMG_RING_EXPERIMENT" python mima_b200/build.py --force | tail -1
echo "--- lw_rtrn TMA ring: default 2 stages x 4 layers; x5=1: 4x2, 2: 3x2, 3: 3x3, 4: 6x1, 5: 5x2" | tee gpurun_out/r2q_sweep.txt
sweep "" "x5=1" "x5=2" "x5=3" "x5=4" "x5=5"
RRTMG_B200_DEFS="-DRRTMG_ABLATE=4" python mima_b200/build.py --force | tail -1
echo "--- lw_taumol writes [band][col][lay][ng] (timing only)" | tee -a gpurun_out/r2q_sweep.txt
sweep ""
RRTMG_B200_DEFS="-DRRTMG_ABLATE=6" python mima_b200/build.py --force | tail -1
echo "--- the same without table loads" | tee -a gpurun_out/r2q_sweep.txt
sweep ""
python mima_b200/build.py --force | tail -1
