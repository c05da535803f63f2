#!/bin/bash
# round 2, session p: ablations of lw_taumol (what bounds it?), mbarrier wait hint in lw_rtrn, 32-warp SW solver with the
# loop invariants in shared memory.  The library is rebuilt on the box per variant.
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2p_sweep.txt; }
echo "--- default; x3=1: sw_solver 32 warps/SM, invariants in smem; x3=2: 28 warps, invariants in smem" | tee gpurun_out/r2p_sweep.txt
sweep "" "x3=1" "x3=2"
for H in 200 2000 20000; do
  RRTMG_B200_DEFS="-DRRTMG_MBAR_HINT=$H" python mima_b200/build.py --force | tail -1
  echo "--- mbarrier.try_wait suspend hint $H ns" | tee -a gpurun_out/r2p_sweep.txt
  sweep ""
done
for A in 1 2 3; do
  RRTMG_B200_DEFS="-DRRTMG_ABLATE=$A" python mima_b200/build.py --force | tail -1
  echo "--- lw_taumol ablation $A (1: no global stores, 2: no table loads, 3: neither)" | tee -a gpurun_out/r2p_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
