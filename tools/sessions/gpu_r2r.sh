#!/bin/bash
# round 2, session r: lw_taumol store-layout ablations (timing only): 8 = {taug, fracs} interleaved, 12 = interleaved + band-major
set -u
mkdir -p gpurun_out
sweep() { python tools/gpu_sweep.py T170L60 "$@" 2>&1 | tee -a gpurun_out/r2r_sweep.txt; }
: > gpurun_out/r2r_sweep.txt
for A in 8 12 10 14; do
  RRTMG_B200_DEFS="-DRRTMG_ABLATE=$A" python mima_b200/build.py --force | tail -1
  echo "--- lw_taumol ablation $A (8: interleaved pairs, 4: band-major, 2: no table loads)" | tee -a gpurun_out/r2r_sweep.txt
  sweep ""
done
python mima_b200/build.py --force | tail -1
