#!/bin/bash
# round 2, session aa: fused SW kernel -- tasks of at most 4 g-points (32 tasks) and 16 / 20 / 24 warps per block against the default
# (at most 6 g-points, 16 warps).  Variant libraries are built beforehand (mima_b200/lib/variants) and copied over the default.
set -u
mkdir -p gpurun_out
: > gpurun_out/r2aa_sweep.txt
cp mima_b200/lib/librrtmg_b200.so /tmp/default.so
echo "--- default (n6, 16 warps; LW two 8-warp blocks)" | tee -a gpurun_out/r2aa_sweep.txt
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a gpurun_out/r2aa_sweep.txt
for V in n4w16 n4w20 n4w24 n6w20; do
  cp mima_b200/lib/variants/$V.so mima_b200/lib/librrtmg_b200.so
  echo "--- sw_column variant $V" | tee -a gpurun_out/r2aa_sweep.txt
  python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a gpurun_out/r2aa_sweep.txt
done
cp /tmp/default.so mima_b200/lib/librrtmg_b200.so
