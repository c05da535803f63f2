#!/bin/bash
# round 2, session ak: what the table reads cost the column kernels -- ablations (results wrong on purpose, timing only):
# rows1 = every k-table row read of the band sums becomes row 0 (loop-invariant, the loads leave the layer loop); rows2 = rows folded
# into the first 4 KB of the band table (always L1 hits); exp1 = the exp/tfn gathers read two fixed entries; rows1exp1 = both.
# Variant libraries built beforehand (-DABL_ROWS / -DABL_EXP, removed from the sources afterwards).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2ak_sweep.txt
: > $O
cp mima_b200/lib/librrtmg_b200.so /tmp/default.so
for V in default rows1 rows2 exp1 rows1exp1; do
  [ $V = default ] && cp /tmp/default.so mima_b200/lib/librrtmg_b200.so || cp mima_b200/lib/variants/$V.so mima_b200/lib/librrtmg_b200.so
  echo "--- $V" | tee -a $O
  python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a $O
done
cp /tmp/default.so mima_b200/lib/librrtmg_b200.so
