#!/bin/bash
# round 2, session ai: what bounds the column kernels -- ablations (results wrong on purpose, timing only): no scratch stores in the
# downward sweep; no upward sweep; neither.  Variant libraries built beforehand (sources restored afterwards).
set -u
mkdir -p gpurun_out
: > gpurun_out/r2ai_sweep.txt
cp mima_b200/lib/librrtmg_b200.so /tmp/default.so
for V in default nostore nopass2 compute; do
  [ $V = default ] && cp /tmp/default.so mima_b200/lib/librrtmg_b200.so || cp mima_b200/lib/variants/$V.so mima_b200/lib/librrtmg_b200.so
  echo "--- $V" | tee -a gpurun_out/r2ai_sweep.txt
  python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a gpurun_out/r2ai_sweep.txt
done
cp /tmp/default.so mima_b200/lib/librrtmg_b200.so
