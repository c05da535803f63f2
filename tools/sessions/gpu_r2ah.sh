#!/bin/bash
# round 2, session ah: SW task order by g-point count; super-group size 2048 / 4096 / 8192 columns
set -u
mkdir -p gpurun_out
: > gpurun_out/r2ah_sweep.txt
cp mima_b200/lib/librrtmg_b200.so /tmp/default.so
for V in default sg2048 sg8192; do
  [ $V = default ] && cp /tmp/default.so mima_b200/lib/librrtmg_b200.so || cp mima_b200/lib/variants/$V.so mima_b200/lib/librrtmg_b200.so
  echo "--- $V" | tee -a gpurun_out/r2ah_sweep.txt
  python tools/gpu_sweep.py T170L60 "" "chunk=16384" 2>&1 | tee -a gpurun_out/r2ah_sweep.txt
  python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a gpurun_out/r2ah_sweep.txt
done
cp /tmp/default.so mima_b200/lib/librrtmg_b200.so
