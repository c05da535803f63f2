#!/bin/bash
# round 2, session ad: table gathers of a layer's g-points issued together (branch-free look-up in lw_column, reftra split in sw_column)
set -u
mkdir -p gpurun_out
: > gpurun_out/r2ad_sweep.txt
cp mima_b200/lib/librrtmg_b200.so /tmp/default.so
for V in base lwonly both; do
  cp mima_b200/lib/variants/$V.so mima_b200/lib/librrtmg_b200.so
  echo "--- variant $V (base: branches around the look-ups; lwonly: lw_column branch-free; both: + sw_column direct-beam gathers first)" | tee -a gpurun_out/r2ad_sweep.txt
  python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a gpurun_out/r2ad_sweep.txt
done
cp /tmp/default.so mima_b200/lib/librrtmg_b200.so
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
