#!/bin/bash
# round 2, session ap: lw_column's scratch holds the optical depths and the recipe of the Planck fractions (8 bytes per g-point + 16
# per task and cell) instead of {absorptivity, upward source} (16 bytes per g-point); the upward sweep forms both again.  Plus the
# finish kernels with their 46 loads per level batched.  base.so = the finish change alone.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2ap_sweep.txt
: > $O
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -3 | tee -a $O
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a $O
python tools/gpu_sweep.py T42L40 "" 2>&1 | tee -a $O
cp mima_b200/lib/librrtmg_b200.so /tmp/new.so; cp mima_b200/lib/variants/base.so mima_b200/lib/librrtmg_b200.so
echo "--- base (finish kernels batched only)" | tee -a $O
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee -a $O
cp /tmp/new.so mima_b200/lib/librrtmg_b200.so
