#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants" 2>&1 | tail -3
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from mima_b200 import rrtmg
from mima_b200.columns import make_columns
rrtmg.set_device(0); rrtmg.rrtmg_lw_ini(allow_synthetic_lw=True); rrtmg.rrtmg_sw_ini()
cols = make_columns("T170L60", nlon=64, nlat=8, night=True)
rrtmg.set_option("sw_solver_variant", 4); r4 = rrtmg.sw_from_columns(cols)
rrtmg.set_option("sw_solver_variant", 7)
for w in (20, 24, 28):
    rrtmg.set_option("x0", w); r = rrtmg.sw_from_columns(cols)
    print("v7", w, [float(np.max(np.abs(a - b))) for a, b in zip(r, r4)], flush=True)
PY
timeout 1500 python tools/gpu_sweep.py T170L60 \
  "sw_solver_variant=4" \
  "sw_solver_variant=7,x0=28" \
  "sw_solver_variant=7,x0=24" \
  "sw_solver_variant=7,x0=20" \
  2>&1 | tee gpurun_out/r2i_sweep.txt
RRTMG_TUNE="sw_solver_variant=7,x0=28" timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_solver_sm -s 2 -c 1 -f -o gpurun_out/r2i_v7 \
     python bench.py --steps 1 --warmup 1 --workload T170L60 --no-cpu > gpurun_out/r2i_v7.log 2>&1
ncu -i gpurun_out/r2i_v7.ncu-rep --page raw --csv > gpurun_out/r2i_v7_raw.csv 2>/dev/null
rm -f gpurun_out/r2i_v7.ncu-rep
