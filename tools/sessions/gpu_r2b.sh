#!/bin/bash
# round 2, session b: why is binned taumol not faster, why is the L2-stack solver slower -- ncu of each
set -u
mkdir -p gpurun_out
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from mima_b200 import rrtmg
from mima_b200.columns import make_columns
rrtmg.set_device(0); rrtmg.rrtmg_lw_ini(); rrtmg.rrtmg_sw_ini()
cols = make_columns("T170L60", nlon=64, nlat=8, night=True)
rrtmg.set_option("sw_solver_variant", 4)
r4 = rrtmg.sw_from_columns(cols)
rrtmg.set_option("sw_solver_variant", 5)
for wpb, flags, ns in ((0, 3, 0), (12, 3, 0), (16, 1, 1), (20, 0, 0), (20, 1, 0), (20, 2, 0), (24, 2, 4), (28, 0, 0), (28, 0, 1)):
    for k, v in (("x0", wpb), ("x1", flags), ("x2", ns)):
        rrtmg.set_option(k, v)
    r5 = rrtmg.sw_from_columns(cols)
    print("v5", wpb, flags, ns, [float(np.max(np.abs(a - b))) for a, b in zip(r5, r4)], flush=True)
PY
run_ncu() { # name, tune, kernel regex, skip
  RRTMG_TUNE="$2" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c 1 -f -o gpurun_out/r2b_$1 \
     python bench.py --steps 1 --warmup 1 --workload T170L60 --no-cpu > gpurun_out/r2b_$1.log 2>&1
  ncu -i gpurun_out/r2b_$1.ncu-rep --page raw --csv > gpurun_out/r2b_$1_raw.csv 2>/dev/null
  rm -f gpurun_out/r2b_$1.ncu-rep
}
run_ncu lwtm_bin1 "taumol_bin=1" lw_taumol 2
run_ncu lwtm_bin0 "taumol_bin=0" lw_taumol 2
run_ncu swsolv_v5 "sw_solver_variant=5,x0=20,x1=3" sw_solver_l2 2
run_ncu swsolv_v4 "sw_solver_variant=4" sw_solver_warp 2
ls -la gpurun_out | tail
