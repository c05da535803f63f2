"""Stress test of the TMA-fed lw_rtrn (developer tool): many calls, compared bit for bit with the direct-load kernel.
Needs a development build of the library (RRTMG_B200_DEV_VARIANTS=1 python -m mima_b200.build --force): the default build
carries the TMA-fed kernel only."""
import sys
import numpy as np
sys.path.insert(0, '.')
from mima_b200 import rrtmg
from mima_b200.columns import make_columns
rrtmg.set_device(0); rrtmg.rrtmg_lw_ini(allow_synthetic_lw=True); rrtmg.rrtmg_sw_ini()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for res, rows in (("T170L60", (48, 80)), ("T85L40", (0, 64)), ("T341L80", (100, 116))):
    c = make_columns(res, lat_rows=rows)
    rrtmg.set_option("lw_rtrn_variant", 1)
    ref = rrtmg.lw_from_columns(c)
    for v in (2, 3):
        rrtmg.set_option("lw_rtrn_variant", v)
        bad = 0
        for rep in range(reps):
            lw = rrtmg.lw_from_columns(c)
            for n, a, b in zip(("uflx", "dflx"), lw[:2], ref[:2]):
                d = np.abs(a - b) > 1e-9 * np.abs(b).max()
                if d.any():
                    bad += 1
                    cols, levs = np.nonzero(d)
                    print(res, "variant", v, "rep", rep, n, "cols", np.unique(cols)[:4].tolist(), "levels", sorted(set(levs.tolist()))[:12])
        print(res, c.ncol, "columns, variant", v, ": bad calls", bad, "of", reps)
