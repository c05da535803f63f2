#!/usr/bin/env python3
"""Ad-hoc GPU vs oracle comparison with per-stage diagnostics (developer tool; the judged parity tests
live in tests/).  Usage: python tools/gpu_check.py [RES] [stride]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mima_b200 import rrtmg  # noqa: E402
from mima_b200.columns import make_columns  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def rel(a, b):
    d = np.abs(a - b)
    s = np.maximum(np.abs(b), 1e-300)
    return d.max(), (d / s).max()


def main():
    res = sys.argv[1] if len(sys.argv) > 1 else "T42L40"
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    kw = {}
    if len(sys.argv) > 3 and sys.argv[3] == "c4":
        kw = dict(co2_ppmv=1560.0, ozone="file", secondary_gases=True)
    cols = make_columns(res, **kw)
    cols = cols.take(np.arange(0, cols.ncol, stride))
    print(f"{res}: {cols.ncol} columns x {cols.nlay} layers  kw={kw}")
    orc = Oracle()
    t = time.time()
    olw = orc.rrtmg_lw(cols, stages=True)
    osw = orc.rrtmg_sw(cols, stages=True)
    print(f"oracle: {time.time() - t:.2f}s on {orc.max_threads} threads")

    rrtmg.set_device(0)
    rrtmg.rrtmg_lw_ini(allow_synthetic_lw=True)
    rrtmg.rrtmg_sw_ini()
    rrtmg.set_option("capture_stages", 1)
    rrtmg.set_option("chunk", 1 << 20)

    # ---- reduced tables
    for name in ("lw01.absa", "lw03.absb", "lw03.ka_mn2o", "lw05.fracrefa", "lw08.cfc22adj", "lw16.absa",
                 "sw16.absa", "sw17.absb", "sw24.rayla", "sw24.sfluxref", "sw29.absco2", "sw23.rayl"):
        a, b = rrtmg.get_table(name), orc.table(name)
        print(f"table {name:14s} n={a.size:6d} bit-equal={np.array_equal(a, b)}")

    nc, nl = cols.ncol, cols.nlay
    # ---- LW
    t = time.time()
    glw = rrtmg.lw_from_columns(cols)
    print(f"gpu lw host call: {time.time() - t:.3f}s")
    st = olw["stages"]
    for f in ("laytrop",):
        g = rrtmg.get_stage("lw." + f, (nc,))
        print(f"lw.{f:12s} mismatches: {(g != st[f]).sum()}")
    for f in ("jp", "jt", "jt1", "indself", "indfor", "indminor"):
        g = rrtmg.get_stage("lw." + f, (nc, nl))
        o = st[f]
        if f == "indself":
            lay = np.arange(1, nl + 1)[None, :]
            m = lay <= st["laytrop"][:, None]
            print(f"lw.{f:12s} mismatches (below laytrop): {(g[m] != o[m]).sum()}")
        else:
            print(f"lw.{f:12s} mismatches: {(g != o).sum()}")
    for f in ("fac00", "fac01", "fac10", "fac11", "colh2o", "colco2", "colo3", "coln2o", "colco", "colch4", "colo2",
              "colbrd", "selffac", "forfac", "forfrac", "minorfrac", "scaleminor", "scaleminorn2", "coldry"):
        g = rrtmg.get_stage("lw." + f, (nc, nl))
        print(f"lw.{f:12s} maxabs={rel(g, st[f])[0]:.3e} maxrel={rel(g, st[f])[1]:.3e} biteq={np.array_equal(g, st[f])}")
    for f, shp in (("planklay", (nc, nl, 16)), ("planklev", (nc, nl + 1, 16)), ("plankbnd", (nc, 16))):
        g = rrtmg.get_stage("lw." + f, shp)
        print(f"lw.{f:12s} maxrel={rel(g, st[f])[1]:.3e} biteq={np.array_equal(g, st[f])}")
    for f in ("taug", "fracs"):
        g = rrtmg.get_stage("lw." + f, (nc, nl, 140))
        o = st[f]
        d = np.abs(g - o) / np.maximum(np.abs(o), 1e-300)
        d[o == 0] = np.abs(g[o == 0])
        bad = np.unravel_index(np.argmax(d), d.shape)
        print(f"lw.{f:12s} maxrel={d.max():.3e} at col,lay,g={bad} gpu={g[bad]:.6e} orc={o[bad]:.6e}")
        perband = [d[:, :, a:b].max() for a, b in zip([0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138],
                                                      [10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140])]
        print("     per band:", " ".join(f"{x:.1e}" for x in perband))
    for k, g in zip(("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc"), glw):
        a, r = rel(g, olw[k])
        print(f"LW {k:6s} maxabs={a:.3e} maxrel={r:.3e}")

    # ---- SW
    t = time.time()
    gsw = rrtmg.sw_from_columns(cols)
    print(f"gpu sw host call: {time.time() - t:.3f}s")
    st = osw["stages"]
    g = rrtmg.get_stage("sw.laytrop", (nc,))
    print(f"sw.laytrop      mismatches: {(g != st['laytrop']).sum()}")
    for f in ("jp", "jt", "jt1", "indfor"):
        g = rrtmg.get_stage("sw." + f, (nc, nl))
        print(f"sw.{f:12s} mismatches: {(g != st[f]).sum()}")
    for f in ("fac00", "fac11", "colh2o", "colco2", "colo3", "colch4", "colo2", "colmol", "selffac", "selffrac", "forfac", "forfrac"):
        g = rrtmg.get_stage("sw." + f, (nc, nl))
        print(f"sw.{f:12s} maxrel={rel(g, st[f])[1]:.3e} biteq={np.array_equal(g, st[f])}")
    for f in ("taug", "taur"):
        g = rrtmg.get_stage("sw." + f, (nc, nl, 112))
        o = st[f]
        d = np.abs(g - o) / np.maximum(np.abs(o), 1e-300)
        d[o == 0] = np.abs(g[o == 0])
        bad = np.unravel_index(np.argmax(d), d.shape)
        print(f"sw.{f:12s} maxrel={d.max():.3e} at {bad} gpu={g[bad]:.6e} orc={o[bad]:.6e}")
    g = rrtmg.get_stage("sw.sfluxzen", (nc, 112))
    print(f"sw.sfluxzen     maxrel={rel(g, st['sfluxzen'])[1]:.3e}")
    for k, g in zip(("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc"), gsw):
        a, r = rel(g, osw[k])
        print(f"SW {k:7s} maxabs={a:.3e} maxrel={r:.3e}")
    print("launches:", rrtmg.launch_count())


if __name__ == "__main__":
    main()
