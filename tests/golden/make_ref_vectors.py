#!/usr/bin/env python3
"""Writes tests/golden/ref_t42l40.npz: outputs of the REFERENCE's own RRTMG code for a fixed batch.

The vectors come from oracle/_ref/librrtmg_ref.so, i.e. the reference's non-McICA RRTMG_SW / RRTMG_LW sources read from
/root/reference, translated F90 -> C by tools/f90_to_c.py and compiled with gcc (`make -C oracle _ref`; no Fortran compiler
exists in the authoring image).  SW uses the reference's own coefficients; LW uses the synthetic k-tables (the reference's
LW data file is stripped from the checkout), so the LW vectors pin the algorithm, not the physics.
The batch: 96 columns x 40 layers of the T42L40 generator with a realistic night fraction, 4 x CO2, file ozone and
non-zero secondary gases (config C4 of BASELINE.json, which reaches the minor-gas branches of taumol), idrv = 1 for LW.
Only tests read the file (oracle and GPU are compared against it on the GPU box, where /root/reference does not exist)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mima_b200.columns import make_columns  # noqa: E402


def batch():
    return make_columns("T42L40", nlon=12, nlat=8, night=True, co2_ppmv=1560.0, ozone="file", secondary_gases=True)


if __name__ == "__main__":
    from oracle.pyref import Reference
    c = batch()
    r = Reference()
    sw, lw = r.rrtmg_sw(c, nthreads=1), r.rrtmg_lw(c, nthreads=1, idrv=1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_t42l40.npz")
    np.savez_compressed(out, **sw, **lw)
    print("wrote", out, os.path.getsize(out))
