#!/usr/bin/env python3
"""Writes tests/golden/oracle_t42l40.npz: oracle outputs for a fixed 64-column batch.

These are REGRESSION vectors produced by this repo's own CPU oracle (oracle/), not outputs of the
reference: mjucker/MiMA is Fortran and there is no Fortran compiler in the authoring container, its LW
k-distribution data file is stripped from the checkout, and it ships no SW golden files.  They guard the
oracle (and through it the CUDA path) against accidental drift."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mima_b200.columns import make_columns  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

c = make_columns("T42L40", nlon=8, nlat=8, night=True)
o = Oracle()
lw, sw = o.rrtmg_lw(c, nthreads=1), o.rrtmg_sw(c, nthreads=1)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_t42l40.npz")
np.savez_compressed(out, **{"lw_" + k: lw[k] for k in ("uflx", "dflx", "hr")},
                    **{"sw_" + k: sw[k] for k in ("swuflx", "swdflx", "swhr")})
print("wrote", out, os.path.getsize(out))
