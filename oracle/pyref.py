"""ctypes binding to oracle/_ref/librrtmg_ref.so: the reference's own RRTMG sources, machine-translated F90 -> C.

TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU legs).  The library is built by `make -C oracle _ref` in the
authoring container, where /root/reference exists: tools/f90_to_c.py translates the non-McICA RRTMG_LW / RRTMG_SW
sources statement by statement and gcc compiles the result with the oracle's flags (-O2 -ffp-contract=off).  On the
GPU box there is no /root/reference; the prebuilt .so travels with the repository snapshot.

Same call interface as oracle.pyoracle.Oracle (rrtmg_lw / rrtmg_sw on a Columns batch), so that the tests compare the
hand-written oracle with the reference-derived code on identical inputs, bit for bit.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "librrtmg_ref.so")
_REFERENCE = os.environ.get("MIMA_REFERENCE", "/root/reference")
CP_AIR = 287.04 / (2.0 / 7.0)    # RDGAS/KAPPA, src/shared/constants/constants.f90:64-67

_dp = C.POINTER(C.c_double)


def available() -> bool:
    return os.path.exists(_LIB) or os.path.isdir(os.path.join(_REFERENCE, "src", "atmos_param", "rrtm_radiation"))


def build(force: bool = False) -> str:
    """Translate + compile when the reference sources are present (authoring container); else use the prebuilt file."""
    have_src = os.path.isdir(os.path.join(_REFERENCE, "src", "atmos_param", "rrtm_radiation"))
    if have_src:
        deps = [os.path.join(_HERE, "ref_harness.c"), os.path.join(_HERE, "..", "tools", "f90_to_c.py"),
                os.path.join(_HERE, "..", "tools", "build_tables.py"), os.path.join(_HERE, "Makefile")]
        stale = not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps)
        if force or stale:
            if os.path.exists(_LIB):
                os.remove(_LIB)
            r = subprocess.run(["make", "-C", _HERE, "_ref", "REF=" + _REFERENCE], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("make _ref failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    if not os.path.exists(_LIB):
        raise FileNotFoundError(_LIB + " (build it with `make -C oracle _ref` where /root/reference exists)")
    return _LIB


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Reference:
    def __init__(self, cpdair: float = CP_AIR):
        self.lib = C.CDLL(build())
        self.lib.ref_get_var.restype = C.c_long
        self.lib.ref_get_var.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        self.lib.ref_var_name.restype = C.c_char_p
        self.lib.ref_sw_init.argtypes = [C.c_double]
        self.lib.ref_lw_init.argtypes = [C.c_double]
        if self.lib.ref_sw_init(cpdair) or self.lib.ref_lw_init(cpdair):
            raise RuntimeError("reference init stopped")
        self.max_threads = os.cpu_count() or 1

    def names(self):
        return [self.lib.ref_var_name(i).decode() for i in range(self.lib.ref_var_count())]

    def var(self, name: str) -> np.ndarray:
        """Module variable `module.name` of the translated code (flat, Fortran order), e.g. 'rrsw_kg16.absa'."""
        p, isint = C.c_void_p(), C.c_int()
        n = self.lib.ref_get_var(name.encode(), C.byref(p), C.byref(isint))
        if n < 0:
            raise KeyError(name)
        t = C.c_int if isint.value else C.c_double
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), shape=(n,)).copy()

    def rrtmg_lw(self, cols, *, nthreads: int | None = None, tauaer=None, idrv: int = 0, icld: int = 0, clouds=None,
                 inflglw: int = 0, iceflglw: int = 0, liqflglw: int = 0):
        ncol, nlay = cols.ncol, cols.nlay
        out = {k: np.zeros((ncol, nlay + 1), order="F") for k in ("uflx", "dflx", "uflxc", "dflxc")}
        out.update({k: np.zeros((ncol, nlay), order="F") for k in ("hr", "hrc")})
        if idrv:
            out.update({k: np.zeros((ncol, nlay + 1), order="F") for k in ("duflx_dt", "duflxc_dt")})
        ins = [_f(x) for x in (cols.play, cols.plev, cols.tlay, cols.tlev, cols.tsfc, cols.h2o, cols.o3, cols.co2, cols.ch4,
                               cols.n2o, cols.o2, cols.cfc11, cols.cfc12, cols.cfc22, cols.ccl4, cols.emis)]
        clouds = clouds or {}
        cl = [_f(clouds[k]) if k in clouds else None for k in ("cldfr", "taucld", "cicewp", "cliqwp", "reice", "reliq")]
        ta = None if tauaer is None else _f(tauaer)
        rc = self.lib.ref_rrtmg_lw(C.c_int(ncol), C.c_int(nlay), C.c_int(int(icld)), C.c_int(int(idrv)), *[_p(a) for a in ins],
                                   C.c_int(int(inflglw)), C.c_int(int(iceflglw)), C.c_int(int(liqflglw)), *[_p(a) for a in cl], _p(ta),
                                   _p(out["uflx"]), _p(out["dflx"]), _p(out["hr"]), _p(out["uflxc"]), _p(out["dflxc"]), _p(out["hrc"]),
                                   _p(out.get("duflx_dt")), _p(out.get("duflxc_dt")), C.c_int(nthreads or self.max_threads))
        if rc:
            raise RuntimeError(f"ref_rrtmg_lw rc={rc}")
        return out

    def rrtmg_sw(self, cols, *, nthreads: int | None = None, icld: int = 0, iaer: int = 0, clouds=None, aerosols=None,
                 inflgsw: int = 0, iceflgsw: int = 0, liqflgsw: int = 0):
        ncol, nlay = cols.ncol, cols.nlay
        out = {k: np.zeros((ncol, nlay + 1), order="F") for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
        out.update({k: np.zeros((ncol, nlay), order="F") for k in ("swhr", "swhrc")})
        ins = [_f(x) for x in (cols.play, cols.plev, cols.tlay, cols.tlev, cols.tsfc, cols.h2o, cols.o3, cols.co2, cols.ch4,
                               cols.n2o, cols.o2, cols.albedo, cols.albedo, cols.albedo, cols.albedo, cols.coszen)]
        clouds = clouds or {}
        aerosols = aerosols or {}
        cl = [_f(clouds[k]) if k in clouds else None for k in ("cldfr", "taucld", "ssacld", "asmcld", "fsfcld", "cicewp", "cliqwp", "reice", "reliq")]
        ae = [_f(aerosols[k]) if k in aerosols else None for k in ("tauaer", "ssaaer", "asmaer", "ecaer")]
        rc = self.lib.ref_rrtmg_sw(C.c_int(ncol), C.c_int(nlay), C.c_int(int(icld)), C.c_int(int(iaer)), *[_p(a) for a in ins],
                                   C.c_double(cols.adjes), C.c_int(cols.dyofyr), C.c_double(cols.scon),
                                   C.c_int(int(inflgsw)), C.c_int(int(iceflgsw)), C.c_int(int(liqflgsw)),
                                   *[_p(a) for a in cl], *[_p(a) for a in ae],
                                   _p(out["swuflx"]), _p(out["swdflx"]), _p(out["swhr"]), _p(out["swuflxc"]), _p(out["swdflxc"]),
                                   _p(out["swhrc"]), C.c_int(nthreads or self.max_threads))
        if rc:
            raise RuntimeError(f"ref_rrtmg_sw rc={rc}")
        return out
