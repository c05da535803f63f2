"""ctypes binding to the CPU oracle (oracle/_build/librrtmg_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the mima_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(_HERE, "..", "mima_b200", "data")
_LIB = os.path.join(_HERE, "_build", "librrtmg_oracle.so")
CP_AIR = 287.04 / (2.0 / 7.0)    # RDGAS/KAPPA, src/shared/constants/constants.f90:64-67

NG_LW, NG_SW = 140, 112


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in ("tables.c", "lw.c", "sw.c", "rrtmg_oracle.h")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _LwStages(C.Structure):
    _int_fields = ["laytrop", "jp", "jt", "jt1", "indself", "indfor", "indminor"]
    _dbl_fields = ["fac00", "fac01", "fac10", "fac11", "colh2o", "colco2", "colo3", "coln2o", "colco", "colch4",
                   "colo2", "colbrd", "selffac", "selffrac", "forfac", "forfrac", "minorfrac", "scaleminor",
                   "scaleminorn2", "coldry", "pwvcm", "planklay", "planklev", "plankbnd", "taug", "fracs"]
    _fields_ = [(n, _ip) for n in _int_fields] + [(n, _dp) for n in _dbl_fields]


class _SwStages(C.Structure):
    _int_fields = ["laytrop", "jp", "jt", "jt1", "indself", "indfor"]
    _dbl_fields = ["fac00", "fac01", "fac10", "fac11", "colh2o", "colco2", "colo3", "coln2o", "colch4", "colo2",
                   "colmol", "selffac", "selffrac", "forfac", "forfrac", "taug", "taur", "sfluxzen"]
    _fields_ = [(n, _ip) for n in _int_fields] + [(n, _dp) for n in _dbl_fields]


def _f(a):
    a = np.asfortranarray(a, dtype=np.float64)
    return a


def _p(a):
    return a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self, cpdair: float = CP_AIR, lw_kg: str | None = None):
        self.lib = C.CDLL(build())
        self.lib.orc_init.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_double]
        self.lib.orc_get_table.restype = C.c_long
        self.lib.orc_get_table.argtypes = [C.c_char_p, C.POINTER(_dp)]
        rc = self.lib.orc_init(os.path.join(_DATA, "rrtmg_lw_ref.bin").encode(),
                               (lw_kg or os.path.join(_DATA, "rrtmg_lw_kg_synth.bin")).encode(),
                               os.path.join(_DATA, "rrtmg_sw_kg.bin").encode(), cpdair)
        if rc:
            raise RuntimeError(f"orc_init failed rc={rc}")
        self.max_threads = self.lib.orc_max_threads()

    def table(self, name: str) -> np.ndarray:
        p = _dp()
        n = self.lib.orc_get_table(name.encode(), C.byref(p))
        if n < 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    # ------------------------------------------------------------------ LW
    def rrtmg_lw(self, cols, *, stages: bool = False, nthreads: int | None = None, tauaer=None, idrv: int = 0,
                 icld: int = 0, clouds=None, inflglw: int = 0, iceflglw: int = 0, liqflglw: int = 0):
        """clouds = dict(cldfr (ncol,nlay), taucld (16,ncol,nlay)) for icld >= 1 (inflglw = 0): icld = 1 random overlap
        (rtrn), 2/3 maximum/random overlap (rtrnmr); with inflglw = 1, 2 also cicewp, cliqwp (g/m2) and reice, reliq
        (microns), all (ncol,nlay) (cldprop's parameterisations).  A Fortran `stop` of cldprop comes back as rc = 10 + n."""
        ncol, nlay = cols.ncol, cols.nlay
        nthreads = nthreads or self.max_threads
        out = {k: np.zeros((ncol, nlay + 1), order="F") for k in ("uflx", "dflx", "uflxc", "dflxc")}
        out.update({k: np.zeros((ncol, nlay), order="F") for k in ("hr", "hrc")})
        st = None
        stp = None
        if stages:
            st = {}
            s = _LwStages()
            for n in _LwStages._int_fields:
                st[n] = np.zeros((ncol,) if n == "laytrop" else (ncol, nlay), dtype=np.int32, order="F")
                setattr(s, n, st[n].ctypes.data_as(_ip))
            for n in _LwStages._dbl_fields:
                shape = {"pwvcm": (ncol,), "planklay": (ncol, nlay, 16), "planklev": (ncol, nlay + 1, 16),
                         "plankbnd": (ncol, 16), "taug": (ncol, nlay, NG_LW), "fracs": (ncol, nlay, NG_LW)}.get(n, (ncol, nlay))
                st[n] = np.zeros(shape, order="F")
                setattr(s, n, _p(st[n]))
            stp = C.byref(s)
        if tauaer is None:
            tauaer = np.zeros((ncol, nlay, 16), order="F")
        ins = [_f(x) for x in (cols.play, cols.plev, cols.tlay, cols.tlev, cols.tsfc, cols.h2o, cols.o3, cols.co2,
                               cols.ch4, cols.n2o, cols.o2, cols.cfc11, cols.cfc12, cols.cfc22, cols.ccl4,
                               cols.emis, tauaer)]
        if idrv:
            out.update({k: np.zeros((ncol, nlay + 1), order="F") for k in ("duflx_dt", "duflxc_dt")})
        clouds = clouds or {}
        if clouds and "taucld" not in clouds:
            clouds = dict(clouds, taucld=np.zeros((16, ncol, nlay), order="F"))
        cl = [_f(clouds[k]) if k in clouds else None for k in ("cldfr", "taucld")]
        wp = [_f(clouds[k]) if k in clouds else None for k in ("cicewp", "cliqwp", "reice", "reliq")]
        rc = self.lib.orc_rrtmg_lw(C.c_int(ncol), C.c_int(nlay), C.c_int(int(icld)), C.c_int(int(idrv)), *[_p(a) for a in ins],
                                   C.c_int(int(inflglw)), *[None if a is None else _p(a) for a in cl],
                                   C.c_int(int(iceflglw)), C.c_int(int(liqflglw)), *[None if a is None else _p(a) for a in wp],
                                   _p(out["uflx"]), _p(out["dflx"]), _p(out["hr"]), _p(out["uflxc"]),
                                   _p(out["dflxc"]), _p(out["hrc"]),
                                   _p(out["duflx_dt"]) if idrv else None, _p(out["duflxc_dt"]) if idrv else None,
                                   stp, C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"orc_rrtmg_lw rc={rc}")
        if stages:
            out["stages"] = st
        return out

    # ------------------------------------------------------------------ SW
    def rrtmg_sw(self, cols, *, stages: bool = False, nthreads: int | None = None, icld: int = 0, iaer: int = 0,
                 clouds=None, aerosols=None, inflgsw: int = 0, iceflgsw: int = 0, liqflgsw: int = 0):
        """clouds = dict(cldfr (ncol,nlay), taucld/ssacld/asmcld/fsfcld (14,ncol,nlay)) for icld >= 1 (inflgsw = 0);
        aerosols = dict(tauaer/ssaaer/asmaer (ncol,nlay,14)) for iaer = 10, dict(ecaer (ncol,nlay,6)) for iaer = 6."""
        ncol, nlay = cols.ncol, cols.nlay
        nthreads = nthreads or self.max_threads
        out = {k: np.zeros((ncol, nlay + 1), order="F") for k in ("swuflx", "swdflx", "swuflxc", "swdflxc")}
        out.update({k: np.zeros((ncol, nlay), order="F") for k in ("swhr", "swhrc")})
        st = None
        stp = None
        if stages:
            st = {}
            s = _SwStages()
            for n in _SwStages._int_fields:
                st[n] = np.zeros((ncol,) if n == "laytrop" else (ncol, nlay), dtype=np.int32, order="F")
                setattr(s, n, st[n].ctypes.data_as(_ip))
            for n in _SwStages._dbl_fields:
                shape = {"taug": (ncol, nlay, NG_SW), "taur": (ncol, nlay, NG_SW), "sfluxzen": (ncol, NG_SW)}.get(n, (ncol, nlay))
                st[n] = np.zeros(shape, order="F")
                setattr(s, n, _p(st[n]))
            stp = C.byref(s)
        ins = [_f(x) for x in (cols.play, cols.plev, cols.tlay, cols.tlev, cols.tsfc, cols.h2o, cols.o3, cols.co2,
                               cols.ch4, cols.n2o, cols.o2, cols.albedo, cols.albedo, cols.albedo, cols.albedo,
                               cols.coszen)]
        extra = []
        if clouds is not None and "taucld" not in clouds:      # water-path input: the optical-property dummies still exist
            z = np.zeros((14, ncol, nlay), order="F")
            clouds = dict(clouds, taucld=z, ssacld=z, asmcld=z, fsfcld=z)
        for src, keys in ((clouds, ("cldfr", "taucld", "ssacld", "asmcld", "fsfcld")), (aerosols, ("tauaer", "ssaaer", "asmaer", "ecaer"))):
            for k in keys:
                extra.append(_f(src[k]) if src is not None and k in src else None)
        wp = [_f(clouds[k]) if clouds is not None and k in clouds else None for k in ("cicewp", "cliqwp", "reice", "reliq")]
        rc = self.lib.orc_rrtmg_sw(C.c_int(ncol), C.c_int(nlay), C.c_int(int(icld)), C.c_int(int(iaer)), *[_p(a) for a in ins],
                                   C.c_double(cols.adjes), C.c_int(cols.dyofyr), C.c_double(cols.scon),
                                   C.c_int(int(inflgsw)), *[None if a is None else _p(a) for a in extra],
                                   C.c_int(int(iceflgsw)), C.c_int(int(liqflgsw)), *[None if a is None else _p(a) for a in wp],
                                   _p(out["swuflx"]), _p(out["swdflx"]), _p(out["swhr"]), _p(out["swuflxc"]),
                                   _p(out["swdflxc"]), _p(out["swhrc"]), stp, C.c_int(nthreads))
        if rc:
            raise RuntimeError(f"orc_rrtmg_sw rc={rc}")
        if stages:
            out["stages"] = st
        return out
