"""Host-side mirror of the reference's RRTMG interface over the C ABI of librrtmg_b200.so.

Same procedure names, argument order and meaning as the Fortran module procedures MiMA calls:

    rrtmg_lw_ini(cpdair)   LW/src/rrtmg_lw_init.f90:28            (physics_driver.f90:577)
    rrtmg_sw_ini(cpdair)   SW/src/rrtmg_sw_init.f90:28            (physics_driver.f90:578)
    rrtmg_lw(...)          LW/src/rrtmg_lw_rad.nomcica.f90:80-89  (rrtm_radiation.f90:722-748)
    rrtmg_sw(...)          SW/src/rrtmg_sw_rad.nomcica.f90:78-88  (rrtm_radiation.f90:686-712)

The Fortran intent(out) dummies are returned instead of passed.  Arrays are numpy float64; anything not
already Fortran-ordered is converted.  There is no CPU path: if the shared library or a CUDA device is
missing the calls raise.  Error codes of the C ABI become exceptions (RRTMGError); the reference's
`stop 'PARTIAL CLOUD NOT ALLOWED'` and its unbuilt branches surface the same way.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

CP_AIR = 287.04 / (2.0 / 7.0)      # RDGAS/KAPPA (src/shared/constants/constants.f90:64-67)
NBNDLW, NGPTLW, NBNDSW, NGPTSW = 16, 140, 14, 112
DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

_dp = C.POINTER(C.c_double)
_ERR = {1: "not initialised", 2: "unsupported option", 3: "PARTIAL CLOUD NOT ALLOWED", 4: "bad argument",
        5: "CUDA error", 6: "coefficient tables", 7: "cloud input out of range (a Fortran stop of cldprop)"}


class RRTMGError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rrtmg_b200 error {code} ({_ERR.get(code, '?')}): {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load librrtmg_b200.so (built in-tree by mima_b200.build).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            raise RRTMGError(5, f"{path} is missing: run `python -m mima_b200.build` (nvcc) first; there is no fallback")
        L = C.CDLL(path)
        L.rrtmg_b200_last_error.restype = C.c_char_p
        L.rrtmg_b200_launch_count.restype = C.c_long
        L.rrtmg_b200_get_table.restype = C.c_long
        L.rrtmg_b200_get_table.argtypes = [C.c_char_p, _dp, C.c_long]
        L.rrtmg_b200_get_stage.restype = C.c_long
        L.rrtmg_b200_get_stage.argtypes = [C.c_char_p, _dp, C.c_long]
        L.rrtmg_b200_set_table.argtypes = [C.c_char_p, _dp, C.c_int, C.POINTER(C.c_int)]
        L.rrtmg_b200_load_tables.argtypes = [C.c_char_p]
        L.rrtmg_b200_lw_init.argtypes = [C.c_double]
        L.rrtmg_b200_sw_init.argtypes = [C.c_double]
        L.rrtmg_b200_set_option.argtypes = [C.c_char_p, C.c_long]
        _lib = L
        # developer knob: RRTMG_TUNE="key=value[,key=value]" applies library options at load time
        for kv in filter(None, os.environ.get("RRTMG_TUNE", "").split(",")):
            k, v = kv.split("=")
            rc = L.rrtmg_b200_set_option(k.encode(), int(v))
            if rc:
                raise RRTMGError(rc, L.rrtmg_b200_last_error().decode())
    return _lib


def _check(rc: int):
    if rc:
        raise RRTMGError(rc, lib().rrtmg_b200_last_error().decode())


def set_device(local_rank: int = 0):
    _check(lib().rrtmg_b200_set_device(int(local_rank)))


def load_tables(path: str):
    _check(lib().rrtmg_b200_load_tables(path.encode()))


def set_table(name: str, array):
    a = np.asfortranarray(array, dtype=np.float64)
    dims = (C.c_int * a.ndim)(*a.shape)
    _check(lib().rrtmg_b200_set_table(name.encode(), a.ctypes.data_as(_dp), a.ndim, dims))


def set_option(key: str, value: int):
    _check(lib().rrtmg_b200_set_option(key.encode(), int(value)))


def launch_count() -> int:
    return int(lib().rrtmg_b200_launch_count())


_loaded = set()


def _default_tables(kind: str):
    # the real LW coefficients (tools/build_tables.py with MIMA_LW_KG=.../rrtmg_lw_k_g.f90) win over the synthetic blob
    lw_kg = "rrtmg_lw_kg.bin" if os.path.exists(os.path.join(DATA_DIR, "rrtmg_lw_kg.bin")) else "rrtmg_lw_kg_synth.bin"
    files = {"lw": ["rrtmg_lw_ref.bin", lw_kg], "sw": ["rrtmg_sw_kg.bin"]}[kind]
    for f in files:
        if f not in _loaded:
            load_tables(os.path.join(DATA_DIR, f))
            _loaded.add(f)


class SyntheticTablesWarning(UserWarning):
    pass


def tables_info() -> dict:
    a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
    _check(lib().rrtmg_b200_tables_info(C.byref(a), C.byref(b), C.byref(c)))
    return dict(lw_synthetic=bool(a.value), lw_ready=bool(b.value), sw_ready=bool(c.value))


def rrtmg_lw_ini(cpdair: float = CP_AIR, *, default_tables: bool = True, allow_synthetic_lw: bool | None = None):
    """rrtmg_lw_ini(cpdair) (LW/src/rrtmg_lw_init.f90:28).  With default_tables the packaged blobs are registered first.
    The packaged LW k-distribution is SYNTHETIC (the reference checkout carries no rrtmg_lw_k_g.f90): fluxes computed with
    it exercise the algorithm and mean nothing physically.  It is only used when the caller says so --
    allow_synthetic_lw=True (tests, bench.py, smoke) or MIMA_B200_ALLOW_SYNTHETIC_LW=1 -- otherwise this raises; a real
    coefficient set (mima_b200/data/rrtmg_lw_kg.bin from tools/build_tables.py, or arrays registered with set_table() /
    load_tables() and default_tables=False) needs no switch.  tables_info() reports which set is active."""
    if allow_synthetic_lw is None:
        allow_synthetic_lw = os.environ.get("MIMA_B200_ALLOW_SYNTHETIC_LW", "") not in ("", "0")
    if default_tables:
        _default_tables("lw")
    if tables_info()["lw_synthetic"]:
        if not allow_synthetic_lw:
            raise RRTMGError(6, "the registered LW k-distribution is the packaged synthetic stand-in; pass "
                                "allow_synthetic_lw=True (or MIMA_B200_ALLOW_SYNTHETIC_LW=1) to run on it, or register the "
                                "real rrtmg_lw_k_g coefficients (tools/build_tables.py, set_table/load_tables)")
        import warnings
        warnings.warn("RRTMG_LW runs on SYNTHETIC k-distribution tables: LW fluxes and heating rates are not physical",
                      SyntheticTablesWarning, stacklevel=2)
    _check(lib().rrtmg_b200_lw_init(float(cpdair)))


def rrtmg_sw_ini(cpdair: float = CP_AIR, *, default_tables: bool = True):
    if default_tables:
        _default_tables("sw")
    _check(lib().rrtmg_b200_sw_init(float(cpdair)))


def finalize():
    _check(lib().rrtmg_b200_finalize())
    _loaded.clear()


def get_table(name: str) -> np.ndarray:
    n = lib().rrtmg_b200_get_table(name.encode(), None, 0)
    if n < 0:
        raise KeyError(name)
    out = np.empty(n)
    lib().rrtmg_b200_get_table(name.encode(), out.ctypes.data_as(_dp), n)
    return out


def get_stage(name: str, shape) -> np.ndarray:
    n = lib().rrtmg_b200_get_stage(name.encode(), None, 0)
    if n < 0:
        raise KeyError(f"stage {name} unavailable (single-pass calls only; lw.taug/fracs need capture_stages)")
    out = np.empty(n)
    lib().rrtmg_b200_get_stage(name.encode(), out.ctypes.data_as(_dp), n)
    return out.reshape(shape, order="F")


def _in(a, shape, name, optional=False):
    if a is None:
        if optional:
            return None, None
        raise RRTMGError(4, f"{name} is required")
    arr = np.asfortranarray(a, dtype=np.float64)
    if arr.shape != tuple(shape):
        raise RRTMGError(4, f"{name}: expected shape {tuple(shape)}, got {arr.shape}")
    return arr, arr.ctypes.data_as(_dp)


def rrtmg_lw(ncol, nlay, icld, idrv,
             play, plev, tlay, tlev, tsfc,
             h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
             cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, emis,
             inflglw=0, iceflglw=0, liqflglw=0, cldfr=None,
             taucld=None, cicewp=None, cliqwp=None, reice=None, reliq=None,
             tauaer=None, clear_sky=True):
    """Returns (uflx, dflx, hr, uflxc, dflxc, hrc); fluxes (ncol, nlay+1) W/m2, heating (ncol, nlay) K/day; with
    idrv = 1 also (duflx_dt, duflxc_dt), the change of the upward flux per K of surface temperature (W/m2/K,
    rad.nomcica:143-152, the Fortran's optional dummies).  ch4vmr..ccl4vmr, emis, tauaer may be None (zeros /
    emissivity 1).  icld >= 1 takes the cloud fraction cldfr (ncol,nlay) and the band optical depths taucld
    (16,ncol,nlay) (inflglw = 0): icld = 1 random overlap, 2/3 maximum/random overlap; inflglw = 1, 2 take the ice / liquid
    water paths cicewp, cliqwp (g/m2) and, for 2, the effective radii reice, reliq (microns) with iceflglw 0..3 and
    liqflglw 0..1 (cldprop's parameterisations; a radius outside the parameterisation's range raises RRTMGError(7)).  clear_sky=False does not fetch uflxc, dflxc, hrc (duflxc_dt): they come back as
    None (MiMA never reads them; the C ABI takes NULL for them)."""
    L = (ncol, nlay)
    V = (ncol, nlay + 1)
    keep = []
    ptrs = []
    for a, shp, nm, opt in ((play, L, "play", False), (plev, V, "plev", False), (tlay, L, "tlay", False),
                            (tlev, V, "tlev", False), (tsfc, (ncol,), "tsfc", False),
                            (h2ovmr, L, "h2ovmr", False), (o3vmr, L, "o3vmr", False), (co2vmr, L, "co2vmr", False),
                            (ch4vmr, L, "ch4vmr", True), (n2ovmr, L, "n2ovmr", True), (o2vmr, L, "o2vmr", True),
                            (cfc11vmr, L, "cfc11vmr", True), (cfc12vmr, L, "cfc12vmr", True),
                            (cfc22vmr, L, "cfc22vmr", True), (ccl4vmr, L, "ccl4vmr", True),
                            (emis, (ncol, NBNDLW), "emis", True)):
        arr, p = _in(a, shp, nm, opt)
        keep.append(arr)
        ptrs.append(p)
    taer, ptaer = _in(tauaer, (ncol, nlay, NBNDLW), "tauaer", True)
    if int(icld) != 0 and taucld is None and cldfr is not None:
        taucld = np.zeros((NBNDLW, ncol, nlay), order="F")        # the Fortran dummy always exists
    cloudp = []
    for a, shp, nm in ((cldfr, L, "cldfr"), (taucld, (NBNDLW, ncol, nlay), "taucld"), (cicewp, L, "cicewp"),
                       (cliqwp, L, "cliqwp"), (reice, L, "reice"), (reliq, L, "reliq")):
        arr, p = _in(a, shp, nm, True)
        keep.append(arr)
        cloudp.append(p)
    out = [np.empty(V, order="F"), np.empty(V, order="F"), np.empty(L, order="F"),
           np.empty(V, order="F"), np.empty(V, order="F"), np.empty(L, order="F")]
    if int(idrv) == 1:
        out += [np.empty(V, order="F"), np.empty(V, order="F")]
    if not clear_sky:
        for i in (3, 4, 5, 7):
            if i < len(out):
                out[i] = None
    icld_c = C.c_int(int(icld))
    rc = lib().rrtmg_b200_lw(C.c_int(ncol), C.c_int(nlay), C.byref(icld_c), C.c_int(int(idrv)), *ptrs,
                             C.c_int(inflglw), C.c_int(iceflglw), C.c_int(liqflglw), *cloudp,
                             ptaer, *[None if o is None else o.ctypes.data_as(_dp) for o in out],
                             *([] if int(idrv) == 1 else [None, None]))
    _check(rc)
    return tuple(out)


def rrtmg_sw(ncol, nlay, icld, iaer,
             play, plev, tlay, tlev, tsfc,
             h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
             asdir, asdif, aldir, aldif,
             coszen, adjes, dyofyr, scon,
             inflgsw=0, iceflgsw=0, liqflgsw=0, cldfr=None,
             taucld=None, ssacld=None, asmcld=None, fsfcld=None,
             cicewp=None, cliqwp=None, reice=None, reliq=None,
             tauaer=None, ssaaer=None, asmaer=None, ecaer=None, clear_sky=True):
    """Returns (swuflx, swdflx, swhr, swuflxc, swdflxc, swhrc); the last three are None with clear_sky=False.  icld >= 1 takes cloud optical properties
    (inflgsw = 0: cldfr (ncol,nlay) 0 or 1, taucld/ssacld/asmcld/fsfcld (14,ncol,nlay)); iaer = 10 takes
    tauaer/ssaaer/asmaer (ncol,nlay,14); iaer = 6 takes ecaer (ncol,nlay,6), the optical depth at 0.55 micron of the six
    ECMWF aerosol types.  inflgsw = 2 takes the water paths cicewp, cliqwp (g/m2) and effective radii reice, reliq (microns)
    with iceflgsw 1..3 and liqflgsw = 1 (cldprop_sw's parameterisations; a radius outside its range raises RRTMGError(7));
    a partially cloudy layer raises RRTMGError(3) like the reference's stop."""
    L = (ncol, nlay)
    V = (ncol, nlay + 1)
    B = (NBNDSW, ncol, nlay)
    A = (ncol, nlay, NBNDSW)
    keep = []
    ptrs = []
    optional = []
    for a, shp, nm in ((cldfr, L, "cldfr"), (taucld, B, "taucld"), (ssacld, B, "ssacld"), (asmcld, B, "asmcld"),
                       (fsfcld, B, "fsfcld"), (cicewp, L, "cicewp"), (cliqwp, L, "cliqwp"), (reice, L, "reice"),
                       (reliq, L, "reliq"), (tauaer, A, "tauaer"), (ssaaer, A, "ssaaer"), (asmaer, A, "asmaer"),
                       (ecaer, (ncol, nlay, 6), "ecaer")):
        arr, p = _in(a, shp, nm, True)
        keep.append(arr)
        optional.append(p)
    for a, shp, nm, opt in ((play, L, "play", False), (plev, V, "plev", False), (tlay, L, "tlay", False),
                            (tlev, V, "tlev", False), (tsfc, (ncol,), "tsfc", False),
                            (h2ovmr, L, "h2ovmr", False), (o3vmr, L, "o3vmr", False), (co2vmr, L, "co2vmr", False),
                            (ch4vmr, L, "ch4vmr", True), (n2ovmr, L, "n2ovmr", True), (o2vmr, L, "o2vmr", True),
                            (asdir, (ncol,), "asdir", False), (asdif, (ncol,), "asdif", False),
                            (aldir, (ncol,), "aldir", False), (aldif, (ncol,), "aldif", False),
                            (coszen, (ncol,), "coszen", False)):
        arr, p = _in(a, shp, nm, opt)
        keep.append(arr)
        ptrs.append(p)
    out = [np.empty(V, order="F"), np.empty(V, order="F"), np.empty(L, order="F"),
           np.empty(V, order="F"), np.empty(V, order="F"), np.empty(L, order="F")]
    if not clear_sky:
        out[3:] = [None, None, None]
    icld_c, iaer_c = C.c_int(int(icld)), C.c_int(int(iaer))
    rc = lib().rrtmg_b200_sw(C.c_int(ncol), C.c_int(nlay), C.byref(icld_c), C.byref(iaer_c), *ptrs,
                             C.c_double(float(adjes)), C.c_int(int(dyofyr)), C.c_double(float(scon)),
                             C.c_int(inflgsw), C.c_int(iceflgsw), C.c_int(liqflgsw),
                             *optional, *[None if o is None else o.ctypes.data_as(_dp) for o in out])
    _check(rc)
    return tuple(out)


# ---- convenience over a Columns batch (mima_b200.columns) ---------------------------------------------
def _opt(a):
    """MiMA passes literal zeros for the secondary gases by default; forward None so the ABI skips them."""
    return None if (a is None or not np.any(a)) else a


def lw_from_columns(c, tauaer=None, idrv=0, icld=0, clouds=None, inflglw=0, clear_sky=True, iceflglw=0, liqflglw=0):
    return rrtmg_lw(c.ncol, c.nlay, icld, idrv, c.play, c.plev, c.tlay, c.tlev, c.tsfc, c.h2o, c.o3, c.co2,
                    _opt(c.ch4), _opt(c.n2o), _opt(c.o2), _opt(c.cfc11), _opt(c.cfc12), _opt(c.cfc22), _opt(c.ccl4),
                    None if np.all(c.emis == 1.0) else c.emis, tauaer=tauaer, inflglw=inflglw, iceflglw=iceflglw, liqflglw=liqflglw, clear_sky=clear_sky, **(clouds or {}))


def sw_from_columns(c, icld=0, iaer=0, clouds=None, aerosols=None, inflgsw=0, clear_sky=True, iceflgsw=0, liqflgsw=0):
    return rrtmg_sw(c.ncol, c.nlay, icld, iaer, c.play, c.plev, c.tlay, c.tlev, c.tsfc, c.h2o, c.o3, c.co2,
                    _opt(c.ch4), _opt(c.n2o), _opt(c.o2), c.albedo, c.albedo, c.albedo, c.albedo,
                    c.coszen, c.adjes, c.dyofyr, c.scon, inflgsw=inflgsw, iceflgsw=iceflgsw, liqflgsw=liqflgsw, clear_sky=clear_sky,
                    **(clouds or {}), **(aerosols or {}))
