"""Host-side mirror of MiMA's radiation driver module over the C ABI (device-side marshaling).

Same procedure names and argument meaning as src/atmos_param/rrtm_radiation/rrtm_radiation.f90 and astro.f90:

    interp_temp(z_full, z_half, t_surf_rad, t)                        rrtm_radiation.f90:422-461
    compute_zenith(Time, equinox_day, dt, lat, lon) -> cosz, dyofyr   astro.f90:59-248
    run_rrtmg(is, js, Time, lat, lon, p_full, p_half, albedo, q, t, t_surf_rad, tdt, coszen, flux_sw, flux_lw)
                                                                      rrtm_radiation.f90:471-808

`Time` is the (seconds, days) pair of get_time(); the namelists rrtm_radiation_nml / astro_nml are the
RadConfig dataclass (Fortran defaults).  The alarm (dt_rad), the netCDF interpolators and the diag manager
are FMS control plane and stay with the model: run_rrtmg here is the radiation step proper, and fields the
interpolators would deliver (o3f, a replaced q) are arguments.  Arrays are FMS-ordered (lon, lat, lev),
level 1 = top, Pa / K / kg kg-1.  No CPU path: everything runs in librrtmg_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import rrtmg as _r

_dp = C.POINTER(C.c_double)


class _CConfig(C.Structure):
    """struct rrtmg_b200_rad_config (include/rrtmg_b200.h)."""
    _fields_ = [(n, C.c_int) for n in ("include_secondary_gases", "do_fixed_water", "do_zm_tracers", "do_rad_time_avg",
                                       "dt_rad_avg", "lonstep", "do_zm_rad", "use_dyofyr", "solday", "days_per_year")] + \
               [(n, C.c_double) for n in ("scale_ozone", "o3_val", "ch4_val", "n2o_val", "o2_val", "cfc11_val", "cfc12_val",
                                          "cfc22_val", "ccl4_val", "h2o_lower_limit", "temp_lower_limit", "temp_upper_limit",
                                          "co2ppmv", "fixed_water", "fixed_water_pres", "fixed_water_lat", "slowdown_rad",
                                          "obliq", "solr_cnst", "solrad", "equinox_day")]


@dataclasses.dataclass
class RadConfig:
    """rrtm_radiation_nml (rrtm_radiation.f90:104-197) + astro_nml (astro.f90:24-33), Fortran defaults."""
    include_secondary_gases: bool = False
    scale_ozone: float = 1.0
    o3_val: float = 0.0
    ch4_val: float = 0.0
    n2o_val: float = 0.0
    o2_val: float = 0.0
    cfc11_val: float = 0.0
    cfc12_val: float = 0.0
    cfc22_val: float = 0.0
    ccl4_val: float = 0.0
    h2o_lower_limit: float = 2.0e-7
    temp_lower_limit: float = 100.0
    temp_upper_limit: float = 370.0
    co2ppmv: float = 300.0
    do_fixed_water: bool = False
    fixed_water: float = 2.0e-6
    fixed_water_pres: float = 100.0e2
    fixed_water_lat: float = 90.0
    do_zm_tracers: bool = False
    do_rad_time_avg: bool = True
    dt_rad_avg: int = 86400
    lonstep: int = 1
    slowdown_rad: float = 1.0
    do_zm_rad: bool = False
    obliq: float = 23.439
    use_dyofyr: bool = False
    solr_cnst: float = 1368.22
    solrad: float = 1.0
    solday: int = 0
    equinox_day: float = 0.25
    days_per_year: int = 360

    def to_c(self) -> _CConfig:
        c = _CConfig()
        for name, _ in _CConfig._fields_:
            setattr(c, name, getattr(self, name))
        return c


def default_config() -> RadConfig:
    """The defaults as the library sets them (rrtmg_b200_rad_config_default)."""
    c = _CConfig()
    _r.lib().rrtmg_b200_rad_config_default(C.byref(c))
    return RadConfig(**{n: (bool(getattr(c, n)) if isinstance(getattr(RadConfig(), n), bool) else getattr(c, n))
                        for n, _ in _CConfig._fields_})


def _fa(a, shape, name):
    a = np.asfortranarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def compute_zenith(Time, equinox_day, dt, lat, lon, cfg: RadConfig | None = None):
    """astro.f90 compute_zenith: returns (cosz, dyofyr).  Time = (seconds, days)."""
    cfg = dataclasses.replace(cfg or RadConfig(), equinox_day=float(equinox_day))
    lat = np.asfortranarray(lat, dtype=np.float64)
    lon = _fa(lon, lat.shape, "lon")
    cosz = np.zeros(lat.shape, order="F")
    dy = C.c_int(0)
    c = cfg.to_c()
    _r._check(_r.lib().rrtmg_b200_compute_zenith(C.byref(c), C.c_int(int(Time[0])), C.c_int(int(Time[1])), C.c_int(int(dt)),
                                                 C.c_int(lat.size), _p(lat), _p(lon), _p(cosz), C.byref(dy)))
    return cosz, dy.value


def interp_temp(z_full, z_half, t_surf_rad, t):
    """rrtm_radiation.f90 interp_temp: returns t_half (si, sj, sk+1)."""
    t = np.asfortranarray(t, dtype=np.float64)
    si, sj, sk = t.shape
    z_full = _fa(z_full, (si, sj, sk), "z_full")
    z_half = _fa(z_half, (si, sj, sk + 1), "z_half")
    ts = _fa(t_surf_rad, (si, sj), "t_surf_rad")
    th = np.zeros((si, sj, sk + 1), order="F")
    _r._check(_r.lib().rrtmg_b200_interp_temp(C.c_int(si), C.c_int(sj), C.c_int(sk), _p(z_full), _p(z_half), _p(ts), _p(t), _p(th)))
    return th


def run_rrtmg(is_, js, Time, lat, lon, p_full, p_half, albedo, q, t, t_surf_rad, tdt, *, cfg: RadConfig | None = None,
              z_full=None, z_half=None, t_half=None, o3f=None, diagnostics: bool = False):
    """The radiation step of run_rrtmg.  Returns (tdt, coszen, flux_sw, flux_lw) with tdt = input tdt + radiative
    heating [K/s]; with diagnostics=True also a dict with tdt_rad, tdt_sw, tdt_lw, olr, isr, t_half.
    (is_, js are the index offsets the Fortran passes on to the diag manager; unused.)"""
    cfg = cfg or RadConfig()
    t = np.asfortranarray(t, dtype=np.float64)
    si, sj, sk = t.shape
    lat = _fa(lat, (si, sj), "lat"); lon = _fa(lon, (si, sj), "lon")
    p_full = _fa(p_full, (si, sj, sk), "p_full"); p_half = _fa(p_half, (si, sj, sk + 1), "p_half")
    albedo = _fa(albedo, (si, sj), "albedo"); q = _fa(q, (si, sj, sk), "q"); ts = _fa(t_surf_rad, (si, sj), "t_surf_rad")
    tdt = np.array(_fa(tdt, (si, sj, sk), "tdt"), order="F", copy=True)
    if t_half is not None:
        t_half = _fa(t_half, (si, sj, sk + 1), "t_half")
    else:
        if z_full is None or z_half is None:
            raise ValueError("run_rrtmg needs t_half or z_full + z_half")
        z_full = _fa(z_full, (si, sj, sk), "z_full"); z_half = _fa(z_half, (si, sj, sk + 1), "z_half")
    if o3f is not None:
        o3f = _fa(o3f, (si, sj, sk), "o3f")
    coszen = np.zeros((si, sj), order="F")
    flux_sw = np.zeros((si, sj), order="F"); flux_lw = np.zeros((si, sj), order="F")
    diag = {}
    if diagnostics:
        diag = dict(tdt_rad=np.zeros((si, sj, sk), order="F"), tdt_sw=np.zeros((si, sj, sk), order="F"),
                    tdt_lw=np.zeros((si, sj, sk), order="F"), olr=np.zeros((si, sj), order="F"),
                    isr=np.zeros((si, sj), order="F"), t_half=np.zeros((si, sj, sk + 1), order="F"))
    c = cfg.to_c()
    _r._check(_r.lib().rrtmg_b200_run_rrtmg(
        C.byref(c), C.c_int(si), C.c_int(sj), C.c_int(sk), C.c_int(int(Time[0])), C.c_int(int(Time[1])),
        _p(lat), _p(lon), _p(p_full), _p(p_half), _p(albedo), _p(q), _p(t), _p(ts),
        _p(z_full), _p(z_half), _p(t_half), _p(o3f),
        _p(tdt), _p(coszen), _p(flux_sw), _p(flux_lw),
        _p(diag.get("tdt_rad")), _p(diag.get("tdt_sw")), _p(diag.get("tdt_lw")), _p(diag.get("olr")), _p(diag.get("isr")),
        _p(diag.get("t_half"))))
    if diagnostics:
        return tdt, coszen, flux_sw, flux_lw, diag
    return tdt, coszen, flux_sw, flux_lw
