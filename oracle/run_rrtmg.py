"""CPU restatement (numpy) of MiMA's radiation driver around the RRTMG calls.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the checker for the device-side
`rrtmg_b200_run_rrtmg` path.  Never imported by the mima_b200 package.

Follows, statement by statement and in the same floating-point order,

    interp_temp      src/atmos_param/rrtm_radiation/rrtm_radiation.f90:422-461
    compute_zenith   src/atmos_param/rrtm_radiation/astro.f90:59-248
    run_rrtmg        src/atmos_param/rrtm_radiation/rrtm_radiation.f90:585-808
                     (from "we know now that we want to run radiation": ozone scaling, zonal-mean tracers,
                     fixed water, lon sub-sampling + vertical flip + Pa->hPa + top-interface fix, clamps,
                     the two RRTMG calls, K/day -> K/s, lon re-interpolation, surface fluxes, olr/isr)

The alarm (`dt_rad`), the netCDF interpolators and the diag manager are FMS control plane and stay on the
host of the model; their *results* (o3f, q after do_read_h2o, ...) are inputs here.

Array convention: FMS order (lon, lat, lev) with level 1 = top of the atmosphere, Fortran-ordered float64.
MiMA is compiled with -r8 (bin/mkmf.template.*), so every default `real` below is a double.

Parity status: no golden vectors exist in the reference for this wrapper; pinned by
restatement only, plus the analytic checks in tests/test_run_rrtmg_oracle.py (equinox/solstice declination,
daily-mean insolation integral, partition of unity of the lon re-interpolation).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

PI = 3.14159265358979323846      # src/shared/constants/constants.f90 (PI)


@dataclasses.dataclass
class RadConfig:
    """rrtm_radiation_nml (rrtm_radiation.f90:104-197) and astro_nml (astro.f90:24-33) entries the path reads."""
    # rrtm_radiation_nml
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
    # astro_nml
    obliq: float = 23.439
    use_dyofyr: bool = False
    solr_cnst: float = 1368.22
    solrad: float = 1.0
    solday: int = 0
    equinox_day: float = 0.25
    # calendar (time_manager: length_of_year); MiMA's default calendar has 360 days
    days_per_year: int = 360


def _seqsum0(a):
    """sum(a, 1) over the first (lon) index, added in index order (what the device kernel does; numpy's own
    sum is pairwise)."""
    s = np.zeros(a.shape[1:])
    for i in range(a.shape[0]):
        s = s + a[i]
    return s


def interp_temp(z_full, z_half, t_surf_rad, t):
    """rrtm_radiation.f90:422-461.  Returns t_half (si, sj, sk+1)."""
    si, sj, sk = t.shape
    t_half = np.zeros((si, sj, sk + 1), order="F")
    for k in range(1, sk):                       # Fortran k = 2..kend
        dzk2 = 1.0 / (z_full[:, :, k - 1] - z_full[:, :, k])
        dzk = (z_half[:, :, k] - z_full[:, :, k]) * dzk2
        dzk1 = (z_full[:, :, k - 1] - z_half[:, :, k]) * dzk2
        t_half[:, :, k] = t[:, :, k] * dzk1 + t[:, :, k - 1] * dzk
    t_half[:, :, 0] = 0.5 * (3 * t[:, :, 0] - t[:, :, 1])
    t_half[:, :, sk] = t_surf_rad
    return t_half


def local_time(cfg: RadConfig, seconds: int, days: int):
    """Time_loc of run_rrtmg (rrtm_radiation.f90:550-558) as (seconds, days)."""
    if cfg.solday > 0:
        return seconds, cfg.solday
    if cfg.slowdown_rad != 1.0:
        tot = days * 86400 + seconds
        tot = int(tot * cfg.slowdown_rad)
        return tot % 86400, tot // 86400
    return seconds, days


def compute_zenith(cfg: RadConfig, seconds: int, days: int, dt: int, lat, lon):
    """astro.f90:59-248.  lat, lon (si, sj) in radians.  Returns (cosz, dyofyr)."""
    deg2rad = PI / 180.0
    twopi = 2 * PI
    daysperyear = cfg.days_per_year
    radpersec = 2 * PI / 86400.0
    radperday = 2 * PI / daysperyear
    radsec = seconds * radpersec
    dt_pi = dt * radpersec
    time_pi = np.mod(radsec + lon, 2 * PI) - PI          # Fortran modulo: result has the sign of the divisor
    time_pi = np.where(time_pi >= PI, time_pi - twopi, time_pi)
    time_pi = np.where(time_pi < -PI, time_pi + twopi, time_pi)
    days = days - int(cfg.equinox_day * daysperyear)
    dyofyr = days % daysperyear                            # modulo
    radday = dyofyr * radperday
    dec_sin = math.sin(cfg.obliq * deg2rad) * math.sin(radday)
    dec = math.asin(dec_sin)
    dec_cos = math.cos(dec)
    dec_tan = math.tan(dec)
    eps = 1.0e-05
    lat_h = np.where(lat == 0.5 * PI, lat - eps, lat)
    lat_h = np.where(lat_h == -0.5 * PI, lat + eps, lat_h)
    cos_h = -np.tan(lat_h) * dec_tan
    h = np.where(cos_h <= -1.0, PI, np.where(cos_h >= 1.0, 0.0, np.arccos(np.clip(cos_h, -1.0, 1.0))))
    aa = np.sin(lat) * dec_sin
    bb = np.cos(lat) * dec_cos
    if 0 < dt < 86400:
        tt = time_pi + dt_pi
        st = np.sin(time_pi)
        stt = np.sin(tt)
        sh = np.sin(h)
        cosz = np.zeros_like(lat)
        with np.errstate(divide="ignore", invalid="ignore"):
            cosz = np.where((time_pi < -h) & (tt < -h), 0.0, cosz)
            cosz = np.where(((tt + h) != 0.0) & (time_pi < -h) & (np.abs(tt) <= h), aa + bb * (stt + sh) / (tt + h), cosz)
            cosz = np.where((time_pi < -h) & (h != 0.0) & (h < tt), aa + bb * (sh + sh) / (h + h), cosz)
            cosz = np.where((np.abs(time_pi) <= h) & (np.abs(tt) <= h), aa + bb * (stt - st) / (tt - time_pi), cosz)
            cosz = np.where(((h - time_pi) != 0.0) & (np.abs(time_pi) <= h) & (h < tt), aa + bb * (sh - st) / (h - time_pi), cosz)
            cosz = np.where((twopi - h < tt) & ((tt + h - twopi) != 0.0) & (time_pi <= h),
                            (cosz * (h - time_pi) + (aa * (tt + h - twopi) + bb * (stt + sh))) / ((h - time_pi) + (tt + h - twopi)),
                            cosz)
            cosz = np.where((h < time_pi) & (twopi - h >= tt), 0.0, cosz)
            cosz = np.where((h < time_pi) & (twopi - h < tt), aa + bb * (stt + sh) / (tt + h - twopi), cosz)
            # `fracday` is an uninitialised local in the reference; every point falls in one of the first six
            # `where` masks (they partition the (time_pi, tt) plane), so no undefined value is ever used
            fracday = np.zeros_like(lat)
            fracday = np.where((time_pi < -h) & (tt < -h), 0.0, fracday)
            fracday = np.where((time_pi < -h) & (np.abs(tt) <= h), (tt + h) / dt, fracday)
            fracday = np.where((time_pi < -h) & (h < tt), (h + h) / dt, fracday)
            fracday = np.where((np.abs(time_pi) <= h) & (np.abs(tt) <= h), (tt - time_pi) / dt, fracday)
            fracday = np.where((np.abs(time_pi) <= h) & (h < tt), (h - time_pi) / dt, fracday)
            fracday = np.where(h < time_pi, 0.0, fracday)
            fracday = np.where(twopi - h < tt, fracday + (tt + h - twopi) / dt, fracday)
        cosz = cosz * fracday / radpersec
    elif dt >= 86400:
        cosz = (aa * h + bb * np.sin(h)) / PI
    else:
        cosz = np.where(np.abs(time_pi) <= h, aa + bb * np.cos(time_pi), 0.0)
    cosz = np.maximum(0.0, cosz)
    return np.asfortranarray(cosz), int(dyofyr)


def pack_columns(cfg: RadConfig, p_full, p_half, t, t_half, q, o3f, coszen, albedo, t_surf_rad):
    """rrtm_radiation.f90:619-677: what run_rrtmg hands to rrtmg_sw / rrtmg_lw.  Returns a dict of
    (ncols_rrt, nlay) / (ncols_rrt, nlay+1) / (ncols_rrt,) arrays, level 1 = surface, hPa."""
    si, sj, sk = t.shape
    ls = cfg.lonstep
    if cfg.do_zm_tracers:
        q_tmp = np.broadcast_to(_seqsum0(q) / si, q.shape).copy()               # :622-626
    else:
        q_tmp = q.copy()
    # (do_read_h2o: the interpolated field is the caller's q; :631-635)
    if cfg.do_fixed_water:                                                   # :638-646
        raise NotImplementedError("do_fixed_water needs lat; use run_rrtmg")

    def resh(a):        # reshape(a(1:si:lonstep,:,n:1:-1), (/ si*sj/lonstep, n /))
        b = a[::ls, :, ::-1]
        return np.asfortranarray(b.reshape((b.shape[0] * b.shape[1], b.shape[2]), order="F"))

    pfull = resh(p_full) * 0.01
    phalf = resh(p_half) * 0.01
    if np.min(phalf[:, sk]) <= 0.0:                                          # :655-656
        phalf[:, sk] = pfull[:, sk - 1] * 0.5
    tfull = resh(t)
    thalf = resh(t_half)
    h2o = resh(q_tmp)
    ncols = pfull.shape[0]
    if o3f is not None:
        o3 = resh(o3f)
    else:
        o3 = np.full((ncols, sk), cfg.o3_val, order="F")
    cosz_rr = np.asfortranarray(coszen[::ls, :].reshape(-1, order="F"))
    albedo_rr = np.asfortranarray(albedo[::ls, :].reshape(-1, order="F"))
    tsrf = np.asfortranarray(t_surf_rad[::ls, :].reshape(-1, order="F"))
    h2o = np.maximum(h2o, cfg.h2o_lower_limit)                               # :673-677
    tfull = np.minimum(np.maximum(tfull, cfg.temp_lower_limit), cfg.temp_upper_limit)
    thalf = np.minimum(np.maximum(thalf, cfg.temp_lower_limit), cfg.temp_upper_limit)
    return dict(pfull=pfull, phalf=phalf, tfull=tfull, thalf=thalf, h2o=h2o, o3=o3, cosz_rr=cosz_rr,
                albedo_rr=albedo_rr, tsrf=tsrf)


def run_rrtmg(oracle, cfg: RadConfig, seconds: int, days: int, lat, lon, p_full, p_half, albedo, q, t,
              t_surf_rad, tdt, *, z_full=None, z_half=None, t_half=None, o3f=None):
    """The radiation step of run_rrtmg.  `oracle` is oracle.pyoracle.Oracle (the RRTMG restatement).
    tdt is returned updated (tdt + tdt_rrtm), not modified in place."""
    from mima_b200.columns import Columns      # plain container, no compute

    si, sj, sk = t.shape
    ls = cfg.lonstep
    sec_l, day_l = local_time(cfg, seconds, days)
    dt = cfg.dt_rad_avg if cfg.do_rad_time_avg else 0                        # :562-566
    coszen, dyofyr = compute_zenith(cfg, sec_l, day_l, dt, lat, lon)
    if not cfg.use_dyofyr:
        dyofyr = 0                                                           # :592
    if t_half is None:
        t_half = interp_temp(z_full, z_half, t_surf_rad, t)
    if o3f is not None:
        o3f = np.maximum(0.0, o3f * cfg.scale_ozone)                         # :603-605
    qq = q
    if cfg.do_zm_tracers:
        qq = np.broadcast_to(_seqsum0(q) / si, q.shape).copy()
    if cfg.do_fixed_water:                                                   # :638-646
        qq = qq.copy()
        m = (np.abs(lat) <= cfg.fixed_water_lat)[:, :, None] & (p_full <= cfg.fixed_water_pres * 100.0)
        qq[m] = cfg.fixed_water
    c2 = dataclasses.replace(cfg, do_zm_tracers=False, do_fixed_water=False)
    pk = pack_columns(c2, p_full, p_half, t, t_half, qq, o3f, coszen, albedo, t_surf_rad)
    ncols = pk["pfull"].shape[0]
    ones = np.ones((ncols, sk), order="F")
    zeros = np.zeros((ncols, sk), order="F")
    sec = cfg.include_secondary_gases
    cols = Columns(
        ncol=ncols, nlay=sk, nlon=si // ls, nlat=sj,
        play=pk["pfull"], plev=pk["phalf"], tlay=pk["tfull"], tlev=pk["thalf"], tsfc=pk["tsrf"],
        h2o=pk["h2o"], o3=pk["o3"], co2=np.asfortranarray(cfg.co2ppmv * 1.e-6 * ones),
        ch4=cfg.ch4_val * ones if sec else zeros, n2o=cfg.n2o_val * ones if sec else zeros,
        o2=cfg.o2_val * ones if sec else zeros,
        cfc11=cfg.cfc11_val * ones if sec else zeros, cfc12=cfg.cfc12_val * ones if sec else zeros,
        cfc22=cfg.cfc22_val * ones if sec else zeros, ccl4=cfg.ccl4_val * ones if sec else zeros,
        emis=np.ones((ncols, 16), order="F"), albedo=pk["albedo_rr"], coszen=pk["cosz_rr"],
        scon=cfg.solr_cnst, adjes=cfg.solrad, dyofyr=dyofyr)
    sw = oracle.rrtmg_sw(cols)
    lw = oracle.rrtmg_lw(cols)
    daypersec = 1. / 86400
    ni = si // ls

    def unresh(a):      # reshape(a(:,sk:1:-1), (/ si/lonstep, sj, sk /))
        return a[:, ::-1].reshape((ni, sj, a.shape[1]), order="F")

    swijk = unresh(sw["swhr"]) * daypersec                                   # :715
    isrijk = (sw["swdflx"][:, sk] - sw["swuflx"][:, sk]).reshape((ni, sj), order="F")
    lwijk = unresh(lw["hr"]) * daypersec                                     # :751
    olrijk = lw["uflx"][:, sk].reshape((ni, sj), order="F")
    swflxijk = (sw["swdflx"][:, 0] - sw["swuflx"][:, 0]).reshape((ni, sj), order="F")   # :777
    lwflxijk = lw["dflx"][:, 0].reshape((ni, sj), order="F")
    tdt_rrtm = np.zeros((si, sj, sk), order="F")
    tdt_sw = np.zeros_like(tdt_rrtm)
    tdt_lw = np.zeros_like(tdt_rrtm)
    olr = np.zeros((si, sj), order="F")
    isr = np.zeros_like(olr)
    flux_sw = np.zeros_like(olr)
    flux_lw = np.zeros_like(olr)
    dlon = 1. / ls
    for i in range(ni):                                                      # :759-780
        i1 = i + 1
        if i1 > ni - 1:
            i1 = 0
        for ij in range(ls):
            di = ij * dlon
            ij1 = i * ls + ij
            if cfg.do_zm_rad:
                tdt_rrtm[ij1] = _seqsum0(swijk + lwijk) / max(1, ni)
                flux_sw[ij1] = _seqsum0(swflxijk) / max(1, ni)
                flux_lw[ij1] = _seqsum0(lwflxijk) / max(1, ni)
            else:
                tdt_rrtm[ij1] = di * (swijk[i1] + lwijk[i1]) + (1. - di) * (swijk[i] + lwijk[i])
                flux_sw[ij1] = di * swflxijk[i1] + (1. - di) * swflxijk[i]
                flux_lw[ij1] = di * lwflxijk[i1] + (1. - di) * lwflxijk[i]
            tdt_sw[ij1] = di * swijk[i1] + (1. - di) * swijk[i]
            tdt_lw[ij1] = di * lwijk[i1] + (1. - di) * lwijk[i]
            olr[ij1] = di * olrijk[i1] + (1. - di) * olrijk[i]
            isr[ij1] = di * isrijk[i1] + (1. - di) * isrijk[i]
    return dict(tdt=np.asfortranarray(tdt + tdt_rrtm), tdt_rrtm=tdt_rrtm, coszen=coszen, flux_sw=flux_sw,
                flux_lw=flux_lw, tdt_sw=tdt_sw, tdt_lw=tdt_lw, olr=olr, isr=isr, t_half=t_half,
                packed=pk, dyofyr=dyofyr, sw=sw, lw=lw)
