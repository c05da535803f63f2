"""CPU tests of the numpy restatement of the radiation driver (oracle/run_rrtmg.py): interp_temp
(rrtm_radiation.f90:422-461), compute_zenith (astro.f90:59-248) and the marshaling of run_rrtmg
(rrtm_radiation.f90:585-808).  The reference holds no golden vectors for this wrapper, so the restatement is
pinned by analytic known answers."""
import dataclasses

import numpy as np

from mima_b200.columns import make_columns, make_gcm_state
from oracle import run_rrtmg as R

PI = R.PI


def _grid(n=24, m=13):
    lat = np.asfortranarray(np.broadcast_to(np.linspace(-1.45, 1.45, m)[None, :], (n, m)))
    lon = np.asfortranarray(np.broadcast_to((np.arange(n) * 2 * PI / n)[:, None], (n, m)))
    return lat, lon


def test_zenith_equinox_instantaneous_and_declination():
    cfg = R.RadConfig(days_per_year=360, equinox_day=0.25)
    lat, lon = _grid()
    # day 90 = 0.25*360: declination 0 -> cosz = cos(lat) cos(local time)
    cz, dy = R.compute_zenith(cfg, 6 * 3600, 90, 0, lat, lon)
    assert dy == 0
    tloc = np.mod(6 * 3600 * 2 * PI / 86400 + lon, 2 * PI) - PI
    np.testing.assert_allclose(cz, np.maximum(0.0, np.cos(lat) * np.cos(tloc)), atol=1e-15)
    # a quarter year later: declination = obliquity, the pole is sunlit all day
    cz, dy = R.compute_zenith(cfg, 0, 180, 0, lat, lon)
    assert dy == 90
    dec = np.arcsin(np.sin(np.radians(cfg.obliq)))
    np.testing.assert_allclose(np.degrees(dec), cfg.obliq, rtol=1e-14)
    assert (cz[:, -1] > 0).all() and (cz[:, 0] == 0).all()


def test_zenith_daily_mean_is_the_integral_of_the_instantaneous_value():
    cfg = R.RadConfig()
    lat, lon = _grid(8, 9)
    for day in (90, 140, 200, 300):
        mean, _ = R.compute_zenith(cfg, 0, day, 86400, lat, lon)
        n = 2880
        acc = np.zeros_like(mean)
        for s in range(n):
            acc += R.compute_zenith(cfg, int(s * 86400 / n), day, 0, lat, lon)[0]
        np.testing.assert_allclose(mean, acc / n, atol=2e-4)


def test_zenith_window_average_matches_quadrature():
    cfg = R.RadConfig()
    lat, lon = _grid(12, 7)
    dt = 3 * 3600
    for sec in (0, 5 * 3600, 17 * 3600 + 1800):
        avg, _ = R.compute_zenith(cfg, sec, 250, dt, lat, lon)
        n = 720
        acc = np.zeros_like(avg)
        for s in range(n):
            # midpoint rule over [sec, sec + dt); the day index is held fixed as in the reference
            tsec = sec + (s + 0.5) * dt / n
            c2 = dataclasses.replace(cfg)
            radsec = tsec * 2 * PI / 86400
            tloc = np.mod(radsec + lon, 2 * PI) - PI
            full, _ = R.compute_zenith(c2, 0, 250, 0, lat, tloc + PI)      # lon chosen so that local time = tloc
            acc += full
        np.testing.assert_allclose(avg, acc / n, atol=5e-4)


def test_interp_temp_reproduces_a_linear_profile():
    rng = np.random.default_rng(3)
    si, sj, sk = 5, 4, 9
    z_half = np.zeros((si, sj, sk + 1), order="F")
    z_half[:, :, :-1] = np.cumsum(rng.uniform(300., 900., (si, sj, sk)), axis=2)[:, :, ::-1]
    z_full = 0.5 * (z_half[:, :, :-1] + z_half[:, :, 1:]) + rng.uniform(-50, 50, (si, sj, sk))
    t = 290.0 - 6.5e-3 * z_full
    ts = rng.uniform(280, 300, (si, sj))
    th = R.interp_temp(z_full, z_half, ts, t)
    np.testing.assert_allclose(th[:, :, 1:-1], 290.0 - 6.5e-3 * z_half[:, :, 1:-1], rtol=1e-13)
    np.testing.assert_array_equal(th[:, :, -1], ts)
    np.testing.assert_allclose(th[:, :, 0], 0.5 * (3 * t[:, :, 0] - t[:, :, 1]), rtol=0, atol=0)


def test_pack_reproduces_the_column_layout():
    """The FMS-ordered state goes back to exactly the (ncol, nlay) surface-first hPa arrays of make_columns
    (rrtm_radiation.f90:652-664), including the top-interface replacement (:655-656) and the clamps (:673-677)."""
    kw = dict(nlon=8, nlat=4)
    cols = make_columns("T42L40", **kw)
    g = make_gcm_state("T42L40", **kw)
    cfg = R.RadConfig()
    th = np.asfortranarray(cols.tlev.reshape((8, 4, 41), order="F")[:, :, ::-1])
    cz = np.asfortranarray(cols.coszen.reshape((8, 4), order="F"))
    pk = R.pack_columns(cfg, g["p_full"], g["p_half"], g["t"], th, g["q"], g["o3f"], cz, g["albedo"], g["t_surf"])
    np.testing.assert_allclose(pk["pfull"], cols.play, rtol=4e-16)
    np.testing.assert_allclose(pk["phalf"], cols.plev, rtol=4e-16)
    np.testing.assert_array_equal(pk["tfull"], cols.tlay)
    np.testing.assert_array_equal(pk["thalf"], cols.tlev)
    np.testing.assert_array_equal(pk["h2o"], cols.h2o)
    np.testing.assert_array_equal(pk["o3"], cols.o3)
    np.testing.assert_array_equal(pk["cosz_rr"], cols.coszen)
    np.testing.assert_array_equal(pk["tsrf"], cols.tsfc)
    # lonstep = 2 keeps every other longitude of every row
    pk2 = R.pack_columns(dataclasses.replace(cfg, lonstep=2), g["p_full"], g["p_half"], g["t"], th, g["q"], g["o3f"], cz,
                         g["albedo"], g["t_surf"])
    idx = np.array([i + 8 * j for j in range(4) for i in range(0, 8, 2)])
    np.testing.assert_array_equal(pk2["tfull"], cols.tlay[idx])


def test_run_rrtmg_lon_reinterpolation_and_units(oracle):
    g = make_gcm_state("T42L40", nlon=8, nlat=2)
    args = (g["lat"], g["lon"], g["p_full"], g["p_half"], g["albedo"], g["q"], g["t"], g["t_surf"], g["tdt"])
    kw = dict(z_full=g["z_full"], z_half=g["z_half"], o3f=g["o3f"])
    o1 = R.run_rrtmg(oracle, R.RadConfig(co2ppmv=390.), 0, 100, *args, **kw)
    o2 = R.run_rrtmg(oracle, R.RadConfig(co2ppmv=390., lonstep=2), 0, 100, *args, **kw)
    # sampled longitudes are identical, the ones in between are the mean of their neighbours (closed toroidally)
    np.testing.assert_array_equal(o2["tdt"][0::2], o1["tdt"][0::2])
    nb = 0.5 * (o1["tdt"][0::2] + np.roll(o1["tdt"][0::2], -1, axis=0))
    np.testing.assert_allclose(o2["tdt"][1::2], nb, rtol=1e-13, atol=1e-20)
    # K/day -> K/s and the vertical flip: tdt_sw(top-first) * 86400 = swhr(surface-first) reversed
    sw = o1["sw"]["swhr"].reshape((8, 2, 40), order="F")[:, :, ::-1]
    np.testing.assert_allclose(o1["tdt_sw"] * 86400, sw, rtol=1e-14)
    # surface and top-of-atmosphere diagnostics
    np.testing.assert_array_equal(o1["flux_lw"].ravel(order="F"), o1["lw"]["dflx"][:, 0])
    np.testing.assert_array_equal(o1["olr"].ravel(order="F"), o1["lw"]["uflx"][:, 40])
    # zonal-mean radiation is constant along longitude
    o3 = R.run_rrtmg(oracle, R.RadConfig(co2ppmv=390., do_zm_rad=True), 0, 100, *args, **kw)
    assert np.ptp(o3["tdt"], axis=0).max() == 0.0 and np.ptp(o3["flux_sw"], axis=0).max() == 0.0
