"""GPU parity of the device-side radiation driver (rrtmg_b200_run_rrtmg, _compute_zenith, _interp_temp) against
oracle/run_rrtmg.py, through the C ABI via the host mirror mima_b200.rrtm_radiation.

Tolerances: cos(zenith) 1e-14 absolute (device libm vs host libm, a few ulp on values <= 1); t_half bit-exact
(same operations in the same order, FMA contraction off); heating rates 1e-4 K/day and fluxes 1e-6 relative
(north_star), tighter bounds asserted as regression guards."""
import dataclasses

import numpy as np
import pytest

from mima_b200.columns import make_gcm_state
from oracle import run_rrtmg as R

pytestmark = pytest.mark.gpu


def _cfg_pair(**kw):
    from mima_b200 import rrtm_radiation as rr
    return rr.RadConfig(**kw), R.RadConfig(**kw)


def _lat_lon(n=32, m=17):
    lat = np.asfortranarray(np.broadcast_to(np.linspace(-np.pi / 2, np.pi / 2, m)[None, :], (n, m)))
    lon = np.asfortranarray(np.broadcast_to((np.arange(n) * 2 * np.pi / n)[:, None], (n, m)))
    return lat, lon


@pytest.mark.parametrize("dt", [0, 1800, 3 * 3600, 11 * 3600, 86400])
def test_compute_zenith(gpu, dt):
    from mima_b200 import rrtm_radiation as rr
    lat, lon = _lat_lon()
    for sec, day in ((0, 90), (3 * 3600 + 17, 123), (20 * 3600, 271), (86399, 359), (43200, 0)):
        g, dy = rr.compute_zenith((sec, day), 0.25, dt, lat, lon)
        o, dyo = R.compute_zenith(R.RadConfig(), sec, day, dt, lat, lon)
        assert dy == dyo
        assert np.isfinite(g).all()
        assert np.max(np.abs(g - o)) < 1e-14, (dt, sec, day, float(np.max(np.abs(g - o))))


def test_interp_temp_bit_exact(gpu):
    from mima_b200 import rrtm_radiation as rr
    g = make_gcm_state("T42L40", nlon=16, nlat=8)
    th = rr.interp_temp(g["z_full"], g["z_half"], g["t_surf"], g["t"])
    np.testing.assert_array_equal(th, R.interp_temp(g["z_full"], g["z_half"], g["t_surf"], g["t"]))


def _compare(gpu_out, orc, sk):
    tdt, coszen, fsw, flw, diag = gpu_out
    assert np.max(np.abs(coszen - orc["coszen"])) < 1e-14
    for name, a, b in (("tdt", tdt, orc["tdt"]), ("tdt_rad", diag["tdt_rad"], orc["tdt_rrtm"]),
                       ("tdt_sw", diag["tdt_sw"], orc["tdt_sw"]), ("tdt_lw", diag["tdt_lw"], orc["tdt_lw"])):
        err = np.max(np.abs(a - b)) * 86400.0            # K/day
        assert err < 1e-4, (name, float(err))
        assert err < 1e-7, (name, float(err))
    for name, a, b in (("flux_sw", fsw, orc["flux_sw"]), ("flux_lw", flw, orc["flux_lw"]), ("olr", diag["olr"], orc["olr"]),
                       ("isr", diag["isr"], orc["isr"])):
        scale = np.maximum(np.abs(b), 1e-6 * np.abs(b).max() + 1e-300)
        r = np.max(np.abs(a - b) / scale)
        assert r < 1e-6, (name, float(r))
        assert r < 1e-9, (name, float(r))
    np.testing.assert_array_equal(diag["t_half"], orc["t_half"])


CASES = [
    dict(co2ppmv=390.0),                                                        # MiMA defaults: daily-mean sun
    dict(co2ppmv=390.0, lonstep=2, do_rad_time_avg=False),                      # sub-sampled lon, instantaneous sun (night columns)
    dict(co2ppmv=1560.0, lonstep=4, dt_rad_avg=7200, include_secondary_gases=True, ch4_val=1.8e-6, n2o_val=3.2e-7,
         o2_val=0.209, cfc11_val=2.5e-10, cfc12_val=5.3e-10, cfc22_val=2.0e-10, ccl4_val=9.0e-11, scale_ozone=0.8),
    dict(co2ppmv=300.0, do_zm_tracers=True, do_zm_rad=True),
    dict(co2ppmv=300.0, do_fixed_water=True, fixed_water_lat=0.6, fixed_water_pres=150.0, use_dyofyr=True,
         days_per_year=365, solday=200),
    dict(co2ppmv=300.0, slowdown_rad=0.5, do_rad_time_avg=False, h2o_lower_limit=5e-6, temp_lower_limit=215.0,
         temp_upper_limit=290.0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_run_rrtmg_matches_oracle(gpu, oracle, case):
    from mima_b200 import rrtm_radiation as rr
    gc, oc = _cfg_pair(**CASES[case])
    g = make_gcm_state("T42L40", nlon=16, nlat=8, ozone="file" if case == 2 else "analytic")
    tdt0 = np.asfortranarray(np.random.default_rng(case).normal(0, 1e-5, g["t"].shape))
    sec, day = 7 * 3600 + 11, 140 + 37 * case
    args = (g["lat"], g["lon"], g["p_full"], g["p_half"], g["albedo"], g["q"], g["t"], g["t_surf"], tdt0)
    out = rr.run_rrtmg(1, 1, (sec, day), *args, cfg=gc, z_full=g["z_full"], z_half=g["z_half"], o3f=g["o3f"], diagnostics=True)
    orc = R.run_rrtmg(oracle, oc, sec, day, *args, z_full=g["z_full"], z_half=g["z_half"], o3f=g["o3f"])
    _compare(out, orc, g["sk"])


def test_run_rrtmg_with_given_t_half_and_constant_ozone(gpu, oracle):
    from mima_b200 import rrtm_radiation as rr
    gc, oc = _cfg_pair(co2ppmv=390.0, o3_val=2.0e-6)
    g = make_gcm_state("T42L40", nlon=8, nlat=4)
    th = R.interp_temp(g["z_full"], g["z_half"], g["t_surf"], g["t"]) + 0.25
    args = (g["lat"], g["lon"], g["p_full"], g["p_half"], g["albedo"], g["q"], g["t"], g["t_surf"], g["tdt"])
    out = rr.run_rrtmg(1, 1, (0, 10), *args, cfg=gc, t_half=th, diagnostics=True)
    orc = R.run_rrtmg(oracle, oc, 0, 10, *args, t_half=th)
    _compare(out, orc, g["sk"])


def test_run_rrtmg_argument_errors(gpu):
    from mima_b200 import rrtm_radiation as rr, rrtmg
    g = make_gcm_state("T42L40", nlon=6, nlat=2)
    args = (g["lat"], g["lon"], g["p_full"], g["p_half"], g["albedo"], g["q"], g["t"], g["t_surf"], g["tdt"])
    with pytest.raises(rrtmg.RRTMGError) as e:      # lonstep must divide the number of longitudes
        rr.run_rrtmg(1, 1, (0, 0), *args, cfg=rr.RadConfig(lonstep=4), z_full=g["z_full"], z_half=g["z_half"])
    assert e.value.code == 4
    with pytest.raises(ValueError):
        rr.run_rrtmg(1, 1, (0, 0), *args)


def test_run_rrtmg_row_block_pipeline_is_bitwise_invariant(gpu):
    """The host-pointer entry cuts the grid into blocks of latitude rows (two-stage pipeline); every quantity of
    run_rrtmg is local to a row, so any block size must give bit-identical results -- also with zonal means."""
    from mima_b200 import rrtm_radiation as rr
    g = make_gcm_state("T42L40", nlon=32, nlat=16)
    args = (g["lat"], g["lon"], g["p_full"], g["p_half"], g["albedo"], g["q"], g["t"], g["t_surf"], g["tdt"])
    for cfg in (rr.RadConfig(co2ppmv=390.0, lonstep=2, do_rad_time_avg=False),
                rr.RadConfig(co2ppmv=390.0, do_zm_tracers=True, do_zm_rad=True)):
        ref = rr.run_rrtmg(1, 1, (3600, 77), *args, cfg=cfg, z_full=g["z_full"], z_half=g["z_half"], o3f=g["o3f"], diagnostics=True)
        for chunk in (32, 96, 200):
            gpu.set_option("chunk", chunk)
            try:
                out = rr.run_rrtmg(1, 1, (3600, 77), *args, cfg=cfg, z_full=g["z_full"], z_half=g["z_half"], o3f=g["o3f"],
                                   diagnostics=True)
            finally:
                gpu.set_option("chunk", 0)
            for a, b in zip(ref[:4], out[:4]):
                assert np.array_equal(a, b)
            for k in ref[4]:
                assert np.array_equal(ref[4][k], out[4][k]), k


def test_run_rrtmg_full_size_t170(gpu):
    """BASELINE config 3 (512 x 256 x 60) through the driver: finite, tendency = SW + LW parts, surface and TOA
    diagnostics consistent, daily-mean sun never negative, polar night has no SW heating."""
    from mima_b200 import rrtm_radiation as rr
    g = make_gcm_state("T170L60")
    cfg = rr.RadConfig(co2ppmv=390.0)
    tdt, cz, fsw, flw, d = rr.run_rrtmg(1, 1, (0, 180), g["lat"], g["lon"], g["p_full"], g["p_half"], g["albedo"], g["q"], g["t"],
                                        g["t_surf"], g["tdt"], cfg=cfg, z_full=g["z_full"], z_half=g["z_half"], o3f=g["o3f"],
                                        diagnostics=True)
    for a in (tdt, cz, fsw, flw, d["tdt_sw"], d["tdt_lw"], d["olr"], d["isr"]):
        assert np.isfinite(a).all()
    np.testing.assert_allclose(tdt, d["tdt_sw"] + d["tdt_lw"], rtol=1e-12, atol=1e-18)
    np.testing.assert_array_equal(tdt, d["tdt_rad"])
    assert (cz >= 0).all() and cz.max() <= 1.0
    night = cz == 0.0
    assert night.any()                                  # day 180 - 90 = northern solstice: southern polar night
    assert (fsw[night] == 0).all() and (d["isr"][night] == 0).all() and (d["tdt_sw"][night] == 0).all()
    # isr is the NET incoming SW at the top (down minus reflected, rrtm_radiation.f90:716): positive, below S0 cos(z)
    assert (d["isr"][~night] > 0).all() and (d["isr"] <= cfg.solr_cnst * cz * (1 + 2e-5)).all()
    assert (fsw <= d["isr"] * (1 + 1e-12)).all()            # the surface absorbs no more than enters at the top
    assert (flw > 50).all() and (d["olr"] > 50).all()
