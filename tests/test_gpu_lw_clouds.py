"""GPU parity of the cloudy longwave branches: cloud optical depth given per band (inflglw = 0, cldprop.f90:154-176),
random overlap through rtrn (icld = 1, rrtmg_lw_rtrn.f90:302-485) and maximum/random overlap through rtrnmr
(icld = 2, 3, rrtmg_lw_rtrnmr.f90:316-479, 569-588, 653-674).  MiMA itself runs clear sky (rrtm_radiation.f90:722-748);
these are the branches the rrtmg_lw signature advertises around the hot path."""
import numpy as np
import pytest

from mima_b200.columns import make_columns
from test_gpu_parity import LW_OUT
from test_gpu_parity import _check_outputs as _check
from test_oracle_lw_clouds import cloud_field

pytestmark = pytest.mark.gpu


def _check_outputs(got, ref, names=LW_OUT):
    _check(got, ref, names, hr_tight=1e-6)


@pytest.fixture(scope="module")
def cols():
    return make_columns("T42L40", nlon=64, nlat=8)


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_fractional_clouds(gpu, oracle, cols, icld):
    cl = cloud_field(cols, np.random.default_rng(30 + icld))
    got = gpu.lw_from_columns(cols, icld=icld, clouds=cl)
    _check_outputs(got, oracle.rrtmg_lw(cols, icld=icld, clouds=cl))
    clear = gpu.lw_from_columns(cols)
    cloudy = cl["cldfr"].max(axis=1) >= 1e-6
    assert cloudy.any() and (~cloudy).any()
    for i in (3, 4, 5):          # the clear-sky stream does not see the clouds
        tol = 1e-6 if i == 5 else 1e-9 * np.abs(clear[i]).max()
        assert np.max(np.abs(got[i] - clear[i])) < tol
    for i in (0, 1, 2):          # columns without cloud: total = clear
        assert np.array_equal(got[i][~cloudy], got[i + 3][~cloudy])
    assert np.abs(got[0][cloudy] - got[3][cloudy]).max() > 1.0


def test_overcast_and_thin_clouds(gpu, oracle, cols):
    """cldfrac exactly 1 (facclr2 guard, rtrnmr.f90:350), optical depths on both sides of the 0.06 series switch."""
    rng = np.random.default_rng(9)
    cl = cloud_field(cols, rng)
    cl["cldfr"][...] = np.where(cl["cldfr"] > 0, 1.0, 0.0)
    cl["taucld"][...] = cl["taucld"] * 10.0 ** rng.uniform(-4, 0, cl["taucld"].shape)
    for icld in (1, 2):
        _check_outputs(gpu.lw_from_columns(cols, icld=icld, clouds=cl), oracle.rrtmg_lw(cols, icld=icld, clouds=cl))


def test_clouds_with_aerosol_emissivity_and_derivative(gpu, oracle, cols):
    c = cols.take(np.arange(211))
    rng = np.random.default_rng(12)
    c.emis = np.asfortranarray(rng.uniform(0.85, 1.0, (c.ncol, 16)))
    taer = np.asfortranarray(rng.uniform(0.0, 0.05, (c.ncol, c.nlay, 16)))
    cl = cloud_field(c, rng)
    names = LW_OUT + ("duflx_dt", "duflxc_dt")
    for icld in (1, 2):
        ref = oracle.rrtmg_lw(c, icld=icld, clouds=cl, tauaer=taer, idrv=1)
        _check_outputs(gpu.lw_from_columns(c, tauaer=taer, idrv=1, icld=icld, clouds=cl), ref, names)
    gpu.set_option("host_chunk", 64)
    try:
        _check_outputs(gpu.lw_from_columns(c, tauaer=taer, idrv=1, icld=2, clouds=cl), ref, names)
    finally:
        gpu.set_option("host_chunk", 0)
    gpu.set_option("chunk", 50)
    try:
        _check_outputs(gpu.lw_from_columns(c, tauaer=taer, idrv=1, icld=2, clouds=cl), ref, names)
    finally:
        gpu.set_option("chunk", 0)


def test_zero_cloud_fraction_reproduces_the_clear_kernel(gpu, cols):
    z = np.zeros((cols.ncol, cols.nlay), order="F")
    t = np.full((16, cols.ncol, cols.nlay), 2.0, order="F")
    ref = gpu.lw_from_columns(cols)
    for icld in (1, 2):
        got = gpu.lw_from_columns(cols, icld=icld, clouds=dict(cldfr=z, taucld=t))
        for g, r, n in zip(got, ref, LW_OUT):
            tol = 1e-7 if "hr" in n else 1e-9 * np.abs(r).max()
            assert np.max(np.abs(g - r)) < tol, n


def test_argument_errors(gpu, cols):
    c = cols.take(np.arange(8))
    cl = cloud_field(c, np.random.default_rng(1))
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.lw_from_columns(c, icld=2)                       # no cloud arrays
    assert e.value.code == 4
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.lw_from_columns(c, icld=2, clouds=cl, inflglw=2)  # water-path cloud optics without the water paths
    assert e.value.code == 4
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.lw_from_columns(c, icld=2, clouds=cl, inflglw=3)
    assert e.value.code == 4


def _water_clouds(c, rng):
    shp = (c.ncol, c.nlay)
    cf = np.asfortranarray((rng.uniform(size=shp) < 0.3) * rng.uniform(0.1, 1.0, shp))
    return dict(cldfr=cf,
                cicewp=np.asfortranarray(rng.uniform(0, 30, shp) * (rng.uniform(size=shp) < 0.7)),
                cliqwp=np.asfortranarray(rng.uniform(0, 60, shp) * (rng.uniform(size=shp) < 0.7)),
                reice=np.asfortranarray(rng.uniform(14, 120, shp)), reliq=np.asfortranarray(rng.uniform(3, 50, shp)))


@pytest.mark.parametrize("flags", [(1, 0, 0), (2, 0, 0), (2, 1, 0), (2, 1, 1), (2, 2, 1), (2, 3, 1), (2, 2, 0)])
def test_cloud_optics_from_water_paths(gpu, oracle, cols, flags):
    """cldprop's parameterisations (cldprop.f90:176-270): inflglw = 1 (one absorption coefficient), 2 with the four ice
    and two liquid options; ncbands = 1, 5 or 16 decides which cloud band a spectral band reads (ipat)."""
    infl, ice, liq = flags
    cl = _water_clouds(cols, np.random.default_rng(40 + 10 * ice + liq))
    for icld in (1, 2):
        got = gpu.lw_from_columns(cols, icld=icld, clouds=cl, inflglw=infl, iceflglw=ice, liqflglw=liq, idrv=1)
        ref = oracle.rrtmg_lw(cols, icld=icld, clouds=cl, inflglw=infl, iceflglw=ice, liqflglw=liq, idrv=1)
        _check_outputs(got, ref, LW_OUT + ("duflx_dt", "duflxc_dt"))


def test_radius_out_of_range_is_an_error(gpu, cols):
    c = cols.take(np.arange(40))
    cl = _water_clouds(c, np.random.default_rng(2))
    cl["cldfr"][3, 5] = 0.5; cl["cicewp"][3, 5] = 10.0; cl["reice"][3, 5] = 4.0
    for ice, code in ((0, 7), (1, 7), (2, 7), (3, 7)):
        with pytest.raises(gpu.RRTMGError) as e:
            gpu.lw_from_columns(c, icld=2, clouds=cl, inflglw=2, iceflglw=ice, liqflglw=1)
        assert e.value.code == code
    cl["reice"][3, 5] = 50.0
    cl["cliqwp"][3, 5] = 5.0; cl["reliq"][3, 5] = 70.0
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.lw_from_columns(c, icld=2, clouds=cl, inflglw=2, iceflglw=2, liqflglw=1)
    assert e.value.code == 7 and "LIQUID" in str(e.value)
    gpu.lw_from_columns(c)          # the library stays usable


def test_eighty_layers(gpu, oracle):
    c = make_columns("T341L80", nlon=32, nlat=4)
    cl = cloud_field(c, np.random.default_rng(6))
    for icld in (1, 2):
        _check_outputs(gpu.lw_from_columns(c, icld=icld, clouds=cl), oracle.rrtmg_lw(c, icld=icld, clouds=cl))
