"""GPU parity of the general shortwave path: aerosols given per band (iaer = 10, SW rad.nomcica:633-640) and
clouds given as optical properties (icld >= 1, inflgsw = 0: cldprop_sw's delta-M branch, rrtmg_sw_cldprop.f90:120-166;
total-sky stream of spcvrt_sw.f90:455-548).  MiMA itself never takes these branches (rrtm_radiation.f90:703-717);
they are the part of the rrtmg_sw interface around the hot path.  Same tolerances as test_gpu_parity."""
import numpy as np
import pytest

from mima_b200.columns import make_columns
from test_gpu_parity import SW_OUT
from test_gpu_parity import _check_outputs as _check

pytestmark = pytest.mark.gpu


def _check_outputs(got, ref, names):
    # thick clouds put W/m2-sized flux differences across the thin top layers: the regression guard on heating rates is
    # 1e-6 K/day here (fluxes stay at 1e-9 relative); the contract tolerance of 1e-4 K/day is asserted either way
    _check(got, ref, names, hr_tight=1e-6)


def _aerosols(cols, rng, tau_max=0.3):
    shp = (cols.ncol, cols.nlay, 14)
    return dict(tauaer=np.asfortranarray(rng.uniform(0.0, tau_max, shp)),
                ssaaer=np.asfortranarray(rng.uniform(0.6, 0.999, shp)),
                asmaer=np.asfortranarray(rng.uniform(0.3, 0.8, shp)))


def _clouds(cols, rng, frac=0.3, tau_max=20.0):
    shp = (14, cols.ncol, cols.nlay)
    cld = (rng.uniform(size=(cols.ncol, cols.nlay)) < frac).astype(np.float64)
    asm = rng.uniform(0.7, 0.9, shp)
    return dict(cldfr=np.asfortranarray(cld),
                taucld=np.asfortranarray(rng.uniform(0.0, tau_max, shp) * cld[None]),
                ssacld=np.asfortranarray(rng.uniform(0.5, 0.99999, shp)),
                asmcld=np.asfortranarray(asm),
                fsfcld=np.asfortranarray(asm * asm))


@pytest.fixture(scope="module")
def cols():
    return make_columns("T42L40", nlon=64, nlat=8, night=True)


def test_aerosols(gpu, oracle, cols):
    aer = _aerosols(cols, np.random.default_rng(11))
    got = gpu.sw_from_columns(cols, iaer=10, aerosols=aer)
    _check_outputs(got, oracle.rrtmg_sw(cols, iaer=10, aerosols=aer), SW_OUT)
    assert np.array_equal(got[0], got[3]) and np.array_equal(got[2], got[5])      # no clouds: total = clear
    clear = gpu.sw_from_columns(cols)
    day = cols.coszen > 0.1
    assert (got[1][day, 0] < clear[1][day, 0]).all()            # aerosols dim the surface


def test_ecmwf_aerosol_types(gpu, oracle, cols):
    """iaer = 6: optical depths at 0.55 micron of the six ECMWF aerosol types, spectral properties from the swaerpr tables
    (rrtmg_sw_init.f90:370-470, rad.nomcica:608-640); also together with clouds and in chunks."""
    rng = np.random.default_rng(17)
    aer = dict(ecaer=np.asfortranarray(rng.uniform(0.0, 0.05, (cols.ncol, cols.nlay, 6)) * (rng.uniform(size=(cols.ncol, cols.nlay, 1)) < 0.7)))
    got = gpu.sw_from_columns(cols, iaer=6, aerosols=aer)
    _check_outputs(got, oracle.rrtmg_sw(cols, iaer=6, aerosols=aer), SW_OUT)
    clear = gpu.sw_from_columns(cols)
    day = cols.coszen > 0.1
    assert (got[1][day, 0] < clear[1][day, 0]).all()
    c = cols.take(np.arange(150))
    aer = dict(ecaer=np.asfortranarray(aer["ecaer"][:150]))
    cl = _clouds(c, rng)
    gpu.set_option("host_chunk", 64)
    try:
        _check_outputs(gpu.sw_from_columns(c, icld=2, iaer=6, clouds=cl, aerosols=aer),
                       oracle.rrtmg_sw(c, icld=2, iaer=6, clouds=cl, aerosols=aer), SW_OUT)
    finally:
        gpu.set_option("host_chunk", 0)


def test_zero_aerosol_reproduces_the_clear_path(gpu, cols):
    """tauaer = 0 through the general kernel against MiMA's specialised kernel: the same numbers to rounding
    (the two differ only in reciprocal refinement and summation order)."""
    z = np.zeros((cols.ncol, cols.nlay, 14), order="F")
    got = gpu.sw_from_columns(cols, iaer=10, aerosols=dict(tauaer=z, ssaaer=z + 0.9, asmaer=z + 0.5))
    ref = gpu.sw_from_columns(cols)
    for g, r, n in zip(got, ref, SW_OUT):
        tol = 1e-7 if "hr" in n else 1e-9 * max(np.abs(r).max(), 1.0)
        assert np.max(np.abs(g - r)) < tol, n


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_clouds(gpu, oracle, cols, icld):
    cl = _clouds(cols, np.random.default_rng(20 + icld))
    got = gpu.sw_from_columns(cols, icld=icld, clouds=cl)
    _check_outputs(got, oracle.rrtmg_sw(cols, icld=icld, clouds=cl), SW_OUT)
    clear = gpu.sw_from_columns(cols)
    for i in (3, 4, 5):          # the clear-sky stream does not see the clouds
        tol = 1e-6 if i == 5 else 1e-9 * np.abs(clear[i]).max()
        assert np.max(np.abs(got[i] - clear[i])) < tol
    day = cols.coszen > 0.1
    cloudy = day & (cl["cldfr"].sum(axis=1) > 0) & (cl["taucld"].sum(axis=(0, 2)) > 1.0)
    assert cloudy.any() and (got[1][cloudy, 0] < got[4][cloudy, 0]).all()


def test_clouds_and_aerosols_chunked(gpu, oracle, cols):
    """Both branches at once, with host chunks and device passes smaller than the batch (ragged tail): the banded
    (14, ncol, nlay) arrays are cut by columns."""
    rng = np.random.default_rng(5)
    c = cols.take(np.arange(301))
    cl, aer = _clouds(c, rng), _aerosols(c, rng, 0.1)
    ref = oracle.rrtmg_sw(c, icld=2, iaer=10, clouds=cl, aerosols=aer)
    gpu.set_option("host_chunk", 64)
    try:
        _check_outputs(gpu.sw_from_columns(c, icld=2, iaer=10, clouds=cl, aerosols=aer), ref, SW_OUT)
    finally:
        gpu.set_option("host_chunk", 0)
    gpu.set_option("chunk", 50)
    try:
        _check_outputs(gpu.sw_from_columns(c, icld=2, iaer=10, clouds=cl, aerosols=aer), ref, SW_OUT)
    finally:
        gpu.set_option("chunk", 0)


def test_out_of_range_switches_are_reset(gpu, oracle, cols):
    """icld outside 0..3 becomes 2, iaer outside {0, 6, 10} becomes 0 (rad.nomcica:468-473)."""
    c = cols.take(np.arange(40))
    cl = _clouds(c, np.random.default_rng(3))
    got = gpu.sw_from_columns(c, icld=7, iaer=3, clouds=cl)
    _check_outputs(got, oracle.rrtmg_sw(c, icld=2, clouds=cl), SW_OUT)


def test_partial_cloud_is_an_error(gpu, cols):
    c = cols.take(np.flatnonzero(cols.coszen > 0.1)[:12].tolist() + np.flatnonzero(cols.coszen == 0)[:4].tolist())
    cl = _clouds(c, np.random.default_rng(4))
    cl["cldfr"][14, 7] = 0.5                   # in a night column: not an error, the column is skipped first
    gpu.sw_from_columns(c, icld=2, clouds=cl)
    cl["cldfr"][5, 7] = 0.5
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.sw_from_columns(c, icld=2, clouds=cl)
    assert e.value.code == 3
    gpu.sw_from_columns(c)         # the library stays usable


def test_eighty_layers(gpu, oracle):
    """L80 takes the kernels' 128-layer instantiation."""
    c = make_columns("T341L80", nlon=32, nlat=4, night=True)
    rng = np.random.default_rng(8)
    cl, aer = _clouds(c, rng), _aerosols(c, rng, 0.1)
    _check_outputs(gpu.sw_from_columns(c, icld=2, iaer=10, clouds=cl, aerosols=aer),
                   oracle.rrtmg_sw(c, icld=2, iaer=10, clouds=cl, aerosols=aer), SW_OUT)
    _check_outputs(gpu.sw_from_columns(c, iaer=10, aerosols=aer), oracle.rrtmg_sw(c, iaer=10, aerosols=aer), SW_OUT)


def _water_clouds(c, rng):
    shp = (c.ncol, c.nlay)
    return dict(cldfr=np.asfortranarray((rng.uniform(size=shp) < 0.3).astype(np.float64)),
                cicewp=np.asfortranarray(rng.uniform(0, 30, shp) * (rng.uniform(size=shp) < 0.7)),
                cliqwp=np.asfortranarray(rng.uniform(0, 60, shp) * (rng.uniform(size=shp) < 0.7)),
                reice=np.asfortranarray(rng.uniform(14, 120, shp)), reliq=np.asfortranarray(rng.uniform(3, 50, shp)))


@pytest.mark.parametrize("iceflg", [1, 2, 3])
def test_cloud_optics_from_water_paths(gpu, oracle, cols, iceflg):
    """cldprop_sw's inflag = 2 (rrtmg_sw_cldprop.f90:168-345): Ebert-Curry / Streamer / Fu ice, Hu-Stamnes liquid, and their
    delta-scaled combination; also with aerosols and in chunks."""
    rng = np.random.default_rng(60 + iceflg)
    cl = _water_clouds(cols, rng)
    got = gpu.sw_from_columns(cols, icld=2, inflgsw=2, iceflgsw=iceflg, liqflgsw=1, clouds=cl)
    _check_outputs(got, oracle.rrtmg_sw(cols, icld=2, inflgsw=2, iceflgsw=iceflg, liqflgsw=1, clouds=cl), SW_OUT)
    clear = gpu.sw_from_columns(cols)
    assert np.max(np.abs(got[4] - clear[4])) < 1e-9 * np.abs(clear[4]).max()
    c = cols.take(np.arange(170))
    cl = {k: np.asfortranarray(v[:170]) for k, v in cl.items()}
    aer = dict(ecaer=np.asfortranarray(rng.uniform(0.0, 0.05, (c.ncol, c.nlay, 6))))
    ref = oracle.rrtmg_sw(c, icld=2, iaer=6, inflgsw=2, iceflgsw=iceflg, liqflgsw=1, clouds=cl, aerosols=aer)
    gpu.set_option("host_chunk", 64)
    try:
        _check_outputs(gpu.sw_from_columns(c, icld=2, iaer=6, inflgsw=2, iceflgsw=iceflg, liqflgsw=1, clouds=cl, aerosols=aer), ref, SW_OUT)
    finally:
        gpu.set_option("host_chunk", 0)


def test_radius_out_of_range_is_an_error(gpu, cols):
    c = cols.take(np.flatnonzero(cols.coszen > 0.1)[:30].tolist() + np.flatnonzero(cols.coszen == 0)[:10].tolist())
    cl = _water_clouds(c, np.random.default_rng(2))
    cl["cldfr"][35, 5] = 1.0; cl["cicewp"][35, 5] = 10.0; cl["reice"][35, 5] = 4.0     # a night column: never looked at (rad.nomcica:497-505)
    gpu.sw_from_columns(c, icld=2, inflgsw=2, iceflgsw=2, liqflgsw=1, clouds=cl)
    cl["cldfr"][3, 5] = 1.0; cl["cicewp"][3, 5] = 10.0; cl["reice"][3, 5] = 4.0
    for ice in (1, 2, 3):
        with pytest.raises(gpu.RRTMGError) as e:
            gpu.sw_from_columns(c, icld=2, inflgsw=2, iceflgsw=ice, liqflgsw=1, clouds=cl)
        assert e.value.code == 7 and "ICE" in str(e.value)
    cl["reice"][3, 5] = 50.0
    cl["cliqwp"][3, 5] = 5.0; cl["reliq"][3, 5] = 2.0
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.sw_from_columns(c, icld=2, inflgsw=2, iceflgsw=2, liqflgsw=1, clouds=cl)
    assert e.value.code == 7 and "LIQUID" in str(e.value)
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.sw_from_columns(c, icld=2, inflgsw=2, iceflgsw=0, liqflgsw=1, clouds=cl)      # cldprop_sw defines ice options 1..3 only
    assert e.value.code == 4
    gpu.sw_from_columns(c)
