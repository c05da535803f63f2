"""GPU parity at the edges of the interface: layer counts from 1 to the 128-layer maximum (every template instantiation
and the ragged tails of the 4-layer TMA chunks and 8-level reduction batches), column counts around the tile and pair
sizes, and inputs pushed to the clamps the reference applies (temperatures outside the Planck table, a dry and a
saturated atmosphere, albedo 0 and 1, the sun on the horizon).  All against the oracle, tolerances of north_star."""
import numpy as np
import pytest

from mima_b200.columns import make_columns
from test_gpu_parity import LW_OUT, SW_OUT, _check_outputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nlay", [1, 2, 3, 5, 7, 8, 9, 33, 63, 64, 65, 127, 128])
def test_layer_counts(gpu, oracle, nlay):
    c = make_columns("T42L40", nlon=37, nlat=1, nlay=nlay, night=True, seed=100 + nlay)
    assert c.nlay == nlay and c.ncol == 37
    _check_outputs(gpu.lw_from_columns(c), oracle.rrtmg_lw(c), LW_OUT, hr_tight=1e-6)
    _check_outputs(gpu.sw_from_columns(c), oracle.rrtmg_sw(c), SW_OUT, hr_tight=1e-6)
    ref = oracle.rrtmg_lw(c, idrv=1)
    _check_outputs(gpu.lw_from_columns(c, idrv=1), ref, LW_OUT + ("duflx_dt", "duflxc_dt"), hr_tight=1e-6)


@pytest.mark.parametrize("seed,nlay", [(1, 40), (2, 60)])
def test_wild_columns(gpu, oracle, seed, nlay):
    """Columns far outside the bench climate (mima_b200.columns.wild_columns: every gas and CFC present and varied over orders
    of magnitude, mountains and deep lows, +-35 K, grey surfaces, earth-sun adjustment) at the north-star tolerances."""
    from mima_b200.columns import wild_columns
    c = wild_columns(seed, nlay, nlon=64, nlat=4)
    _check_outputs(gpu.lw_from_columns(c), oracle.rrtmg_lw(c), LW_OUT, tight=False)
    _check_outputs(gpu.sw_from_columns(c), oracle.rrtmg_sw(c), SW_OUT, tight=False)


@pytest.mark.parametrize("ncol", [1, 2, 31, 32, 33, 63, 65, 255, 257])
def test_column_counts(gpu, oracle, ncol):
    base = make_columns("T170L60", nlon=257, nlat=1, night=True, seed=7)
    c = base.take(np.arange(ncol))
    _check_outputs(gpu.lw_from_columns(c), oracle.rrtmg_lw(c), LW_OUT)
    _check_outputs(gpu.sw_from_columns(c), oracle.rrtmg_sw(c), SW_OUT)


def test_inputs_at_the_clamps(gpu, oracle):
    c = make_columns("T42L40", nlon=64, nlat=2, seed=11)
    n = c.ncol
    rng = np.random.default_rng(3)
    # temperatures beyond both ends of the Planck table (159..339 K, setcoef.f90:154-160) and of tref +- 30 K (jt clamps)
    c.tlay[: n // 4] = 100.0 + 10.0 * rng.uniform(size=(n // 4, c.nlay))
    c.tlev[: n // 4] = 100.0 + 10.0 * rng.uniform(size=(n // 4, c.nlay + 1))
    c.tsfc[: n // 4] = 105.0
    c.tlay[n // 4: n // 2] = 355.0 + 15.0 * rng.uniform(size=(n // 2 - n // 4, c.nlay))
    c.tlev[n // 4: n // 2] = 355.0 + 15.0 * rng.uniform(size=(n // 2 - n // 4, c.nlay + 1))
    c.tsfc[n // 4: n // 2] = 370.0
    # bone dry and soaking wet columns, no ozone, 20 x CO2 (adjcol branches of LW bands 3, 6, 7, 8, 9, 13)
    c.h2o[n // 2: n // 2 + 8] = 2.0e-9
    c.h2o[n // 2 + 8: n // 2 + 16] = 0.04
    c.o3[n // 2 + 16: n // 2 + 24] = 0.0
    c.co2[n // 2 + 24: n // 2 + 32] = 20 * 390e-6
    # albedo 0 and 1, sun on the horizon / just below the night threshold / overhead
    c.albedo[-8:-4] = 0.0
    c.albedo[-4:] = 1.0
    c.coszen[-16:-12] = 1.0e-10
    c.coszen[-12:-10] = 0.99e-10
    c.coszen[-10:-8] = 1.0
    for a in (c.tlay, c.tlev, c.h2o, c.o3, c.co2):
        assert a.flags["F_CONTIGUOUS"]
    _check_outputs(gpu.lw_from_columns(c), oracle.rrtmg_lw(c), LW_OUT, hr_tight=1e-5)
    _check_outputs(gpu.sw_from_columns(c), oracle.rrtmg_sw(c), SW_OUT, hr_tight=1e-5)


@pytest.mark.parametrize("nlay", [1, 2, 9, 65, 128])
def test_layer_counts_general_branches(gpu, oracle, nlay):
    """The cloudy / aerosol kernels at the same layer counts (their 4-level reduction batches and the overlap pre-pass)."""
    from test_oracle_lw_clouds import cloud_field
    c = make_columns("T42L40", nlon=41, nlat=1, nlay=nlay, night=True, seed=300 + nlay)
    rng = np.random.default_rng(nlay)
    if nlay >= 9:
        lw_cl = cloud_field(c, rng)
    else:
        cf = np.asfortranarray(rng.uniform(0.0, 1.0, (c.ncol, nlay)) * (rng.uniform(size=(c.ncol, nlay)) < 0.6))
        lw_cl = dict(cldfr=cf, taucld=np.asfortranarray(rng.uniform(0.0, 8.0, (16, c.ncol, nlay)) * (cf > 0)[None]))
    for icld in (1, 2):
        _check_outputs(gpu.lw_from_columns(c, icld=icld, clouds=lw_cl, idrv=1), oracle.rrtmg_lw(c, icld=icld, clouds=lw_cl, idrv=1),
                       LW_OUT + ("duflx_dt", "duflxc_dt"), hr_tight=1e-5)
    shp = (14, c.ncol, nlay)
    cld = (rng.uniform(size=(c.ncol, nlay)) < 0.4).astype(np.float64)
    asm = rng.uniform(0.7, 0.9, shp)
    sw_cl = dict(cldfr=np.asfortranarray(cld), taucld=np.asfortranarray(rng.uniform(0.0, 20.0, shp) * cld[None]),
                 ssacld=np.asfortranarray(rng.uniform(0.5, 0.99999, shp)), asmcld=np.asfortranarray(asm),
                 fsfcld=np.asfortranarray(asm * asm))
    aer = dict(ecaer=np.asfortranarray(rng.uniform(0.0, 0.05, (c.ncol, nlay, 6))))
    _check_outputs(gpu.sw_from_columns(c, icld=2, iaer=6, clouds=sw_cl, aerosols=aer),
                   oracle.rrtmg_sw(c, icld=2, iaer=6, clouds=sw_cl, aerosols=aer), SW_OUT, hr_tight=1e-5)
