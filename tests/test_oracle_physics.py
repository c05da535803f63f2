"""Oracle behaviour: invariants of the reference algorithm and regression
vectors.  CPU only."""
import os

import numpy as np
import pytest

from mima_b200.columns import make_columns

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cols():
    return make_columns("T42L40", nlon=16, nlat=16, night=True)


def test_sw_night_columns_are_zero(oracle, cols):
    out = oracle.rrtmg_sw(cols)
    night = cols.coszen < 1e-10
    assert night.any() and (~night).any()
    for k in ("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc"):
        assert (out[k][night] == 0).all(), k
    assert (out["swdflx"][~night, -1] > 0).all()


def test_clear_equals_total_when_icld0(oracle, cols):
    sw = oracle.rrtmg_sw(cols)
    lw = oracle.rrtmg_lw(cols)
    assert np.array_equal(sw["swuflx"], sw["swuflxc"]) and np.array_equal(sw["swdflx"], sw["swdflxc"])
    assert np.array_equal(sw["swhr"], sw["swhrc"])
    assert np.array_equal(lw["uflx"], lw["uflxc"]) and np.array_equal(lw["dflx"], lw["dflxc"])
    assert np.array_equal(lw["hr"], lw["hrc"])


def test_heating_rate_is_flux_divergence(oracle, cols):
    """hr = heatfac * d(net flux)/dp (rtrnmr.f90:770-771, rad.nomcica:719-723); SW top layer forced to 0
    (MiMA edit, SW rad.nomcica:724-726), LW top layer computed."""
    heatfac = 9.8066 * 86400.0 / (1004.64 * 100.0)
    lw = oracle.rrtmg_lw(cols)
    fnet = lw["uflx"] - lw["dflx"]
    hr = heatfac * (fnet[:, :-1] - fnet[:, 1:]) / (cols.plev[:, :-1] - cols.plev[:, 1:])
    assert np.allclose(lw["hr"], hr, rtol=1e-12, atol=1e-12)
    assert (lw["hr"][:, -1] != 0).all()
    sw = oracle.rrtmg_sw(cols)
    day = cols.coszen >= 1e-10
    net = sw["swdflx"] - sw["swuflx"]
    hr = (net[:, 1:] - net[:, :-1]) * heatfac / (cols.plev[:, :-1] - cols.plev[:, 1:])
    assert np.allclose(sw["swhr"][day, :-1], hr[day, :-1], rtol=1e-12, atol=1e-12)
    assert (sw["swhr"][:, -1] == 0).all()


def test_sw_energy_bounds(oracle, cols):
    sw = oracle.rrtmg_sw(cols)
    day = cols.coszen >= 1e-10
    toa = sw["swdflx"][day, -1]
    assert np.allclose(toa, cols.scon * cols.coszen[day], rtol=2e-5)          # incoming = S0 * cos(zenith)
    assert (sw["swuflx"][day, -1] < toa).all() and (sw["swuflx"][day] >= 0).all()
    assert (sw["swdflx"][day, 0] < toa).all()
    # surface reflection: up = albedo * down at the surface (Lambertian, same albedo for direct and diffuse)
    assert np.allclose(sw["swuflx"][day, 0], cols.albedo[day] * sw["swdflx"][day, 0], rtol=1e-10)
    assert (sw["swhr"][day, :-1] >= -1e-9).all()                               # the sun only heats


def test_lw_surface_emission(oracle, cols):
    """With emissivity 1 the upward surface flux is the band-integrated Planck flux at tsfc."""
    lw = oracle.rrtmg_lw(cols)
    sb = 5.6704e-8 * cols.tsfc ** 4
    assert np.all(lw["uflx"][:, 0] < 1.0001 * sb) and np.all(lw["uflx"][:, 0] > 0.98 * sb)
    assert (lw["dflx"][:, -1] == 0).all()                                      # nothing enters at the top
    assert (lw["dflx"][:, 0] > 0).all()


def test_column_permutation_and_threads(oracle, cols):
    """No cross-column coupling: any permutation / thread count gives bit-identical columns."""
    perm = np.random.default_rng(1).permutation(cols.ncol)
    a = oracle.rrtmg_lw(cols, nthreads=1)
    b = oracle.rrtmg_lw(cols.take(perm), nthreads=4)
    assert np.array_equal(a["uflx"][perm], b["uflx"]) and np.array_equal(a["hr"][perm], b["hr"])
    a = oracle.rrtmg_sw(cols, nthreads=1)
    b = oracle.rrtmg_sw(cols.take(perm), nthreads=3)
    assert np.array_equal(a["swdflx"][perm], b["swdflx"]) and np.array_equal(a["swhr"][perm], b["swhr"])


def test_branch_coverage_of_generated_columns(oracle):
    """The synthetic batches must exercise the branches the reference has: both eta 3-point stencils, the
    thin/thick optical-depth branches, the high-CO2 adjfac branch (4xCO2), both SW reftra branches."""
    c = make_columns("T42L40", nlon=16, nlat=16)
    st = oracle.rrtmg_lw(c, stages=True)["stages"]
    assert st["laytrop"].min() >= 10 and st["laytrop"].max() < c.nlay
    od = 1.66 * st["taug"]
    assert 0.1 < (od <= 0.06).mean() < 0.9
    lay = np.arange(1, c.nlay + 1)[None, :]
    low = lay <= st["laytrop"][:, None]
    # default MiMA: ch4 = n2o = 0 -> eta clamps to oneminus (bands 9,13,16) and ~0 (band 15)
    eta = st["colh2o"] / (st["colh2o"] + 1.0 * st["colch4"])
    assert (eta[low] > 0.875).all()
    c4 = make_columns("T42L40", nlon=16, nlat=16, co2_ppmv=1560.0, ozone="file", secondary_gases=True)
    s4 = oracle.rrtmg_lw(c4, stages=True)["stages"]
    ratco2 = (s4["colco2"] / s4["coldry"]) * 1e20 / 3.55e-4
    assert (ratco2 > 3.0).all()
    sw = oracle.rrtmg_sw(c, stages=True)["stages"]
    w = sw["taur"] / (sw["taur"] + sw["taug"])
    assert (w >= 0.9999995).any() and (w < 0.9999995).any()


def test_regression_vectors(oracle):
    """Pinned oracle outputs (tests/golden/oracle_t42l40.npz, written by tests/golden/make_golden.py).  These are
    self-generated regression vectors -- NOT reference outputs (the reference cannot be run here)."""
    path = os.path.join(GOLD, "oracle_t42l40.npz")
    g = np.load(path)
    c = make_columns("T42L40", nlon=8, nlat=8, night=True)
    lw, sw = oracle.rrtmg_lw(c), oracle.rrtmg_sw(c)
    for k in ("uflx", "dflx", "hr"):
        assert np.allclose(lw[k], g["lw_" + k], rtol=1e-12, atol=1e-12), k
    for k in ("swuflx", "swdflx", "swhr"):
        assert np.allclose(sw[k], g["sw_" + k], rtol=1e-12, atol=1e-12), k


def test_reference_vectors(oracle):
    """tests/golden/ref_t42l40.npz holds outputs of the reference's own RRTMG code (machine-translated F90 -> C, see
    tests/golden/make_ref_vectors.py) for config C4 columns: the oracle must reproduce them bit for bit.  Runs anywhere
    (no oracle/_ref needed)."""
    from golden.make_ref_vectors import batch
    g = np.load(os.path.join(GOLD, "ref_t42l40.npz"))
    c = batch()
    lw, sw = oracle.rrtmg_lw(c, idrv=1), oracle.rrtmg_sw(c)
    for k in ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc", "duflx_dt", "duflxc_dt"):
        assert np.array_equal(lw[k], g[k]), k
    for k in ("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc"):
        assert np.array_equal(sw[k], g[k]), k
