"""The reference's LW known-answer cases (doc_rrtm/runs_std_atm: AER standard atmospheres MLS/MLW/SAW/TROP, clear sky).

tests/golden/std_atm.npz is generated from the reference files by tests/golden/make_std_atm.py (committed).  The
fixture itself is checked here on every run; the flux comparison needs the real LW k-distribution
(mima_b200/data/rrtmg_lw_kg.bin, built by tools/build_tables.py from rrtmg_lw_k_g.f90, which is stripped from the
reference checkout) and is skipped until that blob exists -- with the packaged synthetic tables the LW
*coefficients* stay unpinned.

How a TAPE5 case enters the MiMA-flavoured rrtmg_lw interface (LW rad.nomcica:770-784): H2O volume mixing ratio ->
specific humidity q = w/(amdw + w) (inverse of :775), O3 vmr -> mass mixing ratio vmr/amdo (inverse of :778), other
gases as vmr; level pressures and temperatures as given.  inatm recomputes the dry-air column from the level
pressures instead of taking the file's broadening-gas column, so agreement is at the 1e-2 level, not the printed
precision: tolerance 1 % of the largest flux (the fixture check below shows the per-layer column amounts agree to a few per
cent where the printed level pressures are coarse, the column totals to 4e-3)."""
import dataclasses
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "std_atm.npz")
REAL_LW = os.path.join(ROOT, "mima_b200", "data", "rrtmg_lw_kg.bin")
CASES = ("MLS", "MLW", "SAW", "TROP")
AMDW, AMDO = 1.607793, 0.603428
SIGMA = 5.67e-8 * 1.0   # W m-2 K-4 (value used for the plausibility check only)


def _case(name):
    z = np.load(FIX)
    return {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", CASES)
def test_fixture_is_wellformed(name):
    c = _case(name)
    L = c["pavel"].size
    assert c["pz"].size == L + 1 and c["tz"].size == L + 1 and c["vmr"].shape == (7, L) and c["uflx"].size == L + 1
    assert (np.diff(c["pz"]) < 0).all() and (np.diff(c["pavel"]) < 0).all()
    assert ((c["pavel"] < c["pz"][:-1]) & (c["pavel"] > c["pz"][1:])).all()
    np.testing.assert_allclose(c["plev"], c["pz"], rtol=2e-3)                 # output levels = input levels (printed 4 digits)
    # black surface: upward flux at the ground = sigma T^4 of the boundary temperature; nothing enters at the top
    np.testing.assert_allclose(c["uflx"][0], SIGMA * c["tbound"] ** 4, rtol=2e-3)
    assert c["dflx"][-1] == 0.0
    np.testing.assert_allclose(c["fnet"], c["uflx"] - c["dflx"], atol=2e-4)
    # heating rate = flux divergence (LW rtrnmr.f90:751-777), printed to 5 decimals
    heatfac = 9.8066 * 86400.0 / (1004.64 * 100.0)
    hr = heatfac * (c["fnet"][:-1] - c["fnet"][1:]) / (c["pz"][:-1] - c["pz"][1:])
    np.testing.assert_allclose(c["hr"][:-1], hr, rtol=5e-3, atol=5e-3)
    # the hydrostatic dry-air column inatm derives from the level pressures matches the file's column amounts
    amd, amw, grav, avogad = 28.9660, 18.0160, 9.8066, 6.02214199e+23
    w = c["vmr"][0]
    amm = (1.0 - w) * amd + w * amw
    coldry = (c["pz"][:-1] - c["pz"][1:]) * 1.e3 * avogad / (1.e2 * grav * amm * (1.0 + w))
    np.testing.assert_allclose(coldry, c["coldry"], rtol=8e-2)       # level pressures are printed to 4-5 digits
    np.testing.assert_allclose(coldry.sum(), c["coldry"].sum(), rtol=4e-3)


def _columns(c):
    from mima_b200.columns import Columns
    L = c["pavel"].size
    row = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float64)[None, :])
    w = c["vmr"]
    zeros = np.zeros((1, L), order="F")
    return Columns(ncol=1, nlay=L, nlon=1, nlat=1, play=row(c["pavel"]), plev=row(c["pz"]), tlay=row(c["tavel"]),
                   tlev=row(c["tz"]), tsfc=np.array([float(c["tbound"])]), h2o=row(w[0] / (AMDW + w[0])), o3=row(w[2] / AMDO),
                   co2=row(w[1]), ch4=row(w[5]), n2o=row(w[3]), o2=row(w[6]), cfc11=zeros, cfc12=zeros, cfc22=zeros,
                   ccl4=zeros, emis=np.ones((1, 16), order="F"), albedo=np.array([0.2]), coszen=np.array([0.5]))


def test_columns_from_a_tape5_case_run_through_the_oracle(oracle):
    """The conversion into the rrtmg_lw interface is exercised on every run (synthetic tables: only finiteness and the
    table-independent surface emission can be checked)."""
    for name in CASES:
        c = _case(name)
        out = oracle.rrtmg_lw(_columns(c))
        assert np.isfinite(out["uflx"]).all() and np.isfinite(out["hr"]).all()
        np.testing.assert_allclose(out["uflx"][0, 0], c["uflx"][0], rtol=2e-3)     # Planck table x emissivity 1


NEEDS_REAL = pytest.mark.xfail(not os.path.exists(REAL_LW), strict=True,
                               reason="LW coefficients unpinned: the run uses the packaged SYNTHETIC k-distribution because "
                               "rrtmg_lw_k_g.f90 is stripped from the reference checkout; build mima_b200/data/rrtmg_lw_kg.bin "
                               "with tools/build_tables.py (MIMA_LW_KG=...) and these known answers must pass")


@NEEDS_REAL
@pytest.mark.parametrize("name", CASES)
def test_known_answers_oracle(name):
    """Expected to FAIL (xfail, strict) while only the synthetic tables exist -- reported on every run instead of hidden
    behind a skip."""
    from oracle.pyoracle import Oracle
    c = _case(name)
    out = Oracle(lw_kg=REAL_LW if os.path.exists(REAL_LW) else None).rrtmg_lw(_columns(c))
    scale = c["uflx"].max()
    assert np.max(np.abs(out["uflx"][0] - c["uflx"])) < 1e-2 * scale
    assert np.max(np.abs(out["dflx"][0] - c["dflx"])) < 1e-2 * scale
    tropo = c["pavel"] > 100.0
    assert np.max(np.abs(out["hr"][0][tropo] - c["hr"][:-1][tropo])) < 0.05      # K/day


@pytest.mark.gpu
@NEEDS_REAL
@pytest.mark.parametrize("name", CASES)
def test_known_answers_gpu(gpu, name):
    c = _case(name)
    uflx, dflx, hr = gpu.lw_from_columns(_columns(c))[:3]
    scale = c["uflx"].max()
    assert np.max(np.abs(uflx[0] - c["uflx"])) < 1e-2 * scale
    assert np.max(np.abs(dflx[0] - c["dflx"])) < 1e-2 * scale
