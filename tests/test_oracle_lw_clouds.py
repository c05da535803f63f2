"""CPU checks of the oracle's cloudy longwave branches (rtrn random overlap, rtrnmr maximum/random overlap, cloud
optical depth given per band).  The reference's cloudy known answers (doc_rrtm/runs_std_atm) need the stripped LW
k-tables, so the restatement is anchored on its reduction to the pinned clear path and on the limits in which the two
overlap rules must agree."""
import numpy as np
import pytest

from mima_b200.columns import make_columns

LW = ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")


def cloud_field(cols, rng, p_block=0.5):
    """Per column up to three cloud decks of 1-6 contiguous layers with fractions that rise and fall inside a deck
    (every branch of the maximum/random factors), some columns cloud free, the top two layers always clear."""
    ncol, nlay = cols.ncol, cols.nlay
    cf = np.zeros((ncol, nlay), order="F")
    for i in range(ncol):
        if rng.uniform() < 0.25:
            continue
        for _ in range(rng.integers(1, 4)):
            if rng.uniform() > p_block:
                continue
            lo = int(rng.integers(0, nlay - 8))
            n = int(rng.integers(1, 7))
            f = rng.uniform(0.05, 1.0, n)
            if rng.uniform() < 0.2:
                f[rng.integers(0, n)] = 1.0
            if n > 2 and rng.uniform() < 0.3:
                f[1] = f[0]                                   # equal neighbours
            cf[i, lo:lo + n] = f
    tau = rng.uniform(0.0, 8.0, (16, ncol, nlay)) * (cf > 0)[None]
    return dict(cldfr=np.asfortranarray(cf), taucld=np.asfortranarray(tau))


@pytest.fixture(scope="module")
def cols():
    return make_columns("T42L40", nlon=16, nlat=4)


def test_zero_cloud_is_the_clear_path_bitwise(oracle, cols):
    ref = oracle.rrtmg_lw(cols, idrv=1)
    z = np.zeros((cols.ncol, cols.nlay), order="F")
    t = np.full((16, cols.ncol, cols.nlay), 3.0, order="F")
    for icld in (1, 2, 3):
        got = oracle.rrtmg_lw(cols, icld=icld, clouds=dict(cldfr=z, taucld=t), idrv=1)
        for k in ref:
            assert np.array_equal(got[k], ref[k]), (icld, k)


def test_clear_stream_and_cloud_free_columns(oracle, cols):
    cl = cloud_field(cols, np.random.default_rng(1))
    ref = oracle.rrtmg_lw(cols)
    cloudy = cl["cldfr"].max(axis=1) > 0
    assert cloudy.any() and (~cloudy).any()
    for icld in (1, 2):
        got = oracle.rrtmg_lw(cols, icld=icld, clouds=cl)
        for k in ("uflxc", "dflxc", "hrc"):
            assert np.array_equal(got[k], ref[k]), k
        for k in ("uflx", "dflx", "hr"):
            assert np.array_equal(got[k][~cloudy], ref[k][~cloudy]), k
            assert np.isfinite(got[k]).all()
        assert np.abs(got["dflx"][cloudy, 0] - ref["dflx"][cloudy, 0]).max() > 1.0


def test_single_cloud_layer_overlap_rules_agree(oracle, cols):
    """With one cloudy layer per column there is nothing to overlap: random and maximum/random differ only in how the
    cloud transmittance is evaluated (exp() against the table), i.e. by the table's interpolation error."""
    rng = np.random.default_rng(2)
    cf = np.zeros((cols.ncol, cols.nlay), order="F")
    cf[np.arange(cols.ncol), rng.integers(2, cols.nlay - 4, cols.ncol)] = rng.uniform(0.1, 1.0, cols.ncol)
    tau = np.asfortranarray(rng.uniform(0.0, 8.0, (16, cols.ncol, cols.nlay)) * (cf > 0)[None])
    a = oracle.rrtmg_lw(cols, icld=1, clouds=dict(cldfr=cf, taucld=tau))
    b = oracle.rrtmg_lw(cols, icld=2, clouds=dict(cldfr=cf, taucld=tau))
    for k in ("uflx", "dflx"):
        assert np.max(np.abs(a[k] - b[k])) < 5e-2, k
        assert np.max(np.abs(a[k] - b[k])) > 0.0


def test_maximum_overlap_shields_less_than_random(oracle, cols):
    """Two adjacent half-covered opaque layers: maximally overlapped they cover half the sky, randomly overlapped three
    quarters, so the random rule lets less surface emission through in the window bands."""
    cf = np.zeros((cols.ncol, cols.nlay), order="F")
    cf[:, 5:7] = 0.5
    tau = np.asfortranarray(np.full((16, cols.ncol, cols.nlay), 50.0) * (cf > 0)[None])
    rnd = oracle.rrtmg_lw(cols, icld=1, clouds=dict(cldfr=cf, taucld=tau))
    mro = oracle.rrtmg_lw(cols, icld=2, clouds=dict(cldfr=cf, taucld=tau))
    clr = oracle.rrtmg_lw(cols)
    # downward flux at the surface: more cloud cover seen from below under random overlap
    assert (rnd["dflx"][:, 0] >= mro["dflx"][:, 0] - 1e-9).all()
    assert (mro["dflx"][:, 0] >= clr["dflx"][:, 0] - 1e-9).all()
    assert (rnd["dflx"][:, 0] - mro["dflx"][:, 0]).max() > 0.1


def test_switch_handling(oracle, cols):
    c = cols.take(np.arange(6))
    cl = cloud_field(c, np.random.default_rng(3))
    a = oracle.rrtmg_lw(c, icld=8, clouds=cl)           # reset to 2 (rad.nomcica:437)
    b = oracle.rrtmg_lw(c, icld=2, clouds=cl)
    d = oracle.rrtmg_lw(c, icld=3, clouds=cl)           # 2 and 3 both go through rtrnmr (:527-541)
    for k in LW:
        assert np.array_equal(a[k], b[k]) and np.array_equal(d[k], b[k])
    with pytest.raises(RuntimeError, match="rc=3"):
        oracle.rrtmg_lw(c, icld=2)


def test_cldprop_parameterisations(oracle, cols):
    """inflglw = 1 is inflglw = 0 fed with abscld1 * (ice + liquid water path) in every band (cldprop.f90:176-180); an
    iceflglw = 0, liqflglw = 0 cloud is grey as well (one coefficient for all bands, icb(:,0) = 1); the Fortran stops come
    back as return codes."""
    rng = np.random.default_rng(5)
    shp = (cols.ncol, cols.nlay)
    cf = np.asfortranarray((rng.uniform(size=shp) < 0.3) * rng.uniform(0.1, 1.0, shp))
    iwp = np.asfortranarray(rng.uniform(0, 30, shp))
    lwp = np.asfortranarray(rng.uniform(0, 60, shp))
    cl = dict(cldfr=cf, cicewp=iwp, cliqwp=lwp, reice=np.full(shp, 40.0, order="F"), reliq=np.full(shp, 10.0, order="F"))
    a = oracle.rrtmg_lw(cols, icld=2, clouds=cl, inflglw=1)
    tau = np.asfortranarray(np.broadcast_to((0.0602410 * (iwp + lwp))[None], (16,) + shp))
    b = oracle.rrtmg_lw(cols, icld=2, clouds=dict(cldfr=cf, taucld=tau))
    for k in LW:
        assert np.array_equal(a[k], b[k]), k
    g = oracle.rrtmg_lw(cols, icld=2, clouds=cl, inflglw=2, iceflglw=0, liqflglw=0)
    tau = np.asfortranarray(np.broadcast_to((iwp * (0.005 + 1.0 / 40.0) + lwp * 0.0903614)[None], (16,) + shp))
    h = oracle.rrtmg_lw(cols, icld=2, clouds=dict(cldfr=cf, taucld=tau))
    for k in ("uflx", "dflx"):
        # same optical depths, but ncbands = 1: every band reads secdiff(1) * taucloud(:,1) instead of its own secant
        assert np.max(np.abs(g[k] - h[k])) < 3.0 and np.max(np.abs(g[k] - h[k])) > 0.0, k
    cl["cldfr"][2, 3] = 0.5; cl["cldfr"][1, 1] = 0.5            # make sure the two cells below are cloudy
    cl["reice"][2, 3] = 4.0
    for ice, rc in ((0, 11), (1, 12), (2, 12), (3, 13)):
        with pytest.raises(RuntimeError, match=f"rc={rc}"):
            oracle.rrtmg_lw(cols, icld=2, clouds=cl, inflglw=2, iceflglw=ice, liqflglw=1)
    cl["reice"][2, 3] = 40.0
    cl["reliq"][1, 1] = 61.0
    with pytest.raises(RuntimeError, match="rc=14"):
        oracle.rrtmg_lw(cols, icld=2, clouds=cl, inflglw=2, iceflglw=2, liqflglw=1)
