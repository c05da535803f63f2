"""CPU checks of the oracle's general shortwave path (aerosols iaer = 10, clouds icld >= 1 with inflgsw = 0).
The reference holds no vectors for these branches (MiMA never takes them), so the restatement is anchored on its
reduction to the pinned clear-sky path and on the physics the two-stream equations guarantee."""
import numpy as np
import pytest

from mima_b200.columns import make_columns

SW = ("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc")


@pytest.fixture(scope="module")
def cols():
    return make_columns("T42L40", nlon=16, nlat=4, night=True)


def _clouds(cols, rng, frac=0.3):
    shp = (14, cols.ncol, cols.nlay)
    cld = (rng.uniform(size=(cols.ncol, cols.nlay)) < frac).astype(np.float64)
    asm = rng.uniform(0.7, 0.9, shp)
    return dict(cldfr=np.asfortranarray(cld), taucld=np.asfortranarray(rng.uniform(0.0, 20.0, shp) * cld[None]),
                ssacld=np.asfortranarray(rng.uniform(0.5, 0.99999, shp)), asmcld=np.asfortranarray(asm),
                fsfcld=np.asfortranarray(asm * asm))


def test_zero_optical_depth_is_the_clear_path_bitwise(oracle, cols):
    ref = oracle.rrtmg_sw(cols)
    z = np.zeros((cols.ncol, cols.nlay, 14), order="F")
    a = oracle.rrtmg_sw(cols, iaer=10, aerosols=dict(tauaer=z, ssaaer=z + 0.9, asmaer=z + 0.6))
    zc = np.zeros((14, cols.ncol, cols.nlay), order="F")
    c = oracle.rrtmg_sw(cols, icld=2, clouds=dict(cldfr=np.zeros((cols.ncol, cols.nlay), order="F"), taucld=zc + 5.0,
                                                   ssacld=zc + 0.9, asmcld=zc + 0.8, fsfcld=zc + 0.64))
    for k in SW:
        assert np.array_equal(a[k], ref[k]), k
        assert np.array_equal(c[k], ref[k]), k


def test_clouds_leave_the_clear_stream_alone_and_dim_the_surface(oracle, cols):
    cl = _clouds(cols, np.random.default_rng(1))
    ref = oracle.rrtmg_sw(cols)
    got = oracle.rrtmg_sw(cols, icld=2, clouds=cl)
    for k in ("swuflxc", "swdflxc", "swhrc"):
        assert np.array_equal(got[k], ref[k]), k
    day = cols.coszen > 0.1
    cloudy = day & (cl["taucld"].sum(axis=(0, 2)) > 1.0)
    assert cloudy.any()
    assert (got["swdflx"][cloudy, 0] < got["swdflxc"][cloudy, 0]).all()
    # energy: nothing exceeds the incoming flux, absorbed + reflected = incoming at the top
    top = got["swdflx"][:, -1]
    assert (got["swuflx"] <= top[:, None] * (1 + 1e-12)).all()
    assert (got["swdflx"] <= top[:, None] * (1 + 1e-12)).all()


def test_absorbing_aerosol_heats_scattering_aerosol_reflects(oracle, cols):
    shp = (cols.ncol, cols.nlay, 14)
    tau = np.full(shp, 0.02, order="F")
    ref = oracle.rrtmg_sw(cols)
    absorbing = oracle.rrtmg_sw(cols, iaer=10, aerosols=dict(tauaer=tau, ssaaer=np.full(shp, 0.3, order="F"),
                                                             asmaer=np.full(shp, 0.6, order="F")))
    scattering = oracle.rrtmg_sw(cols, iaer=10, aerosols=dict(tauaer=tau, ssaaer=np.full(shp, 0.999999, order="F"),
                                                              asmaer=np.full(shp, 0.6, order="F")))
    day = cols.coszen > 0.2
    net = lambda o: o["swdflx"] - o["swuflx"]
    assert (net(absorbing)[day, -1] - net(absorbing)[day, 0] > net(ref)[day, -1] - net(ref)[day, 0]).all()
    assert (scattering["swuflx"][day, -1] > ref["swuflx"][day, -1]).all()


def test_switch_handling(oracle, cols):
    c = cols.take(np.arange(8))
    cl = _clouds(c, np.random.default_rng(2))
    a = oracle.rrtmg_sw(c, icld=9, iaer=4, clouds=cl)           # reset to icld = 2, iaer = 0 (rad.nomcica:468-473)
    b = oracle.rrtmg_sw(c, icld=2, clouds=cl)
    for k in SW:
        assert np.array_equal(a[k], b[k])
    night, day = int(np.flatnonzero(c.coszen == 0)[0]), int(np.flatnonzero(c.coszen > 0)[0])
    cl["cldfr"][night, 3] = 0.4                                  # a night column is skipped before the test (:497-505)
    oracle.rrtmg_sw(c, icld=2, clouds=cl)
    cl["cldfr"][day, 3] = 0.4
    with pytest.raises(RuntimeError, match="rc=4"):             # stop 'PARTIAL CLOUD NOT ALLOWED' (:537)
        oracle.rrtmg_sw(c, icld=2, clouds=cl)
    with pytest.raises(RuntimeError, match="rc=3"):
        oracle.rrtmg_sw(c, iaer=6)                               # needs ecaer


def test_ecmwf_aerosol_types(oracle, cols):
    """iaer = 6 (rad.nomcica:608-640): zero amounts are the clear path bit for bit; one type alone reproduces iaer = 10
    fed with that type's band properties from the swaerpr tables."""
    ref = oracle.rrtmg_sw(cols)
    z = np.zeros((cols.ncol, cols.nlay, 6), order="F")
    got = oracle.rrtmg_sw(cols, iaer=6, aerosols=dict(ecaer=z))
    for k in SW:
        assert np.array_equal(got[k], ref[k]), k
    tau = oracle.table("swaer.rsrtaua").reshape(14, 6, order="F")
    piz = oracle.table("swaer.rsrpiza").reshape(14, 6, order="F")
    asy = oracle.table("swaer.rsrasya").reshape(14, 6, order="F")
    assert tau[9, 0] == 1.69446 and piz[0, 5] == .2355667 and asy[13, 1] == 0.818871      # swaerpr literals
    e = np.zeros((cols.ncol, cols.nlay, 6), order="F")
    e[:, :12, 2] = 0.03                                          # type 3 in the lowest 12 layers
    six = oracle.rrtmg_sw(cols, iaer=6, aerosols=dict(ecaer=e))
    shp = (cols.ncol, cols.nlay, 14)
    on = (e[:, :, 2] > 0)[:, :, None]
    ten = oracle.rrtmg_sw(cols, iaer=10, aerosols=dict(
        tauaer=np.asfortranarray(np.broadcast_to(tau[None, None, :, 2], shp) * e[:, :, 2:3]),
        ssaaer=np.asfortranarray(np.where(on, np.broadcast_to(piz[None, None, :, 2], shp), 1.0)),
        asmaer=np.asfortranarray(np.where(on, np.broadcast_to(asy[None, None, :, 2], shp), 0.0))))
    for k in SW:
        scale = max(np.abs(ten[k]).max(), 1.0)
        assert np.max(np.abs(six[k] - ten[k])) < 1e-12 * scale, k
    assert (six["swdflx"][cols.coszen > 0.1, 0] < ref["swdflx"][cols.coszen > 0.1, 0]).all()


def test_cldprop_sw_parameterisations(oracle, cols):
    """inflgsw = 2: a pure liquid cloud reproduces inflgsw = 0 fed with the Hu-Stamnes properties read off the table at an
    integer radius (extinction x path, ssa, g, forward fraction g^2); the Fortran stops come back as return codes."""
    shp = (cols.ncol, cols.nlay)
    cf = np.zeros(shp, order="F"); cf[:, 4:7] = 1.0
    lwp = np.asfortranarray(np.full(shp, 25.0) * cf)
    cl = dict(cldfr=cf, cicewp=np.zeros(shp, order="F"), cliqwp=lwp, reice=np.full(shp, 40.0, order="F"),
              reliq=np.full(shp, 11.5, order="F"))                       # index = int(11.5 - 1.5) = 10, fint = 0
    a = oracle.rrtmg_sw(cols, icld=2, inflgsw=2, iceflgsw=2, liqflgsw=1, clouds=cl)
    ext = oracle.table("swcld.extliq1").reshape(58, 14, order="F")[9]
    ssa = oracle.table("swcld.ssaliq1").reshape(58, 14, order="F")[9]
    asy = oracle.table("swcld.asyliq1").reshape(58, 14, order="F")[9]
    full = (14,) + shp
    b = oracle.rrtmg_sw(cols, icld=2, clouds=dict(
        cldfr=cf, taucld=np.asfortranarray(ext[:, None, None] * lwp[None]), ssacld=np.asfortranarray(np.broadcast_to(ssa[:, None, None], full)),
        asmcld=np.asfortranarray(np.broadcast_to(asy[:, None, None], full)),
        fsfcld=np.asfortranarray(np.broadcast_to((asy * asy)[:, None, None], full))))
    for k in SW:
        scale = max(np.abs(b[k]).max(), 1.0)
        assert np.max(np.abs(a[k] - b[k])) < 1e-9 * scale, k
    d0, d1 = np.flatnonzero(cols.coszen > 0.1)[:2]                        # night columns never reach cldprop_sw
    cl["cicewp"][d0, 5] = 3.0
    cl["reice"][d0, 5] = 4.0
    for ice, rc in ((1, 11), (2, 11), (3, 12)):
        with pytest.raises(RuntimeError, match=f"rc={rc}"):
            oracle.rrtmg_sw(cols, icld=2, inflgsw=2, iceflgsw=ice, liqflgsw=1, clouds=cl)
    cl["reice"][d0, 5] = 40.0
    cl["reliq"][d1, 4] = 61.0
    with pytest.raises(RuntimeError, match="rc=13"):
        oracle.rrtmg_sw(cols, icld=2, inflgsw=2, iceflgsw=2, liqflgsw=1, clouds=cl)
