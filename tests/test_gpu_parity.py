"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(mima_b200.rrtmg -> librrtmg_b200.so); the oracle is only the checker.

Tolerances (BASELINE.json north_star): fluxes 1e-6 relative, heating rates 1e-4 K/day.  Integer/index work
(jp, jt, jt1, indself, indfor, indminor, laytrop, laysolfr) and the reduced coefficient tables must be
bit-exact.  The CUDA kernels reproduce the oracle far more tightly than the tolerance; the tighter bounds
asserted below (1e-10 on optical depths, 1e-9 on fluxes) guard against regressions hiding in the slack.
"""
import numpy as np
import pytest

from mima_b200.columns import make_columns

pytestmark = pytest.mark.gpu

FLUX_RTOL = 1e-6       # north_star
HR_ATOL = 1e-4         # K/day, north_star
TIGHT = 1e-9

LW_OUT = ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")
SW_OUT = ("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc")
DEFAULT_SW_VARIANT = 4
LW_NGS = [0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140]


def _check_outputs(got, ref, names, tight=True, hr_tight=1e-7):
    for g, n in zip(got, names):
        o = ref[n]
        assert np.isfinite(g).all(), n
        if "hr" in n:
            assert np.max(np.abs(g - o)) < HR_ATOL, (n, float(np.max(np.abs(g - o))))
            if tight:
                assert np.max(np.abs(g - o)) < hr_tight, (n, float(np.max(np.abs(g - o))))
        else:
            scale = np.maximum(np.maximum(np.abs(o), 1e-6 * np.abs(o).max()), 1e-300)   # all-night batches are all zero
            r = np.max(np.abs(g - o) / scale)
            assert r < FLUX_RTOL, (n, float(r))
            if tight:
                assert r < TIGHT, (n, float(r))


def _relerr(g, o):
    d = np.abs(g - o)
    return np.where(o != 0, d / np.maximum(np.abs(o), 1e-300), d)


@pytest.fixture(scope="module")
def t42(oracle):
    cols = make_columns("T42L40", nlon=64, nlat=16)
    return cols, oracle.rrtmg_lw(cols, stages=True), oracle.rrtmg_sw(cols, stages=True)


def test_reduced_tables_bit_exact(gpu, oracle):
    """rrtmg_lw_ini / rrtmg_sw_ini: the 16 -> ngc reduction (cmbgbNN) and lookup tables, bit for bit."""
    names = []
    for b in range(1, 17):
        names += [f"lw{b:02d}.absa", f"lw{b:02d}.selfref", f"lw{b:02d}.forref", f"lw{b:02d}.fracrefa"]
    names += ["lw01.ka_mn2", "lw01.kb_mn2", "lw03.absb", "lw03.ka_mn2o", "lw03.kb_mn2o", "lw03.fracrefb", "lw05.ka_mo3",
              "lw05.ccl4", "lw06.ka_mco2", "lw06.cfc11adj", "lw06.cfc12", "lw07.kb_mco2", "lw08.ka_mo3", "lw08.cfc22adj",
              "lw09.kb_mn2o", "lw11.ka_mo2", "lw13.ka_mco", "lw13.kb_mo3", "lw15.ka_mn2", "lw16.absb"]
    for b in range(16, 30):
        names += [f"sw{b}.sfluxref"]
    names += ["sw16.absa", "sw16.absb", "sw17.absb", "sw17.forref", "sw20.absch4", "sw22.selfref", "sw23.rayl",
              "sw24.rayla", "sw24.raylb", "sw24.abso3a", "sw25.abso3b", "sw27.absb", "sw28.absa", "sw29.absh2o", "sw29.absco2",
              "lw.exp_tbl", "lw.tfn_tbl", "sw.exp_tbl"]
    for n in names:
        a, b = gpu.get_table(n), oracle.table(n)
        assert a.shape == b.shape and np.array_equal(a, b), n


def test_lw_stages(gpu, t42):
    cols, olw, _ = t42
    gpu.set_option("capture_stages", 1)
    gpu.set_option("chunk", 1 << 20)
    try:
        got = gpu.lw_from_columns(cols)
        nc, nl = cols.ncol, cols.nlay
        st = olw["stages"]
        assert np.array_equal(gpu.get_stage("lw.laytrop", (nc,)), st["laytrop"])
        low = np.arange(1, nl + 1)[None, :] <= st["laytrop"][:, None]
        for f in ("jp", "jt", "jt1", "indfor", "indminor"):
            assert np.array_equal(gpu.get_stage("lw." + f, (nc, nl)), st[f]), f
        assert np.array_equal(gpu.get_stage("lw.indself", (nc, nl))[low], st["indself"][low])
        for f in ("colh2o", "colco2", "colo3", "coln2o", "colco", "colch4", "colo2", "colbrd", "selffac", "forfac",
                  "forfrac", "minorfrac", "scaleminor", "scaleminorn2", "coldry"):
            assert np.array_equal(gpu.get_stage("lw." + f, (nc, nl)), st[f]), f      # no transcendental involved
        for f in ("fac00", "fac01", "fac10", "fac11"):                               # depend on log(p): ulp-level
            assert np.allclose(gpu.get_stage("lw." + f, (nc, nl)), st[f], rtol=1e-11, atol=1e-14), f
        for f, shp in (("planklay", (nc, nl, 16)), ("planklev", (nc, nl + 1, 16)), ("plankbnd", (nc, 16))):
            assert np.array_equal(gpu.get_stage("lw." + f, shp), st[f]), f
        for f in ("taug", "fracs"):
            g = gpu.get_stage("lw." + f, (nc, nl, 140))
            o = st[f]
            scale = np.maximum(np.abs(o), 1e-12 * np.abs(o).max())
            for b in range(16):
                sl = slice(LW_NGS[b], LW_NGS[b + 1])
                r = np.max(np.abs(g[:, :, sl] - o[:, :, sl]) / scale[:, :, sl])
                assert r < 1e-10, (f, "band", b + 1, float(r))
        _check_outputs(got, olw, LW_OUT)
    finally:
        gpu.set_option("capture_stages", 0)
        gpu.set_option("chunk", 0)


def test_sw_stages(gpu, t42):
    cols, _, osw = t42
    gpu.set_option("capture_stages", 1)
    gpu.set_option("chunk", 1 << 20)
    try:
        got = gpu.sw_from_columns(cols)
        nc, nl = cols.ncol, cols.nlay
        st = osw["stages"]
        assert np.array_equal(gpu.get_stage("sw.laytrop", (nc,)), st["laytrop"])
        for f in ("jp", "jt", "jt1", "indfor", "indself"):
            assert np.array_equal(gpu.get_stage("sw." + f, (nc, nl)), st[f]), f
        for f in ("colh2o", "colco2", "colo3", "colch4", "colo2", "colmol", "selffac", "selffrac", "forfac", "forfrac"):
            assert np.array_equal(gpu.get_stage("sw." + f, (nc, nl)), st[f]), f
        for f in ("taug", "taur"):
            g = gpu.get_stage("sw." + f, (nc, nl, 112))
            o = st[f]
            scale = np.maximum(np.abs(o), 1e-12 * np.abs(o).max())
            assert np.max(np.abs(g - o) / scale) < 1e-10, f
        assert np.allclose(gpu.get_stage("sw.sfluxzen", (nc, 112)), st["sfluxzen"], rtol=1e-14)
        _check_outputs(got, osw, SW_OUT)
    finally:
        gpu.set_option("capture_stages", 0)
        gpu.set_option("chunk", 0)


@pytest.mark.parametrize("kw", [
    dict(resolution="T42L40", nlon=128, nlat=8, night=True),                                        # config 1 slice, with night
    dict(resolution="T42L40", nlon=64, nlat=16, co2_ppmv=1560.0, ozone="file", secondary_gases=True),  # config 4
    dict(resolution="T170L60", nlon=64, nlat=8),                                                    # L60
    dict(resolution="T341L80", nlon=32, nlat=8, secondary_gases=True),                              # L80 (LMAX=128 path)
])
def test_end_to_end_parity(gpu, oracle, kw):
    cols = make_columns(**kw)
    _check_outputs(gpu.lw_from_columns(cols), oracle.rrtmg_lw(cols), LW_OUT)
    _check_outputs(gpu.sw_from_columns(cols), oracle.rrtmg_sw(cols), SW_OUT)


def test_set_table_path_gives_the_same_reduced_tables(gpu):
    """What the Fortran shim's rrtmg_lw_ini / rrtmg_sw_ini do (shim/rrtmg_b200_tables.f90): every array of the generated list
    through rrtmg_b200_set_table, then init -- the reduced tables must equal those of the load_tables path bit for bit."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_shim_tables", os.path.join(root, "tools", "gen_shim_tables.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    probe = ["lw01.absa", "lw03.absb", "lw05.ka_mo3", "lw07.kb_mco2", "lw13.ka_mco", "lw16.fracrefa", "lw.exp_tbl", "lw.tfn_tbl",
             "sw16.absa", "sw17.absb", "sw24.rayla", "sw29.absco2", "sw27.sfluxref", "sw.exp_tbl"]
    before = {n: gpu.get_table(n) for n in probe}
    blobs = {}
    for f in ("rrtmg_lw_ref.bin", "rrtmg_lw_kg_synth.bin", "rrtmg_sw_kg.bin"):
        blobs.update(g.bt.read_blob(os.path.join(gpu.DATA_DIR, f)))
    try:
        gpu.finalize()
        for kind in ("lw", "sw"):
            for name, _, _, shape in g.table_list(kind):
                gpu.set_table(name, np.asfortranarray(blobs[name]).reshape(shape, order="F"))
        gpu.set_table("lwmeta.synthetic", np.array([1.0]))
        gpu.rrtmg_lw_ini(default_tables=False, allow_synthetic_lw=True)
        gpu.rrtmg_sw_ini(default_tables=False)
        for n in probe:
            assert np.array_equal(gpu.get_table(n), before[n]), n
        cols = make_columns("T42L40", nlon=16, nlat=4)
        a = gpu.lw_from_columns(cols) + gpu.sw_from_columns(cols)
    finally:
        gpu.finalize()
        gpu.rrtmg_lw_ini(allow_synthetic_lw=True)
        gpu.rrtmg_sw_ini()
    b = gpu.lw_from_columns(cols) + gpu.sw_from_columns(cols)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_sw_ignores_cloud_switches_when_icld_is_zero(gpu):
    """run_rrtmg passes the LONGWAVE cloud switches inflglw, iceflglw, liqflglw (and LW-shaped cloud/aerosol arrays) in the
    shortwave call (rrtm_radiation.f90:692-694, 706-708).  With icld = 0 the reference never looks at them; neither does
    the library: any switch values, no cloud arrays, same bits."""
    cols = make_columns("T42L40", nlon=32, nlat=4, night=True)
    ref = gpu.sw_from_columns(cols)
    for flags in ((2, 3, 1), (1, 0, 0), (7, -1, 9)):
        got = gpu.sw_from_columns(cols, inflgsw=flags[0], iceflgsw=flags[1], liqflgsw=flags[2])
        assert all(np.array_equal(a, b) for a, b in zip(got, ref)), flags


def test_share_inputs_option(gpu):
    """Option share_inputs: rrtmg_b200_lw reuses the device copies rrtmg_b200_sw made of the eleven arrays both read -- when
    it is called right after it with the same host arrays; with other arrays (or after another LW call) it uploads its own.
    Results are bitwise those of the plain calls, also with night columns, chunked blocks and a ragged column count."""
    cols = make_columns("T170L60", nlon=83, nlat=5, night=True)
    other = make_columns("T170L60", nlon=83, nlat=5, night=True, seed=99)
    plain = gpu.sw_from_columns(cols) + gpu.lw_from_columns(cols)
    plain_other = gpu.lw_from_columns(other)
    try:
        gpu.set_option("share_inputs", 1)
        for hc, slots in ((0, 4), (64, 4), (64, 2), (64, 3), (32, 6)):
            gpu.set_option("host_chunk", hc)
            gpu.set_option("host_slots", slots)                                    # pipeline depth of the host-pointer calls
            got = gpu.sw_from_columns(cols) + gpu.lw_from_columns(cols)            # LW reuses the SW uploads
            assert all(np.array_equal(a, b) for a, b in zip(got, plain)), (hc, slots)
            again = gpu.lw_from_columns(cols)                                      # one shot: this one uploads itself
            assert all(np.array_equal(a, b) for a, b in zip(again, plain[6:]))
            gpu.sw_from_columns(cols)
            got_other = gpu.lw_from_columns(other)                                 # other arrays: not the shared copies
            assert all(np.array_equal(a, b) for a, b in zip(got_other, plain_other))
    finally:
        gpu.set_option("share_inputs", 0)
        gpu.set_option("host_chunk", 0)
        gpu.set_option("host_slots", 4)


def test_reference_golden_vectors(gpu):
    """The CUDA path against outputs of the reference's own code (tests/golden/ref_t42l40.npz: the reference's RRTMG
    sources machine-translated F90 -> C and run on config C4 columns, tests/golden/make_ref_vectors.py) -- no oracle in
    between.  north_star tolerances, and the tighter regression bounds."""
    import os
    from golden.make_ref_vectors import batch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_t42l40.npz"))
    c = batch()
    _check_outputs(gpu.lw_from_columns(c), g, LW_OUT)
    _check_outputs(gpu.sw_from_columns(c), g, SW_OUT)


def _dev_variants(gpu):
    try:
        gpu.set_option("dev_variants", 1)
        return True
    except gpu.RRTMGError:
        return False


def test_sw_solver_variants_agree(gpu, oracle):
    """Development builds only (RRTMG_B200_DEV_VARIANTS=1): the default SW solver (variant 4, one-warp blocks: the reference's
    top-down recurrence first, then the upward flux from coefficients kept by that sweep, algebraically vrtqdr_sw :103-150)
    against variant 3 (the same scheme in 7-warp blocks, bitwise equal sums), 2 (bottom-up first), 1 (both of the
    reference's recurrences literally), 0 (the first version), 5 and 6 (the kept coefficients in an L2-resident scratch /
    shared memory instead of local memory).  All must match the oracle, and each other to rounding."""
    if not _dev_variants(gpu):
        pytest.skip("the library was built without RRTMG_B200_DEV_VARIANTS: it carries the default solver only")
    cols = make_columns("T170L60", nlon=64, nlat=8, night=True)
    ref = oracle.rrtmg_sw(cols)
    res = {}
    try:
        for v in (6, 5, 4, 3, 2, 1, 0):
            gpu.set_option("sw_solver_variant", v)
            res[v] = gpu.sw_from_columns(cols)
            _check_outputs(res[v], ref, SW_OUT)
        gpu.set_option("sw_solver_variant", 5)
        for wpb, flags, ns in ((12, 3, 0), (16, 1, 1), (24, 2, 4), (28, 0, 0)):    # x0 blocks/SM, x1 policy|discard, x2 shared levels + 1
            for k, v in (("x0", wpb), ("x1", flags), ("x2", ns)):
                gpu.set_option(k, v)
            _check_outputs(gpu.sw_from_columns(cols), ref, SW_OUT)
        gpu.set_option("sw_solver_variant", 6)                                     # exp table in shared memory, one block per SM
        for wps, flags in ((16, 3), (20, 0), (24, 1), (28, 2)):
            gpu.set_option("x0", wps)
            gpu.set_option("x1", flags)
            _check_outputs(gpu.sw_from_columns(cols), ref, SW_OUT)
    finally:
        gpu.set_option("sw_solver_variant", DEFAULT_SW_VARIANT)
        for k, v in (("x0", 0), ("x1", 3), ("x2", 0)):
            gpu.set_option(k, v)
    for v in (5, 6):
        for a, b, n in zip(res[v], res[4], SW_OUT):                       # stack in L2 / shared memory: same formulas
            assert np.max(np.abs(a - b)) < (1e-7 if "hr" in n else 1e-8), (v, n)
    assert all(np.array_equal(a, b) for a, b in zip(res[4], res[3]))      # same sums in the same order
    for v in (2, 3):
        for a, b, n in zip(res[v], res[1], SW_OUT):
            if "hr" in n:
                assert np.max(np.abs(a - b)) < 1e-7, (v, n)          # K/day
            else:
                assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * np.abs(b).max())) < 1e-10, (v, n)


def test_emissivity_and_aerosol_inputs(gpu, oracle):
    """Spectrally varying emissivity (reflection term, rtrnmr.f90:628-636) and a non-zero LW tauaer
    (iaer = 10 is forced, rad.nomcica:442, :514-519)."""
    cols = make_columns("T42L40", nlon=32, nlat=4)
    rng = np.random.default_rng(7)
    cols.emis = np.asfortranarray(rng.uniform(0.85, 1.0, (cols.ncol, 16)))
    taer = np.asfortranarray(rng.uniform(0.0, 0.05, (cols.ncol, cols.nlay, 16)))
    _check_outputs(gpu.lw_from_columns(cols, tauaer=taer), oracle.rrtmg_lw(cols, tauaer=taer), LW_OUT)


def test_ragged_sizes_and_chunking(gpu, oracle):
    """ncol not a multiple of any tile, single column, and a chunk smaller than the batch."""
    base = make_columns("T42L40", nlon=64, nlat=4, night=True)
    for n in (1, 3, 129, 200):
        c = base.take(np.arange(n))
        _check_outputs(gpu.lw_from_columns(c), oracle.rrtmg_lw(c), LW_OUT)
        _check_outputs(gpu.sw_from_columns(c), oracle.rrtmg_sw(c), SW_OUT)
    gpu.set_option("chunk", 50)
    try:
        c = base.take(np.arange(173))
        _check_outputs(gpu.lw_from_columns(c), oracle.rrtmg_lw(c), LW_OUT)
        _check_outputs(gpu.sw_from_columns(c), oracle.rrtmg_sw(c), SW_OUT)
    finally:
        gpu.set_option("chunk", 0)
    # empty batch is a no-op
    e = base.take(np.arange(0))
    assert gpu.lw_from_columns(e)[0].shape == (0, 41)


def test_night_columns_and_clear_equals_total(gpu):
    c = make_columns("T42L40", nlon=64, nlat=4, night=True)
    sw = gpu.sw_from_columns(c)
    night = c.coszen < 1e-10
    assert night.any()
    for a in sw:
        assert (a[night] == 0).all()
    assert np.array_equal(sw[0], sw[3]) and np.array_equal(sw[1], sw[4]) and np.array_equal(sw[2], sw[5])
    lw = gpu.lw_from_columns(c)
    assert np.array_equal(lw[0], lw[3]) and np.array_equal(lw[2], lw[5])


def test_unsupported_options_fail_loudly(gpu):
    c = make_columns("T42L40", nlon=4, nlat=2)
    args = (c.play, c.plev, c.tlay, c.tlev, c.tsfc, c.h2o, c.o3, c.co2, None, None, None)
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_lw(c.ncol, c.nlay, 1, 0, *args, None, None, None, None, None)
    assert e.value.code == 4                      # icld > 0 without the cloud arrays
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_lw(c.ncol, c.nlay, 1, 0, *args, None, None, None, None, None, inflglw=2, cldfr=c.tlay * 0)
    assert e.value.code == 4                      # inflglw = 2 without the water paths / radii
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_lw(c.ncol, c.nlay, 0, 2, *args, None, None, None, None, None)      # idrv must be 0 or 1
    assert e.value.code == 4
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_sw(c.ncol, c.nlay, 0, 6, *args, c.albedo, c.albedo, c.albedo, c.albedo, c.coszen, 1.0, 0, 1370.0)
    assert e.value.code == 4                      # iaer = 6 without ecaer
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_sw(c.ncol, c.nlay, 0, 10, *args, c.albedo, c.albedo, c.albedo, c.albedo, c.coszen, 1.0, 0, 1370.0)
    assert e.value.code == 4                      # iaer = 10 without the aerosol arrays
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_sw(c.ncol, c.nlay, 2, 0, *args, c.albedo, c.albedo, c.albedo, c.albedo, c.coszen, 1.0, 0, 1370.0, inflgsw=2,
                     cldfr=c.tlay * 0)
    assert e.value.code == 4                      # inflgsw = 2 without the water paths / radii
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_sw(c.ncol, c.nlay, 2, 0, *args, c.albedo, c.albedo, c.albedo, c.albedo, c.coszen, 1.0, 0, 1370.0, inflgsw=1,
                     cldfr=c.tlay * 0)
    assert e.value.code == 2                      # cldprop_sw has no inflag = 1 branch
    with pytest.raises(gpu.RRTMGError) as e:
        gpu.rrtmg_lw(c.ncol, 200, 0, 0, *args, None, None, None, None, None)
    assert e.value.code == 4


@pytest.mark.parametrize("res", ["T85L40", "T170L60"])
def test_full_size_properties(gpu, res):
    """BASELINE configs 2 and 3 at full size (32768 x 40, 131072 x 60: two device passes), checked through
    size-independent properties: shard invariance (any split of the batch gives bit-identical columns), heating =
    flux divergence, TOA insolation = S0 cos(z), surface reflection, clear == total."""
    c = make_columns(res)
    lw = gpu.lw_from_columns(c)
    sw = gpu.sw_from_columns(c)
    for a in lw + sw:
        assert np.isfinite(a).all()
    heatfac = 9.8066 * 86400.0 / (1004.64 * 100.0)
    fnet = lw[0] - lw[1]
    hr = heatfac * (fnet[:, :-1] - fnet[:, 1:]) / (c.plev[:, :-1] - c.plev[:, 1:])
    assert np.allclose(lw[2], hr, rtol=1e-10, atol=1e-10)
    assert np.allclose(sw[1][:, -1], c.scon * c.coszen, rtol=2e-5)
    assert np.allclose(sw[0][:, 0], c.albedo * sw[1][:, 0], rtol=1e-9)
    assert np.array_equal(sw[0], sw[3]) and np.array_equal(lw[1], lw[4])
    # shard invariance: a block of latitude rows computed alone equals the same rows of the full call (the block
    # straddles a device-pass boundary of the full call at T170L60)
    j0, j1 = c.nlat // 4 - 16, c.nlat // 4 + 16
    blk = c.rows(j0, j1)
    lwb, swb = gpu.lw_from_columns(blk), gpu.sw_from_columns(blk)
    s = slice(j0 * c.nlon, j1 * c.nlon)
    for a, b in zip(lw + sw, lwb + swb):
        assert np.array_equal(a[s], b)


def test_repeated_calls_are_bitwise_reproducible(gpu):
    """Guards the mbarrier/TMA stage ring of lw_rtrn (a stage refilled under outstanding shared-memory loads showed
    up as 1-3 wrong columns per 16384 in one call out of ten) and the warp-local sums: 25 calls, identical bits."""
    c = make_columns("T170L60", lat_rows=(48, 80))
    lw0, sw0 = gpu.lw_from_columns(c), gpu.sw_from_columns(c)
    for _ in range(24):
        lw = gpu.lw_from_columns(c)
        for a, b in zip(lw0, lw):
            assert np.array_equal(a, b)
    for _ in range(4):
        sw = gpu.sw_from_columns(c)
        for a, b in zip(sw0, sw):
            assert np.array_equal(a, b)


def test_idrv1_flux_derivative(gpu, oracle):
    """idrv = 1 (rad.nomcica:143-152; setcoef.f90:197-201; rtrnmr.f90:629-746): the upward-flux derivative with respect
    to the surface temperature.  GPU vs oracle at 1e-9, the six standard outputs unchanged, and the
    derivative against a centred finite difference of the forward model (known answer independent of the restatement;
    3e-3: the Planck table is piecewise linear in 1 K steps)."""
    import dataclasses
    c = make_columns("T170L60", nlon=64, nlat=4)
    rng = np.random.default_rng(11)
    c.emis = np.asfortranarray(rng.uniform(0.9, 1.0, (c.ncol, 16)))
    got = gpu.lw_from_columns(c, idrv=1)
    ref = oracle.rrtmg_lw(c, idrv=1)
    # idrv = 1 runs the staged kernels, a clear-sky call without derivatives the fused column kernel (another summation
    # order over the g-points): bit for bit against the staged kernels, to rounding against the fused one
    gpu.set_option("lw_fused", 0)
    try:
        base = gpu.lw_from_columns(c)
    finally:
        gpu.set_option("lw_fused", 1)
    for a, b in zip(got[:6], base):
        assert np.array_equal(a, b)
    for a, b in zip(got[:6], gpu.lw_from_columns(c)):
        assert np.max(np.abs(a - b)) <= 1e-11 * np.abs(b).max()
    for g, n in zip(got[6:], ("duflx_dt", "duflxc_dt")):
        r = np.max(np.abs(g - ref[n]) / np.abs(ref[n]).max())
        assert r < 1e-9, (n, float(r))
    up = gpu.lw_from_columns(dataclasses.replace(c, tsfc=c.tsfc + 0.5))[0]
    dn = gpu.lw_from_columns(dataclasses.replace(c, tsfc=c.tsfc - 0.5))[0]
    fd = up - dn
    assert np.max(np.abs(fd - got[6])) < 3e-3 * np.abs(fd).max()
    assert (got[6] > 0).all() and (np.diff(got[6], axis=1) <= 1e-12).all()      # attenuated on the way up


def test_clear_sky_outputs_are_optional(gpu):
    """NULL for the clear-sky output arrays of the host-pointer ABI: the total-sky results are unchanged, bit for bit."""
    c = make_columns("T42L40", nlon=64, nlat=4, night=True)
    gpu.set_option("host_chunk", 100)
    try:
        for full, part in ((gpu.lw_from_columns(c, idrv=1), gpu.lw_from_columns(c, idrv=1, clear_sky=False)),
                           (gpu.sw_from_columns(c), gpu.sw_from_columns(c, clear_sky=False))):
            assert all(p is None for p in (part[3], part[4], part[5]))
            for i in (0, 1, 2):
                assert np.array_equal(full[i], part[i])
            if len(full) == 8:
                assert np.array_equal(full[6], part[6]) and part[7] is None
    finally:
        gpu.set_option("host_chunk", 0)
