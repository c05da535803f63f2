"""The hand-written oracle (oracle/*.c) against oracle/_ref: the reference's own RRTMG sources, machine-translated
F90 -> C by tools/f90_to_c.py and compiled with the oracle's flags (`make -C oracle _ref`).  Every comparison is bit for
bit: the two are independent renderings of the same arithmetic in the same order, so any difference is a transcription
error in the oracle (the first run of this file found one: the upper-atmosphere foreign continuum of SW bands 17 and 21
is colh2o*forfac*(...) evaluated left to right, rrtmg_sw_taumol.f90:451-454, not colh2o*(forfac*(...))).

SW runs on the reference's own coefficient file (SW/src/rrtmg_sw_k_g.f90, translated like the code); LW on the
coefficients the oracle loads, written out in the layout of the stripped LW/src/rrtmg_lw_k_g.f90 and translated the
same way -- so LW is pinned as an algorithm (init reduction included), not as data.

CPU only.  Where neither the prebuilt library nor /root/reference exists the module is skipped.
"""
import numpy as np
import pytest

from mima_b200.columns import make_columns
from oracle import pyref

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built and /root/reference is absent")

LW = ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")
SW = ("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc")


@pytest.fixture(scope="module")
def ref():
    return pyref.Reference()


def _same(a, b, keys, what):
    for k in keys:
        assert np.isfinite(b[k]).all(), (what, k)
        assert np.array_equal(a[k], b[k]), (what, k, float(np.max(np.abs(a[k] - b[k]))))


# ---------------------------------------------------------------------------------------------- translator units
def test_translator_expressions():
    """Lexing and operator rules the RRTMG sources rely on."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import f90_to_c as t
    toks = t.tokenize("if (iout.gt.0.and.iband.ge.2) x = 1.e-20_rb*y**2 - 5*(plog+0.04_rb)")
    kinds = [k for k, _ in toks]
    assert ("int", "0") in toks and ("op", ".and.") in toks and ("real", "1.e-20") in toks and "end" == kinds[-1]
    e = t.parse_expr("-a**2*b")                         # -( (a**2) * b )
    assert e[0] == "un" and e[2][0] == "bin" and e[2][1] == "*" and e[2][2][1] == "**"
    e = t.parse_expr("a**b**c")                         # right associative
    assert e[1] == "**" and e[3][1] == "**"
    e = t.parse_expr("x .lt. 1 .or. .not. y .and. z")   # .not. binds tighter than .and., .and. tighter than .or.
    assert e[1] == ".or." and e[3][1] == ".and." and e[3][2][0] == "un"
    assert t.parse_arg("13:59")[0] == "sec" and t.parse_arg(":")[0] == "sec"
    assert t.parse_expr("(/ 1._rb, 2.5e+01_rb /)")[0] == "ctor"
    assert t.find_assign("a(i,j) = b(k) .le. c") == 7 and t.find_assign("if (a == b) c") < 0


def test_translator_semantics(tmp_path):
    """A small module through translator + gcc: lower bounds, column-major sections, integer division, real -> integer
    truncation, x**n, mod, min/max, goto, internal subroutine with host association, optional argument."""
    import os, subprocess, ctypes as C, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import f90_to_c as t
    src = tmp_path / "unit.f90"
    src.write_text("""
      module unit
      implicit none
      integer, parameter :: n = 4
      real(kind=8) :: tab(0:n, 2:3)
      contains
      subroutine run(k, x, out, extra)
      integer, intent(in) :: k
      real(kind=8), intent(in) :: x
      real(kind=8), intent(out) :: out(:)
      real(kind=8), intent(in), optional :: extra(:)
      integer :: i, j
      real(kind=8) :: acc, w(0:k)
      tab(:, 2) = (/ 1.0d0, 2.0d0, 3.0d0, 4.0d0, 5.0d0 /)
      tab(1:3, 3) = 7.
      w(:) = 0.5
      i = -7.9
      out(1) = i
      out(2) = 7/2 + (-7)/2
      out(3) = x**3 + 2**k
      out(4) = mod(7.5d0, 2.0d0) + mod(-7, 3)
      out(5) = min(3, max(1, int(x))) + tab(4,2) + tab(2,3) + tab(0,3)
      acc = 0.
      do j = k, 0, -1
         acc = acc + w(j) * j
         if (j .eq. 2) goto 10
      enddo
 10   continue
      out(6) = acc
      call inner
      out(8) = 0.
      if (present(extra)) out(8) = extra(2)
      contains
      subroutine inner
      out(7) = acc * k + n
      end subroutine inner
      end subroutine run
      end module unit
""")
    out = tmp_path / "c"
    assert t.main(["f90_to_c", "-o", str(out), str(src)]) == 0
    (out / "stop.c").write_text('#include "f90ref.h"\nvoid f90_stop(const char *m) { abort(); }\n')
    so = tmp_path / "unit.so"
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(out / "unit.c"),
                    str(out / "f90ref_globals.c"), str(out / "stop.c"), "-lm"], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    res = np.zeros(8)
    ex = np.array([1.5, 2.5])
    dp = C.POINTER(C.c_double)
    lib.unit__run(C.c_int(4), C.c_double(2.5), res.ctypes.data_as(dp), C.c_int(8), ex.ctypes.data_as(dp), C.c_int(2))
    acc = 0.5 * 4 + 0.5 * 3 + 0.5 * 2
    assert res.tolist() == [-7.0, 3 - 3, 2.5 * 2.5 * 2.5 + 16, 1.5 + -1, 2 + 5.0 + 7.0 + 0.0, acc, acc * 4 + 4, 2.5]
    lib.unit__run(C.c_int(4), C.c_double(2.5), res.ctypes.data_as(dp), C.c_int(8), None, C.c_int(0))
    assert res[7] == 0.0


# ---------------------------------------------------------------------------------------------- init: reduced tables
def test_reduced_tables(ref, oracle):
    """rrtmg_sw_ini / rrtmg_lw_ini of the translated reference (cmbgb16s..29, cmbgb1..16, lookup tables) against the
    oracle's tables."""
    pairs = [("rrsw_tbl.exp_tbl", "sw.exp_tbl"), ("rrlw_tbl.exp_tbl", "lw.exp_tbl"), ("rrlw_tbl.tfn_tbl", "lw.tfn_tbl")]
    for b in range(16, 30):
        for n in ("absa", "absb", "selfref", "forref", "sfluxref"):
            pairs.append((f"rrsw_kg{b}.{n}", f"sw{b}.{n}"))
    for b in range(1, 17):
        for n in ("absa", "absb", "selfref", "forref", "fracrefa", "fracrefb"):
            pairs.append((f"rrlw_kg{b:02d}.{n}", f"lw{b:02d}.{n}"))
    names = set(ref.names())
    checked = 0
    for rn, on in pairs:
        if rn not in names:
            continue
        try:
            o = oracle.table(on)
        except KeyError:
            continue
        r = ref.var(rn)
        assert r.size == o.size and np.array_equal(r, o), (rn, on)
        checked += 1
    assert checked >= 100, checked


# ---------------------------------------------------------------------------------------------- end to end
@pytest.mark.parametrize("kw", [dict(resolution="T42L40", nlon=64, nlat=16, night=True),
                                dict(resolution="T170L60", nlon=48, nlat=8),
                                dict(resolution="T341L80", nlon=32, nlat=6, night=True),
                                dict(resolution="T42L40", nlon=48, nlat=16, co2_ppmv=1560.0, ozone="file", secondary_gases=True),
                                dict(resolution="T42L40", nlon=16, nlat=4, nlay=1), dict(resolution="T42L40", nlon=16, nlat=4, nlay=7)],
                         ids=["T42L40-night", "T170L60", "T341L80", "C4-4xCO2-ozone-file", "one-layer", "seven-layers"])
def test_clear_sky(ref, oracle, kw):
    cols = make_columns(**kw)
    _same(oracle.rrtmg_sw(cols), ref.rrtmg_sw(cols), SW, "sw")
    _same(oracle.rrtmg_lw(cols), ref.rrtmg_lw(cols), LW, "lw")


@pytest.mark.parametrize("seed,nlay", [(1, 40), (2, 60), (3, 25), (4, 80)])
def test_wild_columns(ref, oracle, seed, nlay):
    """Columns far outside the bench climate (mima_b200.columns.wild_columns): every gas and CFC present and varied over
    orders of magnitude, mountains and deep lows, +-35 K, grey surfaces, earth-sun adjustment."""
    from mima_b200.columns import wild_columns
    cols = wild_columns(seed, nlay)
    _same(oracle.rrtmg_sw(cols), ref.rrtmg_sw(cols), SW, "sw")
    _same(oracle.rrtmg_lw(cols, idrv=1), ref.rrtmg_lw(cols, idrv=1), LW + ("duflx_dt", "duflxc_dt"), "lw")


def test_idrv_emissivity_aerosol(ref, oracle):
    cols = make_columns("T42L40", nlon=32, nlat=4)
    rng = np.random.default_rng(7)
    cols.emis = np.asfortranarray(rng.uniform(0.85, 1.0, (cols.ncol, 16)))
    taer = np.asfortranarray(rng.uniform(0.0, 0.05, (cols.ncol, cols.nlay, 16)))
    _same(oracle.rrtmg_lw(cols, tauaer=taer, idrv=1), ref.rrtmg_lw(cols, tauaer=taer, idrv=1), LW + ("duflx_dt", "duflxc_dt"), "lw")


def test_extreme_columns(ref, oracle):
    """Temperatures beyond both ends of the Planck table, dry / saturated / ozone-free / 20 x CO2 columns, albedo 0 and 1,
    the sun at and just below the night threshold."""
    cols = make_columns("T42L40", nlon=32, nlat=2)
    n = cols.ncol
    cols.tlay[0:4] = 150.0; cols.tlev[0:4] = 150.0; cols.tsfc[0:4] = 150.0
    cols.tlay[4:8] = 345.0; cols.tlev[4:8] = 345.0; cols.tsfc[4:8] = 345.0
    cols.h2o[8:12] = 2e-7
    cols.h2o[12:16] = 0.04
    cols.o3[16:20] = 0.0
    cols.co2[20:24] *= 20.0
    cols.albedo[24:28] = 0.0
    cols.albedo[28:32] = 1.0
    cols.coszen[32:36] = 1e-10
    cols.coszen[36:40] = 0.99e-10
    cols.coszen[40:44] = 1.0
    _same(oracle.rrtmg_sw(cols), ref.rrtmg_sw(cols), SW, "sw")
    _same(oracle.rrtmg_lw(cols, idrv=1), ref.rrtmg_lw(cols, idrv=1), LW, "lw")
    assert n >= 44


def _lw_cloud_field(cols, rng):
    from test_oracle_lw_clouds import cloud_field
    return cloud_field(cols, rng)


@pytest.mark.parametrize("icld", [1, 2, 3])
def test_lw_clouds(ref, oracle, icld):
    """rtrn (random overlap) and rtrnmr (maximum/random): cloud optical depth per band as given (inflglw = 0)."""
    cols = make_columns("T42L40", nlon=32, nlat=6)
    cl = _lw_cloud_field(cols, np.random.default_rng(30 + icld))
    _same(oracle.rrtmg_lw(cols, icld=icld, clouds=cl, idrv=1), ref.rrtmg_lw(cols, icld=icld, clouds=cl, idrv=1),
          LW + ("duflx_dt", "duflxc_dt"), icld)


def _water_clouds(c, rng, partial):
    shp = (c.ncol, c.nlay)
    cf = (rng.uniform(size=shp) < 0.3) * (rng.uniform(0.1, 1.0, shp) if partial else 1.0)
    return dict(cldfr=np.asfortranarray(cf.astype(np.float64)),
                cicewp=np.asfortranarray(rng.uniform(0, 30, shp) * (rng.uniform(size=shp) < 0.7)),
                cliqwp=np.asfortranarray(rng.uniform(0, 60, shp) * (rng.uniform(size=shp) < 0.7)),
                reice=np.asfortranarray(rng.uniform(14, 120, shp)), reliq=np.asfortranarray(rng.uniform(3, 50, shp)))


@pytest.mark.parametrize("flags", [(1, 0, 0), (2, 0, 0), (2, 1, 0), (2, 1, 1), (2, 2, 1), (2, 3, 1), (2, 2, 0)])
def test_lw_cloud_optics(ref, oracle, flags):
    """cldprop (rrtmg_lw_cldprop.f90:151-270): every ice and liquid option, both overlap rules."""
    infl, ice, liq = flags
    cols = make_columns("T42L40", nlon=32, nlat=4)
    cl = _water_clouds(cols, np.random.default_rng(40 + 10 * ice + liq), partial=True)
    for icld in (1, 2):
        kw = dict(icld=icld, clouds=cl, inflglw=infl, iceflglw=ice, liqflglw=liq)
        _same(oracle.rrtmg_lw(cols, **kw), ref.rrtmg_lw(cols, **kw), LW, (flags, icld))


def test_sw_clouds_and_aerosols(ref, oracle):
    """spcvrt with clouds given by optical properties (inflgsw = 0, delta-M scaling of cldprop_sw) and aerosols given per band
    (iaer = 10) or by ECMWF type (iaer = 6)."""
    cols = make_columns("T42L40", nlon=32, nlat=8, night=True)
    rng = np.random.default_rng(11)
    shp = (14, cols.ncol, cols.nlay)
    cld = (rng.uniform(size=(cols.ncol, cols.nlay)) < 0.3).astype(np.float64)
    asm = rng.uniform(0.7, 0.9, shp)
    cl = dict(cldfr=np.asfortranarray(cld), taucld=np.asfortranarray(rng.uniform(0.0, 20.0, shp) * cld[None]),
              ssacld=np.asfortranarray(rng.uniform(0.5, 0.99999, shp)), asmcld=np.asfortranarray(asm), fsfcld=np.asfortranarray(asm * asm))
    a3 = (cols.ncol, cols.nlay, 14)
    aer = dict(tauaer=np.asfortranarray(rng.uniform(0.0, 0.3, a3)), ssaaer=np.asfortranarray(rng.uniform(0.6, 0.999, a3)),
               asmaer=np.asfortranarray(rng.uniform(0.3, 0.8, a3)))
    ec = dict(ecaer=np.asfortranarray(rng.uniform(0.0, 0.05, (cols.ncol, cols.nlay, 6))))
    for kw in (dict(icld=2, clouds=cl), dict(iaer=10, aerosols=aer), dict(icld=1, iaer=10, clouds=cl, aerosols=aer),
               dict(iaer=6, aerosols=ec), dict(icld=3, iaer=6, clouds=cl, aerosols=ec)):
        _same(oracle.rrtmg_sw(cols, **kw), ref.rrtmg_sw(cols, **kw), SW, sorted(kw))


@pytest.mark.parametrize("iceflg", [1, 2, 3])
def test_sw_cloud_optics(ref, oracle, iceflg):
    """cldprop_sw's inflag = 2 (rrtmg_sw_cldprop.f90:168-345)."""
    cols = make_columns("T42L40", nlon=32, nlat=8, night=True)
    cl = _water_clouds(cols, np.random.default_rng(60 + iceflg), partial=False)
    kw = dict(icld=2, inflgsw=2, iceflgsw=iceflg, liqflgsw=1, clouds=cl)
    _same(oracle.rrtmg_sw(cols, **kw), ref.rrtmg_sw(cols, **kw), SW, iceflg)


def test_stops(ref, oracle):
    """The Fortran `stop`s come back as errors from both."""
    cols = make_columns("T42L40", nlon=8, nlat=2)
    cl = _water_clouds(cols, np.random.default_rng(2), partial=False)
    cl["cldfr"][3, 5] = 1.0; cl["cicewp"][3, 5] = 10.0; cl["reice"][3, 5] = 4.0
    for f in (ref.rrtmg_sw, oracle.rrtmg_sw):
        with pytest.raises(RuntimeError):
            f(cols, icld=2, inflgsw=2, iceflgsw=2, liqflgsw=1, clouds=cl)
    for f in (ref.rrtmg_lw, oracle.rrtmg_lw):
        with pytest.raises(RuntimeError):
            f(cols, icld=2, inflglw=2, iceflglw=2, liqflglw=1, clouds=cl)
    part = dict(cl, cldfr=np.asfortranarray(np.full((cols.ncol, cols.nlay), 0.5)))
    for f in (ref.rrtmg_sw, oracle.rrtmg_sw):
        with pytest.raises(RuntimeError):                 # stop 'PARTIAL CLOUD NOT ALLOWED' (rad.nomcica:537)
            f(cols, icld=2, inflgsw=2, iceflgsw=2, liqflgsw=1, clouds=dict(part, reice=np.asfortranarray(np.full((cols.ncol, cols.nlay), 50.0))))


def test_golden_vectors_come_from_the_reference(ref):
    """tests/golden/ref_t42l40.npz was written by tests/golden/make_ref_vectors.py from the translated reference; it must
    still be what the translated reference computes (the GPU-box tests compare against the file)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "ref_t42l40.npz")
    g = np.load(path)
    from golden.make_ref_vectors import batch
    cols = batch()
    sw, lw = ref.rrtmg_sw(cols), ref.rrtmg_lw(cols, idrv=1)
    for k in SW:
        assert np.array_equal(g[k], sw[k]), k
    for k in LW + ("duflx_dt",):
        assert np.array_equal(g[k], lw[k]), k
