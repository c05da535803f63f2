"""The fused clear-sky column kernels (lw_column.cu, sw_column.cu) against the staged kernels they replace, and their
invariances.  Both paths evaluate the same band formulas (lw_bands.cuh / sw_bands.cuh) and the same recurrences; they
differ in where the sums over g-points are cut (per task and then in task order, against per warp of g-points), so they agree
to rounding, not bit for bit.  Each path is checked against the oracle by the rest of the GPU suite (the fused one by
default, the staged one wherever a test asks for clouds, aerosols, idrv = 1 or stage capture)."""
import numpy as np
import pytest

from mima_b200.columns import make_columns
from test_gpu_parity import LW_OUT, SW_OUT, _check_outputs

pytestmark = pytest.mark.gpu


def _both(gpu, fn, c, **kw):
    fused = fn(c, **kw)
    gpu.set_option("lw_fused", 0)
    gpu.set_option("sw_fused", 0)
    try:
        staged = fn(c, **kw)
    finally:
        gpu.set_option("lw_fused", 1)
        gpu.set_option("sw_fused", 1)
    return fused, staged


@pytest.mark.parametrize("kw", [dict(nlon=64, nlat=8, night=True),
                                dict(nlon=37, nlat=3, night=True),                      # 111 columns: a ragged last tile
                                dict(nlon=5, nlat=1),                                   # less than one tile
                                dict(nlon=32, nlat=4, co2_ppmv=1560.0, ozone="file", secondary_gases=True)])
def test_fused_and_staged_kernels_agree(gpu, oracle, kw):
    c = make_columns("T42L40", **kw)
    for fn, names, ref in ((gpu.lw_from_columns, LW_OUT, oracle.rrtmg_lw(c)), (gpu.sw_from_columns, SW_OUT, oracle.rrtmg_sw(c))):
        fused, staged = _both(gpu, fn, c)
        _check_outputs(fused, ref, names)
        _check_outputs(staged, ref, names)
        for a, b, n in zip(fused, staged, names):
            if "hr" in n:        # flux differences over layers of down to 0.1 hPa: K/day, against the 1e-4 of north_star
                assert np.max(np.abs(a - b)) <= 1e-7, n
            else:
                assert np.max(np.abs(a - b)) <= 1e-11 * max(np.abs(b).max(), 1e-300), n
            if c.ncol >= 256:
                assert not np.array_equal(a, b), n                   # two different kernels did run


def test_fused_lw_with_aerosol_optical_depth(gpu, oracle):
    """The AER instantiation of lw_column_kernel (tauaer passed, iaer = 10 forced: rad.nomcica:442, :514-519)."""
    c = make_columns("T85L40", nlon=48, nlat=3)
    rng = np.random.default_rng(5)
    taer = np.asfortranarray(rng.uniform(0.0, 0.08, (c.ncol, c.nlay, 16)))
    c.emis = np.asfortranarray(rng.uniform(0.8, 1.0, (c.ncol, 16)))
    fused, staged = _both(gpu, gpu.lw_from_columns, c, tauaer=taer)
    ref = oracle.rrtmg_lw(c, tauaer=taer)
    _check_outputs(fused, ref, LW_OUT)
    _check_outputs(staged, ref, LW_OUT)


def test_block_shape_and_pass_size_do_not_change_a_bit(gpu):
    """A column's arithmetic does not depend on which warp, block or device pass it lands in: the results are bitwise
    the same for both block shapes of the column kernels and for any pass size (option chunk)."""
    c = make_columns("T170L60", nlon=96, nlat=5, night=True)      # 480 columns, 15 tiles
    base = gpu.lw_from_columns(c), gpu.sw_from_columns(c)
    try:
        for key, val in (("col_warps", 8), ("col_warps", 16), ("chunk", 64), ("chunk", 100), ("chunk", 479)):
            gpu.set_option(key, val)
            got = gpu.lw_from_columns(c), gpu.sw_from_columns(c)
            gpu.set_option(key, 0)
            for x, y in zip(base[0] + base[1], got[0] + got[1]):
                assert np.array_equal(x, y), (key, val)
    finally:
        gpu.set_option("col_warps", 0)
        gpu.set_option("chunk", 0)


def test_night_tiles_and_mixed_tiles(gpu, oracle):
    """Tiles of 32 columns that are all night (the column kernel skips them), all day, and mixed (night lanes are carried along
    and discarded): zeros where the sun is down, parity elsewhere."""
    c = make_columns("T42L40", nlon=64, nlat=3)
    c.coszen[:40] = 0.0                     # tile 0 all night, tile 1 mixed
    c.coszen[100:101] = 0.0                 # one night column inside a sunlit tile
    c.coszen[-3:] = 0.5e-10                 # below the threshold of rad.nomcica:502
    got = gpu.sw_from_columns(c)
    _check_outputs(got, oracle.rrtmg_sw(c), SW_OUT)
    night = c.coszen < 1e-10
    assert night.sum() == 44
    for a in got:
        assert (a[night] == 0).all() and np.isfinite(a).all()
    assert (got[1][~night, -1] > 0).all()   # sunlit columns do receive flux at the top
