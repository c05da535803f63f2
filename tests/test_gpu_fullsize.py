"""Full-size batches on the GPU against the oracle on a column sample: the whole T170L60 batch (two device
passes of 65536 columns, and the default single pass), the whole T85L40 batch and a two-pass T341L80 slice go through the C ABI in
ONE block, and sampled latitude rows -- both polar rows, the rows either side of the 65536-column pass boundary, random rows -- are
compared with the oracle (itself bit-identical to the translated reference, tests/test_ref_translation.py)."""
import numpy as np
import pytest

from mima_b200.columns import RESOLUTIONS, make_columns

pytestmark = pytest.mark.gpu

FLUX_RTOL, HR_ATOL = 1e-6, 1e-4          # north_star
LW_OUT = ("uflx", "dflx", "hr", "uflxc", "dflxc", "hrc")
SW_OUT = ("swuflx", "swdflx", "swhr", "swuflxc", "swdflxc", "swhrc")


def _check(got, ref, names, idx):
    for g, n in zip(got, names):
        g, o = g[idx], ref[n]
        assert np.isfinite(g).all(), n
        if "hr" in n:
            assert np.max(np.abs(g - o)) < HR_ATOL, (n, float(np.max(np.abs(g - o))))
            assert np.max(np.abs(g - o)) < 1e-7, (n, float(np.max(np.abs(g - o))))
        else:
            r = np.max(np.abs(g - o) / np.maximum(np.abs(o), 1e-6 * np.abs(o).max()))
            assert r < FLUX_RTOL and r < 1e-9, (n, float(r))


@pytest.mark.parametrize("res, rows, nsample_rows, chunk",
                         [("T170L60", None, 16, 0), ("T170L60", None, 8, 65536), ("T85L40", None, 16, 0), ("T341L80", (192, 320), 8, 65536)],
                         ids=["T170L60-whole-one-pass", "T170L60-whole-two-passes", "T85L40-whole", "T341L80-two-passes"])
def test_full_batch_sample_against_oracle(gpu, oracle, res, rows, nsample_rows, chunk):
    nlon, nlat, nlay = RESOLUTIONS[res]
    cols = make_columns(res, lat_rows=rows, night=(res == "T85L40"))
    ncol = cols.ncol
    nrow = ncol // nlon
    rng = np.random.default_rng(5)
    picks = {0, nrow - 1}
    if chunk and ncol > chunk:               # rows either side of the device-pass boundary
        b = chunk // nlon
        picks |= {b - 1, b}
    while len(picks) < nsample_rows:
        picks.add(int(rng.integers(0, nrow)))
    idx = np.concatenate([np.arange(j * nlon, (j + 1) * nlon) for j in sorted(picks)])
    gpu.set_option("host_chunk", ncol)       # the whole batch as one block: one device pass (default, up to 131072 columns) or two
    gpu.set_option("chunk", chunk)
    try:
        lw = gpu.lw_from_columns(cols)
        sw = gpu.sw_from_columns(cols)
    finally:
        gpu.set_option("host_chunk", 0)
        gpu.set_option("chunk", 0)
    sample = cols.take(idx)
    _check(lw, oracle.rrtmg_lw(sample), LW_OUT, idx)
    _check(sw, oracle.rrtmg_sw(sample), SW_OUT, idx)
    # the host pipeline in its default blocks gives the same bits
    lw2 = gpu.lw_from_columns(cols)
    assert all(np.array_equal(a, b) for a, b in zip(lw, lw2))
