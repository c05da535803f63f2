"""CPU checks of the compile-time tables the clear-sky column kernels are built on (mima_b200/csrc/rrtmg_dev.cuh,
lw_column.cu, sw_column.cu): the (band, first g-point, count) tasks must tile every band exactly, in band order; counts and
offsets must be even (optical depths and table rows move in 16-byte pairs); the row stride of a task slice must be an odd
number of 16-byte units that holds the band's widest task (bank-conflict-free LDS.128 of distinct rows); the launch orders
must be permutations; and the state masks must only name fields that exist.  A wrong table here shows up on the GPU as wrong
fluxes at best -- these run everywhere."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mima_b200", "csrc")


def _read(name):
    return open(os.path.join(CSRC, name)).read()


def _tasks(fn):
    """the ColTask table of lw_task / sw_task in rrtmg_dev.cuh"""
    txt = _read("rrtmg_dev.cuh")
    body = txt[txt.index("constexpr ColTask %s(int t)" % fn):]
    body = body[body.index("tk[COL_NTASK] = {"):body.index("};")]
    return [tuple(int(x) for x in m) for m in re.findall(r"\{(\d+),\s*(\d+),\s*(\d+)\}", body)]


def _int_table(txt, decl):
    body = txt[txt.index(decl):]
    body = body[body.index("{") + 1:body.index("}")]
    return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", body)]


def _slice_rs(nmax):
    # col_slice_rs in rrtmg_dev.cuh
    return 10 if nmax > 6 else (6 if nmax > 2 else 2)


def _check_tasks(tasks, ng):
    assert len(tasks) == 23
    assert tasks == sorted(tasks), "tasks must come in (band, first g-point) order: the scratch slot of a task is its index"
    for band, n_band in enumerate(ng):
        mine = [(g0, n) for b, g0, n in tasks if b == band]
        assert mine, band
        pos = 0
        for g0, n in mine:
            assert g0 == pos, (band, g0, pos)              # no gap, no overlap
            assert n % 2 == 0 and g0 % 2 == 0 and 2 <= n <= 8
            pos += n
        assert pos == n_band, (band, pos, n_band)
        rs = _slice_rs(max(n for _, n in mine))
        assert rs >= max(n for _, n in mine)
        assert (rs * 8) % 16 == 0 and ((rs * 8) // 16) % 2 == 1, "slice stride: an odd number of 16-byte units"
    assert sum(ng) == sum(n for _, _, n in tasks)


def test_lw_tasks_tile_the_bands():
    ng = _int_table(_read("lw_bands.cuh"), "constexpr int ng[16]")
    assert sum(ng) == 140
    tasks = _tasks("lw_task")
    _check_tasks(tasks, ng)
    # lw_column's scratch: slot of a task = (first g-point of the task in the 140-vector) / 2 + task index, LW_CSLOT in total
    first = [sum(ng[:b]) + g0 for b, g0, _ in tasks]
    slots = [f // 2 + t for t, f in enumerate(first)]
    width = [n // 2 + 1 for _, _, n in tasks]
    for t in range(22):
        assert slots[t] + width[t] == slots[t + 1], t
    assert slots[-1] + width[-1] == 140 // 2 + 23
    assert "LW_CSLOT = NGPTLW / 2 + LW_NTASK" in _read("rrtmg_dev.cuh")


def test_sw_tasks_tile_the_bands():
    ng = _int_table(_read("sw_bands.cuh"), "constexpr int ng[14]")
    assert sum(ng) == 112
    tasks = _tasks("sw_task")
    _check_tasks(tasks, ng)
    assert all(n <= 6 for _, _, n in tasks)
    g0 = _int_table(_read("sw_column.cu"), "constexpr int g0[14]")
    assert g0 == [sum(ng[:b]) for b in range(14)]
    # scratch slots: three per g-point and one per task, SW_NSLOT in total
    slots = [3 * (g0[b] + o) + t for t, (b, o, _) in enumerate(tasks)]
    width = [3 * n + 1 for _, _, n in tasks]
    for t in range(22):
        assert slots[t] + width[t] == slots[t + 1], t
    assert slots[-1] + width[-1] == 3 * 112 + 23


def test_launch_orders_are_permutations():
    for unit, name in (("lw_column.cu", "c_task_order[LW_NTASK]"), ("sw_column.cu", "c_sw_task_order[SW_NTASK]")):
        order = _int_table(_read(unit), name + " = ")
        assert sorted(order) == list(range(23)), unit
    # longest first: the first task launched has the most g-points of its code
    lw, sw = _tasks("lw_task"), _tasks("sw_task")
    assert lw[_int_table(_read("lw_column.cu"), "c_task_order[LW_NTASK] = ")[0]][2] == max(n for _, _, n in lw)
    assert sw[_int_table(_read("sw_column.cu"), "c_sw_task_order[SW_NTASK] = ")[0]][2] == max(n for _, _, n in sw)


def test_largest_slices_fit_the_shared_memory_of_an_sm():
    """Rows of the largest band tables (RRTMG_LW bands 3-5: absa 9x5x13, absb 5x5x47, the eta-resolved minor species, self,
    foreign, Planck fractions) times the slice stride must fit the 227 KB a block may use."""
    rows_band3 = 585 + 1175 + 171 + 95 + 10 + 4 + 9 + 5
    assert rows_band3 * _slice_rs(8) * 8 + 1024 <= 227 * 1024
