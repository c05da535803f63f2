"""N > 1 path on CPU: two gloo ranks each compute their latitude-row block (with the oracle standing in for
the GPU, tests may use it) and the gathered result equals the unsharded call bit for bit; the timing
reduction is the max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mima_b200 import sharding
from mima_b200.columns import make_columns


def test_row_blocks_partition_the_grid():
    for nlat, world in ((64, 1), (64, 2), (256, 8), (128, 4)):
        cover = []
        for r in range(world):
            c0, c1 = sharding.column_range(128, nlat, world, r)
            cover += list(range(c0, c1, 128))
        assert cover == list(range(0, 128 * nlat, 128))
    with pytest.raises(ValueError):
        sharding.lat_row_block(64, 3, 0)
    with pytest.raises(ValueError):
        sharding.lat_row_block(64, 2, 2)
    assert sharding.aggregate_rate(1000, 4, 2.0) == 2.0e6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nlon, nlat, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.pyoracle import Oracle
        j0, j1 = sharding.lat_row_block(nlat, world, rank)
        blk = make_columns("T42L40", nlon=nlon, nlat=nlat, lat_rows=(j0, j1), night=True)
        orc = Oracle()
        lw, sw = orc.rrtmg_lw(blk, nthreads=1), orc.rrtmg_sw(blk, nthreads=1)
        mine = torch.from_numpy(np.ascontiguousarray(np.concatenate([lw["uflx"], lw["hr"], sw["swdflx"], sw["swhr"]], axis=1)))
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)                       # test-side gather only; the product path has no collective
        t = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            q.put((torch.cat(parts, 0).numpy(), t))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_unsharded_call(oracle):
    nlon, nlat, world = 8, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nlon, nlat, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = make_columns("T42L40", nlon=nlon, nlat=nlat, night=True)
    lw, sw = oracle.rrtmg_lw(full, nthreads=1), oracle.rrtmg_sw(full, nthreads=1)
    ref = np.concatenate([lw["uflx"], lw["hr"], sw["swdflx"], sw["swhr"]], axis=1)
    assert np.array_equal(got, ref)
    assert tmax == 11.0
