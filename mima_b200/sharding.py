"""Latitude-row sharding of a radiation call over ranks (one rank <-> one GPU), and the timing reduction.

MiMA decomposes grid-point work by latitude rows (`layout = (/1,npes/)`, src/atmos_spectral/tools/spec_mpp.f90:42-44)
and `run_rrtmg` flattens (lon, lat) with longitude fastest (rrtm_radiation.f90:652), so the block of rows
owned by a rank is a contiguous column range: sharding is a pointer offset, there is no collective on the
data path.  The only communication is the max-reduction of the per-rank device time.
"""
from __future__ import annotations


def lat_row_block(nlat: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [j0, j1) owned by `rank`.  nlat must divide evenly, as in spec_mpp.f90:49."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    if nlat % world:
        raise ValueError(f"{nlat} latitude rows do not divide over {world} ranks (spec_mpp.f90:49)")
    per = nlat // world
    return rank * per, (rank + 1) * per


def column_range(nlon: int, nlat: int, world: int, rank: int) -> tuple[int, int]:
    """Columns [c0, c1) of the flattened (lon fastest) grid owned by `rank`."""
    j0, j1 = lat_row_block(nlat, world, rank)
    return j0 * nlon, j1 * nlon


def max_over_ranks(x: float, device=None) -> float:
    """max of x over the ranks of the default process group (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_rate(columns_per_rank: int, world: int, ms_max: float) -> float:
    """Whole-job columns/s: all ranks' columns over the slowest rank's time."""
    return columns_per_rank * world / (ms_max * 1e-3)
