"""Multi-GPU host logic: one process per GPU (torch.distributed), work sharded by
independent electrode channel (the reference sorts one channel per
process, src/hmmsort.jl:79-83) with NO data-path collective.  Only results and
timings are gathered/reduced.  The same functions run under the `gloo` backend on
CPU (tests/test_sharding_gloo.py)."""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `n_items` for `rank`; the first
    n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    q, r = divmod(n_items, world)
    start = rank * q + min(rank, r)
    return start, start + q + (1 if rank < r else 0)


def shard_time(T: int, world: int, rank: int, halo: int, align: int = 256) -> Tuple[int, int, int, int]:
    """Contiguous time span [start, stop) for `rank` (boundaries aligned to `align`
    samples) plus the halo-extended span [lo, hi) a speculative start / look-ahead
    needs (config 5 style time sharding)."""
    per = -(-T // world)
    per = -(-per // align) * align
    start = min(T, rank * per)
    stop = min(T, start + per)
    return start, stop, max(0, start - halo), min(T, stop + halo)


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def world_info() -> Tuple[int, int]:
    d = _dist()
    return (d.get_world_size(), d.get_rank()) if d else (1, 0)


def all_reduce_max(value: float, device=None) -> float:
    """Max over ranks (timings are reported as the slowest rank's)."""
    d = _dist()
    if d is None:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def decode_channels_sharded(decode_fn: Callable, Y: np.ndarray, models: Sequence, gather: bool = True):
    """Shard the columns of Y [T x C] (one model per column) over the ranks, call
    `decode_fn(Y_shard, models_shard) -> (x [T x c], ll [c])` on this rank's
    shard, and (optionally) gather the per-channel results on every rank.

    Returns (x_local, ll_local, (start, stop)) when gather=False, else
    (x_all [T x C], ll_all [C])."""
    world, rank = world_info()
    C = Y.shape[1]
    a, b = shard_range(C, world, rank)
    x_loc, ll_loc = decode_fn(Y[:, a:b], list(models[a:b])) if b > a else (
        np.empty((Y.shape[0], 0), dtype=np.int16, order="F"), np.empty(0))
    if not gather:
        return x_loc, ll_loc, (a, b)
    d = _dist()
    if d is None:
        return x_loc, ll_loc
    parts: List = [None] * world
    d.all_gather_object(parts, (a, b, np.ascontiguousarray(x_loc.T), np.asarray(ll_loc)))
    x_all = np.empty((Y.shape[0], C), dtype=np.int16, order="F")
    ll_all = np.empty(C, dtype=np.float64)
    for pa, pb, px, pl in parts:
        x_all[:, pa:pb] = px.T
        ll_all[pa:pb] = pl
    return x_all, ll_all
