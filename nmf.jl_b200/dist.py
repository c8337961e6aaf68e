"""Multi-GPU plumbing: one process per GPU (torchrun), rows of X and W sharded across ranks, H
replicated.  torch.distributed is used only to hand the NCCL unique id
around; the data-path all-reduce runs inside libnmfb200 on the solver's own stream."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def row_shard(p: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [start, stop) of rank `rank`: the first p % world ranks get one extra row."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(p, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def h_row_ownership(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [start, stop) of H' (= columns of H) that `rank` OWNS in the row-sharded tensor-core MultUpdate(:mse) solve: it alone
    applies the multiplicative ratio to them (csrc/tc_shard.cuh).  H' is cut into tiles of 128 rows (n < 128: one tile), each rank
    owns ceil(tiles / world) consecutive tiles; trailing ranks may own nothing.  Python mirror of nmfb200_shard_geometry."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    tile = 128 if n >= 128 else -(-max(n, 8) // 8) * 8
    tiles = -(-n // tile)
    tpo = -(-tiles // world)
    return min(n, rank * tpo * tile), min(n, (rank + 1) * tpo * tile)


def shard_geometry(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """The library's own answer (host-only C entry point nmfb200_shard_geometry): (own_row0, own_row1, tile_rows)."""
    import ctypes

    from . import _lib
    a, b, t = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    st = _lib.load().nmfb200_shard_geometry(n, world, rank, ctypes.byref(a), ctypes.byref(b), ctypes.byref(t))
    if st != _lib.OK:
        raise ValueError("nmfb200_shard_geometry: invalid arguments")
    return a.value, b.value, t.value


def init_comm(session, group=None) -> None:
    """Create the library's NCCL communicator over the ranks of `group` (default: WORLD)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    box = [type(session).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    session.comm_init(rank, world, box[0])


def solve_sharded(alg, session, X_rows: np.ndarray, W_rows: np.ndarray, H: np.ndarray):
    """NMF.solve!(alg, X, W, H) with X = vcat(X_rows of every rank), W likewise, H identical on all ranks.
    Returns this rank's Result (W = its rows, H = the full replicated factor; niters/converged/objvalue
    are global and identical on every rank)."""
    session.set_X(X_rows)
    return session.solve(alg, W_rows, H)
