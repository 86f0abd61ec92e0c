"""Multi-GPU plumbing: shard walkers, replicate the graph, one all-gather at the end.

Walkers never interact (reference pecanpy.py:189-206 writes only row ``i``) and the graph is read-only,
so the path shards by rows of the (host-shuffled) start array.  Each rank walks the contiguous block
``[rank * R, (rank + 1) * R)`` straight into its slice of a full-size device buffer and a single
``all_gather_into_tensor`` (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic)
collects the matrix.  Philox is keyed by the GLOBAL row index, so the result is independent of the
number of ranks.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_rows(total_rows: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Rows [lo, hi) of this rank and the padded block size R (= ceil(total / world))."""
    R = (total_rows + world - 1) // world
    lo = min(total_rows, rank * R)
    hi = min(total_rows, (rank + 1) * R)
    return lo, hi, R


def sharded_walks(walk_block: Callable[[int, int, torch.Tensor], None], total_rows: int, row_len: int,
                  device, group=None) -> torch.Tensor:
    """Run ``walk_block(lo, hi, out_block)`` for this rank's rows and all-gather the full matrix.

    ``out_block`` is this rank's ``[R, row_len]`` int32 slice of the full buffer (rows past ``hi - lo``
    are padding and stay zero).  Returns the ``[total_rows, row_len]`` matrix (a view of the buffer)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi, R = shard_rows(total_rows, world, rank)
    full = torch.zeros((R * world, row_len), dtype=torch.int32, device=device)
    mine = full[rank * R:(rank + 1) * R]
    if hi > lo:
        walk_block(lo, hi, mine)
    if world > 1:
        dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1), group=group)
    return full[:total_rows]


def simulate_walks_distributed(engine, mode, p: float, q: float, start, walk_length: int, seed: int, *,
                               extend: bool = False, flags: int = 0, group=None) -> torch.Tensor:
    """All-rank walk of the (identical, host-shuffled) ``start`` array: every rank holds a replica of the graph
    in ``engine`` (a :class:`pecanpy_b200.engine.WalkEngine` on its own GPU), walks its row block and receives
    the full ``int32[len(start), walk_length + 2]`` matrix (bit-identical for any number of ranks)."""
    import numpy as np
    start = np.ascontiguousarray(start, dtype=np.uint32)

    def walk_block(lo: int, hi: int, out_block: torch.Tensor) -> None:
        engine.walk(mode, p, q, start[lo:hi], walk_length, seed=seed, extend=extend, row0=lo, out=out_block,
                    flags=flags, collect_stats=False)

    return sharded_walks(walk_block, start.size, walk_length + 2, engine.device, group=group)
