"""Multi-process multi-GPU plumbing (one rank per GPU, ``torch.distributed``): shard walkers, replicate the graph,
all-gather the walk matrix.

Walkers never interact (reference pecanpy.py:189-206 writes only row ``i``) and the graph is read-only, so the path
shards by rows of the (host-shuffled) start array.  The rows are cut into ``NB`` batches; batch ``b`` of rank ``r`` is
the contiguous row block ``[(b * world + r) * B, ... + B)``, walked straight into its place in a full-size device
buffer.  As soon as a batch is walked it is all-gathered (in place: the gathered batch is the contiguous block
``[b * world * B, (b + 1) * world * B)``) on a side stream while the next batch is being walked, so the NVLink
transfer hides behind the kernel instead of trailing it (NCCL on GPUs; gloo, staged through the host, in the CPU
tests of the host logic).  Philox is keyed by the GLOBAL row index, so the result is independent of the number of
ranks and of ``NB``.  ``NB = 1`` is the single all-gather at the end.

Single-process hosts (PecanPy itself is one) use ``multi.py`` / ``b2w_walk_multi`` instead.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_rows(total_rows: int, world: int, rank: int, batches: int = 1) -> Tuple[List[Tuple[int, int]], int]:
    """This rank's row blocks ``[(lo, hi), ...]`` (one per batch; ``hi - lo`` may be 0 at the tail) and the block
    size ``B = ceil(total / (world * batches))``."""
    batches = max(1, int(batches))
    B = (total_rows + world * batches - 1) // (world * batches)
    blocks = []
    for b in range(batches):
        lo = min(total_rows, (b * world + rank) * B)
        hi = min(total_rows, lo + B)
        blocks.append((lo, hi))
    return blocks, B


def _all_gather_block(seg: torch.Tensor, mine: torch.Tensor, group) -> None:
    """All-gather ``mine`` (this rank's rows of the batch) into ``seg`` (the batch, rank-major)."""
    if dist.get_backend(group) == "nccl" or not seg.is_cuda:
        dist.all_gather_into_tensor(seg.view(-1), mine.reshape(-1), group=group)
    else:                                   # gloo has no device collectives: stage through the host (tests)
        h = torch.empty(seg.shape, dtype=seg.dtype)
        dist.all_gather_into_tensor(h.view(-1), mine.reshape(-1).cpu(), group=group)
        seg.copy_(h)


def sharded_walks(walk_block: Callable[[int, int, torch.Tensor], None], total_rows: int, row_len: int,
                  device, group=None, batches: int = 1) -> torch.Tensor:
    """Run ``walk_block(lo, hi, out_block)`` for each of this rank's row blocks and all-gather the full matrix,
    batch by batch, overlapped with the walk of the next batch.  Returns the ``[total_rows, row_len]`` matrix
    (a view of the padded buffer; rows past ``total_rows`` stay zero)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    blocks, B = shard_rows(total_rows, world, rank, batches)
    full = torch.zeros((B * world * len(blocks), row_len), dtype=torch.int32, device=device)
    on_gpu = full.is_cuda
    cur = torch.cuda.current_stream(full.device) if on_gpu else None
    comm = torch.cuda.Stream(device=full.device) if (on_gpu and world > 1) else None
    for b, (lo, hi) in enumerate(blocks):
        slot = (b * world + rank) * B
        if hi > lo:
            walk_block(lo, hi, full[slot:slot + (hi - lo)])
        if world > 1:
            seg, mine = full[b * world * B:(b + 1) * world * B], full[slot:slot + B]
            if comm is not None:
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    _all_gather_block(seg, mine, group)
            else:
                _all_gather_block(seg, mine, group)
    if comm is not None:
        cur.wait_stream(comm)
    return full[:total_rows]


def simulate_walks_distributed(engine, mode, p: float, q: float, start, walk_length: int, seed: int, *,
                               extend: bool = False, flags: int = 0, group=None, batches: int = 1) -> torch.Tensor:
    """All-rank walk of the (identical, host-shuffled) ``start`` array: every rank holds a replica of the graph
    in ``engine`` (a :class:`pecanpy_b200.engine.WalkEngine` on its own GPU), walks its row blocks and receives
    the full ``int32[len(start), walk_length + 2]`` matrix (bit-identical for any number of ranks / batches)."""
    import numpy as np
    start = np.ascontiguousarray(start, dtype=np.uint32)

    def walk_block(lo: int, hi: int, out_block: torch.Tensor) -> None:
        engine.walk(mode, p, q, start[lo:hi], walk_length, seed=seed, extend=extend, row0=lo, out=out_block,
                    flags=flags, collect_stats=False)

    return sharded_walks(walk_block, start.size, walk_length + 2, engine.device, group=group, batches=batches)
