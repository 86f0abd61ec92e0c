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


class _DevBuffer:
    """A raw device allocation seen by torch (``torch.as_tensor`` on the CUDA array interface, zero copy)."""

    def __init__(self, ptr: int, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
        self._owner = owner                    # keeps the allocation alive as long as a tensor refers to it


class PeerMatrix:
    """The full-size walk matrix of this rank, allocated so that the other ranks of a one-node job can map it
    (C ABI: b2w_shared_alloc / b2w_shared_open), for an all-gather by the copy engines: after a batch has been walked,
    ``push`` copies this rank's rows straight into the same rows of every peer's matrix with device-to-device DMA over
    NVLink (b2w_push_rows: plain cudaMemcpyAsync on a local side stream).  It runs no kernel, so it overlaps with a
    walk kernel that fills the chip without taking SMs from it.  ``finish`` waits for this rank's copies and for every
    peer's (a barrier): then ``self.full`` holds all rows.  ``close`` unmaps and frees (collective).

    The same mapped matrices serve the FUSED gather: ``mirror_ptrs(row)`` are the addresses the walk kernel stores its
    rows to itself (engine.walk(..., mirrors=...) -> b2w_walk_mirrored); ``finish`` is then just the barrier.
    ``ok`` is False on every rank if any rank could not allocate or map (no peer access): fall back to NCCL."""

    def __init__(self, rows: int, row_len: int, device, group=None):
        import ctypes as C
        from . import _capi as capi
        self.lib, self.capi, self.C = capi.lib(), capi, C
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.row_bytes = 4 * row_len
        nbytes = max(1, rows * self.row_bytes)
        ptr = C.c_void_p(None)
        handle = C.create_string_buffer(64)
        self.ok, self.error = True, ""
        self._ptr, self.full, self._mapped = None, None, [None] * self.world
        # every step below is collective: a rank whose allocation or mapping fails keeps taking part and the ranks
        # agree on `ok` at the end (a rank that raised here would leave the others waiting in a collective)
        if self.lib.b2w_shared_alloc(self.device.index, nbytes, C.byref(ptr), handle) != capi.OK:
            self.ok, self.error = False, self.lib.b2w_last_error().decode("utf-8", "replace")
        else:
            self._ptr = ptr.value
        with torch.cuda.device(self.device):
            if self._ptr is not None:
                self.full = torch.as_tensor(_DevBuffer(self._ptr, (rows, row_len), self), device=self.device)
                self.full.zero_()
                torch.cuda.current_stream(self.device).synchronize()
            self.stream = torch.cuda.Stream(device=self.device)
            # one stream per peer: the copies to different peers run on different copy engines at the same time
            self.streams = [torch.cuda.Stream(device=self.device) if r != self.rank else self.stream
                            for r in range(self.world)]
            self._stream_ptrs = (C.c_void_p * self.world)(*[s.cuda_stream for s in self.streams])
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw if self._ptr is not None else None, group=group)
        for r, h in enumerate(handles):
            if r == self.rank:
                self._mapped[r] = self._ptr
            elif h is None or self._ptr is None:
                self.ok = False
            else:
                p = C.c_void_p(None)
                if self.lib.b2w_shared_open(self.device.index, h, C.byref(p)) != capi.OK:
                    self.ok, self.error = False, self.lib.b2w_last_error().decode("utf-8", "replace")
                    continue                                   # (no peer access on this box: the caller falls back)
                self._mapped[r] = p.value
        self._peers = (C.c_void_p * self.world)(*self._mapped)
        self._token = torch.zeros(1, device=self.device)
        flag = torch.tensor([1 if self.ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # usable only if every rank mapped every peer
        self.ok = bool(int(flag.item()))

    def push(self, lo: int, hi: int) -> None:
        """Rows [lo, hi) of the local matrix -> the same rows of every peer, ordered after the work queued so far on
        the current stream (the walk of that block)."""
        if hi <= lo or self.world == 1:
            return
        cur = torch.cuda.current_stream(self.device)
        for r, st in enumerate(self.streams):
            if r != self.rank:
                st.wait_stream(cur)
        self.capi.check(self.lib.b2w_push_rows_streams(self.device.index, self._peers, self.world, self.rank, lo, hi - lo,
                                                       self.row_bytes, self._stream_ptrs), "b2w_push_rows_streams")

    def mirror_ptrs(self, row: int) -> list:
        """Addresses of row ``row`` in every PEER's matrix (for engine.walk(..., mirrors=...): the walk kernel stores
        its rows there itself, the all-gather fused into the kernel); ``finish`` completes it."""
        return [p + row * self.row_bytes for r, p in enumerate(self._mapped) if r != self.rank]

    def finish(self) -> None:
        for r, st in enumerate(self.streams):                  # my copies have landed in the peers
            if r != self.rank:
                st.synchronize()
        dist.all_reduce(self._token, group=self.group)         # ... and everybody else's in mine
        torch.cuda.current_stream(self.device).synchronize()

    def close(self) -> None:
        """Collective (also on a rank whose allocation failed): unmap, barrier, free."""
        if self._mapped is None:
            return
        torch.cuda.synchronize(self.device)
        for r, p in enumerate(self._mapped):
            if r != self.rank and p:
                self.capi.check(self.lib.b2w_shared_close(self.device.index, self.C.c_void_p(p)), "b2w_shared_close")
        self._mapped = None
        dist.barrier(self.group)                               # nobody frees before everybody has unmapped
        self.full = None
        if self._ptr is not None:
            self.capi.check(self.lib.b2w_shared_free(self.device.index, self.C.c_void_p(self._ptr)), "b2w_shared_free")
        self._ptr = None


def _all_gather_block(seg: torch.Tensor, mine: torch.Tensor, group) -> None:
    """All-gather ``mine`` (this rank's rows of the batch) into ``seg`` (the batch, rank-major)."""
    if dist.get_backend(group) == "nccl" or not seg.is_cuda:
        dist.all_gather_into_tensor(seg.view(-1), mine.reshape(-1), group=group)
    else:                                   # gloo has no device collectives: stage through the host (tests)
        h = torch.empty(seg.shape, dtype=seg.dtype)
        dist.all_gather_into_tensor(h.view(-1), mine.reshape(-1).cpu(), group=group)
        seg.copy_(h)


def sharded_walks(walk_block: Callable[[int, int, torch.Tensor], None], total_rows: int, row_len: int,
                  device, group=None, batches: int = 1, gather: str = "nccl") -> torch.Tensor:
    """Run ``walk_block(lo, hi, out_block)`` for each of this rank's row blocks and all-gather the full matrix,
    batch by batch, overlapped with the walk of the next batch.  Returns the ``[total_rows, row_len]`` matrix
    (a view of the padded buffer; rows past ``total_rows`` stay zero)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    blocks, B = shard_rows(total_rows, world, rank, batches)
    on_gpu = torch.device(device).type == "cuda"
    peer = PeerMatrix(B * world * len(blocks), row_len, device, group) if (gather in ("push", "mirror") and on_gpu and world > 1) else None
    if peer is not None and not peer.ok:       # no peer access between these GPUs: NCCL does it
        peer.close()
        peer = None
    mirror = peer is not None and gather == "mirror" and world <= 8
    if peer is not None:
        full = peer.full
    else:
        full = torch.zeros((B * world * len(blocks), row_len), dtype=torch.int32, device=device)
    cur = torch.cuda.current_stream(full.device) if on_gpu else None
    comm = torch.cuda.Stream(device=full.device) if (on_gpu and world > 1) else None
    for b, (lo, hi) in enumerate(blocks):
        slot = (b * world + rank) * B
        if hi > lo:
            if mirror:
                walk_block(lo, hi, full[slot:slot + (hi - lo)], mirrors=peer.mirror_ptrs(slot))
            else:
                walk_block(lo, hi, full[slot:slot + (hi - lo)])
        if mirror:
            pass                                # the kernel has stored the rows in the peers' matrices itself
        elif peer is not None:
            peer.push(slot, slot + (hi - lo))
        elif world > 1:
            seg, mine = full[b * world * B:(b + 1) * world * B], full[slot:slot + B]
            if comm is not None:
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    _all_gather_block(seg, mine, group)
            else:
                _all_gather_block(seg, mine, group)
    if peer is not None:
        peer.finish()
        out = full[:total_rows].clone()    # the shared allocation is released here (collective); hand back a copy
        del full
        peer.close()
        return out
    if comm is not None:
        cur.wait_stream(comm)
    return full[:total_rows]


def simulate_walks_distributed(engine, mode, p: float, q: float, start, walk_length: int, seed: int, *,
                               extend: bool = False, flags: int = 0, group=None, batches: int = 1,
                               gather: str = "nccl") -> torch.Tensor:
    """All-rank walk of the (identical, host-shuffled) ``start`` array: every rank holds a replica of the graph
    in ``engine`` (a :class:`pecanpy_b200.engine.WalkEngine` on its own GPU), walks its row blocks and receives
    the full ``int32[len(start), walk_length + 2]`` matrix (bit-identical for any number of ranks / batches)."""
    import numpy as np
    start = np.ascontiguousarray(start, dtype=np.uint32)

    def walk_block(lo: int, hi: int, out_block: torch.Tensor, mirrors=None) -> None:
        engine.walk(mode, p, q, start[lo:hi], walk_length, seed=seed, extend=extend, row0=lo, out=out_block,
                    flags=flags, collect_stats=False, mirrors=mirrors)

    if gather == "mirror":                      # only the unweighted SparseOTF edge-index kernel mirrors its rows
        if engine.prepare(mode, p, q, extend, flags) != "walk_uw_edge_kernel":
            gather = "push"

    return sharded_walks(walk_block, start.size, walk_length + 2, engine.device, group=group, batches=batches,
                         gather=gather)
