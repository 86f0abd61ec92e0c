"""Device-side glue: torch tensors as device buffers, libb2w.so for every kernel.

PyTorch is plumbing here (allocation, streams, ``torch.distributed``); all computation is in
the hand-written CUDA library behind ``include/b2w.h``.  There is no CPU code path: without a
CUDA device or without the built extension every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import _capi as capi

MODES = {
    "SparseOTF": capi.MODE_SPARSE_OTF,
    "PreComp": capi.MODE_PRECOMP,
    "DenseOTF": capi.MODE_DENSE_OTF,
    "FirstOrderUnweighted": capi.MODE_FIRST_ORDER_UNWEIGHTED,
    "PreCompFirstOrder": capi.MODE_PRECOMP_FIRST_ORDER,
}


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pecanpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise ValueError(f"CUDA device required, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _to_dev(arr: np.ndarray, dev: torch.device, view_dtype=None) -> torch.Tensor:
    a = np.ascontiguousarray(arr)
    if view_dtype is not None:
        a = a.view(view_dtype)
    return torch.from_numpy(a).to(dev, non_blocking=False)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class WalkEngine:
    """One graph resident on one GPU + the kernels that walk it."""

    def __init__(self, device=None):
        self.lib = capi.lib()
        self.device = _require_cuda(device)
        self.handle = C.c_void_p(None)
        self.kind = None
        self.n = 0
        self._keep = {}          # device tensors borrowed by the handle
        self.thr: Optional[torch.Tensor] = None
        self._work: dict = {}
        self.alias: Optional[Tuple[np.ndarray, torch.Tensor, torch.Tensor]] = None
        self.last_stats = None
        # per-edge index (b2w_edge_index.cu): "auto" = built on the first walk that can use it
        self.edge_index_policy = os.environ.get("B2W_EDGE_INDEX", "auto")
        self.edge_index_ms: Optional[float] = None
        self._edge_index_failed = False
        self.windex_ms: Optional[float] = None
        self._windex_key = None
        self.edge_ckpt_ms: Optional[float] = None
        self.edge_ckpt_floats = 0
        self._edge_ckpt_key = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_csr(cls, indptr, indices, data, device=None) -> "WalkEngine":
        e = cls(device)
        indptr = np.ascontiguousarray(indptr, dtype=np.uint32)
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        data = np.ascontiguousarray(data, dtype=np.float32)
        n, nnz = indptr.size - 1, int(indptr[-1])
        if indices.size != nnz or data.size != nnz:
            raise ValueError("CSR arrays are inconsistent: indptr[-1] != len(indices) / len(data)")
        padded = np.zeros(nnz + 1, dtype=np.uint32)       # one element for the unchecked choice == deg read
        padded[:nnz] = indices
        with torch.cuda.device(e.device):
            e._keep["indptr"] = _to_dev(indptr, e.device, np.int32)
            e._keep["indices"] = _to_dev(padded, e.device, np.int32)
            e._keep["data"] = _to_dev(data if nnz else np.zeros(1, np.float32), e.device)
            h = C.c_void_p(None)
            capi.check(e.lib.b2w_graph_csr_create(e.device.index, n, nnz, _ptr(e._keep["indptr"]),
                                                  _ptr(e._keep["indices"]), _ptr(e._keep["data"]), C.byref(h)),
                       "b2w_graph_csr_create")
        e.handle, e.kind, e.n = h, "csr", n
        return e

    @classmethod
    def from_dense(cls, data, nonzero, device=None) -> "WalkEngine":
        e = cls(device)
        data = np.ascontiguousarray(data, dtype=np.float64)
        nz = np.ascontiguousarray(nonzero).view(np.uint8)
        if data.ndim != 2 or data.shape[0] != data.shape[1] or nz.shape != data.shape:
            raise ValueError("dense graph: data must be square and nonzero must have the same shape")
        n = data.shape[0]
        with torch.cuda.device(e.device):
            e._keep["dense"] = _to_dev(data, e.device)
            e._keep["nonzero"] = _to_dev(nz, e.device)
            h = C.c_void_p(None)
            capi.check(e.lib.b2w_graph_dense_create(e.device.index, n, _ptr(e._keep["dense"]),
                                                    _ptr(e._keep["nonzero"]), C.byref(h)), "b2w_graph_dense_create")
        e.handle, e.kind, e.n = h, "dense", n
        return e

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.b2w_graph_destroy(self.handle)
            self.handle = C.c_void_p(None)
        self._keep.clear()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> capi.GraphInfo:
        gi = capi.GraphInfo()
        capi.check(self.lib.b2w_graph_info_get(self.handle, C.byref(gi)), "b2w_graph_info_get")
        return gi

    def set_thresholds(self, thr: Optional[np.ndarray]):
        """Attach caller-computed node2vec+ noise thresholds (float32[n]); see compute_thresholds()."""
        self.thr = None if thr is None else _to_dev(np.ascontiguousarray(thr, dtype=np.float32), self.device)

    def compute_thresholds(self, gamma: float) -> torch.Tensor:
        """node2vec+ noise thresholds on the GPU (b2w_noise_thresholds), bit-identical to the reference's NumPy
        loop (rw/sparse_rw.py:22-35, rw/dense_rw.py:11-19); the result is attached to the engine and returned
        (float32[n], device)."""
        with torch.cuda.device(self.device):
            thr = torch.empty(self.n, dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            capi.check(self.lib.b2w_noise_thresholds(self.handle, float(gamma), _ptr(thr), C.c_void_p(stream)),
                       "b2w_noise_thresholds")
            torch.cuda.current_stream(self.device).synchronize()   # consumers may run on other streams (b2w_walk_host)
        self.thr = thr
        return thr

    # ------------------------------------------------------------------ per-edge index
    @property
    def has_edge_index(self) -> bool:
        return bool(self.info().flags & capi.GRAPH_HAS_EDGE_INDEX)

    def build_edge_index(self) -> bool:
        """Build the per-edge index on the GPU (b2w_edge_index_prepare / _finish) and attach it to the handle.
        Returns False (and leaves the on-the-fly kernels in charge) when the lists do not fit 32-bit offsets or
        the device memory does not hold them."""
        if self.kind != "csr":
            raise ValueError("the edge index needs a CSR graph")
        if self.has_edge_index:
            return True
        gi = self.info()
        with torch.cuda.device(self.device):
            free, _ = torch.cuda.mem_get_info(self.device)
            rec_bytes = 16 * (gi.nnz + 1)
            wb = int(self.lib.b2w_edge_index_work_bytes(self.handle))
            if rec_bytes + wb > 0.6 * free:
                self._edge_index_failed = True
                return False
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            t0.record()
            rec = torch.empty(4 * (gi.nnz + 1), dtype=torch.int32, device=self.device)
            work = torch.empty(wb, dtype=torch.uint8, device=self.device)
            words = C.c_uint64(0)
            rc = self.lib.b2w_edge_index_prepare(self.handle, _ptr(rec), _ptr(work), wb, C.byref(words),
                                                 C.c_void_p(stream))
            if rc == capi.ERR_UNSUPPORTED or (rc == capi.OK and 4 * words.value > 0.6 * free):
                self._edge_index_failed = True
                return False
            capi.check(rc, "b2w_edge_index_prepare")
            tri = torch.empty(max(int(words.value), 1), dtype=torch.int32, device=self.device)
            capi.check(self.lib.b2w_edge_index_finish(self.handle, _ptr(rec), _ptr(tri), int(words.value), _ptr(work),
                                                      wb, C.c_void_p(stream)), "b2w_edge_index_finish")
            t1.record()
            t1.synchronize()
            self.edge_index_ms = t0.elapsed_time(t1)
        self._keep["edge_rec"], self._keep["edge_tri"] = rec, tri
        self.edge_index_words = int(words.value)
        return True

    def build_edge_ckpt(self, p: float, q: float) -> bool:
        """Checkpoints of the exact cdf for the replays of the unweighted edge-index kernel (b2w_edge_ckpt_*), for
        these p, q.  Returns False when they do not fit (the kernel then replays from the start of the row)."""
        if not self.has_edge_index:
            return False
        key = (float(p), float(q))
        if self._edge_ckpt_key == key:
            return True
        self.drop_edge_ckpt()
        gi = self.info()
        with torch.cuda.device(self.device):
            free, _ = torch.cuda.mem_get_info(self.device)
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            t0.record()
            wb = int(self.lib.b2w_edge_ckpt_work_bytes(self.handle))
            work = torch.empty(wb, dtype=torch.uint8, device=self.device)
            ckb = torch.empty(gi.num_nodes + 1, dtype=torch.int32, device=self.device)
            n_ck = C.c_uint64(0)
            rc = self.lib.b2w_edge_ckpt_prepare(self.handle, _ptr(ckb), _ptr(work), wb, C.byref(n_ck), C.c_void_p(stream))
            if rc == capi.ERR_UNSUPPORTED or (rc == capi.OK and 4 * n_ck.value > 0.4 * free):
                self._edge_ckpt_key = None
                return False
            capi.check(rc, "b2w_edge_ckpt_prepare")
            ck = torch.empty(max(int(n_ck.value), 1), dtype=torch.float32, device=self.device)
            capi.check(self.lib.b2w_edge_ckpt_finish(self.handle, float(p), float(q), _ptr(ckb), _ptr(ck), int(n_ck.value),
                                                     C.c_void_p(stream)), "b2w_edge_ckpt_finish")
            t1.record()
            t1.synchronize()
            self.edge_ckpt_ms = t0.elapsed_time(t1)
        self._keep["edge_ckb"], self._keep["edge_ckpt"] = ckb, ck
        self._edge_ckpt_key = key
        self.edge_ckpt_floats = int(n_ck.value)
        return True

    def drop_edge_ckpt(self):
        if self.kind == "csr" and self.handle:
            capi.check(self.lib.b2w_graph_clear_edge_ckpt(self.handle), "b2w_graph_clear_edge_ckpt")
        self._keep.pop("edge_ckb", None)
        self._keep.pop("edge_ckpt", None)
        self._edge_ckpt_key = None

    def build_windex(self, p: float, q: float, extend: bool = False) -> bool:
        """Build the WEIGHTED per-edge index for these bias parameters (b2w_windex_prepare / _finish) and attach it:
        SparseOTF on weighted graphs, node2vec+ and arbitrary p, q then run the lane-per-walker kernel of
        b2w_wedge.cu.  Returns False (the weight-streaming kernel stays in charge) when it does not fit."""
        if self.kind != "csr":
            raise ValueError("the weighted edge index needs a CSR graph")
        if extend and self.thr is None:
            raise ValueError("extend=True needs set_thresholds() / compute_thresholds() first")
        gi = self.info()
        key = (float(p), float(q), bool(extend), self.thr.data_ptr() if extend else 0)
        if gi.flags & capi.GRAPH_HAS_WINDEX and self._windex_key == key:
            return True
        self.drop_windex()
        with torch.cuda.device(self.device):
            free, _ = torch.cuda.mem_get_info(self.device)
            nnz = int(gi.nnz)
            wb = int(self.lib.b2w_windex_work_bytes(self.handle))
            if 44 * (nnz + 1) + wb > 0.5 * free:
                return False
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            t0.record()
            rec = torch.empty(8 * (nnz + 1), dtype=torch.int32, device=self.device)
            bw = torch.empty(max(nnz, 1), dtype=torch.float32, device=self.device)
            bq = torch.empty(max(nnz, 1), dtype=torch.float64, device=self.device)
            work = torch.empty(wb, dtype=torch.uint8, device=self.device)
            ckb = torch.empty(gi.num_nodes + 1, dtype=torch.int32, device=self.device)
            n_exc, n_ck = C.c_uint64(0), C.c_uint64(0)
            thr = _ptr(self.thr if extend else None)
            rc = self.lib.b2w_windex_prepare(self.handle, float(p), float(q), int(bool(extend)), thr, _ptr(rec), _ptr(bw),
                                             _ptr(bq), _ptr(ckb), _ptr(work), wb, C.byref(n_exc), C.byref(n_ck),
                                             C.c_void_p(stream))
            if rc == capi.ERR_UNSUPPORTED or (rc == capi.OK and 24 * n_exc.value + 4 * n_ck.value > 0.5 * free):
                return False
            capi.check(rc, "b2w_windex_prepare")
            exc = torch.empty(3 * max(int(n_exc.value), 1), dtype=torch.int64, device=self.device)
            ck = torch.empty(max(int(n_ck.value), 1), dtype=torch.float32, device=self.device)
            capi.check(self.lib.b2w_windex_finish(self.handle, float(p), float(q), int(bool(extend)), thr, _ptr(rec),
                                                  _ptr(bw), _ptr(bq), _ptr(ckb), _ptr(exc), int(n_exc.value), _ptr(ck),
                                                  int(n_ck.value), _ptr(work), wb, C.c_void_p(stream)), "b2w_windex_finish")
            t1.record()
            t1.synchronize()
            self.windex_ms = t0.elapsed_time(t1)
        self._keep["w_rec"], self._keep["w_bw"], self._keep["w_bq"], self._keep["w_exc"], self._keep["w_ckpt"] = rec, bw, bq, exc, ck
        self._keep["w_ckb"] = ckb
        self._windex_key = key
        self.windex_bytes = 32 * (nnz + 1) + 12 * nnz + 24 * int(n_exc.value) + 4 * int(n_ck.value)
        self.windex_counts = dict(exceptions=int(n_exc.value), checkpoints=int(n_ck.value))
        return True

    def drop_windex(self):
        if self.kind == "csr" and self.handle:
            capi.check(self.lib.b2w_graph_clear_windex(self.handle), "b2w_graph_clear_windex")
        for k in ("w_rec", "w_bw", "w_bq", "w_exc", "w_ckpt", "w_ckb"):
            self._keep.pop(k, None)
        self._windex_key = None

    def drop_edge_index(self):
        self.drop_edge_ckpt()
        if self.kind == "csr" and self.handle:
            capi.check(self.lib.b2w_graph_set_edge_index(self.handle, None, None, 0), "b2w_graph_set_edge_index")
        self._keep.pop("edge_rec", None)
        self._keep.pop("edge_tri", None)

    def _maybe_edge_index(self, mode: int, p: float, q: float, extend: bool, flags: int):
        """Policy "auto": build the index before the first walk whose kernel can use it (unweighted SparseOTF with
        exactly representable biases; PreComp)."""
        if self.kind != "csr" or self.edge_index_policy in ("0", "off", "never") or self._edge_index_failed:
            return
        if flags & capi.FLAG_NO_EDGE_INDEX:
            return
        if mode == capi.MODE_SPARSE_OTF:
            name = self.lib.b2w_walk_kernel_name(self.handle, mode, float(p), float(q), int(bool(extend)), int(flags)).decode()
            if name == "walk_uw_kernel" and not self.has_edge_index:
                if self.build_edge_index():
                    name = "walk_uw_edge_kernel"
            if name == "walk_uw_edge_kernel":
                if not (flags & capi.FLAG_NO_CKPT) and self._edge_ckpt_key != (float(p), float(q)):
                    self.build_edge_ckpt(p, q)
            elif name == "walk_sparse_warp_kernel" and not (flags & (capi.FLAG_NO_UNWEIGHTED_KERNEL | capi.FLAG_COOP | 0xFF00)):
                # weighted graph / node2vec+ / biases off the exact grid: the weighted index for THESE parameters
                if not extend or self.thr is not None:
                    self.build_windex(p, q, extend)
        elif mode == capi.MODE_PRECOMP and not self.has_edge_index:
            self.build_edge_index()

    def _scratch(self, key: str, nbytes: int) -> Optional[torch.Tensor]:
        if nbytes == 0:
            return None
        t = self._work.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._work[key] = t
        return t

    # ------------------------------------------------------------------ PreComp tables
    @staticmethod
    def alias_indptr(indptr: np.ndarray) -> np.ndarray:
        """[0, cumsum(deg^2)] as uint64 (reference pecanpy.py:469-476)."""
        deg = (indptr[1:] - indptr[:-1]).astype(np.uint64)
        out = np.zeros(indptr.size, dtype=np.uint64)
        out[1:] = np.cumsum(deg * deg)
        return out

    def build_alias(self, indptr: np.ndarray, p: float, q: float, extend: bool = False, first_order: bool = False,
                    packed: Optional[bool] = None):
        """Build the alias tables on the GPU and attach them to the handle.

        ``packed`` (default for PreComp): ONE table of 8-byte ``{q, j}`` entries (b2w_alias_build_packed) so that a
        draw touches one memory sector; ``self.alias`` still presents the reference's two arrays -- as strided views
        of the packed table, bit-identical to ``alias_j`` / ``alias_q`` of pecanpy.py:442-507.  The call returns
        after the build kernel has finished (the walk entry points may run on other streams)."""
        if self.kind != "csr":
            raise ValueError("alias tables need a CSR graph")
        if packed is None:
            packed = not first_order
        gi = self.info()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            wb = int(self.lib.b2w_alias_build_work_bytes(self.handle))
            work = self._scratch("alias", wb)
            aqj = None
            if first_order:
                tot = int(indptr[-1])
                aj = torch.zeros(tot + gi.max_degree + 1, dtype=torch.int32, device=self.device)
                aq = torch.zeros(tot + gi.max_degree + 1, dtype=torch.float32, device=self.device)
                capi.check(self.lib.b2w_alias_build_first_order(self.handle, _ptr(aj), _ptr(aq), _ptr(work), wb,
                                                                C.c_void_p(stream.cuda_stream)), "b2w_alias_build_first_order")
                aip_h, aip_d = None, None
            else:
                aip_h = self.alias_indptr(np.ascontiguousarray(indptr, dtype=np.uint32))
                tot = int(aip_h[-1])
                aip_d = _to_dev(aip_h, self.device, np.int64)
                if extend and self.thr is None:
                    raise ValueError("extend=True needs set_thresholds() first")
                thr = _ptr(self.thr if extend else None)
                if packed:
                    aqj = torch.zeros(tot + gi.max_degree + 1, dtype=torch.int64, device=self.device)
                    capi.check(self.lib.b2w_alias_build_packed(self.handle, float(p), float(q), int(bool(extend)), thr,
                                                               _ptr(aip_d), _ptr(aqj), _ptr(work), wb,
                                                               C.c_void_p(stream.cuda_stream)), "b2w_alias_build_packed")
                    pair = aqj.view(torch.int32).view(-1, 2)          # [:, 0] = q bits, [:, 1] = j
                    aq, aj = pair[:, 0].view(torch.float32), pair[:, 1]
                else:
                    aj = torch.zeros(tot + gi.max_degree + 1, dtype=torch.int32, device=self.device)
                    aq = torch.zeros(tot + gi.max_degree + 1, dtype=torch.float32, device=self.device)
                    capi.check(self.lib.b2w_alias_build(self.handle, float(p), float(q), int(bool(extend)), thr,
                                                        _ptr(aip_d), _ptr(aj), _ptr(aq), _ptr(work), wb,
                                                        C.c_void_p(stream.cuda_stream)), "b2w_alias_build")
            if aqj is not None:
                capi.check(self.lib.b2w_graph_set_alias_packed(self.handle, _ptr(aip_d), _ptr(aqj)), "b2w_graph_set_alias_packed")
            else:
                capi.check(self.lib.b2w_graph_set_alias(self.handle, _ptr(aip_d), _ptr(aj), _ptr(aq)), "b2w_graph_set_alias")
            stream.synchronize()          # b2w_walk_host runs on its own streams: the tables must be complete
        self._keep["alias_indptr"], self._keep["alias_j"], self._keep["alias_q"], self._keep["alias_qj"] = aip_d, aj, aq, aqj
        self.alias = (aip_h, aj[:tot], aq[:tot])
        return self.alias

    # ------------------------------------------------------------------ walking
    def walk(self, mode, p: float, q: float, start, walk_length: int, seed: int = 0, *, extend: bool = False,
             rng: int = capi.RNG_PHILOX, feed=None, row0: int = 0, flags: int = 0,
             out: Optional[torch.Tensor] = None, collect_stats: bool = True, mirrors=None) -> torch.Tensor:
        """Walk ``len(start)`` rows on this GPU; returns an int32 *view* of the uint32 matrix
        ``[n_rows, walk_length + 2]`` (layout of reference pecanpy.py:182-206) on the device.

        ``mirrors``: device addresses (ints) of the place of ``out`` in up to 7 peer matrices mapped into this process
        (dist.PeerMatrix.mirror_ptrs): the kernel stores every row there too, over NVLink, while it walks
        (b2w_walk_mirrored; raises for kernels that do not mirror)."""
        mode = MODES[mode] if isinstance(mode, str) else int(mode)
        with torch.cuda.device(self.device):
            if isinstance(start, torch.Tensor):
                d_start = start.to(self.device)
                if d_start.dtype not in (torch.int32, torch.uint32):
                    d_start = d_start.to(torch.int32)
                d_start = d_start.contiguous()
            else:
                d_start = _to_dev(np.ascontiguousarray(start, dtype=np.uint32), self.device, np.int32)
            n_rows = d_start.numel()
            ld = walk_length + 2
            if out is None:
                out = torch.empty((n_rows, ld), dtype=torch.int32, device=self.device)
            elif out.shape[0] < n_rows or out.stride(0) < ld or out.stride(1) != 1 or out.element_size() != 4:
                raise ValueError("out must be a row-major 4-byte integer tensor of at least [n_rows, L+2]")
            d_feed = None
            if rng == capi.RNG_FEED:
                d_feed = feed if isinstance(feed, torch.Tensor) else _to_dev(np.ascontiguousarray(feed, np.float64), self.device)
                if d_feed.numel() != n_rows * walk_length:
                    raise ValueError("feed must hold n_rows * walk_length doubles")
            self._maybe_edge_index(mode, p, q, bool(extend), int(flags))
            wb = int(self.lib.b2w_walk_work_bytes(self.handle, mode))
            work = self._scratch("walk", wb)
            stats_t = torch.zeros(4, dtype=torch.int64, device=self.device) if collect_stats else None
            thr = self.thr if extend else None
            if extend and thr is None and mode in (capi.MODE_SPARSE_OTF, capi.MODE_DENSE_OTF):
                raise ValueError("extend=True needs set_thresholds() first")
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if mirrors:
                if rng != capi.RNG_PHILOX:
                    raise ValueError("mirrored walks run in the Philox regime only")
                arr = (C.c_void_p * len(mirrors))(*[int(m) for m in mirrors])
                capi.check(self.lib.b2w_walk_mirrored(self.handle, mode, float(p), float(q), int(bool(extend)), _ptr(thr),
                                                      _ptr(d_start), int(row0), n_rows, int(walk_length),
                                                      int(seed) & (2 ** 64 - 1), _ptr(out), out.stride(0), _ptr(work),
                                                      wb, _ptr(stats_t), int(flags), C.c_void_p(stream), len(mirrors),
                                                      arr), "b2w_walk_mirrored")
            else:
                capi.check(self.lib.b2w_walk(self.handle, mode, float(p), float(q), int(bool(extend)), _ptr(thr),
                                             _ptr(d_start), int(row0), n_rows, int(walk_length), int(seed) & (2 ** 64 - 1),
                                             int(rng), _ptr(d_feed), _ptr(out), out.stride(0), _ptr(work), wb,
                                             _ptr(stats_t), int(flags), C.c_void_p(stream)), "b2w_walk")
            self.last_stats = stats_t
        return out[:n_rows]

    def prepare(self, mode, p: float, q: float, extend: bool = False, flags: int = 0) -> str:
        """Build whatever per-graph index ``walk`` would build for these parameters now (it is built on first use
        otherwise) and return the name of the kernel that will serve them."""
        m = MODES[mode] if isinstance(mode, str) else int(mode)
        with torch.cuda.device(self.device):
            self._maybe_edge_index(m, p, q, bool(extend), int(flags))
        return self.kernel_name(m, p, q, extend, flags)

    def kernel_name(self, mode, p: float, q: float, extend: bool = False, flags: int = 0) -> str:
        mode = MODES[mode] if isinstance(mode, str) else int(mode)
        return self.lib.b2w_walk_kernel_name(self.handle, mode, float(p), float(q), int(bool(extend)), int(flags)).decode()

    def stats(self) -> dict:
        if self.last_stats is None:
            return {}
        s = self.last_stats.cpu().tolist()
        return dict(steps=s[0], exact_replays=s[1], seq_sums=s[2], overflow_choices=s[3])

    def walk_host(self, mode, p: float, q: float, start: np.ndarray, walk_length: int, seed: int = 0, *,
                  extend: bool = False, row0: int = 0, flags: int = 0, batch_rows: int = 0,
                  out: Optional[np.ndarray] = None) -> np.ndarray:
        """End-to-end call with HOST buffers (b2w_walk_host): H2D of the start nodes, kernels and D2H of
        the walk matrix are pipelined inside the library."""
        mode = MODES[mode] if isinstance(mode, str) else int(mode)
        start = np.ascontiguousarray(start, dtype=np.uint32)
        n_rows = start.size
        if out is None:
            out = np.empty((n_rows, walk_length + 2), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.shape == (n_rows, walk_length + 2)
        st = capi.WalkStats()
        thr = self.thr if extend else None
        with torch.cuda.device(self.device):
            self._maybe_edge_index(mode, p, q, bool(extend), int(flags))
            capi.check(self.lib.b2w_walk_host(self.handle, mode, float(p), float(q), int(bool(extend)), _ptr(thr),
                                              C.c_void_p(start.ctypes.data), int(row0), n_rows, int(walk_length),
                                              int(seed) & (2 ** 64 - 1), C.c_void_p(out.ctypes.data), int(batch_rows),
                                              C.byref(st), int(flags)), "b2w_walk_host")
        self.last_host_stats = dict(steps=st.steps, exact_replays=st.exact_replays, seq_sums=st.seq_sums,
                                    overflow_choices=st.overflow_choices)
        return out

    def count_steps(self, walks: torch.Tensor, walk_length: int) -> int:
        with torch.cuda.device(self.device):
            acc = torch.zeros(1, dtype=torch.int64, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            capi.check(self.lib.b2w_count_steps(_ptr(walks), walks.shape[0], walk_length, walks.stride(0), _ptr(acc),
                                                C.c_void_p(stream)), "b2w_count_steps")
            return int(acc.item())


def new_seed() -> int:
    """Seed for random_state=None (reference: nondeterministic)."""
    return int.from_bytes(os.urandom(8), "little")
