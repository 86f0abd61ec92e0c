"""Edge list -> CSR on the GPU: ``b2w_csr_from_edges`` behind a NumPy-in / NumPy-out call.

The reference builds a dict of dicts edge by edge and sorts every row in Python (graph.py:160-341, minutes at 10^7
edges).  The text parsing and the first-appearance node numbering stay on the host (graph._parse_edge_list); the
de-duplication ("a later line wins"), the symmetrisation and the row sort are one stable radix sort on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _capi as capi
from .engine import _ptr, _require_cuda, _to_dev


def csr_from_edges_device(num_nodes: int, src, dst, weight: Optional[np.ndarray], directed: bool,
                          device=None, return_tensors: bool = False):
    """CSR (indptr uint32[n+1], indices uint32[nnz], data float32[nnz]) of the edges ``(src[e], dst[e], weight[e])``
    given in file order; ``weight=None`` means unweighted (all ones).  Same arrays as the reference's
    ``AdjlstGraph.read`` + ``to_csr`` for the same lines."""
    lib = capi.lib()
    dev = _require_cuda(device)
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    m = int(src.size)
    if dst.size != m or (weight is not None and np.size(weight) != m):
        raise ValueError("src, dst and weight must have the same length")
    if m and (int(src.max()) >= num_nodes or int(dst.max()) >= num_nodes):
        raise ValueError("edge endpoint out of range")
    cap = m if directed else 2 * m
    with torch.cuda.device(dev):
        d_src = _to_dev(src, dev, np.int32) if m else None
        d_dst = _to_dev(dst, dev, np.int32) if m else None
        d_w = _to_dev(np.ascontiguousarray(weight, dtype=np.float64), dev) if (weight is not None and m) else None
        indptr = torch.zeros(num_nodes + 1, dtype=torch.int32, device=dev)
        indices = torch.zeros(cap + 1, dtype=torch.int32, device=dev)      # + the pad element b2w_graph_csr_create wants
        data = torch.zeros(max(cap, 1), dtype=torch.float32, device=dev)
        wb = int(lib.b2w_csr_from_edges_work_bytes(num_nodes, m, int(bool(directed))))
        if wb == 0:
            raise capi.B2WError("b2w_csr_from_edges_work_bytes failed")
        work = torch.empty(wb, dtype=torch.uint8, device=dev)
        nnz = C.c_uint64(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        capi.check(lib.b2w_csr_from_edges(dev.index, num_nodes, m, _ptr(d_src), _ptr(d_dst), _ptr(d_w),
                                          int(bool(directed)), _ptr(indptr), _ptr(indices), _ptr(data), C.byref(nnz),
                                          _ptr(work), wb, C.c_void_p(stream)), "b2w_csr_from_edges")
        k = int(nnz.value)
        if return_tensors:
            return indptr, indices[:k + 1], data[:k], k
        return (indptr.cpu().numpy().view(np.uint32), indices[:k].cpu().numpy().view(np.uint32),
                data[:k].cpu().numpy())
