"""Single-process multi-GPU walks behind the drop-in classes: ``model.devices = ["cuda:0", "cuda:1", ...]`` makes
``simulate_walks`` / ``simulate_walks_array`` shard the (host-shuffled) start array over those GPUs from THIS process
(C ABI: ``b2w_walk_multi`` -- one host thread per device, graph replicated, rows delivered into one host matrix).
The reference is single-process, so its drop-in has to be too; multi-process jobs (torchrun, one rank per GPU, NCCL
all-gather of the device blocks) live in ``dist.py``."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi as capi
from .engine import MODES


def replicas(model):
    """One WalkEngine per device of ``model.devices`` (built on first use, cached in the model), each with the
    thresholds / alias tables the mode needs."""
    devs = [torch.device(d) for d in model.devices]
    first = model.engine                                   # the model's own engine serves its device
    out = []
    for d in devs:
        if d.index is None:
            d = torch.device("cuda", torch.cuda.current_device())
        if d == first.device:
            out.append(first)
            continue
        key = str(d)
        eng = model._replicas.get(key)
        if eng is None:
            eng = model._make_engine(device=d)
            if model.extend and model._MODE in ("SparseOTF", "DenseOTF", "PreComp"):
                eng.compute_thresholds(model.gamma)
            if model._MODE == "PreComp":
                eng.build_alias(model.indptr, model.p, model.q, extend=bool(model.extend))
            elif model._MODE == "PreCompFirstOrder":
                eng.build_alias(model.indptr, 1.0, 1.0, first_order=True)
            model._replicas[key] = eng
        out.append(eng)
    return out


def walk_host_engines(engines, mode, p: float, q: float, extend: bool, start: np.ndarray, walk_length: int, seed: int,
                      out: np.ndarray = None, flags: int = 0) -> np.ndarray:
    """``b2w_walk_multi`` over ready replicas (one WalkEngine per device): rows sharded in contiguous blocks, one host
    thread per GPU inside the library, ONE host matrix ``uint32[len(start), walk_length + 2]``."""
    mode = MODES[mode] if isinstance(mode, str) else int(mode)
    start = np.ascontiguousarray(start, dtype=np.uint32)
    n_rows = start.size
    if out is None:
        out = np.empty((n_rows, walk_length + 2), dtype=np.uint32)
    assert out.dtype == np.uint32 and out.flags.c_contiguous and out.shape == (n_rows, walk_length + 2)
    extend = bool(extend)
    for e in engines:                                      # index policy runs per replica
        with torch.cuda.device(e.device):
            e._maybe_edge_index(mode, p, q, extend, flags)
    n = len(engines)
    handles = (C.c_void_p * n)(*[e.handle for e in engines])
    thr = (C.c_void_p * n)(*[C.c_void_p(e.thr.data_ptr()) if (extend and e.thr is not None) else None for e in engines])
    st = capi.WalkStats()
    capi.check(capi.lib().b2w_walk_multi(n, handles, mode, float(p), float(q), int(extend), thr,
                                         C.c_void_p(start.ctypes.data), n_rows, int(walk_length),
                                         int(seed) & (2 ** 64 - 1), C.c_void_p(out.ctypes.data), 0, C.byref(st), int(flags)),
               "b2w_walk_multi")
    walk_host_engines.last_stats = dict(steps=st.steps, exact_replays=st.exact_replays, seq_sums=st.seq_sums,
                                        overflow_choices=st.overflow_choices, devices=[str(e.device) for e in engines])
    return out


def walk_host_multi(model, start: np.ndarray, walk_length: int, seed: int, out: np.ndarray = None) -> np.ndarray:
    engines = replicas(model)
    out = walk_host_engines(engines, model._MODE, model.p, model.q, bool(model.extend), start, walk_length, seed, out)
    model.last_multi_stats = walk_host_engines.last_stats
    return out
