"""ctypes binding of libb2w.so (include/b2w.h).  Thin by design: argument marshalling only.

The CUDA library is the product; if it is missing this module raises -- there is no CPU
fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2W_LIBRARY: developer knob for A/B runs of tuning variants built by tools/build_variants.py (same ABI, same
# CUDA code base compiled with a different -D); unset in normal use.
LIB_PATH = os.environ.get("B2W_LIBRARY") or os.path.join(_HERE, "lib", "libb2w.so")

OK = 0
ERR_UNSUPPORTED = -4
ERR_NOMEM = -5
MODE_SPARSE_OTF, MODE_PRECOMP, MODE_DENSE_OTF, MODE_FIRST_ORDER_UNWEIGHTED, MODE_PRECOMP_FIRST_ORDER = range(5)
RNG_PHILOX, RNG_FEED = 0, 1
FLAG_FORCE_EXACT_REPLAY = 0x1
FLAG_NO_FILTER_STATS = 0x2
FLAG_THREAD_PER_WALKER = 0x4
FLAG_NO_UNWEIGHTED_KERNEL = 0x8
FLAG_NO_TMA = 0x10
FLAG_COOP = 0x20
FLAG_NO_EDGE_INDEX = 0x40
FLAG_NO_CKPT = 0x80
FLAG_OFFEDGE_WARP = 0x1000000
FLAG_OFFEDGE_LANE = 0x2000000


def FLAG_GROUP(n: int) -> int:
    return (n & 0xFF) << 8

GRAPH_CSR, GRAPH_DENSE, GRAPH_UNWEIGHTED, GRAPH_HAS_ALIAS, GRAPH_HAS_EDGE_INDEX = 0x1, 0x2, 0x4, 0x8, 0x10
GRAPH_HAS_WINDEX = 0x20

EXPORTS = [
    "b2w_version", "b2w_last_error", "b2w_device_count", "b2w_graph_csr_create", "b2w_graph_dense_create",
    "b2w_graph_info_get", "b2w_graph_destroy", "b2w_alias_build_work_bytes", "b2w_alias_build",
    "b2w_alias_build_first_order", "b2w_graph_set_alias", "b2w_walk_work_bytes", "b2w_walk", "b2w_walk_host",
    "b2w_count_steps", "b2w_philox_selftest", "b2w_walk_kernel_name", "b2w_noise_thresholds", "b2w_csr_from_edges_work_bytes", "b2w_csr_from_edges",
    "b2w_edgelist_parse", "b2w_edgelist_fetch", "b2w_edgelist_free",
    "b2w_edge_ckpt_work_bytes", "b2w_edge_ckpt_prepare", "b2w_edge_ckpt_finish", "b2w_graph_clear_edge_ckpt",
    "b2w_windex_work_bytes", "b2w_windex_prepare", "b2w_windex_finish", "b2w_graph_clear_windex",
    "b2w_shared_alloc", "b2w_shared_free", "b2w_shared_open", "b2w_shared_close", "b2w_push_rows", "b2w_push_rows_streams",
    "b2w_walk_multi", "b2w_walk_mirrored", "b2w_allgather_rows", "b2w_shuffled_start", "b2w_alias_build_packed", "b2w_graph_set_alias_packed",
    "b2w_edge_index_work_bytes", "b2w_edge_index_prepare", "b2w_edge_index_finish", "b2w_graph_set_edge_index",
]


class GraphInfo(C.Structure):
    _fields_ = [("num_nodes", C.c_uint32), ("nnz", C.c_uint64), ("max_degree", C.c_uint32), ("flags", C.c_uint32)]


class WalkStats(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("exact_replays", C.c_uint64), ("seq_sums", C.c_uint64),
                ("overflow_choices", C.c_uint64)]


class B2WError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libb2w.so; raise loudly (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run "
            "`python -m pecanpy_b200.build` (needs nvcc; cross-compiles for sm_100a without a GPU).")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, dbl, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double, C.c_size_t
    L.b2w_version.restype = i32
    L.b2w_last_error.restype = C.c_char_p
    L.b2w_device_count.argtypes = [C.POINTER(i32)]
    L.b2w_graph_csr_create.argtypes = [i32, u32, u64, vp, vp, vp, C.POINTER(vp)]
    L.b2w_graph_dense_create.argtypes = [i32, u32, vp, vp, C.POINTER(vp)]
    L.b2w_graph_info_get.argtypes = [vp, C.POINTER(GraphInfo)]
    L.b2w_graph_destroy.argtypes = [vp]
    L.b2w_graph_destroy.restype = None
    L.b2w_alias_build_work_bytes.argtypes = [vp]
    L.b2w_alias_build_work_bytes.restype = sz
    L.b2w_alias_build.argtypes = [vp, dbl, dbl, i32, vp, vp, vp, vp, vp, sz, vp]
    L.b2w_alias_build_first_order.argtypes = [vp, vp, vp, vp, sz, vp]
    L.b2w_graph_set_alias.argtypes = [vp, vp, vp, vp]
    L.b2w_alias_build_packed.argtypes = [vp, dbl, dbl, i32, vp, vp, vp, vp, sz, vp]
    L.b2w_graph_set_alias_packed.argtypes = [vp, vp, vp]
    L.b2w_walk_work_bytes.argtypes = [vp, i32]
    L.b2w_walk_work_bytes.restype = sz
    L.b2w_walk.argtypes = [vp, i32, dbl, dbl, i32, vp, vp, u64, u64, u32, u64, i32, vp, vp, u64, vp, sz, vp, u32, vp]
    L.b2w_walk_mirrored.argtypes = [vp, i32, dbl, dbl, i32, vp, vp, u64, u64, u32, u64, vp, u64, vp, sz, vp, u32, vp, i32,
                                    C.POINTER(vp)]
    L.b2w_walk_host.argtypes = [vp, i32, dbl, dbl, i32, vp, vp, u64, u64, u32, u64, vp, u64, C.POINTER(WalkStats), u32]
    L.b2w_walk_multi.argtypes = [i32, C.POINTER(vp), i32, dbl, dbl, i32, C.POINTER(vp), vp, u64, u32, u64, vp, u64,
                                 C.POINTER(WalkStats), u32]
    L.b2w_edge_ckpt_work_bytes.argtypes = [vp]
    L.b2w_edge_ckpt_work_bytes.restype = sz
    L.b2w_edge_ckpt_prepare.argtypes = [vp, vp, vp, sz, C.POINTER(u64), vp]
    L.b2w_edge_ckpt_finish.argtypes = [vp, dbl, dbl, vp, vp, u64, vp]
    L.b2w_graph_clear_edge_ckpt.argtypes = [vp]
    L.b2w_windex_work_bytes.argtypes = [vp]
    L.b2w_windex_work_bytes.restype = sz
    L.b2w_windex_prepare.argtypes = [vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, sz, C.POINTER(u64), C.POINTER(u64), vp]
    L.b2w_windex_finish.argtypes = [vp, dbl, dbl, i32, vp, vp, vp, vp, vp, vp, u64, vp, u64, vp, sz, vp]
    L.b2w_graph_clear_windex.argtypes = [vp]
    L.b2w_shared_alloc.argtypes = [i32, sz, C.POINTER(vp), C.c_char_p]
    L.b2w_shared_free.argtypes = [i32, vp]
    L.b2w_shared_open.argtypes = [i32, C.c_char_p, C.POINTER(vp)]
    L.b2w_shared_close.argtypes = [i32, vp]
    L.b2w_push_rows.argtypes = [i32, C.POINTER(vp), i32, i32, u64, u64, u64, vp]
    L.b2w_push_rows_streams.argtypes = [i32, C.POINTER(vp), i32, i32, u64, u64, u64, C.POINTER(vp)]
    L.b2w_allgather_rows.argtypes = [i32, C.POINTER(vp), i32, i32, u64, u32, vp]
    L.b2w_shuffled_start.argtypes = [u32, u32, vp, C.POINTER(C.c_int32), vp]
    L.b2w_walk_kernel_name.argtypes = [vp, i32, dbl, dbl, i32, u32]
    L.b2w_walk_kernel_name.restype = C.c_char_p
    L.b2w_noise_thresholds.argtypes = [vp, dbl, vp, vp]
    L.b2w_csr_from_edges_work_bytes.argtypes = [u32, u64, i32]
    L.b2w_csr_from_edges_work_bytes.restype = sz
    L.b2w_csr_from_edges.argtypes = [i32, u32, u64, vp, vp, vp, i32, vp, vp, vp, C.POINTER(u64), vp, sz, vp]
    L.b2w_edgelist_parse.argtypes = [C.c_char_p, i32, C.c_char_p, C.POINTER(vp), C.POINTER(u64), C.POINTER(u32),
                                     C.POINTER(u64), C.POINTER(u64)]
    L.b2w_edgelist_fetch.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(u32)]
    L.b2w_edgelist_free.argtypes = [vp]
    L.b2w_edgelist_free.restype = None
    L.b2w_edge_index_work_bytes.argtypes = [vp]
    L.b2w_edge_index_work_bytes.restype = sz
    L.b2w_edge_index_prepare.argtypes = [vp, vp, vp, sz, C.POINTER(u64), vp]
    L.b2w_edge_index_finish.argtypes = [vp, vp, vp, u64, vp, sz, vp]
    L.b2w_graph_set_edge_index.argtypes = [vp, vp, vp, u64]
    L.b2w_count_steps.argtypes = [vp, u64, u32, u64, vp, vp]
    L.b2w_philox_selftest.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    for name in EXPORTS:
        getattr(L, name)   # AttributeError here == the library does not export what b2w.h declares
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != OK:
        msg = lib().b2w_last_error().decode("utf-8", "replace")
        raise B2WError(f"{what or 'libb2w'} failed ({rc}): {msg}")
