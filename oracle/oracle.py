"""ctypes front end of the CPU parity oracle (oracle/walk_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``pecanpy_b200/`` imports this module.

The functions mirror the reference entry points they restate (file:line relative to
/root/reference/src/pecanpy/):

* :func:`walk_csr` / :func:`walk_dense`  -> ``Base._random_walks`` (pecanpy.py:164-210) with the
  ``move_forward`` of SparseOTF (:522-561), PreComp (:384-440), DenseOTF (:576-614),
  FirstOrderUnweighted (:299-309) and PreCompFirstOrder (:319-334)
* :func:`alias_build`  -> ``PreComp.preprocess_transition_probs`` (pecanpy.py:442-507)
* :func:`sparse_probs` / :func:`dense_probs` -> ``get_(extended_)normalized_probs``
  (rw/sparse_rw.py:51-130, rw/dense_rw.py:34-118)
* :func:`noise_thresholds_csr` / ``_dense`` -> ``get_noise_thresholds`` (plain NumPy in the
  reference too: rw/sparse_rw.py:22-35, rw/dense_rw.py:11-19)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

RNG_WORDS, RNG_FEED, RNG_PHILOX = 0, 1, 2
MODE_SPARSE_OTF, MODE_PRECOMP, MODE_DENSE_OTF, MODE_FIRST_ORDER_UNWEIGHTED, MODE_PRECOMP_FIRST_ORDER = range(5)
MODES = {
    "SparseOTF": MODE_SPARSE_OTF,
    "PreComp": MODE_PRECOMP,
    "DenseOTF": MODE_DENSE_OTF,
    "FirstOrderUnweighted": MODE_FIRST_ORDER_UNWEIGHTED,
    "PreCompFirstOrder": MODE_PRECOMP_FIRST_ORDER,
}


def build(force: bool = False) -> str:
    """Compile walk_oracle.c with the system gcc (the image's $CC has no libgomp)."""
    src = os.path.join(_HERE, "walk_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-std=c11",
           "-shared", "-o", _LIB_PATH, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_walk_csr.restype = C.c_int
        _lib.orc_walk_dense.restype = C.c_int
        _lib.orc_sparse_probs.restype = C.c_uint32
        _lib.orc_dense_probs.restype = C.c_uint32
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pad_indices(indices: np.ndarray) -> np.ndarray:
    """One trailing element for the reference's unchecked indices[indptr[cur]+deg] read."""
    out = np.zeros(indices.size + 1, dtype=np.uint32)
    out[:-1] = indices
    return out


def mt_words(seed: int, n: int) -> np.ndarray:
    """Raw MT19937 32-bit outputs of Numba's generator after ``np.random.seed(seed)``.

    Numba's seeded stream equals ``np.random.RandomState(seed)``; full-range
    uint32 ``randint`` returns the raw tempered words without rejection.
    """
    return np.random.RandomState(seed).randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)


def mt_uniform_feed(seed: int, n_rows: int, walk_length: int) -> np.ndarray:
    """U[i, j-1] for the MT-replay regime R1 (each walker consumes exactly L doubles)."""
    return np.random.RandomState(seed).random_sample(n_rows * walk_length).reshape(n_rows, walk_length)


def shuffled_start(num_nodes: int, num_walks: int, seed) -> np.ndarray:
    """pecanpy.py:135-141, verbatim semantics (NumPy legacy global generator)."""
    nodes = np.array(range(num_nodes), dtype=np.uint32)
    start = np.concatenate([nodes] * num_walks)
    np.random.seed(seed)
    np.random.shuffle(start)
    return start


def noise_thresholds_csr(indptr, data, gamma: float) -> np.ndarray:
    n = indptr.size - 1
    thr = np.zeros(n, dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(n):
            row = data[indptr[i]:indptr[i + 1]]
            thr[i] = row.mean() + gamma * row.std()
    return np.maximum(thr, 0)


def noise_thresholds_dense(data, nonzero, gamma: float) -> np.ndarray:
    n = data.shape[0]
    thr = np.zeros(n, dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(n):
            w = data[i, nonzero[i]]
            thr[i] = w.mean() + gamma * w.std()
    return np.maximum(thr, 0)


def alias_indptr(indptr: np.ndarray) -> np.ndarray:
    deg = (indptr[1:] - indptr[:-1]).astype(np.uint64)
    out = np.zeros(indptr.size, dtype=np.uint64)
    out[1:] = np.cumsum(deg * deg)
    return out


def alias_build(indptr, indices, data, p, q, extend=False, thr=None):
    n = indptr.size - 1
    aip = alias_indptr(indptr)
    tot = int(aip[-1])
    j = np.zeros(tot + 1, dtype=np.uint32)
    qq = np.zeros(tot + 1, dtype=np.float32)
    idx = pad_indices(np.ascontiguousarray(indices, dtype=np.uint32))
    lib().orc_alias_build(C.c_uint32(n), _p(indptr), _p(idx), _p(data), C.c_double(p), C.c_double(q),
                          C.c_int(int(extend)), _p(thr), _p(aip), _p(j), _p(qq))
    return aip, j[:tot], qq[:tot]


def alias_build_first_order(indptr, indices, data):
    n = indptr.size - 1
    nnz = int(indptr[-1])
    j = np.zeros(nnz + 1, dtype=np.uint32)
    qq = np.zeros(nnz + 1, dtype=np.float32)
    idx = pad_indices(np.ascontiguousarray(indices, dtype=np.uint32))
    lib().orc_alias_build_first_order(C.c_uint32(n), _p(indptr), _p(idx), _p(data), _p(j), _p(qq))
    return j[:nnz], qq[:nnz]


def sparse_probs(indptr, indices, data, p, q, cur, prev=None, extend=False, thr=None):
    n = indptr.size - 1
    deg = int(indptr[cur + 1] - indptr[cur])
    out = np.zeros(max(deg, 1), dtype=np.float32)
    idx = pad_indices(np.ascontiguousarray(indices, dtype=np.uint32))
    lib().orc_sparse_probs(C.c_uint32(n), _p(indptr), _p(idx), _p(data), C.c_double(p), C.c_double(q),
                           C.c_int(int(extend)), _p(thr), C.c_uint32(cur),
                           C.c_int64(-1 if prev is None else int(prev)), _p(out))
    return out[:deg]


def dense_probs(data, nonzero, p, q, cur, prev=None, extend=False, thr=None):
    n = data.shape[0]
    out = np.zeros(n, dtype=np.float64)
    cols = np.zeros(n, dtype=np.uint32)
    nz8 = np.ascontiguousarray(nonzero).view(np.uint8)
    d = lib().orc_dense_probs(C.c_uint32(n), _p(data), _p(nz8), C.c_double(p), C.c_double(q),
                              C.c_int(int(extend)), _p(thr), C.c_uint32(cur),
                              C.c_int64(-1 if prev is None else int(prev)), _p(out), _p(cols))
    return out[:d], cols[:d]


def _rng_args(rng, seed, words, feed):
    if rng == RNG_WORDS:
        assert words is not None
        return words, None
    if rng == RNG_FEED:
        assert feed is not None and feed.dtype == np.float64 and feed.flags.c_contiguous
        return None, feed
    return None, None


def walk_csr(mode, indptr, indices, data, p, q, start, walk_length, *, extend=False, thr=None,
             alias=None, rng=RNG_PHILOX, seed=0, words=None, feed=None, row0=0, nthreads=0,
             return_words_used=False):
    """Walk matrix ``uint32[len(start), L+2]`` in the layout of pecanpy.py:182-187."""
    if isinstance(mode, str):
        mode = MODES[mode]
    indptr = np.ascontiguousarray(indptr, dtype=np.uint32)
    idx = pad_indices(np.ascontiguousarray(indices, dtype=np.uint32))
    data = np.ascontiguousarray(data, dtype=np.float32)
    start = np.ascontiguousarray(start, dtype=np.uint32)
    n = indptr.size - 1
    out = np.zeros((start.size, walk_length + 2), dtype=np.uint32)
    words, feed = _rng_args(rng, seed, words, feed)
    aip = aj = aq = None
    if alias is not None:
        aip, aj, aq = alias
        # slack for the reference's unchecked table offset after a failed neighbour search
        aj = np.concatenate([aj, np.zeros(1, np.uint32)])
        aq = np.concatenate([aq, np.zeros(1, np.float32)])
    used = C.c_uint64(0)
    rc = lib().orc_walk_csr(C.c_int(mode), C.c_uint32(n), _p(indptr), _p(idx), _p(data),
                            C.c_double(p), C.c_double(q), C.c_int(int(extend)), _p(thr),
                            _p(aip), _p(aj), _p(aq), _p(start), C.c_uint64(row0),
                            C.c_uint64(start.size), C.c_uint32(walk_length), C.c_int(rng),
                            C.c_uint64(seed), _p(words), C.c_uint64(0 if words is None else words.size),
                            _p(feed), _p(out), C.c_int(nthreads), C.byref(used))
    if rc != 0:
        raise RuntimeError("oracle: random word stream exhausted")
    return (out, used.value) if return_words_used else out


def walk_dense(data, nonzero, p, q, start, walk_length, *, extend=False, thr=None, rng=RNG_PHILOX,
               seed=0, words=None, feed=None, row0=0, nthreads=0):
    data = np.ascontiguousarray(data, dtype=np.float64)
    nz8 = np.ascontiguousarray(nonzero).view(np.uint8)
    start = np.ascontiguousarray(start, dtype=np.uint32)
    n = data.shape[0]
    out = np.zeros((start.size, walk_length + 2), dtype=np.uint32)
    words, feed = _rng_args(rng, seed, words, feed)
    used = C.c_uint64(0)
    rc = lib().orc_walk_dense(C.c_uint32(n), _p(data), _p(nz8), C.c_double(p), C.c_double(q),
                              C.c_int(int(extend)), _p(thr), _p(start), C.c_uint64(row0),
                              C.c_uint64(start.size), C.c_uint32(walk_length), C.c_int(rng),
                              C.c_uint64(seed), _p(words), C.c_uint64(0 if words is None else words.size),
                              _p(feed), _p(out), C.c_int(nthreads), C.byref(used))
    if rc != 0:
        raise RuntimeError("oracle: random word stream exhausted")
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())
