"""Seeded synthetic graphs of the BASELINE.json shapes.  Host NumPy; deterministic."""
from __future__ import annotations

import numpy as np


def _undirected_csr(n: int, lo: np.ndarray, hi: np.ndarray, w=None):
    """Symmetrise unique (lo < hi) pairs and return sorted-row CSR (uint32, uint32, float32)."""
    rows = np.concatenate([lo, hi])
    cols = np.concatenate([hi, lo])
    ww = None if w is None else np.concatenate([w, w])
    key = rows.astype(np.int64) * n + cols.astype(np.int64)
    order = np.argsort(key, kind="stable")
    rows, cols = rows[order], cols[order]
    indptr = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    data = np.ones(cols.size, dtype=np.float32) if ww is None else ww[order].astype(np.float32)
    return indptr, cols.astype(np.uint32), data


def _dedupe_pairs(n: int, a: np.ndarray, b: np.ndarray, m: int, rng):
    keep = a != b
    a, b = a[keep], b[keep]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    keys = np.unique(lo.astype(np.int64) * n + hi.astype(np.int64))
    if keys.size > m:
        keys = keys[rng.permutation(keys.size)[:m]]
    return (keys // n).astype(np.int64), (keys % n).astype(np.int64)


def erdos_renyi_csr(n: int, m: int, seed: int = 0, weighted: bool = False):
    """Config #2 / #4 generator: ceil(1.05 m) uniform pairs, drop loops, dedupe, keep m, symmetrise."""
    rng = np.random.default_rng(seed)
    k = int(np.ceil(1.05 * m))
    a = rng.integers(0, n, size=k)
    b = rng.integers(0, n, size=k)
    lo, hi = _dedupe_pairs(n, a, b, m, rng)
    w = None
    if weighted:
        w = np.float32(0.01) + np.float32(0.99) * rng.random(lo.size, dtype=np.float32)
    return _undirected_csr(n, lo, hi, w)


def power_law_csr(n: int, m: int, seed: int = 1, weighted: bool = False):
    """Config #3 generator: Chung-Lu, theta_i = (i + 65)^(-2/3) (degree exponent 2.5)."""
    rng = np.random.default_rng(seed)
    theta = (np.arange(n, dtype=np.float64) + 65.0) ** (-2.0 / 3.0)
    cdf = np.cumsum(theta)
    cdf /= cdf[-1]
    k = int(np.ceil(1.1 * m))
    a = np.searchsorted(cdf, rng.random(k), side="right")
    b = np.searchsorted(cdf, rng.random(k), side="right")
    np.minimum(a, n - 1, out=a)
    np.minimum(b, n - 1, out=b)
    lo, hi = _dedupe_pairs(n, a, b, m, rng)
    w = None
    if weighted:
        w = np.float32(0.01) + np.float32(0.99) * rng.random(lo.size, dtype=np.float32)
    return _undirected_csr(n, lo, hi, w)


def dense_weighted(n: int, density: float = 0.3, seed: int = 3):
    """Config #5 generator: symmetric float64 matrix, `density` of the upper triangle non-zero."""
    rng = np.random.default_rng(seed)
    mat = np.zeros((n, n), dtype=np.float64)
    blk = 2048
    for r0 in range(0, n, blk):
        r1 = min(n, r0 + blk)
        mask = rng.random((r1 - r0, n), dtype=np.float32) < density
        w = (np.float32(0.01) + np.float32(0.99) * rng.random((r1 - r0, n), dtype=np.float32)).astype(np.float64)
        cols = np.arange(n)[None, :]
        rows = np.arange(r0, r1)[:, None]
        upper = mask & (cols > rows)
        mat[r0:r1] = np.where(upper, w, 0.0)
    mat = mat + mat.T
    return mat, mat != 0


def shuffled_start(num_nodes: int, num_walks: int, seed) -> np.ndarray:
    """The reference's start array (pecanpy.py:135-141): NumPy legacy global generator."""
    nodes = np.array(range(num_nodes), dtype=np.uint32)
    start = np.concatenate([nodes] * num_walks)
    np.random.seed(seed)
    np.random.shuffle(start)
    return start


def sparse_otf_algorithmic_bytes(indptr: np.ndarray, walks: np.ndarray) -> int:
    """per step j>=2: 20 + 8 d_cur + 4 d_prev; step 1: 12 + 8 d_cur; + 4 per walker."""
    deg = (indptr[1:].astype(np.int64) - indptr[:-1].astype(np.int64))
    L = walks.shape[1] - 2
    eff = walks[:, -1].astype(np.int64)          # number of valid entries
    total = 4 * walks.shape[0]
    cols = np.arange(L + 2)[None, :]
    valid_cur = cols[:, :L] < (eff[:, None] - 1)     # entry c is a `cur` of step c+1 iff c + 1 <= eff - 1
    dcur = deg[walks[:, :L]] * valid_cur
    total += int((8 * dcur).sum())
    nsteps = (eff - 1)
    total += int((12 * (nsteps >= 1)).sum() + (20 * np.maximum(nsteps - 1, 0)).sum())
    # prev of step j (j >= 2) is entry j-2
    valid_prev = cols[:, :L] < (eff[:, None] - 2)
    total += int((4 * deg[walks[:, :L]] * valid_prev).sum())
    return total
