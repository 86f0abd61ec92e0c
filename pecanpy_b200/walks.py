"""Hand-off of the walk matrix to Python consumers.

The kernels deliver ``uint32[rows, L+2]`` (layout of ``Base._random_walks``, reference pecanpy.py:182-206);
the reference turns it into ``List[List[str]]`` with one Python comprehension per walk (``_map_walk``,
pecanpy.py:103-114, ~25 us per walk -- 2-3x its own kernel time, and everything once the kernel runs on a
GPU).  Here that is one C loop (``csrc/b2w_pylists.c``) plus a lazy, restartable corpus for consumers that
only iterate (gensim's Word2Vec accepts any restartable iterable of token lists, pecanpy.py:275-288).
"""
from __future__ import annotations

import importlib.util
from typing import Iterator, List, Sequence

import numpy as np

_ext = None


def _pylists():
    """Load the in-tree CPython extension; raise loudly when it has not been built."""
    global _ext
    if _ext is None:
        from .build import pylists_path
        path = pylists_path()
        spec = importlib.util.spec_from_file_location("_b2w_pylists", path)
        if spec is None or spec.loader is None:
            raise ImportError(f"{path} not found: run `python -m pecanpy_b200.build`")
        try:
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        except (ImportError, OSError) as exc:
            raise ImportError(f"{path} not found or not loadable: run `python -m pecanpy_b200.build`") from exc
        _ext = mod
    return _ext


def _check(mat: np.ndarray) -> np.ndarray:
    mat = np.asarray(mat)
    if mat.ndim != 2 or mat.dtype.itemsize != 4 or mat.dtype.kind not in "ui" or mat.shape[1] < 3:
        raise ValueError("walk matrix must be a 2-D array of 4-byte integers with walk_length + 2 columns")
    if mat.strides[1] != 4:
        mat = np.ascontiguousarray(mat)
    return mat


def map_walks(mat: np.ndarray, nodes: Sequence[str]) -> List[List[str]]:
    """All rows of ``mat`` as lists of node ids, truncated at each row's effective length (last column):
    ``[_map_walk(row) for row in mat]`` of the reference (pecanpy.py:160) in one C call."""
    mat = _check(mat)
    return _pylists().rows_to_lists(mat, nodes, mat.shape[1] - 2, 0, mat.shape[0])


class WalkCorpus:
    """Restartable iterable over the walks of a matrix (for ``gensim.models.Word2Vec``): materialises
    ``block`` rows at a time instead of 10^7 Python lists at once."""

    def __init__(self, mat: np.ndarray, nodes: Sequence[str], block: int = 8192):
        self.mat = _check(mat)
        self.nodes = nodes if isinstance(nodes, (list, tuple)) else list(nodes)
        self.block = int(block)

    def __len__(self) -> int:
        return self.mat.shape[0]

    def __iter__(self) -> Iterator[List[str]]:
        f, L, n = _pylists().rows_to_lists, self.mat.shape[1] - 2, self.mat.shape[0]
        for r0 in range(0, n, self.block):
            yield from f(self.mat, self.nodes, L, r0, min(n, r0 + self.block))

    def __getitem__(self, i: int) -> List[str]:
        n = self.mat.shape[0]
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return _pylists().rows_to_lists(self.mat, self.nodes, self.mat.shape[1] - 2, i, i + 1)[0]
