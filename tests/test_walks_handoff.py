"""Hand-off of the walk matrix to Python (pecanpy_b200/walks.py, csrc/b2w_pylists.c) against the reference's
``_map_walk`` comprehension (pecanpy.py:103-114,160)."""
import numpy as np
import pytest

from pecanpy_b200.walks import WalkCorpus, map_walks


def _reference_map(mat, nodes):
    # pecanpy.py:103-114 applied to every row (pecanpy.py:160)
    return [[nodes[i] for i in row[:row[-1]]] for row in mat]


def _matrix(rng, rows, L, n_ids, dead_every=7):
    mat = rng.integers(0, n_ids, (rows, L + 2), dtype=np.uint32)
    mat[:, -1] = L + 1
    for r in range(0, rows, dead_every):          # dead ends: shorter effective length, zero tail
        k = int(rng.integers(1, L + 1))
        mat[r, k:L + 1] = 0
        mat[r, -1] = k
    return mat


@pytest.mark.parametrize("dtype", [np.uint32, np.int32])
def test_map_walks_equals_reference_comprehension(dtype):
    rng = np.random.default_rng(0)
    nodes = [f"n{i}" for i in range(500)]
    mat = _matrix(rng, 1000, 20, len(nodes)).astype(dtype)
    got = map_walks(mat, nodes)
    assert got == _reference_map(mat, nodes)
    assert all(isinstance(w, list) for w in got) and isinstance(got[0][0], str)
    assert got[0][0] is nodes[mat[0, 0]]          # ids are shared objects, not copies


def test_map_walks_empty_and_strided():
    nodes = [str(i) for i in range(10)]
    assert map_walks(np.zeros((0, 12), np.uint32), nodes) == []
    rng = np.random.default_rng(1)
    big = _matrix(rng, 64, 10, 10)
    view = big[::2]                               # row-strided view: no copy needed
    assert map_walks(view, nodes) == _reference_map(view, nodes)
    wide = np.zeros((8, 16), np.uint32)
    wide[:, :12] = _matrix(rng, 8, 10, 10)
    assert map_walks(wide[:, :12], nodes) == _reference_map(wide[:, :12], nodes)


def test_map_walks_rejects_bad_input():
    nodes = [str(i) for i in range(4)]
    mat = np.zeros((2, 6), np.uint32)
    mat[:, -1] = 5
    mat[1, 2] = 9                                 # node index out of range
    with pytest.raises(IndexError):
        map_walks(mat, nodes)
    mat[1, 2] = 0
    mat[0, -1] = 7                                # effective length > L + 1
    with pytest.raises(ValueError):
        map_walks(mat, nodes)
    with pytest.raises(ValueError):
        map_walks(np.zeros((2, 6), np.uint64), nodes)


def test_walk_corpus_is_restartable_and_lazy():
    rng = np.random.default_rng(2)
    nodes = [f"v{i}" for i in range(100)]
    mat = _matrix(rng, 257, 15, len(nodes))
    want = _reference_map(mat, nodes)
    corpus = WalkCorpus(mat, nodes, block=50)
    assert len(corpus) == 257
    assert list(corpus) == want
    assert list(corpus) == want                   # second epoch (gensim iterates the corpus epochs + 1 times)
    assert corpus[3] == want[3] and corpus[-1] == want[-1]
    with pytest.raises(IndexError):
        corpus[257]
