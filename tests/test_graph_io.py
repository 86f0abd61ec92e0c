"""Host-side containers: the input layout the kernels consume (reference graph.py conventions)."""
import os

import numpy as np
import pytest

from pecanpy_b200.graph import DenseGraph, SparseGraph

KARATE = "/root/reference/demo/karate.edg"


def _reference_style_parse(path, weighted, directed, delimiter="\t"):
    """Straightforward restatement of AdjlstGraph.read + to_csr (graph.py:217-341): dict of dicts."""
    ids, data = {}, []
    for line in open(path):
        t = line.strip().split(delimiter)
        if not line.strip():
            continue
        a, b = t[0].strip(), t[1].strip()
        w = float(t[-1]) if weighted else 1.0
        if w <= 0:
            continue
        for x in (a, b):
            if x not in ids:
                ids[x] = len(ids); data.append({})
        data[ids[a]][ids[b]] = w
        if not directed:
            data[ids[b]][ids[a]] = w
    indptr = np.zeros(len(ids) + 1, np.uint32)
    idx, dat = [], []
    for i, row in enumerate(data):
        indptr[i + 1] = indptr[i] + len(row)
        for j in sorted(row):
            idx.append(j); dat.append(row[j])
    names = [None] * len(ids)
    for k, v in ids.items():
        names[v] = k
    return names, indptr, np.array(idx, np.uint32), np.array(dat, np.float32)


@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_read_edg_matches_reference_conventions(tmp_path, weighted, directed):
    rng = np.random.default_rng(3)
    path = tmp_path / "g.edg"
    with open(path, "w") as f:
        for _ in range(400):
            a, b = rng.integers(0, 40, size=2)
            w = rng.choice([0.5, 1.5, 2.0, -1.0, 0.0])
            f.write(f"n{a}\tn{b}\t{w}\n" if weighted else f"n{a}\tn{b}\n")
    g = SparseGraph()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.read_edg(str(path), weighted=weighted, directed=directed)
    names, indptr, indices, data = _reference_style_parse(str(path), weighted, directed)
    assert g.nodes == names
    assert np.array_equal(g.indptr, indptr) and np.array_equal(g.indices, indices) and np.array_equal(g.data, data)
    assert g.indptr.dtype == np.uint32 and g.indices.dtype == np.uint32 and g.data.dtype == np.float32
    d = DenseGraph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d.read_edg(str(path), weighted=weighted, directed=directed)
    dense = np.zeros((len(names), len(names)))
    for i in range(len(names)):
        dense[i, indices[indptr[i]:indptr[i + 1]]] = data[indptr[i]:indptr[i + 1]]
    assert np.array_equal(d.data, dense) and np.array_equal(d.nonzero, dense != 0) and d.data.dtype == np.float64


@pytest.mark.skipif(not os.path.exists(KARATE), reason="reference demo file only exists in the build container")
def test_karate_matches_golden_fixture():
    g = SparseGraph()
    g.read_edg(KARATE, weighted=False, directed=False)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "karate_sparseotf_p1_q1.npz"))
    assert np.array_equal(g.indptr, z["indptr"]) and np.array_equal(g.indices, z["indices"])
    assert np.array_equal(g.data, z["data"])


def test_npz_round_trip_and_from_mat(tmp_path):
    mat = np.array([[0, 2, 0], [2, 0, 1], [0, 1, 0]], dtype=float)
    g = SparseGraph.from_mat(mat, ["a", "b", "c"])
    g.save(str(tmp_path / "g.csr.npz"))
    h = SparseGraph()
    h.read_npz(str(tmp_path / "g.csr.npz"), weighted=True)
    assert h.nodes == ["a", "b", "c"] and np.array_equal(h.indptr, g.indptr) and np.array_equal(h.data, g.data)
    u = SparseGraph()
    u.read_npz(str(tmp_path / "g.csr.npz"), weighted=False)
    assert np.all(u.data == 1.0)
    d = DenseGraph.from_mat(mat, ["a", "b", "c"])
    assert d.num_edges == 4 and d.nonzero.dtype == bool
