"""b2w_csr_from_edges (device CSR build) against the UNMODIFIED reference's ingest (graph.py:160-341: dict of dicts
filled line by line -- later lines overwrite -- rows emitted with sorted columns): the fixtures
tests/golden/graph_*.npz hold edge-list texts and the arrays the reference's AdjlstGraph.read / to_csr made of them
(oracle/gen_golden_graph.py).  The random multigraphs further down use a restatement of add_edge + to_csr on
integer endpoints as an additional, larger check."""
import glob
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPH_FIXTURES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "graph_*.npz"))
                        if "literals" not in f)


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_read_edg_on_device_equals_reference(tmp_path, name):
    """read_edg(device=...) = native parser + b2w_csr_from_edges: nodes and CSR arrays byte-identical to what the
    reference built from the same file."""
    from pecanpy_b200.graph import SparseGraph
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    path = tmp_path / (name + ".edg")
    with open(path, "w", newline="", encoding="utf-8") as f:
        f.write(str(z["text"]))
    g = SparseGraph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.read_edg(str(path), weighted=bool(z["weighted"]), directed=bool(z["directed"]), delimiter=str(z["delimiter"]),
                   device="cuda:0")
    assert g.nodes == [str(x) for x in z["nodes"]]
    for k in ("indptr", "indices", "data"):
        got = getattr(g, k)
        assert got.dtype == z[k].dtype and np.array_equal(got, z[k]), k


def _reference_csr(n, src, dst, w, directed):
    """AdjlstGraph.add_edge + to_csr restated on integer endpoints (graph.py:273-341)."""
    rows = [dict() for _ in range(n)]
    for e in range(len(src)):
        a, b = int(src[e]), int(dst[e])
        x = 1.0 if w is None else float(w[e])
        rows[a][b] = x
        if not directed:
            rows[b][a] = x
    indptr = np.zeros(n + 1, np.uint32)
    idx, dat = [], []
    for i, r in enumerate(rows):
        indptr[i + 1] = indptr[i] + len(r)
        for j in sorted(r):
            idx.append(j)
            dat.append(r[j])
    return indptr, np.array(idx, np.uint32), np.array(dat, np.float32)


@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("n,m", [(1, 0), (5, 3), (40, 400), (3000, 50000)])
def test_device_csr_equals_reference_build(n, m, weighted, directed):
    from pecanpy_b200.ingest import csr_from_edges_device
    rng = np.random.default_rng(n + m)
    src = rng.integers(0, n, m).astype(np.uint32)          # duplicates, both orientations and self loops on purpose
    dst = rng.integers(0, n, m).astype(np.uint32)
    w = (0.01 + rng.random(m)) if weighted else None
    got = csr_from_edges_device(n, src, dst, w, directed)
    want = _reference_csr(n, src, dst, w, directed)
    for g, x in zip(got, want):
        assert g.dtype == x.dtype and np.array_equal(g, x)


def test_read_edg_on_device_equals_host_path(tmp_path):
    import warnings
    from pecanpy_b200.graph import SparseGraph
    rng = np.random.default_rng(9)
    path = tmp_path / "g.edg"
    with open(path, "w") as f:
        for _ in range(3000):
            a, b = rng.integers(0, 200, size=2)
            f.write(f"n{a}\tn{b}\t{rng.choice([0.5, 1.5, 2.0, -1.0, 0.0, 0.123456789])}\n")
    host, dev = SparseGraph(), SparseGraph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        host.read_edg(str(path), weighted=True, directed=False)
        dev.read_edg(str(path), weighted=True, directed=False, device="cuda:0")
    assert host.nodes == dev.nodes
    assert np.array_equal(host.indptr, dev.indptr) and np.array_equal(host.indices, dev.indices)
    assert np.array_equal(host.data.view(np.uint32), dev.data.view(np.uint32))
    # the device-built arrays feed the engine unchanged
    from pecanpy_b200.engine import WalkEngine
    eng = WalkEngine.from_csr(dev.indptr, dev.indices, dev.data)
    assert eng.info().nnz == dev.indptr[-1]
    eng.close()
