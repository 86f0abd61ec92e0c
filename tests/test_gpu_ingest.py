"""b2w_csr_from_edges (device CSR build) against the reference's ingest conventions (graph.py:160-341): dict of
dicts filled line by line -- later lines overwrite -- rows emitted with sorted columns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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
