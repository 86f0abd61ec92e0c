"""The per-edge index (b2w_edge_index.cu) against a NumPy restatement of what it must hold, and the lane-per-walker
kernel that walks it (b2w_walk_edge.cu) against the oracle and the on-the-fly kernels."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def restate_edge_index(indptr, indices):
    """For every stored edge e = (a -> b): lower_bound of a in row(b) (+ found), and the ascending positions k of
    row(b) with row(b)[k] in N(a), row(b)[k] != a   (rw/sparse_rw.py:84, :201-230; pecanpy.py:429)."""
    ip = indptr.astype(np.int64)
    n = ip.size - 1
    rows = [indices[ip[i]:ip[i + 1]] for i in range(n)]
    sets = [set(r.tolist()) for r in rows]
    out = []
    for a in range(n):
        for b in rows[a]:
            rb = rows[int(b)]
            pos = int(np.searchsorted(rb, a))
            found = pos < rb.size and rb[pos] == a
            lst = [k for k, x in enumerate(rb.tolist()) if x != a and x in sets[a]]
            out.append((int(b), pos, found, lst, rb.size))
    return out


def fetch_index(eng):
    rec = eng._keep["edge_rec"].cpu().numpy().view(np.uint32).reshape(-1, 4)
    tri = eng._keep["edge_tri"].cpu().numpy().view(np.uint32)
    return rec, tri


def graphs():
    from pecanpy_b200.synth import erdos_renyi_csr, power_law_csr
    yield "karate", load("karate_sparseotf_p1_q1")
    yield "hub400", load("uhub400_sparseotf_n2v")
    yield "directed-deadends", load("dir150_sparseotf_deadends")
    ip, ix, dt = erdos_renyi_csr(3000, 30000, seed=3)
    yield "er3000", dict(indptr=ip, indices=ix, data=dt)
    ip, ix, dt = power_law_csr(4000, 120000, seed=4)
    yield "powerlaw4000", dict(indptr=ip, indices=ix, data=dt)


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", ["karate", "hub400", "directed-deadends", "er3000", "powerlaw4000"])
def test_edge_index_content(name):
    from pecanpy_b200.engine import WalkEngine
    c = dict(graphs())[name]
    indptr, indices = c["indptr"], c["indices"]
    eng = WalkEngine.from_csr(indptr, indices, np.ones(indices.size, np.float32))
    assert eng.build_edge_index()
    rec, tri = fetch_index(eng)
    want = restate_edge_index(indptr, indices)
    nnz = indices.size
    assert rec.shape[0] == nnz + 1
    words = 0
    for e, (b, pos, found, lst, degb) in enumerate(want):
        nxt, kpf, off, deg = (int(v) for v in rec[e])
        assert nxt == b and deg == degb, e
        if degb == 0:
            continue
        assert (kpf & 0x3FFFFFFF) == pos and bool(kpf & 0x40000000) == (not found), (e, kpf, pos, found)
        assert bool(kpf & 0x80000000) == (len(lst) > 0), e
        if lst:
            assert int(tri[off]) == len(lst) and tri[off + 1: off + 1 + len(lst)].tolist() == lst, e
            words += len(lst) + 1
    assert words == eng.edge_index_words
    # the pad record serves the reference's unchecked indices[nnz] read (node 0)
    assert int(rec[nnz, 0]) == 0 and int(rec[nnz, 3]) == int(indptr[1] - indptr[0])
    eng.close()


@pytest.mark.parametrize("pq", [(4.0, 0.25), (0.5, 2.0), (1.0, 1.0), (0.25, 4.0), (2.0, 0.5)])
@pytest.mark.parametrize("flags", [0, 1, 0x80, 0x81, 0x2000000, 0x2000001],
                         ids=["filter", "forced-replay", "filter-no-ckpt", "forced-replay-no-ckpt", "lane-loops", "lane-loops-replay"])
def test_edge_kernel_equals_oracle_and_membership_kernel(pq, flags):
    """Hub rows above 1024 neighbours, many common neighbours per edge (dense core), both filter and replay."""
    import torch
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.synth import power_law_csr
    indptr, indices, data = power_law_csr(20000, 800000, seed=5)
    start = orc.shuffled_start(20000, 2, 3)
    p, q = pq
    eng = WalkEngine.from_csr(indptr, indices, data)
    got = eng.walk("SparseOTF", p, q, start, 40, seed=77, flags=flags)
    assert eng.kernel_name("SparseOTF", p, q) == "walk_uw_edge_kernel"
    assert (eng.edge_ckpt_floats > 0) == (not flags & 0x80)           # checkpoints are built on demand, per (p, q)
    st = eng.stats()
    assert int((indptr[1:] - indptr[:-1]).max()) > 256               # default here: converged warps (long rows)
    ref = eng.walk("SparseOTF", p, q, start, 40, seed=77, flags=(flags & 1) | 0x40)
    assert eng.kernel_name("SparseOTF", p, q, flags=0x40) == "walk_uw_kernel"
    assert torch.equal(got, ref)
    st2 = eng.stats()                                                 # (the replay counts differ by design)
    assert (st["steps"], st["overflow_choices"]) == (st2["steps"], st2["overflow_choices"])
    k = 3000
    want = orc.walk_csr("SparseOTF", indptr, indices, data, p, q, start[:k], 40, rng=orc.RNG_PHILOX, seed=77)
    assert np.array_equal(got[:k].cpu().numpy().view(np.uint32), want)
    eng.close()


def test_edge_kernel_directed_graph_and_dead_ends():
    """Directed graph: prev is usually NOT a neighbour of cur (no return bias, rw/sparse_rw.py:79), dead ends stop
    walkers; the edge index must encode both."""
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    c = load("dir150_sparseotf_deadends")
    ones = np.ones(c["indices"].size, np.float32)
    start = orc.shuffled_start(c["indptr"].size - 1, 20, 1)
    eng = WalkEngine.from_csr(c["indptr"], c["indices"], ones)
    for flags in (0, 1):
        got = eng.walk("SparseOTF", 0.5, 2.0, start, 30, seed=5, flags=flags).cpu().numpy().view(np.uint32)
        want = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], ones, 0.5, 2.0, start, 30, rng=orc.RNG_PHILOX, seed=5)
        assert np.array_equal(got, want)
    assert eng.kernel_name("SparseOTF", 0.5, 2.0) == "walk_uw_edge_kernel"
    eng.close()


def test_edge_kernel_after_overflow_read():
    """u so close to 1 that cdf[-1] < u: the reference reads indices[indptr[cur] + deg] unchecked (pecanpy.py:559);
    the next step did not arrive over a stored edge and must still equal the oracle (fed uniforms force it)."""
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import _capi as capi
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.synth import erdos_renyi_csr
    indptr, indices, data = erdos_renyi_csr(500, 6000, seed=9)
    rng = np.random.default_rng(0)
    rows, L = 4000, 12
    start = rng.integers(0, 500, rows).astype(np.uint32)
    feed = rng.random((rows, L))
    feed[rng.random((rows, L)) < 0.08] = 1.0 - 2.0 ** -53             # the largest double below 1
    eng = WalkEngine.from_csr(indptr, indices, data)
    for p, q in ((4.0, 0.25), (0.5, 2.0)):
        want = orc.walk_csr("SparseOTF", indptr, indices, data, p, q, start, L, rng=orc.RNG_FEED, feed=feed)
        # the step without an edge: by the walker's lane (plain loops) and by its whole warp (converged loops)
        for flags in (0, capi.FLAG_OFFEDGE_LANE, capi.FLAG_OFFEDGE_WARP):
            got = eng.walk("SparseOTF", p, q, start, L, rng=capi.RNG_FEED, feed=feed.ravel(), flags=flags).cpu().numpy().view(np.uint32)
            assert eng.kernel_name("SparseOTF", p, q) == "walk_uw_edge_kernel"
            assert eng.stats()["overflow_choices"] > 0
            assert np.array_equal(got, want), flags
    eng.close()


def test_edge_checkpoints_equal_the_reference_cdf():
    """The checkpoints are the reference's own float32 cdf (sequential cumsum of w / w.sum(), rw/sparse_rw.py:89,
    pecanpy.py:556) after elements 127, 255, ... of row(cur), for every stored edge into a row of >= 128 slots."""
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.synth import power_law_csr
    indptr, indices, data = power_law_csr(3000, 60000, seed=11)
    eng = WalkEngine.from_csr(indptr, indices, data)
    p, q = 4.0, 0.25
    assert eng.build_edge_index() and eng.build_edge_ckpt(p, q)
    ckb = eng._keep["edge_ckb"].cpu().numpy().view(np.uint32)
    ck = eng._keep["edge_ckpt"].cpu().numpy()
    ip = indptr.astype(np.int64)
    deg = ip[1:] - ip[:-1]
    checked = 0
    for cur in np.argsort(-deg)[:12]:
        d = int(deg[cur])
        if d < 128:
            continue
        nck = d // 128
        row = indices[ip[cur]:ip[cur + 1]]
        for kp in (0, d // 3, d - 1):
            prev = int(row[kp])
            probs = orc.sparse_probs(indptr, indices, data, p, q, int(cur), prev)
            cdf = np.float32(0)
            want = []
            for k in range(nck * 128):
                cdf = np.float32(cdf + probs[k])
                if k % 128 == 127:
                    want.append(cdf)
            got = ck[int(ckb[cur]) + kp * nck: int(ckb[cur]) + (kp + 1) * nck]
            assert np.array_equal(got.view(np.uint32), np.array(want, np.float32).view(np.uint32)), (cur, kp)
            checked += 1
    assert checked >= 6
    eng.close()
