"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot walk
10^7 rows in test time): every step is an edge of the graph, rows are well formed, independent kernel
variants (edge-index kernel, membership-bitmap kernel, generic weight-streaming kernel) produce IDENTICAL matrices, a
prefix of the rows equals the oracle, and the result does not depend on how rows are sharded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_rows_are_walks(torch, dev, indptr, indices, walks_t, L):
    """walks_t: int32 device tensor [rows, L+2].  Checks edge validity of every step and the row layout."""
    ip = torch.from_numpy(indptr.astype(np.int64)).to(dev)
    ix = torch.from_numpy(indices.astype(np.int64)).to(dev)
    rows = walks_t.shape[0]
    non_edges = 0
    chunk = 1 << 18
    for r0 in range(0, rows, chunk):
        w = walks_t[r0:r0 + chunk].to(torch.int64)
        eff = w[:, L + 1]
        assert int(eff.min()) >= 1 and int(eff.max()) <= L + 1
        cols = torch.arange(L + 1, device=dev)[None, :]
        body = w[:, :L + 1]
        assert bool(((cols >= eff[:, None]) <= (body == 0)).all()), "non-zero entries after the effective length"
        src, dst = body[:, :-1], body[:, 1:]
        valid = cols[:, 1:] < eff[:, None]                       # step j exists iff j < eff
        s, d = src[valid], dst[valid]
        lo, hi = ip[s], ip[s + 1]
        # binary search of d in the sorted row [lo, hi)
        pos = lo.clone()
        span = int((hi - lo).max())
        step = 1 << max(span.bit_length() - 1, 0)
        while step:
            cand = pos + step
            ok = (cand <= hi) & (ix[torch.clamp(cand - 1, max=ix.numel() - 1)] < d)
            pos = torch.where(ok, cand, pos)
            step >>= 1
        found = (pos < hi) & (ix[torch.clamp(pos, max=ix.numel() - 1)] == d)
        non_edges += int((~found).sum())
        # a walker stops early only at a node without neighbours
        stopped = eff < L + 1
        if bool(stopped.any()):
            lastn = body[stopped, :].gather(1, (eff[stopped] - 1)[:, None]).squeeze(1)
            assert bool((ip[lastn + 1] == ip[lastn]).all())
    return non_edges


@pytest.mark.parametrize("workload", ["er", "powerlaw"])
def test_full_size_sparse_otf(workload):
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import synth
    from pecanpy_b200.engine import WalkEngine
    dev = torch.device("cuda", 0)
    if workload == "er":       # BASELINE config #2, full size
        indptr, indices, data = synth.erdos_renyi_csr(100_000, 1_000_000, seed=0)
        p, q, num_walks = 0.5, 2.0, 10
    else:                      # BASELINE config #3 graph, one walk per node (10^6 walkers)
        indptr, indices, data = synth.power_law_csr(1_000_000, 10_000_000, seed=1)
        p, q, num_walks = 4.0, 0.25, 1
    n, L = indptr.size - 1, 80
    start = synth.shuffled_start(n, num_walks, 0)
    eng = WalkEngine.from_csr(indptr, indices, data, device=dev)
    a = eng.walk("SparseOTF", p, q, start, L, seed=5)                      # edge-index kernel (index built on demand)
    assert eng.kernel_name("SparseOTF", p, q) == "walk_uw_edge_kernel"
    st = eng.stats()
    m = eng.walk("SparseOTF", p, q, start, L, seed=5, flags=0x40)          # membership-bitmap kernel
    assert eng.kernel_name("SparseOTF", p, q, flags=0x40) == "walk_uw_kernel"
    assert torch.equal(a, m), "edge-index and membership kernels disagree at full size"
    del m
    b = eng.walk("SparseOTF", p, q, start, L, seed=5, flags=8)             # generic weight-streaming kernel
    assert torch.equal(a, b), "kernel variants disagree at full size"
    del b
    assert st["steps"] == eng.count_steps(a, L)
    # the only steps that may leave the edge set are the reference's own unchecked `choice == deg` reads
    # (cdf[-1] < u, ~1e-7 per step), which the engine reproduces and counts
    assert _check_rows_are_walks(torch, dev, indptr, indices, a, L) <= st["overflow_choices"]
    # determinism and shard invariance (rows keyed by the global row index)
    h = start.size // 3
    lo = eng.walk("SparseOTF", p, q, start[:h], L, seed=5, row0=0)
    hi = eng.walk("SparseOTF", p, q, start[h:], L, seed=5, row0=h)
    assert torch.equal(torch.cat([lo, hi]), a)
    # a prefix of the rows against the oracle
    k = 20000
    want = orc.walk_csr("SparseOTF", indptr, indices, data, p, q, start[:k], L, rng=orc.RNG_PHILOX, seed=5)
    assert np.array_equal(a[:k].cpu().numpy().view(np.uint32), want)
    eng.close()


def test_full_size_precomp_config4():
    """BASELINE config #4 at full size: ER 50k nodes / 1M edges, weighted, PreComp p=0.25 q=4 -- 8.2e7 table entries
    (656 MB).  Tables byte-identical to the oracle's (packed layout, de-interleaved view), a 20k-row prefix of the
    walks equals the oracle, and both walk kernels / both table layouts give the same matrix for all 5e5 rows."""
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import synth
    from pecanpy_b200.engine import WalkEngine
    indptr, indices, data = synth.erdos_renyi_csr(50_000, 1_000_000, seed=2, weighted=True)
    p, q, L = 0.25, 4.0, 80
    start = synth.shuffled_start(50_000, 10, 0)
    eng = WalkEngine.from_csr(indptr, indices, data, device="cuda:0")
    aip, aj, aq = eng.build_alias(indptr, p, q)
    o_aip, o_j, o_q = orc.alias_build(indptr, indices, data, p, q)
    assert int(aip[-1]) > 80_000_000 and np.array_equal(aip, o_aip)
    assert np.array_equal(aj.cpu().numpy().view(np.uint32), o_j)
    assert np.array_equal(aq.cpu().numpy().view(np.uint32), o_q.view(np.uint32))
    a = eng.walk("PreComp", p, q, start, L, seed=11)
    assert eng.kernel_name("PreComp", p, q) == "walk_precomp_edge_kernel"
    st = eng.stats()
    b = eng.walk("PreComp", p, q, start, L, seed=11, flags=0x40)      # bisecting kernel, packed tables
    assert torch.equal(a, b)
    assert st["steps"] == eng.count_steps(a, L) == 50_000 * 10 * L
    k = 20000
    want = orc.walk_csr("PreComp", indptr, indices, data, p, q, start[:k], L, alias=(o_aip, o_j, o_q),
                        rng=orc.RNG_PHILOX, seed=11)
    assert np.array_equal(a[:k].cpu().numpy().view(np.uint32), want)
    host = eng.walk_host("PreComp", p, q, start, L, seed=11)
    assert np.array_equal(host, a.cpu().numpy().view(np.uint32))
    eng2 = WalkEngine.from_csr(indptr, indices, data, device="cuda:0")
    eng2.build_alias(indptr, p, q, packed=False)                      # the reference's two-array layout
    c = eng2.walk("PreComp", p, q, start, L, seed=11, flags=0x40)
    assert torch.equal(a, c)
    eng.close(); eng2.close()


def test_full_size_dense_config5():
    """BASELINE config #5 at full size: dense 20 000 nodes, 30 % density, weighted, DenseOTF node2vec+ (gamma 0,
    p=0.5 q=2): 20 TMA tiles per row with the 3-stage ring wrapping.  Thresholds equal the oracle's, a 2k-row prefix
    of the walks equals the oracle, TMA and LDG load paths agree on every row."""
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import synth
    from pecanpy_b200.engine import WalkEngine
    n, L = 20_000, 80
    data, nz = synth.dense_weighted(n, 0.3, 3)
    eng = WalkEngine.from_dense(data, nz, device="cuda:0")
    thr = eng.compute_thresholds(0.0).cpu().numpy()
    want_thr = orc.noise_thresholds_dense(data, nz, 0.0)
    assert np.array_equal(thr.view(np.uint32), want_thr.view(np.uint32))
    start = synth.shuffled_start(n, 2, 0)                            # 40 000 walkers
    a = eng.walk("DenseOTF", 0.5, 2.0, start, L, seed=13, extend=True)             # TMA staged
    st = eng.stats()
    b = eng.walk("DenseOTF", 0.5, 2.0, start, L, seed=13, extend=True, flags=0x10)  # per-lane vector loads
    assert torch.equal(a, b)
    assert st["steps"] == eng.count_steps(a, L) == start.size * L
    k = 1500
    want = orc.walk_dense(data, nz, 0.5, 2.0, start[:k], L, extend=True, thr=want_thr, rng=orc.RNG_PHILOX, seed=13)
    assert np.array_equal(a[:k].cpu().numpy().view(np.uint32), want)
    eng.close()


@pytest.mark.parametrize("extend", [False, True], ids=["n2v", "n2v+"])
def test_full_size_weighted_sparse_otf(extend):
    """BASELINE config #3's topology with random weights, one walk per node (10^6 walkers): the weighted edge-index
    kernel and the weight-streaming kernel give IDENTICAL matrices; a prefix equals the oracle."""
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import synth
    from pecanpy_b200.engine import WalkEngine
    indptr, indices, data = synth.power_law_csr(1_000_000, 10_000_000, seed=1, weighted=True)
    p, q, L = 4.0, 0.25, 80
    start = synth.shuffled_start(1_000_000, 1, 0)
    eng = WalkEngine.from_csr(indptr, indices, data, device="cuda:0")
    thr = eng.compute_thresholds(0.0).cpu().numpy() if extend else None
    a = eng.walk("SparseOTF", p, q, start, L, seed=5, extend=extend)
    assert eng.kernel_name("SparseOTF", p, q, extend) == "walk_wedge_kernel"
    st = eng.stats()
    b = eng.walk("SparseOTF", p, q, start, L, seed=5, extend=extend, flags=0x40)
    assert eng.kernel_name("SparseOTF", p, q, extend, flags=0x40) == "walk_sparse_warp_kernel"
    assert torch.equal(a, b), "weighted edge-index kernel and weight-streaming kernel disagree at full size"
    assert st["steps"] == eng.count_steps(a, L)
    assert _check_rows_are_walks(torch, torch.device("cuda", 0), indptr, indices, a, L) <= st["overflow_choices"]
    k = 10000
    want = orc.walk_csr("SparseOTF", indptr, indices, data, p, q, start[:k], L, extend=extend, thr=thr,
                        rng=orc.RNG_PHILOX, seed=5)
    assert np.array_equal(a[:k].cpu().numpy().view(np.uint32), want)
    eng.close()
