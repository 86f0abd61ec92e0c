"""The weighted per-edge index (b2w_wedge.cu) against a NumPy restatement of what it must hold, and the lane-per-walker
kernel that walks it against the oracle and the weight-streaming kernel: weighted graphs, node2vec+, p and q off the
power-of-two grid, hub rows (checkpointed replays), dead ends, the step after the reference's unchecked read."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def restate(indptr, indices, data, p, q, extend, thr):
    """Per row: base weights and their prefix; per stored edge (a -> b): return-edge position / weight, exceptions
    (position, weight) and the reference's sequential float32 sum of the biased weights (rw/sparse_rw.py:51-130)."""
    ip = indptr.astype(np.int64)
    n = ip.size - 1
    invq = 1.0 / q
    supp = min(1.0, invq)
    rows = [indices[ip[i]:ip[i + 1]] for i in range(n)]
    wts = [data[ip[i]:ip[i + 1]] for i in range(n)]

    def base(w, thr_cur):
        if not extend:
            return f32(np.float64(w) / q)
        alpha = invq + (1.0 - invq) * 0.0
        if w < thr_cur:
            alpha = supp
        return f32(np.float64(w) * alpha)

    bw = np.concatenate([[base(w, thr[i] if extend else 0) for w in wts[i]] for i in range(n)] + [[]]).astype(f32)
    out = []
    for a in range(n):
        pa = {int(x): float(w) for x, w in zip(rows[a], wts[a])}
        for b in rows[a]:
            b = int(b)
            rb, wb = rows[b], wts[b]
            pos = int(np.searchsorted(rb, a))
            found = pos < rb.size and rb[pos] == a
            exc = []
            biased = bw[ip[b]:ip[b + 1]].copy()
            for k, (x, w) in enumerate(zip(rb.tolist(), wb.tolist())):
                w = f32(w)
                if x == a:
                    biased[k] = f32(np.float64(w) / p)
                    continue
                if x in pa:
                    if not extend:
                        v = w
                    else:
                        wp, th = f32(pa[x]), f32(thr[x])
                        if wp >= th:
                            v = w
                        else:
                            t = f32(wp / th)
                            alpha = invq + (1.0 - invq) * np.float64(t)
                            if w < thr[b]:
                                alpha = supp
                            v = f32(np.float64(w) * alpha)
                    if v.view(np.uint32) != biased[k].view(np.uint32):
                        exc.append((k, v))
                    biased[k] = v
            S = f32(0)
            for v in biased:
                S = f32(S + v)
            out.append(dict(nxt=b, pos=pos, found=bool(found), exc=exc, S=S, deg=rb.size, cs=int(ip[b]),
                            vkp=f32(np.float64(wb[pos]) / p) if found else f32(0)))
    return bw, out


@pytest.mark.parametrize("case", ["w200-n2v", "w200-ext", "hub400-ext", "dir150"])
def test_windex_content(case):
    from pecanpy_b200.engine import WalkEngine
    name, p, q, extend, gamma = {"w200-n2v": ("w200_sparseotf_n2v", 0.3, 0.7, False, 0.0),
                                 "w200-ext": ("w200_sparseotf_ext_g05", 0.3, 3.0, True, 0.5),
                                 "hub400-ext": ("hub400_sparseotf_ext", 0.5, 2.0, True, 0.0),
                                 "dir150": ("dir150_sparseotf_deadends", 2.0, 0.5, False, 0.0)}[case]
    c = load(name)
    indptr, indices, data = c["indptr"], c["indices"], c["data"]
    eng = WalkEngine.from_csr(indptr, indices, data)
    thr = None
    if extend:
        thr = eng.compute_thresholds(gamma).cpu().numpy()
    assert eng.build_windex(p, q, extend)
    nnz = indices.size
    rec = eng._keep["w_rec"].cpu().numpy().view(np.uint32).reshape(-1, 8)
    bw = eng._keep["w_bw"].cpu().numpy()[:nnz]
    bq = eng._keep["w_bq"].cpu().numpy()[:nnz]
    exc = eng._keep["w_exc"].cpu().numpy().view(np.uint8).reshape(-1, 24)
    want_bw, want = restate(indptr, indices, data, p, q, extend, thr)
    assert np.array_equal(bw.view(np.uint32), want_bw.view(np.uint32))
    for i in range(indptr.size - 1):
        s, e = int(indptr[i]), int(indptr[i + 1])
        assert np.allclose(bq[s:e], np.cumsum(bw[s:e].astype(np.float64)), rtol=1e-13, atol=0)
    assert rec.shape[0] == nnz + 1
    for e, w in enumerate(want):
        nxt, kpf, off, deg, cs, S, vkp, bkp = (int(v) for v in rec[e])
        assert (nxt, deg, cs) == (w["nxt"], w["deg"], w["cs"]), e
        if deg == 0:
            continue
        assert (kpf & 0x3FFFFFFF) == w["pos"] and bool(kpf & 0x40000000) == (not w["found"]), e
        assert bool(kpf & 0x80000000) == (len(w["exc"]) > 0), e
        assert S == int(w["S"].view(np.uint32)), (e, np.uint32(S).view(f32), w["S"])
        if w["found"]:
            assert vkp == int(w["vkp"].view(np.uint32)), e
            assert bkp == int(want_bw[cs + w["pos"]].view(np.uint32)), e
        if w["exc"]:
            hdr = exc[off]
            assert int(hdr[:4].view(np.uint32)[0]) == len(w["exc"]), e
            dev = 0.0
            for t, (k, v) in enumerate(w["exc"]):
                ent = exc[off + 1 + t]
                assert int(ent[:4].view(np.uint32)[0]) == k and ent[4:8].view(np.uint32)[0] == v.view(np.uint32), (e, t)
                dev += float(v) - float(bw[cs + k])
                assert np.isclose(ent[16:24].view(np.float64)[0], dev, rtol=1e-12, atol=1e-300)
                assert np.isclose(ent[8:16].view(np.float64)[0], bq[cs + k] + dev, rtol=1e-12, atol=1e-300)
    eng.close()


CASES = [(0.5, 2.0, False), (4.0, 0.25, False), (0.3, 0.7, False), (1.0, 1.0, False), (0.5, 2.0, True), (0.3, 3.1, True),
         (4.0, 0.25, True)]


@pytest.mark.parametrize("p,q,extend", CASES)
@pytest.mark.parametrize("flags", [0, 1, 0x1000000], ids=["filter", "forced-replay", "warp-loops"])
def test_wedge_kernel_equals_oracle_and_streaming_kernel(p, q, extend, flags):
    """Weighted power-law graph with hub rows above 1024 slots (checkpointed replays, multi-exception edges)."""
    import torch
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.synth import power_law_csr
    indptr, indices, data = power_law_csr(20000, 800000, seed=6, weighted=True)
    start = orc.shuffled_start(20000, 2, 4)
    eng = WalkEngine.from_csr(indptr, indices, data)
    thr = eng.compute_thresholds(0.25).cpu().numpy() if extend else None
    got = eng.walk("SparseOTF", p, q, start, 40, seed=31, extend=extend, flags=flags)
    assert eng.kernel_name("SparseOTF", p, q, extend) == "walk_wedge_kernel"
    st = eng.stats()
    ref = eng.walk("SparseOTF", p, q, start, 40, seed=31, extend=extend, flags=(flags & 1) | 0x40)
    assert eng.kernel_name("SparseOTF", p, q, extend, flags=0x40) == "walk_sparse_warp_kernel"
    assert torch.equal(got, ref)
    st2 = eng.stats()
    assert (st["steps"], st["overflow_choices"]) == (st2["steps"], st2["overflow_choices"])
    k = 2500
    want = orc.walk_csr("SparseOTF", indptr, indices, data, p, q, start[:k], 40, extend=extend, thr=thr,
                        rng=orc.RNG_PHILOX, seed=31)
    assert np.array_equal(got[:k].cpu().numpy().view(np.uint32), want)
    eng.close()


def test_wedge_kernel_unweighted_graph_with_biases_off_the_grid():
    """p = 0.3, q = 0.7 on an unweighted graph: not eligible for the integer kernels, served by the weighted index."""
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.synth import power_law_csr
    indptr, indices, data = power_law_csr(20000, 400000, seed=5)
    start = orc.shuffled_start(20000, 1, 2)[:8000]
    eng = WalkEngine.from_csr(indptr, indices, data)
    got = eng.walk("SparseOTF", 0.3, 0.7, start, 30, seed=3).cpu().numpy().view(np.uint32)
    assert eng.kernel_name("SparseOTF", 0.3, 0.7) == "walk_wedge_kernel"
    want = orc.walk_csr("SparseOTF", indptr, indices, data, 0.3, 0.7, start, 30, rng=orc.RNG_PHILOX, seed=3)
    assert np.array_equal(got, want)
    eng.close()


def test_wedge_kernel_after_overflow_read_and_dead_ends():
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import _capi as capi
    from pecanpy_b200.engine import WalkEngine
    from pecanpy_b200.synth import erdos_renyi_csr
    indptr, indices, data = erdos_renyi_csr(500, 6000, seed=9, weighted=True)
    rng = np.random.default_rng(0)
    rows, L = 4000, 12
    start = rng.integers(0, 500, rows).astype(np.uint32)
    feed = rng.random((rows, L))
    feed[rng.random((rows, L)) < 0.08] = 1.0 - 2.0 ** -53             # the largest double below 1
    eng = WalkEngine.from_csr(indptr, indices, data)
    want = orc.walk_csr("SparseOTF", indptr, indices, data, 0.5, 2.0, start, L, rng=orc.RNG_FEED, feed=feed)
    for flags in (0, capi.FLAG_OFFEDGE_WARP):                        # steps without an edge: by the lane / by the warp
        got = eng.walk("SparseOTF", 0.5, 2.0, start, L, rng=capi.RNG_FEED, feed=feed.ravel(), flags=flags).cpu().numpy().view(np.uint32)
        assert eng.kernel_name("SparseOTF", 0.5, 2.0) == "walk_wedge_kernel" and eng.stats()["overflow_choices"] > 0
        assert np.array_equal(got, want), flags
    eng.close()
    c = load("dir150_sparseotf_deadends")
    start = orc.shuffled_start(c["indptr"].size - 1, 20, 1)
    eng = WalkEngine.from_csr(c["indptr"], c["indices"], c["data"])
    got = eng.walk("SparseOTF", 2.0, 0.5, start, 30, seed=5).cpu().numpy().view(np.uint32)
    want = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], 2.0, 0.5, start, 30, rng=orc.RNG_PHILOX, seed=5)
    assert np.array_equal(got, want)
    eng.close()


def test_windex_follows_the_parameters():
    """The index is valid for one (p, q, extend, thresholds): other parameters rebuild it (or fall back)."""
    from oracle import oracle as orc
    from pecanpy_b200.engine import WalkEngine
    c = load("hub400_sparseotf_n2v")
    eng = WalkEngine.from_csr(c["indptr"], c["indices"], c["data"])
    start = c["start"][:300]
    for p, q in ((4.0, 0.25), (0.5, 2.0), (4.0, 0.25)):
        got = eng.walk("SparseOTF", p, q, start, 20, seed=8).cpu().numpy().view(np.uint32)
        want = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], p, q, start, 20, rng=orc.RNG_PHILOX, seed=8)
        assert np.array_equal(got, want)
        assert eng.kernel_name("SparseOTF", p, q) == "walk_wedge_kernel"
    # an index built for other parameters is never used
    assert eng.kernel_name("SparseOTF", 0.7, 0.7) == "walk_sparse_warp_kernel"
    eng.close()
