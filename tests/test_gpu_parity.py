"""GPU parity: the CUDA engine (through the C ABI) vs the pinned CPU oracle and the committed
golden fixtures produced by the reference.  Bit-exact: walks are integer node indices.

Regimes:
  R1  B2W_RNG_FEED with the reference's own MT19937 uniforms  -> equals the UNMODIFIED reference
      (tests/golden/*.npz, written by oracle/gen_golden.py) for the OTF modes without dead ends;
  R2  B2W_RNG_PHILOX -> equals the oracle driven by the same Philox stream, every mode, dead ends,
      isolated nodes, node2vec+, weighted, hubs.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def mods():
    import torch
    from oracle import oracle as orc
    from pecanpy_b200 import _capi as capi
    from pecanpy_b200.engine import WalkEngine
    assert torch.cuda.is_available()
    return dict(torch=torch, orc=orc, capi=capi, WalkEngine=WalkEngine)


def to_np(t):
    return t.cpu().numpy().view(np.uint32)


def first_diff(a, b):
    bad = np.argwhere((a != b).any(axis=1)).ravel()
    if bad.size == 0:
        return "equal"
    r = bad[0]
    c = np.argwhere(a[r] != b[r]).ravel()[0]
    return f"{bad.size} bad rows; first row {r} col {c}: got {a[r, max(0, c - 2):c + 3]} want {b[r, max(0, c - 2):c + 3]}"


def test_philox_device_known_answers(mods):
    import ctypes as C
    lib = mods["capi"].lib()
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        mods["capi"].check(lib.b2w_philox_selftest(c, k, o))
        assert list(o) == want


# b2w_walk flags: 1 forced replay, 4 lane-per-walker kernel, 8 generic (weight-streaming) kernel even on
# unweighted graphs, 0x20 warp-cooperative state-machine kernel (sub-warp groups), 0x40 ignore the per-edge index
# (on-the-fly membership kernels), bits 8..15 lanes per walker of the membership-bitmap kernel (which also routes
# around the edge-index kernel).  "auto" on an unweighted graph with representable biases = the edge-index kernel.
FLAG_SETS = {"auto": 0, "auto-replay": 1, "lane-per-walker": 4, "generic": 8, "generic-replay": 9,
             "membership": 0x40, "membership-replay": 0x41,
             "uw-g8": 8 << 8, "uw-g16": 16 << 8, "uw-g32": 32 << 8, "uw-g8-replay": (8 << 8) | 1,
             "uw-g32-replay": (32 << 8) | 1, "uw-coop": 0x20, "generic-coop": 0x28, "uw-g16-coop": (16 << 8) | 0x20}

SPARSE_R1 = ["testwalk_SparseOTF", "karate_sparseotf_p1_q1", "karate_sparseotf_p05_q2", "karate_sparseotf_p03_q07",
             "w200_sparseotf_n2v", "w200_sparseotf_ext_g0", "w200_sparseotf_ext_g05", "hub400_sparseotf_n2v",
             "hub400_sparseotf_ext", "uhub400_sparseotf_n2v"]


@pytest.mark.parametrize("flags", list(FLAG_SETS.values()), ids=list(FLAG_SETS))
@pytest.mark.parametrize("name", SPARSE_R1)
def test_sparse_otf_replays_reference_R1(mods, name, flags):
    """MT-replay: feed the reference's uniforms; rows that never dead-end consume exactly L doubles in
    row order, so the whole matrix must equal the unmodified reference's."""
    c = load(name)
    L = int(c["walk_length"])
    if (c["walks"][:, -1] != L + 1).any():
        pytest.skip("fixture has dead ends: per-row feed not defined (covered by R2)")
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    if bool(c["extend"]):
        eng.set_thresholds(c["thr"])
    feed = mods["orc"].mt_uniform_feed(int(c["seed"]), c["start"].size, L)
    got = to_np(eng.walk("SparseOTF", float(c["p"]), float(c["q"]), c["start"], L, extend=bool(c["extend"]),
                         rng=mods["capi"].RNG_FEED, feed=feed, flags=flags))
    assert np.array_equal(got, c["walks"]), first_diff(got, c["walks"])
    eng.close()


@pytest.mark.parametrize("name", ["testwalk_DenseOTF", "karate_denseotf_p05_q2", "w200_denseotf_n2v", "w200_denseotf_ext"])
@pytest.mark.parametrize("flags", [0, 1], ids=["filter", "forced-replay"])
def test_dense_otf_replays_reference_R1(mods, name, flags):
    c = load(name)
    L = int(c["walk_length"])
    if (c["walks"][:, -1] != L + 1).any():
        pytest.skip("fixture has dead ends")
    eng = mods["WalkEngine"].from_dense(c["dense"], c["nonzero"])
    if bool(c["extend"]):
        eng.set_thresholds(c["thr"])
    feed = mods["orc"].mt_uniform_feed(int(c["seed"]), c["start"].size, L)
    got = to_np(eng.walk("DenseOTF", float(c["p"]), float(c["q"]), c["start"], L, extend=bool(c["extend"]),
                         rng=mods["capi"].RNG_FEED, feed=feed, flags=flags))
    assert np.array_equal(got, c["walks"]), first_diff(got, c["walks"])
    eng.close()


ALL_SPARSE = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                    if "sparseotf" in f or "SparseOTF" in f)


@pytest.mark.parametrize("flags", list(FLAG_SETS.values()), ids=list(FLAG_SETS))
@pytest.mark.parametrize("name", ALL_SPARSE)
def test_sparse_otf_philox_R2(mods, name, flags):
    c = load(name)
    orc = mods["orc"]
    L = int(c["walk_length"])
    ext = bool(c["extend"])
    want = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], float(c["p"]), float(c["q"]), c["start"], L,
                        extend=ext, thr=c.get("thr"), rng=orc.RNG_PHILOX, seed=1234 + int(c["seed"]))
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    if ext:
        eng.set_thresholds(c["thr"])
    got = to_np(eng.walk("SparseOTF", float(c["p"]), float(c["q"]), c["start"], L, seed=1234 + int(c["seed"]),
                         extend=ext, flags=flags))
    assert np.array_equal(got, want), first_diff(got, want)
    st = eng.stats()
    assert st["steps"] == int((want[:, -1].astype(np.int64) - 1).sum())
    eng.close()


PRECOMP_VARIANTS = {"packed+edge-index": (True, 0), "packed": (True, 0x40), "two-arrays+edge-index": (False, 0),
                    "two-arrays": (False, 0x40)}


@pytest.mark.parametrize("variant", list(PRECOMP_VARIANTS), ids=list(PRECOMP_VARIANTS))
@pytest.mark.parametrize("name", ["karate_precomp_p025_q4", "w200_precomp_n2v", "w200_precomp_ext",
                                  "dir150_precomp_deadends", "testwalk_PreComp"])
def test_precomp_tables_and_walks(mods, name, variant):
    """Alias tables byte-identical to the reference's arrays (RNG-free parity) in both table layouts (the
    reference's two arrays / one packed {q, j} table), then Philox walks vs the oracle through both walk kernels
    (per-edge records / bisection of prev in row(cur))."""
    packed, flags = PRECOMP_VARIANTS[variant]
    c = load(name)
    orc = mods["orc"]
    L = int(c["walk_length"])
    ext = bool(c["extend"])
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    if ext:
        eng.set_thresholds(c["thr"])
    aip, aj, aq = eng.build_alias(c["indptr"], float(c["p"]), float(c["q"]), extend=ext, packed=packed)
    assert np.array_equal(aip, c["alias_indptr"])
    assert np.array_equal(to_np(aj), c["alias_j"])
    assert np.array_equal(to_np(aq), c["alias_q"].view(np.uint32))
    want = orc.walk_csr("PreComp", c["indptr"], c["indices"], c["data"], float(c["p"]), float(c["q"]), c["start"], L,
                        alias=(c["alias_indptr"], c["alias_j"], c["alias_q"]), rng=orc.RNG_PHILOX, seed=99)
    got = to_np(eng.walk("PreComp", float(c["p"]), float(c["q"]), c["start"], L, seed=99, flags=flags))
    assert eng.kernel_name("PreComp", float(c["p"]), float(c["q"]), flags=flags) == \
        ("walk_thread_kernel<PRECOMP>" if flags else "walk_precomp_edge_kernel")
    assert np.array_equal(got, want), first_diff(got, want)
    assert eng.stats()["steps"] == int((want[:, -1].astype(np.int64) - 1).sum())
    eng.close()


@pytest.mark.parametrize("extend", [False, True], ids=["n2v", "n2v+"])
def test_precomp_tables_hub_graph(mods, extend):
    """max degree > 64: the global-scratch table builder (the fixtures above use the shared-memory one)."""
    c = load("hub400_sparseotf_ext")
    orc = mods["orc"]
    assert int((c["indptr"][1:] - c["indptr"][:-1]).max()) > 64
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    if extend:
        eng.set_thresholds(c["thr"])
    aip, aj, aq = eng.build_alias(c["indptr"], 0.7, 0.3, extend=extend)
    o_aip, o_j, o_q = orc.alias_build(c["indptr"], c["indices"], c["data"], 0.7, 0.3, extend, c["thr"] if extend else None)
    assert np.array_equal(aip, o_aip)
    assert np.array_equal(to_np(aj), o_j)
    assert np.array_equal(to_np(aq), o_q.view(np.uint32))
    want = orc.walk_csr("PreComp", c["indptr"], c["indices"], c["data"], 0.7, 0.3, c["start"], 20,
                        alias=(o_aip, o_j, o_q), rng=orc.RNG_PHILOX, seed=123)
    got = to_np(eng.walk("PreComp", 0.7, 0.3, c["start"], 20, seed=123))
    assert np.array_equal(got, want), first_diff(got, want)
    got = to_np(eng.walk("PreComp", 0.7, 0.3, c["start"], 20, seed=123, flags=0x40))   # without the edge records
    assert np.array_equal(got, want), first_diff(got, want)
    eng.close()


@pytest.mark.parametrize("name", ["testwalk_FirstOrderUnweighted", "karate_firstorder"])
def test_first_order_unweighted(mods, name):
    c = load(name)
    orc = mods["orc"]
    L = int(c["walk_length"])
    want = orc.walk_csr("FirstOrderUnweighted", c["indptr"], c["indices"], c["data"], 1, 1, c["start"], L,
                        rng=orc.RNG_PHILOX, seed=5)
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    got = to_np(eng.walk("FirstOrderUnweighted", 1, 1, c["start"], L, seed=5))
    assert np.array_equal(got, want), first_diff(got, want)
    eng.close()


@pytest.mark.parametrize("name", ["testwalk_PreCompFirstOrder", "w200_precompfirstorder"])
def test_precomp_first_order(mods, name):
    c = load(name)
    orc = mods["orc"]
    L = int(c["walk_length"])
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    _, aj, aq = eng.build_alias(c["indptr"], 1, 1, first_order=True)
    assert np.array_equal(to_np(aj), c["alias_j"])
    assert np.array_equal(to_np(aq), c["alias_q"].view(np.uint32))
    want = orc.walk_csr("PreCompFirstOrder", c["indptr"], c["indices"], c["data"], 1, 1, c["start"], L,
                        alias=(None, c["alias_j"], c["alias_q"]), rng=orc.RNG_PHILOX, seed=6)
    got = to_np(eng.walk("PreCompFirstOrder", 1, 1, c["start"], L, seed=6))
    assert np.array_equal(got, want), first_diff(got, want)
    eng.close()


@pytest.mark.parametrize("flags", [0, 1], ids=["filter", "forced-replay"])
@pytest.mark.parametrize("name", ["testwalk_DenseOTF", "karate_denseotf_p05_q2", "w200_denseotf_n2v",
                                  "w200_denseotf_ext", "dir150_denseotf_deadends"])
def test_dense_otf_philox_R2(mods, name, flags):
    c = load(name)
    orc = mods["orc"]
    L = int(c["walk_length"])
    ext = bool(c["extend"])
    want = orc.walk_dense(c["dense"], c["nonzero"], float(c["p"]), float(c["q"]), c["start"], L, extend=ext,
                          thr=c.get("thr"), rng=orc.RNG_PHILOX, seed=77)
    eng = mods["WalkEngine"].from_dense(c["dense"], c["nonzero"])
    if ext:
        eng.set_thresholds(c["thr"])
    got = to_np(eng.walk("DenseOTF", float(c["p"]), float(c["q"]), c["start"], L, seed=77, extend=ext, flags=flags))
    assert np.array_equal(got, want), first_diff(got, want)
    eng.close()


@pytest.mark.parametrize("flags", list(FLAG_SETS.values()), ids=list(FLAG_SETS))
@pytest.mark.parametrize("pq", [(4.0, 0.25), (0.5, 2.0), (1.0, 1.0), (0.3, 0.7)])
def test_power_law_hubs_unweighted(mods, pq, flags):
    """Unweighted power-law graph with hub rows above 1024 neighbours (global bitmap / scratch rows, both
    search directions, multi-word prefix); every kernel variant must equal the oracle."""
    from pecanpy_b200.synth import power_law_csr
    orc = mods["orc"]
    indptr, indices, data = power_law_csr(20000, 800000, seed=5)
    assert int((indptr[1:] - indptr[:-1]).max()) > 1100
    start = orc.shuffled_start(20000, 1, 3)[:6000]
    p, q = pq
    want = orc.walk_csr("SparseOTF", indptr, indices, data, p, q, start, 24, rng=orc.RNG_PHILOX, seed=21)
    eng = mods["WalkEngine"].from_csr(indptr, indices, data)
    got = to_np(eng.walk("SparseOTF", p, q, start, 24, seed=21, flags=flags))
    assert np.array_equal(got, want), first_diff(got, want)
    eng.close()


def test_power_law_hubs_weighted(mods):
    from pecanpy_b200.synth import power_law_csr
    orc = mods["orc"]
    indptr, indices, data = power_law_csr(20000, 800000, seed=6, weighted=True)
    start = orc.shuffled_start(20000, 1, 4)[:4000]
    for flags in (0, 1, 4):
        want = orc.walk_csr("SparseOTF", indptr, indices, data, 0.5, 2.0, start, 16, rng=orc.RNG_PHILOX, seed=22)
        eng = mods["WalkEngine"].from_csr(indptr, indices, data)
        got = to_np(eng.walk("SparseOTF", 0.5, 2.0, start, 16, seed=22, flags=flags))
        assert np.array_equal(got, want), first_diff(got, want)
        eng.close()


@pytest.mark.parametrize("flags", [0, 0x10, 1], ids=["tma", "no-tma", "forced-replay"])
@pytest.mark.parametrize("extend", [False, True], ids=["n2v", "n2v+"])
@pytest.mark.parametrize("n", [256, 1040, 2064])
def test_dense_tma_staging(mods, n, extend, flags):
    """N % 16 == 0 selects the cp.async.bulk (TMA) staged pass 1: several tiles, a short tail tile, ring reuse
    across steps; must equal the oracle and the LDG variant."""
    from pecanpy_b200.synth import dense_weighted
    orc = mods["orc"]
    data, nz = dense_weighted(n, 0.3, seed=40 + n)
    data[n - 3, :] = 0.0; data[:, n - 3] = 0.0                       # one isolated node -> dead ends
    nz = data != 0
    thr = orc.noise_thresholds_dense(data, nz, 0.25) if extend else None
    start = orc.shuffled_start(n, 1, 9)[:300]
    want = orc.walk_dense(data, nz, 0.5, 2.0, start, 12, extend=extend, thr=thr, rng=orc.RNG_PHILOX, seed=31)
    eng = mods["WalkEngine"].from_dense(data, nz)
    if extend:
        eng.set_thresholds(thr)
    got = to_np(eng.walk("DenseOTF", 0.5, 2.0, start, 12, seed=31, extend=extend, flags=flags))
    assert np.array_equal(got, want), first_diff(got, want)
    eng.close()


def test_row_sharding_invariance_and_host_wrapper(mods):
    """Rows are keyed by the GLOBAL row index: any split of the start array gives the same matrix; the
    host-buffer entry point (b2w_walk_host, batched + pipelined) returns the same rows."""
    c = load("hub400_sparseotf_n2v")
    eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    start = c["start"]
    full = to_np(eng.walk("SparseOTF", 4, 0.25, start, 30, seed=3))
    h = 317
    a = to_np(eng.walk("SparseOTF", 4, 0.25, start[:h], 30, seed=3, row0=0))
    b = to_np(eng.walk("SparseOTF", 4, 0.25, start[h:], 30, seed=3, row0=h))
    assert np.array_equal(np.vstack([a, b]), full)
    host = eng.walk_host("SparseOTF", 4, 0.25, start, 30, seed=3, batch_rows=100)
    assert np.array_equal(host, full)
    # the staging buffers are cached in the handle: repeated calls with other shapes must still be right
    host2 = eng.walk_host("SparseOTF", 4, 0.25, start, 30, seed=3, batch_rows=333)
    assert np.array_equal(host2, full)
    long = to_np(eng.walk("SparseOTF", 4, 0.25, start[:50], 70, seed=4))
    assert np.array_equal(eng.walk_host("SparseOTF", 4, 0.25, start[:50], 70, seed=4), long)
    assert eng.last_host_stats["steps"] == int((long[:, -1].astype(np.int64) - 1).sum())
    eng.close()


def test_invalid_graphs_are_rejected(mods):
    capi = mods["capi"]
    indptr = np.array([0, 2, 3], dtype=np.uint32)
    with pytest.raises(capi.B2WError, match="sorted"):
        mods["WalkEngine"].from_csr(indptr, np.array([1, 0, 0], np.uint32), np.ones(3, np.float32))
    with pytest.raises(capi.B2WError, match="weight"):
        mods["WalkEngine"].from_csr(indptr, np.array([0, 1, 0], np.uint32), np.array([1, -1, 1], np.float32))


def test_drop_in_classes_match_oracle(mods):
    """The reference-shaped classes: from_mat + simulate_walks (list of id lists)."""
    from pecanpy_b200 import pecanpy as b2
    orc = mods["orc"]
    MAT = np.array([[0, 1, 0, 0, 0], [1, 0, 1, 0, 0], [0, 1, 0, 1, 1], [0, 0, 1, 0, 1], [0, 0, 1, 1, 0]])
    IDS = list("abcde")
    for name in ["SparseOTF", "PreComp", "DenseOTF", "FirstOrderUnweighted", "PreCompFirstOrder"]:
        g = getattr(b2, name).from_mat(MAT, IDS, p=1, q=1, random_state=0)
        walks = g.simulate_walks(2, 3)
        start = orc.shuffled_start(5, 2, 0)
        if name == "DenseOTF":
            want = orc.walk_dense(g.data, g.nonzero, 1, 1, start, 3, rng=orc.RNG_PHILOX, seed=0)
        else:
            alias = None
            if name == "PreComp":
                alias = orc.alias_build(g.indptr, g.indices, g.data, 1, 1)
            elif name == "PreCompFirstOrder":
                j, q = orc.alias_build_first_order(g.indptr, g.indices, g.data)
                alias = (None, j, q)
            want = orc.walk_csr(name, g.indptr, g.indices, g.data, 1, 1, start, 3, alias=alias,
                                rng=orc.RNG_PHILOX, seed=0)
        assert walks == [[IDS[i] for i in row[: row[-1]]] for row in want], name
        g.release()


# ---------------------------------------------------------------------------------------------- thresholds
def _same_f32(a, b):
    nan = np.isnan(a)
    return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


THR_FIXTURES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                      if "_ext" in os.path.basename(f))


@pytest.mark.parametrize("name", THR_FIXTURES)
def test_noise_thresholds_equal_reference_fixture(mods, name):
    """b2w_noise_thresholds vs the `thr` array the unmodified reference computed (get_noise_thresholds)."""
    c = load(name)
    gamma = float(c["gamma"]) if "gamma" in c else 0.0
    if "dense" in c:
        eng = mods["WalkEngine"].from_dense(c["dense"], c["nonzero"])
    else:
        eng = mods["WalkEngine"].from_csr(c["indptr"], c["indices"], c["data"])
    got = eng.compute_thresholds(gamma).cpu().numpy()
    assert _same_f32(got, c["thr"].astype(np.float32))
    eng.close()


@pytest.mark.parametrize("gamma", [0.0, 0.25, -0.7, 3.3])
def test_noise_thresholds_csr_vs_numpy(mods, gamma):
    """Rows of every length class of NumPy's pairwise summation (< 8, <= 128, recursive), empty rows, a hub."""
    rng = np.random.default_rng(5)
    n = 6000
    deg = np.concatenate([rng.integers(0, 40, n - 60), rng.integers(100, 700, 50), [0, 1, 7, 8, 9, 127, 128, 129, 5000, 3]])
    # any sorted duplicate-free rows will do: the thresholds only read the weights
    indptr = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(deg, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(n, int(d), replace=False)) for d in deg]).astype(np.uint32)
    data = (np.float32(0.01) + np.float32(0.99) * rng.random(indices.size, dtype=np.float32)).astype(np.float32)
    eng = mods["WalkEngine"].from_csr(indptr, indices, data)
    got = eng.compute_thresholds(gamma).cpu().numpy()
    want = mods["orc"].noise_thresholds_csr(indptr, data, gamma)
    assert _same_f32(got, want)
    eng.close()


@pytest.mark.parametrize("n", [333, 1024])
@pytest.mark.parametrize("gamma", [0.0, 0.5, -2.0])
def test_noise_thresholds_dense_vs_numpy(mods, n, gamma):
    rng = np.random.default_rng(6)
    nz = rng.random((n, n)) < 0.4
    nz[3] = False
    nz[4] = False; nz[4, 10] = True
    nz[5] = True
    data = np.where(nz, 0.01 + 0.99 * rng.random((n, n)), 0.0)
    eng = mods["WalkEngine"].from_dense(data, nz)
    got = eng.compute_thresholds(gamma).cpu().numpy()
    want = mods["orc"].noise_thresholds_dense(data, nz, gamma)
    assert _same_f32(got, want)
    eng.close()


def test_dropin_extend_uses_device_thresholds(mods):
    """SparseOTF(extend=True): thresholds come from the GPU kernel and the walks equal the oracle's."""
    from pecanpy_b200 import pecanpy as node2vec
    c = load("w200_sparseotf_ext_g05")
    g = node2vec.SparseOTF(p=float(c["p"]), q=float(c["q"]), extend=True, gamma=float(c["gamma"]), random_state=11)
    g.indptr, g.indices, g.data = c["indptr"], c["indices"], c["data"]
    g.set_node_ids(None, implicit_ids=True, num_nodes=c["indptr"].size - 1)
    thr = g.get_noise_thresholds()
    assert _same_f32(thr, c["thr"].astype(np.float32))
    mat = g.simulate_walks_array(2, 15)
    start = mods["orc"].shuffled_start(g.num_nodes, 2, 11)
    want = mods["orc"].walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], float(c["p"]), float(c["q"]), start, 15,
                                extend=True, thr=c["thr"], rng=mods["orc"].RNG_PHILOX, seed=11)
    assert np.array_equal(mat, want), first_diff(mat, want)


def test_precomp_dropin_right_after_table_build(mods):
    """Round-1 advice: preprocess (alias build, tens of ms) followed at once by the host-buffer walk on the
    library's own streams -- the tables must be complete when the walk kernel reads them.  Also the lazily
    copied alias_j / alias_q attributes of the reference surface."""
    from pecanpy_b200 import pecanpy as b2
    from pecanpy_b200.synth import erdos_renyi_csr
    orc = mods["orc"]
    indptr, indices, data = erdos_renyi_csr(20000, 600000, seed=8, weighted=True)
    for extend in (False, True):
        g = b2.PreComp(p=0.25, q=4, extend=extend, gamma=0.3, random_state=5)
        g.indptr, g.indices, g.data = indptr, indices, data
        g.set_node_ids(None, implicit_ids=True, num_nodes=indptr.size - 1)
        mat = g.simulate_walks_array(1, 20)                 # builds thresholds + tables, walks immediately
        thr = orc.noise_thresholds_csr(indptr, data, 0.3) if extend else None
        alias = orc.alias_build(indptr, indices, data, 0.25, 4, extend, thr)
        start = orc.shuffled_start(indptr.size - 1, 1, 5)
        k = 6000
        want = orc.walk_csr("PreComp", indptr, indices, data, 0.25, 4, start[:k], 20, alias=alias, rng=orc.RNG_PHILOX, seed=5)
        assert np.array_equal(mat[:k], want), first_diff(mat[:k], want)
        assert np.array_equal(g.alias_j, alias[1]) and np.array_equal(g.alias_q.view(np.uint32), alias[2].view(np.uint32))
        assert np.array_equal(g.alias_indptr, alias[0])
        g.release()


def test_dropin_rebuilds_device_copy_when_the_graph_changes(mods):
    """Round-1 advice: the cached engine must not outlive the arrays it was made from (the reference rebuilds its
    closures on every call, pecanpy.py:143-144)."""
    from pecanpy_b200 import pecanpy as b2
    orc = mods["orc"]
    a = load("w200_sparseotf_n2v")
    b = load("hub400_sparseotf_n2v")
    g = b2.PreComp(p=0.5, q=2, random_state=3)
    for c in (a, b, a):
        g.indptr, g.indices, g.data = c["indptr"], c["indices"], c["data"]
        g.set_node_ids(None, implicit_ids=True, num_nodes=c["indptr"].size - 1)
        mat = g.simulate_walks_array(2, 12)
        start = orc.shuffled_start(g.num_nodes, 2, 3)
        alias = orc.alias_build(c["indptr"], c["indices"], c["data"], 0.5, 2)
        want = orc.walk_csr("PreComp", c["indptr"], c["indices"], c["data"], 0.5, 2, start, 12, alias=alias,
                            rng=orc.RNG_PHILOX, seed=3)
        assert np.array_equal(mat, want), first_diff(mat, want)
    # re-weighting the same topology and changing p: tables must follow
    g.data = (g.data * np.float32(1.5)).astype(np.float32)
    g.p = 2.0
    mat = g.simulate_walks_array(1, 10)
    alias = orc.alias_build(g.indptr, g.indices, g.data, 2.0, 2)
    want = orc.walk_csr("PreComp", g.indptr, g.indices, g.data, 2.0, 2, orc.shuffled_start(g.num_nodes, 1, 3), 10,
                        alias=alias, rng=orc.RNG_PHILOX, seed=3)
    assert np.array_equal(mat, want)
    s = b2.SparseOTF(p=1, q=1, extend=True, gamma=0.0, random_state=1)
    s.indptr, s.indices, s.data = a["indptr"], a["indices"], a["data"]
    s.set_node_ids(None, implicit_ids=True, num_nodes=a["indptr"].size - 1)
    t0 = s.get_noise_thresholds()
    s.gamma = 1.0                                            # thresholds depend on gamma
    m1 = s.simulate_walks_array(1, 8)
    thr = orc.noise_thresholds_csr(a["indptr"], a["data"], 1.0)
    want = orc.walk_csr("SparseOTF", a["indptr"], a["indices"], a["data"], 1, 1, orc.shuffled_start(s.num_nodes, 1, 1), 8,
                        extend=True, thr=thr, rng=orc.RNG_PHILOX, seed=1)
    assert np.array_equal(m1, want) and not np.array_equal(t0, thr)
    g.release(); s.release()


@pytest.mark.gpu
def test_random_walks_entry_keeps_the_reference_signature():
    """Base._random_walks(tot, L, random_state, start, has_nbrs, move_forward, progress) -- reference
    pecanpy.py:164-210, the call SURVEY 8d times -- returns the matrix simulate_walks_array builds from the same start
    array and seed; the two callbacks are accepted and ignored."""
    from pecanpy_b200 import pecanpy as pp
    from pecanpy_b200.synth import power_law_csr
    indptr, indices, data = power_law_csr(3000, 30000, seed=4)
    g = pp.SparseOTF(p=0.5, q=2, random_state=7)
    g.indptr, g.indices, g.data = indptr, indices, data
    g.set_node_ids(None, implicit_ids=True, num_nodes=3000)
    want = g.simulate_walks_array(2, 20)
    start = g._start_nodes(2)
    got = g._random_walks(start.size, 20, 7, start, g.get_has_nbrs(), None, None)
    assert got.dtype == np.uint32 and np.array_equal(got, want)
    part = g._random_walks(100, 20, 7, start)                  # a prefix of the jobs: the same rows (Philox by row)
    assert np.array_equal(part, want[:100])
    g.release()
