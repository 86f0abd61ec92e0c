"""Pin the CPU oracle (oracle/walk_oracle.c) against the reference itself.

tests/golden/*.npz were written by oracle/gen_golden.py, which imports the unmodified
reference (PecanPy, Numba, 1 thread).  The oracle replays the same MT19937 word stream
(ORC_RNG_WORDS) and must reproduce every walk matrix, alias table and probability vector
bit for bit -- including the reference's own known-answer vectors (test/test_walk.py:21-82).
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WALK_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                    if not os.path.basename(f).startswith(("probs_", "graph_")))

# test/test_walk.py:21-82 (ids a..e -> 0..4); PreCompFirstOrder row uses pecanpy.PreComp (:89)
REFERENCE_KNOWN_ANSWERS = {
    "FirstOrderUnweighted": ["cbcd", "dcde", "edcb", "edcb", "baba", "babc", "cede", "dcbc", "abcd", "abcb"],
    "PreComp": ["cded", "dcde", "edce", "edec", "bcec", "bcdc", "cded", "dced", "abab", "abce"],
    "SparseOTF": ["cded", "decd", "eced", "eced", "bcec", "babc", "cede", "dece", "abcb", "abcd"],
    "DenseOTF": ["cded", "decd", "eced", "eced", "bcec", "babc", "cede", "dece", "abcb", "abcd"],
}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def run_oracle_words(c):
    mode = str(c["mode"])
    L = int(c["walk_length"])
    seed = int(c["seed"])
    start = c["start"]
    words = orc.mt_words(seed, 64 + 6 * start.size * L)
    kw = dict(extend=bool(c["extend"]), thr=c.get("thr"), rng=orc.RNG_WORDS, words=words)
    if mode == "DenseOTF":
        return orc.walk_dense(c["dense"], c["nonzero"], float(c["p"]), float(c["q"]), start, L, **kw)
    alias = None
    if mode == "PreComp":
        alias = (c["alias_indptr"], c["alias_j"], c["alias_q"])
    elif mode == "PreCompFirstOrder":
        alias = (None, c["alias_j"], c["alias_q"])
    return orc.walk_csr(mode, c["indptr"], c["indices"], c["data"], float(c["p"]), float(c["q"]),
                        start, L, alias=alias, **kw)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert orc.philox4x32_10(ctr, key).tolist() == want


def test_mt_words_reproduce_random_sample():
    w = orc.mt_words(123, 20).astype(np.uint64)
    u = ((w[0::2] >> 5) * 67108864.0 + (w[1::2] >> 6)) / 9007199254740992.0
    assert np.array_equal(u, np.random.RandomState(123).random_sample(10))


def test_shuffled_start_matches_fixture():
    c = load("karate_sparseotf_p1_q1")
    n = c["indptr"].size - 1
    assert np.array_equal(orc.shuffled_start(n, int(c["num_walks"]), int(c["seed"])), c["start"])


@pytest.mark.parametrize("name", WALK_CASES)
def test_oracle_replays_reference_walks(name):
    c = load(name)
    got = run_oracle_words(c)
    assert got.dtype == np.uint32 and got.shape == c["walks"].shape
    assert np.array_equal(got, c["walks"]), f"{name}: first bad row {np.argwhere((got != c['walks']).any(1))[:3].ravel()}"


@pytest.mark.parametrize("mode", sorted(REFERENCE_KNOWN_ANSWERS))
def test_reference_test_walk_vectors(mode):
    """The literal expected walks of the reference's test/test_walk.py."""
    c = load("testwalk_" + mode)
    got = run_oracle_words(c)
    walks = ["".join("abcde"[i] for i in row[: row[-1]]) for row in got]
    assert walks == REFERENCE_KNOWN_ANSWERS[mode]


@pytest.mark.parametrize("name", ["karate_sparseotf_p1_q1", "karate_sparseotf_p05_q2", "karate_sparseotf_p03_q07",
                                  "w200_sparseotf_ext_g05", "hub400_sparseotf_n2v"])
def test_feed_regime_equals_word_regime(name):
    """R1: per-row uniform feed == sequential MT replay when every walker draws exactly L doubles."""
    c = load(name)
    if (c["walks"][:, -1] != int(c["walk_length"]) + 1).any():
        pytest.skip("graph has dead ends; feed regime not applicable")
    L = int(c["walk_length"])
    feed = orc.mt_uniform_feed(int(c["seed"]), c["start"].size, L)
    got = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], float(c["p"]), float(c["q"]),
                       c["start"], L, extend=bool(c["extend"]), thr=c.get("thr"), rng=orc.RNG_FEED, feed=feed)
    assert np.array_equal(got, c["walks"])


@pytest.mark.parametrize("name", ["karate_precomp_p025_q4", "w200_precomp_n2v", "w200_precomp_ext",
                                  "dir150_precomp_deadends"])
def test_alias_tables_bit_exact(name):
    c = load(name)
    aip, j, q = orc.alias_build(c["indptr"], c["indices"], c["data"], float(c["p"]), float(c["q"]),
                                bool(c["extend"]), c.get("thr"))
    assert np.array_equal(aip, c["alias_indptr"])
    assert np.array_equal(j, c["alias_j"])
    assert np.array_equal(q.view(np.uint32), c["alias_q"].view(np.uint32))


def test_first_order_alias_tables_bit_exact():
    c = load("w200_precompfirstorder")
    j, q = orc.alias_build_first_order(c["indptr"], c["indices"], c["data"])
    assert np.array_equal(j, c["alias_j"])
    assert np.array_equal(q.view(np.uint32), c["alias_q"].view(np.uint32))


def test_noise_thresholds_match_reference():
    c = load("w200_sparseotf_ext_g05")
    thr = orc.noise_thresholds_csr(c["indptr"], c["data"], float(c["gamma"]))
    assert np.array_equal(thr.view(np.uint32), c["thr"].view(np.uint32))
    d = load("w200_denseotf_ext")
    thr = orc.noise_thresholds_dense(d["dense"], d["nonzero"].astype(bool), float(d["gamma"]))
    assert np.array_equal(thr.view(np.uint32), d["thr"].view(np.uint32))


def test_probability_vectors_bit_exact():
    c = load("probs_hub600")
    offs = c["offsets"]
    for k, (cur, prev) in enumerate(c["pairs"]):
        sl = slice(offs[k], offs[k + 1])
        a = orc.sparse_probs(c["indptr"], c["indices"], c["data"], 0.3, 0.7, int(cur), int(prev))
        b = orc.sparse_probs(c["indptr"], c["indices"], c["data"], 0.3, 0.7, int(cur), int(prev), True, c["thr"])
        f = orc.sparse_probs(c["indptr"], c["indices"], c["data"], 0.3, 0.7, int(cur), None)
        assert np.array_equal(a.view(np.uint32), c["probs_n2v"][sl].view(np.uint32)), (cur, prev)
        assert np.array_equal(b.view(np.uint32), c["probs_ext"][sl].view(np.uint32)), (cur, prev)
        assert np.array_equal(f.view(np.uint32), c["probs_first"][sl].view(np.uint32)), (cur, prev)


def test_dense_probability_vectors_bit_exact():
    c = load("probs_hub600")
    n = c["indptr"].size - 1
    dense = np.zeros((n, n))
    for i in range(n):
        s, e = c["indptr"][i], c["indptr"][i + 1]
        dense[i, c["indices"][s:e]] = c["data"][s:e]
    nonzero = dense != 0
    offs = c["dense_offsets"]
    for k, (cur, prev) in enumerate(c["pairs"][:20]):
        sl = slice(offs[k], offs[k + 1])
        a, _ = orc.dense_probs(dense, nonzero, 0.3, 0.7, int(cur), int(prev))
        b, _ = orc.dense_probs(dense, nonzero, 0.3, 0.7, int(cur), int(prev), True, c["dense_thr"])
        assert np.array_equal(a.view(np.uint64), c["dense_probs_n2v"][sl].view(np.uint64)), (cur, prev)
        assert np.array_equal(b.view(np.uint64), c["dense_probs_ext"][sl].view(np.uint64)), (cur, prev)


def test_philox_regime_is_thread_count_invariant():
    c = load("hub400_sparseotf_n2v")
    a = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], 4, 0.25, c["start"], 20,
                     rng=orc.RNG_PHILOX, seed=7, nthreads=1)
    b = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], 4, 0.25, c["start"], 20,
                     rng=orc.RNG_PHILOX, seed=7, nthreads=4)
    assert np.array_equal(a, b)
    # sharding invariance: rows keyed by the global row index
    h = c["start"].size // 2
    lo = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], 4, 0.25, c["start"][:h], 20,
                      rng=orc.RNG_PHILOX, seed=7, row0=0)
    hi = orc.walk_csr("SparseOTF", c["indptr"], c["indices"], c["data"], 4, 0.25, c["start"][h:], 20,
                      rng=orc.RNG_PHILOX, seed=7, row0=h)
    assert np.array_equal(np.vstack([lo, hi]), a)
