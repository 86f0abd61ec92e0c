"""Generate tests/golden/*.npz by running the UNMODIFIED reference (PecanPy @ /root/reference).

Run in the build container only (the reference is a Python/Numba package that cannot
travel to the GPU box):  ``python oracle/gen_golden.py``.

The reference is imported from /root/reference/src with three import stubs
(oracle/ref_stubs: gensim, numba_progress, nptyping -- none touches the walk arithmetic)
and executed at ``numba.set_num_threads(1)``, the only configuration in which its seeded
walks are reproducible (pecanpy.py:51-55, test/test_walk.py:8).

Every fixture stores the graph arrays, the parameters, the shuffled start array and the raw
``uint32[tot, L+2]`` matrix returned by ``Base._random_walks`` (pecanpy.py:164-210), plus --
for PreComp -- the alias tables of ``preprocess_transition_probs`` (pecanpy.py:442-507) and --
for the probability fixtures -- outputs of ``get_(extended_)normalized_probs``.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_stubs"), "/root/reference/src"]

import numba  # noqa: E402
import numpy as np  # noqa: E402

numba.set_num_threads(1)

from numba_progress import ProgressBar  # noqa: E402  (stub)
from pecanpy import pecanpy  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def ref_walk_matrix(g, num_walks, walk_length):
    """simulate_walks (pecanpy.py:116-157) up to, not including, the id mapping."""
    g._preprocess_transition_probs()
    nodes = np.array(range(g.num_nodes), dtype=np.uint32)
    start = np.concatenate([nodes] * num_walks)
    np.random.seed(g.random_state)
    np.random.shuffle(start)
    mf = g.get_move_forward()
    hn = g.get_has_nbrs()
    with ProgressBar(total=start.size, disable=True) as progress:
        mat = g._random_walks(start.size, walk_length, g.random_state, start, hn, mf, progress)
    return start, mat


def sym_weighted_graph(n, m, seed, weighted=True, isolated=0):
    """Small undirected graph as a dense matrix (float64), optional isolated tail nodes."""
    rng = np.random.default_rng(seed)
    mat = np.zeros((n, n))
    live = n - isolated
    a = rng.integers(0, live, size=m)
    b = rng.integers(0, live, size=m)
    keep = a != b
    a, b = a[keep], b[keep]
    w = (np.float32(0.01) + np.float32(0.99) * rng.random(a.size, dtype=np.float32)).astype(np.float64) \
        if weighted else np.ones(a.size)
    mat[a, b] = w
    mat[b, a] = w
    return mat


def hub_graph(n, seed):
    """Power-law-ish weighted undirected graph with a few high-degree hubs."""
    rng = np.random.default_rng(seed)
    theta = (np.arange(n) + 2.0) ** (-0.9)
    theta /= theta.sum()
    m = 12 * n
    a = rng.choice(n, size=m, p=theta)
    b = rng.choice(n, size=m, p=theta)
    keep = a != b
    a, b = a[keep], b[keep]
    mat = np.zeros((n, n))
    w = (np.float32(0.05) + rng.random(a.size, dtype=np.float32)).astype(np.float64)
    mat[a, b] = w
    mat[b, a] = w
    return mat


def directed_with_dead_ends(n, m, seed):
    rng = np.random.default_rng(seed)
    mat = np.zeros((n, n))
    a = rng.integers(0, n - 5, size=m)          # last 5 nodes never have out-edges
    b = rng.integers(0, n, size=m)
    keep = a != b
    w = (np.float32(0.1) + rng.random(m, dtype=np.float32)).astype(np.float64)
    mat[a[keep], b[keep]] = w[keep]
    return mat


def save_case(name, cls_name, mat, p, q, extend, gamma, seed, num_walks, walk_length, ids=None):
    ids = ids or [str(i) for i in range(mat.shape[0])]
    cls = getattr(pecanpy, cls_name)
    g = cls.from_mat(mat, ids, p=p, q=q, extend=extend, gamma=gamma, random_state=seed)
    start, walks = ref_walk_matrix(g, num_walks, walk_length)
    rec = dict(mode=cls_name, p=float(p), q=float(q), extend=bool(extend), gamma=float(gamma),
               seed=int(seed), num_walks=int(num_walks), walk_length=int(walk_length),
               start=start, walks=walks)
    if cls_name == "DenseOTF":
        rec.update(dense=np.asarray(g.data), nonzero=np.asarray(g.nonzero))
        if extend:
            rec["thr"] = g.get_noise_thresholds()
    else:
        rec.update(indptr=g.indptr, indices=g.indices, data=g.data)
        if extend:
            rec["thr"] = g.get_noise_thresholds()
        if cls_name == "PreComp":
            rec.update(alias_j=g.alias_j, alias_q=g.alias_q, alias_indptr=g.alias_indptr,
                       alias_dim=g.alias_dim)
        if cls_name == "PreCompFirstOrder":
            rec.update(alias_j=g.alias_j, alias_q=g.alias_q)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    steps = int((walks[:, -1].astype(np.int64) - 1).sum())
    print(f"{name}: {cls_name} n={mat.shape[0]} rows={walks.shape[0]} steps={steps}")


def karate_matrix():
    edges = np.loadtxt("/root/reference/demo/karate.edg", dtype=np.int64)
    # first-seen node order, as AdjlstGraph.read does (graph.py:222-236, 270-305)
    order = {}
    for a, b in edges:
        for x in (a, b):
            if x not in order:
                order[x] = len(order)
    n = len(order)
    mat = np.zeros((n, n))
    for a, b in edges:
        mat[order[a], order[b]] = 1.0
        mat[order[b], order[a]] = 1.0
    ids = [None] * n
    for k, v in order.items():
        ids[v] = str(k)
    return mat, ids


def probs_fixture():
    """Outputs of the reference's njit probability functions on a hub graph."""
    mat = hub_graph(600, 11)
    ids = [str(i) for i in range(mat.shape[0])]
    g = pecanpy.SparseOTF.from_mat(mat, ids, p=0.3, q=0.7, extend=True, gamma=0.5)
    thr = g.get_noise_thresholds()
    rng = np.random.default_rng(5)
    deg = g.indptr[1:] - g.indptr[:-1]
    pairs = []
    hubs = np.argsort(-deg.astype(np.int64))[:8]
    for cur in list(hubs) + list(rng.integers(0, 600, size=24)):
        if deg[cur] == 0:
            continue
        nb = g.indices[g.indptr[cur]:g.indptr[cur + 1]]
        for prev in rng.choice(nb, size=min(3, nb.size), replace=False):
            pairs.append((int(cur), int(prev)))
    rec = dict(indptr=g.indptr, indices=g.indices, data=g.data, thr=thr, gamma=0.5,
               pairs=np.array(pairs, dtype=np.int64))
    out_n2v, out_ext, out_first, offs = [], [], [], [0]
    for (p_, q_) in [(0.3, 0.7)]:
        for cur, prev in pairs:
            a = g.get_normalized_probs(g.data, g.indices, g.indptr, p_, q_, cur, prev, None)
            b = g.get_extended_normalized_probs(g.data, g.indices, g.indptr, p_, q_, cur, prev, thr)
            c = g.get_normalized_probs(g.data, g.indices, g.indptr, p_, q_, cur, None, None)
            out_n2v.append(a); out_ext.append(b); out_first.append(c)
            offs.append(offs[-1] + a.size)
    rec.update(p=0.3, q=0.7, probs_n2v=np.concatenate(out_n2v), probs_ext=np.concatenate(out_ext),
               probs_first=np.concatenate(out_first), offsets=np.array(offs, dtype=np.int64))
    # dense flavour of the same pairs
    gd = pecanpy.DenseOTF.from_mat(mat, ids, p=0.3, q=0.7, extend=True, gamma=0.5)
    thr_d = gd.get_noise_thresholds()
    dn, de, doffs = [], [], [0]
    for cur, prev in pairs[:20]:
        a = gd.get_normalized_probs(gd.data, gd.nonzero, 0.3, 0.7, cur, prev, None)
        b = gd.get_extended_normalized_probs(gd.data, gd.nonzero, 0.3, 0.7, cur, prev, thr_d)
        dn.append(a); de.append(b); doffs.append(doffs[-1] + a.size)
    rec.update(dense_thr=thr_d, dense_probs_n2v=np.concatenate(dn), dense_probs_ext=np.concatenate(de),
               dense_offsets=np.array(doffs, dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, "probs_hub600.npz"), **rec)
    print("probs_hub600:", len(pairs), "pairs; max degree", int(deg.max()))


def main():
    # (1) the reference's own known-answer configuration (test/test_walk.py:10-98)
    MAT = np.array([[0, 1, 0, 0, 0], [1, 0, 1, 0, 0], [0, 1, 0, 1, 1], [0, 0, 1, 0, 1], [0, 0, 1, 1, 0]])
    for cls in ["FirstOrderUnweighted", "PreCompFirstOrder", "PreComp", "SparseOTF", "DenseOTF"]:
        save_case(f"testwalk_{cls}", cls, MAT, 1, 1, False, 0, 0, 2, 3, ids=list("abcde"))

    # (2) BASELINE config #1: karate, 10 x 80
    kmat, kids = karate_matrix()
    save_case("karate_sparseotf_p1_q1", "SparseOTF", kmat, 1, 1, False, 0, 0, 10, 80, ids=kids)
    save_case("karate_sparseotf_p05_q2", "SparseOTF", kmat, 0.5, 2, False, 0, 0, 10, 80, ids=kids)
    save_case("karate_sparseotf_p03_q07", "SparseOTF", kmat, 0.3, 0.7, False, 0, 1, 10, 80, ids=kids)
    save_case("karate_precomp_p025_q4", "PreComp", kmat, 0.25, 4, False, 0, 2, 10, 80, ids=kids)
    save_case("karate_denseotf_p05_q2", "DenseOTF", kmat, 0.5, 2, False, 0, 3, 10, 80, ids=kids)
    save_case("karate_firstorder", "FirstOrderUnweighted", kmat, 1, 1, False, 0, 4, 10, 80, ids=kids)

    # (3) weighted graphs, p,q not powers of two, node2vec+, isolated nodes, dead ends
    w200 = sym_weighted_graph(200, 1500, 21, weighted=True, isolated=3)
    save_case("w200_sparseotf_n2v", "SparseOTF", w200, 0.3, 0.7, False, 0, 5, 4, 40)
    save_case("w200_sparseotf_ext_g0", "SparseOTF", w200, 0.5, 2, True, 0, 6, 4, 40)
    save_case("w200_sparseotf_ext_g05", "SparseOTF", w200, 0.3, 3.0, True, 0.5, 7, 4, 40)
    save_case("w200_precomp_n2v", "PreComp", w200, 0.25, 4, False, 0, 8, 4, 40)
    save_case("w200_precomp_ext", "PreComp", w200, 0.7, 0.3, True, 0.25, 9, 4, 40)
    save_case("w200_precompfirstorder", "PreCompFirstOrder", w200, 1, 1, False, 0, 10, 4, 40)
    save_case("w200_denseotf_n2v", "DenseOTF", w200, 0.3, 0.7, False, 0, 11, 3, 30)
    save_case("w200_denseotf_ext", "DenseOTF", w200, 0.5, 2, True, 0.5, 12, 3, 30)
    dd = directed_with_dead_ends(150, 900, 31)
    save_case("dir150_sparseotf_deadends", "SparseOTF", dd, 2, 0.5, False, 0, 13, 4, 30)
    save_case("dir150_precomp_deadends", "PreComp", dd, 2, 0.5, False, 0, 14, 4, 30)
    save_case("dir150_denseotf_deadends", "DenseOTF", dd, 2, 0.5, False, 0, 15, 3, 20)
    hub = hub_graph(400, 41)
    save_case("hub400_sparseotf_n2v", "SparseOTF", hub, 4, 0.25, False, 0, 16, 2, 40)
    save_case("hub400_sparseotf_ext", "SparseOTF", hub, 0.5, 2, True, 0, 17, 2, 40)
    uhub = (hub != 0) * 1.0
    save_case("uhub400_sparseotf_n2v", "SparseOTF", uhub, 4, 0.25, False, 0, 18, 2, 40)

    # (4) probability vectors
    probs_fixture()


if __name__ == "__main__":
    main()
