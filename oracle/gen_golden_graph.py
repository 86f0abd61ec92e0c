"""Generate tests/golden/graph_*.npz by running the UNMODIFIED reference's ingest (PecanPy @ /root/reference,
src/pecanpy/graph.py): AdjlstGraph.read -> to_csr / to_dense (graph.py:160-362), SparseGraph / DenseGraph.read_edg
and .from_mat (graph.py:423-528, 587-657).  Run in the build container only: ``python oracle/gen_golden_graph.py``.

Each fixture stores the TEXT of an edge-list file (so the tests can re-create the file byte for byte), the read
parameters, and what the reference made of it: node ids in index order, CSR arrays, the dense matrix and its mask.
The literal matrices of the reference's own test/test_graph.py:16-78 go through from_mat the same way and are stored
next to the INDPTR/INDICES/DATA literals that file asserts (so the fixture pins both).
"""
import os
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_stubs"), "/root/reference/src"]

import numpy as np  # noqa: E402
from pecanpy.graph import AdjlstGraph, DenseGraph, SparseGraph  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ref_read(text, weighted, directed, delimiter):
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "g.edg")
        with open(path, "w", newline="", encoding="utf-8") as f:
            f.write(text)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            adj = AdjlstGraph()
            adj.read(path, weighted, directed, delimiter)
            indptr, indices, data = adj.to_csr()
            dense = adj.to_dense()
            sp = SparseGraph()
            sp.read_edg(path, weighted, directed, delimiter)
            dn = DenseGraph()
            dn.read_edg(path, weighted, directed, delimiter)
    assert sp.nodes == adj.nodes and dn.nodes == adj.nodes
    assert np.array_equal(sp.indptr, indptr) and np.array_equal(sp.indices, indices) and np.array_equal(sp.data, data)
    assert np.array_equal(dn.data, dense)
    return dict(nodes=np.array(adj.nodes), indptr=indptr, indices=indices, data=data,
                dense=np.asarray(dn.data), nonzero=np.asarray(dn.nonzero), num_edges=sp.num_edges)


def edge_text(seed, lines, n_ids, weighted, delimiter, messy):
    """Duplicates, both orientations, self loops, weights <= 0 (dropped, ids not registered), scientific notation,
    ids that need stripping, CRLF line ends, extra columns on unweighted lines, no trailing newline."""
    rng = np.random.default_rng(seed)
    weights = ["0.5", "1.5", "2", "-1.0", "0", "1e-3", "3.25E+2", " 7.125 ", ".5", "5.", "+4", "0.1234567890123456789",
               "1e-40", "16777217"]
    out = []
    for i in range(lines):
        a, b = rng.integers(0, n_ids, size=2)
        ida = f" n{a} " if messy and i % 7 == 0 else f"n{a}"
        eol = "\r\n" if messy and i % 5 == 0 else "\n"
        if weighted:
            out.append(f"{ida}{delimiter}n{b}{delimiter}{weights[int(rng.integers(0, len(weights)))]}{eol}")
        else:
            extra = f"{delimiter}ignored" if messy and i % 11 == 0 else ""
            out.append(f"{ida}{delimiter}n{b}{extra}{eol}")
    out.append(f"last{delimiter}line" + (f"{delimiter}1.0" if weighted else ""))      # no trailing newline
    return "".join(out)


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {}
    k = 0
    for weighted in (False, True):
        for directed in (False, True):
            for delimiter, messy in (("\t", True), (",", False)):
                text = edge_text(100 + k, 600, 50, weighted, delimiter, messy)
                rec = ref_read(text, weighted, directed, delimiter)
                rec.update(text=np.array(text), weighted=weighted, directed=directed, delimiter=np.array(delimiter))
                cases[f"edg{k}_w{int(weighted)}_d{int(directed)}"] = rec
                k += 1
    karate = open("/root/reference/demo/karate.edg").read()
    rec = ref_read(karate, False, False, "\t")
    rec.update(text=np.array(karate), weighted=False, directed=False, delimiter=np.array("\t"))
    cases["karate"] = rec
    for name, rec in cases.items():
        np.savez_compressed(os.path.join(OUT, f"graph_{name}.npz"), **rec)
        print(f"graph_{name}: {len(rec['nodes'])} nodes, nnz {rec['indices'].size}")

    # the reference's own literals (test/test_graph.py:16-78)
    MAT = np.array([[0, 1, 1], [1, 0, 0], [1, 0, 0]], dtype=float)
    MAT2 = np.array([[0, 1, 0, 0, 0], [1, 0, 1, 1, 0], [0, 1, 0, 0, 0], [0, 1, 0, 0, 1], [0, 0, 0, 1, 0]], dtype=float)
    MAT3 = np.array([[0, 1, 0, 0], [1, 0, 0, 1], [0, 0, 0, 0], [0, 1, 1, 0]])
    lit = dict(
        INDPTR1=np.array([0, 2, 3, 4], dtype=np.uint32), INDICES1=np.array([1, 2, 0, 0], dtype=np.uint32),
        INDPTR2=np.array([0, 1, 4, 5, 7, 8], dtype=np.uint32), INDICES2=np.array([1, 0, 2, 3, 1, 1, 4, 3], dtype=np.uint32),
        INDPTR3=np.array([0, 1, 3, 3, 5], dtype=np.uint32), INDICES3=np.array([1, 0, 3, 1, 2], dtype=np.uint32))
    rec = {}
    for i, (mat, ids) in enumerate([(MAT, list("abc")), (MAT2, list("abcde")), (MAT3, list("abcd"))], start=1):
        sp = SparseGraph.from_mat(mat, ids)
        dn = DenseGraph.from_mat(mat, ids)
        assert np.array_equal(sp.indptr, lit[f"INDPTR{i}"]) and np.array_equal(sp.indices, lit[f"INDICES{i}"])
        rec.update({f"mat{i}": mat, f"ids{i}": np.array(ids), f"indptr{i}": sp.indptr,
                    f"indices{i}": sp.indices, f"data{i}": sp.data, f"dense{i}": np.asarray(dn.data),
                    f"nonzero{i}": np.asarray(dn.nonzero), f"num_edges{i}": sp.num_edges, f"density{i}": sp.density})
    np.savez_compressed(os.path.join(OUT, "graph_testgraph_literals.npz"), **rec)
    print("graph_testgraph_literals: 3 matrices")


if __name__ == "__main__":
    main()
