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


# ------------------------------------------------------------------------------------------ native parser
def _reference_read_lines(path, weighted, delimiter="\t"):
    """AdjlstGraph.read restated line by line (graph.py:160-193, 217-236, 258-305): ids, endpoints, weights in
    file order; lines with weight <= 0 are ignored and do not register their ids."""
    ids, src, dst, w = {}, [], [], []
    with open(path, encoding="utf-8") as f:
        for line in f:
            if not line.strip():
                continue
            terms = line.strip().split(delimiter)
            id1, id2 = terms[0].strip(), terms[1].strip()
            weight = 1.0
            if weighted:
                if len(terms) != 3:
                    raise ValueError("Expecting three columns")
                weight = float(terms[-1])
            if weight <= 0:
                continue
            for x in (id1, id2):
                if x not in ids:
                    ids[x] = len(ids)
            src.append(ids[id1]); dst.append(ids[id2]); w.append(weight)
    return list(ids), np.array(src, np.int64), np.array(dst, np.int64), np.array(w, np.float64)


def _same_parse(a, b):
    return a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))


@pytest.mark.parametrize("weighted", [False, True])
def test_native_parser_equals_reference_conventions(tmp_path, weighted):
    import warnings
    from pecanpy_b200.graph import _parse_edge_list_native, _parse_edge_list_python
    rng = np.random.default_rng(5)
    path = tmp_path / "g.edg"
    weights = ["0.5", "1.5", "2", "-1.0", "0", "1e-3", "3.25E+2", " 7.125 ", ".5", "5.", "+4", "inf", "0.1234567890123456789"]
    with open(path, "w", newline="") as f:
        for i in range(2000):
            a, b = rng.integers(0, 60, size=2)
            ida = f" n{a} " if i % 7 == 0 else f"n{a}"                 # ids are stripped
            eol = "\r\n" if i % 5 == 0 else "\n"                       # strip() eats the CR
            if weighted:
                f.write(f"{ida}\tn{b}\t{weights[i % len(weights)]}{eol}")
            else:
                extra = "\tignored" if i % 11 == 0 else ""            # extra columns are ignored when unweighted
                f.write(f"{ida}\tn{b}{extra}{eol}")
            if i % 97 == 0:
                f.write("\n")                                          # blank lines are skipped
        f.write("last\tline" + ("\t1.0" if weighted else ""))          # no trailing newline
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nat = _parse_edge_list_native(str(path), weighted, "\t")
        py = _parse_edge_list_python(str(path), weighted, "\t")
    assert nat is not None, "libb2w.so must be built for this test"
    ref = _reference_read_lines(str(path), weighted)
    assert _same_parse(nat, ref) and _same_parse(py, ref)


def test_native_parser_other_delimiter_errors_and_unicode_fallback(tmp_path):
    import warnings
    from pecanpy_b200.graph import SparseGraph, _parse_edge_list_native, _parse_edge_list_python
    p = tmp_path / "c.edg"
    p.write_text("a, b, 1.5\nb,c,2\nc , a,0.25\n")
    nat = _parse_edge_list_native(str(p), True, ",")
    assert _same_parse(nat, _reference_read_lines(str(p), True, ","))
    assert nat[0] == ["a", "b", "c"]
    bad = tmp_path / "bad.edg"
    bad.write_text("a\tb\n")
    with pytest.raises(ValueError, match="three columns"):
        _parse_edge_list_native(str(bad), True, "\t")
    bad.write_text("a\tb\tx1\n")
    with pytest.raises(ValueError, match="float"):
        _parse_edge_list_native(str(bad), True, "\t")
    bad.write_text("lonely\n")
    with pytest.raises(ValueError):
        _parse_edge_list_native(str(bad), False, "\t")
    uni = tmp_path / "u.edg"
    uni.write_text("café\tnaïve\n x \tcafé\n", encoding="utf-8")
    assert _parse_edge_list_native(str(uni), False, "\t") is None      # refused: Python's strip() is Unicode aware
    g = SparseGraph()
    g.read_edg(str(uni), weighted=False, directed=False)                # falls back transparently
    assert g.nodes == ["café", "naïve", "x"]
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        drop = tmp_path / "d.edg"
        drop.write_text("a\tb\t-1\nb\tc\t2\n")
        names, src, dst, w = _parse_edge_list_native(str(drop), True, "\t")
    assert names == ["b", "c"] and len(rec) == 1 and "Non-positive" in str(rec[0].message)
