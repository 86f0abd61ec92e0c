"""Host-side containers and ingest against the UNMODIFIED reference: tests/golden/graph_*.npz hold the text of
edge-list files and what the reference's AdjlstGraph.read / to_csr / to_dense / read_edg / from_mat
(graph.py:160-362, 423-528, 587-657) made of them (written by oracle/gen_golden_graph.py, which imports the
reference in the build container)."""
import glob
import os
import warnings

import numpy as np
import pytest

from pecanpy_b200.graph import DenseGraph, SparseGraph

KARATE = "/root/reference/demo/karate.edg"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPH_FIXTURES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "graph_*.npz"))
                        if "literals" not in f)


def load_graph_fixture(name, tmp_path):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = {k: z[k] for k in z.files}
    path = tmp_path / (name + ".edg")
    with open(path, "w", newline="", encoding="utf-8") as f:
        f.write(str(c["text"]))
    c["path"] = str(path)
    c["weighted"], c["directed"], c["delimiter"] = bool(c["weighted"]), bool(c["directed"]), str(c["delimiter"])
    c["nodes"] = [str(x) for x in c["nodes"]]
    return c


def test_fixtures_exist():
    assert len(GRAPH_FIXTURES) >= 9


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_read_edg_equals_reference(tmp_path, name):
    """SparseGraph / DenseGraph.read_edg (native parser + host CSR build) == the reference's, byte for byte:
    duplicates and both orientations (later line wins), self loops, weights <= 0 dropped without registering their
    ids, ids stripped, CRLF, extra columns, no trailing newline, other delimiters."""
    c = load_graph_fixture(name, tmp_path)
    g = SparseGraph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.read_edg(c["path"], weighted=c["weighted"], directed=c["directed"], delimiter=c["delimiter"])
    assert g.nodes == c["nodes"]
    for k in ("indptr", "indices", "data"):
        got = getattr(g, k)
        assert got.dtype == c[k].dtype and np.array_equal(got, c[k]), k
    assert g.num_edges == int(c["num_edges"])
    d = DenseGraph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d.read_edg(c["path"], weighted=c["weighted"], directed=c["directed"], delimiter=c["delimiter"])
    assert d.nodes == c["nodes"]
    assert d.data.dtype == c["dense"].dtype and np.array_equal(d.data, c["dense"])
    assert np.array_equal(d.nonzero, c["nonzero"])


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_both_parsers_equal_reference(tmp_path, name, monkeypatch):
    """The native (C) parser and the Python fallback feed the same CSR build: both must reproduce the reference."""
    from pecanpy_b200 import graph as G
    c = load_graph_fixture(name, tmp_path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nat = G._parse_edge_list_native(c["path"], c["weighted"], c["delimiter"])
        py = G._parse_edge_list_python(c["path"], c["weighted"], c["delimiter"])
    assert nat is not None, "libb2w.so must be built for this test"
    assert nat[0] == py[0] == c["nodes"]
    for x, y in zip(nat[1:], py[1:]):
        assert np.array_equal(x, y)
    monkeypatch.setattr(G, "_parse_edge_list_native", lambda *a, **k: None)     # force the fallback through read_edg
    g = SparseGraph()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g.read_edg(c["path"], weighted=c["weighted"], directed=c["directed"], delimiter=c["delimiter"])
    assert g.nodes == c["nodes"]
    assert np.array_equal(g.indptr, c["indptr"]) and np.array_equal(g.indices, c["indices"]) and np.array_equal(g.data, c["data"])


def test_dropped_edge_warning_is_the_reference_message(tmp_path):
    """graph.py:184-192: 'Non-positive edge ignored: w(id1,id2) = weight' from either parser."""
    from pecanpy_b200 import graph as G
    p = tmp_path / "d.edg"
    p.write_text("a\tb\t-1\n x \tc\t0\nb\tc\t2\n")
    for parse in (G._parse_edge_list_native, G._parse_edge_list_python):
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            names, *_ = parse(str(p), True, "\t")
        assert names == ["b", "c"]
        assert [str(r.message) for r in rec] == ["Non-positive edge ignored: w(a,b) = -1.0",
                                                 "Non-positive edge ignored: w(x,c) = 0.0"]


def test_native_parser_leaves_ambiguous_inputs_to_python(tmp_path):
    """NUL bytes, bare CR line ends, '_' / '(' in a weight: the byte-level parser steps aside (B2W_ERR_UNSUPPORTED)
    and read_edg gives what Python's text-mode read + float() give."""
    from pecanpy_b200 import graph as G
    cases = {"nul.edg": ("a\x00b\tc\nc\td\n", False), "cr.edg": ("a\tb\rb\tc\r", False),
             "under.edg": ("a\tb\t1_0\n", True)}
    for fn, (text, weighted) in cases.items():
        p = tmp_path / fn
        with open(p, "w", newline="") as f:
            f.write(text)
        assert G._parse_edge_list_native(str(p), weighted, "\t") is None, fn
    g = SparseGraph()
    g.read_edg(str(tmp_path / "cr.edg"), weighted=False, directed=True)
    assert g.nodes == ["a", "b", "c"] and g.num_edges == 2
    g.read_edg(str(tmp_path / "under.edg"), weighted=True, directed=True)
    assert g.data.tolist() == [10.0]
    g.read_edg(str(tmp_path / "nul.edg"), weighted=False, directed=True)
    assert g.nodes == ["a\x00b", "c", "d"]
    bad = tmp_path / "nanp.edg"
    bad.write_text("a\tb\tnan(123)\n")
    with pytest.raises(ValueError):
        g.read_edg(str(bad), weighted=True, directed=True)
    # a line separator that only str.splitlines() honours must not split a line (the reference iterates the file)
    sep = tmp_path / "sep.edg"
    sep.write_text("a\x1cb\tc\n")
    g.read_edg(str(sep), weighted=False, directed=True)
    assert g.nodes == ["a\x1cb", "c"]
    assert G._parse_edge_list_python(str(sep), False, "\t")[0] == ["a\x1cb", "c"]


def test_from_mat_equals_reference_literals():
    """The matrices and CSR literals of the reference's test/test_graph.py:16-78 through from_mat."""
    z = np.load(os.path.join(GOLDEN, "graph_testgraph_literals.npz"))
    for i in (1, 2, 3):
        ids = [str(x) for x in z[f"ids{i}"]]
        sp = SparseGraph.from_mat(z[f"mat{i}"], ids)
        assert sp.nodes == ids
        for k in ("indptr", "indices", "data"):
            assert getattr(sp, k).dtype == z[f"{k}{i}"].dtype and np.array_equal(getattr(sp, k), z[f"{k}{i}"])
        assert sp.num_edges == int(z[f"num_edges{i}"]) and sp.density == float(z[f"density{i}"])
        dn = DenseGraph.from_mat(z[f"mat{i}"], ids)
        assert np.array_equal(dn.data, z[f"dense{i}"]) and dn.data.dtype == z[f"dense{i}"].dtype
        assert np.array_equal(dn.nonzero, z[f"nonzero{i}"])
        assert dn.num_edges == int(z[f"num_edges{i}"])


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
    # malformed lines: the native parser steps aside and the Python parser raises what the reference raises
    # (graph.py:166-178: IndexError from terms[1], ValueError for the column count with the line's repr, float()'s own)
    bad = tmp_path / "bad.edg"
    for text, weighted, exc, match in [("a\tb\n", True, ValueError, r"got 2 instead: 'a\\tb\\n'"),
                                       ("a\tb\tx1\n", True, ValueError, "could not convert string to float: 'x1'"),
                                       ("a\tb\nlonely\n", False, IndexError, "list index out of range"),
                                       ("a\tb\t1\t2", True, ValueError, r"got 4 instead: 'a\\tb\\t1\\t2'")]:
        bad.write_text(text)
        assert _parse_edge_list_native(str(bad), weighted, "\t") is None
        with pytest.raises(exc, match=match):
            SparseGraph().read_edg(str(bad), weighted=weighted, directed=False)
        with pytest.raises(exc, match=match):
            _parse_edge_list_python(str(bad), weighted, "\t")
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


def test_strategy_seams_of_the_reference_api():
    """get_has_nbrs() keeps the meaning of rw/sparse_rw.py:12-20 / rw/dense_rw.py:21-32 as a host callable;
    get_move_forward() (an njit per-step callback in the reference) does not exist on a GPU and says so."""
    from pecanpy_b200 import pecanpy as pp
    adj = np.array([[0, 1, 0, 0], [1, 0, 2, 0], [0, 2, 0, 0], [0, 0, 0, 0]], dtype=float)
    ids = ["a", "b", "c", "d"]
    for cls in (pp.SparseOTF, pp.PreComp, pp.DenseOTF):
        g = cls.from_mat(adj, ids, p=1, q=1)
        has = g.get_has_nbrs()
        assert [has(i) for i in range(4)] == [True, True, True, False]
        with pytest.raises(NotImplementedError):
            g.get_move_forward()


def test_adjlst_graph_behaviour(tmp_path):
    """AdjlstGraph / from_adjlst_graph (reference graph.py:108-387, 498-511, 631-643): first-appearance node order,
    overwrite and non-positive warnings, the insertion counter, sorted CSR rows, save -> read round trip.  (The full
    comparison with the reference's class runs in oracle/fuzz_reference_graph.py.)"""
    from pecanpy_b200.graph import AdjlstGraph, DenseGraph, SparseGraph
    g = AdjlstGraph()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        g.add_edge("b", "a", 2.0)
        g.add_edge("a", "c", 0.5, directed=True)
        g.add_edge("x", "y", -1.0)                          # ignored: neither node is created
        g.add_edge("b", "a", 3.0)                           # overwrites both directions
    assert [str(r.message) for r in rec] == ["Non-positive edge ignored: w(x,y) = -1.0",
                                             "edge from b to a exists, with value of 2.00. Now overwrite to 3.00."]
    assert g.nodes == ["b", "a", "c"] and g.num_edges == 5                    # insertions, repeats included
    assert g.edges == [(0, 1, 3.0), (1, 0, 3.0), (1, 2, 0.5)]
    indptr, indices, data = g.to_csr()
    assert indptr.dtype == np.uint32 and indptr.tolist() == [0, 1, 3, 3]
    assert indices.tolist() == [1, 0, 2] and data.dtype == np.float32 and data.tolist() == [3.0, 3.0, 0.5]
    assert np.array_equal(g.to_dense(), np.array([[0, 3, 0], [3, 0, 0.5], [0, 0, 0]]))
    sp = SparseGraph.from_adjlst_graph(g)
    dn = DenseGraph.from_adjlst_graph(g)
    assert sp.nodes == g.nodes and np.array_equal(sp.indptr, indptr) and np.array_equal(sp.data, data)
    assert dn.nodes == g.nodes and np.array_equal(dn.data, g.to_dense()) and dn.num_edges == 3
    p = tmp_path / "out.edg"
    g.save(str(p))
    assert p.read_text() == "b\ta\t3.0\na\tb\t3.0\na\tc\t0.5\n"
    h = AdjlstGraph()
    h.read(str(p), weighted=True, directed=True)
    assert h.nodes == g.nodes and h.edges == g.edges
    m = AdjlstGraph.from_mat(np.array([[0, -2.0], [1.5, 0]]), ["u", "v"])     # no validity check: the sign stays
    assert m.edges == [(0, 1, -2.0), (1, 0, 1.5)] and m.num_edges == 2
