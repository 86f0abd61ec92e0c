"""Differential check of the ingest row against the UNMODIFIED reference on random edge lists.

TEST INFRASTRUCTURE: build container only (imports /root/reference).  Random ``.edg`` texts -- ids with inner / outer
blanks, numeric-looking and non-ASCII ids, duplicate edges with other weights, both orientations, self loops, weights in
every spelling ``float()`` accepts (exponents, signs, ``.5``, ``5.``, ``inf``, ``1_0``), non-positive weights, CRLF or LF,
with or without a final newline, tab / comma / blank delimiters -- are read by the reference
(``SparseGraph.read_edg`` = ``AdjlstGraph.read`` + ``to_csr``, graph.py:270-341, 423-445) and by this repo's loaders:
the native parser (``b2w_edgelist_parse``), the Python parser, and ``read_edg`` itself.  Node list, ``indptr``,
``indices`` and ``data`` must be identical (bit for bit), and so must the dropped-edge warnings and -- one file in
twelve carries a malformed line -- the class and text of the exception.
``python oracle/fuzz_reference_graph.py --cases 200 --seed 0``; exit code 0 = all equal.
"""
import argparse
import json
import os
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_stubs"), "/root/reference/src", os.path.dirname(HERE)]

import numpy as np  # noqa: E402

from pecanpy import graph as ref_graph  # noqa: E402  (the reference)
from pecanpy_b200 import graph as our_graph  # noqa: E402

WEIGHTS = ["1", "2.5", "0.125", "1e-3", "3E2", "+4", ".5", "5.", " 7 ", "0", "-1", "-0.0", "0.0", "1e400", "inf",
           "1_0", "0x10" if False else "16", "1.0000001", "123456789.125", "4.9e-324"]
IDS = ["a", "b", "c", "n1", "n2", "007", "7", "x y", "gene-1", "P53", "é", "Ω3", "node_with_long_name_0123456789", "A", "a ",
       " lead", "t ", "z.z", "0", "-1", "1e3"]


def draw_text(rng):
    weighted = bool(rng.integers(2))
    directed = bool(rng.integers(2))
    delim = ["\t", "\t", ",", " "][int(rng.integers(4))]
    pool_size = int(rng.integers(2, 14))
    ascii_only = bool(rng.integers(3))                     # two thirds of the files stay on the native parser's turf
    ids = [i for i in IDS if (i.isascii() or not ascii_only) and (delim not in i.strip())]
    pool = [ids[int(k)] for k in rng.choice(len(ids), size=min(pool_size, len(ids)), replace=False)]
    pool += [str(int(x)) for x in rng.integers(0, 50, size=int(rng.integers(0, 8)))]
    if delim == " ":
        pool = [p.strip() for p in pool if " " not in p.strip() and " " not in p]
    pool = [p for p in pool if p.strip()] or ["a", "b"]
    n_lines = int(rng.integers(1, 40))
    eol = "\r\n" if rng.integers(4) == 0 else "\n"
    lines = []
    for _ in range(n_lines):
        a = pool[int(rng.integers(len(pool)))]
        b = pool[int(rng.integers(len(pool)))]
        cols = [a, b]
        if weighted:
            w = WEIGHTS[int(rng.integers(len(WEIGHTS)))]
            if delim == " ":
                w = w.strip()
            cols.append(w)
        elif rng.integers(8) == 0:
            cols.append("extra")                           # unweighted files may carry more columns (ignored)
        lines.append(delim.join(cols))
    if rng.integers(12) == 0:                              # one malformed line: the exception CLASS must agree as well
        a, b = pool[0], pool[-1]
        kind = int(rng.integers(4))
        if weighted:
            bad = [delim.join([a, b, "abc"]), delim.join([a, b]), delim.join([a, b, "1", "2"]), delim.join([a, b, "1e"])][kind]
        else:
            bad = [a, a + delim.strip(), a, a][kind] if delim.strip() else a
        lines.insert(int(rng.integers(len(lines) + 1)), bad)
    text = eol.join(lines) + (eol if rng.integers(3) else "")
    return text, weighted, directed, delim


def read_with(fn):
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        try:
            out = fn()
            err = None
        except Exception as exc:                           # noqa: BLE001 -- the exception class is part of the behaviour
            out, err = None, f"{type(exc).__name__}: {exc}"
    msgs = [str(r.message) for r in rec if "Non-positive edge ignored" in str(r.message)]
    return out, err, msgs


def csr_of(g):
    # (num_nodes / num_edges / density, graph.py:38-53, 416-421, ride along as a 3-vector)
    props = np.array([g.num_nodes, g.num_edges, g.density], dtype=np.float64)
    return list(g.nodes), np.asarray(g.indptr), np.asarray(g.indices), np.asarray(g.data), props


def same(a, b):
    return (a[0] == b[0] and all(x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x.view(np.uint8), y.view(np.uint8))
                                 for x, y in zip(a[1:], b[1:])))


def check(text, weighted, directed, delim, path):
    with open(path, "w", encoding="utf-8", newline="") as f:
        f.write(text)

    def ref():
        g = ref_graph.SparseGraph()
        g.read_edg(path, weighted, directed, delim)
        return csr_of(g)

    def ours(force_python):
        def run():
            g = our_graph.SparseGraph()
            if force_python:
                saved = our_graph._parse_edge_list_native
                our_graph._parse_edge_list_native = lambda *a, **k: None
                try:
                    g.read_edg(path, weighted, directed, delim)
                finally:
                    our_graph._parse_edge_list_native = saved
            else:
                g.read_edg(path, weighted, directed, delim)
            return csr_of(g)
        return run

    def dense_of(mod):
        def run():
            g = mod.DenseGraph()
            g.read_edg(path, weighted, directed, delim)
            return (list(g.nodes), np.asarray(g.data), np.asarray(g.nonzero),
                    np.array([g.num_nodes, g.num_edges, g.density], dtype=np.float64))
        return run

    def npz_and_mat(mod, src_mod):
        """save (written by src_mod) -> read_npz (by mod), weighted and not; from_mat of the dense matrix."""
        def run():
            g0 = src_mod.SparseGraph()
            g0.read_edg(path, weighted, directed, delim)
            npz = path + "." + src_mod.__name__.split(".")[0] + ".npz"
            g0.save(npz)
            out = []
            for wflag in (True, False):
                g = mod.SparseGraph()
                g.read_npz(npz, wflag)
                out.append(([str(x) for x in g.nodes], np.asarray(g.indptr), np.asarray(g.indices), np.asarray(g.data)))
            d = src_mod.DenseGraph()
            d.read_edg(path, weighted, directed, delim)
            g = mod.SparseGraph.from_mat(np.asarray(d.data), list(d.nodes))
            out.append((list(g.nodes), np.asarray(g.indptr), np.asarray(g.indices), np.asarray(g.data)))
            return out
        return run

    def adjlst(mod):
        """AdjlstGraph (graph.py:108-387): read, the stored edges, the insertion counter, both conversions, save,
        from_mat on the dense matrix (negative entries kept), from_adjlst_graph of both graph classes."""
        def run():
            g = mod.AdjlstGraph()
            g.read(path, weighted, directed, delim)
            out_path = path + "." + mod.__name__.split(".")[0] + ".saved"
            mod.AdjlstGraph.save(g, out_path, not weighted, delim)
            with open(out_path, encoding="utf-8") as f:
                saved = f.read()
            dense = g.to_dense()
            signed = dense.copy()
            signed[::2] *= -1.0
            g2 = mod.AdjlstGraph.from_mat(signed, list(g.nodes))
            sp = mod.SparseGraph.from_adjlst_graph(g)
            dn = mod.DenseGraph.from_adjlst_graph(g)
            return (list(g.nodes), [tuple(map(float, e)) for e in g.edges], int(g.num_edges), saved, g.to_csr(), dense,
                    [tuple(map(float, e)) for e in g2.edges], int(g2.num_edges), g2.to_csr(),
                    (list(sp.nodes), sp.indptr, sp.indices, sp.data), (list(dn.nodes), np.asarray(dn.data), np.asarray(dn.nonzero)))
        return run

    def deep_same(a, b):
        if isinstance(a, (tuple, list)) and isinstance(b, (tuple, list)):
            return len(a) == len(b) and all(deep_same(x, y) for x, y in zip(a, b))
        if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
            a, b = np.asarray(a), np.asarray(b)
            return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
        return a == b and type(a) is type(b)

    def all_warnings(fn):
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            try:
                out, err = fn(), None
            except Exception as exc:                       # noqa: BLE001
                out, err = None, f"{type(exc).__name__}: {exc}"
        return out, err, [str(r.message) for r in rec]

    want, werr, wmsg = read_with(ref)
    bad = []
    ra, raerr, rawarn = all_warnings(adjlst(ref_graph))
    oa, oaerr, oawarn = all_warnings(adjlst(our_graph))
    if raerr != oaerr or rawarn != oawarn or (ra is not None and not deep_same(ra, oa)):
        bad.append(f"AdjlstGraph differs (reference: {raerr}, this repo: {oaerr}; warnings equal: {rawarn == oawarn})")
    if not werr and want[1].size > 1:                      # .npz round trips (graph.py:447-497) and from_mat (:513-528)
        rr, _, _ = read_with(npz_and_mat(ref_graph, ref_graph))
        for label, m, src in (("read_npz of a reference file / from_mat", our_graph, ref_graph),
                              ("reference read_npz of this repo's file", ref_graph, our_graph)):
            oo, oerr, _ = read_with(npz_and_mat(m, src))
            if oerr or rr is None or any(not same(a, b) for a, b in zip(oo, rr)):
                bad.append(f"{label} differs ({oerr})")
    if not werr:                                           # DenseGraph.read_edg (graph.py:613-625: read + to_dense)
        dw, _, _ = read_with(dense_of(ref_graph))
        dg, derr, _ = read_with(dense_of(our_graph))
        if derr or not same(dg, dw):
            bad.append(f"DenseGraph.read_edg differs ({derr})")
    for label, fn in (("read_edg", ours(False)), ("python parser", ours(True))):
        got, gerr, gmsg = read_with(fn)
        if werr or gerr:
            if werr != gerr:
                bad.append(f"{label}: reference raised {werr}, this repo {gerr}")
            continue
        if not same(got, want):
            bad.append(f"{label}: CSR / node list differs")
        if gmsg[:20] != wmsg[:20]:
            bad.append(f"{label}: dropped-edge warnings differ")
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    failures = []
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "g.edg")
        for k in range(args.cases):
            text, weighted, directed, delim = draw_text(rng)
            bad = check(text, weighted, directed, delim, path)
            if bad or args.verbose:
                print(("MISMATCH " if bad else "ok ") + f"case {k} weighted={weighted} directed={directed} delim={delim!r}: "
                      + "; ".join(bad) + ("\n" + repr(text) if bad else ""), flush=True)
            if bad:
                failures.append(k)
    print(json.dumps({"cases": args.cases, "seed": args.seed, "failures": failures}))
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
