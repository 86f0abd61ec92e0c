"""Graph containers: the input layout of the walk kernels.

The reference's graph I/O (src/pecanpy/graph.py) is out of scope for the B200 engine and is not
re-implemented feature for feature; these light containers exist so that the drop-in classes
in :mod:`pecanpy_b200.pecanpy` can be constructed and loaded exactly like the reference's
(``read_edg`` / ``read_npz`` / ``from_mat`` / ``from_adjlst_graph`` / ``save``; same attribute names and dtypes):

* ``AdjlstGraph``: the editable adjacency-list graph of reference graph.py:108-387 (``add_node`` / ``add_edge`` /
  ``read`` / ``save`` / ``to_csr`` / ``to_dense`` / ``from_mat``)

* ``SparseGraph``: CSR ``indptr`` uint32[n+1], ``indices`` uint32[nnz] (rows sorted, unique),
  ``data`` float32[nnz]   (reference graph.py:389-528, typing.py:31)
* ``DenseGraph``:  ``data`` float64[n,n], ``nonzero`` bool[n,n] == (data != 0)  (graph.py:531-657)

Node order follows the reference: first appearance in the edge list (graph.py:217-236).
"""
from __future__ import annotations

import os
import warnings
from typing import Dict, List, Optional, Sequence

import numpy as np


class BaseGraph:
    """Node-id bookkeeping (mirrors reference graph.py:19-105)."""

    def __init__(self):
        self._node_ids: List[str] = []
        self._node_idmap: Dict[str, int] = {}

    @property
    def nodes(self) -> List[str]:
        return self._node_ids

    @property
    def num_nodes(self) -> int:
        return len(self._node_ids)

    @property
    def num_edges(self) -> int:
        raise NotImplementedError

    @property
    def density(self) -> float:
        return self.num_edges / self.num_nodes / (self.num_nodes - 1)

    def get_has_nbrs(self):
        raise NotImplementedError                # abstract, as graph.py:99-101

    def get_move_forward(self):
        raise NotImplementedError                # abstract, as graph.py:103-105

    def set_node_ids(self, node_ids: Optional[Sequence[str]], implicit_ids: bool = False,
                     num_nodes: Optional[int] = None):
        if node_ids is not None and not implicit_ids:
            self._node_ids = list(node_ids)
        elif num_nodes is None:
            raise ValueError("Need to specify `num_nodes` when setting implicit node IDs.")
        else:
            self._node_ids = [str(i) for i in range(num_nodes)]
            if not implicit_ids:
                warnings.warn("WARNING: Implicitly set node IDs to the canonical node ordering due to missing "
                              "IDs field in the raw CSR npz file.", stacklevel=2)
        self._node_idmap = {j: i for i, j in enumerate(self._node_ids)}


def _parse_edge_list(path: str, weighted: bool, delimiter: str = "\t"):
    """Parse an ``.edg`` file into (ids, src, dst, weights) in FILE ORDER with the reference's conventions
    (graph.py:160-305): node order = first appearance (id1 then id2, line by line), non-positive weights dropped
    with a warning.  Uses the native parser of libb2w (``b2w_edgelist_parse``, ~50x faster) when the library is built
    and the file is plain ASCII; otherwise the NumPy-vectorised Python below, which gives the same arrays."""
    native = _parse_edge_list_native(path, weighted, delimiter)
    if native is not None:
        return native
    return _parse_edge_list_python(path, weighted, delimiter)


def _parse_edge_list_native(path: str, weighted: bool, delimiter: str):
    import ctypes as C
    try:
        from . import _capi as capi
        lib = capi.lib()
    except (ImportError, OSError):
        return None
    with open(path, "rb"):                # same exceptions as the reference's open() for a missing / unreadable file
        pass
    h = C.c_void_p(None)
    m, n, nb, nd = C.c_uint64(0), C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
    rc = lib.b2w_edgelist_parse(os.fsencode(path), int(bool(weighted)), delimiter.encode("utf-8"), C.byref(h), C.byref(m),
                                C.byref(n), C.byref(nb), C.byref(nd))
    if rc == capi.ERR_UNSUPPORTED:
        return None                       # non-ASCII file: Python's Unicode-aware strip() decides
    if rc == capi.ERR_NOMEM:
        raise MemoryError(lib.b2w_last_error().decode("utf-8", "replace"))
    if rc != capi.OK:
        return None                       # a malformed line: the Python parser raises the reference's own exception
    try:
        src = np.empty(m.value, dtype=np.uint32)
        dst = np.empty(m.value, dtype=np.uint32)
        w = np.empty(m.value, dtype=np.float64)
        blob = C.create_string_buffer(max(int(nb.value), 1))
        lines = (C.c_uint64 * 20)()
        nl = C.c_uint32(0)
        capi.check(lib.b2w_edgelist_fetch(h, C.c_void_p(src.ctypes.data), C.c_void_p(dst.ctypes.data),
                                          C.c_void_p(w.ctypes.data), C.cast(blob, C.c_void_p), C.cast(lines, C.c_void_p),
                                          C.byref(nl)), "b2w_edgelist_fetch")
    finally:
        lib.b2w_edgelist_free(h)
    if nl.value:
        # the same message as the reference / the Python parser (graph.py:187-192): edge and weight of the line
        want = {int(lines[k]) for k in range(nl.value)}
        with open(path, encoding="utf-8", newline="\n") as f:
            for no, line in enumerate(f, start=1):
                if no in want:
                    t = line.strip().split(delimiter)
                    warnings.warn(f"Non-positive edge ignored: w({t[0].strip()},{t[1].strip()}) = {float(t[-1])}",
                                  RuntimeWarning, stacklevel=3)
    names = blob.raw[:max(int(nb.value) - 1, 0)].decode("ascii").split("\0") if n.value else []
    return names, src.astype(np.int64), dst.astype(np.int64), w


def _parse_edge_list_python(path: str, weighted: bool, delimiter: str = "\t"):
    """The same parse in NumPy-vectorised Python (Unicode aware; used when the native parser is unavailable)."""
    with open(path, encoding="utf-8") as f:                 # universal newlines, as the reference's `for line in f`
        pieces = f.read().split("\n")                       # (str.splitlines would also break at \x0b, \x1c, \x85, ...)
    # the lines as the reference's loop sees them (each with its "\n", except an unterminated last one);
    # blank lines are skipped (the reference raises IndexError on them, graph.py:166-167: a documented tolerance)
    lines = [ln + "\n" for ln in pieces[:-1]] + ([pieces[-1]] if pieces[-1] else [])
    lines = [ln for ln in lines if ln.strip()]
    if not lines:
        return [], np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64)
    parts = [ln.strip().split(delimiter) for ln in lines]
    wl = []
    for ln, t in zip(lines, parts):                         # errors in file order, of the reference's classes and texts
        if len(t) < 2:
            raise IndexError("list index out of range")     # terms[1] (graph.py:168)
        if weighted:
            if len(t) != 3:
                raise ValueError(f"Expecting three columns in the edge list file for a weighted graph, "
                                 f"got {len(t)} instead: {ln!r}")
            wl.append(float(t[-1]))
    w = np.array(wl, dtype=np.float64) if weighted else np.ones(len(parts), dtype=np.float64)
    a = np.array([t[0].strip() for t in parts], dtype=object)
    b = np.array([t[1].strip() for t in parts], dtype=object)
    bad = w <= 0
    if bad.any():
        for i in np.flatnonzero(bad)[:20]:
            warnings.warn(f"Non-positive edge ignored: w({a[i]},{b[i]}) = {w[i]}", RuntimeWarning, stacklevel=2)
        a, b, w = a[~bad], b[~bad], w[~bad]
    # first-appearance order over the interleaved sequence a0, b0, a1, b1, ...
    inter = np.empty(2 * a.size, dtype=object)
    inter[0::2], inter[1::2] = a, b
    uniq, first, inv = np.unique(inter, return_index=True, return_inverse=True)
    rank = np.empty(uniq.size, dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(uniq.size)
    names = [None] * uniq.size
    for u, r in zip(uniq, rank):
        names[r] = u
    return names, rank[inv[0::2]], rank[inv[1::2]], w


def _read_edge_list(path: str, weighted: bool, directed: bool, delimiter: str = "\t"):
    """(ids, rows, cols, weights) with duplicates resolved on the host: a later definition of the same edge
    overwrites an earlier one (graph.py:273-305)."""
    names, ia, ib, w = _parse_edge_list(path, weighted, delimiter)
    seq = np.arange(ia.size, dtype=np.int64)
    if directed:
        rows, cols, ww, order = ia, ib, w, seq
    else:
        rows = np.concatenate([ia, ib]); cols = np.concatenate([ib, ia])
        ww = np.concatenate([w, w]); order = np.concatenate([seq, seq])
    # later definitions win: keep the last occurrence of every (row, col)
    n = max(len(names), 1)
    key = rows * n + cols
    srt = np.lexsort((order, key))
    key_s = key[srt]
    last = np.ones(key_s.size, dtype=bool)
    last[:-1] = key_s[1:] != key_s[:-1]
    keep = srt[last]
    return names, rows[keep], cols[keep], ww[keep]


def _coo_to_csr(n: int, rows, cols, w):
    order = np.lexsort((cols, rows))
    rows, cols, w = rows[order], cols[order], w[order]
    indptr = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    return indptr, cols.astype(np.uint32), w.astype(np.float32)


class AdjlstGraph(BaseGraph):
    """Editable adjacency-list graph with the observable behaviour of the reference's (graph.py:108-387): the class
    users build or read a graph with before handing it to ``SparseGraph.from_adjlst_graph`` /
    ``DenseGraph.from_adjlst_graph``.  Per node one dict ``neighbour index -> weight``; node order = first appearance;
    a repeated edge overwrites (with the reference's warning when the weight changes); non-positive weights are
    ignored with a warning; ``num_edges`` counts INSERTIONS (both directions of an undirected edge, repeats included),
    as the reference's counter does (graph.py:238-241)."""

    def __init__(self):
        super().__init__()
        self._data: List[Dict[int, float]] = []
        self._num_edges = 0

    @property
    def edges_iter(self):
        for head, nbrs in enumerate(self._data):
            for tail in sorted(nbrs):
                yield head, tail, nbrs[tail]

    @property
    def edges(self):
        return list(self.edges_iter)

    @property
    def num_edges(self):
        return self._num_edges

    def add_node(self, node_id: str):
        if node_id not in self._node_idmap:
            self._node_idmap[node_id] = len(self._node_ids)
            self._node_ids.append(node_id)
            self._data.append({})

    def get_node_idx(self, node_id: str) -> int:
        self.add_node(node_id)
        return self._node_idmap[node_id]

    def _add_edge_from_idx(self, idx1: int, idx2: int, weight: float):
        self._data[idx1][idx2] = weight
        self._num_edges += 1

    def add_edge(self, id1: str, id2: str, weight: float = 1.0, directed: bool = False):
        if weight <= 0:                                      # graph.py:182-192
            warnings.warn(f"Non-positive edge ignored: w({id1},{id2}) = {weight}", RuntimeWarning, stacklevel=2)
            return
        i, j = self.get_node_idx(id1), self.get_node_idx(id2)
        old = self._data[i].get(j)
        if old is not None and old != weight:                # graph.py:194-215 (checked for the forward direction only)
            warnings.warn(f"edge from {id1} to {id2} exists, with value of {old:.2f}. Now overwrite to {weight:.2f}.",
                          RuntimeWarning, stacklevel=2)
        self._add_edge_from_idx(i, j, weight)
        if not directed:
            self._add_edge_from_idx(j, i, weight)

    def read(self, path: str, weighted: bool, directed: bool, delimiter: str = "\t"):
        """One ``add_edge`` per line, in file order (graph.py:160-180, 270-305): ``id1 <delim> id2 [<delim> weight]``."""
        with open(path, encoding="utf-8") as f:
            for line in f:
                terms = line.strip().split(delimiter)
                id1, id2 = terms[0].strip(), terms[1].strip()
                weight = 1.0
                if weighted:
                    if len(terms) != 3:
                        raise ValueError(f"Expecting three columns in the edge list file for a weighted graph, "
                                         f"got {len(terms)} instead: {line!r}")
                    weight = float(terms[-1])
                self.add_edge(id1, id2, weight, directed)

    def save(self, path: str, unweighted: bool = False, delimiter: str = "\t"):
        """Edge list, one line per stored (head, tail) in index order (graph.py:307-321)."""
        with open(path, "w", encoding="utf-8") as f:
            for h, t, w in self.edges_iter:
                cols = (self._node_ids[h], self._node_ids[t]) if unweighted else (self._node_ids[h], self._node_ids[t], str(w))
                f.write(delimiter.join(cols) + "\n")

    def to_csr(self):
        """(indptr uint32, indices uint32, data float32), rows sorted (graph.py:323-341)."""
        deg = np.fromiter((len(r) for r in self._data), dtype=np.int64, count=len(self._data))
        indptr = np.zeros(len(self._data) + 1, dtype=np.uint32)
        np.cumsum(deg, out=indptr[1:])
        indices = np.zeros(int(indptr[-1]), dtype=np.uint32)
        data = np.zeros(int(indptr[-1]), dtype=np.float32)
        for i, nbrs in enumerate(self._data):
            if nbrs:
                keys = sorted(nbrs)
                lo = int(indptr[i])
                indices[lo:lo + len(keys)] = keys
                data[lo:lo + len(keys)] = [nbrs[k] for k in keys]
        return indptr, indices, data

    def to_dense(self):
        """Full float64 adjacency matrix in ``nodes`` order (graph.py:343-362)."""
        n = len(self._node_ids)
        mat = np.zeros((n, n))
        for i, nbrs in enumerate(self._data):
            if nbrs:
                mat[i, list(nbrs)] = list(nbrs.values())
        return mat

    @classmethod
    def from_mat(cls, adj_mat, node_ids: List[str], **kwargs):
        """Every NONZERO entry becomes an edge, sign included, without the checks of ``add_edge`` (graph.py:364-387)."""
        g = cls(**kwargs)
        for node_id in node_ids:
            g.add_node(node_id)
        adj = np.asarray(adj_mat)
        for i, j in zip(*np.nonzero(adj != 0)):
            g._add_edge_from_idx(i, j, adj[i, j])
        return g


class SparseGraph(BaseGraph):
    def __init__(self):
        super().__init__()
        self.data: Optional[np.ndarray] = None
        self.indptr: Optional[np.ndarray] = None
        self.indices: Optional[np.ndarray] = None

    @property
    def num_edges(self) -> int:
        if self.indptr is None:
            raise ValueError("Empty graph.")
        return self.indptr[-1]                 # (a NumPy scalar, as in the reference, graph.py:416-421)

    def read_edg(self, path: str, weighted: bool, directed: bool, delimiter: str = "\t", device=None):
        """Load an edge list (reference graph.py:447-486).  With ``device`` (e.g. ``"cuda:0"``) the CSR is built on
        that GPU (``b2w_csr_from_edges``: one stable radix sort instead of the host lexsort); the arrays are
        identical either way."""
        if device is not None:
            from .ingest import csr_from_edges_device
            names, ia, ib, w = _parse_edge_list(path, weighted, delimiter)
            self.set_node_ids(names)
            self.indptr, self.indices, self.data = csr_from_edges_device(len(names), ia, ib, w if weighted else None,
                                                                         directed, device=device)
            return
        names, rows, cols, w = _read_edge_list(path, weighted, directed, delimiter)
        self.set_node_ids(names)
        self.indptr, self.indices, self.data = _coo_to_csr(len(names), rows, cols, w)

    def read_npz(self, path: str, weighted: bool, implicit_ids: bool = False):
        raw = np.load(path)
        self.indptr = raw["indptr"].astype(np.uint32)
        self.indices = raw["indices"].astype(np.uint32)
        self.data = raw["data"].astype(np.float32)
        if not weighted:
            self.data[:] = 1.0
        ids = raw["IDs"] if "IDs" in raw.files else None
        self.set_node_ids(ids, implicit_ids=implicit_ids, num_nodes=int(self.indptr.size - 1))

    def save(self, path: str):
        np.savez(path, IDs=self.nodes, data=self.data, indptr=self.indptr, indices=self.indices)

    @classmethod
    def from_mat(cls, adj_mat, node_ids: List[str], **kwargs):
        g = cls(**kwargs)
        g.set_node_ids(node_ids)
        adj = np.asarray(adj_mat)
        rows, cols = np.nonzero(adj)
        g.indptr, g.indices, g.data = _coo_to_csr(adj.shape[0], rows, cols, adj[rows, cols].astype(np.float64))
        return g

    @classmethod
    def from_adjlst_graph(cls, adjlst_graph, **kwargs):
        """graph.py:498-511: node ids and CSR of an ``AdjlstGraph`` (this module's or the reference's)."""
        g = cls(**kwargs)
        g.set_node_ids(adjlst_graph.nodes)
        g.indptr, g.indices, g.data = adjlst_graph.to_csr()
        return g

    @classmethod
    def from_csr(cls, indptr, indices, data, node_ids: Optional[List[str]] = None, **kwargs):
        """Adopt ready CSR arrays (rows must be sorted and duplicate-free)."""
        g = cls(**kwargs)
        g.indptr = np.ascontiguousarray(indptr, dtype=np.uint32)
        g.indices = np.ascontiguousarray(indices, dtype=np.uint32)
        g.data = np.ascontiguousarray(data, dtype=np.float32)
        g.set_node_ids(node_ids, implicit_ids=node_ids is None, num_nodes=int(g.indptr.size - 1))
        return g


class DenseGraph(BaseGraph):
    def __init__(self):
        super().__init__()
        self._data: Optional[np.ndarray] = None
        self._nonzero: Optional[np.ndarray] = None

    @property
    def num_edges(self) -> int:
        if self._nonzero is None:
            raise ValueError("Empty graph.")
        return self._nonzero.sum()             # (a NumPy scalar, as in the reference, graph.py:563-569)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, data):
        self._data = np.ascontiguousarray(np.asarray(data).astype(float))
        self._nonzero = np.array(self._data != 0, dtype=bool)

    @property
    def nonzero(self):
        return self._nonzero

    def read_npz(self, path: str, weighted: bool, implicit_ids: bool = False):
        raw = np.load(path)
        self.data = raw["data"]
        if not weighted:
            self.data = self.nonzero * 1.0
        ids = raw["IDs"] if "IDs" in raw.files else None
        self.set_node_ids(ids, implicit_ids=implicit_ids, num_nodes=self.data.shape[0])

    def read_edg(self, path: str, weighted: bool, directed: bool, delimiter: str = "\t"):
        names, rows, cols, w = _read_edge_list(path, weighted, directed, delimiter)
        n = len(names)
        mat = np.zeros((n, n))
        mat[rows, cols] = w
        self.set_node_ids(names)
        self.data = mat

    def save(self, path: str):
        np.savez(path, data=self.data, IDs=self.nodes)

    @classmethod
    def from_adjlst_graph(cls, adjlst_graph, **kwargs):
        """graph.py:631-643: node ids and dense matrix of an ``AdjlstGraph`` (this module's or the reference's)."""
        g = cls(**kwargs)
        g.set_node_ids(adjlst_graph.nodes)
        g.data = adjlst_graph.to_dense()
        return g

    @classmethod
    def from_mat(cls, adj_mat, node_ids: List[str], **kwargs):
        g = cls(**kwargs)
        g.data = adj_mat
        g.set_node_ids(node_ids)
        return g
