"""Drop-in walk generators with the API surface of ``pecanpy.pecanpy`` (reference
src/pecanpy/pecanpy.py), backed by the B200 CUDA engine.

Same class names, constructor signature ``(p, q, workers, verbose, extend, gamma, random_state)``
(pecanpy.py:83-92), loaders (``read_edg`` / ``read_npz`` / ``from_mat``), and the methods
``preprocess_transition_probs()``, ``simulate_walks(num_walks, walk_length) -> List[List[str]]``
(pecanpy.py:116-162) and ``embed(...)`` (pecanpy.py:240-290).  ``simulate_walks_array`` returns
the raw ``uint32[tot, L+2]`` matrix in the layout of ``Base._random_walks`` (pecanpy.py:182-206).

What changes with respect to the reference:

* the njit closure pair ``get_move_forward()`` / ``get_has_nbrs()`` cannot be called from a GPU;
  the strategy is selected by the class (``_MODE``) instead.  ``_random_walks`` keeps the reference's positional
  signature and ignores the two callbacks; ``get_has_nbrs()`` returns a plain host callable, ``get_move_forward()``
  raises.  The njit helpers those closures call (``get_normalized_probs``, ``get_extended_normalized_probs``,
  ``get_normalized_probs_first_order``, rw/sparse_rw.py:39-130, rw/dense_rw.py:34-118) exist only inside the CUDA
  kernels and are not exposed as host methods: a host copy of them would be a second, CPU implementation of the path;
* random numbers: Philox4x32-10 keyed by ``(random_state, global walker row, step)`` instead of a
  per-thread MT19937, so seeded walks are reproducible for ANY thread / GPU count (the reference
  is reproducible only at one thread, pecanpy.py:51-55).  The start-node shuffle stays on the host
  and is the reference's, verbatim (pecanpy.py:135-141);
* there is no CPU fallback: without a CUDA device or the built extension the methods raise.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .graph import BaseGraph, DenseGraph, SparseGraph


def shuffled_start(num_nodes: int, num_walks: int, random_state) -> np.ndarray:
    """The start array of pecanpy.py:135-141 -- ``num_walks`` copies of every node index, shuffled by NumPy's legacy
    GLOBAL generator seeded with ``random_state`` -- bit for bit, including the state the global generator is left in.
    NumPy seeds (so every seed type, ``None`` included, means what it means in the reference); the Fisher-Yates loop
    itself runs natively (``b2w_shuffled_start``: 3-6x faster at 10^7 walkers, where NumPy's shuffle would cost several
    times the GPU's whole walk).  Plain NumPy when the library is not built or the array has 2^32 entries or more."""
    np.random.seed(random_state)
    tot = int(num_nodes) * int(num_walks)
    lib = None
    if 2 <= tot < 2 ** 32:
        try:
            from . import _capi as capi
            lib = capi.lib()
        except (ImportError, OSError):
            lib = None
    if lib is None:
        nodes = np.arange(num_nodes, dtype=np.uint32)                 # == np.array(range(n), dtype=np.uint32)
        start = np.concatenate([nodes] * num_walks) if num_walks else np.zeros(0, np.uint32)
        np.random.shuffle(start)
        return start
    import ctypes as C
    name, key, pos, has_gauss, cached = np.random.get_state()
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    cpos = C.c_int32(int(pos))
    start = np.empty(tot, dtype=np.uint32)
    capi.check(lib.b2w_shuffled_start(int(num_nodes), int(num_walks), C.c_void_p(key.ctypes.data), C.byref(cpos),
                                      C.c_void_p(start.ctypes.data)), "b2w_shuffled_start")
    np.random.set_state((name, key, int(cpos.value), has_gauss, cached))
    return start


class Base(BaseGraph):
    _MODE = ""

    def __init__(self, p: float = 1, q: float = 1, workers: int = 1, verbose: bool = False, extend: bool = False,
                 gamma: float = 0, random_state: Optional[int] = None):
        super().__init__()
        self.p = p
        self.q = q
        self.workers = workers
        self.verbose = verbose
        self.extend = extend
        self.gamma = gamma
        self.random_state = random_state
        self._preprocessed = False
        self._preprocess_key = None
        self._engine = None
        self._engine_key = None
        self._replicas = {}
        self._alias_dev = None
        self.device = None          # torch device string/obj; None = current CUDA device
        self.devices = None         # list of CUDA devices for single-process multi-GPU walks (None = one GPU)
        self.last_seed = None

    # -- engine ------------------------------------------------------------------------------
    def _make_engine(self):
        raise NotImplementedError

    def _graph_key(self):
        """Identity of what the device copy was made from.  The reference rebuilds its closures from the current
        arrays on every simulate_walks call (pecanpy.py:143-144); here the device copy is cached and must be dropped
        when the graph is reloaded (read_edg / read_npz / from_mat), an array is re-assigned, or gamma changes."""
        raise NotImplementedError

    def _engine_stale(self) -> bool:
        return self._engine is not None and self._engine_key != (self._graph_key(), self.extend, self.gamma)

    @property
    def engine(self):
        if self._engine_stale():
            self.release()
        if self._engine is None:
            self._engine = self._make_engine()
            self._engine_key = (self._graph_key(), self.extend, self.gamma)
            if self.extend and self._MODE in ("SparseOTF", "DenseOTF", "PreComp"):
                self._engine.compute_thresholds(self.gamma)
        return self._engine

    def get_noise_thresholds(self) -> np.ndarray:
        """``max(mean + gamma * std, 0)`` of every row's weights (rw/sparse_rw.py:22-35, rw/dense_rw.py:11-19),
        computed by the device kernel (bit-identical to the reference's NumPy loop) and returned as float32[n]."""
        eng = self.engine
        thr = eng.thr if (self.extend and eng.thr is not None) else eng.compute_thresholds(self.gamma)
        return thr.cpu().numpy()

    def release(self):
        """Drop the device copy of the graph (and everything derived from it: thresholds, alias tables, edge index)."""
        if self._engine is not None:
            self._engine.close()
        for e in getattr(self, "_replicas", {}).values():
            e.close()
        self._replicas = {}
        self._engine = None
        self._engine_key = None
        self._preprocessed = False
        self._preprocess_key = None
        self._alias_dev = None
        if hasattr(self, "_alias_host"):
            self._alias_host = None

    # -- reference surface ---------------------------------------------------------------------
    def preprocess_transition_probs(self):
        """Null default (pecanpy.py:231-233)."""

    def _preprocess_transition_probs(self):
        key = (self._graph_key(), self.extend, self.gamma, float(self.p), float(self.q))
        if self._engine_stale() or not self._preprocessed or self._preprocess_key != key:
            _ = self.engine                      # (re)creates the device copy if the graph changed
            self.preprocess_transition_probs()
            self._preprocessed = True
            self._preprocess_key = key

    def _start_nodes(self, num_walks: int) -> np.ndarray:
        return shuffled_start(self.num_nodes, num_walks, self.random_state)

    def _seed(self) -> int:
        from .engine import new_seed
        self.last_seed = new_seed() if self.random_state is None else int(self.random_state)
        return self.last_seed

    def simulate_walks_array(self, num_walks: int, walk_length: int) -> np.ndarray:
        """Raw walk matrix ``uint32[num_nodes * num_walks, walk_length + 2]`` (host), layout of pecanpy.py:182-206.
        With ``self.devices`` set to several CUDA devices the rows are sharded over them from THIS process (one
        host thread per GPU, graph replicated, Philox keyed by the global row: the matrix is identical for any
        number of GPUs) -- simulate_walks stays a plain call, no torchrun."""
        self._preprocess_transition_probs()
        start = self._start_nodes(num_walks)
        seed = self._seed()
        if self.devices and len(self.devices) > 1:
            from .multi import walk_host_multi
            return walk_host_multi(self, start, walk_length, seed)
        return self.engine.walk_host(self._MODE, self.p, self.q, start, walk_length, seed, extend=bool(self.extend))

    def _random_walks(self, tot_num_jobs: int, walk_length: int, random_state: Optional[int], start_node_idx_ary,
                      has_nbrs=None, move_forward=None, progress_proxy=None) -> np.ndarray:
        """The reference's kernel entry (pecanpy.py:164-210; a staticmethod there, called ``self._random_walks(...)``)
        with its positional arguments and its return value: ``uint32[tot_num_jobs, walk_length + 2]``, row i = the
        walk from ``start_node_idx_ary[i]``.  ``has_nbrs`` / ``move_forward`` (the njit strategy callbacks) are
        accepted and ignored: the strategy is this class's mode, evaluated by the CUDA kernels.
        ``random_state=None`` draws a fresh seed, like the reference's unseeded run."""
        from .engine import new_seed
        self._preprocess_transition_probs()
        start = np.ascontiguousarray(np.asarray(start_node_idx_ary)[:tot_num_jobs], dtype=np.uint32)
        if start.size != tot_num_jobs:
            raise ValueError("start_node_idx_ary holds fewer than tot_num_jobs entries")
        self.last_seed = new_seed() if random_state is None else int(random_state)
        return self.engine.walk_host(self._MODE, self.p, self.q, start, walk_length, self.last_seed,
                                     extend=bool(self.extend))

    def get_move_forward(self):
        """The reference returns an njit closure here (graph.py:103-105, pecanpy.py:384-440, 522-561, 576-614) that
        ``_random_walks`` calls once per step.  A GPU kernel cannot call back into the host per step: the strategy
        is the class (mode enum of include/b2w.h) and this seam does not exist."""
        raise NotImplementedError("pecanpy_b200 evaluates the walk strategy inside its CUDA kernels; there is no per-step "
                                  "move_forward callback -- call simulate_walks / simulate_walks_array / _random_walks")

    def _map_walk(self, walk_idx_ary) -> List[str]:
        end_idx = walk_idx_ary[-1]
        return [self.nodes[i] for i in walk_idx_ary[:end_idx]]

    def simulate_walks(self, num_walks: int, walk_length: int) -> List[List[str]]:
        """Same return value as the reference (pecanpy.py:116-162): one list of node ids per walk, truncated at
        dead ends.  The id mapping of pecanpy.py:160 is one C loop over the matrix (walks.map_walks)."""
        from .walks import map_walks
        return map_walks(self.simulate_walks_array(num_walks, walk_length), self.nodes)

    def simulate_walks_corpus(self, num_walks: int, walk_length: int, block: int = 8192):
        """The same walks as a lazy, restartable iterable (walks.WalkCorpus): what a streaming consumer such as
        gensim's Word2Vec needs, without 10^7 Python lists alive at once."""
        from .walks import WalkCorpus
        return WalkCorpus(self.simulate_walks_array(num_walks, walk_length), self.nodes, block=block)

    def embed(self, dim: int = 128, num_walks: int = 10, walk_length: int = 80, window_size: int = 10,
              epochs: int = 1, verbose: bool = False):
        try:
            from gensim.models import Word2Vec
        except ImportError as exc:  # pragma: no cover - gensim is not part of this image
            raise ImportError("embed() needs gensim (the downstream Word2Vec consumer is out of scope "
                              "for the B200 walk engine; simulate_walks works without it)") from exc
        walks = self.simulate_walks_corpus(num_walks, walk_length)
        w2v = Word2Vec(walks, vector_size=dim, window=window_size, sg=1, min_count=0, workers=self.workers,
                       epochs=epochs, seed=self.random_state)
        return w2v.wv[self.nodes]


class _SparseBase(Base, SparseGraph):
    def __init__(self, *args, **kwargs):
        Base.__init__(self, *args, **kwargs)
        self.data = None
        self.indptr = None
        self.indices = None

    def _graph_key(self):
        return tuple((id(a), getattr(a, "shape", None)) for a in (self.indptr, self.indices, self.data))

    def _make_engine(self, device=None):
        from .engine import WalkEngine
        return WalkEngine.from_csr(self.indptr, self.indices, self.data, device=device or self.device)

    def get_has_nbrs(self):
        """Host callable with the meaning of rw/sparse_rw.py:12-20 (the kernels test the same thing themselves)."""
        indptr = self.indptr

        def has_nbrs(idx):
            return bool(indptr[idx] != indptr[idx + 1])

        return has_nbrs


class SparseOTF(_SparseBase):
    """Sparse graph, transition probabilities on the fly (reference pecanpy.py:510-561)."""
    _MODE = "SparseOTF"


class FirstOrderUnweighted(_SparseBase):
    """Uniform first-order walks (reference pecanpy.py:293-309)."""
    _MODE = "FirstOrderUnweighted"


class _AliasTables:
    """alias_j / alias_q as in the reference (populated attributes after preprocess_transition_probs,
    pecanpy.py:505-507): copied from the device on first access."""

    def _fetch_alias(self):
        if self._alias_host is None:
            if self._alias_dev is None:
                return None
            aj, aq = self._alias_dev
            self._alias_host = (aj.cpu().numpy().view(np.uint32).copy(), aq.cpu().numpy().copy())
        return self._alias_host

    @property
    def alias_j(self):
        t = self._fetch_alias()
        return None if t is None else t[0]

    @property
    def alias_q(self):
        t = self._fetch_alias()
        return None if t is None else t[1]

    def fetch_alias_tables(self):
        """(alias_j uint32, alias_q float32) on the host -- same arrays as the two properties."""
        return self.alias_j, self.alias_q


class PreComp(_AliasTables, _SparseBase):
    """Pre-computed 2nd-order alias tables (reference pecanpy.py:364-507); tables are built on the GPU."""
    _MODE = "PreComp"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alias_dim = None
        self.alias_indptr = None
        self._alias_host = None

    def preprocess_transition_probs(self):
        eng = self.engine
        aip, aj, aq = eng.build_alias(self.indptr, self.p, self.q, extend=bool(self.extend))
        self.alias_dim = (self.indptr[1:] - self.indptr[:-1]).astype(np.uint32)
        self.alias_indptr = aip
        self._alias_dev = (aj, aq)
        self._alias_host = None


class PreCompFirstOrder(_AliasTables, _SparseBase):
    """Pre-computed first-order alias tables (reference pecanpy.py:312-361)."""
    _MODE = "PreCompFirstOrder"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._alias_host = None

    def preprocess_transition_probs(self):
        _, aj, aq = self.engine.build_alias(self.indptr, 1.0, 1.0, first_order=True)
        self._alias_dev = (aj, aq)
        self._alias_host = None


class DenseOTF(Base, DenseGraph):
    """Dense graph, transition probabilities on the fly (reference pecanpy.py:564-614)."""
    _MODE = "DenseOTF"

    def __init__(self, *args, **kwargs):
        Base.__init__(self, *args, **kwargs)
        self._data = None
        self._nonzero = None

    def _graph_key(self):
        return ((id(self._data), getattr(self._data, "shape", None)),)

    def _make_engine(self, device=None):
        from .engine import WalkEngine
        return WalkEngine.from_dense(self.data, self.nonzero, device=device or self.device)

    def get_has_nbrs(self):
        """Host callable with the meaning of rw/dense_rw.py:21-32."""
        nonzero = self.nonzero

        def has_nbrs(idx):
            return bool(nonzero[idx].any())

        return has_nbrs
