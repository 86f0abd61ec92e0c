"""pecanpy_b200 -- a B200-native (sm_100a) biased random-walk engine behind PecanPy's
``simulate_walks`` API.

Only the walk-generation hot path of krishnanlab/PecanPy is implemented here:
``pecanpy_b200.pecanpy.{SparseOTF, PreComp, DenseOTF, FirstOrderUnweighted, PreCompFirstOrder}``
keep the reference's constructors, loaders and ``simulate_walks`` / ``preprocess_transition_probs``
/ ``embed`` surface; the work is done by hand-written CUDA kernels in ``csrc/`` behind the C ABI
declared in ``include/b2w.h``.
"""
from . import graph  # noqa: F401
from . import pecanpy  # noqa: F401

version = "0.1.0"
__all__ = ["graph", "pecanpy"]
