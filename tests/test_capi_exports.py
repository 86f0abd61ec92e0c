"""The C-ABI library loads and exports every symbol that include/b2w.h declares (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2w.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2w_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from pecanpy_b200 import _capi
    lib = _capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"libb2w.so does not export {s}"
    assert sorted(_capi.EXPORTS) == syms
    assert lib.b2w_version() == 100


def test_no_cpu_fallback_without_gpu():
    """Product entry points fail loudly when no CUDA device is present."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from pecanpy_b200 import pecanpy as b2
    g = b2.SparseOTF.from_mat(np.array([[0, 1], [1, 0]]), ["a", "b"], p=1, q=1, random_state=0)
    with pytest.raises(RuntimeError, match="CUDA"):
        g.simulate_walks(1, 2)


def test_product_never_imports_the_oracle():
    """Nothing under pecanpy_b200/ may import, load or link the CPU oracle."""
    pkg = os.path.join(ROOT, "pecanpy_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|liboracle|walk_oracle|oracle\.(oracle|walk_|alias_)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} references the oracle"
