"""Compare the public API surface of the drop-in modules with the reference's by introspection (build container only):
every public attribute of the reference's walk classes and graph containers, and for callables the parameter names,
kinds and defaults.  Prints a JSON object {"missing": {...}, "signatures": {...}}; tests/test_oracle_vs_reference_live.py
pins the (short, deliberate) list of differences."""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_stubs"), "/root/reference/src", os.path.dirname(HERE)]

from pecanpy import graph as RG, pecanpy as R  # noqa: E402  (the reference)
from pecanpy_b200 import graph as OG, pecanpy as O  # noqa: E402


def public(cls):
    return {n for n in dir(cls) if not n.startswith("__")}


def params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([p.name, p.kind.name, d])
    return out


def main():
    missing, sigs = {}, {}
    pairs = [(getattr(R, n), getattr(O, n), n) for n in ("SparseOTF", "PreComp", "DenseOTF", "FirstOrderUnweighted",
                                                          "PreCompFirstOrder")]
    pairs += [(getattr(RG, n), getattr(OG, n), n) for n in ("BaseGraph", "AdjlstGraph", "SparseGraph", "DenseGraph")]
    for r, o, name in pairs:
        gone = sorted(public(r) - public(o))
        if gone:
            missing[name] = gone
        for m in sorted(public(r) & public(o)):
            a, b = inspect.getattr_static(r, m), inspect.getattr_static(o, m)
            fa = a.__func__ if isinstance(a, (staticmethod, classmethod)) else a
            fb = b.__func__ if isinstance(b, (staticmethod, classmethod)) else b
            fa = getattr(fa, "py_func", fa)                    # numba dispatcher -> the Python function
            if inspect.isfunction(fa) and inspect.isfunction(fb):
                pa, pb = params(fa), params(fb)
                if pa != pb:
                    sigs[f"{name}.{m}"] = {"reference": pa, "ours": pb}
    print(json.dumps({"missing": missing, "signatures": sigs}))


if __name__ == "__main__":
    main()
