#!/usr/bin/env python
"""Build tuning variants of libb2w.so (same sources, different -D) for A/B runs on the GPU box.

usage: python tools/build_variants.py NAME=-DFOO=1,-DBAR=0 [NAME2=...]
Each variant lands in pecanpy_b200/lib/variants/libb2w_NAME.so; run it with B2W_LIBRARY=<path> python bench.py ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pecanpy_b200 import build as b  # noqa: E402

for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(b.LIB_DIR, "variants", f"libb2w_{name}.so")
    b.build(force=True, extra_flags=[f for f in flags.split(",") if f], out=out,
            obj_dir=os.path.join(ROOT, "build", f"obj_{name}"))
    print(out)
