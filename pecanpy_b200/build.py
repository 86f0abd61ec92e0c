"""In-tree build of libb2w.so (hand-written CUDA for sm_100a) with nvcc.

``python -m pecanpy_b200.build`` or ``__graft_entry__.build()``.  The shared library is written
next to the package (``pecanpy_b200/lib/libb2w.so``) so that it travels with the source tree;
nothing is installed into site-packages and nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libb2w.so")
OBJ_DIR = os.path.join(os.path.dirname(HERE), "build", "obj")

SOURCES = ["b2w_api.cu", "b2w_alias.cu", "b2w_dense.cu", "b2w_walk_thread.cu", "b2w_walk_warp.cu", "b2w_walk_uw.cu",
           "b2w_thresholds.cu", "b2w_csr_build.cu", "b2w_edgelist.cu", "b2w_edge_index.cu", "b2w_walk_edge.cu", "b2w_shared.cu", "b2w_wedge.cu", "b2w_start.cu"]
# -fmad=false: the reference rounds every multiply and add separately (no FMA contraction);
# no fast-math: IEEE division and no flush-to-zero are part of the bit-exactness contract.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


def pylists_path() -> str:
    import sysconfig
    return os.path.join(LIB_DIR, "_b2w_pylists" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pylists(force: bool = False) -> str:
    """The CPython extension that turns a walk matrix into List[List[str]] (csrc/b2w_pylists.c), with gcc."""
    import sysconfig
    src, out = os.path.join(CSRC, "b2w_pylists.c"), pylists_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", "-I", sysconfig.get_paths()["include"], src, "-o",
           out + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed for b2w_pylists.c:\n{r.stdout}\n{r.stderr}")
    os.replace(out + ".tmp", out)
    return out


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = LIB, obj_dir: str = OBJ_DIR) -> str:
    """Compile every .cu for sm_100a and link the shared library.  `extra_flags` / `out` / `obj_dir` build a tuning
    variant (e.g. ``-DB2W_UW_NESTED=0``) next to the product library; see tools/build_variants.py."""
    build_pylists(force)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _deps_mtime():
        return out
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    env.pop("CC", None)   # the image's $CC is not a usable nvcc host compiler selector

    def compile_one(src: str):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        return src, obj, r

    objs = []
    log = []
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            log.append(f"== {src}\n{r.stderr}")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            objs.append(obj)
    with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    tmp = out + ".tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
