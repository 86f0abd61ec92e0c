"""b2w_replay.cuh (the exact emulation of the reference's sequential float32 cumsum used by the unweighted SparseOTF
kernels when their filter is inconclusive) against genuine float32 additions, on the CPU: the header is compiled
with g++ through a shim for the CUDA intrinsics (tests/cuda_host_shim.h).  The kernels that run the same header are
checked on the GPU against the oracle (tests/test_gpu_parity.py, forced-replay variants)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rp") / "replay_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "replay_harness.cpp")], check=True)
    L = C.CDLL(so)
    L.h_udiv24.restype = C.c_uint32
    L.h_udiv24.argtypes = [C.c_uint32, C.c_uint32]
    L.h_upper_float.restype = C.c_float
    L.h_upper_float.argtypes = [C.c_double]
    for f in (L.h_advance_run, L.h_brute_run):
        f.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_uint32, C.c_float, C.c_double, C.POINTER(C.c_uint32)]
    return L


def test_udiv24_is_exact(lib):
    rng = np.random.default_rng(0)
    for _ in range(200000):
        b = int(rng.integers(1, 1 << 23))
        a = int(rng.integers(0, 1 << 24)) if b > 1 else int(rng.integers(0, 1 << 24))
        if b > 1 and rng.random() < 0.3:              # exact multiples and their neighbours
            a = min((1 << 24) - 1, b * int(rng.integers(0, (1 << 24) // b + 1)) + int(rng.integers(-1, 2)))
            a = max(a, 0)
        assert lib.h_udiv24(a, b) == a // b, (a, b)
    for a in (0, 1, (1 << 24) - 1, (1 << 23), (1 << 23) - 1):
        for b in (1, 2, 3, 7, (1 << 23) - 1):
            assert lib.h_udiv24(a, b) == a // b


def test_upper_float_is_the_smallest_float_not_below_u(lib):
    rng = np.random.default_rng(1)
    for u in list(rng.random(2000)) + [0.0, 1.0 - 2.0 ** -53, 0.5, float(np.float32(0.3)), 2.0 ** -30]:
        ub = np.float32(lib.h_upper_float(u))
        assert float(ub) >= u
        if ub > 0:
            assert float(np.nextafter(ub, np.float32(0))) < u


def run(lib, fn, cdf0, k0, n, fo, u):
    cdf, k, ch = C.c_float(cdf0), C.c_uint32(k0), C.c_uint32(0xFFFFFFFF)
    found = fn(C.byref(cdf), C.byref(k), n, C.c_float(fo), u, C.byref(ch))
    return (1, ch.value) if found else (0, np.float32(cdf.value).view(np.uint32), k.value)


def test_advance_run_equals_genuine_additions(lib):
    """Runs of one addend from many starting prefixes, thresholds inside / outside the run and exactly on values the
    prefix takes; addends shaped like the kernels' (w / S for small integer w and S up to 2^24 grid units)."""
    rng = np.random.default_rng(2)
    for it in range(4000):
        Wd = int(rng.integers(1, 70000))
        a = int(rng.choice([1, 2, 3, 4, 16]))
        g = np.float32(2.0 ** int(rng.integers(-4, 1)))
        S = np.float32(np.float32(Wd) * g)
        fo = np.float32(np.float32(np.float32(a) * g) / S)
        # a starting prefix the recurrence can actually reach: a few genuine additions of another addend
        cdf0 = np.float32(0.0)
        other = np.float32(np.float32(np.float32(int(rng.choice([1, 4, 16]))) * g) / S)
        for _ in range(int(rng.integers(0, 40))):
            cdf0 = np.float32(cdf0 + other)
        n = int(rng.integers(1, 6000))
        end = float(cdf0) + n * float(fo)
        kind = rng.integers(0, 4)
        if kind == 0:
            u = float(rng.random())
        elif kind == 1:
            u = min(float(cdf0) + rng.random() * n * float(fo), 1.0 - 2.0 ** -53)
        elif kind == 2:
            u = min(end * (1.0 + 1e-6), 1.0 - 2.0 ** -53)
        else:                                          # exactly a value of the prefix (ties of the comparison)
            c = np.float32(cdf0)
            for _ in range(int(rng.integers(1, min(n, 300) + 1))):
                c = np.float32(c + fo)
            u = min(float(c), 1.0 - 2.0 ** -53)
        if not float(cdf0) < u:
            continue
        got = run(lib, lib.h_advance_run, cdf0, 7, n, fo, u)
        want = run(lib, lib.h_brute_run, cdf0, 7, n, fo, u)
        assert got == want, (Wd, a, float(g), float(cdf0), n, float(fo), u, got, want)
