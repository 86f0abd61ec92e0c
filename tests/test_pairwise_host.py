"""b2w_pairwise.cuh (the arithmetic of the noise-threshold kernels) against NumPy itself, on the CPU.

The header is __host__ __device__; this test compiles it with g++ (no GPU needed) and requires bit equality
with the reference's expression ``row.mean() + gamma * row.std()`` (rw/sparse_rw.py:22-35,
rw/dense_rw.py:11-19) evaluated by NumPy.  The CUDA kernels that run the same header are checked on the GPU
in tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pw") / "pairwise_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "pairwise_harness.cpp")], check=True)
    lib = C.CDLL(so)
    lib.h_sum_f32.restype = C.c_float
    return lib


def _numpy_thr_csr(indptr, data, gamma):
    thr = np.zeros(indptr.size - 1, dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(thr.size):
            row = data[indptr[i]:indptr[i + 1]]
            thr[i] = row.mean() + gamma * row.std()
    return np.maximum(thr, 0)


def _numpy_thr_dense(data, nz, gamma):
    thr = np.zeros(data.shape[0], dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(thr.size):
            w = data[i, nz[i]]
            thr[i] = w.mean() + gamma * w.std()
    return np.maximum(thr, 0)


def _same(a, b):
    nan = np.isnan(a)
    return np.array_equal(nan, np.isnan(b)) and np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


def test_pairwise_sum_matches_numpy_add_reduce(harness):
    rng = np.random.default_rng(0)
    for n in list(range(0, 300)) + [1000, 1023, 1024, 1025, 4097, 8191, 8192, 8193, 20000, 100001]:
        a = (rng.random(n + 3, dtype=np.float32) * np.float32(3) + np.float32(0.01))[n % 3:][:n]   # misaligned too
        a = np.ascontiguousarray(a)
        got = np.float32(harness.h_sum_f32(a.ctypes.data_as(C.c_void_p), C.c_uint32(n)))
        want = a.sum() if n else np.float32(0)
        assert got.view(np.uint32) == np.float32(want).view(np.uint32), n


@pytest.mark.parametrize("gamma", [0, 0.0, 0.25, 1.0, -0.7, 3.3])
def test_csr_thresholds_match_numpy(harness, gamma):
    rng = np.random.default_rng(1)
    deg = np.concatenate([rng.integers(0, 40, 400), rng.integers(100, 700, 40), [0, 1, 7, 8, 9, 127, 128, 129, 5000]])
    indptr = np.zeros(deg.size + 1, dtype=np.uint32)
    np.cumsum(deg, out=indptr[1:])
    data = (np.float32(0.01) + np.float32(0.99) * rng.random(int(indptr[-1]), dtype=np.float32)).astype(np.float32)
    data[indptr[5]:indptr[6]] = np.float32(2.5)            # a constant row: std exactly 0
    got = np.zeros(deg.size, dtype=np.float32)
    harness.h_thr_csr(C.c_uint32(deg.size), indptr.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                      C.c_double(gamma), got.ctypes.data_as(C.c_void_p))
    assert _same(got, _numpy_thr_csr(indptr, data, gamma))


@pytest.mark.parametrize("gamma", [0, 0.5, -2.0])
def test_dense_thresholds_match_numpy(harness, gamma):
    rng = np.random.default_rng(2)
    n = 300
    nz = rng.random((n, n)) < 0.4
    nz[3] = False                                          # a row without neighbours: NaN
    nz[4] = False; nz[4, 10] = True
    nz[5] = True
    data = np.where(nz, 0.01 + 0.99 * rng.random((n, n)), 0.0)
    got = np.zeros(n, dtype=np.float32)
    harness.h_thr_dense(C.c_uint32(n), data.ctypes.data_as(C.c_void_p), nz.view(np.uint8).ctypes.data_as(C.c_void_p),
                        C.c_double(gamma), got.ctypes.data_as(C.c_void_p))
    assert _same(got, _numpy_thr_dense(data, nz, gamma))


def test_thresholds_equal_reference_fixtures(harness, golden_dir):
    """`thr` arrays written by the UNMODIFIED reference's get_noise_thresholds (oracle/gen_golden.py)."""
    import glob
    seen = 0
    for f in sorted(glob.glob(os.path.join(golden_dir, "*_ext*.npz"))):
        c = np.load(f)
        gamma = float(c["gamma"])
        want = c["thr"].astype(np.float32)
        got = np.zeros_like(want)
        if "dense" in c.files:
            data = np.ascontiguousarray(c["dense"], dtype=np.float64)
            nz = np.ascontiguousarray(c["nonzero"]).view(np.uint8)
            harness.h_thr_dense(C.c_uint32(want.size), data.ctypes.data_as(C.c_void_p), nz.ctypes.data_as(C.c_void_p),
                                C.c_double(gamma), got.ctypes.data_as(C.c_void_p))
        else:
            indptr = np.ascontiguousarray(c["indptr"], dtype=np.uint32)
            data = np.ascontiguousarray(c["data"], dtype=np.float32)
            harness.h_thr_csr(C.c_uint32(want.size), indptr.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p),
                              C.c_double(gamma), got.ctypes.data_as(C.c_void_p))
        assert _same(got, want), os.path.basename(f)
        seen += 1
    assert seen >= 4
