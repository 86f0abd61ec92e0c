"""The row writers of the lane-per-walker kernels (pecanpy_b200/csrc/b2w_rowout.cuh) executed on the CPU: the header is
compiled with g++ (tests/rowtile_harness.cpp supplies threadIdx / __stcs / __syncwarp and runs the 32 lanes of a warp one
after the other; every global store is recorded).  WarpRowTile -- the coalesced writer of the mirrored multi-GPU kernel
-- must store every word of every row exactly once, with the right value, in the local matrix and in every mirror,
and touch nothing else; RowWriter must produce the rows with 16-byte stores that are aligned and never overlap, for every row phase / row length / leading
dimension.  The GPU suite checks the same through the kernels (tests/test_gpu_multi.py, tests/test_gpu_parity.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rt") / "rowtile_harness.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "rowtile_harness.cpp")],
                   check=True)
    L = C.CDLL(so)
    u32 = C.c_uint32
    L.h_rowtile_run.argtypes = [u32, u32, u32, u32, u32, C.c_int, u32, C.c_void_p, C.c_void_p]
    L.h_rowwriter_run.argtypes = [u32, u32, u32, u32, u32, u32, C.c_void_p, C.c_void_p]
    return L


def expected(L, ld, row0, n_rows, total_rows):
    want = np.full((total_rows, ld), 0xDEADBEEF, dtype=np.uint32)
    cnt = np.zeros((total_rows, ld), dtype=np.uint32)
    j = np.arange(L + 2, dtype=np.uint32)
    for r in range(row0, row0 + n_rows):
        want[r, :L + 2] = 0x10000 * (r + 1) + j + 1
        cnt[r, :L + 2] = 1
    return want, cnt


@pytest.mark.parametrize("L", [1, 5, 21, 22, 23, 30, 31, 46, 47, 80, 95, 200])
@pytest.mark.parametrize("pad", [0, 3])
def test_warp_row_tile_stores_every_word_once(lib, L, pad):
    ld = L + 2 + pad
    for n_rows, n_mirrors in ((32, 0), (32, 1), (7, 3), (1, 7), (32, 7)):
        for phase in range(8):
            row0, total = 2, 2 + 32 + 1
            words = total * ld
            vals = np.zeros((1 + n_mirrors, words), dtype=np.uint32)
            cnts = np.zeros((1 + n_mirrors, words), dtype=np.uint32)
            stray = lib.h_rowtile_run(L, ld, row0, n_rows, total, n_mirrors, phase, vals.ctypes.data, cnts.ctypes.data)
            want, cnt = expected(L, ld, row0, n_rows, total)
            assert stray == 0
            for b in range(1 + n_mirrors):
                assert np.array_equal(vals[b].reshape(total, ld), want), (L, pad, n_rows, n_mirrors, phase, b)
                assert np.array_equal(cnts[b].reshape(total, ld), cnt), (L, pad, n_rows, n_mirrors, phase, b)


@pytest.mark.parametrize("L", [1, 5, 6, 7, 8, 30, 31, 80, 81])
@pytest.mark.parametrize("pad", [0, 1, 5])
def test_row_writer_stores_every_row(lib, L, pad):
    ld = L + 2 + pad
    for phase in range(8):
        row0, n_rows, total = 1, 32, 34
        words = total * ld
        vals = np.zeros(words, dtype=np.uint32)
        cnts = np.zeros(words, dtype=np.uint32)
        stray = lib.h_rowwriter_run(L, ld, row0, n_rows, total, phase, vals.ctypes.data, cnts.ctypes.data)
        want, cnt = expected(L, ld, row0, n_rows, total)
        assert stray == 0                                  # (also: every 16-byte store was 16-byte aligned)
        assert np.array_equal(vals.reshape(total, ld), want), (L, pad, phase)
        # (the head and the tail of a row leave by plain word stores, which the harness does not count: the sector
        # stores must not write any word twice or outside the rows)
        assert (cnts.reshape(total, ld) <= cnt).all(), (L, pad, phase)
