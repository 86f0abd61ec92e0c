"""The start array of Base.simulate_walks (reference pecanpy.py:135-141) built natively (b2w_shuffled_start) against
NumPy's own legacy shuffle: the array and the state the global generator is left in must be identical -- the shuffle
fixes the row order of the walk matrix."""
import numpy as np
import pytest

from pecanpy_b200.pecanpy import shuffled_start


def reference_start(n, num_walks, seed):
    nodes = np.array(range(n), dtype=np.uint32)
    start = np.concatenate([nodes] * num_walks)
    np.random.seed(seed)
    np.random.shuffle(start)
    return start


@pytest.mark.parametrize("n,num_walks", [(1, 1), (1, 2), (2, 1), (3, 1), (5, 3), (31, 1), (33, 1), (34, 10), (1000, 7),
                                          (65536, 1), (65537, 2), (300000, 3)])
@pytest.mark.parametrize("seed", [0, 1, 12345, 2 ** 32 - 1])
def test_native_shuffle_equals_numpy(n, num_walks, seed):
    from pecanpy_b200 import _capi
    _capi.lib()                                            # the native path must be the one under test
    want = reference_start(n, num_walks, seed)
    state_want = np.random.get_state()
    after_want = np.random.random(3)
    got = shuffled_start(n, num_walks, seed)
    state_got = np.random.get_state()
    after_got = np.random.random(3)
    assert got.dtype == np.uint32 and np.array_equal(got, want)
    assert state_got[0] == state_want[0] and state_got[2] == state_want[2] and np.array_equal(state_got[1], state_want[1])
    assert np.array_equal(after_got, after_want)


def test_native_shuffle_continues_any_generator_state():
    """Seeding is NumPy's (any seed type); the native loop continues from whatever state it left -- checked from a
    state in the middle of a block of 624 words and for an array seed."""
    for seed in ([1, 2, 3, 4], np.arange(7, dtype=np.uint32)):
        want = reference_start(4097, 3, seed)
        got = shuffled_start(4097, 3, seed)
        assert np.array_equal(got, want)
    np.random.seed(5)
    np.random.random(1000)                                 # pos is somewhere inside the block
    st = np.random.get_state()
    a = np.concatenate([np.arange(999, dtype=np.uint32)] * 4)
    np.random.shuffle(a)
    after = np.random.get_state()
    import ctypes as C
    from pecanpy_b200 import _capi
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(st[2]))
    b = np.empty(999 * 4, dtype=np.uint32)
    assert _capi.lib().b2w_shuffled_start(999, 4, C.c_void_p(key.ctypes.data), C.byref(pos), C.c_void_p(b.ctypes.data)) == 0
    assert np.array_equal(a, b) and pos.value == after[2] and np.array_equal(key, after[1])


def test_unseeded_start_is_a_permutation():
    s = shuffled_start(1000, 3, None)                      # random_state=None: nondeterministic, like the reference
    assert np.array_equal(np.sort(s), np.repeat(np.arange(1000, dtype=np.uint32), 3))


def test_matches_committed_fixture():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "karate_sparseotf_p1_q1.npz"))
    n = z["indptr"].size - 1
    assert np.array_equal(shuffled_start(n, int(z["num_walks"]), int(z["seed"])), z["start"])
