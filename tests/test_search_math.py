"""The branch-free search of the membership phase (pecanpy_b200/csrc/b2w_membership.cuh, lower_bound_eq), restated
in Python and checked exhaustively: k + 1 probes, none out of bounds, the answer equals bisect_left, and "x is in
the row" equals "the last probe that loaded a value >= x loaded x" (no final dependent load)."""
import bisect

import numpy as np

NONE = 0xFFFFFFFF


def lower_bound_eq(row, x):
    n = len(row)
    k = n.bit_length() - 1                      # floor(log2 n), n >= 1
    top = 1 << k
    probes = [top - 1]
    v = row[top - 1]
    ge = NONE if v < x else v
    lo = n - top + 1 if v < x else 0
    s = top >> 1
    while s:
        i = lo + s - 1
        assert 0 <= i < n, "probe out of bounds"
        probes.append(i)
        v = row[i]
        if v < x:
            lo += s
        else:
            ge = v
        s >>= 1
    assert len(probes) == k + 1
    return lo, ge == x


def test_search_is_exact_for_every_row_length_and_key():
    rng = np.random.default_rng(0)
    for n in list(range(1, 70)) + [127, 128, 129, 255, 256, 257, 1000]:
        row = sorted(int(v) for v in rng.choice(4 * n + 8, size=n, replace=False))
        keys = set(row) | {v + 1 for v in row} | {v - 1 for v in row if v > 0} | {0, 4 * n + 9}
        for x in keys:
            pos, found = lower_bound_eq(row, x)
            assert pos == bisect.bisect_left(row, x), (n, x)
            assert found == (x in row), (n, x)


def test_invalid_lane_key_walks_to_the_end():
    # idle lanes search for 0xFFFFFFFF (larger than every node index): position n; the "found" flag compares the
    # sentinel with itself, which is why every caller masks it with the lane's `valid` bit
    row = [3, 9, 27, 81]
    pos, found = lower_bound_eq(row, NONE)
    assert pos == 4 and found


def inreg_found(prow, x):
    """B2W_INREG_MEMBERSHIP (off by default): row(prev) padded to 32 register lanes with the sentinel, five
    shuffle-probes of a power-of-two lower_bound, then one equality probe."""
    pv = list(prow) + [NONE] * (32 - len(prow))
    base, half = 0, 16
    while half:
        if pv[base + half - 1] < x:
            base += half
        half >>= 1
    assert 0 <= base < 32
    return pv[base] == x


def test_in_register_membership_is_exact_for_rows_up_to_32():
    rng = np.random.default_rng(1)
    for n in range(1, 33):
        for _ in range(20):
            row = sorted(int(v) for v in rng.choice(200, size=n, replace=False))
            for x in range(0, 201):
                assert inreg_found(row, x) == (x in row), (row, x)
