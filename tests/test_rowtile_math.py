"""Index arithmetic of WarpRowTile (pecanpy_b200/csrc/b2w_rowout.cuh), restated: a ring of 32 words per walker, flushed
every MIRROR_PERIOD words in sector-aligned pieces of at most 32 words.  For every row phase and row length: each word
of the row is stored exactly once, with the value that was put, never from a ring slot that has been overwritten, and
every piece except the head and the tail of the row starts and ends on a 32-byte boundary."""
import pytest

PERIOD = 24


def run(ph, L):
    ld = L + 2
    ring = [None] * 32
    flushed = 0
    stored = {}
    pieces = []

    def flush(have, last):
        nonlocal flushed
        a = flushed - ((ph + flushed) & 7) if flushed else 0
        b = have if last else have - ((ph + have) & 7)
        assert b - a <= 32
        for lane in range(32):
            t = a + lane
            if t < b:
                v = ring[t & 31]
                assert v == ("w", t), (ph, L, t, v)
                assert t not in stored
                stored[t] = v
        if b > a:
            pieces.append((a, b))
        flushed = have

    ring[0] = ("w", 0)
    since = 1
    for j in range(1, L + 1):
        ring[j & 31] = ("w", j)
        since += 1
        if since == PERIOD:
            flush(j + 1, False)
            since = 0
    ring[(L + 1) & 31] = ("w", L + 1)
    flush(L + 2, True)
    assert sorted(stored) == list(range(ld))
    for a, b in pieces:
        assert a == 0 or (ph + a) % 8 == 0
        assert b == ld or (ph + b) % 8 == 0


@pytest.mark.parametrize("L", [1, 2, 5, 6, 21, 22, 23, 24, 30, 31, 46, 47, 70, 80, 81, 94, 95, 200])
def test_every_word_stored_once(L):
    for ph in range(8):
        run(ph, L)
