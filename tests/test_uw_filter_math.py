"""The integer-domain filter of the unweighted SparseOTF kernel (pecanpy_b200/csrc/b2w_walk_uw.cu, phase 3),
restated in NumPy and checked against the reference's own arithmetic on the CPU.

The kernel never forms the reference's f32 cumulative distribution on its common path; it proves the result of
``searchsorted(cumsum(probs), u)`` (pecanpy.py:556-557) from exact integer prefixes W_k and two thresholds

    W_k <  u W_d (1 - e)            =>  cdf_k <  u
    W_k >= u W_d (1 + e + 2 e^2)    =>  cdf_k >= u        e = 1.02 (k_max + 3) 2^-24

and replays the f32 recurrence only when neither holds.  This test evaluates exactly those formulas (same
constants, same word / position structure) and requires that every decision the filter would PROVE equals the
reference computed with sequential float32 sums -- including uniforms placed on and one ulp around cdf values."""
import math

import numpy as np
import pytest

EC = 1.02 * 2.0 ** -24          # the kernel's constant
PUSH = 2.9e-14                  # the kernel's outward push of the thresholds


def reference_choice(weights_f32, u):
    """rw/sparse_rw.py:89 + pecanpy.py:556-557 with Numba's semantics: sequential f32 sum, f32 cumsum, bisect left."""
    total = np.float32(0)
    for x in weights_f32:
        total = np.float32(total + x)
    probs = (weights_f32 / total).astype(np.float32)
    cdf = np.cumsum(probs, dtype=np.float32)            # sequential in the array dtype
    return int(np.searchsorted(cdf, u, side="left")), cdf


def kernel_filter(d, common, kp, a_in, a_out, a_ret, u, has_prev=True):
    """uw_step phase 2 + 3.  Returns the proven choice or None (= the kernel would run the exact replay)."""
    a_o = a_out if has_prev else a_in
    da, dr = a_in - a_o, a_ret - a_o
    h = 0 if kp is None else 1
    Wd = d * a_o + int(common.sum()) * da + h * dr
    A = u * float(Wd)
    k = np.arange(d)
    Wk = (k + 1) * a_o + np.cumsum(common) * da + (((k >= kp) * dr) if kp is not None else 0)
    assert Wk[-1] == Wd and Wd < 2 ** 24
    nwords = (d + 31) // 32
    wsel = 0
    if nwords > 1:
        t_row = A * (1.0 - EC * (d + 2) - PUSH)
        wsel = None
        for w in range(nwords):
            kend = min(d, (w + 1) * 32) - 1
            if float(Wk[kend]) >= t_row:
                wsel = w
                break
        if wsel is None:
            return None
    nb = min(32, d - wsel * 32)
    e_w = EC * (wsel * 32 + nb + 2)
    t_poss = A * (1.0 - e_w - PUSH)
    t_sure = A * (1.0 + e_w + 2.0 * e_w * e_w + PUSH)
    for b in range(nb):
        kk = wsel * 32 + b
        if float(Wk[kk]) >= t_poss:
            return kk if float(Wk[kk]) >= t_sure else None
    return None


@pytest.mark.parametrize("p,q", [(4, 0.25), (0.5, 2), (1, 1), (0.25, 4), (2, 0.5), (8, 0.125), (1, 0.5)])
def test_integer_filter_only_proves_what_the_reference_computes(p, q):
    rng = np.random.default_rng(int(p * 1000 + q * 10))
    w_in, w_out, w_ret = np.float32(1), np.float32(1 / q), np.float32(1 / p)
    g = min(1.0, float(w_out), float(w_ret))           # the weight grid (a power of two for these p, q)
    a_in, a_out, a_ret = int(1 / g), int(float(w_out) / g), int(float(w_ret) / g)
    assert a_in * g == 1 and a_out * g == float(w_out) and a_ret * g == float(w_ret)
    proven = replays = 0
    for d in [1, 2, 5, 17, 31, 32, 33, 64, 65, 100, 500, 1500]:
        for density in [0.0, 0.05, 0.3, 0.9]:
            common = (rng.random(d) < density).astype(np.int64)
            kp = int(rng.integers(0, d)) if rng.random() < 0.9 else None
            if kp is not None:
                common[kp] = 0
            w = np.where(common == 1, w_in, w_out).astype(np.float32)
            if kp is not None:
                w[kp] = w_ret
            _, cdf = reference_choice(w, 0.5)
            us = list(rng.random(12))
            for kk in rng.integers(0, d, size=6):       # adversarial uniforms: on and around a cdf value
                c = float(cdf[kk])
                us += [c, float(np.nextafter(np.float64(c), 0.0)), float(np.nextafter(np.float64(c), 2.0)),
                       c * (1 - 1e-9), c * (1 + 1e-9)]
            for u in us:
                if not 0.0 <= u < 1.0:
                    continue
                want = int(np.searchsorted(cdf, u, side="left"))
                got = kernel_filter(d, common, kp, a_in, a_out, a_ret, float(u))
                if got is None:
                    replays += 1
                else:
                    proven += 1
                    assert got == want, (d, kp, u, got, want)
    assert proven > replays / 4                          # the filter is not vacuous


def test_random_uniforms_rarely_need_the_replay():
    """With u uniform the replay probability is ~ e * W_d per step: a handful in thousands of steps."""
    rng = np.random.default_rng(3)
    a_in, a_out, a_ret = 4, 16, 1                        # p = 4, q = 0.25 (BASELINE config #3), g = 0.25
    replays = total = 0
    for _ in range(300):
        d = int(rng.integers(1, 200))
        common = (rng.random(d) < 0.2).astype(np.int64)
        kp = int(rng.integers(0, d))
        common[kp] = 0
        w = np.where(common == 1, np.float32(1), np.float32(4)).astype(np.float32)
        w[kp] = np.float32(0.25)
        _, cdf = reference_choice(w, 0.5)
        for u in rng.random(10):
            got = kernel_filter(d, common, kp, a_in, a_out, a_ret, float(u))
            total += 1
            if got is None:
                replays += 1
            else:
                assert got == int(np.searchsorted(cdf, u, side="left"))
    assert replays <= max(3, total // 200)
