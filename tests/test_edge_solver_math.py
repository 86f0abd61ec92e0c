"""CPU restatement of the closed-form step of b2w_walk_edge.cu (StepDist::first_at_least and the two thresholds of
edge_step) checked against (a) a brute-force scan of the exact integer prefix W(k) and (b) the reference's own
arithmetic -- float32 probabilities, sequential float32 cumsum, searchsorted (pecanpy.py:556-557,
rw/sparse_rw.py:86-89) -- including uniforms placed right at the cdf boundaries.  Whenever the filter claims a
choice it must be the reference's; otherwise the kernel replays exactly (covered on the GPU)."""
import math
import random

import numpy as np

NONE = 0xFFFFFFFF
EC = 1.02 * 2.0 ** -24


class StepDist:
    def __init__(self, d, lst, kp, a_in, a_out, a_ret, has_prev=True):
        self.d, self.lst, self.m, self.kp = d, lst, len(lst), kp
        self.a_o = a_out if has_prev else a_in
        self.da, self.dr = a_in - self.a_o, a_ret - self.a_o
        self.inv = 1.0 / self.a_o

    def W(self, k, c):
        return (k + 1) * self.a_o + c * self.da + (self.dr if self.kp <= k else 0)

    def ceil_div(self, need):
        qv = int(need * self.inv)                       # __double2int_rz
        if qv * self.a_o < need:
            qv += 1
        if (qv - 1) * self.a_o >= need:
            qv -= 1
        return qv

    def range(self, lo, hi, B0, T):
        kp_in = hi >= 0 and self.kp <= hi
        kp_i = self.kp if self.kp != NONE else -1
        hiA = kp_i - 1 if kp_in else hi
        kA = max(self.ceil_div(T - B0) - 1, lo)
        kB = max(self.ceil_div(T - B0 - self.dr) - 1, max(lo, kp_i))
        return kA if kA <= hiA else (kB if (kp_in and kB <= hi) else -1)

    def first_at_least(self, T):
        lo, hi = 0, self.m
        while lo < hi:
            mid = (lo + hi) >> 1
            if self.W(self.lst[mid], mid + 1) >= T:
                hi = mid
            else:
                lo = mid + 1
        i = lo
        klo = self.lst[i - 1] + 1 if i else 0
        pi = self.lst[i] if i < self.m else self.d
        r = self.range(klo, pi - 1, i * self.da, T)
        if r >= 0:
            k, c = r, i
        elif i < self.m:
            k, c = pi, i + 1
        else:
            k, c = self.d - 1, self.m
        return k, self.W(k, c)

    def brute(self, T):
        c = 0
        for k in range(self.d):
            c += k in self.lst
            if self.W(k, c) >= T:
                return k, self.W(k, c)
        return self.d - 1, self.W(self.d - 1, self.m)

    def filter(self, u):
        """edge_step without the replay: (choice, proven?)"""
        Wd = self.d * self.a_o + self.m * self.da + (self.dr if self.kp != NONE else 0)
        A = u * float(Wd)
        e_row = EC * (self.d + 2)
        t_hi = A * (1.0 + e_row + 2.0 * e_row * e_row + 2.9e-14)
        slack = (self.m * -self.da if self.da < 0 else 0) + (-self.dr if (self.dr < 0 and self.kp != NONE) else 0)
        k_up = (t_hi + slack) * self.inv + 1.0
        k_hi = int(k_up) if k_up < self.d - 1 else self.d - 1
        e = EC * (k_hi + 3)
        T_poss = math.ceil(A * (1.0 - e - 2.9e-14))
        t_sure = math.ceil(A * (1.0 + e + 2.0 * e * e + 2.9e-14))
        k, Wk = self.first_at_least(T_poss)
        assert k <= k_hi
        return k, Wk >= t_sure


def random_case(rng, dmax):
    d = rng.randint(1, dmax)
    a_in, a_out, a_ret = rng.choice([(4, 16, 1), (2, 1, 4), (1, 1, 1), (1, 4, 16), (2, 4, 1), (3, 2, 6), (16, 1, 4)])
    kp = rng.choice([NONE] + list(range(d)))
    dens = rng.choice([0.0, 0.0, 0.1, 0.5, 0.95])
    lst = [k for k in range(d) if k != kp and rng.random() < dens]
    return d, lst, kp, a_in, a_out, a_ret


def test_first_at_least_equals_brute_force():
    rng = random.Random(1)
    for _ in range(60000):
        d, lst, kp, a_in, a_out, a_ret = random_case(rng, 48)
        s = StepDist(d, lst, kp, a_in, a_out, a_ret)
        Wd = s.W(d - 1, len(lst))
        T = rng.randint(-3, Wd)
        assert s.first_at_least(T) == s.brute(T), (d, lst, kp, (a_in, a_out, a_ret), T)


def reference_choice(d, lst, kp, w_in, w_out, w_ret, u):
    w = np.full(d, w_out, dtype=np.float32)
    w[lst] = w_in
    if kp != NONE:
        w[kp] = w_ret
    s = np.float32(0.0)
    for x in w:                                        # numba's sequential float32 sum
        s = np.float32(s + x)
    probs = (w / s).astype(np.float32)
    cdf = np.float32(0.0)
    out = np.empty(d, dtype=np.float32)
    for i, x in enumerate(probs):                      # sequential float32 cumsum
        cdf = np.float32(cdf + x)
        out[i] = cdf
    return int(np.searchsorted(out.astype(np.float64), u, side="left")), out


def test_proven_choices_are_the_reference_choices():
    rng = random.Random(2)
    proven = total = 0
    for _ in range(3000):
        d, lst, kp, a_in, a_out, a_ret = random_case(rng, 300)
        g = 2.0 ** rng.choice([-4, -2, 0])            # the common grid of the three weights
        s = StepDist(d, lst, kp, a_in, a_out, a_ret)
        _, cdf = reference_choice(d, lst, kp, a_in * g, a_out * g, a_ret * g, 0.5)
        us = [(rng.random(), True) for _ in range(6)] + [(0.0, False), (1.0 - 2.0 ** -53, False)]
        for k in rng.sample(range(d), min(d, 4)):      # adversarial: at and next to a cdf boundary
            c = float(cdf[k])
            us += [(c, False), (float(np.nextafter(c, 0.0)), False), (float(np.nextafter(c, 1.0)), False)]
        for u, is_random in us:
            if not (0.0 <= u < 1.0):
                continue
            want = int(np.searchsorted(cdf.astype(np.float64), u, side="left"))
            k, sure = s.filter(u)
            total += is_random
            if sure:
                proven += is_random
                assert k == want, (d, lst, kp, (a_in, a_out, a_ret), u, k, want)
    assert proven > 0.97 * total                       # random uniforms: the filter decides; boundaries go to the replay
