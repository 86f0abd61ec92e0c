#!/usr/bin/env python
"""Offline cost model of the unweighted SparseOTF step on BASELINE config #3 (no GPU needed).

Walks a sample of the real job with the CPU oracle (tools are not product code), then replays every step through the
kernel's membership policy (b2w_membership.cuh: direction choice, chunks of G keys, k+1 probes per search, N chunks
interleaved per iteration) and reports, per step class, the number of probe instructions and the length of the
dependent-load chain a warp waits for.  This is a MODEL (it counts, it does not time); it exists to size ideas
before they are written -- e.g. the early exit that was not built.

usage: python tools/cost_model.py [rows=3000]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from pecanpy_b200 import synth  # noqa: E402

G = 32


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    cache = "/tmp/b2w_bench_cache/powerlaw_1000000_10000000_1_0.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        indptr, indices, data = z["indptr"], z["indices"], z["data"]
    else:
        indptr, indices, data = synth.power_law_csr(1_000_000, 10_000_000, 1, False)
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.savez(cache, indptr=indptr, indices=indices, data=data)
    n, L = indptr.size - 1, 80
    start = orc.shuffled_start(n, 1, 0)[:rows]
    W = orc.walk_csr("SparseOTF", indptr, indices, data, 4.0, 0.25, start, L, rng=orc.RNG_PHILOX, seed=1)
    ip = indptr.astype(np.int64)
    cls = {}
    steps = 0
    for w in W:
        for j in range(2, int(w[-1])):
            prev, cur, nxt = int(w[j - 2]), int(w[j - 1]), int(w[j])
            d, pd = int(ip[cur + 1] - ip[cur]), int(ip[prev + 1] - ip[prev])
            lgp, lgd = pd.bit_length(), d.bit_length()
            nwords = (d + 31) // 32
            fwd = ((d + G - 1) // G) * (lgp + 2)
            rev = ((pd + G) // G) * (lgd + 2) + (nwords + G - 1) // G
            if d <= 64 or fwd <= rev:
                chunks, depth = (d + G - 1) // G, lgp          # k + 1 probes per search = bit_length
                name = "fwd d<=32" if d <= 32 else ("fwd d<=64" if d <= 64 else "fwd d>64")
            else:
                chunks, depth = (pd + 1 + G - 1) // G, lgd
                name = "rev"
            # early exit: chunks actually needed to reach the chosen position
            crow = indices[ip[cur]:ip[cur + 1]]
            k = int(np.searchsorted(crow, nxt))
            if name == "rev":
                prow = indices[ip[prev]:ip[prev + 1]]
                need = min(int(np.searchsorted(prow, crow[min(k, d - 1)], side="right")) // G + 1, chunks)
            else:
                need = min(k // G + 1, chunks)
            a = cls.setdefault(name, np.zeros(8))
            a += [1, chunks * depth, depth * chunks,                   # steps, probes, chain with N = 1
                  depth * ((chunks + 1) // 2), depth * ((chunks + 3) // 4), need * depth, chunks, depth]
            steps += 1
    tot = sum(v for v in cls.values())
    print(f"config #3 sample: {rows} walkers, {steps} second-order steps, G = {G}")
    print(f"{'class':10s} {'% steps':>8s} {'probes/step':>12s} {'chain N=1':>10s} {'chain N=2':>10s} {'chain N=4':>10s} "
          f"{'probes w/ early exit':>21s} {'chunks':>7s} {'depth':>6s}")
    for name, v in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:10s} {100 * v[0] / steps:8.1f} {v[1] / steps:12.2f} {v[2] / steps:10.2f} {v[3] / steps:10.2f} "
              f"{v[4] / steps:10.2f} {v[5] / steps:21.2f} {v[6] / v[0]:7.2f} {v[7] / v[0]:6.2f}")
    print(f"{'all':10s} {100.0:8.1f} {tot[1] / steps:12.2f} {tot[2] / steps:10.2f} {tot[3] / steps:10.2f} "
          f"{tot[4] / steps:10.2f} {tot[5] / steps:21.2f}")
    print("probes/step: warp-level probe iterations (5 instructions each in the unrolled search); chain: dependent loads a "
          "warp waits for per step when N chunks are interleaved; columns are averages over ALL steps (contribution of "
          "the class), chunks / depth are per step of the class")


if __name__ == "__main__":
    main()
