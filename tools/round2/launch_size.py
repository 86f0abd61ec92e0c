"""How does the lane-per-walker kernel's time depend on the number of rows per launch?  (round 2: batching the
all-gather made the walk slower: 8 launches of 625k rows took 24 ms where one launch of 5M rows took 12.8 ms.)"""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pecanpy_b200 import synth
from pecanpy_b200.engine import WalkEngine
indptr, indices, data = synth.power_law_csr(1_000_000, 10_000_000, 1, False)
eng = WalkEngine.from_csr(indptr, indices, data, device="cuda:0")
start = synth.shuffled_start(1_000_000, 10, 0)
d_start = torch.from_numpy(start.view(np.int32)).cuda()
out = torch.empty((start.size, 82), dtype=torch.int32, device="cuda")
eng.walk("SparseOTF", 4.0, 0.25, d_start, 80, seed=1, out=out, collect_stats=False)
torch.cuda.synchronize()
for rows in [10_000_000, 5_000_000, 2_500_000, 1_250_000, 625_000, 312_500, 156_250]:
    n = start.size // rows
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in range(n):
            eng.walk("SparseOTF", 4.0, 0.25, d_start[b * rows:(b + 1) * rows], 80, seed=2, row0=b * rows,
                     out=out[b * rows:(b + 1) * rows], collect_stats=False)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # the same launches timed one by one
    per = []
    for b in range(min(n, 4)):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        eng.walk("SparseOTF", 4.0, 0.25, d_start[b * rows:(b + 1) * rows], 80, seed=2, row0=b * rows,
                 out=out[b * rows:(b + 1) * rows], collect_stats=False)
        a1.record(); torch.cuda.synchronize()
        per.append(round(a0.elapsed_time(a1), 3))
    print(f"rows/launch {rows:>9}  launches {n:>3}  total {ms:8.3f} ms  per-launch {per}", flush=True)
