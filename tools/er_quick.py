"""Quick ER (BASELINE config #2) timing of the walk kernel alone: 5 timed passes, L2 flushed between them."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pecanpy_b200 import synth
from pecanpy_b200.engine import WalkEngine

indptr, indices, data = synth.erdos_renyi_csr(100_000, 1_000_000, seed=0)
start = synth.shuffled_start(100_000, 10, 0)
eng = WalkEngine.from_csr(indptr, indices, data)
d_start = torch.from_numpy(start.view(np.int32)).cuda()
out = torch.empty((start.size, 82), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(3):
    eng.walk("SparseOTF", 0.5, 2.0, d_start, 80, seed=100 + i, out=out, collect_stats=False)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
for i, (a, b) in enumerate(ev):
    flush.fill_(i)
    a.record()
    eng.walk("SparseOTF", 0.5, 2.0, d_start, 80, seed=i, out=out, collect_stats=False)
    b.record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
steps = eng.count_steps(out, 80)
print("ER kernel ms", [round(x, 2) for x in ms], "G steps/s", round(steps / (sum(ms) / len(ms) * 1e-3) / 1e9, 3))
