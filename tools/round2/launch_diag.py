"""Where do the ~1.6 ms per launch of the lane-per-walker kernel go?  Host time of the call vs GPU time of the kernel,
and the kernel alone for several launch sizes (run under `ncu --metrics gpu__time_duration.sum` for true durations)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pecanpy_b200 import synth
from pecanpy_b200.engine import WalkEngine
gen = sys.argv[1] if len(sys.argv) > 1 else "powerlaw"
if gen == "powerlaw":
    indptr, indices, data = synth.power_law_csr(1_000_000, 10_000_000, 1, False); p, q = 4.0, 0.25
else:
    indptr, indices, data = synth.erdos_renyi_csr(1_000_000, 10_000_000, 1, False); p, q = 4.0, 0.25
eng = WalkEngine.from_csr(indptr, indices, data, device="cuda:0")
n = indptr.size - 1
start = synth.shuffled_start(n, 2, 0)
d_start = torch.from_numpy(start.view(np.int32)).cuda()
out = torch.empty((start.size, 82), dtype=torch.int32, device="cuda")
eng.walk("SparseOTF", p, q, d_start, 80, seed=1, out=out, collect_stats=False)
torch.cuda.synchronize()
for rows in [2_000_000, 500_000, 156_250, 20_000, 2_000]:
    nl = min(8, start.size // rows)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in range(nl):
        eng.walk("SparseOTF", p, q, d_start[b * rows:(b + 1) * rows], 80, seed=2, row0=b * rows,
                 out=out[b * rows:(b + 1) * rows], collect_stats=False)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{gen} rows/launch {rows:>8} launches {nl}: host {1e3 * t_host / nl:7.3f} ms/call, gpu {e0.elapsed_time(e1) / nl:7.3f} ms/launch", flush=True)
# which walkers are slow?  time single-row-block launches of the hub-start rows vs ordinary rows
deg = (indptr[1:] - indptr[:-1]).astype(np.int64)
hubs = np.argsort(-deg)[:2000].astype(np.uint32)
low = np.flatnonzero(deg <= 20)[:2000].astype(np.uint32)
for name, arr in (("hub starts", hubs), ("low-degree starts", low)):
    ds = torch.from_numpy(arr.view(np.int32)).cuda()
    o = torch.empty((arr.size, 82), dtype=torch.int32, device="cuda")
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.walk("SparseOTF", p, q, ds, 80, seed=3, out=o, collect_stats=True)
        e1.record(); torch.cuda.synchronize()
    print(f"{gen} {name}: 2000 rows {e0.elapsed_time(e1):7.3f} ms  stats {eng.stats()}", flush=True)
