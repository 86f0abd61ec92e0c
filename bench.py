#!/usr/bin/env python
"""bench.py -- walk-steps/s of the B200 walk engine on the BASELINE.json workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference] [--no-extra]

One "step" = one pass of the hot path (Base._random_walks, reference pecanpy.py:164-210) over the
whole job: num_walks x num_nodes walkers x walk_length steps.  Default workload = BASELINE config #3
(the configuration the north-star target is quoted on): synthetic power-law graph, 1M nodes / 10M
edges, SparseOTF p=4 q=0.25, 10 x 80.  With N > 1 (torchrun, one rank per GPU) the graph is replicated,
the shuffled start array is sharded over the ranks and every rank ends with the whole walk matrix
("scaling": "strong": the job is fixed).  How the matrix gets everywhere is --gather: `mirror` = the walk kernel
stores its rows into every peer's CUDA-IPC-mapped matrix itself, over NVLink, while it walks (b2w_walk_mirrored; no
gather phase), `nccl` = one all-gather after the walk (or per batch with --batches), `push` = copy engines; `auto`
(default) = mirror where the kernel supports it, chosen against NCCL by a warm-up probe beyond 4 GPUs.

Prints ONE JSON line (rank 0).  `value` = steps of the whole job / device time (max over ranks),
inputs resident in HBM.  `e2e` = the same through the host-buffer C-ABI call (b2w_walk_host; at N > 1
b2w_walk_multi from ONE process over the N GPUs): start nodes in pinned host memory, ONE walk matrix
delivered to pinned host memory, copies inside the timed region.  `roofline` = algorithmic HBM bytes of
the walk kernel / its CUDA-event time, against the measured copy bandwidth in
MEASURED_PEAKS.json; `roofline.dram_frac` = the kernel's real DRAM traffic (ncu, profiles/traffic.json)
on the same scale.  `cpu_baseline` = the C port of the reference's algorithm (oracle/walk_oracle.c, OpenMP
over walkers) on the host cores, bounded sample.  `extra` (N = 1, default workload): the same measurements
for the other BASELINE configurations (#2 ER, #4 PreComp, #5 DenseOTF node2vec+) and for the weighted /
node2vec+ SparseOTF kernels, each a short sub-benchmark with its own roofline, cpu_baseline, e2e and clocks.

--impl reference: the CPU arm on the same workload/metric (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "powerlaw-1M-10M-sparseotf": dict(mode="SparseOTF", p=4.0, q=0.25, extend=False, gen="powerlaw", n=1_000_000,
                                      m=10_000_000, weighted=False, num_walks=10, L=80, seed=1, config="#3"),
    "powerlaw-1M-10M-sparseotf-weighted": dict(mode="SparseOTF", p=4.0, q=0.25, extend=False, gen="powerlaw",
                                               n=1_000_000, m=10_000_000, weighted=True, num_walks=10, L=80, seed=1,
                                               config="#3 topology, random weights (generic kernel)"),
    "powerlaw-1M-10M-sparseotf-n2vplus": dict(mode="SparseOTF", p=4.0, q=0.25, extend=True, gamma=0.0, gen="powerlaw",
                                              n=1_000_000, m=10_000_000, weighted=True, num_walks=10, L=80, seed=1,
                                              config="#3 topology, random weights, node2vec+ (--extend)"),
    "er-100k-1M-sparseotf": dict(mode="SparseOTF", p=0.5, q=2.0, extend=False, gen="er", n=100_000, m=1_000_000,
                                 weighted=False, num_walks=10, L=80, seed=0, config="#2"),
    "er-50k-1M-precomp": dict(mode="PreComp", p=0.25, q=4.0, extend=False, gen="er", n=50_000, m=1_000_000,
                              weighted=True, num_walks=10, L=80, seed=2, config="#4"),
    "dense-20k-denseotf-n2vplus": dict(mode="DenseOTF", p=0.5, q=2.0, extend=True, gamma=0.0, gen="dense", n=20_000, m=0,
                                       weighted=True, num_walks=10, L=80, seed=3, density=0.3, config="#5"),
}
DEFAULT_WORKLOAD = "powerlaw-1M-10M-sparseotf"
# sub-benchmarks appended to the default N = 1 line: (workload, warm-up passes, minimum timed passes)
EXTRA = [("er-100k-1M-sparseotf", 3, 5), ("er-50k-1M-precomp", 3, 5), ("dense-20k-denseotf-n2vplus", 3, 2),
         ("powerlaw-1M-10M-sparseotf-weighted", 3, 2), ("powerlaw-1M-10M-sparseotf-n2vplus", 3, 2)]
METRIC = "walk-steps/s"
CHECK_SEED = 12345          # the pass whose matrix is checksummed: the same for every N, K, W
MIN_TIMED_S = 1.0           # sub-benchmarks: a timed region shorter than this is repeated


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ graphs
def make_graph(wl: dict, scale: float, rank: int, barrier):
    """Deterministic synthetic graph; rank 0 generates into /tmp, the other ranks load it."""
    from pecanpy_b200 import synth
    n = max(64, int(wl["n"] * scale))
    m = int(wl["m"] * scale)
    cache = os.path.join("/tmp", "b2w_bench_cache")
    os.makedirs(cache, exist_ok=True)
    tag = f"{wl['gen']}_{n}_{m}_{wl['seed']}_{int(wl['weighted'])}"
    path = os.path.join(cache, tag + ".npz")
    if wl["gen"] == "dense":
        # 3.2 GB: every rank generates its own copy (deterministic), no cache file
        data, nz = synth.dense_weighted(n, wl.get("density", 0.3), wl["seed"])
        return dict(kind="dense", n=n, data=data, nonzero=nz)
    if rank == 0 and not os.path.exists(path):
        t0 = time.time()
        if wl["gen"] == "powerlaw":
            indptr, indices, data = synth.power_law_csr(n, m, wl["seed"], wl["weighted"])
        else:
            indptr, indices, data = synth.erdos_renyi_csr(n, m, wl["seed"], wl["weighted"])
        tmp = path + ".tmp.npz"
        np.savez(tmp, indptr=indptr, indices=indices, data=data)
        os.replace(tmp, path)
        log(f"[bench] generated {tag} in {time.time() - t0:.1f}s")
    barrier()
    z = np.load(path)
    return dict(kind="csr", n=n, indptr=z["indptr"], indices=z["indices"], data=z["data"])


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 20):
        self.idx = gpu_index
        self.period = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period), "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs a moment to start: do not open the timed region before its first sample."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def samples_since(self, t0: float) -> int:
        return sum(1 for t, _ in self.lines if t >= t0)

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.lines:
            if not (t0 <= t <= t1 + 0.05):
                continue                                    # only samples taken during the timed region
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ bytes
def algorithmic_bytes_sparse_gpu(torch, deg_t, walks_t, L: int, extend: bool = False) -> int:
    """SURVEY.md 8d, SparseOTF node2vec: step 1: 12 + 8 d_cur; step j>=2: 20 + 8 d_cur + 4 d_prev;
    + 4 per walker (start).  node2vec+: + 4 d_prev (weights of row(prev)) + 4 (thr[cur]) per step j>=2 (the 4 c
    threshold gathers of the c common neighbours are not counted: a lower bound).  Evaluated exactly from the walk
    matrix, on the device, in row chunks."""
    total = 0
    rows = walks_t.shape[0]
    chunk = 1 << 19
    for r0 in range(0, rows, chunk):
        w = walks_t[r0:r0 + chunk]
        eff = w[:, L + 1].to(torch.int64)
        nsteps = eff - 1
        cols = torch.arange(L, device=w.device)[None, :]
        d = deg_t[w[:, :L].to(torch.int64)]
        cur_ok = cols < nsteps[:, None]                 # entry c is `cur` of step c + 1
        prev_ok = cols < (nsteps[:, None] - 1)          # entry c is `prev` of step c + 2
        total += int((8 * (d * cur_ok).sum() + (8 if extend else 4) * (d * prev_ok).sum()).item())
        total += int((12 * (nsteps >= 1).sum() + (24 if extend else 20) * torch.clamp(nsteps - 1, min=0).sum()).item())
        total += 4 * w.shape[0]
    return total


def algorithmic_bytes_precomp(torch, deg_t, walks_t, L: int) -> int:
    """SURVEY.md 8d, PreComp step j>=2: 32 + 4 (ceil(log2 d_cur) + 1); step 1 as SparseOTF first order."""
    total = 0
    rows = walks_t.shape[0]
    chunk = 1 << 19
    for r0 in range(0, rows, chunk):
        w = walks_t[r0:r0 + chunk]
        nsteps = w[:, L + 1].to(torch.int64) - 1
        cols = torch.arange(L, device=w.device)[None, :]
        d = deg_t[w[:, :L].to(torch.int64)]
        later = (cols >= 1) & (cols < nsteps[:, None])
        lg = torch.ceil(torch.log2(torch.clamp(d.to(torch.float64), min=1.0))).to(torch.int64)
        total += int(((32 + 4 * (lg + 1)) * later).sum().item())
        total += int(((12 + 8 * d[:, 0]) * (nsteps >= 1)).sum().item()) + 4 * w.shape[0]
    return total


def matrix_checksum(torch, walks_t) -> str:
    """Order-sensitive 64-bit checksum of a device walk matrix (wrap-around int64 arithmetic): equal for equal
    matrices whatever the number of GPUs that produced them."""
    acc = 0
    rows, ld = walks_t.shape
    chunk = 1 << 19
    for r0 in range(0, rows, chunk):
        w = walks_t[r0:r0 + chunk].to(torch.int64) & 0xFFFFFFFF
        idx = (torch.arange(r0, r0 + w.shape[0], device=w.device, dtype=torch.int64)[:, None] * ld +
               torch.arange(ld, device=w.device, dtype=torch.int64)[None, :])
        acc = (acc + int(((w + 1) * ((idx * 0x9E3779B1) % 0x7FFFFFFF + 1)).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return f"{acc:016x}"


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_walk(wl, g, start, L, seed, cores):
    from oracle import oracle as orc
    if g["kind"] == "dense":
        return orc.walk_dense(g["data"], g["nonzero"], wl["p"], wl["q"], start, L, extend=wl["extend"],
                              thr=g.get("thr"), rng=orc.RNG_PHILOX, seed=seed, nthreads=cores)
    return orc.walk_csr(wl["mode"], g["indptr"], g["indices"], g["data"], wl["p"], wl["q"], start, L,
                        extend=wl["extend"], thr=g.get("thr"), alias=g.get("alias"), rng=orc.RNG_PHILOX, seed=seed,
                        nthreads=cores)


def cpu_port_rate(wl, g, start, L, budget_s: float, seed: int):
    """Time the C port of the reference on a bounded sample; returns (steps/s, cores, sample description, rows)."""
    cores = host_threads()

    def run(rows, nt):
        t0 = time.perf_counter()
        out = cpu_walk(wl, g, start[:rows], L, seed, nt)
        dt = time.perf_counter() - t0
        return int((out[:, -1].astype(np.int64) - 1).sum()), dt

    # thread count: all hardware threads, or one per physical core if that is faster (SMT often hurts this
    # latency-bound gather loop); a short probe decides
    rows = min(start.size, 40000 if g["kind"] != "dense" else 512)
    run(rows, cores)                                        # warm-up (thread pool, page faults, clocks)
    cores_all, half = cores, max(1, cores // 2)
    best = {}
    for nt in (cores_all, half, cores_all, half):
        s_, dt_ = run(rows, nt)
        best[nt] = max(best.get(nt, 0.0), s_ / dt_)
    cores = cores_all if best[cores_all] >= best[half] else half
    # grow the sample until it runs for at least ~40% of the budget (a tiny probe is a poor predictor),
    # capped by the budget and by the job size
    while True:
        s, dt = run(rows, cores)
        if dt >= 0.4 * budget_s or rows >= start.size:
            break
        rate = s / max(dt, 1e-9)
        nxt = int(min(start.size, max(2 * rows, rate * budget_s / max(L, 1))))
        if nxt <= rows:
            break
        rows = nxt
    return s / dt, cores, f"first {rows} rows of the shuffled start array x {L} steps ({s} steps in {dt:.2f}s)", rows


# ------------------------------------------------------------------------------------------ one workload, our arm
def run_workload(name, args, K, W, rank, world, local_rank, *, do_e2e=True, do_cpu=True, cpu_budget=15.0,
                 min_timed_s=0.0):
    import torch
    import torch.distributed as dist
    from pecanpy_b200 import synth
    from pecanpy_b200.engine import WalkEngine

    dev = torch.device("cuda", local_rank)
    wl = dict(WORKLOADS[name])
    if args.num_walks:
        wl["num_walks"] = args.num_walks
    L = wl["L"]

    def barrier():
        if world > 1:
            dist.barrier()

    g = make_graph(wl, args.scale, rank, barrier)
    n = g["n"]
    if g["kind"] == "dense":
        eng = WalkEngine.from_dense(g["data"], g["nonzero"], device=dev)
    else:
        eng = WalkEngine.from_csr(g["indptr"], g["indices"], g["data"], device=dev)
    extras = {}
    if wl["extend"]:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.compute_thresholds(wl.get("gamma", 0.0))      # b2w_noise_thresholds (device), not the oracle
        torch.cuda.synchronize()
        extras["thresholds_ms"] = 1e3 * (time.perf_counter() - t0)
    if wl["mode"] == "PreComp":
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.build_alias(g["indptr"], wl["p"], wl["q"], extend=wl["extend"])
        torch.cuda.synchronize()
        extras["alias_build_ms"] = 1e3 * (time.perf_counter() - t0)
        extras["alias_entries"] = int(eng.alias[0][-1])

    start = synth.shuffled_start(n, wl["num_walks"], 0)
    tot = start.size
    ld = L + 2
    # Row blocks.  One GPU: the whole array.  N GPUs: NB batches; batch b of rank r = rows [(b N + r) B, +B), so that
    # the all-gather of batch b fills the contiguous rows [b N B, (b + 1) N B) while batch b + 1 is being walked.
    # (default: ONE batch = the single all-gather at the end; measured on 2 x B200, more batches only cost: every
    # extra launch of the lane-per-walker kernel adds ~1.4 ms, more than the overlap hides -- DESIGN.md 5)
    auto_nb = 1
    NB = 1 if world == 1 else max(1, min(args.batches or auto_nb, (tot + world * 65536 - 1) // (world * 65536)))
    B = (tot + world * NB - 1) // (world * NB)
    tot_pad = B * world * NB
    peer, gather = None, "none"
    if world > 1:
        # auto: the all-gather fused into the walk kernel (mirrored stores) where the kernel can do it -- measured
        # 2 GPUs 68-70 G vs 62 G (NCCL), 4 GPUs 113 G vs 88 G; at 8 GPUs (the coalesced variant has not been measured
        # there) a short probe in the warm-up decides between it and NCCL
        gather = "mirror" if args.gather == "auto" else args.gather
        if gather == "mirror" and (world > 8 or eng.prepare(wl["mode"], wl["p"], wl["q"], wl["extend"], args.flags) != "walk_uw_edge_kernel"):
            gather = "nccl"                                 # only the unweighted edge-index kernel mirrors its rows
        if gather in ("push", "mirror"):
            from pecanpy_b200.dist import PeerMatrix
            peer = PeerMatrix(tot_pad, ld, dev)             # collective; .ok is agreed between the ranks
            if not peer.ok:                                 # no CUDA IPC / peer access on this box: NCCL does it
                log(f"[bench] rank {rank}: peer mapping unavailable ({peer.error}); falling back to NCCL")
                peer.close()
                peer, gather = None, "nccl"
    full = peer.full if peer is not None else torch.zeros((tot_pad, ld), dtype=torch.int32, device=dev)
    start_pad = np.zeros(tot_pad, dtype=np.uint32)
    start_pad[:tot] = start
    d_start_all = torch.from_numpy(start_pad.view(np.int32)).to(dev)
    blocks = []                                             # (row0, rows actually walked) of this rank, per batch
    for b in range(NB):
        r0 = (b * world + rank) * B
        blocks.append((r0, max(0, min(B, tot - r0))))
    my_rows = sum(r for _, r in blocks)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB)
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    def one_pass(seed, events=None):
        """Walk this rank's blocks.  N > 1, by `gather`: "mirror" -- the walk kernel itself stores the rows into every
        peer's mapped matrix (nothing to do afterwards but the barrier in peer.finish); "push" -- each walked batch is
        copied into the peers' matrices by the copy engines while the next one is walked; "nccl" -- all-gathered by
        NCCL on a side stream."""
        cur = torch.cuda.current_stream(dev)
        for b, (r0, rows) in enumerate(blocks):
            if rows:
                eng.walk(wl["mode"], wl["p"], wl["q"], d_start_all[r0:r0 + rows], L, seed=seed, extend=wl["extend"],
                         row0=r0, out=full[r0:r0 + rows], flags=args.flags, collect_stats=False,
                         mirrors=peer.mirror_ptrs(r0) if gather == "mirror" else None)
            if gather == "mirror":
                pass                                        # the kernel stored the rows in every peer's matrix itself
            elif gather == "push":
                peer.push(r0, r0 + rows)
            elif world > 1:
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    seg = full[b * world * B:(b + 1) * world * B]
                    dist.all_gather_into_tensor(seg.view(-1), full[(b * world + rank) * B:(b * world + rank + 1) * B].view(-1))
        if events is not None:
            events[1].record(cur)                           # this rank's walk kernels are done
        if gather in ("mirror", "push"):
            peer.finish()
        elif world > 1:
            cur.wait_stream(comm)

    if world > 4 and args.gather == "auto" and gather == "mirror":
        def run_with(cand, seed):                           # (NCCL gathers into the same, IPC-shared, matrix)
            nonlocal gather
            gather = cand
            one_pass(seed)

        def timed_ms(fn):
            torch.cuda.synchronize()
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            fn()
            p1.record()
            torch.cuda.synchronize()
            pt = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)       # the same number, hence the same choice, on every rank
            return float(pt[0])

        gather, probe = choose_gather(["mirror", "nccl"], run_with, timed_ms)
        extras["gather_probe_ms"] = probe
        log(f"[bench] rank {rank}: gather probe {probe} -> {gather}")
    for it in range(W):
        one_pass(1000 + it)
    torch.cuda.synchronize()
    barrier()
    if getattr(eng, "edge_index_ms", None) is not None:
        # one-time graph preparation (like the upload of the CSR): per-edge records + common-neighbour lists
        extras["edge_index_build_ms"] = eng.edge_index_ms
        extras["edge_index_bytes"] = 16 * (int(g["indptr"][-1]) + 1) + 4 * int(getattr(eng, "edge_index_words", 0))
    if getattr(eng, "edge_ckpt_ms", None) is not None:
        extras["edge_ckpt_build_ms"] = eng.edge_ckpt_ms
        extras["edge_ckpt_bytes"] = 4 * eng.edge_ckpt_floats
    if getattr(eng, "windex_ms", None) is not None:
        extras["weighted_index_build_ms"] = eng.windex_ms
        extras["weighted_index_bytes"] = eng.windex_bytes
        extras["weighted_index_counts"] = eng.windex_counts
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    torch.cuda.synchronize()
    barrier()
    # timed region: EXACTLY K passes for the headline (min_timed_s = 0; clocks are sampled every 20 ms).  The `extra`
    # sub-benchmarks pass min_timed_s > 0: their K is a minimum and a region shorter than that (a 1 ms PreComp pass)
    # is extended so that the sampler sees the GPU under load -- the number of passes actually timed is reported
    ev = []
    t_wall0 = time.perf_counter()
    it = 0
    while True:
        flush.fill_(it & 0xFF)                              # evict L2 between timed iterations (untimed)
        e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
             torch.cuda.Event(enable_timing=True))
        e[0].record()
        one_pass(it, e)
        e[2].record()
        ev.append(e)
        it += 1
        if it >= K:
            if it % 4 == 0 or it == K:                      # decide together (the loop must end on every rank at once)
                torch.cuda.synchronize()
                done = torch.tensor([1 if (time.perf_counter() - t_wall0 >= min_timed_s or it >= 400 * max(K, 1)) else 0],
                                    device=dev)
                if world > 1:
                    dist.all_reduce(done, op=dist.ReduceOp.MIN)
                if int(done.item()):
                    break
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    K_eff = len(ev)
    t_total = sum(a.elapsed_time(c) for a, b, c in ev) * 1e-3
    t_kernel = sum(a.elapsed_time(b) for a, b, c in ev) * 1e-3
    tt = torch.tensor([t_total, t_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_kernel_max = float(tt[0]), float(tt[1])

    # one extra (untimed) pass with a fixed seed: checksum of the whole matrix (identical for every N), kernel-side
    # counters of this rank's share
    one_pass(CHECK_SEED)
    torch.cuda.synchronize()
    steps_job = eng.count_steps(full[:tot], L) if world > 1 else eng.count_steps(full[:tot], L)
    checksum = matrix_checksum(torch, full[:tot]) if rank == 0 else None
    walk_stats = None
    if rank == 0 and blocks[0][1]:
        r0, rows = blocks[0]
        eng.walk(wl["mode"], wl["p"], wl["q"], d_start_all[r0:r0 + rows], L, seed=CHECK_SEED, extend=wl["extend"], row0=r0,
                 out=full[r0:r0 + rows], flags=args.flags, collect_stats=True)
        walk_stats = eng.stats()
        walk_stats["rows"] = rows
    steps_mine = sum(eng.count_steps(full[r0:r0 + rows], L) for r0, rows in blocks if rows)
    value = steps_job * K_eff / t_total

    # roofline of the walk kernel on this rank (rank 0 reports)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    roofline = None
    kname = eng.kernel_name(wl["mode"], wl["p"], wl["q"], wl["extend"], args.flags)
    if rank == 0 and my_rows:
        kernel_s = t_kernel / K_eff                     # this rank's own kernel time per pass
        mine = torch.cat([full[r0:r0 + rows] for r0, rows in blocks if rows]) if len(blocks) > 1 else full[:my_rows]
        if g["kind"] == "csr":
            deg_t = torch.from_numpy((g["indptr"][1:].astype(np.int64) - g["indptr"][:-1].astype(np.int64))).to(dev)
            if wl["mode"] == "PreComp":
                alg = algorithmic_bytes_precomp(torch, deg_t, mine, L)
                formula = "PreComp: 32 + 4(ceil(log2 d_cur)+1) per step j>=2; 12 + 8 d_cur step 1; 4 per walker"
            else:
                alg = algorithmic_bytes_sparse_gpu(torch, deg_t, mine, L, wl["extend"])
                formula = ("SparseOTF: 20 + 8 d_cur + 4 d_prev per step j>=2; 12 + 8 d_cur step 1; 4 per walker" +
                           ("; node2vec+: + 4 d_prev + 4 per step j>=2 (thr gathers of common neighbours not counted)"
                            if wl["extend"] else ""))
        else:
            per = (21 if wl["extend"] else 10) * n + 4
            alg = per * steps_mine
            formula = f"DenseOTF: {21 if wl['extend'] else 10} N + 4 per step"
        del mine
        achieved = alg / kernel_s / 1e9
        traffic, note = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            rec = json.load(open(tpath)).get(name, {})
            if rec.get("kernel", kname) == kname:
                traffic = rec.get("dram_bytes_per_launch")
                if traffic and rec.get("num_walks"):        # captured with fewer walks per node than this launch
                    traffic = int(traffic * wl["num_walks"] / rec["num_walks"])
                note = rec.get("note")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": kname, "kernel_ms": 1e3 * kernel_s,
                    "algorithmic_bytes_per_launch": alg, "bytes_per_step": alg / max(steps_mine, 1),
                    "formula": formula, "peak_source": peak_src}
        if traffic and world == 1:
            # the kernel's REAL DRAM traffic on the same scale: what the HBM system actually delivers
            roofline["dram_gbs"] = traffic / kernel_s / 1e9
            roofline["dram_frac"] = traffic / kernel_s / 1e9 / peak
            roofline["dram_bytes_per_step"] = traffic / max(steps_mine, 1)
        if kname in ("walk_uw_edge_kernel", "walk_uw_kernel"):
            roofline["note"] = ("`frac` is the contract's EFFECTIVE bandwidth: SURVEY 8d's row-streaming bytes / time. "
                                "This kernel never streams rows (per-edge index: one 16-byte record per step), so frac "
                                "can exceed 1; its own ceiling is the random-sector DRAM rate, see dram_frac")
        elif note:
            roofline["note"] = note

    # ---------------- e2e through the host-buffer C-ABI entry points: ONE host matrix
    e2e = None
    del full
    if peer is not None:                                    # unmap the peers' matrices, then free the own one
        peer.close()
        peer = None
    torch.cuda.empty_cache()
    if do_e2e:
        e2e = run_e2e(torch, dist, eng, g, wl, start, L, K, rank, world, local_rank, steps_job, args)

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and do_cpu:
        if wl["extend"] and eng.thr is not None:
            # input of the CPU port: the thresholds (bit-identical to the reference's NumPy loop, tests/) from the
            # device instead of 13 us per node of Python
            g["thr"] = eng.thr.cpu().numpy()
        prep_reference_extras(wl, g)
        rate, cores, sample, _ = cpu_port_rate(wl, g, start, L, cpu_budget, seed=0)
        cpu = {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}

    line = None
    if rank == 0:
        launches = K_eff * sum(1 for _, r in blocks if r)
        line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_total / max(K_eff, 1), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": dtype_of(wl), "data": "synthetic",
                "config": config_of(args, name, wl, g, world, NB, {"push": "copy engines over NVLink into CUDA-IPC mapped peer matrices", "mirror": "fused: the walk kernel stores every row sector into the CUDA-IPC mapped matrices of all peers over NVLink", "nccl": "NCCL", "none": ""}[gather]), "clocks": clocks, "gpu_launches": launches,
                "timed_passes": K_eff, "steps_per_pass": steps_job,
                "kernel_ms_max_over_ranks": 1e3 * t_kernel_max / max(K_eff, 1),
                "checksum": {"seed": CHECK_SEED, "fnv_like_u64": checksum},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "walk_stats_rank0": walk_stats}
        line.update(extras)
    eng.close()
    del eng, d_start_all, flush
    torch.cuda.empty_cache()
    return line


def choose_gather(candidates, run_pass, timed_ms):
    """Warm-up probe between gather mechanisms (N > 4, --gather auto): ``run_pass(cand, seed)`` runs one whole pass
    with mechanism ``cand`` (collective: every rank calls it in the same order), ``timed_ms(fn)`` returns the device
    time of ``fn()`` in ms, max over ranks (so every rank sees the same numbers and chooses the same).  One untimed
    pass, then two timed ones per candidate; the first candidate wins ties.  Returns (choice, {cand: ms per pass})."""
    probe = {}
    for cand in candidates:
        run_pass(cand, 900)
        probe[cand] = timed_ms(lambda c=cand: (run_pass(c, 901), run_pass(c, 902))) / 2
    best = min(candidates, key=lambda c: (probe[c], candidates.index(c)))
    return best, probe


def on_rank0_while_others_sleep(dist, rank, work, timeout_s=1800.0):
    """Rank 0 runs ``work()``; the other ranks SLEEP ON THE CPU until it is done (a flag file on this node -- the job
    is one node by contract), then everybody meets in a barrier.  Waiting inside the barrier instead would leave an NCCL
    kernel spinning on every other GPU, and rank 0's own kernels on those GPUs (b2w_walk_multi drives all of them from
    one process) would have to time-slice against it."""
    import tempfile
    import uuid
    tok = [os.path.join(tempfile.gettempdir(), "b2w_bench_" + uuid.uuid4().hex) if rank == 0 else None]
    dist.broadcast_object_list(tok, src=0)
    err = None
    if rank == 0:
        try:
            work()
        except BaseException as exc:                        # let the others go before failing
            err = exc
        open(tok[0], "w").close()
    else:
        t_end = time.time() + timeout_s
        while not os.path.exists(tok[0]) and time.time() < t_end:
            time.sleep(0.005)
    dist.barrier()
    if rank == 0:
        try:
            os.remove(tok[0])
        except OSError:
            pass
    if err is not None:
        raise err


def run_e2e(torch, dist, eng, g, wl, start, L, K, rank, world, local_rank, steps_job, args):
    """The call a user makes, host buffers in and out: b2w_walk_host on one GPU; with N > 1, b2w_walk_multi from
    rank 0's process over all N GPUs of the box (one host thread per GPU, one pinned host matrix) while the other
    ranks wait -- the single-process multi-GPU path of the drop-in classes (pecanpy_b200/multi.py)."""
    from pecanpy_b200.engine import WalkEngine
    tot, ld = start.size, L + 2
    reps = max(2, min(K, 3))
    dt = None
    api = "WalkEngine.walk_host -> b2w_walk_host (pinned host start[] in, pinned host walk matrix out)"
    def rank0_work():
        nonlocal reps, dt, api
        h_start = torch.from_numpy(start.view(np.int32).copy()).pin_memory()
        h_out = torch.empty((tot, ld), dtype=torch.int32).pin_memory()
        np_start = h_start.numpy().view(np.uint32)
        np_out = h_out.numpy().view(np.uint32)
        if world == 1:
            def call(seed):
                eng.walk_host(wl["mode"], wl["p"], wl["q"], np_start, L, seed=seed, extend=wl["extend"], out=np_out,
                              flags=args.flags)
        else:
            from pecanpy_b200.multi import walk_host_engines
            engines = [eng]
            for d in range(world):
                if d == local_rank:
                    continue
                dv = torch.device("cuda", d)
                e2 = (WalkEngine.from_dense(g["data"], g["nonzero"], device=dv) if g["kind"] == "dense"
                      else WalkEngine.from_csr(g["indptr"], g["indices"], g["data"], device=dv))
                if wl["extend"]:
                    e2.compute_thresholds(wl.get("gamma", 0.0))
                if wl["mode"] == "PreComp":
                    e2.build_alias(g["indptr"], wl["p"], wl["q"], extend=wl["extend"])
                engines.append(e2)
            api = (f"walk_host_engines -> b2w_walk_multi: one process, {world} GPUs, one host thread per GPU, pinned host "
                   "start[] in, ONE pinned host walk matrix out")

            def call(seed):
                walk_host_engines(engines, wl["mode"], wl["p"], wl["q"], wl["extend"], np_start, L, seed, out=np_out,
                                  flags=args.flags)
        t0 = time.perf_counter()
        call(7)                                             # warm-up: staging buffers, edge index of the replicas
        if time.perf_counter() - t0 < 0.5:
            call(8)
        else:
            reps = 2                                        # long passes: keep the default run within minutes
        for d in range(world if world > 1 else 1):
            torch.cuda.synchronize(d if world > 1 else local_rank)
        t0 = time.perf_counter()
        for r in range(reps):
            call(r)
        dt = time.perf_counter() - t0
        if world > 1:
            for e2 in engines[1:]:
                e2.close()
        del h_out, h_start

    if world == 1:
        rank0_work()
    else:
        on_rank0_while_others_sleep(dist, rank, rank0_work)
    if rank != 0:
        return None
    return {"value": steps_job * reps / dt, "unit": "steps/s", "h2d_bytes_per_step": int(4 * tot),
            "d2h_bytes_per_step": int(4 * ld * tot), "reps": reps, "api": api}


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debug only; invalidates the number)")
    ap.add_argument("--num-walks", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the sub-benchmarks of the other BASELINE configs")
    ap.add_argument("--batches", type=int, default=0, help="N > 1: all-gather batches overlapped with the walk (0 = auto)")
    ap.add_argument("--gather", default="auto", choices=["auto", "push", "nccl", "mirror"],
                    help="N > 1: how every rank gets the whole matrix -- mirror: the walk kernel stores its rows into the peers' matrices itself (CUDA IPC over NVLink); push: copy engines; nccl: all-gather; auto: mirror where the kernel supports it (probe against nccl beyond 4 GPUs)")
    ap.add_argument("--flags", type=int, default=0, help="b2w_walk flags (debug)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(args.steps, 1), max(args.warmup, 0)

    # ---------------- reference arm: CPU port on rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = dict(WORKLOADS[args.workload])
        if args.num_walks:
            wl["num_walks"] = args.num_walks
        L = wl["L"]
        g = make_graph(wl, args.scale, 0, lambda: None)
        prep_reference_extras(wl, g)
        from pecanpy_b200 import synth
        start = synth.shuffled_start(g["n"], wl["num_walks"], 0)
        per_step_budget = max(2.0, 150.0 / max(K + W, 1))
        rate0, cores, _, rows = cpu_port_rate(wl, g, start, L, per_step_budget, seed=0)
        times, steps = [], 0
        for it in range(W + K):
            t0 = time.perf_counter()
            out = cpu_walk(wl, g, start[:rows], L, it, cores)
            dt = time.perf_counter() - t0
            if it >= W:
                times.append(dt)
                steps += int((out[:, -1].astype(np.int64) - 1).sum())
        value = steps / sum(times)
        sample = f"each step = first {rows} rows of the shuffled start array x {L} steps"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * sum(times) / max(K, 1), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": dtype_of(wl), "data": "synthetic",
                "config": config_of(args, args.workload, wl, g, 1, 1),
                "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------- our arm
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    line = run_workload(args.workload, args, K, W, rank, world, local_rank, do_e2e=not args.no_e2e,
                        do_cpu=not args.no_cpu)
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_extra and not args.num_walks \
            and args.scale == 1.0 and not args.flags:
        extra = {}
        for name, w_, k_ in EXTRA:
            t0 = time.time()
            try:
                sub = run_workload(name, args, k_, w_, rank, world, local_rank, do_e2e=not args.no_e2e,
                                   do_cpu=not args.no_cpu, cpu_budget=4.0, min_timed_s=MIN_TIMED_S)
                sub["steps"] = sub["timed_passes"]
                for key in ("higher_is_better", "scaling", "vs_baseline", "data", "n_gpus", "metric", "unit"):
                    sub.pop(key, None)
                sub["bench_wall_s"] = round(time.time() - t0, 1)
                extra[name] = sub
            except Exception as exc:                          # a sub-benchmark must not take the headline down
                extra[name] = {"error": repr(exc)}
            log(f"[bench] extra {name}: {time.time() - t0:.1f}s")
        line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def host_threads() -> int:
    """Host threads the CPU arm may use (torchrun exports OMP_NUM_THREADS=1; ask the OS instead)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def dtype_of(wl):
    return "f64" if wl["mode"] == "DenseOTF" else "f32"


def config_of(args, name, wl, g, world, nb, gather="none"):
    return {"workload": name, "baseline_config": wl.get("config"), "mode": wl["mode"], "p": wl["p"], "q": wl["q"],
            "extend": wl["extend"], "weighted": wl["weighted"],
            "num_nodes": g["n"], "nnz": int(g["indptr"][-1]) if g["kind"] == "csr" else None,
            "num_walks": wl["num_walks"], "walk_length": wl["L"], "rng": "philox4x32-10 keyed by (seed, global row, step)",
            "parallelism": (f"graph replicated, walkers sharded over {world} GPU(s), all-gather in {nb} batch(es) "
                            f"overlapped with the walk of the next batch ({gather})") if world > 1 else "1 GPU",
            "l2": "256 MiB buffer written between timed iterations (L2 flush); graph+walk matrix exceed L2",
            "scale": args.scale}


def prep_reference_extras(wl, g):
    """Host-side inputs the CPU arm needs: node2vec+ thresholds, PreComp tables (built by the port itself)."""
    from oracle import oracle as orc
    if wl["extend"] and "thr" not in g:
        if g["kind"] == "dense":
            g["thr"] = orc.noise_thresholds_dense(g["data"], g["nonzero"], wl.get("gamma", 0.0))
        else:
            g["thr"] = orc.noise_thresholds_csr(g["indptr"], g["data"], wl.get("gamma", 0.0))
    if wl["mode"] == "PreComp" and "alias" not in g:
        g["alias"] = orc.alias_build(g["indptr"], g["indices"], g["data"], wl["p"], wl["q"], wl["extend"], g.get("thr"))


if __name__ == "__main__":
    sys.exit(main())
