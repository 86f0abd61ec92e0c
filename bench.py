#!/usr/bin/env python
"""bench.py -- walk-steps/s of the B200 walk engine on the BASELINE.json workloads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one pass of the hot path (Base._random_walks, reference pecanpy.py:164-210) over the
whole job: num_walks x num_nodes walkers x walk_length steps.  Default workload = BASELINE config #3
(the configuration the north-star target is quoted on): synthetic power-law graph, 1M nodes / 10M
edges, SparseOTF p=4 q=0.25, 10 x 80.  With N > 1 (torchrun, one rank per GPU) the graph is replicated,
the shuffled start array is sharded into N contiguous row blocks, each rank walks its block into its
slice of the full matrix and ONE NCCL all-gather collects it ("scaling": "strong": the job is fixed).

Prints ONE JSON line (rank 0).  `value` = steps of the whole job / device time (max over ranks),
inputs resident in HBM.  `e2e` = the same through the host-buffer C-ABI call (b2w_walk_host): start
nodes in pinned host memory, walk matrix delivered to pinned host memory, copies inside the timed
region.  `roofline` = algorithmic HBM bytes of the walk kernel / its CUDA-event time,
against the measured copy bandwidth in MEASURED_PEAKS.json.  `cpu_baseline` = the C port of the
reference's algorithm (oracle/walk_oracle.c, OpenMP over walkers) on the host cores, bounded sample.

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
    # name: (mode, p, q, extend, gamma, weighted, generator, n, m, num_walks, L)
    "powerlaw-1M-10M-sparseotf": dict(mode="SparseOTF", p=4.0, q=0.25, extend=False, gen="powerlaw", n=1_000_000,
                                      m=10_000_000, weighted=False, num_walks=10, L=80, seed=1),
    "powerlaw-1M-10M-sparseotf-weighted": dict(mode="SparseOTF", p=4.0, q=0.25, extend=False, gen="powerlaw",
                                               n=1_000_000, m=10_000_000, weighted=True, num_walks=10, L=80, seed=1),
    "er-100k-1M-sparseotf": dict(mode="SparseOTF", p=0.5, q=2.0, extend=False, gen="er", n=100_000, m=1_000_000,
                                 weighted=False, num_walks=10, L=80, seed=0),
    "er-50k-1M-precomp": dict(mode="PreComp", p=0.25, q=4.0, extend=False, gen="er", n=50_000, m=1_000_000,
                              weighted=True, num_walks=10, L=80, seed=2),
    "dense-20k-denseotf-n2vplus": dict(mode="DenseOTF", p=0.5, q=2.0, extend=True, gen="dense", n=20_000, m=0,
                                       weighted=True, num_walks=10, L=80, seed=3, density=0.3),
}
DEFAULT_WORKLOAD = "powerlaw-1M-10M-sparseotf"
METRIC = "walk-steps/s"


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

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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
def algorithmic_bytes_sparse_gpu(torch, deg_t, walks_t, L: int) -> int:
    """SURVEY.md 8d, SparseOTF node2vec: step 1: 12 + 8 d_cur; step j>=2: 20 + 8 d_cur + 4 d_prev;
    + 4 per walker (start).  Evaluated exactly from the walk matrix, on the device, in row chunks."""
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
        total += int((8 * (d * cur_ok).sum() + 4 * (d * prev_ok).sum()).item())
        total += int((12 * (nsteps >= 1).sum() + 20 * torch.clamp(nsteps - 1, min=0).sum()).item())
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


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(wl, g, start, L, budget_s: float, seed: int):
    """Time the C port of the reference on a bounded sample; returns (steps/s, cores, sample description)."""
    from oracle import oracle as orc
    cores = host_threads()

    def run(rows):
        t0 = time.perf_counter()
        if g["kind"] == "dense":
            out = orc.walk_dense(g["data"], g["nonzero"], wl["p"], wl["q"], start[:rows], L, extend=wl["extend"],
                                 thr=g.get("thr"), rng=orc.RNG_PHILOX, seed=seed, nthreads=cores)
        else:
            out = orc.walk_csr(wl["mode"], g["indptr"], g["indices"], g["data"], wl["p"], wl["q"], start[:rows], L,
                               extend=wl["extend"], thr=g.get("thr"), alias=g.get("alias"), rng=orc.RNG_PHILOX, seed=seed,
                               nthreads=cores)
        dt = time.perf_counter() - t0
        return int((out[:, -1].astype(np.int64) - 1).sum()), dt

    # thread count: all hardware threads, or one per physical core if that is faster (SMT often hurts this
    # latency-bound gather loop); a short probe decides
    rows = min(start.size, 40000 if g["kind"] != "dense" else 512)
    run(rows)                                               # warm-up (thread pool, page faults, clocks)
    cores_all, half = cores, max(1, cores // 2)
    best = {}
    for nt in (cores_all, half, cores_all, half):
        cores = nt
        s_, dt_ = run(rows)
        best[nt] = max(best.get(nt, 0.0), s_ / dt_)
    cores = cores_all if best[cores_all] >= best[half] else half
    # grow the sample until it runs for at least ~40% of the budget (a tiny probe is a poor predictor),
    # capped by the budget and by the job size
    while True:
        s, dt = run(rows)
        if dt >= 0.4 * budget_s or rows >= start.size:
            break
        rate = s / max(dt, 1e-9)
        nxt = int(min(start.size, max(2 * rows, rate * budget_s / max(L, 1))))
        if nxt <= rows:
            break
        rows = nxt
    return s / dt, cores, f"first {rows} rows of the shuffled start array x {L} steps ({s} steps in {dt:.2f}s)", rows


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
    ap.add_argument("--flags", type=int, default=0, help="b2w_walk flags (debug)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.num_walks:
        wl["num_walks"] = args.num_walks
    L = wl["L"]
    K, W = args.steps, max(args.warmup, 0)

    # ---------------- reference arm: CPU port on rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        g = make_graph(wl, args.scale, 0, lambda: None)
        prep_reference_extras(wl, g)
        from pecanpy_b200 import synth
        start = synth.shuffled_start(g["n"], wl["num_walks"], 0)
        per_step_budget = max(2.0, 150.0 / max(K + W, 1))
        rate0, cores, _, rows = cpu_port_rate(wl, g, start, L, per_step_budget, seed=0)
        times, steps = [], 0
        from oracle import oracle as orc
        for it in range(W + K):
            t0 = time.perf_counter()
            if g["kind"] == "dense":
                out = orc.walk_dense(g["data"], g["nonzero"], wl["p"], wl["q"], start[:rows], L, extend=wl["extend"],
                                     thr=g.get("thr"), rng=orc.RNG_PHILOX, seed=it, nthreads=cores)
            else:
                out = orc.walk_csr(wl["mode"], g["indptr"], g["indices"], g["data"], wl["p"], wl["q"], start[:rows], L,
                                   extend=wl["extend"], thr=g.get("thr"), alias=g.get("alias"), rng=orc.RNG_PHILOX, seed=it,
                                   nthreads=cores)
            dt = time.perf_counter() - t0
            if it >= W:
                times.append(dt)
                steps += int((out[:, -1].astype(np.int64) - 1).sum())
        value = steps / sum(times)
        sample = f"each step = first {rows} rows of the shuffled start array x {L} steps"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * sum(times) / max(K, 1), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": dtype_of(wl), "data": "synthetic",
                "config": config_of(args, wl, g, world=1),
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
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    from pecanpy_b200 import _capi as capi
    from pecanpy_b200 import synth
    from pecanpy_b200.engine import WalkEngine

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
    R = (tot + world - 1) // world                     # rows per rank (last block padded)
    tot_pad = R * world
    lo, hi = rank * R, min(tot, (rank + 1) * R)
    my_rows = max(hi - lo, 0)
    ld = L + 2
    full = torch.zeros((tot_pad, ld), dtype=torch.int32, device=dev)
    mine = full[rank * R:(rank + 1) * R]
    d_start = torch.from_numpy(start[lo:hi].view(np.int32)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def one_pass(seed):
        if my_rows:
            eng.walk(wl["mode"], wl["p"], wl["q"], d_start, L, seed=seed, extend=wl["extend"], row0=lo,
                     out=mine, flags=args.flags, collect_stats=False)
        if world > 1:
            dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1))

    for it in range(W):
        one_pass(1000 + it)
    torch.cuda.synchronize()
    barrier()
    if getattr(eng, "edge_index_ms", None) is not None:
        # one-time graph preparation (like the upload of the CSR): per-edge records + common-neighbour lists
        extras["edge_index_build_ms"] = eng.edge_index_ms
        extras["edge_index_bytes"] = 16 * (int(g["indptr"][-1]) + 1) + 4 * int(getattr(eng, "edge_index_words", 0))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    barrier()
    for it in range(K):
        flush.fill_(it & 0xFF)                          # evict L2 between timed iterations (untimed)
        e0, e1, e2 = ev[it]
        e0.record()
        if my_rows:
            eng.walk(wl["mode"], wl["p"], wl["q"], d_start, L, seed=it, extend=wl["extend"], row0=lo, out=mine,
                     flags=args.flags, collect_stats=False)
        e1.record()
        if world > 1:
            dist.all_gather_into_tensor(full.view(-1), mine.reshape(-1))
        e2.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_total = sum(a.elapsed_time(c) for a, b, c in ev) * 1e-3
    t_kernel = sum(a.elapsed_time(b) for a, b, c in ev) * 1e-3
    tt = torch.tensor([t_total, t_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_kernel_max = float(tt[0]), float(tt[1])

    # kernel-side counters of one extra (untimed) pass: exact replays, reference-overflow choices
    walk_stats = None
    if my_rows and rank == 0:
        eng.walk(wl["mode"], wl["p"], wl["q"], d_start, L, seed=K - 1 if K else 0, extend=wl["extend"], row0=lo,
                 out=mine, flags=args.flags, collect_stats=True)
        walk_stats = eng.stats()

    # steps of the whole job, from the last timed pass (every rank holds the full matrix when world > 1)
    steps_job = eng.count_steps(full[:tot] if world > 1 else mine[:my_rows], L)
    steps_mine = eng.count_steps(mine[:my_rows], L) if my_rows else 0
    value = steps_job * K / t_total

    # roofline of the walk kernel on this rank (rank 0 reports)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    roofline = None
    if rank == 0 and my_rows:
        kernel_s = t_kernel / K                         # this rank's own kernel time per launch
        if g["kind"] == "csr":
            deg_t = torch.from_numpy((g["indptr"][1:].astype(np.int64) - g["indptr"][:-1].astype(np.int64))).to(dev)
            if wl["mode"] == "PreComp":
                alg = algorithmic_bytes_precomp(torch, deg_t, mine[:my_rows], L)
                formula = "PreComp: 32 + 4(ceil(log2 d_cur)+1) per step j>=2; 12 + 8 d_cur step 1; 4 per walker"
            else:
                alg = algorithmic_bytes_sparse_gpu(torch, deg_t, mine[:my_rows], L)
                formula = "SparseOTF: 20 + 8 d_cur + 4 d_prev per step j>=2; 12 + 8 d_cur step 1; 4 per walker"
        else:
            per = (21 if wl["extend"] else 10) * n + 4
            alg = per * steps_mine
            formula = f"DenseOTF: {21 if wl['extend'] else 10} N + 4 per step"
        achieved = alg / kernel_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": eng.kernel_name(wl["mode"], wl["p"], wl["q"], wl["extend"], args.flags), "kernel_ms": 1e3 * kernel_s,
                    "algorithmic_bytes_per_launch": alg, "bytes_per_step": alg / max(steps_mine, 1),
                    "formula": formula, "peak_source": peak_src}

    # ---------------- e2e through the host-buffer C-ABI entry point
    e2e = None
    if not args.no_e2e and my_rows >= 0:
        h_start = torch.from_numpy(start[lo:hi].view(np.int32).copy()).pin_memory()
        h_out = torch.empty((max(my_rows, 1), ld), dtype=torch.int32).pin_memory()
        np_start = h_start.numpy().view(np.uint32)
        np_out = h_out.numpy().view(np.uint32)[:my_rows]
        del full, mine
        torch.cuda.empty_cache()
        reps = max(2, min(K, 3))
        if my_rows:
            eng.walk_host(wl["mode"], wl["p"], wl["q"], np_start, L, seed=7, extend=wl["extend"], row0=lo, out=np_out,
                          flags=args.flags)
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        for r in range(reps):
            if my_rows:
                eng.walk_host(wl["mode"], wl["p"], wl["q"], np_start, L, seed=r, extend=wl["extend"], row0=lo,
                              out=np_out, flags=args.flags)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": steps_job * reps / float(td[0]), "unit": "steps/s",
               "h2d_bytes_per_step": int(4 * tot), "d2h_bytes_per_step": int(4 * ld * tot), "reps": reps,
               "api": "WalkEngine.walk_host -> b2w_walk_host (pinned host start[] in, pinned host walk matrix out)"}

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        prep_reference_extras(wl, g, eng)
        rate, cores, sample, _ = cpu_port_rate(wl, g, start, L, 15.0, seed=0)
        cpu = {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t_total / max(K, 1), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": dtype_of(wl), "data": "synthetic",
                "config": config_of(args, wl, g, world), "clocks": clocks, "gpu_launches": K * (1 if my_rows else 0),
                "steps_per_pass": steps_job, "kernel_ms_max_over_ranks": 1e3 * t_kernel_max / max(K, 1),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "walk_stats_rank0": walk_stats}
        line.update(extras)
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


def config_of(args, wl, g, world):
    return {"workload": args.workload, "mode": wl["mode"], "p": wl["p"], "q": wl["q"], "extend": wl["extend"],
            "num_nodes": g["n"], "nnz": int(g["indptr"][-1]) if g["kind"] == "csr" else None,
            "num_walks": wl["num_walks"], "walk_length": wl["L"], "rng": "philox4x32-10 keyed by (seed, global row, step)",
            "parallelism": f"graph replicated, walkers sharded over {world} GPU(s), one NCCL all-gather" if world > 1
            else "1 GPU", "l2": "256 MiB buffer written between timed iterations (L2 flush); graph+walk matrix exceed L2",
            "scale": args.scale}


def prep_reference_extras(wl, g, eng=None):
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
