#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME.txt [steps_per_launch]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:70s} {d[k]:>18s} {u.get(k, '')}")
        stalls = [(float(d[h]), h) for h in hdr if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct") and d[h]]
        for v, h in sorted(stalls, reverse=True)[:8]:
            lines.append(f"  stall {h.split('issue_stalled_')[1].replace('_per_warp_active.pct', ''):40s} {v:8.2f} % of active warps")
        # newer ncu: average number of warps per scheduler in each stall state, per issued instruction
        ratios = [(float(d[h]), h) for h in hdr if "average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and d[h]]
        for v, h in sorted(ratios, reverse=True)[:8]:
            lines.append(f"  stall {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):40s} {v:8.2f} warps per issue")
        try:
            t = float(d["gpu__time_duration.sum"])
            tu = u["gpu__time_duration.sum"]
            t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
            rb, wb = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            tb = rb * sc[u["dram__bytes_read.sum"]] + wb * sc[u["dram__bytes_write.sum"]]
            lines.append(f"  derived: dram traffic {tb / 1e9:.3f} GB per launch, {tb / t_s / 1e9:.1f} GB/s under the profiler")
            inst = float(d["smsp__inst_executed.sum"])
            if steps:
                lines.append(f"  derived: {inst / steps:.1f} warp-instructions per walk step, {tb / steps:.1f} DRAM bytes per walk step "
                             f"({steps:.0f} steps per launch)")
        except Exception as e:  # pragma: no cover
            lines.append(f"  (derived metrics unavailable: {e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
