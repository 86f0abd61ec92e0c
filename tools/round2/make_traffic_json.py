"""profiles/traffic.json from the CSVs of tools/round2/collect_traffic.sh (ncu, full bench launches of the final tree)."""
import csv, glob, json, os, sys
out = {}
for f in sorted(glob.glob("gpurun_out/traffic_*.csv")):
    wl = os.path.basename(f)[len("traffic_"):-4]
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and "walk_" in "".join(r)]
    if not rows:
        continue
    hdr = next(r for r in csv.reader(open(f)) if "Metric Name" in r)
    i_name, i_val, i_unit, i_k = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    m = {}
    for r in rows:
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}.get(u, 1)
        m[r[i_name]] = v * scale
        kern = r[i_k]
    import re
    name = re.search(r"walk_\w+", kern).group(0)
    if name == "walk_thread_kernel":
        name = "walk_thread_kernel<PRECOMP>"
    rec = {"kernel": name, "dram_bytes_per_launch": int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]),
           "dram_read_bytes": int(m["dram__bytes_read.sum"]), "dram_write_bytes": int(m["dram__bytes_write.sum"]),
           "ncu_kernel_ms": 1e3 * m["gpu__time_duration.sum"], "warp_instructions": int(m["smsp__inst_executed.sum"]),
           "l2_hit_pct": m.get("lts__t_sector_hit_rate.pct"), "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... -c 1 on `bench.py --workload %s --steps 1 --warmup 0` (round 2, final tree, cold first launch)" % wl}
    if wl.startswith("dense"):
        rec["note"] = "captured with --num-walks 1 (20 000 walkers): per-launch traffic of the 10-walk bench launch is 10x this"
        rec["num_walks"] = 1
    out[wl] = rec
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
