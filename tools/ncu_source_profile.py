#!/usr/bin/env python
"""Per-source-line dynamic instruction profile of one kernel from an .ncu-rep (read on the build box).

usage: python tools/ncu_source_profile.py REPORT.ncu-rep STEPS_PER_LAUNCH [TOP_N] > profiles/NAME_src.txt

Uses `ncu -i REPORT --page source --csv --print-source cuda,sass` (needs -lineinfo at build time and
--import-source on at capture time) and prints warp-instructions per walk step and the share of stall samples for the
hottest source lines -- the view that found the 13-instruction binary-search probe in round 1."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur, hdr, kernel = None, None, ""
    inst, samp, text = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            kernel = r[1]
            continue
        if len(r) > 2 and r[0] == "Line No":
            hdr = r
            i_inst, i_samp = r.index("Instructions Executed"), r.index("# Samples")
            continue
        if hdr is None or len(r) <= i_inst or not r[0]:
            continue
        try:
            key = (cur, int(r[0]))
            v, s = float(r[i_inst]), float(r[i_samp])
        except ValueError:
            continue
        inst[key] += v
        samp[key] += s
        text.setdefault(key, r[1].strip()[:100])
    tot, ts = sum(inst.values()), max(sum(samp.values()), 1.0)
    print(f"kernel: {kernel}")
    print(f"warp-instructions executed: {tot:.0f} = {tot / steps:.1f} per walk step ({steps:.0f} steps); stall samples {ts:.0f}")
    per_file = collections.Counter()
    for (f, _), v in inst.items():
        per_file[f] += v
    print("per file: " + ", ".join(f"{f} {v / steps:.1f}" for f, v in per_file.most_common()))
    print(f"{'file':24s} {'line':>4s} {'inst/step':>9s} {'% inst':>6s} {'% stall':>7s}  source")
    for key, v in inst.most_common(top):
        print(f"{key[0]:24s} {key[1]:4d} {v / steps:9.2f} {100 * v / tot:6.1f} {100 * samp[key] / ts:7.1f}  {text[key]}")


if __name__ == "__main__":
    main()
