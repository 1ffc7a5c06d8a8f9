#!/usr/bin/env python3
"""Per-kernel roofline table from an ncu CSV taken with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,\\
smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,\\
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file X python tools/profile_g1.py 1
Aggregates the launches of each kernel: time, DRAM bytes, achieved GB/s against the measured HBM peak, and the time-weighted
issue-slot / multiplier-pipe / ALU-pipe occupancy.  (ncu serialises launches and flushes caches: use SHARES and ratios.)"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6550.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
per = defaultdict(dict)
names = {}
for r in rows:
    per[r[0]][r[12]] = (float(r[14].replace(",", "")), r[13])
    names[r[0]] = re.sub(r"\(.*", "", r[4]).replace("void ", "")
agg = defaultdict(lambda: defaultdict(float))
for lid, m in per.items():
    k = names[lid]
    def get(name, scale=1.0):
        v, unit = m.get(name, (0.0, ""))
        mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        return v * mult
    t = get("gpu__time_duration.sum")
    a = agg[k]
    a["n"] += 1; a["t"] += t
    a["bytes"] += get("dram__bytes_read.sum") + get("dram__bytes_write.sum")
    a["inst"] += m.get("smsp__inst_executed.sum", (0, ""))[0]
    a["issue_t"] += t * m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
    a["fmah_t"] += t * m.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
    a["alu_t"] += t * m.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
    a["lanes_i"] += m.get("smsp__thread_inst_executed_per_inst_executed.ratio", (0, ""))[0] * m.get("smsp__inst_executed.sum", (0, ""))[0]   # optional metric
total = sum(a["t"] for a in agg.values())
print("HBM peak %.1f GB/s (MEASURED_PEAKS.json); one G1 proof incl. trace generation; per kernel, all launches" % peak)
print("%-34s %6s %9s %6s %9s %8s %7s %7s %7s %7s %6s" % ("kernel", "n", "ms", "share", "DRAM MB", "GB/s", "HBM%", "issue%", "mulpipe%", "alu%", "lanes"))
for k in sorted(agg, key=lambda k: -agg[k]["t"]):
    a = agg[k]
    if a["t"] < total * 0.002:
        continue
    gbs = a["bytes"] / a["t"] / 1e9
    print("%-34s %6d %9.3f %5.1f%% %9.1f %8.1f %6.1f%% %6.1f%% %7.1f%% %6.1f%% %6.1f" % (k[:34], a["n"], a["t"] * 1e3, 100 * a["t"] / total, a["bytes"] / 1e6, gbs,
          100 * gbs / peak, a["issue_t"] / a["t"], a["fmah_t"] / a["t"], a["alu_t"] / a["t"], a["lanes_i"] / a["inst"] if a["inst"] else 0.0))   # lanes = active threads per warp instruction
print("%-34s %6d %9.3f" % ("total", sum(a["n"] for a in agg.values()), total * 1e3))
