#!/usr/bin/env python3
"""Stall-sample breakdown of one kernel from `ncu -i rep --page source --csv --kernel-name K`:
segments between barriers / branches and the most-sampled SASS instructions."""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
first = True
data = []
for r in rows:
    if r and r[0].startswith("Kernel Name"):
        if data:
            break
    if len(r) > 10 and r[0].startswith("0x"):
        data.append(r)
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
totE = sum(int(r[iE]) for r in data)
print(len(data), "SASS instructions; samples", tot, "warp-instructions executed", totE)
seg, cur = [], {"s": 0, "e": 0, "n": 0, "start": 0}
for k, r in enumerate(data):
    src = r[iSrc].strip()
    cur["s"] += int(r[iS]); cur["e"] += int(r[iE]); cur["n"] += 1
    w = src.split()
    op = w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "")
    if op.startswith("BAR") or op.startswith("BRA") or op.startswith("EXIT"):
        cur["end"] = k; cur["mark"] = src; seg.append(cur); cur = {"s": 0, "e": 0, "n": 0, "start": k + 1}
seg.append(cur)
for s in seg:
    if s["s"] > tot * 0.01:
        print("  sass %5d-%-5s n=%4d  samples %5.1f%%  executed %5.1f%%  ends: %s" % (s["start"], s.get("end"), s["n"], 100 * s["s"] / tot, 100 * s["e"] / totE, s.get("mark", "")[:60]))
print("top instructions by samples")
for k, r in sorted(enumerate(data), key=lambda x: -int(x[1][iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print("  %5d  %6s  %8s  %s" % (k, r[iS], r[iE], r[iSrc].strip()[:100]))
