#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*launch_list_summary*.txt)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4])
    tot[name] += float(r[14].replace(",", "")); cnt[name] += 1
total = sum(tot.values())
print("%-44s %8s %14s %7s" % ("kernel", "launches", "total_ns", "share"))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-44s %8d %14d %6.1f%%" % (k[:44], cnt[k], tot[k], 100 * tot[k] / total))
print("%-44s %8d %14d" % ("total", sum(cnt.values()), total))
