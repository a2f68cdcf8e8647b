#!/usr/bin/env python
"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list:  python scripts/launch_table.py file.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi or not r[0].isdigit():
        continue
    t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
    n = re.sub(r"\(.*", "", r[ki])
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(v[1] for v in agg.values())
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}% x{c:5d} {n[:120]}")
print(f"{tot:9.3f} ms total, {sum(v[0] for v in agg.values())} launches")
