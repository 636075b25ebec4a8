#!/usr/bin/env python
"""Aggregates an ncu launch list (gpu__time_duration.sum CSV) over the last `n` launches (= one incremental frame)."""
import collections
import csv
import re
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 153
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
last = [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("void <unnamed>::", ""), float(r["Metric Value"].replace(",", "")))
        for r in rows][-n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for name, v in last:
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in last)
print(f"last {n} launches: {tot / 1e3:.1f} us")
for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / tot * 100:5.1f}%  {v / 1e3:9.1f} us  x{c:3d}  avg {v / c / 1e3:8.1f} us  {name[:100]}")
