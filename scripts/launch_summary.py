"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (scripts/launch_summary.py FILE [--all])."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.defaultdict(list)
order = []
for r in rows[hi + 2:]:
    if len(r) > vi:
        name = r[ki].split("(")[0].replace("void ", "").replace("gcnn::", "")
        t = float(r[vi].replace(",", "")) / 1e3
        agg[name].append(t)
        order.append((name, r[gi], t))
total = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:48s} n={len(v):4d} sum={sum(v):9.1f} us share={sum(v) / total:6.3f} avg={sum(v) / len(v):7.2f} min={min(v):7.2f} max={max(v):7.2f}")
if "--all" in sys.argv:
    for name, grid, t in order:
        print(f"{name:48s} grid={grid:16s} {t:8.2f} us")
