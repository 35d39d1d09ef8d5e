"""Top stalled SASS instructions and per-source-line totals from `ncu --page source --csv` (arg: csv [instance])."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[hdr_idx[inst]]
end = hdr_idx[inst + 1] - 1 if inst + 1 < len(hdr_idx) else len(rows)
data = [r for r in rows[hdr_idx[inst] + 1:end] if len(r) == len(h)]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[si] or 0) for r in data)
print(rows[hdr_idx[inst] - 1][:2], "instances", len(hdr_idx), "samples", tot, "instrs", len(data),
      "warp-instrs executed", sum(int(r[ie] or 0) for r in data))
agg = collections.Counter()
for r in data:
    for i, c in stall_cols:
        agg[c] += int(r[i] or 0)
print("stall totals:", [(c, n) for c, n in agg.most_common(8)])
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:40]:
    st = sorted(((int(r[i] or 0), c) for i, c in stall_cols), reverse=True)[:2]
    print(f"{int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% exec={r[ie]:>7s} {r[src][:80]:80s} {st}")
