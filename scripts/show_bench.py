"""Print the per-kernel-class table of a bench.py JSON line (scripts/show_bench.py FILE)."""
import json
import sys

d = json.load(open(sys.argv[1]))
tot = 0.0
for k in d["kernels"]:
    print(f"{k['kernel']:22s} n={k['launches_per_step']:5.1f} ms={k['ms_per_step']:.4f} MB={k['algorithmic_mb_per_step']:8.2f} "
          f"GB/s={k['achieved_gbs']:7.1f} frac={k['frac']:.3f}")
    tot += k["ms_per_step"]
print(f"sum of kernel classes {tot:.4f} ms; step {d['ms_per_step']:.4f} ms; {d['value']:.0f} {d['unit']}; "
      f"e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.4f} ms); launches {d['gpu_launches']}")
print("roofline:", {k: v for k, v in d["roofline"].items() if k != "bytes"})
