"""Print the per-kernel-class table of bench.py JSON lines (scripts/show_bench.py FILE [FILE ...])."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as exc:
        print(f"== {path}: unreadable ({exc})")
        continue
    print(f"== {path}")
    tot = 0.0
    for k in d.get("kernels", []):
        print(f"{k['kernel']:22s} n={k['launches_per_step']:5.1f} ms={k['ms_per_step']:.4f} MB={k['algorithmic_mb_per_step']:8.2f} "
              f"GB/s={k['achieved_gbs']:7.1f} frac={k['frac']:.3f}")
        tot += k["ms_per_step"]
    print(f"sum of kernel classes {tot:.4f} ms; step {d['ms_per_step']:.4f} ms; {d['value']:.0f} {d['unit']}; "
          f"e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.4f} ms); launches {d['gpu_launches']}")
    if "e2e_records" in d:
        print(f"e2e_records {d['e2e_records']['value']:.0f} ({d['e2e_records']['ms_per_step']:.4f} ms)")
    print("roofline:", {k: v for k, v in d["roofline"].items() if k != "bytes"})
