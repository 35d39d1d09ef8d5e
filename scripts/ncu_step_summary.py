"""Summarise an `ncu --set full` capture of one whole training step (raw-page CSV) into a markdown table and a JSON of
per-kernel DRAM traffic that bench.py reads for the `roofline.traffic` field.
Usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_step_summary.py raw.csv profiles/r1_step"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
CLASS = [("edge_block_forward", "edge_forward"), ("edge_block_backward", "edge_backward"), ("transpose_blocks", "csr_scatter"),
         ("assemble_records", "assemble_records"), ("dp_allreduce_adam", "adam"), ("check_init_hist", "csr_check"), ("digit_scan", "csr_scan"), ("radix_scatter", "csr_scatter"),
         ("finalize_layout", "csr_finalize"), ("tc_embed_forward", "embed_forward_chain"),
         ("tc_conv_forward", "conv_forward_chain"), ("head_loss", "head2"),
         ("edge_forward", "edge_forward"), ("head2", "head2"), ("edge_backward", "edge_backward"),
         ("reduce_partials", "reduce_partials"), ("mse_seed", "mse_seed"), ("adam", "adam"), ("pack_weights", "pack_weights"),
         ("tc_conv_backward", "conv_backward_chain"), ("tc_embed_backward", "embed_backward_chain")]


def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def to_bytes(r, k):
    u = units[col[k]]
    return val(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


# The capture may hold more than one step and kernels of other libraries (torch's fill of the L2-flush buffer): keep
# this library's kernels and cut out ONE step, from the first CSR-build kernel that follows an Adam update to the next
# Adam update.
def short(r):
    return r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("gcnn::", "")


ours = [r for r in rows[2:] if any(key in short(r) for key, _ in CLASS)]
starts = [i for i, r in enumerate(ours) if "check_init_hist" in short(r) and (i == 0 or "adam" in short(ours[i - 1]))]
window = ours
for s0 in starts:
    ends = [i for i in range(s0, len(ours)) if "adam" in short(ours[i])]
    if ends and (s0 > 0 or len(starts) == 1):
        window = ours[s0:ends[0] + 1]
        break

out, per_class = [], {}
for r in window:
    name = short(r)
    cls = next((c for key, c in CLASS if key in name), "other")
    dur = val(r, "gpu__time_duration.sum")
    dur_us = dur * {"usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(units[col["gpu__time_duration.sum"]], 1)
    dram = to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    out.append({"kernel": name, "class": cls, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]], "us": dur_us,
                "dram_mb": dram / 1e6, "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
                "tensor_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "regs": val(r, "launch__registers_per_thread")})
    c = per_class.setdefault(cls, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    c["launches"] += 1
    c["us"] += dur_us
    c["dram_bytes"] += dram
total_us = sum(o["us"] for o in out)
with open(sys.argv[2] + "_kernels.md", "w") as f:
    f.write("# One training step under `ncu --set full` (cold caches, serialised; compare shares, not absolutes)\n\n")
    f.write("Workload: bench.py default (32 setcov graphs per step, per-sample counts -> block edge kernels and per-block transposes).  DRAM = dram__bytes_read.sum + dram__bytes_write.sum.\n\n")
    f.write("| # | kernel | grid | block | us | share | DRAM MB | DRAM % | L2 hit % | tensor % | warps active % | issue active % | regs |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for i, o in enumerate(out):
        f.write(f"| {i} | {o['kernel']} | {o['grid']} | {o['block']} | {o['us']:.1f} | {o['us'] / total_us:.3f} | {o['dram_mb']:.2f} | "
                f"{o['dram_pct']:.1f} | {o['l2_hit_pct']:.1f} | {o['tensor_pct']:.1f} | {o['warps_active_pct']:.1f} | "
                f"{o['issue_active_pct']:.1f} | {o['regs']:.0f} |\n")
    f.write(f"\nTotal {total_us:.1f} us over {len(out)} launches.\n\n## Per kernel class (bench.py classes)\n\n")
    f.write("| class | launches | us | share | DRAM MB per step | DRAM MB per launch |\n|---|---|---|---|---|---|\n")
    for cls, c in sorted(per_class.items(), key=lambda kv: -kv[1]["us"]):
        f.write(f"| {cls} | {c['launches']} | {c['us']:.1f} | {c['us'] / total_us:.3f} | {c['dram_bytes'] / 1e6:.2f} | "
                f"{c['dram_bytes'] / 1e6 / c['launches']:.2f} |\n")
json.dump({cls: {"launches_per_step": c["launches"], "dram_bytes_per_launch": c["dram_bytes"] / c["launches"],
                 "us_per_step_under_ncu": c["us"]} for cls, c in per_class.items()},
          open(sys.argv[2] + "_traffic.json", "w"), indent=1)
print(open(sys.argv[2] + "_kernels.md").read())
