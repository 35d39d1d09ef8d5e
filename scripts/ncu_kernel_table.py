"""Per-kernel table (times, throughputs, pipe utilisation, shared-memory traffic, stall reasons per issue) from the raw page
of an `ncu --set full` report:  ncu -i rep.ncu-rep --page raw --csv > raw.csv; python scripts/ncu_kernel_table.py raw.csv out.md "title" """
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}


def val(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


def scaled(r, k, table):
    if k not in col:
        return float("nan")
    unit = units[col[k]].split("/")[0]
    return val(r, k) * table.get(unit, 1)


US = {"usecond": 1, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "barrier", "branch_resolving", "dispatch_stall", "no_instruction", "sleeping", "membar"]

out = []
title = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full, per kernel"
out.append(f"# {title}\n")
out.append("`ncu --set full --clock-control none` (cold caches, one kernel at a time: compare shares and ratios, not absolute "
           "times).  Stall columns are `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`: warps waiting for that "
           "reason per issued instruction.\n")
out.append("| # | kernel | grid | block | regs | smem KB | us | DRAM MB | DRAM % | L2 hit % | issue % | warps active % | "
           "FMA pipe % | ALU pipe % | LSU pipe % | smem wavefronts (M) | bank conflicts (M) | inst (M) |")
out.append("|" + "---|" * 18)
stall_rows = []
for i, r in enumerate(rows[2:]):
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("gcnn::", "")
    us = scaled(r, "gpu__time_duration.sum", US)
    dram = scaled(r, "dram__bytes_read.sum", BYTES) + scaled(r, "dram__bytes_write.sum", BYTES)
    out.append(
        f"| {i} | {name} | {r[col['Grid Size']]} | {r[col['Block Size']]} | {val(r, 'launch__registers_per_thread'):.0f} | "
        f"{scaled(r, 'launch__shared_mem_per_block_dynamic', BYTES) / 1024:.0f} | {us:.1f} | {dram / 1e6:.2f} | "
        f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {val(r, 'lts__t_sector_hit_rate.pct'):.1f} | "
        f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{val(r, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{val(r, 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{val(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f} | "
        f"{val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:.2f} | "
        f"{val(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:.2f} | {val(r, 'smsp__inst_executed.sum') / 1e6:.2f} |")
    stall_rows.append((i, name, [val(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio") for s in STALLS]))
out.append("\n## Stall reasons (warps stalled per issued instruction)\n")
out.append("| # | kernel | " + " | ".join(s.replace("_", " ") for s in STALLS) + " |")
out.append("|" + "---|" * (2 + len(STALLS)))
for i, name, st in stall_rows:
    out.append(f"| {i} | {name} | " + " | ".join(f"{x:.2f}" for x in st) + " |")
text = "\n".join(out) + "\n"
open(sys.argv[2], "w").write(text)
print(text)
