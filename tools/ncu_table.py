#!/usr/bin/env python
"""ncu `--page raw --csv` output (ncu --set full over tools/prof_ops.py) -> the per-kernel table kept under profiles/.
   usage: python tools/ncu_table.py <raw.csv> <out.md> <title>"""
import csv
import sys

raw, out, title = sys.argv[1:4]
rows = list(csv.reader(open(raw, errors="ignore")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]


def col(r, name, scale=1.0, default=0.0):
    if name not in h:
        return default
    try:
        return float(r[h.index(name)].replace(",", "")) * scale
    except ValueError:
        return default


with open(out, "w") as fh:
    fh.write(f"# {title}\n\n| # | kernel | us | grid | regs | dram rd MB | dram wr MB | dram % | tensor % | issue % | warps % |\n"
             "|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    n = 0
    for r in rows[hi + 2:]:
        if len(r) != len(h):
            continue
        name = r[h.index("Kernel Name")].replace("mlsp::", "")
        fh.write(f"| {n} | `{name[:70]}` | {col(r, 'gpu__time_duration.sum', 1e-3):.1f} | {int(col(r, 'launch__grid_size'))} | "
                 f"{int(col(r, 'launch__registers_per_thread'))} | {col(r, 'dram__bytes_read.sum', 1e-6):.1f} | "
                 f"{col(r, 'dram__bytes_write.sum', 1e-6):.1f} | {col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                 f"{col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{col(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |\n")
        n += 1
print("wrote", out, n, "kernels")
