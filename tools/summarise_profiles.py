#!/usr/bin/env python
"""Turn gpurun_out/ artefacts (ncu launch list CSV, ncu --set full report, bench JSON) into the small tracked
summaries under profiles/.   usage: python tools/summarise_profiles.py <round-tag> <launches.csv> <prof.ncu-rep> <bench.json>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep, bench = sys.argv[1:5]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -> per-kernel share table
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if v == v:                                   # ncu prints n/a (parsed as nan) for launches it could not time
        agg[r[ki]].append(v)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"launches_{tag}.md"), "w") as fh:
    fh.write(f"# ncu launch list, {tag}: `ncu --metrics gpu__time_duration.sum --clock-control none` over "
             "`python bench.py --steps 2 --warmup 3`\n\n")
    fh.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    fh.write("| share | launches | avg us | min us | max us | kernel |\n|---:|---:|---:|---:|---:|---|\n")
    for k, v in sorted(((k, v) for k, v in agg.items() if v), key=lambda kv: -sum(kv[1])):
        fh.write(f"| {sum(v)/tot*100:.1f}% | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {min(v)/1e3:.1f} | {max(v)/1e3:.1f} | `{k[:110]}` |\n")
    own = {k: v for k, v in agg.items() if v and ("mlsp::" in k or "gemm3" in k)}
    tot_own = sum(sum(v) for v in own.values())
    fh.write(f"\nThis library's kernels only ({tot_own / tot * 100:.1f}% of all kernel time in the command; the rest is torch: input "
             "generation with randn, fills, copies):\n\n| share | launches | avg us | kernel |\n|---:|---:|---:|---|\n")
    for k, v in sorted(own.items(), key=lambda kv: -sum(kv[1])):
        fh.write(f"| {sum(v)/tot_own*100:.1f}% | {len(v)} | {sum(v)/len(v)/1e3:.1f} | `{k[:110]}` |\n")

# ---- ncu --set full -> key metrics per captured kernel
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
    with open(os.path.join(out, f"ncu_{tag}.md"), "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, {tag} ({os.path.basename(rep)}; the .ncu-rep itself is scratch)\n\n")
        for r in rr[2:]:
            fh.write(f"## `{r[h.index('Kernel Name')][:120]}`\n\n")
            for w in want:
                if w in h:
                    fh.write(f"- {w}: {r[h.index(w)]} {units[h.index(w)]}\n")
            fh.write("\n")

# ---- bench line
if os.path.exists(bench):
    d = json.load(open(bench))
    json.dump(d, open(os.path.join(out, f"bench_{tag}.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(out)))
