#!/usr/bin/env python
"""dram__bytes_read+write per launch of the modelled kernels, from an `ncu --set full` report of tools/opbench.py
(workload A) -> profiles/ncu_traffic.json, keyed like bench.py's op names ("A:edge_bwd_C64", ...).
   usage: python tools/traffic.py <tag> <report.ncu-rep> [more reports ...]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, reps = sys.argv[1], sys.argv[2:]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
B, N, k = 32, 1024, 20
rows = []                                               # (kernel name, dram bytes, microseconds)
for rep in reps:
    # a .ncu-rep, or the `ncu -i ... --page raw --csv` export of one (reports over 64 MiB do not travel back from the GPU box)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    kn, rd, wr, tm = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    ur, uw = rr[1][rd], rr[1][wr]
    for r in rr[2:]:
        rows.append((r[kn], float(r[rd].replace(",", "")) * scale[ur] + float(r[wr].replace(",", "")) * scale[uw],
                     float(r[tm].replace(",", ""))))


def alg(C):
    return 4 * B * C * N + 8 * B * N * k + 8 * B * C * N * k


groups = collections.defaultdict(list)
for name, tot, us in rows:
    if "edge_fwd_vec" in name or "edge_bwd_vec" in name:
        op = "edge_fwd" if "fwd" in name else "edge_bwd"
        C = 64 if us < 90 else 128                      # the two shapes of workload A differ 2x in time
        groups[f"A:{op}_C{C}"].append(tot)
    elif "knn_refine_kernel" in name:
        C = 128 if ", 128>" in name else 64
        if us > (80 if C == 128 else 50):               # with the fused edge gather (ggf*); the ranking alone is 30-40 us
            groups[f"A:k_rank_gather_C{C}"].append(tot)
    elif "knn_tensor_kernel" in name:
        groups[f"A:knn_C{64 if us < 40 else 128}"].append(tot)
        groups[f"A:k_filter_C{64 if us < 40 else 128}"].append(tot)
    elif "knn3_kernel" in name:
        groups["A:knn_C3"].append(tot)
    elif "edge_fwd3" in name:
        groups["A:edge_fwd_C3"].append(tot)
    elif "edge_bwd3_kernel" in name:
        groups["A:edge_bwd_C3"].append(tot)
out = {}
for key, v in sorted(groups.items()):
    out[key] = {"dram_bytes": sum(v) / len(v), "launches": len(v),
                "source": f"profiles/ncu_{tag}.md (ncu --set full --clock-control none over tools/prof_ops.py; main kernel of the op, "
                          "dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
    if (key.startswith("A:edge_") or key.startswith("A:k_rank_gather_")) and not key.endswith("C3"):
        out[key]["algorithmic_bytes"] = alg(int(key.split("_C")[1]))
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
