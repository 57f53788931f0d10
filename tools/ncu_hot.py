#!/usr/bin/env python
"""Hot SASS lines of one kernel from an .ncu-rep (ncu --set full --import-source on).
   usage: python tools/ncu_hot.py <rep> <kernel-regex> [launch-skip] [top] [--range lo hi]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(h):
        break
    body.append(r)
si = h.index("# Samples"); src = h.index("Source"); ex = h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") or n.startswith("Stall")]
# stall columns are unnamed in some versions: find via names containing 'stall'
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[si] or 0) for r in body)
print(rows[0][1][:100], "total samples", tot, "instructions", len(body))
if "--all" in sys.argv:
    for n, r in enumerate(body):
        s = int(r[si] or 0)
        print(f"{n:5d} {s:6d} {r[ex]:>8s}  {r[src].strip()[:110]}")
    sys.exit()
order = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
for i in sorted(order):
    r = body[i]
    s = int(r[si] or 0)
    why = sorted(((int(r[c] or 0), h[c]) for c in stall_cols if (r[c] or "0").isdigit() and int(r[c] or 0) > 0), reverse=True)[:3]
    print(f"{i:5d} {s:6d} {100*s/max(tot,1):5.1f}% ex={r[ex]:>8s}  {r[src].strip()[:80]:80s} {' '.join(f'{n.replace('stall_','')}={v}' for v,n in why)}")
