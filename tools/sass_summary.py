#!/usr/bin/env python
"""SASS mnemonic summary per kernel of libmlsp_b200.so -> profiles/sass_summary_<tag>.txt
   usage: python tools/sass_summary.py <tag>      (cuobjdump -sass; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mlsp_b200", "lib", "libmlsp_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "SYNCS", "UCGABAR", "REDUX", "REDG", "RED", "ATOMG", "ATOMS", "LDG", "STG", "LDS", "STS",
         "FFMA", "FMNMX", "SHFL", "VOTE", "BAR", "LDL", "STL", "UTCATOMSWS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        total[fn] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        total[fn] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("RED", "BAR") and op == w):
                counts[fn][w] += 1
                break
path = os.path.join(ROOT, "profiles", f"sass_summary_{tag}.txt")
with open(path, "w") as fh:
    fh.write(f"# SASS mnemonic summary per kernel of mlsp_b200/lib/libmlsp_b200.so (sm_100a), {tag}\n")
    fh.write("# produced by: python tools/sass_summary.py (cuobjdump -sass, mnemonics counted per function)\n")
    fh.write("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA -> UTMALDG, mbarrier -> SYNCS,\n")
    fh.write("# redux.sync -> REDUX, red.global.add.v4.f32 -> RED / ATOMG\n\n")
    for f in sorted(counts):
        fh.write(f"{f}\n    instructions={total[f]}  " + "  ".join(f"{k}={v}" for k, v in counts[f].items()) + "\n")
print(path)
