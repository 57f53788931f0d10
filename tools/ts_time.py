#!/usr/bin/env python
"""Times target_structure (one launch) against knn + estimate_normals(idx) + cal_density (three launches), config A and S."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import synth
dev = torch.device("cuda:0")
def span(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for name, near, radius, pg, sh in (("A", 20, 0.13, 2, 0), ("S", 10, 0.115, 5, 10)):
    B, N, k = synth.CONFIGS[name]
    x = synth.surface_clouds(B, N, 1).to(dev)
    pts = x.permute(0, 2, 1).contiguous()
    def sep():
        idx = M.knn(x, near); M.estimate_normals(pts, near, idx=idx); M.cal_density(pts, radius, 16, pg, sh)
    print(name, "fused %.1f us   separate %.1f us   knn alone %.1f us" % (span(lambda: M.target_structure(pts, near, radius, 16, pg, sh)), span(sep), span(lambda: M.knn(x, near))))
