#!/usr/bin/env python
"""FPS variant sweep (MLSP_FPS_VARIANT="P,W" tuning hook of mlsp_fps) with a parity check against the oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import synth
import oracle
dev = torch.device("cuda:0")
for (B, N, npoint) in ((32, 1024, 512), (16, 2048, 512)):
    x = synth.surface_clouds(B, N, 7)
    xd = x.to(dev)
    start = (torch.arange(B) * 13) % N
    sd = start.to(dev)
    rc, rv = oracle.fps(x.numpy(), npoint, start.numpy())
    for var in (None, "4,8", "8,4", "16,2", "32,1", "2,16", "8,8", "16,4", "4,16"):
        if var is None:
            os.environ.pop("MLSP_FPS_VARIANT", None)
        else:
            P, W = map(int, var.split(","))
            if 32 * P * W < N:
                continue
            os.environ["MLSP_FPS_VARIANT"] = var
        cen, vals = M.fps_from_start(xd, npoint, sd)
        ok = np.array_equal(cen.cpu().numpy(), rc) and np.array_equal(vals.cpu().numpy(), rv)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            M.fps_from_start(xd, npoint, sd)
        b.record()
        torch.cuda.synchronize()
        print(f"B={B} N={N} npoint={npoint} variant={var}: {a.elapsed_time(b)/20*1e3:.1f} us  parity={'ok' if ok else 'FAIL'}", flush=True)
