#!/usr/bin/env python
"""Tensor-path kNN tuning: MLSP_KT_STAGES sweep with parity against the exact kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import synth
dev = torch.device("cuda:0")
for (B, C, N, k) in ((32, 64, 1024, 20), (32, 128, 1024, 20), (16, 64, 2048, 20)):
    x = synth.smooth_features(B, C, N, 1244).to(dev)
    ref = M.knn(x, k, flags=1)
    for st in sys.argv[1:] or ["2", "3", "4", "6"]:
        os.environ["MLSP_KT_STAGES"] = st
        try:
            idx, stats = M.knn(x, k, return_stats=True)
        except Exception as e:
            print(B, C, N, "stages", st, "FAILED", str(e)[:100]); continue
        ok = torch.equal(idx, ref)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            M.knn(x, k)
        b.record(); torch.cuda.synchronize()
        print(f"B={B} C={C} N={N} stages={st}: {a.elapsed_time(b)/20*1e3:.1f} us/call parity={'ok' if ok else 'FAIL'} {stats}", flush=True)
