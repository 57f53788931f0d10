#!/usr/bin/env python
"""Phase timeline of the tcgen05 filter kernel (mlsp_knn_tensor_timeline): per-phase mean/max over the CTAs and the wave structure.
   usage: python tools/kt_timeline.py [B N k]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsp_b200 import _lib, synth  # noqa: E402
from mlsp_b200.ops import _ptr, _stream, _workspace  # noqa: E402

B, N, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else synth.CONFIGS["A"]
dev = torch.device("cuda:0")
names = ["prologue", "first acc", "pass 1", "sort", "xchg", "merge+thr", "pass 2", "lists done", "copy out"]
cluster = int(os.environ.get("KT_CLUSTER", "0"))
for C in (64, 128):
    x = synth.smooth_features(B, C, N, 1244 + C).to(dev)
    idx = torch.empty((B, N, k), dtype=torch.int64, device=dev)
    ws = _workspace(_lib.OP_KNN, B, C, N, k, dev)
    nct = B * ((N + 127) // 128)
    ts = torch.zeros((nct, 16), dtype=torch.int64, device=dev)
    for _ in range(3):
        _lib.call("mlsp_knn_tensor_timeline", _ptr(x), B, C, N, k, _ptr(idx), _ptr(ws), ws.numel(), _ptr(ts), cluster, _stream(dev))
    torch.cuda.synchronize()
    t = ts.cpu().numpy()
    t0 = t[:, 0].min()
    d = np.diff(t[:, :10], axis=1)
    print(f"C={C} cluster={cluster}: {nct} CTAs, kernel span {(t[:, 9].max() - t0) / 1e3:.1f} us, CTA lifetime mean {(t[:, 9] - t[:, 0]).mean() / 1e3:.1f} us")
    for i, n in enumerate(names):
        print(f"   {n:12s} mean {d[:, i].mean() / 1e3:7.2f} us   max {d[:, i].max() / 1e3:7.2f} us")
    start = np.sort(t[:, 0] - t0) / 1e3
    print("   CTA start times (us): first wave <", f"{start[min(147, nct - 1)]:.1f}", " last start", f"{start[-1]:.1f}")
