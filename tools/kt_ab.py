#!/usr/bin/env python
"""A/B timing of the feature-space kNN (tcgen05 path) under MLSP_KT_MODE switches, per workload shape.
   usage: python tools/kt_ab.py [--reps 30] [--workloads A,S,X] [--modes 0,4]
   mode 0 = shipped (pass 1 on the bf16 heads only), 4 = three-term pass 1 (the round-1e kernel)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--workloads", default="A,S,X")
ap.add_argument("--modes", default="0,4")
args = ap.parse_args()
dev = torch.device("cuda:0")
for w in args.workloads.split(","):
    B, N, k = synth.CONFIGS[w]
    for C in (64, 128):
        x = (synth.features(B, C, N, 5) if w == "X" else synth.smooth_features(B, C, N, 1244 + C)).to(dev)
        ref = None
        for mode in args.modes.split(","):
            os.environ["MLSP_KT_MODE"] = mode
            idx, stats = M.knn(x, k, return_stats=True)
            if ref is None:
                ref = idx
            same = bool(torch.equal(ref, idx))
            for _ in range(3):
                M.knn(x, k)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                M.knn(x, k)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / args.reps * 1e3
            fl = 2.0 * B * N * N * C
            print(f"{w} B={B} N={N} k={k} C={C} mode={mode}: {us:9.1f} us/call  {fl / us * 1e-6:7.1f} TFLOP/s algorithmic  "
                  f"fallback_rows={stats['fallback_rows']} same_idx={same}", flush=True)
        os.environ.pop("MLSP_KT_MODE", None)
        if w == "X" and C == 128:
            continue                                     # the 43 GB edge tensor is not part of this probe
        for fused in ("1", "0"):                         # get_graph_feature(idx=None): ranking + gather in one kernel, or two
            os.environ["MLSP_GGF_FUSED"] = fused
            xs = x[:64] if w == "X" else x
            for _ in range(3):
                out = M.get_graph_feature(xs, None, k=k)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.reps):
                out = M.get_graph_feature(xs, None, k=k)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / args.reps * 1e3
            gb = (4 * xs.numel() + 8 * xs.shape[0] * N * k + 8 * xs.numel() * k) / 1e9
            print(f"{w} B={xs.shape[0]} N={N} k={k} C={C} ggf fused={fused}: {us:9.1f} us/call  {gb / us * 1e6:7.0f} GB/s algorithmic (edge bytes)",
                  flush=True)
            del out
        os.environ.pop("MLSP_GGF_FUSED", None)
