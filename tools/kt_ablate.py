#!/usr/bin/env python
"""Ablation of the tcgen05 filter kernel alone (mlsp_graph_feature_fwd_stage, stage 2) under MLSP_KT_MODE:
   0 = shipped, 1 = pass-1 epilogue math off, 2 = pass-2 epilogue math off, 3 = both off (TMA + MMA + TMEM hand-off
   skeleton), 4 = three-term pass 1.  Also times the prep (stage 1) and ranking + gather (stage 4) kernels.
   usage: python tools/kt_ablate.py [--reps 20] [--profile]      (--profile: one launch of each filter shape between
   cudaProfilerStart/Stop, for `ncu --profile-from-start off`)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
shapes = [("A", 32, 1024, 20, 64), ("A", 32, 1024, 20, 128), ("X/4", 64, 4096, 40, 64), ("X/4", 64, 4096, 40, 128)]


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


handles = []
for name, B, N, k, C in shapes:
    x = (synth.features(B, C, N, 5) if name.startswith("X") else synth.smooth_features(B, C, N, 1244 + C)).to(dev)
    h = M.ops.GraphFeatureStages(x, k)
    handles.append(h)
    if args.profile:
        continue
    flops = 2.0 * B * N * N * C
    line = [f"{name} B={B} N={N} k={k} C={C}: prep {timed(lambda: h.run(1), args.reps):8.1f} us"]
    for cl in ("0", "1"):                                      # CTA pairs sharing the candidate blocks by TMA multicast, or not
        os.environ["MLSP_KT_CLUSTER"] = cl
        for mode in ("0", "2", "3", "4") if cl == "0" else ("0", "3"):
            os.environ["MLSP_KT_MODE"] = mode
            us = timed(lambda: h.run(2), args.reps)
            line.append(f"filter[pair {cl} mode {mode}] {us:8.1f} us ({flops / us * 1e-6:6.1f} TF/s alg)")
    os.environ["MLSP_KT_MODE"] = "0"
    os.environ["MLSP_KT_CLUSTER"] = "0"
    h.run(2)                                                   # valid lists again
    line.append(f"rank+gather {timed(lambda: h.run(4), args.reps):8.1f} us")
    print("  ".join(line), flush=True)
if args.profile:
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for h in handles:
        h.run(2)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
