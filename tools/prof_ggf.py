#!/usr/bin/env python
"""Minimal driver for `ncu --profile-from-start off`: one get_graph_feature(idx=None) call per feature width of
workload A (knn_prep + knn_tensor + knn_refine with the fused edge gather), plus one bare knn call.
   ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof python tools/prof_ggf.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402

B, N, k = synth.CONFIGS["A"]
dev = torch.device("cuda:0")
clouds = synth.surface_clouds(B, N, 1234).to(dev)
f64 = synth.smooth_features(B, 64, N, 1244).to(dev)
f128 = synth.smooth_features(B, 128, N, 1246).to(dev)
for f in (clouds, f64, f128):
    M.get_graph_feature(f, None, k=k)
    M.knn(f, k)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in (clouds, f64, f128):
    out = M.get_graph_feature(f, None, k=k)
M.knn(f64, k)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
