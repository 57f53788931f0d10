#!/usr/bin/env python
"""Per-op device timing (CUDA events around R back-to-back calls) and host enqueue cost, for kernel iteration.
   usage: python tools/opbench.py [--workload A|S] [--reps 50] [--ops knn3,fps,...]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="A")
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--ops", default="")
args = ap.parse_args()
B, N, k = synth.CONFIGS[args.workload]
dev = torch.device("cuda:0")
want = set(args.ops.split(",")) if args.ops else None

clouds = synth.surface_clouds(B, N, 1234).to(dev)
pts = clouds.permute(0, 2, 1).contiguous()
f64 = synth.smooth_features(B, 64, N, 1244).to(dev)
f128 = synth.smooth_features(B, 128, N, 1246).to(dev)
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, reps=args.reps, flush_l2=False):
    if want and name not in want:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if flush_l2:
        ev = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        dev_us = float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3
        host_us = float("nan")
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        host_us = (time.perf_counter() - t0) / reps * 1e6
        torch.cuda.synchronize()
        dev_us = a.elapsed_time(b) / reps * 1e3
    print(f"{name:24s} dev {dev_us:9.1f} us/call   host-enqueue {host_us:8.1f} us/call", flush=True)


idx3 = M.knn(clouds, k)
idx64 = M.knn(f64, k)
idx128 = M.knn(f128, k)
timeit("knn3", lambda: M.knn(clouds, k))
timeit("knn64", lambda: M.knn(f64, k))
timeit("knn128", lambda: M.knn(f128, k))
timeit("knn64_exact", lambda: M.knn(f64, k, flags=1), reps=10)
timeit("ggf3", lambda: M.get_graph_feature(clouds, None, k=k), flush_l2=True)       # knn + gather in one call (the DGCNN call)
timeit("ggf64", lambda: M.get_graph_feature(f64, None, k=k), flush_l2=True)
timeit("ggf128", lambda: M.get_graph_feature(f128, None, k=k), flush_l2=True)
for C, f, idx in ((3, clouds, idx3), (64, f64, idx64), (128, f128, idx128)):
    g = torch.randn(B, N, k, 2 * C, device=dev).permute(0, 3, 1, 2)
    fr = f.detach().requires_grad_(True)
    out = M.get_graph_feature(fr, None, k=k, idx=idx)
    timeit(f"edge_fwd{C}", lambda: M.get_graph_feature(f, None, k=k, idx=idx), flush_l2=True)
    timeit(f"edge_bwd{C}", lambda: torch.autograd.grad(out, fr, g, retain_graph=True), flush_l2=True)
    del out, g
start = torch.arange(B) * 7 % N
timeit("fps512", lambda: M.fps_from_start(clouds, 512, start.to(dev)))
timeit("fps1024", lambda: M.fps_from_start(clouds, 1024, start.to(dev)))
timeit("fps_api512", lambda: M.farthest_point_sample(None, clouds, 512))
timeit("normals", lambda: M.estimate_normals(pts, 20))
timeit("density", lambda: M.cal_density(pts, 0.13, 16))
timeit("deform_voxels", lambda: M.deform_input(clouds.clone(), lookup, "volume_based_voxels", dev))
X = clouds.clone()
_, mask = M.deform_input(X, lookup, "volume_based_voxels", dev)
pred = (clouds.permute(0, 2, 1) + 0.05 * torch.randn(B, N, 3, device=dev)).contiguous().requires_grad_(True)


def chamfer():
    loss = M.reconstruction_loss(pred, clouds, mask)
    loss.backward()


timeit("chamfer_fwd", lambda: M.reconstruction_loss(pred, clouds, mask))
timeit("chamfer_fwd_bwd", chamfer)
