#!/usr/bin/env python
"""Minimal driver for `ncu --profile-from-start off`: every op of one hotpath-A step exactly once (after a warm-up pass),
so that a `--set full` capture stays short and each kernel of the step appears once per shape in the report.
   ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof python tools/prof_ops.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402

B, N, k = synth.CONFIGS["A"]
dev = torch.device("cuda:0")
clouds = synth.surface_clouds(B, N, 1234).to(dev)
pts = clouds.permute(0, 2, 1).contiguous()
feats = {3: clouds, 64: synth.smooth_features(B, 64, N, 1244).to(dev), 128: synth.smooth_features(B, 128, N, 1246).to(dev)}
grads = {C: torch.randn(B, N, k, 2 * C, device=dev).permute(0, 3, 1, 2) for C in feats}
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
pred = (pts + 0.05 * torch.randn(B, N, 3, device=dev)).contiguous()
start = (torch.arange(B) * 7 % N).to(dev)


from mlsp_b200 import bn as mbn  # noqa: E402
bn_cases = [(torch.randn(B, N, k, 128, device=dev).permute(0, 3, 1, 2), torch.nn.BatchNorm2d(128).to(dev).train()),
            (torch.randn(B, 1024, N, device=dev), torch.nn.BatchNorm1d(1024).to(dev).train())]
clouds_s = synth.surface_clouds(16, 2048, 77).to(dev)
start_s = (torch.arange(16) * 7 % 2048).to(dev)


def one_pass():
    np.random.seed(1)
    for C, f in feats.items():                                   # the DGCNN call: knn + gather fused, then its backward
        x = f.detach().requires_grad_(True)
        M.get_graph_feature(x, None, k=k).backward(grads[C])
    M.fps_from_start(clouds, 512, start)
    M.target_structure(pts, 20, 0.13, 16, 2, 0)                  # 8f rank 2: kNN(near) + normals + cardinality in one launch
    M.estimate_normals(pts, 20)                                  # the stand-alone ops, for reference: knn3 + pca_normals
    M.cal_density(pts, 0.13, 16, 2, 0)
    M.scan_input(pts.clone(), dev)                               # 8f rank 3: z-buffer scan
    X, mask = M.deform_input(clouds.clone(), lookup, "volume_based_voxels", dev)
    p = pred.detach().requires_grad_(True)
    M.reconstruction_loss(p, clouds, mask).backward()
    M.knn(feats[64], k)                                          # the ranking kernels without the fused gather, for reference
    M.get_graph_feature(feats[64], None, k=k, idx=M.knn(feats[64], k))   # explicit-idx gather (edge_fwd_vec_kernel)
    # late round 2: BatchNorm + LeakyReLU pass pairs on the training step's largest maps (rows and (B,C,L) layouts), FPS at the
    # PointSegDA shape (SM reserved, two clouds per CTA), kNN with k > 64 (two rounds of the exact kernel)
    for xb, mod in bn_cases:
        xin = xb.detach().requires_grad_(True)
        mbn.bn_act(xin, mod, 0.2).backward(xb)
    M.fps_from_start(clouds_s, 1024, start_s)
    M.knn(feats[64][:4].contiguous(), 100)


one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.start()
one_pass()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
