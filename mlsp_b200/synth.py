"""Seeded synthetic inputs of the PointDA-10 / PointSegDA shapes.

There is no network for datasets, so every test and benchmark runs on these.  The generator
follows the reference's own normalisation (`scale_to_unit_cube`, utils/pc_utils.py:263-277:
subtract the centroid, divide by the largest norm), which is what its data loaders feed the
model (PointDA/data/dataloader.py).  CPU generator, so the same seed gives the same clouds on
every box.
"""
from __future__ import annotations

import torch

CONFIGS = {
    # name: (clouds per GPU, points, k)
    "A": (32, 1024, 20),    # PointDA-10 shape (BASELINE.json configs[0..2])
    "S": (16, 2048, 20),    # PointSegDA shape (configs[3])
    "X": (256, 4096, 40),   # scaling sweep (configs[4])
}


def clouds(B: int, N: int, seed: int = 1234, quantised: bool = False) -> torch.Tensor:
    """(B,3,N) float32 unit-ball clouds.  quantised=True rounds to multiples of 2^-9, the set on
    which every fp32 evaluation order of the distance formulas agrees exactly (ties are real)."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, N, 3, generator=g)
    p = p - p.mean(dim=1, keepdim=True)
    p = p / p.norm(dim=2).max(dim=1).values.view(B, 1, 1)
    if quantised:
        p = torch.round(p * 512.0) / 512.0
    return p.permute(0, 2, 1).contiguous()


def surface_clouds(B: int, N: int, seed: int = 1234) -> torch.Tensor:
    """(B,3,N) clouds sampled on a noisy ellipsoid surface -- closer to CAD/scan data than a
    Gaussian blob (well-defined normals, region histograms with >= 40 points per voxel)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(B, N, 3, generator=g)
    d = d / d.norm(dim=2, keepdim=True)
    axes = 0.5 + 0.5 * torch.rand(B, 1, 3, generator=g)
    p = d * axes + 0.01 * torch.randn(B, N, 3, generator=g)
    p = p - p.mean(dim=1, keepdim=True)
    p = p / p.norm(dim=2).max(dim=1).values.view(B, 1, 1)
    return p.permute(0, 2, 1).contiguous()


def features(B: int, C: int, N: int, seed: int = 1234, quantised: bool = False) -> torch.Tensor:
    """(B,C,N) LeakyReLU(0.2)-shaped activations like DGCNN's x1..x3 (PointDA/Models.py:116-125).
    quantised=True: multiples of 2^-6 with |x| <= 4, exact under any fp32 summation order for C <= 256."""
    g = torch.Generator().manual_seed(seed)
    x = torch.nn.functional.leaky_relu(torch.randn(B, C, N, generator=g), 0.2)
    if quantised:
        x = torch.clamp(torch.round(x * 64.0) / 64.0, -4.0, 4.0)
    return x.contiguous()


def smooth_features(B: int, C: int, N: int, seed: int = 1234) -> torch.Tensor:
    """(B,C,N) features that are a smooth random map of a 3-D cloud (a random 2-layer MLP), so the
    feature-space neighbourhoods have low intrinsic dimension like real DGCNN activations."""
    g = torch.Generator().manual_seed(seed)
    p = clouds(B, N, seed + 7)
    w1 = torch.randn(64, 3, generator=g)
    w2 = torch.randn(C, 64, generator=g) / 8.0
    h = torch.nn.functional.leaky_relu(torch.einsum("oc,bcn->bon", w1, p), 0.2)
    return torch.nn.functional.leaky_relu(torch.einsum("oc,bcn->bon", w2, h), 0.2).contiguous()
