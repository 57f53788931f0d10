"""PCM mix-up on the device: mirror of MLSP/PCM.py with the same signatures.

`mix_shapes(args, X, Y)` (MLSP/PCM.py:6-38) combines two shapes of the batch: farthest_point_sample of round(lam N) points of
every cloud and of N - round(lam N) points of a partner cloud, concatenated and point-permuted.  Here the two FPS calls and
the cat / permute run as ONE launch (mlsp_pcm_mix: 2B CTAs instead of 2 x B, no intermediate tensors).  The host RNG streams
are consumed exactly like the reference does -- torch.randperm(B), np.random.beta, torch.randint x 2 (inside the two FPS calls,
utils/pc_utils.py:150), torch.randperm(N) -- so a seeded run mixes the same shapes into the same clouds, bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .ops import _DeviceGuard, _ptr, _require_cuda_f32, _stream


def mix_shapes(args, X: torch.Tensor, Y: torch.Tensor):
    """mix_shapes(args, X, Y): MLSP/PCM.py:6-38.  X (B,3,N) CUDA float32, Y (B) labels -> (mixed_X (B,3,N), (Y_a, Y_b, lam))."""
    _require_cuda_f32(X, "mix_shapes")
    B, C, N = X.shape
    if C != 3:
        raise _lib.MlspError("mix_shapes: expected (B,3,N)")
    X = X.detach().contiguous()
    dev = X.device
    index = torch.randperm(B)                                                     # PCM.py:20 (CPU generator)
    mixup = getattr(args, "mixup_params", 1.0)
    lam = np.random.beta(mixup, mixup) if mixup > 0 else 1.0                      # PCM.py:23
    num_pts_a = round(lam * N)
    start_a = torch.randint(0, N, (B,), dtype=torch.long)                         # the draws of the two FPS calls (pc_utils.py:150)
    start_b = torch.randint(0, N, (B,), dtype=torch.long)
    points_perm = torch.randperm(N)                                               # PCM.py:32
    inv = torch.empty(N, dtype=torch.int32)
    inv[points_perm] = torch.arange(N, dtype=torch.int32)
    host = torch.cat([index, start_a, start_b]).pin_memory() if dev.type == "cuda" else torch.cat([index, start_a, start_b])
    meta = host.to(dev, non_blocking=True)
    inv_d = inv.pin_memory().to(dev, non_blocking=True)
    out = torch.empty_like(X)
    with _DeviceGuard(dev):
        _lib.call("mlsp_pcm_mix", _ptr(X), B, N, int(num_pts_a), _ptr(meta[:B]), _ptr(meta[B:]), _ptr(inv_d), _ptr(out), _stream(dev))
    index_d = meta[:B]
    return out, (Y.clone(), Y[index_d].clone(), lam)


def calc_loss(args, logits, mixup_vals, criterion):
    """calc_loss(args, logits, mixup_vals, criterion): MLSP/PCM.py:76-89 (convex combination of the two label losses)."""
    Y_a, Y_b, lam = mixup_vals
    loss = lam * criterion(logits["cls"], Y_a) + (1 - lam) * criterion(logits["cls"], Y_b)
    return loss * (1 - args.DefRec_weight)
