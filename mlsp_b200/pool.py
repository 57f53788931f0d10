"""Max pooling of the DGCNN's feature maps on the streaming kernels of mlsp_b200/csrc/pool.cu (forward with argmax, backward).

Stand-ins for `x.max(dim=-1)[0]` over the neighbours of a channels-last edge tensor, `torch.max(x, dim=2)[0]` over the
points of a channels-last map (transform_net, PointDA/model_utils.py:116-121) and `F.adaptive_max_pool1d(x, 1)`
(PointDA/Models.py:133).  Same values as torch; the gradient goes to the first maximal element.  No CPU path."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import MlspError
from .ops import _DeviceGuard, _ptr, _require_cuda_f32, _stream


class _MaxMid(torch.autograd.Function):
    """x (R,K,C) contiguous -> (R,C) = max over K."""

    @staticmethod
    def forward(ctx, x):
        R, K, C = x.shape
        val = torch.empty((R, C), dtype=torch.float32, device=x.device)
        arg = torch.empty((R, C), dtype=torch.int32, device=x.device)
        with _DeviceGuard(x.device):
            _lib.call("mlsp_max_mid_fwd", _ptr(x), R, K, C, _ptr(val), _ptr(arg), _stream(x.device))
        ctx.save_for_backward(arg)
        ctx.K = K
        return val

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        R, C = arg.shape
        g = g.contiguous()
        gin = torch.empty((R, ctx.K, C), dtype=torch.float32, device=g.device)
        with _DeviceGuard(g.device):
            _lib.call("mlsp_max_mid_bwd", _ptr(g), _ptr(arg), R, ctx.K, C, _ptr(gin), _stream(g.device))
        return gin


class _MaxRow(torch.autograd.Function):
    """x (R,K) contiguous -> (R) = max over K."""

    @staticmethod
    def forward(ctx, x):
        R, K = x.shape
        val = torch.empty(R, dtype=torch.float32, device=x.device)
        arg = torch.empty(R, dtype=torch.int32, device=x.device)
        with _DeviceGuard(x.device):
            _lib.call("mlsp_max_row_fwd", _ptr(x), R, K, _ptr(val), _ptr(arg), _stream(x.device))
        ctx.save_for_backward(arg)
        ctx.K = K
        return val

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        R = arg.shape[0]
        g = g.contiguous()
        gin = torch.empty((R, ctx.K), dtype=torch.float32, device=g.device)
        with _DeviceGuard(g.device):
            _lib.call("mlsp_max_row_bwd", _ptr(g), _ptr(arg), R, ctx.K, _ptr(gin), _stream(g.device))
        return gin


def max_over_neighbours(x: torch.Tensor) -> torch.Tensor:
    """x.max(dim=-1)[0] for x (B,C,N,k) in channels-last strides (memory (B,N,k,C), what conv1x1 / BatchNorm2d / LeakyReLU keep)
    -> (B,C,N) in channels-last strides (memory (B,N,C)).  Other layouts are made channels-last first (one copy)."""
    _require_cuda_f32(x, "max_over_neighbours")
    B, C, N, k = x.shape
    if C % 4:
        raise MlspError("max_over_neighbours: channels must be a multiple of 4")
    xm = x.permute(0, 2, 3, 1)
    if not xm.is_contiguous():
        xm = xm.contiguous()
    return _MaxMid.apply(xm.reshape(B * N, k, C)).view(B, N, C).permute(0, 2, 1)


def max_over_points_cl(x: torch.Tensor) -> torch.Tensor:
    """torch.max(x, dim=2)[0] for x (B,C,N,1) in channels-last strides (memory (B,N,C)) -> (B,C,1)."""
    _require_cuda_f32(x, "max_over_points_cl")
    B, C, N, one = x.shape
    if one != 1 or C % 4:
        raise MlspError("max_over_points_cl: expected (B,C,N,1) with C a multiple of 4")
    xm = x.permute(0, 2, 3, 1).reshape(B, N, C)
    if not xm.is_contiguous():
        xm = xm.contiguous()
    return _MaxMid.apply(xm).view(B, C, 1)


def global_max_pool(x: torch.Tensor) -> torch.Tensor:
    """F.adaptive_max_pool1d(x, 1) for x (B,C,N) -> (B,C,1)."""
    _require_cuda_f32(x, "global_max_pool")
    B, C, N = x.shape
    return _MaxRow.apply(x.contiguous().view(B * C, N)).view(B, C, 1)
