"""BatchNorm (training mode) fused with the activation that follows it, on mlsp_b200/csrc/bn.cu.

Stand-in for `act(bn(x))` in the reference's blocks -- conv_2d / fc_layer of PointDA/model_utils.py:45-89 (BatchNorm +
LeakyReLU(0.2)) and the Conv1d heads of PointDA/Models.py:165-285, PointSegDA/Models.py:245-392 (BatchNorm1d + ReLU): 3 passes
over the activation forward and 5 backward instead of torch's 5 + 8 (cuDNN BatchNorm plus an elementwise activation kernel).
The module stays the caller's nn.BatchNorm1d / nn.BatchNorm2d: its weight, bias and running statistics are read and updated
with torch's semantics (biased variance normalises, the unbiased one goes into running_var, num_batches_tracked counts).
In eval mode, with track_running_stats off, without a CUDA fp32 input or for a layout / channel count the kernels do not
take, the torch module and activation run as before (they are the caller's layers, not a hot-path op).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .ops import _DeviceGuard, _ptr, _stream


_scratch: dict = {}


def _scratch_doubles(C: int) -> int:
    n = _scratch.get(C)
    if n is None:
        fn = _lib.load().mlsp_bn_scratch_bytes
        fn.argtypes, fn.restype = [_lib._I], _lib._Z
        n = _scratch[C] = (int(fn(C)) + 7) // 8
    return n


def _layout(x: torch.Tensor):
    """-> (layout, R, C, L) for the kernels or None.  0: (R,C) rows (channels innermost); 1: (B,C,L) contiguous."""
    if x.dim() == 2:
        if x.is_contiguous() and x.shape[1] % 4 == 0 and x.shape[1] <= 1024:
            return 0, x.shape[0], x.shape[1], 1
        return None
    if x.dim() not in (3, 4):
        return None
    B, C = x.shape[0], x.shape[1]
    inner = x.numel() // (B * C)
    if x.is_contiguous():
        return 1, B, C, inner
    if x.dim() == 3 and x.stride(2) == 1 and x.stride(1) == inner and x.stride(0) >= C * inner:
        return 1, B, C, inner                                 # a channel slice of a wider (B,C',N) map (torch.split of the merged heads)
    last = x.permute(0, 2, 1) if x.dim() == 3 else x.permute(0, 2, 3, 1)
    if last.is_contiguous() and C % 4 == 0 and C <= 1024:
        return 0, B * inner, C, 1
    return None


class _BnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, slope, lay):
        layout, R, C, L = lay
        dense = x.is_contiguous() or layout == 0
        y = torch.empty_like(x) if dense else torch.empty(x.shape, dtype=x.dtype, device=x.device)   # dense: same strides as x
        xbs = 0 if dense else x.stride(0)
        save_mean = torch.empty(C, dtype=torch.float32, device=x.device)
        save_invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        acc = torch.empty(_scratch_doubles(C), dtype=torch.float64, device=x.device)
        with _DeviceGuard(x.device):
            _lib.call("mlsp_bn_act_fwd", _ptr(x), _ptr(y), R, C, L, layout, xbs, 0, _ptr(weight) if weight is not None else None,
                      _ptr(bias) if bias is not None else None, _ptr(running_mean) if running_mean is not None else None,
                      _ptr(running_var) if running_var is not None else None, float(momentum), float(eps), float(slope),
                      _ptr(save_mean), _ptr(save_invstd), _ptr(acc), _stream(x.device))
        ctx.save_for_backward(x, weight, bias, save_mean, save_invstd)
        ctx.lay, ctx.slope, ctx.xbs = lay, float(slope), xbs
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, save_mean, save_invstd = ctx.saved_tensors
        layout, R, C, L = ctx.lay
        if ctx.xbs:                                           # x is a channel slice: dy / dx are contiguous like y was
            dy = dy.contiguous()
            dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        else:
            if dy.stride() != x.stride():                     # the kernels walk x and dy with one index
                dy = torch.empty_like(x).copy_(dy)
            dx = torch.empty_like(x)
        need_w = weight is not None and ctx.needs_input_grad[1]
        need_b = bias is not None and ctx.needs_input_grad[2]
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device) if need_w else None
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device) if need_b else None
        acc = torch.empty(_scratch_doubles(C), dtype=torch.float64, device=x.device)
        with _DeviceGuard(x.device):
            _lib.call("mlsp_bn_act_bwd", _ptr(x), _ptr(dy), _ptr(dx), R, C, L, layout, ctx.xbs, 0, _ptr(weight) if weight is not None else None,
                      _ptr(bias) if bias is not None else None, _ptr(save_mean), _ptr(save_invstd), ctx.slope,
                      _ptr(dgamma) if need_w else None, _ptr(dbeta) if need_b else None, _ptr(acc), _stream(x.device))
        return dx, dgamma, dbeta, None, None, None, None, None, None


def bn_act(x: torch.Tensor, bn: torch.nn.modules.batchnorm._BatchNorm, slope: float = 1.0) -> torch.Tensor:
    """leaky_relu(bn(x), slope) -- slope 0.2: the reference's LeakyReLU, 0: ReLU, 1: BatchNorm alone.  x (B,C), (B,C,N) or
    (B,C,N,k), contiguous or channels-last.  Same values as the torch modules up to fp32 rounding (statistics in fp64)."""
    lay = _layout(x) if (x.is_cuda and x.dtype == torch.float32) else None
    fused = (lay is not None and bn.training and bn.track_running_stats and bn.momentum is not None
             and x.numel() // x.shape[1] > 1 and x.numel() > 0)
    if not fused:
        y = bn(x)
        if slope == 1.0:
            return y
        return F.relu(y) if slope == 0.0 else F.leaky_relu(y, slope)
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return _BnAct.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, slope, lay)
