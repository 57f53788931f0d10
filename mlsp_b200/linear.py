"""The DGCNN's point-wise layers (nn.Conv1d / nn.Conv2d with kernel 1, nn.Linear) on the hand-written tcgen05 GEMM
`mlsp_gemm_f32` (mlsp_b200/csrc/gemm.cu, include/mlsp_b200.h) -- SURVEY.md section 8f ranks 1 and 4.

Reference call sites: conv_2d / fc_layer / transform_net (PointDA/model_utils.py:45-130), the EdgeConv layers, conv5 and
the heads (PointDA/Models.py:114-131, 156-160, 165-285).  The result is an fp32 product up to summation order (three
bf16 pieces per operand, six piece products, fp32 accumulation), so these functions stand in for torch's fp32 layers with
TF32 off -- what the reference's trainers run.

`gemm_nt` takes strided VIEWS: transposes are expressed with `.transpose()` and cost nothing -- the kernel reads either
operand K-major or MN-major and writes D in either orientation.  No CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import MlspError
from .ops import _DeviceGuard, _ptr, _stream


def _mat(t: torch.Tensor, name: str = ""):
    """(rows, K) view [optionally batched (Z, rows, K)] -> (kmajor, ld, batch stride) or None if it needs a copy."""
    st = t.stride()
    sh = t.shape
    if len(st) == 2:
        s_r, s_k = st
        s_z = 0
        rows, K = sh
    else:
        s_z, s_r, s_k = st
        rows, K = sh[1], sh[2]
    if s_k == 1 and (s_r >= K or rows == 1):
        return 1, (s_r if s_r > K else K), s_z
    if s_r == 1 and (s_k >= rows or K == 1):
        return 0, (s_k if s_k > rows else rows), s_z
    return None


def gemm_nt(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None, out: torch.Tensor | None = None,
            out_colmajor: bool = False) -> torch.Tensor:
    """D = a @ b^T (+ bias over the last dimension):  a (M,K) or (Z,M,K), b (N,K) or (Z,N,K) [2-D operands are shared by
    all Z], any strides with one unit stride per matrix (a view that has none is copied).  out: (M,N) / (Z,M,N) tensor or
    view whose last two strides are (ld,1) or (1,ld); default a new contiguous tensor, or -- out_colmajor=True -- a tensor
    stored (Z,N,M) and returned as its (Z,M,N) transpose view (how a (B,O,N) feature map wants a points x channels product).
    float32 CUDA only.  (Kept lean: the training step calls it ~120 times and is host-bound.)"""
    da, db = a.dim(), b.dim()
    if not (a.is_cuda and b.is_cuda) or a.dtype != torch.float32 or b.dtype != torch.float32 or da not in (2, 3) or db not in (2, 3):
        raise MlspError("gemm_nt: a and b must be 2-D or 3-D float32 CUDA tensors")
    batched = da == 3 or db == 3
    M, K = a.shape[-2], a.shape[-1]
    N, Kb = b.shape[-2], b.shape[-1]
    Z = a.shape[0] if da == 3 else (b.shape[0] if db == 3 else 1)
    if K != Kb or (da == 3 and db == 3 and a.shape[0] != b.shape[0]):
        raise MlspError(f"gemm_nt: shapes {tuple(a.shape)} x {tuple(b.shape)}^T do not match")
    if K == 0:
        raise MlspError("gemm_nt: K = 0")
    la = _mat(a)
    if la is None:
        a = a.contiguous()
        la = _mat(a)
    lb = _mat(b)
    if lb is None:
        b = b.contiguous()
        lb = _mat(b)
    dev = a.device
    if out is None:
        if out_colmajor:
            out = torch.empty(((Z, N, M) if batched else (N, M)), dtype=torch.float32, device=dev).transpose(-1, -2)
        else:
            out = torch.empty(((Z, M, N) if batched else (M, N)), dtype=torch.float32, device=dev)
    elif out.dtype != torch.float32 or out.device != dev or tuple(out.shape[-2:]) != (M, N) or (out.dim() == 3) != batched:
        raise MlspError("gemm_nt: out has the wrong shape / dtype / device")
    ld_ = _mat(out)                             # "kmajor" here means the last dimension (n) is contiguous = row-major
    if ld_ is None:
        raise MlspError("gemm_nt: out needs a unit stride along m or n")
    if bias is not None:
        if bias.shape != (N,) or bias.dtype != torch.float32 or bias.device != dev:
            raise MlspError("gemm_nt: bias must be float32 (N,) on the operands' device")
        if not bias.is_contiguous():
            bias = bias.contiguous()
    if M == 0 or N == 0 or Z == 0:
        return out
    with _DeviceGuard(dev):
        _lib.call("mlsp_gemm_f32", a.data_ptr(), la[0], la[1], la[2] if da == 3 else 0, b.data_ptr(), lb[0], lb[1],
                  lb[2] if db == 3 else 0, out.data_ptr(), ld_[0], ld_[1], ld_[2] if batched else 0,
                  bias.data_ptr() if bias is not None else None, M, N, K, Z, _stream(dev))
    return out


def _gemm_small_m(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """gemm_nt for a product with few output tiles and a long K (the Linear layers on B rows: M = 32, K = 256..1024): as ONE
    m-tile it would be a serial chain of K/64 chunks on a handful of CTAs (25 us for 0.03 GFLOP), so K is cut into 64-wide
    chunks that run as one batched product (every chunk a CTA) and are summed -- 3x faster, same fp32 arithmetic."""
    M, K = a.shape
    N = b.shape[0]
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    if a.dim() != 2 or b.dim() != 2 or tiles > 16 or K < 256 or K % 64:
        return gemm_nt(a, b, bias)
    Z = K // 64
    while Z * tiles > 296 and Z % 2 == 0:
        Z //= 2
    kc = K // Z
    part = gemm_nt(a.unflatten(1, (Z, kc)).permute(1, 0, 2), b.unflatten(1, (Z, kc)).permute(1, 0, 2))    # (Z,M,N)
    out = part.sum(dim=0)
    return out if bias is None else out.add_(bias)


class _Conv1x1(torch.autograd.Function):
    """y (B,O,N) = W (O,C) x (B,C,N) + bias  -- nn.Conv1d(kernel_size=1) on the channel-major feature maps of the model.
    The kernel sees points as M (x[b] is its M-major A operand), channels as N/K; y comes out channel-major directly.
    Backward: dx = W^T dy in the same layout; dW = sum_b dy[b] x[b]^T as B partial products (K = N points each) summed."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, C, N = x.shape
        y = gemm_nt(x.transpose(1, 2), weight, bias, out_colmajor=True).transpose(1, 2)      # (B,O,N) contiguous storage
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gx = gw = gb = None
        if _mat(gy, "gy") is None:
            gy = gy.contiguous()
        if ctx.needs_input_grad[0]:
            gx = gemm_nt(gy.transpose(1, 2), weight.t(), out_colmajor=True).transpose(1, 2)  # (B,C,N)
        if ctx.needs_input_grad[1]:
            gw = gemm_nt(gy, x).sum(dim=0)                                                   # (B,O,C) partials -> (O,C)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(dim=(0, 2))
        return gx, gw, gb


class _Linear(torch.autograd.Function):
    """y (R,O) = x (R,C) W^T (O,C) + bias -- nn.Linear, and nn.Conv2d(kernel_size=1) on the channels-last edge tensor.
    dW (O,C) = dy^T x is a reduction over all R rows: split into chunks of rows (partial products, summed) so that the
    whole machine works on it."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = _gemm_small_m(x, weight, bias) if x.shape[0] <= 128 else gemm_nt(x, weight, bias)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gx = gw = gb = None
        if _mat(gy, "gy") is None:
            gy = gy.contiguous()
        R = x.shape[0]
        if ctx.needs_input_grad[0]:
            gx = _gemm_small_m(gy, weight.t()) if R <= 128 else gemm_nt(gy, weight.t())      # (R,C)
        if ctx.needs_input_grad[1]:
            gw = _reduce_rows_product(gy, x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(dim=0)
        return gx, gw, gb


def _reduce_rows_product(gy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """gy^T x : (R,O),(R,C) -> (O,C), the R rows cut into Z chunks that run as one batched product."""
    R, O = gy.shape
    C = x.shape[1]
    if not (gy.is_contiguous() and x.is_contiguous()):
        gy, x = gy.contiguous(), x.contiguous()
    tiles = ((O + 127) // 128) * ((C + 127) // 128)
    Z = max(1, min(R // 512, (296 + tiles - 1) // tiles))
    while Z > 1 and R % Z:
        Z -= 1
    if Z == 1:
        return gemm_nt(gy.t(), x.t())
    part = gemm_nt(gy.view(Z, R // Z, O).transpose(1, 2), x.view(Z, R // Z, C).transpose(1, 2))   # (Z,O,C)
    return part.sum(dim=0)


class _ApplyTransform(torch.autograd.Function):
    """y[b] (O,N) = T[b] (O,C) x[b] (C,N) with a per-cloud matrix -- the input-transform product of PointDA/Models.py:113."""

    @staticmethod
    def forward(ctx, T, x):
        ctx.save_for_backward(T, x)
        return gemm_nt(x.transpose(1, 2), T, out_colmajor=True).transpose(1, 2)

    @staticmethod
    def backward(ctx, gy):
        T, x = ctx.saved_tensors
        gT = gx = None
        if _mat(gy, "gy") is None:
            gy = gy.contiguous()
        if ctx.needs_input_grad[0]:
            gT = gemm_nt(gy, x)                                                              # (B,O,C), K = N points
        if ctx.needs_input_grad[1]:
            gx = gemm_nt(gy.transpose(1, 2), T.transpose(1, 2), out_colmajor=True).transpose(1, 2)
        return gT, gx


def apply_transform(T: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """torch.matmul(T, x) for T (B,O,C), x (B,C,N)."""
    return _ApplyTransform.apply(T, x if _mat(x.transpose(1, 2), "x") is not None else x.contiguous())


def conv1x1(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """nn.Conv1d(kernel_size=1) / nn.Conv2d(kernel_size=1):  x (B,C,N) -> (B,O,N);  x (B,C,N,k) in channels-last strides (what
    get_graph_feature returns) -> (B,O,N,k) in channels-last strides.  weight (O,C[,1[,1]])."""
    W = weight.flatten(1)
    if x.dim() == 3:
        return _Conv1x1.apply(x if _mat(x.transpose(1, 2), "x") is not None else x.contiguous(), W, bias)
    B, C, N, k = x.shape
    xf = x.permute(0, 2, 3, 1).reshape(B * N * k, C)
    y = _Linear.apply(xf, W, bias)
    return y.view(B, N, k, -1).permute(0, 3, 1, 2)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """torch.nn.functional.linear for 2-D x."""
    return _Linear.apply(x, weight, bias)
