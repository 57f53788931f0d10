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
from .ops import _ptr, _stream


def _mat(t: torch.Tensor, name: str):
    """(rows, K) view [optionally batched (Z, rows, K)] -> (kmajor, ld, batch stride) or None if it needs a copy."""
    if t.dim() == 2:
        s_r, s_k, s_z = t.stride(0), t.stride(1), 0
        rows, K = t.shape
    else:
        s_z, s_r, s_k = t.stride()
        _, rows, K = t.shape
    if s_k == 1 and (s_r >= K or rows == 1):
        return 1, max(s_r, K), s_z
    if s_r == 1 and (s_k >= rows or K == 1):
        return 0, max(s_k, rows), s_z
    return None


def gemm_nt(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None, out: torch.Tensor | None = None,
            out_colmajor: bool = False) -> torch.Tensor:
    """D = a @ b^T (+ bias over the last dimension):  a (M,K) or (Z,M,K), b (N,K) or (Z,N,K) [2-D operands are shared by
    all Z], any strides with one unit stride per matrix (a view that has none is copied).  out: (M,N) / (Z,M,N) tensor or
    view whose last two strides are (ld,1) or (1,ld); default a new contiguous tensor, or -- out_colmajor=True -- a tensor
    stored (Z,N,M) and returned as its (Z,M,N) transpose view (how a (B,O,N) feature map wants a points x channels product).
    float32 CUDA only."""
    for t, nm in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != torch.float32 or t.dim() not in (2, 3):
            raise MlspError(f"gemm_nt: {nm} must be a 2-D or 3-D float32 CUDA tensor")
    Z = a.shape[0] if a.dim() == 3 else (b.shape[0] if b.dim() == 3 else (out.shape[0] if out is not None and out.dim() == 3 else 0))
    batched = a.dim() == 3 or b.dim() == 3
    M, K = a.shape[-2:]
    N, Kb = b.shape[-2:]
    if K != Kb or (a.dim() == 3 and b.dim() == 3 and a.shape[0] != b.shape[0]):
        raise MlspError(f"gemm_nt: shapes {tuple(a.shape)} x {tuple(b.shape)}^T do not match")
    if K == 0:
        raise MlspError("gemm_nt: K = 0")
    a = a.detach()
    b = b.detach()
    la = _mat(a, "a")
    if la is None:
        a = a.contiguous()
        la = _mat(a, "a")
    lb = _mat(b, "b")
    if lb is None:
        b = b.contiguous()
        lb = _mat(b, "b")
    dev = a.device
    if out is None:
        if out_colmajor:
            store = torch.empty(((Z, N, M) if batched else (N, M)), dtype=torch.float32, device=dev)
            out = store.transpose(-1, -2)
        else:
            out = torch.empty(((Z, M, N) if batched else (M, N)), dtype=torch.float32, device=dev)
    else:
        if out.dtype != torch.float32 or out.device != dev or tuple(out.shape[-2:]) != (M, N) or (out.dim() == 3) != batched:
            raise MlspError("gemm_nt: out has the wrong shape / dtype / device")
    ld_ = _mat(out, "out")                      # "kmajor" here means the last dimension (n) is contiguous = row-major
    if ld_ is None:
        raise MlspError("gemm_nt: out needs a unit stride along m or n")
    if bias is not None:
        if bias.shape != (N,) or bias.dtype != torch.float32 or bias.device != dev:
            raise MlspError("gemm_nt: bias must be float32 (N,) on the operands' device")
        bias = bias.detach().contiguous()
    if M == 0 or N == 0 or (batched and Z == 0):
        return out
    with torch.cuda.device(dev):
        _lib.call("mlsp_gemm_f32", _ptr(a), la[0], la[1], la[2] if a.dim() == 3 else 0, _ptr(b), lb[0], lb[1],
                  lb[2] if b.dim() == 3 else 0, _ptr(out), ld_[0], ld_[1], ld_[2] if out.dim() == 3 else 0, _ptr(bias),
                  M, N, K, Z if batched else 1, _stream(dev))
    return out


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
        y = gemm_nt(x, weight, bias)
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
            gx = gemm_nt(gy, weight.t())                                                     # (R,C)
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
