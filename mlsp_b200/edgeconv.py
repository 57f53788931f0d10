"""EdgeConv without the edge tensor -- an opt-in, Models.py-level swap.

The reference builds every DGCNN layer as

    x = get_graph_feature(x, args, k)        # (B,2C,N,k), 0.34-0.67 GB per layer at 32x1024
    x = conv_2d(x)                           # 1x1 Conv2d [+ BatchNorm2d] + LeakyReLU      PointDA/model_utils.py:45-63
    x = x.max(dim=-1, keepdim=False)[0]      # PointDA/Models.py:114-128; PointSegDA/Models.py:171-184 (stacked plain convs)

`edge_conv` computes the same (B,O,N) result and the same gradients without materialising anything of size
B*N*k: the convolution is linear in [x_j - x_i | x_i], so  h_ij = Y[idx_ij] + Z[i]  with two point-wise products
(one tcgen05 GEMM on (B*N,C) x (C,2O), mlsp_gemm_f32); BatchNorm in training mode needs only sum h and sum h^2 over the edges;
BatchNorm + LeakyReLU are monotone per channel, so the max over k commutes with them (min where gamma < 0).
Kernels: mlsp_b200/csrc/edgeconv.cu behind mlsp_edgeconv_* (include/mlsp_b200.h).  No CPU path.

Results agree with the reference composition to fp32 rounding (a different, equally valid association of the same
sums: tests/test_gpu_parity.py::test_edge_conv_*, tolerance 1e-5 relative as north_star states for floating point).
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib, linear
from ._lib import MlspError
from .ops import _DeviceGuard, _ptr, _require_cuda_f32, _stream, knn


class _EdgeConvReduce(torch.autograd.Function):
    """(yz, idx, p0, p1) -> out (B,O,N) = lrelu(a * max_j h_j + c), h_j = Y[idx_j] + Z.
    mode "bn_train": p0 = gamma >= 0, p1 = beta (None = 1 / 0); batch statistics over the B*N*k edges.
    mode "affine"  : p0 = a >= 0, p1 = c per channel (eval-mode BatchNorm folded, or a = 1 / c = 0).
    The caller (edge_conv) folds the sign of gamma / a into the rows of the weight, so the extreme is always a max.
    running = (running_mean, running_var, sign or None, momentum): updated in place by the coefficient kernel."""

    @staticmethod
    def forward(ctx, yz, idx, p0, p1, mode, eps, slope, running=None):
        B, N, O2 = yz.shape
        O = O2 // 2
        k = idx.shape[2]
        dev = yz.device
        yz = yz.contiguous()
        hsel = torch.empty((B, N, O), dtype=torch.float32, device=dev)
        slot = torch.empty((B, N, O), dtype=torch.uint8, device=dev)
        coef = torch.empty((4, O), dtype=torch.float32, device=dev)
        out = torch.empty((B, O, N), dtype=torch.float32, device=dev)
        train = mode == "bn_train"
        rowsum = None
        p0c = p0.detach().float().contiguous() if p0 is not None else None
        p1c = p1.detach().float().contiguous() if p1 is not None else None
        with _DeviceGuard(dev):
            s = _stream(dev)
            if train:
                rowsum = torch.empty((B, N, O), dtype=torch.float32, device=dev)
                stats = torch.empty((2, O), dtype=torch.float64, device=dev)
                rmean, rvar, rsign, mom = running if running is not None else (None, None, None, 0.0)
                _lib.call("mlsp_edgeconv_reduce_fwd", _ptr(yz), _ptr(idx), B, N, O, k, _ptr(hsel), _ptr(slot),
                          _ptr(rowsum), _ptr(stats), s)
                _lib.call("mlsp_edgeconv_bn_coeffs", _ptr(stats), _ptr(p0c), _ptr(p1c), O, float(B * N * k), float(eps),
                          _ptr(coef), _ptr(rmean), _ptr(rvar), _ptr(rsign), float(mom), s)
            else:
                coef[0] = p0c if p0c is not None else 1.0
                coef[1] = p1c if p1c is not None else 0.0
                coef[2] = 0.0          # with mean = 0, invstd = 1 the backward sums are the gradients of (a, c)
                coef[3] = 1.0
                _lib.call("mlsp_edgeconv_reduce_fwd", _ptr(yz), _ptr(idx), B, N, O, k, _ptr(hsel), _ptr(slot),
                          ctypes.c_void_p(0), ctypes.c_void_p(0), s)
            _lib.call("mlsp_edgeconv_apply_fwd", _ptr(hsel), _ptr(coef), B, N, O, float(slope), _ptr(out), s)
        ctx.save_for_backward(yz, idx, hsel, slot, coef, rowsum if train else hsel)
        ctx.cfg = (B, N, O, k, float(slope), train, p0 is not None, p1 is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        yz, idx, hsel, slot, coef, rowsum = ctx.saved_tensors
        B, N, O, k, slope, train, has0, has1 = ctx.cfg
        dev = yz.device
        g = g.float()
        if g.stride(2) != 1 or g.stride(1) != N or g.stride(0) < O * N:   # a channel slice of a cat gradient is used in place
            g = g.contiguous()
        dyz = torch.empty((B, N, 2 * O), dtype=torch.float32, device=dev)
        need_p = train or has0 or has1
        dp = torch.empty((2, O), dtype=torch.float32, device=dev) if need_p else None
        with _DeviceGuard(dev):
            nbytes = max(_lib.workspace_bytes(_lib.OP_EDGECONV_BWD, B, O, N, k), 16)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("mlsp_edgeconv_bwd", _ptr(g), int(g.stride(0)), _ptr(yz), _ptr(idx), _ptr(hsel), _ptr(slot),
                      _ptr(rowsum) if train else ctypes.c_void_p(0), _ptr(coef), B, N, O, k, slope, 1 if train else 0,
                      _ptr(dyz), _ptr(dp), _ptr(ws), ws.numel(), _stream(dev))
        return dyz, None, (dp[0] if has0 else None), (dp[1] if has1 else None), None, None, None, None


class _PointwiseYZ(torch.autograd.Function):
    """yz (B,N,2O) = [Y | Z],  Y = s*Wa x,  Z = s*(Wb - Wa) x + s*bias  -- the layer's one GEMM (mlsp_gemm_f32: tcgen05, fp32
    operands as three bf16 pieces, fp32-faithful) together with the weight split W = [Wa | Wb] -> [Wa ; Wb - Wa] and the
    sign fold (s = sign of the BatchNorm / affine scale per output channel, made by the same prep launch).  One autograd node: every operand is read in the
    layout it already has (x and grad_x are (B,C,N), yz and its gradient (B,N,2O): no transposed copies), the weight
    gradient is made from B partial products (K = N points each) summed afterwards, and un-split."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale):
        B, C, _ = x.shape
        W = weight.detach().reshape(weight.shape[0], -1).contiguous()
        O = W.shape[0]
        if W.shape[1] != 2 * C:
            raise MlspError(f"edge_conv: weight has {W.shape[1]} input channels, expected 2*C = {2 * C}")
        dev = x.device
        Wcat = torch.empty((2 * O, C), dtype=torch.float32, device=dev)
        sgn = torch.empty(O, dtype=torch.float32, device=dev)
        zb = torch.empty(2 * O, dtype=torch.float32, device=dev) if bias is not None else None
        with _DeviceGuard(dev):                                           # weight split, sign fold, bias: one launch
            _lib.call("mlsp_edgeconv_weight_prep", _ptr(W), _ptr(scale.detach().float().contiguous() if scale is not None else None),
                      _ptr(bias.detach().float().contiguous() if bias is not None else None), O, C, _ptr(Wcat), _ptr(sgn), _ptr(zb),
                      _stream(dev))
        # x[b] (C,N) is the GEMM's M-major A operand as it lies in memory; yz comes out point-major (B,N,2O)
        yz = linear.gemm_nt(x.transpose(1, 2), Wcat, zb)
        ctx.save_for_backward(x, Wcat, sgn)
        ctx.wshape = weight.shape
        ctx.mark_non_differentiable(sgn)
        return yz, sgn

    @staticmethod
    def backward(ctx, dyz, _dsgn):
        x, Wcat, sgn = ctx.saved_tensors
        O, C = Wcat.shape[0] // 2, Wcat.shape[1]
        gx = gw = gb = None
        dyz = dyz.contiguous()
        if ctx.needs_input_grad[0]:
            gx = linear.gemm_nt(dyz, Wcat.t(), out_colmajor=True).transpose(1, 2)   # (B,N,2O) x (2O,C) -> stored (B,C,N)
        if ctx.needs_input_grad[1]:
            part = linear.gemm_nt(dyz.transpose(1, 2), x)                         # (B,2O,C) partial products, K = N points each
            gw = torch.empty((O, 2 * C), dtype=torch.float32, device=x.device)
            with _DeviceGuard(x.device):                                     # sum over B, sign, un-split: one launch
                _lib.call("mlsp_edgeconv_weight_grad", _ptr(part), part.shape[0], _ptr(sgn), O, C, _ptr(gw), _stream(x.device))
            gw = gw.reshape(ctx.wshape)
        if ctx.needs_input_grad[2]:
            gb = dyz[..., O:].sum(dim=(0, 1)) * sgn
        return gx, gw, gb, None


def _split_weight(weight: torch.Tensor, C: int) -> torch.Tensor:
    """W (O,2C[,1,1]) = [Wa | Wb] over [x_j - x_i | x_i]  ->  (2O, C) = [Wa ; Wb - Wa]  (rows of Y, then rows of Z)."""
    W = weight.reshape(weight.shape[0], -1)
    if W.shape[1] != 2 * C:
        raise MlspError(f"edge_conv: weight has {W.shape[1]} input channels, expected 2*C = {2 * C}")
    Wa, Wb = W[:, :C], W[:, C:]
    return torch.cat((Wa, Wb - Wa), dim=0)


class BNParams:
    """BatchNorm2d as torch.nn.functional.batch_norm sees it: the tensors plus `batch_stats` (normalise with the statistics
    of this batch), `update_running` (write the running statistics back with `momentum`) and `eps`."""
    __slots__ = ("weight", "bias", "running_mean", "running_var", "batch_stats", "update_running", "momentum", "eps")

    def __init__(self, weight, bias, running_mean, running_var, batch_stats, update_running, momentum, eps):
        self.weight, self.bias, self.running_mean, self.running_var = weight, bias, running_mean, running_var
        self.batch_stats, self.update_running, self.momentum, self.eps = batch_stats, update_running, momentum, eps

    @classmethod
    def from_module(cls, bn: nn.BatchNorm2d):
        """What nn.BatchNorm2d.forward hands to F.batch_norm (torch/nn/modules/batchnorm.py), side effect included:
        num_batches_tracked is incremented when the running statistics are going to be updated."""
        has_running = bn.track_running_stats and bn.running_mean is not None
        batch_stats = bn.training or not has_running
        update = bn.training and has_running
        momentum = 0.0
        if update:
            with torch.no_grad():
                bn.num_batches_tracked += 1
            momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        return cls(bn.weight, bn.bias, bn.running_mean if has_running else None, bn.running_var if has_running else None,
                   batch_stats, update, momentum, bn.eps)


def edge_conv(x: torch.Tensor, weight: torch.Tensor, k: int = 20, *, bias: torch.Tensor | None = None,
              bn: nn.BatchNorm2d | None = None, negative_slope: float | None = 0.2, idx: torch.Tensor | None = None):
    """max_j act(bn(conv1x1(get_graph_feature(x, args, k))))  ->  (B,O,N), without the (B,2C,N,k) tensor.

    x (B,C,N) or (B,C,N,1) float32 CUDA; weight (O,2C) or (O,2C,1,1); bias (O) or None; bn: the layer's
    nn.BatchNorm2d (training mode: batch statistics over all edges and running-statistics update exactly like
    torch.nn.functional.batch_norm; eval mode: running statistics) or None; negative_slope: LeakyReLU slope
    (PointDA uses 0.2), 0.0 = ReLU, None = no activation (PointSegDA's shared layers).  idx (B,N,k) int64 overrides the
    kNN graph (default: knn(x, k), PointDA/model_utils.py:9-16, same bits as the reference's ranking).
    Differentiable w.r.t. x, weight, bias and the BatchNorm affine parameters."""
    return edge_conv_functional(x, weight, k, bias, BNParams.from_module(bn) if bn is not None else None, negative_slope, idx)


def edge_conv_functional(x, weight, k, bias, bnp: BNParams | None, negative_slope, idx=None):
    """edge_conv with BatchNorm given as tensors (BNParams) instead of a module -- the form mlsp_b200.lazy needs when it
    meets F.batch_norm on a deferred graph feature."""
    _require_cuda_f32(x, "edge_conv")
    B, N = x.size(0), x.size(2)
    x = x.reshape(B, -1, N)
    C = x.size(1)
    O = weight.shape[0]
    if O % 4 != 0 or O > 1024:
        raise MlspError(f"edge_conv: output channels must be a multiple of 4 and <= 1024, got {O}")
    if idx is None:
        idx = knn(x, int(k))
    elif idx.shape != (B, N, k) or idx.dtype != torch.int64 or idx.device != x.device:
        raise MlspError("edge_conv: idx must be int64 (B,N,k) on x's device")
    else:
        from .ops import _check_idx
        _check_idx(idx, N, "edge_conv")
    slope = 1.0 if negative_slope is None else float(negative_slope)
    if slope < 0.0:
        raise MlspError("edge_conv: the activation must be non-decreasing (negative_slope >= 0)")
    idx = idx.contiguous()
    train = bnp is not None and bnp.batch_stats
    # per-channel affine map after the convolution: batch-statistics BatchNorm -> (gamma, beta), running-statistics
    # BatchNorm -> a = gamma / sqrt(running_var + eps), c = beta - a * running_mean, no BatchNorm -> (1, 0).
    if bnp is None:
        p0 = p1 = None
    elif train:
        p0, p1 = bnp.weight, bnp.bias
    else:
        invstd = torch.rsqrt(bnp.running_var + bnp.eps)
        p0 = invstd * bnp.weight if bnp.weight is not None else invstd
        p1 = -bnp.running_mean * p0
        if bnp.bias is not None:
            p1 = p1 + bnp.bias
    # a negative scale turns the max over k into a min: its sign is folded into the rows of W (and the bias) by the weight
    # prep kernel, so that the kernels always take a max and see a non-negative scale:  p0 * h = |p0| * (sign(p0) * h)
    yz, sgn = _PointwiseYZ.apply(x, weight, bias, p0)                 # (B,N,2O) = [Y | Z]: the layer's one GEMM
    if p0 is not None:
        p0 = p0 * sgn
    else:
        sgn = None
    if not train:
        return _EdgeConvReduce.apply(yz, idx, p0, p1, "affine", 0.0, slope)
    running = None
    if bnp.update_running:
        rm, rv = bnp.running_mean, bnp.running_var
        if rm.dtype != torch.float32 or not rm.is_contiguous() or not rv.is_contiguous():
            raise MlspError("edge_conv: BatchNorm running statistics must be contiguous float32")
        running = (rm, rv, sgn, float(bnp.momentum))                 # updated in place by the coefficient kernel
    return _EdgeConvReduce.apply(yz, idx, p0, p1, "bn_train", bnp.eps, slope, running)


def fold_convs(layers):
    """A stack of linear 1x1 convolutions [(W (O,I[,1,1]), b or None), ...] applied in order -> one (W, b)."""
    W, b = layers[0][0].flatten(1), layers[0][1]
    for Wn, bn_ in layers[1:]:
        Wn = Wn.flatten(1)
        b = Wn @ b if b is not None else None
        if bn_ is not None:
            b = bn_ if b is None else b + bn_
        W = Wn @ W
    return W, b


class FusedEdgeConv(nn.Module):
    """One DGCNN layer (graph feature -> 1x1 convs [-> BatchNorm2d] -> activation -> max over k) that SHARES the
    parameters of the reference modules it is built from, so a trained / freshly initialised reference model can be
    switched over in place and its optimiser keeps working.

        FusedEdgeConv.from_reference(model.conv2, k=20)                    # PointDA conv_2d (Conv2d, BatchNorm2d, LeakyReLU)
        FusedEdgeConv.from_reference([sl.conv1, sl.conv2], k=20)           # PointSegDA: plain Conv2d stack, no activation
    """

    def __init__(self, convs, bn=None, negative_slope=None, k: int = 20, register: bool = True):
        """register=False (what from_reference uses): the Conv2d / BatchNorm2d modules stay owned by the model they come
        from -- they are not submodules of this layer, so the model's state_dict keys, parameter list, .train() / .to()
        are exactly what they were (reference checkpoints load with strict=True; the optimiser sees each parameter once)."""
        super().__init__()
        convs = list(convs)
        for c in convs:
            if not isinstance(c, nn.Conv2d) or c.kernel_size != (1, 1) or c.groups != 1:
                raise MlspError("FusedEdgeConv: only 1x1 ungrouped Conv2d layers can be fused")
        if register:
            self.convs = nn.ModuleList(convs)
            self.bn = bn
        else:
            object.__setattr__(self, "convs", tuple(convs))
            object.__setattr__(self, "bn", bn)
        self.negative_slope = negative_slope
        self.k = k

    @classmethod
    def from_reference(cls, module, k: int = 20):
        mods = list(module) if isinstance(module, (list, tuple)) else [module]
        flat = []
        for m in mods:
            inner = getattr(m, "conv", m)                   # conv_2d wraps an nn.Sequential called .conv
            flat.extend(list(inner) if isinstance(inner, nn.Sequential) else [inner])
        convs, bn, slope = [], None, None
        for m in flat:
            if isinstance(m, nn.Conv2d):
                if bn is not None or slope is not None:
                    raise MlspError("FusedEdgeConv: a convolution after BatchNorm / activation cannot be folded")
                convs.append(m)
            elif isinstance(m, nn.BatchNorm2d):
                bn = m
            elif isinstance(m, nn.LeakyReLU):
                slope = m.negative_slope
            elif isinstance(m, nn.ReLU):
                slope = 0.0
            else:
                raise MlspError(f"FusedEdgeConv: cannot fuse {type(m).__name__}")
        return cls(convs, bn, slope, k, register=False)

    def effective_weight_bias(self):
        """The stack of linear 1x1 convolutions as one (O, 2C) matrix and bias (autograd reaches every layer's parameters)."""
        return fold_convs([(c.weight, c.bias) for c in self.convs])

    def forward(self, x, idx=None):
        W, b = self.effective_weight_bias()
        return edge_conv(x, W, self.k, bias=b, bn=self.bn, negative_slope=self.negative_slope, idx=idx)


def dgcnn_backbone(model, x: torch.Tensor, k: int = 20) -> torch.Tensor:
    """The four EdgeConv layers of the reference's PointDA DGCNN (PointDA/Models.py:114-130: conv1..conv4 on the
    transformed cloud) through `edge_conv`; returns x_cat (B, 64+64+128+256, N) like `torch.cat((x1,x2,x3,x4), 1)`.
    `model` is the reference's DGCNN instance (its conv_2d modules are used in place)."""
    outs = []
    for name in ("conv1", "conv2", "conv3", "conv4"):
        layer = FusedEdgeConv.from_reference(getattr(model, name), k=k)
        x = layer(x)
        outs.append(x)
    return torch.cat(outs, dim=1)
