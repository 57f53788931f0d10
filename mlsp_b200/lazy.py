"""Deferred graph feature: the EdgeConv fusion of mlsp_b200.edgeconv behind the reference's UNCHANGED model code.

`mlsp_b200.patch.patch(fuse_edgeconv=True)` rebinds `get_graph_feature` to the function below.  It returns a
`LazyGraphFeature`: a tensor subclass with the right shape, dtype and device and no storage, which records what the
model does to it and decides at the last moment what has to be computed:

    x = get_graph_feature(x, self.args, k=self.k)     # nothing runs                        PointDA/Models.py:114
    x = self.conv1(x)                                 # Conv2d 1x1 -> BatchNorm2d -> LeakyReLU: recorded      :115
    x1 = x.max(dim=-1, keepdim=False)[0]              # edge_conv(...): no (B,2C,N,k) tensor ever exists      :116

PointSegDA's `conv2(conv1(graph_feature)).max(-1)` (plain Conv2d stacks, PointSegDA/Models.py:171-174) is folded the
same way.  Anything else -- the input-transform nets run two convolutions with a non-linearity in between
(PointDA/model_utils.py:111-114), user code may index or print the tensor -- materialises the feature with the fused
knn + gather kernels (ops.get_graph_feature), replays the recorded ops with torch's own functions and carries on, so
the result is always what the reference computes.  The interception point is `__torch_function__`; no class of the
reference or of torch is patched.
"""
from __future__ import annotations

import collections

import torch
import torch.nn.functional as F

from . import edgeconv, ops

MaxResult = collections.namedtuple("max", ["values", "indices"])

# the two computations behind the lazy tensor (module attributes so that the interception logic can be exercised with
# the CPU oracle in tests; the product always runs the CUDA ops)
_materialise_backend = lambda x, k, idx: ops.get_graph_feature(x, None, k=k, idx=idx)                    # noqa: E731
_fused_backend = lambda x, W, k, bias, bnp, slope, idx: edgeconv.edge_conv_functional(x, W, k, bias, bnp, slope, idx)  # noqa: E731

_T = torch.Tensor
_METADATA = {
    _T.dim, _T.ndimension, _T.size, _T.numel, _T.nelement, _T.get_device, _T.is_floating_point, _T.is_complex,
    _T.shape.__get__, _T.ndim.__get__, _T.dtype.__get__, _T.device.__get__, _T.is_cuda.__get__, _T.layout.__get__,
    _T.requires_grad.__get__, _T.is_leaf.__get__, _T.grad_fn.__get__, _T.is_sparse.__get__, _T.is_quantized.__get__,
    _T.is_meta.__get__, _T.__len__,
}
counters = collections.Counter()        # "fused" / "materialised": what the lazy features of this process turned into


class _Chain:
    """What has been applied to the graph feature so far: 1x1 convolutions, then at most one BatchNorm, then at most one
    activation."""
    __slots__ = ("convs", "bn", "slope")

    def __init__(self, convs=(), bn=None, slope=None):
        self.convs, self.bn, self.slope = tuple(convs), bn, slope


class LazyGraphFeature(torch.Tensor):
    @staticmethod
    def __new__(cls, x, k, idx, chain=None):
        chain = chain or _Chain()
        B, C, N = x.shape
        ch = chain.convs[-1][0].shape[0] if chain.convs else 2 * C
        r = torch.Tensor._make_wrapper_subclass(cls, (B, ch, N, int(k)), dtype=x.dtype, device=x.device, requires_grad=False)
        r._src = (x, int(k), idx)
        r._chain = chain
        return r

    def __repr__(self):
        c = self._chain
        return (f"LazyGraphFeature(shape={tuple(self.shape)}, convs={len(c.convs)}, bn={c.bn is not None}, "
                f"slope={c.slope}, device={self.device})")

    # ---- the two ways out
    def materialise(self) -> torch.Tensor:
        """The tensor the reference would hold at this point: fused knn + gather kernels, then the recorded ops."""
        x, k, idx = self._src
        counters["materialised"] += 1
        h = _materialise_backend(x, k, idx)
        for W, b in self._chain.convs:
            h = F.conv2d(h, W, b)
        bn = self._chain.bn
        if bn is not None:
            h = F.batch_norm(h, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.batch_stats, bn.momentum, bn.eps)
        if self._chain.slope is not None:
            h = F.leaky_relu(h, self._chain.slope) if self._chain.slope != 0.0 else F.relu(h)
        return h

    def fused_max(self) -> torch.Tensor:
        x, k, idx = self._src
        counters["fused"] += 1
        W, b = edgeconv.fold_convs(list(self._chain.convs))
        return _fused_backend(x, W, k, b, self._chain.bn, self._chain.slope, idx)

    def _with(self, **kw):
        c = self._chain
        return LazyGraphFeature(*self._src, _Chain(kw.get("convs", c.convs), kw.get("bn", c.bn), kw.get("slope", c.slope)))

    # ---- interception
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in _METADATA:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        self = args[0] if args and isinstance(args[0], LazyGraphFeature) else None
        if self is not None:
            c = self._chain
            out = None
            if func in (torch.conv2d, F.conv2d):
                out = self._try_conv(c, *args[1:], **kwargs)
            elif func is F.batch_norm:
                out = self._try_bn(c, *args[1:], **kwargs)
            elif func is F.leaky_relu:
                out = self._try_act(c, kwargs.get("negative_slope", args[1] if len(args) > 1 else 0.01))
            elif func in (F.relu, torch.relu, F.relu_, torch.relu_):
                out = self._try_act(c, 0.0)
            elif func in (_T.max, torch.max):
                out = self._try_max(c, *args[1:], **kwargs)
            if out is not None:
                return out
        # anything else: the reference's tensor, then the real op
        def real(a):
            if isinstance(a, LazyGraphFeature):
                return a.materialise()
            if isinstance(a, (list, tuple)):
                return type(a)(real(e) for e in a)
            return a
        rargs = [real(a) for a in args]
        rkw = {k_: real(v) for k_, v in kwargs.items()}
        return func(*rargs, **rkw)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # only operators that bypassed __torch_function__ (called from C++) get here: run them on the real tensor
        from torch.utils._pytree import tree_map
        real = lambda a: a.materialise() if isinstance(a, LazyGraphFeature) else a          # noqa: E731
        return func(*tree_map(real, args), **tree_map(real, kwargs or {}))

    def _try_conv(self, c, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
        def all_eq(v, want):
            return all(e == want for e in v) if isinstance(v, (tuple, list)) else v == want
        if (c.bn is not None or c.slope is not None or isinstance(weight, LazyGraphFeature) or weight.dim() != 4
                or weight.shape[2:] != (1, 1) or weight.shape[1] != self.shape[1] or groups != 1
                or not all_eq(stride, 1) or not all_eq(padding, 0) or not all_eq(dilation, 1)):
            return None
        return self._with(convs=c.convs + ((weight, bias),))

    def _try_bn(self, c, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
        if not c.convs or c.bn is not None or c.slope is not None:
            return None
        has_running = running_mean is not None and running_var is not None
        if not training and not has_running:
            return None
        bnp = edgeconv.BNParams(weight, bias, running_mean if has_running else None, running_var if has_running else None,
                                bool(training), bool(training) and has_running, float(momentum), float(eps))
        return self._with(bn=bnp)

    def _try_act(self, c, slope):
        if not c.convs or c.slope is not None or float(slope) < 0.0:
            return None
        return self._with(slope=float(slope))

    def _try_max(self, c, dim=None, keepdim=False):
        if not c.convs or dim is None or isinstance(dim, torch.Tensor) or keepdim or dim not in (-1, 3):
            return None
        O = c.convs[-1][0].shape[0]
        # the same support predicate as edge_conv_functional / the C ABI: anything else materialises the reference's tensor
        if O % 4 != 0 or O > 1024 or self.shape[3] > 64:
            return None
        return MaxResult(self.fused_max(), None)      # the reference takes [0]; the arg-max over k is not produced


def get_graph_feature(x: torch.Tensor, args=None, k: int = 20, idx: torch.Tensor | None = None):
    """get_graph_feature(x, args, k, idx) of PointDA/model_utils.py:18-42 == PointSegDA/Models.py:18-45, deferred: returns a
    LazyGraphFeature of shape (B,2C,N,k) (see the module docstring).  Same signature as ops.get_graph_feature."""
    ops._require_cuda_f32(x, "get_graph_feature")
    B, N = x.size(0), x.size(2)
    x = x.reshape(B, -1, N).contiguous()
    if idx is not None and (idx.shape != (B, N, k) or idx.dtype != torch.int64 or idx.device != x.device):
        raise ops.MlspError("get_graph_feature: idx must be int64 (B,N,k) on x's device")
    if not (1 <= int(k) <= N):
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")  # torch.topk's message
    return LazyGraphFeature(x, int(k), idx)
