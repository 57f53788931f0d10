"""mlsp_gemm_f32 (tcgen05 GEMM on three bf16 pieces per fp32 operand) against fp64 ground truth, through the C ABI.
Reference call sites it replaces: nn.Conv2d(kernel 1) in conv_2d PointDA/model_utils.py:45-63, nn.Conv1d(kernel 1) of the heads
PointDA/Models.py:165-285, nn.Linear of fc_layer model_utils.py:65-89 -- fp32 products, tolerance 1e-5 relative (north_star)."""
import ctypes

import numpy as np
import pytest
import torch

needs_gpu = pytest.mark.gpu


def _rel_err(got, ref64, a64, b64, bias64=None):
    """max |got - ref| relative to the magnitude sum_k |a||b| (+ |bias|) of each entry (the scale fp32 rounding acts on)."""
    scale = (a64.abs() @ b64.abs().transpose(-1, -2)).clamp_min(1e-30)
    if bias64 is not None:
        scale = scale + bias64.abs()
    return float(((got.double() - ref64).abs() / scale).max())


def test_gemm_argument_errors_without_gpu():
    from mlsp_b200 import build, _lib
    build.build()
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.mlsp_gemm_f32(None, 1, 8, 0, p, 1, 8, 0, p, 1, 8, 0, None, 4, 4, 8, 1, None) == 1 and b"null" in lib.mlsp_last_error()
    assert lib.mlsp_gemm_f32(p, 1, 8, 0, p, 1, 8, 0, p, 1, 8, 0, None, 4, 4, 0, 1, None) == 1            # K = 0
    assert lib.mlsp_gemm_f32(p, 1, 4, 0, p, 1, 8, 0, p, 1, 8, 0, None, 4, 4, 8, 1, None) == 1            # lda < K
    assert lib.mlsp_gemm_f32(p, 1, 8, 0, p, 1, 8, 0, p, 0, 2, 0, None, 4, 4, 8, 1, None) == 1            # ldd < M (column-major D)
    assert lib.mlsp_gemm_f32(p, 1, 8, 0, p, 1, 8, 0, p, 1, 8, 0, None, 0, 4, 8, 1, None) == 0            # empty product: nothing to do


def test_linear_refuses_cpu_tensors():
    from mlsp_b200 import linear, MlspError
    with pytest.raises(MlspError):
        linear.gemm_nt(torch.zeros(4, 8), torch.zeros(4, 8))


@needs_gpu
@pytest.mark.parametrize("akm", [1, 0])
@pytest.mark.parametrize("bkm", [1, 0])
@pytest.mark.parametrize("drm", [1, 0])
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 128, 128), (300, 200, 136), (1024, 512, 128), (77, 19, 3), (130, 64, 6),
                                   (5000, 300, 200), (9999, 129, 70)])       # the last two run as CTA pairs (cta_group::2)
def test_gemm_all_layouts_vs_fp64(akm, bkm, drm, shape):
    from mlsp_b200 import linear
    M, N, K = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 31 + N * 7 + K + 4 * akm + 2 * bkm + drm)
    a = (torch.randn(M, K, generator=g) * torch.exp(2 * torch.randn(M, 1, generator=g))).to(dev)
    b = torch.randn(N, K, generator=g).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    av = a if akm else a.t().contiguous().t()
    bv = b if bkm else b.t().contiguous().t()
    out = torch.full((M, N), float("nan"), device=dev) if drm else torch.full((N, M), float("nan"), device=dev).t()
    linear.gemm_nt(av, bv, bias, out=out)
    ref = a.double() @ b.double().t() + bias.double()
    err = _rel_err(out, ref, a.double(), b.double(), bias.double())
    assert err < 2e-6, err
    assert torch.isfinite(out).all()


@needs_gpu
def test_gemm_batched_strided_views():
    """The layouts the model uses: x (B,C,N) feature maps as M-major operands, shared weights, column-major output; a
    channel slice of a wider tensor as an operand (leading dimension > extent, batch stride > matrix)."""
    from mlsp_b200 import linear
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    B, C, N, O = 5, 96, 333, 72
    wide = torch.randn(B, C + 32, N, generator=g).to(dev)
    x = wide[:, 16:16 + C]                                              # (B,C,N) view, batch stride (C+32)*N
    W = torch.randn(O, C, generator=g).to(dev)
    y = linear.gemm_nt(x.transpose(1, 2), W, out_colmajor=True).transpose(1, 2)
    assert y.is_contiguous() and y.shape == (B, O, N)
    ref = torch.matmul(W.double(), x.double())
    scale = torch.matmul(W.double().abs(), x.double().abs())
    assert float(((y.double() - ref).abs() / scale).max()) < 2e-6
    part = linear.gemm_nt(y, x)                                         # (B,O,C): K = N points, both K-major
    refp = torch.matmul(ref, x.double().transpose(1, 2))
    sc = torch.matmul(ref.abs(), x.double().abs().transpose(1, 2))
    assert float(((part.double() - refp).abs() / sc).max()) < 4e-6


@needs_gpu
def test_gemm_many_tiles_persistent_loop():
    """More tiles than SMs and several K chunks: every CTA loops over tiles, both accumulator buffers and smem stages wrap."""
    from mlsp_b200 import linear
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(6)
    a = torch.randn(32768, 320, generator=g).to(dev)
    b = torch.randn(512, 320, generator=g).to(dev)
    out = linear.gemm_nt(a, b)
    ref = a.double() @ b.double().t()
    assert _rel_err(out, ref, a.double(), b.double()) < 2e-6


@needs_gpu
def test_conv1x1_and_linear_autograd_match_torch_fp64():
    from mlsp_b200 import linear
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    B, C, N, O, k = 3, 64, 256, 128, 20
    x = torch.randn(B, C, N, generator=g).to(dev).requires_grad_(True)
    W = (0.2 * torch.randn(O, C, 1, generator=g)).to(dev).requires_grad_(True)
    bias = torch.randn(O, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(B, O, N, generator=g).to(dev)
    y = linear.conv1x1(x, W, bias)
    y.backward(gy)
    xd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (x, W, bias))
    yd = torch.nn.functional.conv1d(xd, Wd, bd)
    yd.backward(gy.double())
    for got, ref in ((y, yd), (x.grad, xd.grad), (W.grad, Wd.grad), (bias.grad, bd.grad)):
        assert float((got.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), (got.shape,)
    # channels-last edge tensor (B,C,N,k) -> (B,O,N,k), the transform net's convolutions
    e = torch.randn(B, N, k, C, generator=g).to(dev).permute(0, 3, 1, 2).requires_grad_(True)
    W2 = (0.2 * torch.randn(O, C, 1, 1, generator=g)).to(dev).requires_grad_(True)
    ge = torch.randn(B, N, k, O, generator=g).to(dev).permute(0, 3, 1, 2)
    y2 = linear.conv1x1(e, W2)
    y2.backward(ge)
    ed, W2d = e.detach().double().requires_grad_(True), W2.detach().double().requires_grad_(True)
    y2d = torch.nn.functional.conv2d(ed, W2d)
    y2d.backward(ge.double())
    for got, ref in ((y2, y2d), (e.grad, ed.grad), (W2.grad, W2d.grad)):
        assert float((got.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), (got.shape,)


@needs_gpu
def test_pointwise_yz_node_matches_autograd_of_the_same_expression():
    """The EdgeConv layer's GEMM node (weight split + sign fold + bias + mlsp_gemm_f32) with its hand-written backward: grad_x
    straight in (B,C,N), weight gradient from B partial products, un-split -- against fp64 autograd of the same expression."""
    from mlsp_b200 import edgeconv
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    C, O = 5, 8
    for use_b, use_s in ((True, True), (False, True), (True, False), (False, False)):
        x = torch.randn(3, C, 8, device=dev, requires_grad=True)
        Wf = torch.randn(O, 2 * C, 1, 1, device=dev, requires_grad=True)
        bf = torch.randn(O, device=dev, requires_grad=True) if use_b else None
        sg = torch.tensor([1.0, -1.0] * (O // 2), device=dev) if use_s else None
        gy = torch.randn(3, 8, 2 * O, device=dev)
        y, s_out = edgeconv._PointwiseYZ.apply(x, Wf, bf, 0.7 * sg if use_s else None)    # scale: only its sign matters
        assert torch.equal(s_out, sg if use_s else torch.ones(O, device=dev))
        y.backward(gy)
        got = (x.grad.clone(), Wf.grad.clone(), bf.grad.clone() if use_b else None)
        xd, Wd = x.detach().double().requires_grad_(True), Wf.detach().double().requires_grad_(True)
        bd = bf.detach().double().requires_grad_(True) if use_b else None
        Wc = edgeconv._split_weight(Wd, C)
        if use_s:
            Wc = Wc * sg.double().repeat(2).view(2 * O, 1)
        ref = torch.matmul(xd.transpose(1, 2), Wc.t())
        if use_b:
            zb = bd * sg.double() if use_s else bd
            ref = ref + torch.cat((torch.zeros_like(zb), zb)).view(1, 1, 2 * O)
        ref.backward(gy.double())
        assert torch.allclose(y.double(), ref, atol=1e-5) and got[0].is_contiguous()
        assert torch.allclose(got[0].double(), xd.grad, atol=1e-5) and torch.allclose(got[1].double(), Wd.grad, atol=1e-5)
        assert not use_b or torch.allclose(got[2].double(), bd.grad, atol=1e-5)


@needs_gpu
def test_max_pooling_kernels_match_torch():
    """pool.cu against torch.max / adaptive_max_pool1d: values, and the gradient on the first maximal element -- on the three
    layouts the model uses, with ties (quantised values), a NaN, K smaller than the 8 cooperating warps, ragged channel blocks."""
    from mlsp_b200 import pool
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for B, C, N, k in ((3, 128, 50, 20), (2, 64, 33, 5), (2, 132, 17, 40)):
        e = torch.randint(-3, 4, (B, N, k, C), generator=g).float().to(dev).permute(0, 3, 1, 2).requires_grad_(True)
        out = pool.max_over_neighbours(e)
        ref, ridx = e.detach().max(dim=-1)
        assert torch.equal(out, ref) and out.shape == (B, C, N)
        go = torch.randn(B, C, N, generator=g).to(dev)
        out.backward(go)
        # every gradient value lands on ONE maximal element of its (b,c,n) row: the first
        first = (e.detach() == ref.unsqueeze(-1)).float().argmax(dim=-1)
        want = torch.zeros_like(e.detach()).scatter_(3, first.unsqueeze(-1), go.unsqueeze(-1))
        assert torch.equal(e.grad, want)
    x = torch.randn(4, 256, 300, 1, generator=g).to(dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
    out = pool.max_over_points_cl(x)
    assert torch.equal(out, torch.max(x.detach(), dim=2)[0])
    out.sum().backward()
    assert float(x.grad.sum()) == 4 * 256 and torch.equal((x.grad != 0), x.detach() == out.detach().unsqueeze(2))
    y = torch.randn(3, 40, 1000, generator=g).to(dev).requires_grad_(True)
    out = pool.global_max_pool(y)
    assert torch.equal(out, torch.nn.functional.adaptive_max_pool1d(y.detach(), 1))
    out.backward(torch.ones_like(out))
    assert torch.equal(y.grad != 0, y.detach() == out.detach())
    yn = y.detach().clone()
    yn[1, 2, 77] = float("nan")
    assert torch.isnan(pool.global_max_pool(yn)[1, 2, 0]) and torch.isfinite(pool.global_max_pool(yn)[0]).all()
    for K in (1, 3, 7, 37):                                              # unaligned rows take the scalar path
        z = torch.randn(5, 7, K, generator=g).to(dev)
        assert torch.equal(pool.global_max_pool(z), z.max(dim=2, keepdim=True)[0])


@needs_gpu
def test_linear_on_few_rows_splits_k():
    """nn.Linear on B rows (M = 32, K up to 1024): the K-split batched form agrees with fp64 like the plain form, forward and backward."""
    from mlsp_b200 import linear
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    for R, C, O in ((32, 1024, 512), (32, 512, 256), (32, 256, 9), (4, 1024, 1024), (32, 192, 64)):
        x = torch.randn(R, C, generator=g).to(dev).requires_grad_(True)
        W = (0.1 * torch.randn(O, C, generator=g)).to(dev).requires_grad_(True)
        b = torch.randn(O, generator=g).to(dev).requires_grad_(True)
        gy = torch.randn(R, O, generator=g).to(dev)
        y = linear.linear(x, W, b)
        y.backward(gy)
        xd, Wd, bd = (t.detach().double().requires_grad_(True) for t in (x, W, b))
        yd = torch.nn.functional.linear(xd, Wd, bd)
        yd.backward(gy.double())
        for got, ref in ((y, yd), (x.grad, xd.grad), (W.grad, Wd.grad), (b.grad, bd.grad)):
            assert float((got.detach().double() - ref.detach()).abs().max()) <= 1e-5 * float(ref.detach().abs().max()), (R, C, O, got.shape)


@needs_gpu
def test_bn_act_matches_torch():
    """bn.cu against torch's BatchNorm (training mode) + LeakyReLU / ReLU in fp64: output, running statistics, num_batches_tracked,
    and the gradients of the input, weight and bias -- on the four layouts the models use (channels-last 4-D edge features,
    contiguous (B,C,N) maps, channels-last (B,C,N), 2-D fc inputs), ragged sizes and an odd channel count (NCL only)."""
    import copy
    from mlsp_b200 import bn as mbn
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)

    def make(kind, B, C, N, k):
        if kind == "nhwc":
            return (torch.randn(B, N, k, C, generator=g) * 2 + 0.7).to(dev).permute(0, 3, 1, 2), torch.nn.BatchNorm2d(C)
        if kind == "ncl":
            return (torch.randn(B, C, N, generator=g) * 3 - 1.5).to(dev), torch.nn.BatchNorm1d(C)
        if kind == "nlc":
            return (torch.randn(B, N, C, generator=g) + 5.0).to(dev).permute(0, 2, 1), torch.nn.BatchNorm1d(C)
        return (torch.randn(B, C, generator=g) * 0.5 + 100.0).to(dev), torch.nn.BatchNorm1d(C)       # large mean: cancellation case

    cases = [("nhwc", 3, 64, 50, 20, 0.2), ("nhwc", 2, 128, 33, 5, 0.2), ("ncl", 4, 256, 300, 1, 0.0), ("ncl", 3, 37, 101, 1, 0.0),
             ("nlc", 2, 1024, 77, 1, 0.2), ("fc", 32, 512, 1, 1, 0.2), ("fc", 5, 192, 1, 1, 1.0), ("ncl", 2, 16, 64, 1, 1.0)]
    for kind, B, C, N, k, slope in cases:
        x0, mod = make(kind, B, C, N, k)
        mod = mod.to(dev).train()
        with torch.no_grad():
            mod.weight.copy_(torch.randn(C, generator=g).to(dev))         # negative scales included
            mod.bias.copy_(torch.randn(C, generator=g).to(dev))
            mod.running_mean.copy_(torch.randn(C, generator=g).to(dev))
        ref = copy.deepcopy(mod).double()
        x = x0.clone().requires_grad_(True)
        y = mbn.bn_act(x, mod, slope)
        xr = x0.detach().double().requires_grad_(True)
        yr = torch.nn.functional.leaky_relu(ref(xr), slope) if slope != 1.0 else ref(xr)
        assert y.shape == yr.shape and y.stride() == x.stride(), (kind, y.stride(), x.stride())
        scale = float(yr.abs().max())
        assert float((y.double() - yr).abs().max()) <= 2e-5 * scale, (kind, C, float((y.double() - yr).abs().max()), scale)
        go = torch.randn(y.shape, generator=g).to(dev)
        y.backward(go)
        yr.backward(go.double())
        for name, a, b in (("dx", x.grad, xr.grad), ("dgamma", mod.weight.grad, ref.weight.grad), ("dbeta", mod.bias.grad, ref.bias.grad)):
            err = float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-30)
            assert err <= 1e-4, (kind, C, name, err)     # act'(z) flips for |z| ~ 1e-7: a handful of elements move dgamma by < 1e-4
        assert torch.allclose(mod.running_mean.double(), ref.running_mean, rtol=1e-5, atol=1e-6), kind
        assert torch.allclose(mod.running_var.double(), ref.running_var, rtol=1e-5, atol=1e-6), kind
        assert int(mod.num_batches_tracked) == int(ref.num_batches_tracked) == 1
    # a channel slice of a wider map (torch.split of the heads' merged first layer): x keeps its batch stride, y / dx are contiguous
    wide = torch.randn(3, 96, 50, generator=g).to(dev).requires_grad_(True)
    parts = torch.split(wide, [32, 64], dim=1)
    mods = [torch.nn.BatchNorm1d(32).to(dev).train(), torch.nn.BatchNorm1d(64).to(dev).train()]
    refs = [copy.deepcopy(m).double() for m in mods]
    out = torch.cat([mbn.bn_act(p, m, 0.0) for p, m in zip(parts, mods)], dim=1)
    wr = wide.detach().double().requires_grad_(True)
    outr = torch.cat([torch.relu(m(p)) for p, m in zip(torch.split(wr, [32, 64], dim=1), refs)], dim=1)
    assert float((out.double() - outr).abs().max()) <= 2e-5 * float(outr.abs().max())
    go = torch.randn(out.shape, generator=g).to(dev)
    out.backward(go)
    outr.backward(go.double())
    assert float((wide.grad.double() - wr.grad).abs().max()) <= 1e-4 * float(wr.grad.abs().max())
    assert torch.allclose(mods[1].running_var.double(), refs[1].running_var, rtol=1e-5, atol=1e-6)
    # eval mode and CPU tensors run the torch modules unchanged
    x0, mod = make("ncl", 2, 64, 40, 1)
    mod = mod.to(dev).eval()
    assert torch.equal(mbn.bn_act(x0, mod, 0.0), torch.relu(mod(x0)))
