import os, sys, torch
sys.path.insert(0, "/root/repo")
from mlsp_b200 import linear
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3
# long-K, one small tile per batch: iterations per CTA = ceil(Z/148) * K/64
for Z, K in ((148, 2560), (148, 640), (16, 2560), (148, 10240), (592, 2560)):
    gy = torch.randn(Z, K, 64, device=dev); x = torch.randn(Z, K, 64, device=dev)
    us = t(lambda: linear.gemm_nt(gy.transpose(1, 2), x.transpose(1, 2)))
    iters = -(-Z // 148) * (K // 64)
    print(f"wgrad-like M=64 N=64 K={K} Z={Z}: {us:7.1f} us, {iters} chunk iterations per CTA -> {us/iters*1e3:6.0f} ns per iteration, A+B = {2*Z*K*64*4/1e6:.0f} MB")
# same but K-major operands (rows contiguous along K)
for Z, K in ((148, 2560), (16, 2560)):
    a = torch.randn(Z, 64, K, device=dev); b = torch.randn(Z, 64, K, device=dev)
    us = t(lambda: linear.gemm_nt(a, b))
    iters = -(-Z // 148) * (K // 64)
    print(f"K-major    M=64 N=64 K={K} Z={Z}: {us:7.1f} us -> {us/iters*1e3:6.0f} ns per iteration")
# full tiles 128x128
for Z, K in ((148, 2560), (16, 2560)):
    a = torch.randn(Z, 128, K, device=dev); b = torch.randn(Z, 128, K, device=dev)
    us = t(lambda: linear.gemm_nt(a, b))
    iters = -(-Z // 148) * (K // 64)
    print(f"K-major    M=128 N=128 K={K} Z={Z}: {us:7.1f} us -> {us/iters*1e3:6.0f} ns per iteration (MMA alone: 24 x 64 clk = 780 ns)")
