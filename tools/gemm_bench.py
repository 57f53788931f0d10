#!/usr/bin/env python
"""mlsp_gemm_f32 against torch.matmul (fp32, TF32 off: cuBLAS SIMT SGEMM) on the products of the PointDA training step."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsp_b200 import linear
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
HBM = 6554.0


def t(fn, n=20):
    """device time per call: n calls captured in one CUDA graph (no host launch overhead in the number), replayed 5 times"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n)


B, N = 32, 1024
cases = []
# EdgeConv layer products: yz (B,N,2O) = x^T (B,N,C) Wcat^T
for C, O in ((64, 64), (64, 128), (128, 256)):
    x = torch.randn(B, C, N, device=dev)
    W = torch.randn(2 * O, C, device=dev)
    cases.append((f"yz C={C} 2O={2*O}", lambda x=x, W=W: linear.gemm_nt(x.transpose(1, 2), W),
                  lambda x=x, W=W: torch.matmul(x.transpose(1, 2), W.t()), B * N * (C + 2 * O) * 4, 2.0 * B * N * C * 2 * O))
    dyz = torch.randn(B, N, 2 * O, device=dev)
    cases.append((f"gx  C={C} 2O={2*O}", lambda d=dyz, W=W: linear.gemm_nt(d, W.t(), out_colmajor=True),
                  lambda d=dyz, W=W: torch.matmul(W.t(), d.transpose(1, 2)), B * N * (C + 2 * O) * 4, 2.0 * B * N * C * 2 * O))
    cases.append((f"gw  C={C} 2O={2*O}", lambda d=dyz, x=x: linear.gemm_nt(d.transpose(1, 2), x),
                  lambda d=dyz, x=x: torch.bmm(d.transpose(1, 2), x.transpose(1, 2)), B * N * (C + 2 * O) * 4, 2.0 * B * N * C * 2 * O))
# conv5 / heads first layer: (B,512,N) -> (B,1024,N)
x = torch.randn(B, 512, N, device=dev)
W = torch.randn(1024, 512, device=dev)
cases.append(("conv5 512->1024", lambda: linear.gemm_nt(x.transpose(1, 2), W, out_colmajor=True), lambda: torch.matmul(W, x),
              B * N * (512 + 1024) * 4, 2.0 * B * N * 512 * 1024))
# transform net conv on the edge tensor: (B*N*k, 64) -> 128
e = torch.randn(B * N * 20, 64, device=dev)
W2 = torch.randn(128, 64, device=dev)
cases.append(("tnet 64->128 edges", lambda: linear.gemm_nt(e, W2), lambda: torch.matmul(e, W2.t()), B * N * 20 * (64 + 128) * 4,
              2.0 * B * N * 20 * 64 * 128))
ge = torch.randn(B * N * 20, 128, device=dev)
cases.append(("tnet wgrad 128x64", lambda: linear._reduce_rows_product(ge, e), lambda: torch.matmul(ge.t(), e), B * N * 20 * (64 + 128) * 4,
              2.0 * B * N * 20 * 64 * 128))
print(f"{'case':24s} {'ours ms':>9s} {'torch ms':>9s} {'speedup':>8s} {'GB/s':>8s} {'hbm frac':>8s} {'TFLOP/s':>8s}")
for name, ours, ref, nbytes, flops in cases:
    a, b = t(ours), t(ref)
    print(f"{name:24s} {a:9.4f} {b:9.4f} {b/a:8.2f} {nbytes/a/1e6:8.0f} {nbytes/a/1e6/HBM:8.3f} {flops/a/1e9:8.1f}")
