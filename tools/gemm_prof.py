#!/usr/bin/env python
"""One launch of mlsp_gemm_f32 per shape of interest, for ncu (--set full) captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsp_b200 import linear
dev = torch.device("cuda:0")
B, N = 32, 1024
x = torch.randn(B, 128, N, device=dev); W = torch.randn(512, 128, device=dev)
x5 = torch.randn(B, 512, N, device=dev); W5 = torch.randn(1024, 512, device=dev)
e = torch.randn(B * N * 20, 64, device=dev); W2 = torch.randn(128, 64, device=dev)
x64 = torch.randn(B, 64, N, device=dev); W64 = torch.randn(128, 64, device=dev)
dyz = torch.randn(B, N, 512, device=dev)
for _ in range(2):
    linear.gemm_nt(x.transpose(1, 2), W)                              # yz C=128 2O=512 (row-major out)
    linear.gemm_nt(x5.transpose(1, 2), W5, out_colmajor=True)         # conv5
    linear.gemm_nt(e, W2)                                             # transform net 64 -> 128 on the edge tensor
    linear.gemm_nt(x64.transpose(1, 2), W64)                          # yz C=64 2O=128
    linear.gemm_nt(dyz.transpose(1, 2), x)                            # weight gradient partials, K = 1024 per cloud
torch.cuda.synchronize()
