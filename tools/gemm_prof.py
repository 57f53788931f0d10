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
for _ in range(2):
    linear.gemm_nt(x.transpose(1, 2), W)                              # yz C=128 2O=512 (row-major out)
    linear.gemm_nt(x5.transpose(1, 2), W5, out_colmajor=True)         # conv5
torch.cuda.synchronize()
