#!/usr/bin/env python
"""Per-iteration phase timeline of CTA 0 of mlsp_gemm_f32 (measurement hook mlsp_gemm_f32_timeline): where a K-chunk iteration's
time goes -- loaders' convert phase, barrier hops, MMA issue."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsp_b200 import _lib
dev = torch.device("cuda:0")


def run(name, a, b, akm, bkm, M, N, K, Z, lda, ldb, sa, sb):
    out = torch.empty(Z, M, N, device=dev)
    ts = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        _lib.call("mlsp_gemm_f32_timeline", a.data_ptr(), akm, lda, sa, b.data_ptr(), bkm, ldb, sb, out.data_ptr(), 1, N, M * N, None, M, N, K, Z,
                  ts.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t = ts.view(64, 8).cpu().numpy().astype(float)
    n = min(40, (K + 63) // 64 * ((Z + 147) // 148)) - 1
    sel = range(8, n)
    ghz = 1.965
    free_to_stored = [t[i, 1] - t[i, 0] for i in sel]
    stored_to_arr = [t[i, 2] - t[i, 1] for i in sel]
    arr_to_full = [t[i, 3] - t[i, 2] for i in sel]
    full_to_issued = [t[i, 4] - t[i, 3] for i in sel]
    issued_to_free = [t[i + 2, 0] - t[i, 4] for i in sel if i + 2 < 64]           # MMA execution + commit -> loaders of the same slot
    period = [t[i + 1, 0] - t[i, 0] for i in sel]
    wait_next = [t[i + 1, 0] - t[i, 2] for i in sel]                               # loaders idle between arrive and next stage free
    f = lambda v: sum(v) / len(v) / ghz
    print(f"{name}: period {f(period):6.0f} ns | loaders: convert {f(free_to_stored):5.0f}, fence+arrive {f(stored_to_arr):4.0f}, idle until next stage free "
          f"{f(wait_next):5.0f} | arrive->MMA sees full {f(arr_to_full):5.0f} | MMA issue {f(full_to_issued):4.0f} | issued -> same slot free again {f(issued_to_free):5.0f}")


Z, K = 148, 2560
a = torch.randn(Z, 128, K, device=dev); b = torch.randn(Z, 128, K, device=dev)
run("K-major 128x128 HBM  ", a, b, 1, 1, 128, 128, K, Z, K, K, 128 * K, 128 * K)
a = torch.randn(16, 128, K, device=dev); b = torch.randn(16, 128, K, device=dev)
run("K-major 128x128 L2   ", a, b, 1, 1, 128, 128, K, 16, K, K, 128 * K, 128 * K)
a = torch.randn(16, K, 128, device=dev); b = torch.randn(16, K, 128, device=dev)
run("MN-major 128x128 L2  ", a, b, 0, 0, 128, 128, K, 16, 128, 128, 128 * K, 128 * K)
a = torch.randn(16, 64, K, device=dev); b = torch.randn(16, 64, K, device=dev)
run("K-major 64x64 L2     ", a, b, 1, 1, 64, 64, K, 16, K, K, 64 * K, 64 * K)
