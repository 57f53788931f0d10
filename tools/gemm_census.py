#!/usr/bin/env python
"""Which products one workload-T training step sends to mlsp_gemm_f32: shape census + device time per distinct shape."""
import collections, os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import dgcnn, pcm, synth, _lib
dev = torch.device("cuda:0")
B, N = 32, 1024
targs = types.SimpleNamespace(mixup_params=1.0, DefRec_weight=0.5)
torch.manual_seed(0)
model = dgcnn.DGCNN(dropout=0.5).to(dev).train()
model.Rec_scan.requires_grad_(False)
crit = torch.nn.CrossEntropyLoss()
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
src = synth.surface_clouds(B, N, 1).permute(0, 2, 1).contiguous().to(dev)
trg = synth.surface_clouds(B, N, 2).permute(0, 2, 1).contiguous().to(dev)
lab = (torch.arange(B) % 10).to(dev)
def step():
    mixed, vals = pcm.mix_shapes(targs, src.permute(0, 2, 1), lab)
    pcm.calc_loss(targs, model(mixed), vals, crit).backward()
    dgcnn.target_branch_loss(model, trg.clone(), lookup).backward()
step()
torch.cuda.synchronize()
log = []
orig = _lib.call
def spy(name, *args):
    if name == "mlsp_gemm_f32":
        log.append(args)
    return orig(name, *args)
_lib.call = spy
step()
torch.cuda.synchronize()
_lib.call = orig
cnt = collections.Counter()
keep = {}
for a in log:
    key = (a[13], a[14], a[15], a[16], a[1], a[5], a[9])       # M N K Z akm bkm drm
    cnt[key] += 1
    keep[key] = a
rows = []
for key, n in cnt.items():
    a = keep[key]
    for _ in range(3): orig("mlsp_gemm_f32", *a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): orig("mlsp_gemm_f32", *a)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    rows.append((n * us, n, us, key))
tot = sum(r[0] for r in rows)
print(f"{len(log)} calls per step, {tot/1e3:.2f} ms summed (back-to-back launches, warm)")
for t, n, us, key in sorted(rows, reverse=True)[:40]:
    M_, N_, K_, Z_, akm, bkm, drm = key
    gf = 2.0 * M_ * N_ * K_ * Z_ / 1e9
    print(f"{t/tot*100:5.1f}%  n={n:2d}  {us:7.1f} us  M={M_:6d} N={N_:5d} K={K_:6d} Z={Z_:3d} a{'K' if akm else 'M'} b{'K' if bkm else 'N'} d{'R' if drm else 'C'}  {gf/us*1e3:7.1f} TFLOP/s")
