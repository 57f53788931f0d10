#!/usr/bin/env python
"""cProfile of the host side of one workload-T training step (where the eager Python time goes)."""
import cProfile, os, pstats, sys, types, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import dgcnn, pcm, synth
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
B, N = 32, 1024
targs = types.SimpleNamespace(mixup_params=1.0, DefRec_weight=0.5)
torch.manual_seed(0)
model = dgcnn.DGCNN(dropout=0.5).to(dev).train()
model.Rec_scan.requires_grad_(False)
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=5e-5, fused=True)
crit = torch.nn.CrossEntropyLoss()
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
src = synth.surface_clouds(B, N, 1).permute(0, 2, 1).contiguous().to(dev)
trg = synth.surface_clouds(B, N, 2).permute(0, 2, 1).contiguous().to(dev)
lab = (torch.arange(B) % 10).to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    mixed, vals = pcm.mix_shapes(targs, src.permute(0, 2, 1), lab)
    pcm.calc_loss(targs, model(mixed), vals, crit).backward()
    dgcnn.target_branch_loss(model, trg.clone(), lookup).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter()          # host enqueue time only (no sync)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/5:.2f} ms/step, with final sync {1e3*(t2-t0)/5:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
