"""Sum of the GPU kernel time of one workload-T training step (torch.profiler / CUPTI): the floor a graph-captured step could reach."""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import dgcnn, pcm, synth
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev) / 3 / 1e3
print(f"GPU kernel time per step: {tot:.3f} ms over {sum(e.count for e in ev)/3:.0f} kernels")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:25]:
    print(f"{e.device_time_total/3/1e3:8.3f} ms  {e.count/3:6.1f}x  {e.key[:100]}")
