import os, sys, types, warnings
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import mlsp_b200 as M
from mlsp_b200 import dgcnn, pcm, synth
dev = torch.device("cuda:0")
B, N = 32, 1024
targs = types.SimpleNamespace(mixup_params=1.0, DefRec_weight=0.5)
torch.manual_seed(0)
model = dgcnn.DGCNN(dropout=0.5).to(dev).train()
model.Rec_scan.requires_grad_(False)
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=5e-5)
crit = torch.nn.CrossEntropyLoss()
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
src = synth.surface_clouds(B, N, 1).permute(0, 2, 1).contiguous().to(dev)
trg = synth.surface_clouds(B, N, 2).permute(0, 2, 1).contiguous().to(dev)
lab = (torch.arange(B) % 10).to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    tb = trg.clone()
    pending = M.deform_input_begin(tb.permute(0, 2, 1))
    mixed, vals = pcm.mix_shapes(targs, src.permute(0, 2, 1), lab)
    pcm.calc_loss(targs, model(mixed), vals, crit).backward()
    dgcnn.target_branch_loss(model, tb, lookup, pending=pending).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
import traceback
def showwarning(message, category, filename, lineno, file=None, line=None):
    print("SYNC:", str(message)[:80])
    for l in traceback.format_stack()[-9:-2]:
        if "/root/repo" in l: print("   ", l.strip().splitlines()[0][:150])
warnings.showwarning = showwarning
warnings.simplefilter("always")
step()
torch.cuda.set_sync_debug_mode("default")
