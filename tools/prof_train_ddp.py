#!/usr/bin/env python
"""Host cost of DistributedDataParallel in the workload-T step: the same step with and without the wrapper in one process
(torchrun, 2 ranks), host enqueue time per step and a cProfile of the wrapped variant on rank 0.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/prof_train_ddp.py"""
import contextlib, cProfile, io, os, pstats, sys, time, types
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import dgcnn, pcm, synth
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, N = 32, 1024
targs = types.SimpleNamespace(mixup_params=1.0, DefRec_weight=0.5)
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
src = synth.surface_clouds(B, N, 1 + rank).permute(0, 2, 1).contiguous().to(dev)
trg = synth.surface_clouds(B, N, 2 + rank).permute(0, 2, 1).contiguous().to(dev)
lab = (torch.arange(B) % 10).to(dev)
crit = torch.nn.CrossEntropyLoss()
import gc


def build(wrap, **kw):
    torch.manual_seed(0)
    model = dgcnn.DGCNN(dropout=0.5).to(dev).train()
    model.Rec_scan.requires_grad_(False)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[lr], gradient_as_bucket_view=True, broadcast_buffers=False, **kw) if wrap else model
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=5e-5, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        tb = trg.clone()
        pending = M.deform_input_begin(tb.permute(0, 2, 1))
        with (net.no_sync() if wrap else contextlib.nullcontext()):
            mixed, vals = pcm.mix_shapes(targs, src.permute(0, 2, 1), lab)
            pcm.calc_loss(targs, net(mixed), vals, crit).backward()
        dgcnn.target_branch_loss(net, tb, lookup, pending=pending).backward()
        opt.step()
    return step


for name, wrap, kw in (("no wrapper", False, {}), ("DDP", True, {}), ("no wrapper", False, {})):
    step = build(wrap, **kw)
    gc.collect(); gc.freeze()
    for _ in range(5): step()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        print(f"{name:18s}: host enqueue {1e3*(t1-t0)/10:.2f} ms/step, with final sync {1e3*(t2-t0)/10:.2f} ms/step", flush=True)
    if name == "DDP" and rank == 0:
        pr = cProfile.Profile(); pr.enable()
    if name == "DDP":
        for _ in range(5): step()
        if rank == 0:
            pr.disable()
            s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500], flush=True)
    dist.barrier(); torch.cuda.synchronize()
    gc.unfreeze()
dist.destroy_process_group()
