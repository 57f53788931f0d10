"""tools/bn_bench.py -- the fused BatchNorm + activation kernels (csrc/bn.cu) on the training step's shapes: device time per
call (CUDA graph replay, 10 launches between two events), algorithmic bytes (3 passes forward, 5 backward) against the
measured HBM peak, and torch's BatchNorm + activation modules on the same tensors."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsp_b200 import bn as mbn  # noqa: E402

dev = torch.device("cuda:0")
peak = 6554.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbps"]
except Exception:
    pass


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


shapes = [("transform net conv2d1 (B,64,N,k) channels-last", (32, 1024, 20, 64), "nhwc", 0.2),
          ("transform net conv2d2 (B,128,N,k) channels-last", (32, 1024, 20, 128), "nhwc", 0.2),
          ("conv5 (B,1024,N)", (32, 1024, 1024), "ncl", 0.2),
          ("head bn1 (B,256,N) slice of (B,1024,N)", (32, 1024, 1024), "slice256", 0.0),
          ("head bn2 (B,256,N)", (32, 256, 1024), "ncl", 0.0),
          ("fc (B,512)", (32, 512), "fc", 0.2)]
for name, shp, kind, slope in shapes:
    if kind == "nhwc":
        x = torch.randn(*shp, device=dev).permute(0, 3, 1, 2)
        mod = torch.nn.BatchNorm2d(shp[3])
    elif kind == "slice256":
        x = torch.randn(*shp, device=dev)[:, 256:512]
        mod = torch.nn.BatchNorm1d(256)
    else:
        x = torch.randn(*shp, device=dev)
        mod = torch.nn.BatchNorm1d(shp[1])
    mod = mod.to(dev).train()
    x = x.requires_grad_(True)
    nbytes = x.numel() * 4
    y = mbn.bn_act(x, mod, slope)
    go = torch.randn_like(y)
    f = timed(lambda: mbn.bn_act(x, mod, slope))
    fb = timed(lambda: mbn.bn_act(x, mod, slope).backward(go))
    act = (lambda t: t) if slope == 1.0 else (torch.relu if slope == 0.0 else (lambda t: torch.nn.functional.leaky_relu(t, slope)))
    tf = timed(lambda: act(mod(x)))
    tfb = timed(lambda: act(mod(x)).backward(go))
    print(f"{name:52s} {nbytes/1e6:7.1f} MB | fused fwd {f*1e3:7.1f} us ({3*nbytes/f/1e6:6.0f} GB/s = {3*nbytes/f/1e6/peak*100:4.1f}% of HBM peak)"
          f"  bwd {(fb-f)*1e3:7.1f} us ({5*nbytes/(fb-f)/1e6:6.0f} GB/s = {5*nbytes/(fb-f)/1e6/peak*100:4.1f}%) | torch fwd {tf*1e3:7.1f} us  bwd {(tfb-tf)*1e3:7.1f} us", flush=True)
