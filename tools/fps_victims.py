"""tools/fps_victims.py -- which hot-path kernels slow down while an FPS kernel is resident on other SMs?

Times each model-stream op alone and again while farthest-point sampling (32 CTAs of 256 threads, latency-bound) runs on a
second stream.  python tools/fps_victims.py [groups]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import _lib, synth  # noqa: E402


def main():
    groups = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    bench.set_workload("A")
    B, N, k = synth.CONFIGS["A"]
    device = torch.device("cuda", 0)
    host, dev = bench.make_inputs(B, N, k, 1234, device, pin=True)
    _lib.load().mlsp_fps_set_groups(groups)
    clouds = dev["clouds"]
    start = torch.zeros(B, dtype=torch.int64, device=device)
    side = torch.cuda.Stream(device=device)
    f64, f128 = dev["feats"][2], dev["feats"][4]
    ops = {}
    ops["knn3 + gather (C=3)"] = lambda: M.get_graph_feature(clouds, None, k=k)
    ops["knn only C=64 (centre+prep+filter+refine)"] = lambda: M.knn(f64, k)
    ops["ggf C=64"] = lambda: M.get_graph_feature(f64, None, k=k)
    ops["ggf C=128"] = lambda: M.get_graph_feature(f128, None, k=k)
    e64 = M.get_graph_feature(f64.detach().requires_grad_(True), None, k=k)
    g64 = dev["grads"][2]
    ops["edge_bwd C=64"] = lambda: torch.autograd.grad(e64, e64.grad_fn.next_functions[0][0].variable if False else None) if False else e64.backward(g64, retain_graph=True)
    pts = clouds.permute(0, 2, 1).contiguous()
    ops["target_structure"] = lambda: M.target_structure(pts, bench.NEAR, bench.RADIUS, bench.NUM_CLS, bench.PERGROUP, bench.SHIFT)
    reps = 10
    for name, fn in ops.items():
        res = []
        for with_fps in (False, True):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if with_fps:
                with torch.cuda.stream(side):
                    for _ in range(40):
                        M.fps_from_start(clouds, 512, start)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / reps)
        print(f"groups={groups} {name:45s} alone {res[0]*1e3:8.1f} us   beside FPS {res[1]*1e3:8.1f} us   x{res[1]/res[0]:.2f}", flush=True)


if __name__ == "__main__":
    main()
