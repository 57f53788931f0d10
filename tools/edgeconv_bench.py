"""Times one EdgeConv layer forward+backward on the GPU three ways (CUDA events, after warm-up):
  fused     : mlsp_b200.edgeconv.edge_conv (no edge tensor)
  drop-in   : mlsp_b200.get_graph_feature (fused knn+gather kernels) -> torch Conv2d/BatchNorm2d/LeakyReLU/max
  torch     : oracle.ref_torch.get_graph_feature (the reference's op composition on torch's CUDA kernels) -> same torch layers
and, for the fused path, each stage alone.  python tools/edgeconv_bench.py [--B 32 --N 1024 --k 20 --reps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import edgeconv, synth  # noqa: E402


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def graph_timed(fn, reps):
    """device time of fn's kernels: captured once in a CUDA graph, replayed reps times between two events"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    return timed(gr.replay, reps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--torch-ref", action="store_true")
    ap.add_argument("--fused-only", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.enabled = False           # like the reference's trainers (PointDA/trainer.py:132-134)
    rows = []
    for C, O in ((3, 64), (64, 64), (64, 128), (128, 256)):
        x0 = (synth.clouds(a.B, a.N, 1) if C == 3 else synth.features(a.B, C, a.N, 2)).to(dev)
        seq = torch.nn.Sequential(torch.nn.Conv2d(2 * C, O, 1, bias=False), torch.nn.BatchNorm2d(O), torch.nn.LeakyReLU(0.2)).to(dev)
        g = torch.randn(a.B, O, a.N, device=dev)
        layer = edgeconv.FusedEdgeConv.from_reference(seq, k=a.k)

        def fused():
            x = x0.detach().requires_grad_(True)
            layer(x).backward(g)

        def dropin():
            x = x0.detach().requires_grad_(True)
            seq(M.get_graph_feature(x, None, k=a.k)).max(dim=-1)[0].backward(g)

        row = {"C": C, "O": O, "fused_ms": timed(fused, a.reps), "fused_graph_ms": graph_timed(fused, a.reps)}
        if a.fused_only:
            print(json.dumps(row), flush=True)
            continue
        row["dropin_ms"] = timed(dropin, a.reps)
        if a.torch_ref:
            from oracle import ref_torch

            def ref():
                x = x0.detach().requires_grad_(True)
                seq(ref_torch.get_graph_feature(x, a.k)).max(dim=-1)[0].backward(g)
            row["torch_ms"] = timed(ref, max(a.reps // 4, 2))
        # stages of the fused path
        idx = M.knn(x0, a.k)
        row["knn_ms"] = timed(lambda: M.knn(x0, a.k), a.reps)
        W, _ = layer.effective_weight_bias()
        Wcat = edgeconv._split_weight(W.detach(), C)
        row["gemm_ms"] = timed(lambda: torch.matmul(x0.transpose(1, 2), Wcat.t()), a.reps)
        xg = x0.detach().requires_grad_(True)
        with torch.no_grad():
            row["fwd_given_idx_ms"] = timed(lambda: layer(x0, idx=idx), a.reps)
        out = layer(xg, idx=idx)
        row["bwd_ms"] = timed(lambda: out.backward(g, retain_graph=True), a.reps)
        rows.append({k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in row.items()})
        print(json.dumps(rows[-1]), flush=True)
    if a.fused_only:
        return
    tot = {k_: round(sum(r[k_] for r in rows), 4) for k_ in rows[0] if k_.endswith("_ms")}
    print(json.dumps({"B": a.B, "N": a.N, "k": a.k, "sum_over_layers": tot}))


if __name__ == "__main__":
    main()
