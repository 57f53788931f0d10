"""tools/step_trace.py -- where a hot-path step's time goes between the host and the three streams.

Runs bench.GraphedStep of workload A or S for K steps in two host regimes -- free-running (the bench's resident region) and
throttled (the host waits for the end of step i-1 after enqueuing step i, what the e2e region's loss read-back does) -- and
prints ms/step for each, plus GPU-event spans of the three graphs (gA model layers, gT FPS + structure, gB layer 1 + loss)
measured in a separate pass with the graphs replayed alone.

    python tools/step_trace.py [A|S] [steps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "S"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    bench.set_workload(wl)
    B, N, k = synth.CONFIGS[wl]
    device = torch.device("cuda", 0)
    host, dev = bench.make_inputs(B, N, k, 1234, device, pin=True)
    lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=device)
    np.random.seed(1234)
    torch.manual_seed(1234)
    streams = bench.Streams(device)
    off = bench.OpTimer(False)
    with torch.cuda.stream(streams.model):
        for _ in range(2):
            bench.gpu_step(M, dev, lookup, k, off, streams)
        g = bench.GraphedStep(M, dev, lookup, k, streams)
        for _ in range(5):
            g(off)
        torch.cuda.synchronize()

        def run(throttle, from_host):
            ends = [torch.cuda.Event() for _ in range(K)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(K):
                g(off, host["clouds"] if from_host else None)
                ends[i].record()
                if throttle and i:
                    ends[i - 1].synchronize()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / K * 1e3

        for name, th, fh in (("free-running, resident", False, False), ("throttled, resident", True, False),
                             ("free-running, host clouds", False, True), ("throttled, host clouds", True, True)):
            ms = [run(th, fh) for _ in range(3)]
            print(f"{wl} {name:28s} ms/step {min(ms):.4f} (runs: {' '.join(f'{m:.4f}' for m in ms)})", flush=True)
        # the graphs alone
        for nm, gr in (("gA", g.gA), ("gT", g.gT), ("gB", g.gB)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                gr.replay()
            e0.record()
            for _ in range(10):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
            print(f"{wl} graph {nm} alone: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)
        # deform_input alone, wall clock (host-bound)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            X = dev["clouds"].clone()
            M.deform_input(X, lookup, "volume_based_voxels", device)
        torch.cuda.synchronize()
        print(f"{wl} deform_input alone (wall): {(time.perf_counter() - t0) / 20 * 1e3:.4f} ms", flush=True)


if __name__ == "__main__":
    main()


def overlap_probe():
    """gA and gT replayed together on two streams (no deform_input, no host work in between): do the graphs overlap at all?"""
    wl = sys.argv[1] if len(sys.argv) > 1 else "S"
    bench.set_workload(wl)
    B, N, k = synth.CONFIGS[wl]
    device = torch.device("cuda", 0)
    host, dev = bench.make_inputs(B, N, k, 1234, device, pin=True)
    lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=device)
    for side in (False, True):
        streams = bench.Streams(device, side_model=side)
        off = bench.OpTimer(False)
        with torch.cuda.stream(streams.model):
            for _ in range(2):
                bench.gpu_step(M, dev, lookup, k, off, streams)
            g = bench.GraphedStep(M, dev, lookup, k, streams)
            for _ in range(3):
                g(off)
            torch.cuda.synchronize()
            sm, sa = streams.model, streams.aux
            for label, both in (("gA then gT on one stream", False), ("gA || gT on two streams", True)):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(sm)
                for _ in range(10):
                    start = torch.cuda.Event()
                    start.record(sm)
                    g.gA.replay()
                    if both:
                        with torch.cuda.stream(sa):
                            sa.wait_event(start)
                            g.gT.replay()
                            done = torch.cuda.Event()
                            done.record(sa)
                        sm.wait_event(done)
                    else:
                        g.gT.replay()
                e1.record(sm)
                torch.cuda.synchronize()
                print(f"{wl} model stream {'side' if side else 'default'}: {label}: {e0.elapsed_time(e1) / 10:.4f} ms", flush=True)


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "overlap":
    overlap_probe()


def concurrency_probe():
    """When does gT actually run if it is launched beside gA?  Event offsets relative to the start of gA."""
    wl = sys.argv[1] if len(sys.argv) > 1 else "A"
    bench.set_workload(wl)
    B, N, k = synth.CONFIGS[wl]
    device = torch.device("cuda", 0)
    host, dev = bench.make_inputs(B, N, k, 1234, device, pin=True)
    lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=device)
    streams = bench.Streams(device)
    off = bench.OpTimer(False)
    with torch.cuda.stream(streams.model):
        for _ in range(2):
            bench.gpu_step(M, dev, lookup, k, off, streams)
        g = bench.GraphedStep(M, dev, lookup, k, streams)
        for _ in range(3):
            g(off)
        torch.cuda.synchronize()
        sm, sa = streams.model, streams.aux
        start_dev = g.start_dev[0]

        def fps_only():
            M.fps_from_start(g.clouds, bench.FPS_SPLIT[0], start_dev)

        from mlsp_b200 import _lib

        def fps_groups(n):
            def f():
                _lib.load().mlsp_fps_set_groups(n)
                fps_only()
                _lib.load().mlsp_fps_set_groups(0)
            return f

        for label, side_work in (("gT graph", g.gT.replay), ("one eager FPS call (auto)", fps_only),
                                 ("one eager FPS call, 1 cloud/CTA", fps_groups(1)), ("one eager FPS call, 2 clouds/CTA", fps_groups(2)),
                                 ("one eager FPS call, 4 clouds/CTA", fps_groups(4))):
            for rep in range(3):
                ev = {n: torch.cuda.Event(enable_timing=True) for n in ("a0", "a1", "t0", "t1")}
                torch.cuda.synchronize()
                ev["a0"].record(sm)
                with torch.cuda.stream(sa):
                    sa.wait_event(ev["a0"])
                    ev["t0"].record(sa)
                    side_work()
                    ev["t1"].record(sa)
                g.gA.replay()
                ev["a1"].record(sm)
                torch.cuda.synchronize()
                print(f"{wl} {label}: gA 0 .. {ev['a0'].elapsed_time(ev['a1']):.3f} ms;  side work "
                      f"{ev['a0'].elapsed_time(ev['t0']):.3f} .. {ev['a0'].elapsed_time(ev['t1']):.3f} ms", flush=True)


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "conc":
    concurrency_probe()


def timeline():
    """Per-step event offsets (ms after the step's `ready` event) of the three graphs and deform_input, plus host times."""
    wl = sys.argv[1] if len(sys.argv) > 1 else "S"
    bench.set_workload(wl)
    B, N, k = synth.CONFIGS[wl]
    device = torch.device("cuda", 0)
    host, dev = bench.make_inputs(B, N, k, 1234, device, pin=True)
    lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=device)
    np.random.seed(1234)
    streams = bench.Streams(device)
    off = bench.OpTimer(False)
    with torch.cuda.stream(streams.model):
        for _ in range(2):
            bench.gpu_step(M, dev, lookup, k, off, streams)
        g = bench.GraphedStep(M, dev, lookup, k, streams)
        for _ in range(5):
            g(off)
        torch.cuda.synchronize()
        g.trace = []
        t_host0 = time.perf_counter()
        hosts = []
        for _ in range(8):
            hosts.append(time.perf_counter())
            g(off)
        torch.cuda.synchronize()
        tr = g.trace
        base = tr[0]["ready"]
        for i, t in enumerate(tr):
            r = base.elapsed_time(t["ready"])
            print(f"{wl} step {i}: ready @{r:7.3f} | gA +{t['ready'].elapsed_time(t['gA']):.3f} gT +{t['ready'].elapsed_time(t['gT']):.3f} "
                  f"deformed +{t['ready'].elapsed_time(t['deformed']):.3f} gB +{t['ready'].elapsed_time(t['gB']):.3f} | host: step begins "
                  f"{(hosts[i] - t_host0) * 1e3:.3f}, deform returned {(t['host_deform_done'] - t_host0) * 1e3:.3f}, all enqueued "
                  f"{(t['host_done'] - t_host0) * 1e3:.3f}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "timeline":
    timeline()
