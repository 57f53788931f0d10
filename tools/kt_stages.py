import os, sys, torch
sys.path.insert(0, '/root/repo')
import mlsp_b200 as M
from mlsp_b200 import synth
dev = torch.device("cuda:0")
def timed(fn, reps=20):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for name, B, N, k, C in [("A", 32, 1024, 20, 64), ("A", 32, 1024, 20, 128), ("X/4", 64, 4096, 40, 64), ("X/4", 64, 4096, 40, 128), ("X/4k20", 64, 4096, 20, 128)]:
    x = (synth.features(B, C, N, 5) if name.startswith("X") else synth.smooth_features(B, C, N, 1244 + C)).to(dev)
    h = M.ops.GraphFeatureStages(x, k)
    out = [f"{name} C={C} k={k}:"]
    for st in ("2", "3", "4", "6", "8"):
        os.environ["MLSP_KT_STAGES"] = st
        for mode in ("0", "3"):
            os.environ["MLSP_KT_MODE"] = mode
            try:
                out.append(f"st{st}/m{mode} {timed(lambda: h.run(2)):7.1f}")
            except Exception as e:
                out.append(f"st{st}/m{mode} n/a")
    os.environ.pop("MLSP_KT_STAGES"); os.environ["MLSP_KT_MODE"] = "0"
    print(" ".join(out), flush=True)
