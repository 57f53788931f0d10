#!/usr/bin/env python
"""ncu driver for the tcgen05 filter kernel alone (stage 2 of mlsp_graph_feature_fwd_stage on a prepared workspace).
   ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_filter python tools/prof_filter.py [B N k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M  # noqa: E402
from mlsp_b200 import synth  # noqa: E402
from mlsp_b200.ops import GraphFeatureStages  # noqa: E402

B, N, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else synth.CONFIGS["A"]
dev = torch.device("cuda:0")
hs = [GraphFeatureStages(synth.smooth_features(B, C, N, 1244 + C).to(dev), k) for C in (64, 128)]
for h in hs:
    h.run(2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for h in hs:
    h.run(1)
    h.run(2)
    h.run(4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
