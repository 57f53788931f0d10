#!/usr/bin/env python
"""Host-side cost breakdown of deform_input (voxel mode): cProfile over 200 calls."""
import cProfile, pstats, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mlsp_b200 as M
from mlsp_b200 import synth
dev = torch.device("cuda:0")
clouds = synth.surface_clouds(32, 1024, 1234).to(dev)
lookup = torch.tensor(M.region_mean(3), dtype=torch.float32, device=dev)
for _ in range(10):
    M.deform_input(clouds.clone(), lookup, "volume_based_voxels", dev)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    M.deform_input(clouds.clone(), lookup, "volume_based_voxels", dev)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
