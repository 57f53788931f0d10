"""oracle/ref_real.py -- TEST INFRASTRUCTURE: the hot-path step composed from the reference's OWN functions.

Imports the unmodified reference modules staged by oracle/make_ref.py under oracle/_ref/src (third-party imports the hot
path never reaches are stubbed under oracle/_ref/stubs).  This is what `bench.py --impl reference` and the
`cpu_baseline` leg time on the host cores (kind "reference"); `available()` is False when nothing was staged (then
bench.py falls back to the port oracle/ref_torch.py and says so).

Which reference function runs each op of the step (bench.py's hotpath-A / S op list):
    FPS                      utils.pc_utils.farthest_point_sample          (utils/pc_utils.py:137-161)
    knn + edge gather        PointDA.model_utils.get_graph_feature / knn   (PointDA/model_utils.py:9-42)
    deform_input             MLSP.mlsp.deform_input                        (MLSP/mlsp.py:10-51)
    position loss fwd + bwd  MLSP.mlsp.reconstruction_loss                 (MLSP/mlsp.py:156-182), autograd backward
    cardinality, normals     python-pcl in the reference (mlsp.py:240-272, PointDA/trainer.py:173-188): NOT runnable
                             anywhere -> the dense-torch restatements of oracle/ref_torch.py, labelled "port" in the sample
"""
from __future__ import annotations

import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "_ref", "src")
_STUBS = os.path.join(_HERE, "_ref", "stubs")
_mods = None


def available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "MANIFEST.json"))


def load():
    """-> (pc_utils, mlsp, model_utils) of the staged reference."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("oracle/_ref is empty: run `python -m oracle.make_ref` in the build container")
        for p in (_STUBS, os.path.join(_SRC, "PointDA"), _SRC):     # PointDA/Models.py imports top-level `model_utils`
            if p not in sys.path:
                sys.path.insert(0, p)
        import utils.pc_utils as pc_utils
        from MLSP import mlsp
        from PointDA import model_utils
        _mods = (pc_utils, mlsp, model_utils)
    return _mods


def ref_args(device):
    a = types.SimpleNamespace()
    a.cuda = device.type == "cuda"
    a.gpus = [device.index if device.type == "cuda" else -1]
    return a


def hot_path_step(clouds, feats, grads, pred, lookup, k=20, radius=0.13, num_cls=16, near=20, fps_split=(512, 512),
                  pergroup=2, shift=0):
    """Same op list and argument convention as oracle.ref_torch.hot_path_step, every op but the two python-pcl ones
    executed by the reference's own function."""
    from oracle import ref_torch
    pcu, mlsp, mu = load()
    dev = clouds.device
    args = ref_args(dev)
    pts = clouds.permute(0, 2, 1).contiguous()
    for n in fps_split:
        pcu.farthest_point_sample(args, clouds, n)
    ref_torch.normals_dense(pts, near)                              # python-pcl in the reference: restatement
    ref_torch.cal_density_dense(pts, radius, num_cls, pergroup, shift)
    gold = clouds.clone()
    X, mask = mlsp.deform_input(clouds.clone(), lookup, "volume_based_voxels", dev)
    for f, g in zip(feats, grads):
        f = f.clone().requires_grad_(True)
        out = mu.get_graph_feature(f, args, k=k)
        if g is not None:
            out.backward(g)
    p = pred.clone().requires_grad_(True)
    loss = mlsp.reconstruction_loss(p, gold, mask)
    loss.backward()
    return float(loss.detach())
