"""oracle/gen_golden.py -- TEST INFRASTRUCTURE: writes tests/golden/*.npz.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden
Every array below is produced by calling the REFERENCE'S OWN function (imported through
oracle/ref_import.py) on seeded synthetic inputs; inputs are stored next to outputs so the
fixtures are self-contained on the GPU box, where the reference does not exist.

The reference has no tests or golden vectors of its own, so these are
the pins for a1-a5, a8-a10.  a6/a7 (python-pcl) cannot be run: no fixture, parity unpinned.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlsp_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _args():
    a = types.SimpleNamespace()
    a.cuda = False
    a.gpus = [-1]
    return a


def ref_pd(x):
    """The reference's ranking matrix, by evaluating the expression of knn() (PointSegDA/Models.py:9-11)
    with torch on this CPU -- stored so tie-invariant comparisons are possible (topk is not tie-stable)."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def main():
    torch.set_num_threads(1)          # fixed reduction order for the sgemm inside the reference
    os.makedirs(OUT, exist_ok=True)
    pcu, mlsp, seg = ref_import.load()
    args = _args()

    # ---- a1 knn: quantised (exact) and continuous sets, 3-D and feature space
    cases = {
        "knn_q3": (synth.clouds(2, 160, 11, quantised=True), 20),
        "knn_c3": (synth.clouds(2, 160, 12), 20),
        "knn_q64": (synth.features(2, 64, 96, 13, quantised=True), 20),
        "knn_c64": (synth.smooth_features(2, 64, 96, 14), 20),
        "knn_c128_k40": (synth.smooth_features(1, 128, 128, 15), 40),
        # k > 64 (the reference's args.k is a free flag): rounds of the exact kernel
        "knn_q3_k100": (synth.clouds(1, 200, 16, quantised=True), 100),
        "knn_c16_k130": (synth.smooth_features(1, 16, 160, 17), 130),
    }
    for name, (x, k) in cases.items():
        idx = seg.knn(x, k)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), k=k, idx=idx.numpy(),
                            pd=ref_pd(x).numpy())

    # ---- a2 get_graph_feature (idx supplied, so the fixture is independent of topk tie order)
    for name, x, k in (("ggf_3", synth.clouds(2, 128, 21), 20), ("ggf_16", synth.features(2, 16, 64, 22), 8)):
        idx = seg.knn(x, k)
        out = seg.get_graph_feature(x, args, k=k, idx=idx)
        assert out.stride() == (out.shape[2] * k * out.shape[1], 1, k * out.shape[1], out.shape[1])
        out_default = seg.get_graph_feature(x.view(*x.shape, 1), args, k=k)      # 4-D input path, knn inside
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), k=k, idx=idx.numpy(),
                            out=out.contiguous().numpy(), same_as_default=bool(torch.equal(out, out_default)))
    # backward through the reference graph feature (autograd index_put accumulate)
    x = synth.features(2, 8, 48, 23).requires_grad_(True)
    idx = seg.knn(x.detach(), 6)
    out = seg.get_graph_feature(x, args, k=6, idx=idx)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(g)
    np.savez_compressed(os.path.join(OUT, "ggf_bwd.npz"), x=x.detach().numpy(), idx=idx.numpy(),
                        g=g.numpy(), grad_x=x.grad.numpy())

    # ---- a3 FPS (start index comes from torch.randint on the CPU generator, pc_utils.py:150)
    xyz = synth.clouds(3, 512, 31)
    torch.manual_seed(1)
    cen, vals = pcu.farthest_point_sample(args, xyz, 128)
    torch.manual_seed(1)
    start = torch.randint(0, 512, (3,), dtype=torch.long)
    assert torch.equal(start, cen[:, 0])
    np.savez_compressed(os.path.join(OUT, "fps.npz"), xyz=xyz.numpy(), npoint=128, centroids=cen.numpy(),
                        vals=vals.numpy(), seed=1)
    xyzq = synth.clouds(2, 256, 32, quantised=True)     # duplicates / exact ties
    torch.manual_seed(2)
    cen, vals = pcu.farthest_point_sample(args, xyzq, 256)
    np.savez_compressed(os.path.join(OUT, "fps_q_full.npz"), xyz=xyzq.numpy(), npoint=256, centroids=cen.numpy(),
                        vals=vals.numpy(), seed=2)

    # ---- a4 voxel regions, including points exactly on faces / outside the cube
    X = synth.clouds(3, 1024, 41) * 1.2
    t1 = float(np.float32(-1 + 2 / 3))
    t2 = float(np.float32(-1 + 4 / 3))
    special = torch.tensor([[t1, 0.0, 0.0], [t2, 0.5, -0.5], [1.0, 0.2, 0.2], [-1.0, 0.0, 0.0], [1.5, 0.0, 0.0],
                            [0.9999999, 0.9999999, 0.9999999], [0.0, t1, 0.0], [0.0, 0.0, t2],
                            [float(np.nextafter(np.float32(t1), np.float32(1))), 0.0, 0.0],
                            [float(np.nextafter(np.float32(t1), np.float32(-1))), 0.0, 0.0]])
    X[0, :, : special.shape[0]] = special.t()
    Y = pcu.assign_region_to_point(X, "cpu")
    np.savez_compressed(os.path.join(OUT, "regions.npz"), X=X.numpy(), Y=Y.numpy(),
                        lookup=pcu.region_mean(3))

    # ---- a4/a8 deform_input, voxel mode (numpy RNG: one permutation + one multivariate_normal per cloud)
    lookup = torch.Tensor(pcu.region_mean(3))
    for seed in (1, 7):
        X0 = synth.surface_clouds(4, 1024, 50 + seed)
        X = X0.clone()
        np.random.seed(seed)
        Xd, mask = mlsp.deform_input(X, lookup, "volume_based_voxels", "cpu")
        assert Xd is X
        np.savez_compressed(os.path.join(OUT, f"deform_voxels_s{seed}.npz"), X0=X0.numpy(), X=Xd.numpy(),
                            mask=mask.numpy(), seed=seed)
    # groups=3: the walk over the permuted regions deforms three regions per cloud (mlsp.py:49)
    X0 = synth.surface_clouds(3, 1024, 58)
    X = X0.clone()
    np.random.seed(5)
    Xd, mask = mlsp.deform_input(X, lookup, "volume_based_voxels", "cpu", groups=3)
    np.savez_compressed(os.path.join(OUT, "deform_voxels_g3.npz"), X0=X0.numpy(), X=Xd.numpy(),
                        mask=mask.numpy(), seed=5, groups=3)
    # sparse cloud: no region reaches 40 points -> untouched, empty mask
    X0 = synth.clouds(2, 64, 59)
    X = X0.clone()
    np.random.seed(3)
    Xd, mask = mlsp.deform_input(X, lookup, "volume_based_voxels", "cpu")
    np.savez_compressed(os.path.join(OUT, "deform_voxels_sparse.npz"), X0=X0.numpy(), X=Xd.numpy(),
                        mask=mask.numpy(), seed=3)

    # ---- a5 deform_input, ball mode (collapse_to_point)
    X0 = synth.surface_clouds(3, 512, 61)
    X = X0.clone()
    np.random.seed(2)
    Xd, mask = mlsp.deform_input(X, lookup, "volume_based_radius", "cpu")
    np.savez_compressed(os.path.join(OUT, "deform_radius.npz"), X0=X0.numpy(), X=Xd.numpy(), mask=mask.numpy(),
                        seed=2)

    # ---- a9/a10 masked Chamfer
    gold = synth.surface_clouds(4, 768, 71)                                  # (B,3,N)
    X = gold.clone()
    np.random.seed(4)
    _, mask = mlsp.deform_input(X, lookup, "volume_based_voxels", "cpu")     # realistic region mask
    assert (mask.sum(dim=(1, 2)) > 0).all()
    g = torch.Generator().manual_seed(72)
    pred = (gold.permute(0, 2, 1) + 0.05 * torch.randn(4, 768, 3, generator=g)).contiguous().requires_grad_(True)
    loss = mlsp.reconstruction_loss(pred, gold, mask)
    loss.backward()
    i1 = mlsp.findneareat_index(pred.detach(), gold.permute(0, 2, 1), mask.permute(0, 2, 1))
    i2 = mlsp.findneareat_index(gold.permute(0, 2, 1), pred.detach(), mask.permute(0, 2, 1))
    np.savez_compressed(os.path.join(OUT, "chamfer.npz"), pred=pred.detach().numpy(), gold=gold.numpy(),
                        mask=mask.numpy(), loss=loss.item(), grad=pred.grad.numpy(), idx_pred_gold=i1.numpy(),
                        idx_gold_pred=i2.numpy())
    # far-away predictions: the finite +100 penalty lets unmasked columns win (mlsp.py:143-149)
    pred_far = (gold.permute(0, 2, 1) * 0.0 + 30.0 + torch.randn(4, 768, 3, generator=g)).contiguous().requires_grad_(True)
    loss = mlsp.reconstruction_loss(pred_far, gold, mask)
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "chamfer_far.npz"), pred=pred_far.detach().numpy(), gold=gold.numpy(),
                        mask=mask.numpy(), loss=loss.item(), grad=pred_far.grad.numpy())
    # empty mask -> NaN (0/0)
    loss = mlsp.reconstruction_loss(pred.detach(), gold, torch.zeros_like(mask))
    np.savez_compressed(os.path.join(OUT, "chamfer_empty.npz"), loss=loss.item())

    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {len(os.listdir(OUT))} fixtures, {tot/1e6:.2f} MB -> {OUT}")


if __name__ == "__main__":
    main()
