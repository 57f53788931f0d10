"""CPU: the oracle (C + numpy restatements) against fixtures produced by the reference's own
functions (oracle/gen_golden.py).  This is what pins the oracle before it is trusted as the
checker of the CUDA kernels."""
import numpy as np
import pytest

import oracle
from oracle import np_ops
from conftest import knn_rank_check


@pytest.mark.parametrize("name", ["knn_q3", "knn_q64", "knn_q3_k100"])
def test_knn_quantised_exact(golden, name):
    """Grid-quantised inputs: every fp32 evaluation order agrees, ties are real -> strict equality with
    the stable (lowest-index) ranking of the reference's pd matrix, and value-wise equality with topk."""
    g = golden(name)
    x, k, pd = g["x"], int(g["k"]), g["pd"]
    idx, pdo = oracle.knn(x, k, return_pd=True)
    stable = np.argsort(-pd, axis=2, kind="stable")[:, :, :k]
    assert np.array_equal(idx, stable)
    # same ranked VALUES as the reference's own topk output (tie-invariant statement)
    assert np.array_equal(np.take_along_axis(pd, idx, 2), np.take_along_axis(pd, g["idx"], 2))
    assert np.array_equal(pdo, np.take_along_axis(pd, idx, 2))
    # self is the nearest neighbour
    assert (np.take_along_axis(pd, idx[:, :, :1], 2)[..., 0] == pd.max(axis=2)).all()


@pytest.mark.parametrize("name", ["knn_c3", "knn_c64", "knn_c128_k40", "knn_c16_k130"])
def test_knn_continuous_certified(golden, name):
    g = golden(name)
    x, k = g["x"], int(g["k"])
    idx = oracle.knn(x, k)
    bad, unc = knn_rank_check(x, idx, k)
    assert unc == 0, f"{unc} uncertified mismatches"
    bad_ref, unc_ref = knn_rank_check(x, g["idx"], k)
    assert unc_ref == 0
    # the oracle is no further from fp64 truth than the reference's own sgemm is (plus slack)
    assert bad <= bad_ref + max(4, idx.size // 20000), (bad, bad_ref)
    # and it agrees with the reference everywhere except such near-ties
    assert (idx != g["idx"]).sum() <= bad + bad_ref


@pytest.mark.parametrize("name,layers", [("activations_da", ("x1", "x2", "x3")), ("activations_seg", ("x1", "x2"))])
def test_knn_on_real_backbone_activations(golden, name, layers):
    """The oracle against the reference's own knn on REAL activations (x1/x2/x3 of the seeded reference DGCNN at config-A
    size; PointSegDA's BatchNorm-free, bias-shifted layers): differences only at certified fp64 near-ties."""
    g = golden(name)
    for lay in layers:
        x, ref_idx = g[lay], g["idx_" + lay].astype(np.int64)
        idx = oracle.knn(x, 20)
        bad, unc = knn_rank_check(x, idx, 20)
        bad_ref, unc_ref = knn_rank_check(x, ref_idx, 20)
        # the oracle never leaves the 8-ulp band around fp64 truth; the reference's own sgemm-ordered ranking does on the
        # bias-shifted PointSegDA layers (|x|^2 >> the distances: cancellation in -xx - inner - xx^T), so only its COUNT is used
        assert unc == 0 and (name == "activations_seg" or unc_ref == 0), (lay, unc, unc_ref)
        assert (idx != ref_idx).sum() <= bad + bad_ref, lay


def test_scan_input(golden):
    """oracle/np_ops.scan_input (vectorised z-buffer) against the reference's own scan_input / p_scan (MLSP/mlsp.py:54-94):
    bit-identical clouds and masks for the same seeds, duplicates (ties inside a bin) included."""
    import random
    g = golden("scan_input")
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    out, mask = np_ops.scan_input(g["X"])
    assert np.array_equal(out, g["out"]) and np.array_equal(mask, g["mask"])
    assert (mask[3, :512, 0] == 0).sum() > 0 and (mask[3, 512:, 0] == 0).sum() == 0       # of two equal points the first wins


@pytest.mark.parametrize("name", ["ggf_3", "ggf_16"])
def test_edge_gather(golden, name):
    g = golden(name)
    out = oracle.edge_gather(g["x"], g["idx"])                       # (B,N,k,2C)
    assert np.array_equal(out.transpose(0, 3, 1, 2), g["out"])
    assert bool(g["same_as_default"])


def test_edge_gather_bwd(golden):
    g = golden("ggf_bwd")
    C = g["x"].shape[1]
    gcl = np.ascontiguousarray(g["g"].transpose(0, 2, 3, 1))           # (B,2C,N,k) -> (B,N,k,2C)
    gx = oracle.edge_gather_bwd(gcl, g["idx"], C)
    np.testing.assert_allclose(gx, g["grad_x"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["fps", "fps_q_full"])
def test_fps(golden, name):
    g = golden(name)
    cen, vals = oracle.fps(g["xyz"], int(g["npoint"]), g["centroids"][:, 0])
    assert np.array_equal(cen, g["centroids"])
    assert np.array_equal(vals, g["vals"])


def test_regions(golden):
    g = golden("regions")
    assert np.array_equal(np_ops.assign_region_to_point(g["X"]), g["Y"])
    np.testing.assert_allclose(np_ops.region_mean(3), g["lookup"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name,mode", [("deform_voxels_s1", "volume_based_voxels"),
                                       ("deform_voxels_s7", "volume_based_voxels"),
                                       ("deform_voxels_sparse", "volume_based_voxels"),
                                       ("deform_voxels_g3", "volume_based_voxels"),
                                       ("deform_radius", "volume_based_radius")])
def test_deform_input(golden, name, mode):
    g = golden(name)
    X = g["X0"].copy()
    np.random.seed(int(g["seed"]))
    Xd, mask = np_ops.deform_input(X, np_ops.region_mean(3), mode, groups=int(g["groups"]) if "groups" in g else 1)
    assert np.array_equal(mask, g["mask"])
    assert np.array_equal(Xd, g["X"])


def test_chamfer(golden):
    g = golden("chamfer")
    loss, grad = oracle.reconstruction_loss(g["pred"], g["gold"], g["mask"])
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(grad, g["grad"], rtol=1e-5, atol=1e-7)
    gold_bnc = g["gold"].transpose(0, 2, 1)
    m = g["mask"][:, 0, :]
    _, _, a1 = oracle.chamfer_dir(g["pred"], gold_bnc, m)
    _, _, a2 = oracle.chamfer_dir(gold_bnc, g["pred"], m)
    assert (a1 != g["idx_pred_gold"]).mean() < 1e-3
    assert (a2 != g["idx_gold_pred"]).mean() < 1e-3


def test_chamfer_far_and_empty(golden):
    g = golden("chamfer_far")
    loss, grad = oracle.reconstruction_loss(g["pred"], g["gold"], g["mask"])
    assert abs(loss - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(grad, g["grad"], rtol=1e-5, atol=1e-7)
    loss, _ = oracle.reconstruction_loss(g["pred"], g["gold"], np.zeros_like(g["mask"]))
    assert np.isnan(loss) and np.isnan(float(golden("chamfer_empty")["loss"]))


def test_density_labels_arithmetic():
    cnt = np.array([[0, 1, 2, 3, 29, 30, 31, 100]])
    soft, row = np_ops.density_labels(cnt, 16, pergroup=2, shift=0)
    assert row.tolist() == [[0, 1, 2, 3, 29, 30, 30, 30]]
    assert soft[0, 1, 0] == 0.5 and soft[0, 1, 1] == 0.5 and soft[0, 2, 1] == 1.0
    assert np.allclose(soft.sum(-1), 1.0)
    soft, row = np_ops.density_labels(cnt, 16, pergroup=5, shift=10)
    assert row.tolist() == [[0, 0, 0, 0, 19, 20, 21, 75]]


def test_density_count_semantics():
    """Restated pcl semantics: strict radius, K cap, neighbour index 0 dropped (mlsp.py:252-254)."""
    rng = np.random.default_rng(0)
    P = rng.uniform(-0.3, 0.3, (1, 300, 3)).astype(np.float32)
    r = 0.13
    cnt = oracle.density_count(P, r, K=100)[0]
    d = ((P[0][:, None, :].astype(np.float64) - P[0][None]) ** 2).sum(-1)
    full = (d < r * r).sum(1)
    in0 = d[:, 0] < r * r
    expect = np.minimum(full, 100) - in0
    near = np.abs(d - r * r).min(1) < 1e-6          # fp32-vs-fp64 boundary cases excluded
    assert np.array_equal(cnt[~near & (full <= 100)], expect[~near & (full <= 100)])
    assert cnt[0] == min(full[0], 100) - 1          # point 0 always drops itself
    cntK = oracle.density_count(P, 0.5, K=10)[0]
    assert cntK.max() <= 10 and cntK.min() >= 9


def test_radius_search_is_the_list_form_of_density_count():
    """orc_radius_search (the list pcl's radius_search_for_cloud returns) against brute force, and its
    `(ind != 0).sum(1)` -- the expression of MLSP/mlsp.py:252-253 -- against orc_density_count."""
    rng = np.random.default_rng(1)
    P = rng.uniform(-0.4, 0.4, (2, 400, 3)).astype(np.float32)
    diff = (P[:, :, None, :] - P[:, None, :, :]) ** 2                 # float32, the oracle's own expression
    d = (diff[..., 0] + diff[..., 1]) + diff[..., 2]
    for r, K in ((0.13, 100), (0.3, 20), (0.6, 128)):
        ind, sqd = oracle.radius_search(P, r, K)
        assert np.array_equal((ind != 0).sum(-1), oracle.density_count(P, r, K))
        r2 = np.float32(r * r)
        for b in range(2):
            for i in (0, 7, 399):
                js = np.nonzero(d[b, i] < r2)[0]
                js = js[np.lexsort((js, d[b, i, js]))][:K]
                assert np.array_equal(ind[b, i, :len(js)], js) and not ind[b, i, len(js):].any()
                assert np.array_equal(sqd[b, i, :len(js)], d[b, i, js]) and not sqd[b, i, len(js):].any()


def test_normals_plane():
    rng = np.random.default_rng(1)
    P = np.zeros((1, 400, 3), np.float32)
    P[0, :, :2] = rng.uniform(-1, 1, (400, 2))
    P[0, :, 2] = 0.25
    n = np_ops.pca_normals(P, 10)
    assert np.allclose(np.abs(n[0, :, 2]), 1.0, atol=1e-9)
    assert (n[0, :, 2] < 0).all()                   # flipped towards the origin: n.p <= 0


# ---- the pure-torch port that bench.py times on the CPU (oracle/ref_torch.py) ------------------------------
def test_ref_torch_port_matches_goldens(golden):
    import torch
    from oracle import ref_torch as rt
    g = golden("knn_q3")
    idx = rt.knn(torch.from_numpy(g["x"]), int(g["k"])).numpy()
    assert np.array_equal(np.take_along_axis(g["pd"], idx, 2), np.take_along_axis(g["pd"], g["idx"], 2))
    g = golden("ggf_3")
    out = rt.get_graph_feature(torch.from_numpy(g["x"]), k=int(g["k"]), idx=torch.from_numpy(g["idx"]))
    assert np.array_equal(out.numpy(), g["out"])
    g = golden("fps")
    torch.manual_seed(int(g["seed"]))
    cen, vals = rt.farthest_point_sample(torch.from_numpy(g["xyz"]), int(g["npoint"]))
    assert np.array_equal(cen.numpy(), g["centroids"]) and np.array_equal(vals.numpy(), g["vals"])
    g = golden("regions")
    assert np.array_equal(rt.assign_region_to_point(torch.from_numpy(g["X"])).numpy(), g["Y"])
    g = golden("deform_voxels_s1")
    np.random.seed(int(g["seed"]))
    X, mask = rt.deform_input(torch.from_numpy(g["X0"].copy()), torch.Tensor(np_ops.region_mean(3)))
    assert np.array_equal(X.numpy(), g["X"]) and np.array_equal(mask.numpy(), g["mask"])
    g = golden("chamfer")
    pred = torch.from_numpy(g["pred"]).requires_grad_(True)
    loss = rt.reconstruction_loss(pred, torch.from_numpy(g["gold"]), torch.from_numpy(g["mask"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    np.testing.assert_allclose(pred.grad.numpy(), g["grad"], rtol=1e-5, atol=1e-8)


def test_ref_torch_restatements_match_oracle():
    import torch
    from oracle import ref_torch as rt
    from mlsp_b200 import synth
    pts = synth.surface_clouds(2, 400, 5).permute(0, 2, 1).contiguous()
    lab, row = rt.cal_density_dense(pts, 0.13, 16)
    ol, orow = np_ops.cal_density(pts.numpy(), 0.13, 16)
    assert (row.numpy() != orow).mean() < 5e-3          # cdist rounding at the radius boundary only
    n = rt.normals_dense(pts, 20).numpy()
    on, gap = np_ops.pca_normals(pts.numpy(), 20, return_gap=True)
    ok = gap > 1e-2
    assert (1 - np.abs((n * on).sum(-1))[ok]).max() < 1e-3


# ---- the EdgeConv layer (reference composition restated in oracle/edgeconv_ref.py)
def _t(a):
    import torch
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", ["edgeconv_da_16_32", "edgeconv_da_3_64"])
def test_edgeconv_oracle_pointda_layer(golden, name):
    """oracle.edgeconv_ref.layer == the reference's conv_2d(get_graph_feature(x)).max(-1) (PointDA/Models.py:114-116):
    forward, every gradient, the running statistics after one training step, and the eval-mode forward."""
    import torch
    from oracle import edgeconv_ref
    torch.set_num_threads(1)
    g = golden(name)
    x = _t(g["x"]).requires_grad_(True)
    W = _t(g["weight"]).requires_grad_(True)
    gamma = _t(g["gamma"]).requires_grad_(True)
    beta = _t(g["beta"]).requires_grad_(True)
    rm, rv = _t(g["running_mean0"]).clone(), _t(g["running_var0"]).clone()
    idx = _t(g["idx"])
    out = edgeconv_ref.layer(x, idx, [W], None, gamma, beta, bn=True, eps=float(g["eps"]), slope=float(g["slope"]),
                             running=(rm, rv), momentum=float(g["momentum"]), training=True)
    out.backward(_t(g["g"]))
    assert np.allclose(out.detach().numpy(), g["out"], rtol=0, atol=1e-6 * np.abs(g["out"]).max())
    for got, want in ((x.grad, "grad_x"), (W.grad, "grad_weight"), (gamma.grad, "grad_gamma"), (beta.grad, "grad_beta")):
        assert np.allclose(got.numpy(), g[want], rtol=0, atol=2e-6 * np.abs(g[want]).max()), want
    assert np.allclose(rm.numpy(), g["running_mean1"], atol=1e-6) and np.allclose(rv.numpy(), g["running_var1"], atol=1e-6)
    with torch.no_grad():
        ev = edgeconv_ref.layer(x, idx, [W], None, gamma, beta, bn=True, eps=float(g["eps"]), slope=float(g["slope"]),
                                running=(rm, rv), training=False)
    assert np.allclose(ev.numpy(), g["out_eval"], rtol=0, atol=1e-6 * np.abs(g["out_eval"]).max())


def test_edgeconv_oracle_pointsegda_layer(golden):
    """== conv2(conv1(get_graph_feature(x))).max(-1) of PointSegDA/Models.py:171-174 (plain Conv2d stack with bias)."""
    import torch
    from oracle import edgeconv_ref
    torch.set_num_threads(1)
    g = golden("edgeconv_seg_3_64_64")
    x = _t(g["x"]).requires_grad_(True)
    ps = {n: _t(g[n]).requires_grad_(True) for n in ("w1", "b1", "w2", "b2")}
    out = edgeconv_ref.layer(x, _t(g["idx"]), [ps["w1"], ps["w2"]], [ps["b1"], ps["b2"]])
    out.backward(_t(g["g"]))
    assert np.allclose(out.detach().numpy(), g["out"], rtol=0, atol=1e-6 * np.abs(g["out"]).max())
    assert np.allclose(x.grad.numpy(), g["grad_x"], rtol=0, atol=2e-6 * np.abs(g["grad_x"]).max())
    for n in ps:
        assert np.allclose(ps[n].grad.numpy(), g["grad_" + n], rtol=0, atol=2e-6 * np.abs(g["grad_" + n]).max()), n
