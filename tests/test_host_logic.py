"""CPU: host-side logic of the drop-in layer -- RNG-stream parity helpers, name rebinding, the pcl shim
surface, and the N>1 sharding / timing reduction over gloo (world_size 2)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import mlsp_b200 as M
from mlsp_b200 import dist as mdist
from mlsp_b200 import ops, patch, pcl_shim


def test_draw_gaussians_is_numpy_multivariate_normal():
    """The batched draw must be bit-identical to the reference's per-cloud np.random.multivariate_normal calls
    (utils/pc_utils.py:122) and leave the global RNG in the same state."""
    assert ops._MVN_DIAG is not None
    means = np.array([[1 / 3, -2 / 3, 0.0], [0.1, 0.2, 0.3], [-0.66666669, 0.0, 0.66666669]], dtype=np.float32)
    counts = [47, 0, 61]
    np.random.seed(11)
    ref = [np.random.multivariate_normal(m, np.eye(3) * 0.001, n) for m, n in zip(means, counts)]
    after_ref = np.random.random()
    np.random.seed(11)
    got = ops._draw_gaussians(means, counts)
    after_got = np.random.random()
    assert np.array_equal(np.concatenate(ref, axis=0), got)
    assert after_ref == after_got
    np.random.seed(3)
    assert ops._draw_gaussians(means, [0, 0, 0]).shape == (0, 3)


def test_region_mean_matches_oracle():
    from oracle import np_ops
    assert np.array_equal(M.region_mean(3), np_ops.region_mean(3))
    assert M.region_mean(3).shape == (27, 3)
    assert np.allclose(M.region_mean(3)[0], [-2 / 3, -2 / 3, -2 / 3]) and np.allclose(M.region_mean(3)[26], [2 / 3] * 3)


def test_patch_rebinds_every_namespace():
    def ref_fn(*a, **k):
        return "reference"
    fake = {}
    for name, table in patch.TARGETS.items():
        mod = types.ModuleType(name)
        for attr in table:
            setattr(mod, attr, ref_fn)
        mod.untouched = ref_fn
        fake[name] = mod
    touched = patch.patch(modules=fake)
    assert len(touched) == sum(len(t) for t in patch.TARGETS.values())
    assert fake["PointDA.Models"].get_graph_feature is ops.get_graph_feature
    assert fake["model_utils"].knn is ops.knn and fake["PointSegDA.Models"].knn is ops.knn
    assert fake["utils.pc_utils"].farthest_point_sample is ops.farthest_point_sample
    assert fake["MLSP.mlsp"].deform_input is ops.deform_input
    assert fake["MLSP.mlsp"].cal_density is ops.cal_density
    assert fake["MLSP.mlsp"].untouched is ref_fn
    assert patch.patch(modules=fake) == []                     # idempotent
    patch.patch.originals.clear()


def test_patch_signatures_match_reference_call_sites():
    import inspect
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(ops.knn)[:2] == ["x", "k"]
    assert sig(ops.get_graph_feature) == ["x", "args", "k", "idx"]
    assert sig(ops.farthest_point_sample) == ["args", "xyz", "npoint"]
    assert sig(ops.deform_input) == ["X", "lookup", "DefRec_dist", "device", "groups"]
    assert sig(ops.cal_density) == ["batch_pts", "radius", "num_cls", "pergroup", "shift", "K"]
    assert sig(ops.reconstruction_loss) == ["pred", "gold", "mask"]
    assert sig(ops.chamfer_distance) == ["p1", "p2", "mask"]
    assert sig(ops.calc_loss) == ["args", "logits", "labels", "mask"]
    assert inspect.signature(ops.get_graph_feature).parameters["k"].default == 20
    assert inspect.signature(ops.deform_input).parameters["DefRec_dist"].default == "volume_based_voxels"
    assert inspect.signature(ops.cal_density).parameters["K"].default == 100


@pytest.mark.skipif(not os.path.isdir("/root/reference/MLSP"), reason="reference checkout not mounted")
def test_patch_on_the_real_reference_modules():
    """In the build container: the real reference modules import with the shim and get rebound."""
    saved_pcl = sys.modules.pop("pcl", None)
    try:
        assert pcl_shim.install() is True
        from oracle import ref_import
        pcu, mlsp, seg = ref_import.load()
        touched = patch.patch()
        assert "PointSegDA.Models.get_graph_feature" in touched and "MLSP.mlsp.reconstruction_loss" in touched
        assert "utils.pc_utils.farthest_point_sample" in touched and "MLSP.mlsp.cal_density" in touched
        assert mlsp.reconstruction_loss is ops.reconstruction_loss and seg.knn is ops.knn
        patch.unpatch()
        assert seg.knn is not ops.knn and mlsp.reconstruction_loss is not ops.reconstruction_loss
    finally:
        patch.unpatch()
        if saved_pcl is not None:
            sys.modules["pcl"] = saved_pcl


def test_pcl_shim_surface():
    saved = sys.modules.pop("pcl", None)
    try:
        assert pcl_shim.install(force=True)
        import pcl
        cloud = pcl.PointCloud()
        cloud.from_array(np.zeros((5, 3), np.float32))
        assert cloud.size == 5
        ne = cloud.make_NormalEstimation()
        ne.set_SearchMethod(cloud.make_kdtree())
        ne.set_KSearch(3)
        other = pcl.PointCloud(np.ones((5, 3), np.float32))
        with pytest.raises(NotImplementedError):                       # only the reference's usage: query == indexed cloud
            cloud.make_kdtree_flann().radius_search_for_cloud(other, 0.1, 100)
        with pytest.raises(RuntimeError):                              # no CPU fallback: the search itself needs the GPU op
            cloud.make_kdtree_flann().radius_search_for_cloud(cloud, 0.1, 100)
        with pytest.raises(ValueError):
            cloud.from_array(np.zeros((5, 2), np.float32))
    finally:
        sys.modules.pop("pcl", None)
        if saved is not None:
            sys.modules["pcl"] = saved


def test_deform_input_rejects_unsupported():
    with pytest.raises(M.MlspError):
        M.deform_input(torch.zeros(1, 3, 8), None)


def test_shard_bounds():
    for B, W in [(32, 8), (33, 8), (5, 8), (256, 3)]:
        spans = [mdist.shard_bounds(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mdist.shard_bounds(8, 8, 8)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = mdist.shard_bounds(33, rank, world)
        from mlsp_b200 import synth
        mine = synth.clouds(hi - lo, 64, seed=1234 + rank)          # per-rank seeded shard, no exchange
        t = mdist.max_over_ranks(1.0 + rank)
        counts = mdist.gather_counts(hi - lo)
        q.put((rank, lo, hi, t, counts, float(mine.sum())))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, c0, s0), (r1, lo1, hi1, t1, c1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 17, 17, 33)
    assert t0 == t1 == 2.0                                          # max over ranks
    assert c0 == c1 == [17, 16] and sum(c0) == 33
    assert s0 != s1                                                 # different shards


def test_edgeconv_weight_algebra():
    """Host algebra of mlsp_b200.edgeconv (CPU part only): the [Wa ; Wb-Wa] split reproduces W.[x_j-x_i | x_i], and a
    stack of plain 1x1 convolutions with biases folds into one (W, b) -- the facts edge_conv's single GEMM rests on."""
    from mlsp_b200 import edgeconv
    torch.manual_seed(0)
    C, O = 5, 8
    W = torch.randn(O, 2 * C, 1, 1, dtype=torch.float64)
    xi, xj = torch.randn(C, dtype=torch.float64), torch.randn(C, dtype=torch.float64)
    Wcat = edgeconv._split_weight(W, C)
    want = W.flatten(1) @ torch.cat((xj - xi, xi))
    got = Wcat[:O] @ xj + Wcat[O:] @ xi
    assert torch.allclose(got, want, atol=1e-12)
    with pytest.raises(M.MlspError):
        edgeconv._split_weight(W, C + 1)
    c1 = torch.nn.Conv2d(2 * C, 6, 1, bias=True).double()
    c2 = torch.nn.Conv2d(6, O, 1, bias=False).double()
    c3 = torch.nn.Conv2d(O, O, 1, bias=True).double()
    layer = edgeconv.FusedEdgeConv.from_reference([c1, c2, c3], k=4)
    Weff, beff = layer.effective_weight_bias()
    e = torch.randn(2, 2 * C, 7, 3, dtype=torch.float64)
    want = c3(c2(c1(e)))
    got = torch.einsum("oc,bcnk->bonk", Weff, e) + beff.view(1, -1, 1, 1)
    assert torch.allclose(got, want, atol=1e-12)
    assert layer.bn is None and layer.negative_slope is None
    seq = torch.nn.Sequential(torch.nn.Conv2d(6, 8, 1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.LeakyReLU(0.2))
    wrapped = types.SimpleNamespace(conv=seq)                       # shape of the reference's conv_2d module
    layer = edgeconv.FusedEdgeConv.from_reference(wrapped, k=20)
    assert layer.bn is seq[1] and layer.negative_slope == 0.2 and layer.convs[0] is seq[0]
    with pytest.raises(M.MlspError):
        edgeconv.FusedEdgeConv.from_reference([seq, torch.nn.Conv2d(8, 8, 1)])
    with pytest.raises(M.MlspError):
        edgeconv.edge_conv(torch.zeros(1, 3, 8), torch.zeros(4, 6))   # no CPU path
    # the layer's GEMM with its hand-written backward (grad_x straight in (B,C,N), weight gradient from B partial products)
    x = torch.randn(3, C, 7, dtype=torch.float64, requires_grad=True)
    Wc = torch.randn(2 * O, C, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(3, 7, 2 * O, dtype=torch.float64)
    edgeconv._PointwiseYZ.apply(x, Wc).backward(gy)
    gx, gw = x.grad.clone(), Wc.grad.clone()
    x.grad = Wc.grad = None
    ref = torch.matmul(x.transpose(1, 2), Wc.t())
    ref.backward(gy)
    assert torch.allclose(edgeconv._PointwiseYZ.apply(x, Wc), ref, atol=1e-12)
    assert torch.allclose(gx, x.grad, atol=1e-12) and torch.allclose(gw, Wc.grad, atol=1e-12) and gx.is_contiguous()
