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
    # fuse_edgeconv=True on top: only get_graph_feature moves (to the deferred feature), and unpatch still knows the
    # reference's own function
    from mlsp_b200 import lazy
    touched = patch.patch(modules=fake, fuse_edgeconv=True)
    assert sorted(touched) == sorted(f"{m}.get_graph_feature" for m in patch.TARGETS if "get_graph_feature" in patch.TARGETS[m])
    assert fake["PointDA.Models"].get_graph_feature is lazy.get_graph_feature and fake["model_utils"].knn is ops.knn
    assert patch.patch.originals[("PointDA.Models", "get_graph_feature")] is ref_fn
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
    # the reference modules stay owned by their model: attaching the fused layer to it adds no state_dict key and no
    # parameter, so reference checkpoints still load with strict=True and the optimiser sees each parameter once
    host = torch.nn.Module()
    host.conv1 = seq
    keys, nparam = set(host.state_dict()), len(list(host.parameters()))
    host.edge1 = layer
    assert set(host.state_dict()) == keys and len(list(host.parameters())) == nparam and list(layer.parameters()) == []
    own = edgeconv.FusedEdgeConv([torch.nn.Conv2d(6, 8, 1)], torch.nn.BatchNorm2d(8), 0.2)       # stand-alone: registered
    assert len(list(own.parameters())) == 4 and "bn.running_mean" in own.state_dict()
    with pytest.raises(M.MlspError):
        edgeconv.FusedEdgeConv.from_reference([seq, torch.nn.Conv2d(8, 8, 1)])
    with pytest.raises(M.MlspError):
        edgeconv.edge_conv(torch.zeros(1, 3, 8), torch.zeros(4, 6))   # no CPU path


# ---- mlsp_b200.lazy: the deferred graph feature (EdgeConv fusion behind the reference's unchanged model code)
@pytest.fixture
def lazy_on_cpu(monkeypatch):
    """The interception logic of mlsp_b200.lazy with its two CUDA computations swapped for the oracle (the reference's
    own op composition on the CPU): what is fused, what is materialised, and which arguments reach either."""
    from mlsp_b200 import lazy
    from oracle import edgeconv_ref, ref_torch

    def fused(x, W, k, b, bnp, slope, idx):
        idx = ref_torch.knn(x, k) if idx is None else idx
        if bnp is None:
            return edgeconv_ref.layer(x, idx, [W], [b], slope=slope)
        running = (bnp.running_mean, bnp.running_var) if bnp.update_running else None
        if not bnp.batch_stats:
            running = (bnp.running_mean, bnp.running_var)
        return edgeconv_ref.layer(x, idx, [W], [b], bnp.weight, bnp.bias, bn=True, eps=bnp.eps, slope=slope, running=running,
                                  momentum=bnp.momentum, training=bnp.batch_stats)

    monkeypatch.setattr(lazy.ops, "_require_cuda_f32", lambda *a, **k: None)
    monkeypatch.setattr(lazy, "_materialise_backend", lambda x, k, idx: ref_torch.get_graph_feature(x, k, idx))
    monkeypatch.setattr(lazy, "_fused_backend", fused)
    lazy.counters.clear()
    return lazy


def test_lazy_graph_feature_interception(lazy_on_cpu):
    import copy
    from oracle import ref_torch
    lazy = lazy_on_cpu
    torch.manual_seed(0)
    x = torch.randn(2, 3, 64)
    t = lazy.get_graph_feature(x, None, k=8)
    assert isinstance(t, lazy.LazyGraphFeature) and t.shape == (2, 6, 64, 8) and t.dim() == 4 and t.size(0) == 2
    assert t.get_device() == -1 and t.dtype == torch.float32 and not lazy.counters            # nothing has run
    # PointDA conv_2d + max over k -> one fused layer; BatchNorm's side effects happen exactly once
    seq = torch.nn.Sequential(torch.nn.Conv2d(6, 16, 1, bias=False), torch.nn.BatchNorm2d(16), torch.nn.LeakyReLU(0.2, inplace=True))
    twin = copy.deepcopy(seq)
    y = seq(t)
    assert isinstance(y, lazy.LazyGraphFeature) and y.shape == (2, 16, 64, 8) and not lazy.counters
    out = y.max(dim=-1, keepdim=False)[0]
    want = twin(ref_torch.get_graph_feature(x, 8)).max(dim=-1, keepdim=False)[0]
    assert torch.allclose(out, want, atol=1e-6) and lazy.counters == {"fused": 1}
    assert int(seq[1].num_batches_tracked) == 1 and torch.allclose(seq[1].running_var, twin[1].running_var, atol=1e-6)
    seq.eval(), twin.eval()
    out = seq(lazy.get_graph_feature(x, None, k=8)).max(dim=-1)[0]                               # running statistics
    assert torch.allclose(out, twin(ref_torch.get_graph_feature(x, 8)).max(dim=-1)[0], atol=1e-6)
    assert int(seq[1].num_batches_tracked) == 1 and lazy.counters == {"fused": 2}
    # PointSegDA: plain Conv2d stack with bias, no activation
    c1, c2 = torch.nn.Conv2d(6, 12, 1), torch.nn.Conv2d(12, 8, 1)
    out = c2(c1(lazy.get_graph_feature(x, None, k=8))).max(dim=-1, keepdim=False)[0]
    assert torch.allclose(out, c2(c1(ref_torch.get_graph_feature(x, 8))).max(dim=-1)[0], atol=1e-5) and lazy.counters["fused"] == 3
    # a second convolution after the activation (the input-transform nets) cannot be folded: materialise and carry on
    seq.train(), twin.train()
    tail = torch.nn.Sequential(torch.nn.Conv2d(16, 8, 1), torch.nn.ReLU())
    z = tail(seq(lazy.get_graph_feature(x, None, k=8)))
    assert type(z) is torch.Tensor and lazy.counters["materialised"] == 1
    assert torch.allclose(z, tail(twin(ref_torch.get_graph_feature(x, 8))), atol=1e-6)
    assert torch.allclose(seq[1].running_mean, twin[1].running_mean, atol=1e-6)
    # arbitrary use of the feature: it is the reference's tensor
    feat = ref_torch.get_graph_feature(x, 8)
    assert torch.equal(lazy.get_graph_feature(x, None, k=8) * 2.0, feat * 2.0)
    assert torch.equal(lazy.get_graph_feature(x, None, k=8)[1, :, 3], feat[1, :, 3])
    assert torch.cat((lazy.get_graph_feature(x, None, k=8), feat), 1).shape == (2, 12, 64, 8)
    assert torch.equal(lazy.get_graph_feature(x, None, k=8).max(dim=2)[0], feat.max(dim=2)[0])  # not the max over k
    wide = torch.nn.Conv2d(6, 8, (1, 3), padding=(0, 1))                                         # not a 1x1 convolution
    assert torch.allclose(wide(lazy.get_graph_feature(x, None, k=8)), wide(feat), atol=1e-6)
    with pytest.raises(RuntimeError):
        lazy.get_graph_feature(x, None, k=65)


@pytest.mark.skipif(not os.path.isdir("/root/reference/MLSP"), reason="reference checkout not mounted")
def test_lazy_fusion_runs_the_real_reference_models(lazy_on_cpu):
    """In the build container: the reference's OWN DGCNN (PointDA/Models.py:82-162) and DGCNN_DefRec
    (PointSegDA/Models.py:197-242) with get_graph_feature rebound by patch(fuse_edgeconv=True) -- model code untouched --
    give the same logits and BatchNorm statistics as unpatched, with every EdgeConv layer fused (4 and 3) and only the
    input-transform net's feature materialised."""
    import copy
    import warnings
    from oracle import gen_golden_edgeconv
    lazy = lazy_on_cpu
    gen_golden_edgeconv.load_reference()
    import PointDA.Models as PM
    import PointSegDA.Models as SM
    warnings.filterwarnings("ignore")
    args = types.SimpleNamespace(num_class=10, cuda=False, gpus=[-1], model="dgcnn", dropout=0.5, encoder_type="",
                                 density_num_class=16, pergroup=2)
    torch.manual_seed(0)
    cases = [(PM.DGCNN(args), dict(activate_DefRec=True, activate_normal=True), 4),
             (SM.DGCNN_DefRec(types.SimpleNamespace(cuda=False, gpus=[-1], dropout=0.5, density_num_class=16, pergroup=5),
                              in_size=3, num_classes=8), dict(make_seg=True, activate_DefRec=True), 3)]
    x = M.synth.surface_clouds(2, 128, 5)
    for model, kw, n_fused in cases:
        twin = copy.deepcopy(model)
        torch.manual_seed(1)
        want = twin(x.clone(), **kw)                                   # the reference, untouched
        lazy.counters.clear()
        touched = patch.patch(fuse_edgeconv=True)
        try:
            assert any(t.endswith("get_graph_feature") for t in touched)
            torch.manual_seed(1)
            got = model(x.clone(), **kw)
        finally:
            patch.unpatch()
        assert lazy.counters == {"fused": n_fused, "materialised": 1}, dict(lazy.counters)
        assert set(got) == set(want)
        for name in want:
            assert torch.allclose(got[name], want[name], rtol=1e-4, atol=1e-5), name
        for (n1, b1), (_, b2) in zip(model.named_buffers(), twin.named_buffers()):
            assert torch.allclose(b1.float(), b2.float(), rtol=1e-4, atol=1e-6), n1


def test_dgcnn_mirror_has_the_reference_parameters_and_init():
    """mlsp_b200.dgcnn.DGCNN is a structural mirror of PointDA/Models.py:DGCNN: for the same torch.manual_seed it has the same
    state_dict keys and bit-identical initial weights as the reference's class (digest stored by oracle/gen_golden_dgcnn.py,
    which runs the reference's own DGCNN) -- so reference checkpoints load with strict=True."""
    import hashlib
    import os
    import numpy as np
    import torch
    from mlsp_b200 import dgcnn
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dgcnn_ondef.npz"))
    torch.manual_seed(int(g["seed"]))
    m = dgcnn.DGCNN(num_class=10, density_num_class=16, pergroup=2, dropout=0.0)
    h = hashlib.sha256()
    for name, p in sorted(m.state_dict().items()):
        h.update(name.encode())
        h.update(p.detach().cpu().numpy().tobytes())
    assert h.hexdigest() == bytes(g["digest"]).decode()
    assert sum(p.numel() for p in m.parameters()) == 4548915          # the 18.2 MB all-reduce payload


def test_dgcnn_seg_mirror_has_the_reference_parameters_and_init():
    """mlsp_b200.dgcnn_seg.DGCNN_DefRec mirrors PointSegDA/Models.py:197-242: same state_dict keys and bit-identical initial
    weights for the same seed (digest stored by oracle/gen_golden_dgcnn_seg.py from the reference's own class)."""
    import hashlib
    import os
    import numpy as np
    import torch
    from mlsp_b200 import dgcnn_seg
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dgcnn_seg.npz"))
    torch.manual_seed(int(g["seed"]))
    m = dgcnn_seg.DGCNN_DefRec(in_size=3, num_classes=8, density_num_class=16, pergroup=5, dropout=0.0)
    h = hashlib.sha256()
    for name, p in sorted(m.state_dict().items()):
        h.update(name.encode())
        h.update(p.detach().cpu().numpy().tobytes())
    assert h.hexdigest() == bytes(g["digest"]).decode()
    assert sum(p.numel() for p in m.parameters()) == 3082612
    assert not m.Density_cls.fc2.weight.requires_grad                 # PointSegDA/Models.py:373
