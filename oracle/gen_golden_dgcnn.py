"""oracle/gen_golden_dgcnn.py -- TEST INFRASTRUCTURE: fixtures for mlsp_b200.dgcnn made by the reference's OWN DGCNN class.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden_dgcnn
`torch.manual_seed(SEED)` then `PointDA.Models.DGCNN(args)` gives the reference's initial weights; mlsp_b200.dgcnn.DGCNN is
a structural mirror (same submodules, same construction order), so the same seed gives it the same weights -- the fixture
checks that too (a digest of every parameter) and then pins the training-mode forward with the three MLSP heads
(activate_density_normal_ondef=True, PointDA/Models.py:156-160), a scalar loss and its gradients.  Dropout is 0 so that the
fixture is deterministic."""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlsp_b200 import synth  # noqa: E402
from oracle.gen_golden_edgeconv import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED, B, N = 5, 4, 256


def ref_args():
    a = types.SimpleNamespace()
    a.cuda, a.gpus = False, [-1]
    a.model, a.encoder_type = "dgcnn", None
    a.num_class, a.dropout = 10, 0.0
    a.density_num_class, a.pergroup = 16, 2
    return a


def param_digest(model):
    h = hashlib.sha256()
    for name, p in sorted(model.state_dict().items()):
        h.update(name.encode())
        h.update(p.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def scalar_loss(logits):
    return (logits["DefRec"].square().mean() + logits["Normal"].square().mean() + logits["density_mse"].mean()
            + (logits["density"] * torch.arange(16.0)).sum(1).mean() + logits["cls"].square().mean())


def pcm_fixture():
    """PCM.mix_shapes (MLSP/PCM.py:6-38) by the reference's own function, seeded torch + numpy RNG."""
    from MLSP import PCM
    a = types.SimpleNamespace(cuda=False, gpus=[-1], mixup_params=1.0)
    X = synth.surface_clouds(4, 512, 41)
    Y = torch.arange(4)
    torch.manual_seed(7)
    np.random.seed(7)
    mixed, (Ya, Yb, lam) = PCM.mix_shapes(a, X, Y)
    np.savez_compressed(os.path.join(OUT, "pcm_mix.npz"), X=X.numpy(), Y=Y.numpy(), seed=7, mixed=mixed.numpy(), Ya=Ya.numpy(),
                        Yb=Yb.numpy(), lam=float(lam))
    print("pcm_mix.npz", os.path.getsize(os.path.join(OUT, "pcm_mix.npz")), "lam", float(lam))


def main():
    torch.set_num_threads(1)
    load_reference()
    import PointDA.Models as ref_models
    pcm_fixture()
    torch.manual_seed(SEED)
    model = ref_models.DGCNN(ref_args())
    digest = param_digest(model)
    model.train()
    # the inputs of the four EdgeConv layers and their outputs (after the max over k), for the layer-by-layer comparison: a
    # DGCNN is discontinuous in its inputs (a 1e-6 perturbation can swap the 20th and 21st neighbour of a point), so only the
    # loss is compared end to end; each stage is compared on the reference's own input for it
    import model_utils as mu_top
    stage_in, stage_out = [], []
    orig_knn = mu_top.knn

    def spy(xx, k):
        stage_in.append(xx.detach().clone())
        return orig_knn(xx, k)

    mu_top.knn = spy
    for n in ("conv1", "conv2", "conv3", "conv4"):
        getattr(model, n).register_forward_hook(lambda m, i, o: stage_out.append(o.detach().max(dim=-1)[0].clone()))
    x = synth.surface_clouds(B, N, 31).requires_grad_(True)
    logits = model(x, activate_density_normal_ondef=True)
    mu_top.knn = orig_knn
    loss = scalar_loss(logits)
    loss.backward()
    g = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    np.savez_compressed(
        os.path.join(OUT, "dgcnn_ondef.npz"), seed=SEED, x=x.detach().numpy(), digest=np.frombuffer(digest.encode(), dtype=np.uint8),
        cls=logits["cls"].detach().numpy(), DefRec=logits["DefRec"].detach().numpy(), Normal=logits["Normal"].detach().numpy(),
        density=logits["density"].detach().numpy(), density_mse=logits["density_mse"].detach().numpy(), loss=float(loss.detach()),
        grad_x=x.grad.numpy(), grad_conv1=g["conv1.conv.0.weight"].numpy(), grad_conv4=g["conv4.conv.0.weight"].numpy(),
        grad_fc3=g["input_transform_net.fc3.weight"].numpy(), grad_defrec_conv1=g["DefRec.conv1.weight"].numpy()[:, ::16, 0],
        grad_density_conv1=g["Density_cls.conv1.weight"].numpy()[:, ::16, 0], grad_cls_mlp3=g["C.mlp3.weight"].numpy(),
        bn5_running_mean=model.bn5.running_mean.numpy(), conv2_bn_running_var=model.conv2.conv[1].running_var.numpy(),
        xt=stage_in[1].numpy(), x1=stage_out[0].numpy(), x2=stage_out[1].numpy(), x3=stage_out[2].numpy(), x4=stage_out[3].numpy())
    print("dgcnn_ondef.npz", os.path.getsize(os.path.join(OUT, "dgcnn_ondef.npz")), "loss", float(loss.detach()))


if __name__ == "__main__":
    main()
