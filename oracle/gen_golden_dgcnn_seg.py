"""oracle/gen_golden_dgcnn_seg.py -- TEST INFRASTRUCTURE: fixture for mlsp_b200.dgcnn_seg made by the reference's OWN
PointSegDA DGCNN_DefRec class (PointSegDA/Models.py:197-242).

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden_dgcnn_seg
Same recipe as oracle/gen_golden_dgcnn.py: `torch.manual_seed(SEED)` then the reference's constructor gives the reference's
initial weights, the mirror built after the same seed must have the same ones (parameter digest); the training-mode forward
with every head (make_seg + activate_density_normal_ondef), a scalar loss and its gradients are pinned, and so are the inputs
/ outputs of the three EdgeConv layers for the stage-by-stage comparison (a DGCNN is discontinuous in its inputs).  Dropout
is 0 so that the fixture is deterministic."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlsp_b200 import synth  # noqa: E402
from oracle.gen_golden_dgcnn import param_digest  # noqa: E402
from oracle.gen_golden_edgeconv import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED, B, N = 9, 4, 256


def ref_args():
    return types.SimpleNamespace(cuda=False, gpus=[-1], dropout=0.0, density_num_class=16, pergroup=5)


def scalar_loss(logits):
    return (logits["DefRec"].square().mean() + logits["Normal"].square().mean() + logits["density_mse"].mean()
            + (logits["density"] * torch.arange(16.0, device=logits["density"].device)).sum(1).mean()
            + logits["seg"].square().mean())


def main():
    torch.set_num_threads(1)
    _, seg = load_reference()
    torch.manual_seed(SEED)
    model = seg.DGCNN_DefRec(ref_args(), in_size=3, num_classes=8)
    digest = param_digest(model)
    model.train()
    stage_in = []
    orig_knn = seg.knn

    def spy(xx, k):
        stage_in.append(xx.detach().clone())
        return orig_knn(xx, k)

    seg.knn = spy
    x = synth.surface_clouds(B, N, 37).requires_grad_(True)
    logits = model(x, make_seg=True, activate_DefRec=False, activate_density_normal_ondef=True)
    seg.knn = orig_knn
    # knn is called on: the input (transform net), the aligned cloud, x1, x2
    assert len(stage_in) == 4, len(stage_in)
    with torch.no_grad():
        x123, x5 = model.shared_layers(stage_in[1])
    loss = scalar_loss(logits)
    loss.backward()
    g = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    path = os.path.join(OUT, "dgcnn_seg.npz")
    np.savez_compressed(
        path, seed=SEED, x=x.detach().numpy(), digest=np.frombuffer(digest.encode(), dtype=np.uint8),
        seg=logits["seg"].detach().numpy(), DefRec=logits["DefRec"].detach().numpy(), Normal=logits["Normal"].detach().numpy(),
        density=logits["density"].detach().numpy(), density_mse=logits["density_mse"].detach().numpy(), loss=float(loss.detach()),
        grad_x=x.grad.numpy(), grad_conv1=g["shared_layers.conv1.weight"].numpy(), grad_conv5=g["shared_layers.conv5.weight"].numpy(),
        grad_fc3=g["input_transform_net.fc3.weight"].numpy(),
        grad_seg_conv1=g["seg.conv1.weight"].numpy()[:, ::16, 0], grad_seg_conv4_bias=g["seg.conv4.bias"].numpy(),
        grad_norm_conv1=g["Norm_pred.conv1.weight"].numpy()[:, ::16, 0],
        seg_bn1_running_mean=model.seg.bn1.running_mean.numpy(),
        xt=stage_in[1].numpy(), x1=stage_in[2].numpy(), x2=stage_in[3].numpy(), x123=x123.numpy(), x5=x5.numpy())
    print("dgcnn_seg.npz", os.path.getsize(path), "loss", float(loss.detach()))


if __name__ == "__main__":
    main()
