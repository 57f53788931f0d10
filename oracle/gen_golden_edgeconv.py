"""oracle/gen_golden_edgeconv.py -- TEST INFRASTRUCTURE: writes tests/golden/edgeconv_*.npz.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden_edgeconv
Every output array is produced by the REFERENCE'S OWN modules: PointDA/model_utils.py conv_2d (Conv2d + BatchNorm2d +
LeakyReLU) and get_graph_feature / knn for the PointDA layer, PointSegDA/Models.py get_graph_feature + the plain
nn.Conv2d stack of shared_layers (:159-163, :171-175) for the PointSegDA layer -- forward, the gradients autograd
gives for a fixed upstream gradient, and the BatchNorm running statistics after the step.  Parameters are randomised
(both signs of the BatchNorm weight, so the min-over-k branch is exercised) and stored next to the outputs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlsp_b200 import synth  # noqa: E402

REF = os.environ.get("MLSP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)


def load_reference():
    """PointDA.model_utils needs its never-reached third-party imports stubbed."""
    for p in (REF, os.path.join(REF, "PointDA")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _stub("pcl")
    _stub("termcolor", colored=lambda s, *a, **k: s)
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", DropPath=torch.nn.Identity, trunc_normal_=lambda *a, **k: None)
    _stub("knn_cuda", KNN=object)
    _stub("pointnet2_ops", pointnet2_utils=types.SimpleNamespace())
    import PointDA.model_utils as da
    import PointSegDA.Models as seg
    return da, seg


def _args():
    a = types.SimpleNamespace()
    a.cuda = False
    a.gpus = [-1]
    return a


def main():
    torch.set_num_threads(1)
    da, seg = load_reference()
    args = _args()
    gen = torch.Generator().manual_seed(97)

    # ---- PointDA layer: conv_2d(2C -> O, bias=False, leakyrelu) in training mode, then one more step in eval mode
    for name, C, O, N, k in (("edgeconv_da_16_32", 16, 32, 96, 8), ("edgeconv_da_3_64", 3, 64, 128, 20)):
        x = (synth.clouds(2, N, 5) if C == 3 else synth.smooth_features(2, C, N, 6)).requires_grad_(True)
        torch.manual_seed(1000 + C * 131 + O)                                  # the Conv2d / BatchNorm2d init draws from the global RNG
        layer = da.conv_2d(2 * C, O, kernel=1, bias=False, activation="leakyrelu")
        conv, bn = layer.conv[0], layer.conv[1]
        with torch.no_grad():
            bn.weight.copy_(torch.randn(O, generator=gen))                      # both signs
            bn.bias.copy_(0.5 * torch.randn(O, generator=gen))
            bn.running_mean.copy_(0.1 * torch.randn(O, generator=gen))
            bn.running_var.copy_(0.5 + torch.rand(O, generator=gen))
        rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
        idx = da.knn(x.detach(), k)
        layer.train()
        out = layer(da.get_graph_feature(x, args, k=k, idx=idx)).max(dim=-1, keepdim=False)[0]
        g = torch.randn(out.shape, generator=gen)
        out.backward(g)
        layer.eval()
        with torch.no_grad():
            out_eval = layer(da.get_graph_feature(x.detach(), args, k=k, idx=idx)).max(dim=-1, keepdim=False)[0]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.detach().numpy(), idx=idx.numpy(), k=k,
                            weight=conv.weight.detach().numpy().reshape(O, 2 * C), gamma=bn.weight.detach().numpy(),
                            beta=bn.bias.detach().numpy(), eps=bn.eps, momentum=bn.momentum, slope=0.2,
                            running_mean0=rm0.numpy(), running_var0=rv0.numpy(),
                            running_mean1=bn.running_mean.numpy(), running_var1=bn.running_var.numpy(),
                            out=out.detach().numpy(), g=g.numpy(), grad_x=x.grad.numpy(),
                            grad_weight=conv.weight.grad.numpy().reshape(O, 2 * C), grad_gamma=bn.weight.grad.numpy(),
                            grad_beta=bn.bias.grad.numpy(), out_eval=out_eval.numpy())

    # ---- PointSegDA layer: conv1 -> conv2 (plain nn.Conv2d with bias, no BatchNorm, no activation) -> max over k
    C, H, O, N, k = 3, 64, 64, 128, 20
    x = synth.clouds(2, N, 7).requires_grad_(True)
    torch.manual_seed(11)
    conv1 = torch.nn.Conv2d(2 * C, H, kernel_size=1, bias=True)                  # PointSegDA/Models.py:159-160
    conv2 = torch.nn.Conv2d(H, O, kernel_size=1, bias=True)
    idx = seg.knn(x.detach(), k)
    out = conv2(conv1(seg.get_graph_feature(x, args, k=k, idx=idx))).max(dim=-1, keepdim=False)[0]    # :171-174
    g = torch.randn(out.shape, generator=gen)
    out.backward(g)
    np.savez_compressed(os.path.join(OUT, "edgeconv_seg_3_64_64.npz"), x=x.detach().numpy(), idx=idx.numpy(), k=k,
                        w1=conv1.weight.detach().numpy().reshape(H, 2 * C), b1=conv1.bias.detach().numpy(),
                        w2=conv2.weight.detach().numpy().reshape(O, H), b2=conv2.bias.detach().numpy(),
                        out=out.detach().numpy(), g=g.numpy(), grad_x=x.grad.numpy(),
                        grad_w1=conv1.weight.grad.numpy().reshape(H, 2 * C), grad_b1=conv1.bias.grad.numpy(),
                        grad_w2=conv2.weight.grad.numpy().reshape(O, H), grad_b2=conv2.bias.grad.numpy())
    for f in sorted(os.listdir(OUT)):
        if f.startswith("edgeconv_"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
