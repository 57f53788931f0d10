"""CPU emulation of the tcgen05 filter's pass-1 bound ("Which pass 1"): for the activations a
backbone produces, how many candidates per row does `v <= d_k + eps1 + 2 eps` admit with the bf16-head pass 1
(eps1 = 2^-7 |x_i| max|x_j|), as is and after centring the cloud on the mean of its first 32 points?  More than the
list capacity (64 for k <= 32) means the row falls back to the exact streaming selection.

    python tools/pass1_candidates.py            # PointSegDA shared layers (BatchNorm-free) on 2 x 2048 synthetic clouds
    python tools/pass1_candidates.py --pointda  # PointDA backbone (BatchNorm + LeakyReLU) on 2 x 1024
Needs no GPU (the layers are evaluated with the oracle's torch restatement of the reference)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mlsp_b200 import synth  # noqa: E402
from oracle import ref_torch  # noqa: E402


def main():
    seg = "--pointda" not in sys.argv
    torch.manual_seed(0)
    layers = bench._ec_make_layers("cpu", seg)
    h = synth.surface_clouds(2, 2048 if seg else 1024, 1234)
    k = 20
    with torch.no_grad():
        for li, seq in enumerate(layers):
            h = seq(ref_torch.get_graph_feature(h, k, ref_torch.knn(h, k))).max(dim=-1)[0]
            s = h[:, :, :32]
            ratio = (s ** 2).mean(2).sum(1) / ((s ** 2).mean(2) - s.mean(2) ** 2).sum(1)
            d = torch.cdist(h.transpose(1, 2), h.transpose(1, 2)) ** 2
            dk = d.topk(k, largest=False)[0][:, :, -1]
            row = [f"layer {li + 1} (C={h.shape[1]}): E|x|^2/Var = {ratio.mean():.1f}"]
            for name, hh in (("as is", h), ("centred", h - s.mean(2, keepdim=True))):
                nrm = (hh ** 2).sum(1)
                mx = nrm.max(1, keepdim=True)[0].sqrt()
                eps1, eps = 2 ** -7 * nrm.sqrt() * mx, 2 ** -12 * nrm.sqrt() * mx
                c = (d <= (dk + eps1 + 2 * eps).unsqueeze(-1)).sum(-1).float()
                row.append(f"{name}: {c.mean():.1f} candidates/row, {100 * (c > 64).float().mean():.1f} % rows overflow")
            print(" | ".join(row))


if __name__ == "__main__":
    main()
