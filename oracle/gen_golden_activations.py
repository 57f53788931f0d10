"""oracle/gen_golden_activations.py -- TEST INFRASTRUCTURE: real backbone activations + the reference's own knn output on them.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden_activations
SURVEY.md 8(d) prescribes the x1/x2/x3 activations of the seeded reference DGCNN as the input of the feature-space kNN
(C = 64, 64, 128) -- synthetic features are not what the tensor-core filter meets in a model.  Two fixtures:

* activations_da.npz  : PointDA DGCNN (PointDA/Models.py:82-162, BatchNorm layers: activations centred per channel), seeded
                        random init, training-mode forward of B=2 config-A clouds (N = 1024): the inputs of conv2/conv3/conv4
                        = x1, x2, x3, and knn(x, 20) of PointDA/model_utils.py:9-16 on each, computed by the reference on CPU.
* activations_seg.npz : PointSegDA shared_layers (PointSegDA/Models.py:146-192: plain biased Conv2d stacks, NO BatchNorm --
                        activations sit far from the origin), B=1, N = 2048: x1, x2 and PointSegDA's knn (Models.py:8-15).
idx is stored as int16 (N <= 32767)."""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlsp_b200 import synth  # noqa: E402
from oracle.gen_golden_edgeconv import load_reference  # noqa: E402
from oracle.gen_golden_dgcnn import ref_args  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.set_num_threads(1)                 # one thread: the fixture regenerates bit for bit
    da, seg = load_reference()
    import PointDA.Models as ref_models
    import model_utils as mu_top

    # ---- PointDA
    torch.manual_seed(11)
    model = ref_models.DGCNN(ref_args()).train()
    seen = []
    orig = mu_top.knn

    def spy(xx, k):
        idx = orig(xx, k)
        seen.append((xx.detach().clone(), idx.clone()))
        return idx

    mu_top.knn = spy
    x = synth.surface_clouds(2, 1024, 77)
    with torch.no_grad():
        model(x)
    mu_top.knn = orig
    # calls: [0] raw cloud (transform net), [1] transformed cloud (conv1), [2] x1, [3] x2, [4] x3
    assert [tuple(s[0].shape[1:]) for s in seen] == [(3, 1024), (3, 1024), (64, 1024), (64, 1024), (128, 1024)]
    out = {"seed": 11}
    for name, (xx, idx) in zip(("x1", "x2", "x3"), seen[2:]):
        out[name] = xx.numpy()
        out["idx_" + name] = idx.numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "activations_da.npz"), **out)
    print("activations_da.npz", os.path.getsize(os.path.join(OUT, "activations_da.npz")))

    # ---- PointSegDA
    torch.manual_seed(12)
    a = types.SimpleNamespace(cuda=False, gpus=[-1])
    sl = seg.shared_layers(a, in_size=3)
    seen = []
    orig_seg = seg.knn

    def spy_seg(xx, k):
        idx = orig_seg(xx, k)
        seen.append((xx.detach().clone(), idx.clone()))
        return idx

    seg.knn = spy_seg
    with torch.no_grad():
        sl(synth.surface_clouds(1, 2048, 78))
    seg.knn = orig_seg
    assert [tuple(s[0].shape[1:]) for s in seen] == [(3, 2048), (64, 2048), (64, 2048)]
    out = {"seed": 12}
    for name, (xx, idx) in zip(("x1", "x2"), seen[1:]):
        out[name] = xx.numpy()
        out["idx_" + name] = idx.numpy().astype(np.int16)
        m = xx.mean(dim=2).norm() / xx.std(dim=2).norm()
        print("seg", name, "|mean| / |std| over channels:", float(m))
    np.savez_compressed(os.path.join(OUT, "activations_seg.npz"), **out)
    print("activations_seg.npz", os.path.getsize(os.path.join(OUT, "activations_seg.npz")))


if __name__ == "__main__":
    main()
