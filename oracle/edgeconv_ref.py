"""oracle/edgeconv_ref.py -- TEST INFRASTRUCTURE (see oracle/__init__.py): the reference's EdgeConv layer, restated.

One DGCNN layer of the reference is the composition
    get_graph_feature -> Conv2d 1x1 [-> BatchNorm2d] [-> LeakyReLU] -> max over k
(PointDA/Models.py:114-128 with conv_2d of PointDA/model_utils.py:45-63; PointSegDA/Models.py:171-184 with the plain
Conv2d stacks of PointSegDA/Models.py:159-163).  `layer()` below is that composition, op for op, on top of
oracle.ref_torch.get_graph_feature; tests compare mlsp_b200.edgeconv.edge_conv (which never builds the (B,2C,N,k)
tensor) against it and against the fixtures `python -m oracle.gen_golden_edgeconv` made by running the reference's OWN
conv_2d / Conv2d modules and get_graph_feature (tests/golden/edgeconv_*.npz; pinned in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ref_torch


def layer(x, idx, weights, biases=None, gamma=None, beta=None, bn=False, eps=1e-5, slope=None,
          running=None, momentum=0.1, training=True):
    """x (B,C,N), idx (B,N,k) -> (B,O,N).  weights: list of (O_l, I_l) matrices applied in order (1x1 convolutions);
    biases: matching list (entries may be None); bn: BatchNorm2d after the last convolution (training: batch statistics,
    `running` = (mean, var) tensors updated in place like torch); slope: LeakyReLU slope or None."""
    k = idx.shape[2]
    h = ref_torch.get_graph_feature(x, k, idx)                                   # (B,2C,N,k)
    biases = biases or [None] * len(weights)
    for W, b in zip(weights, biases):
        h = F.conv2d(h, W.view(W.shape[0], W.shape[1], 1, 1), b)
    if bn:
        rm, rv = running if running is not None else (None, None)
        h = F.batch_norm(h, rm, rv, gamma, beta, training or rm is None, momentum, eps)
    if slope is not None:
        h = F.leaky_relu(h, slope)
    return h.max(dim=-1, keepdim=False)[0]
