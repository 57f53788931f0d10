"""The PointSegDA DGCNN_DefRec (segmentation backbone + MLSP heads) on the B200 hot path -- the second caller of a1/a2.

A host-side mirror of PointSegDA/Models.py:48-392 (same submodule tree, parameter names and construction order as the
reference: a reference checkpoint loads with `strict=True`, the same `torch.manual_seed` gives the same initial weights;
tests/golden/dgcnn_seg.npz is made that way with the reference's own class by oracle/gen_golden_dgcnn_seg.py).  Differences
to the PointDA model of mlsp_b200/dgcnn.py: the conv_2d / fc_layer blocks carry no BatchNorm, the three EdgeConv layers are
stacks of plain biased convolutions without activation (folded into one (W, b) per layer, PointSegDA/Models.py:159-184),
the heads' convolutions have biases, and there is a per-point segmentation head.  Execution is the same: fused EdgeConv
layers (no (B,2C,N,k) tensor), every product on mlsp_gemm_f32, the heads' first layers as ONE 192-channel product plus a
per-cloud bias from the 1024 global channels (the reference concatenates and repeats to (B,1216,N) and convolves it four
times, Models.py:224-241), max-poolings on pool.cu.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import edgeconv, linear, ops, pool
from .bn import bn_act
from .dgcnn import DensityHead, FcBlock, conv1x1, density_loss, fc, normal_loss

K = 20  # PointSegDA/Models.py:6


class Conv2dAct(nn.Module):
    """conv_2d of PointSegDA/Models.py:48-64: 1x1 Conv2d + activation, NO BatchNorm; same attribute names."""

    def __init__(self, in_ch, out_ch, kernel=1, activation="leakyrelu", bias=False):
        super().__init__()
        act = nn.LeakyReLU(negative_slope=0.2, inplace=True) if activation == "leakyrelu" else nn.ReLU(inplace=True)
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=kernel, bias=bias), act)

    def forward(self, x):
        return self.conv[1](conv1x1(x, self.conv[0]))


class TransformNet(nn.Module):
    """transform_net of PointSegDA/Models.py:106-143 (fc_layer = Linear + LeakyReLU, no BatchNorm)."""

    def __init__(self, in_ch=6, out=3):
        super().__init__()
        self.K = out
        self.conv2d1 = Conv2dAct(in_ch, 64, bias=False)
        self.conv2d2 = Conv2dAct(64, 128, bias=False)
        self.conv2d3 = Conv2dAct(128, 1024, bias=False)
        self.fc1 = FcBlock(1024, 512, bn=False)
        self.fc2 = FcBlock(512, 256, bn=False)
        self.fc3 = nn.Linear(256, out * out)

    def forward(self, x):
        x = self.conv2d2(self.conv2d1(x))
        x = pool.max_over_neighbours(x).unsqueeze(3)
        x = self.conv2d3(x)
        x = pool.max_over_points_cl(x).view(x.size(0), -1)
        x = fc(self.fc2(self.fc1(x)), self.fc3)
        x = x + torch.eye(self.K, device=x.device, dtype=x.dtype).view(1, self.K * self.K)
        return x.view(x.size(0), self.K, self.K)


class SharedLayers(nn.Module):
    """shared_layers of PointSegDA/Models.py:146-192: three EdgeConv layers of plain biased convolutions, conv6, global max."""

    def __init__(self, in_size=3, k=K):
        super().__init__()
        self.k = k
        self.conv1 = nn.Conv2d(in_size * 2, 64, kernel_size=1, bias=True)
        self.conv2 = nn.Conv2d(64, 64, kernel_size=1, bias=True)
        self.conv3 = nn.Conv2d(64 * 2, 64, kernel_size=1, bias=True)
        self.conv4 = nn.Conv2d(64, 64, kernel_size=1, bias=True)
        self.conv5 = nn.Conv2d(64 * 2, 64, kernel_size=1, bias=True)
        self.conv6 = nn.Conv1d(64 * 3, 1024, kernel_size=1, bias=True)
        object.__setattr__(self, "_edge", (edgeconv.FusedEdgeConv.from_reference([self.conv1, self.conv2], k=k),
                                           edgeconv.FusedEdgeConv.from_reference([self.conv3, self.conv4], k=k),
                                           edgeconv.FusedEdgeConv.from_reference([self.conv5], k=k)))

    def layers_sum(self):
        return 64 * 3

    def forward(self, x):
        feats = []
        h = x
        for layer in self._edge:
            h = layer(h.contiguous())
            feats.append(h)
        x123 = torch.cat(feats, dim=1)
        x5 = pool.global_max_pool(conv1x1(x123, self.conv6))                      # (B,1024,1)
        return x123, x5


class PointHeadSeg(nn.Module):
    """segmentation / DeformationReconstruction / Normal_prediction of PointSegDA/Models.py:245-331."""

    of1, of2, of3 = 256, 256, 128

    def __init__(self, input_size, out_size, dropout=0.5, bias=True):
        super().__init__()
        self.bn1 = nn.BatchNorm1d(self.of1)
        self.bn2 = nn.BatchNorm1d(self.of2)
        self.bn3 = nn.BatchNorm1d(self.of3)
        self.dp1 = nn.Dropout(p=dropout)
        self.dp2 = nn.Dropout(p=dropout)
        self.conv1 = nn.Conv1d(input_size, self.of1, kernel_size=1, bias=bias)
        self.conv2 = nn.Conv1d(self.of1, self.of2, kernel_size=1, bias=bias)
        self.conv3 = nn.Conv1d(self.of2, self.of3, kernel_size=1, bias=bias)
        self.conv4 = nn.Conv1d(self.of3, out_size, kernel_size=1, bias=bias)

    def tail(self, h1):
        x = self.dp1(bn_act(h1, self.bn1, 0.0))
        x = self.dp2(bn_act(conv1x1(x, self.conv2), self.bn2, 0.0))
        x = bn_act(conv1x1(x, self.conv3), self.bn3, 0.0)
        return conv1x1(x, self.conv4).permute(0, 2, 1)

    def forward(self, x):
        return self.tail(conv1x1(x, self.conv1))


class DGCNN_DefRec(nn.Module):
    """DGCNN_DefRec of PointSegDA/Models.py:197-242.  forward(x (B,3,N), make_seg=..., activate_*=...) -> dict of logits."""

    def __init__(self, in_size=3, num_classes=8, density_num_class=16, pergroup=5, dropout=0.5, k=K):
        super().__init__()
        self.k = k
        self.input_transform_net = TransformNet(in_size * 2, in_size)
        self.shared_layers = SharedLayers(in_size=in_size, k=k)
        self.num_f_prev = self.shared_layers.layers_sum()
        self.seg = PointHeadSeg(1024 + self.num_f_prev, num_classes, dropout, bias=True)
        self.DefRec = PointHeadSeg(1024 + self.num_f_prev, in_size, dropout, bias=True)
        self.Norm_pred = PointHeadSeg(self.num_f_prev + 1024, 3, dropout, bias=False)
        self.Density_cls = DensityHead(self.num_f_prev + 1024, density_num_class, pergroup, dropout)

    def heads_first_layer(self, x123, x5, heads):
        """[head.conv1(cat(x123, x5.repeat(1, 1, N))) for head in heads] as ONE product over x123's 192 channels plus a per-cloud
        bias from the 1024 global channels (8f rank 4; PointSegDA/Models.py:224-241)."""
        C = self.num_f_prev
        W = torch.cat([h.conv1.weight.squeeze(-1) for h in heads], dim=0)                 # (sum O, 1216)
        bias = None
        if any(h.conv1.bias is not None for h in heads):
            bias = torch.cat([h.conv1.bias if h.conv1.bias is not None else torch.zeros(h.conv1.out_channels, device=W.device)
                              for h in heads])
        y = linear.conv1x1(x123, W[:, :C]) + linear.linear(x5, W[:, C:], bias).unsqueeze(2)
        return torch.split(y, [h.conv1.out_channels for h in heads], dim=1)

    def forward(self, x, make_seg=True, activate_DefRec=True, activate_normal=False, activate_density=False,
                activate_density_normal_ondef=False):
        logits = {}
        x0 = ops.get_graph_feature(x, None, k=self.k)                                      # fused knn + gather, (B,6,N,k)
        T = self.input_transform_net(x0)
        x = linear.apply_transform(T, x)
        x123, x5 = self.shared_layers(x)
        want = []
        if make_seg:
            want.append(("seg", self.seg))
        if activate_DefRec or activate_density_normal_ondef:
            want.append(("DefRec", self.DefRec))
        if activate_normal or activate_density_normal_ondef:
            want.append(("Normal", self.Norm_pred))
        if activate_density or activate_density_normal_ondef:
            want.append(("density", self.Density_cls))
        if want:
            firsts = self.heads_first_layer(x123, x5.squeeze(2), [h for _, h in want])
            for (name, head), h1 in zip(want, firsts):
                if name == "density":
                    logits["density"], logits["density_mse"] = head.tail(h1)
                else:
                    logits[name] = head.tail(h1)
        return logits


def target_branch_loss(model, trgt_batch, lookup, *, near=10, radius=0.081, density_num_class=16, pergroup=5, shift=10,
                       DefRec_weight=0.02, normal_pred_weight=0.02, Density_weight=0.02, DefRec_dist="volume_based_voxels",
                       defpart=False, pending=None):
    """The target-branch loss of a PointSegDA step, PointSegDA/trainer.py:381-431 (Density_normal_viainput, Normal_ondef,
    Density_ondef; the trainer's default weights).  trgt_batch (B,N,3) as the loader yields it; `pending` as in
    mlsp_b200.dgcnn.target_branch_loss."""
    normal_gt, density_label, density_mse_label = ops.target_structure(trgt_batch, near, radius, density_num_class, pergroup, shift)
    density_label = density_label.reshape(-1, density_num_class)
    density_mse_label = density_mse_label.to(torch.float32).reshape(-1)
    trgt = trgt_batch.permute(0, 2, 1)
    trgt_orig = trgt.clone()
    if pending is not None:
        if pending.X.data_ptr() != trgt.data_ptr() or pending.X.shape != trgt.shape:
            raise ops.MlspError("target_branch_loss: `pending` was begun on a different batch")
        trgt, mask = ops.deform_input_finish(pending, lookup, DefRec_dist)
    else:
        trgt, mask = ops.deform_input(trgt, lookup, DefRec_dist, trgt.device)
    logits = model(trgt.contiguous(), make_seg=False, activate_density_normal_ondef=True)
    loss = DefRec_weight * ops.reconstruction_loss(logits["DefRec"], trgt_orig, mask) * ops.DefRec_SCALER
    mask_cord = mask.permute(0, 2, 1)[:, :, 0]
    if not defpart:
        mask_cord = mask_cord + 1                                                          # trainer.py:412
    loss = loss + normal_loss(logits["Normal"], normal_gt, mask_cord, normal_pred_weight)
    kl, l1 = density_loss(logits, density_mse_label, density_label, mask_cord.reshape(-1), Density_weight)
    loss = loss + kl + l1
    if hasattr(model, "module"):
        # under DistributedDataParallel every parameter must take part in the backward that all-reduces: the segmentation head is
        # not on this forward (make_seg=False, like the reference), so its parameters join with a zero coefficient -- their
        # gradients from the source branch's local backward are then averaged with everything else
        loss = loss + 0.0 * sum(p.sum() for p in model.module.seg.parameters())
    return loss


def source_branch_loss(model, src_batch, src_labels, DefRec_weight=0.02):
    """PointSegDA/trainer.py:298-310: per-point segmentation cross-entropy on the source batch.  src_batch (B,N,3), labels (B,N)."""
    logits = model(src_batch.permute(0, 2, 1).contiguous(), make_seg=True, activate_DefRec=False)
    return (1 - DefRec_weight) * F.cross_entropy(logits["seg"].permute(0, 2, 1), src_labels)
