"""The PointDA DGCNN and its MLSP heads on the B200 hot path -- the caller of a1/a2.

A host-side mirror of the reference's model (PointDA/Models.py:82-285 with the blocks of PointDA/model_utils.py:45-147):
the same submodule tree, parameter names and construction order, so that

* a reference checkpoint loads with `load_state_dict(strict=True)`, and
* a freshly constructed model has the SAME initial weights as the reference's for the same `torch.manual_seed`
  (tests/golden/dgcnn_*.npz are made that way by oracle/gen_golden_dgcnn.py with the reference's own class).

What differs is how the forward is executed:

* the four EdgeConv layers run through `edgeconv.edge_conv` (no (B,2C,N,k) edge tensor; kNN on tcgen05 for C = 64/128);
* the input-transform net takes its (B,6,N,k) edge tensor from the fused knn + gather kernel;
* 8f rank 4 -- the three point-wise heads (RegionReconstruction, Normal_prediction, Density_prediction; Models.py:156-160)
  read the same (B,1536,N) input `cat(x_cat, x5.repeat)`: their first 1x1 convolutions are ONE GEMM on x_cat (512 input
  channels) plus a per-cloud bias from x5 (the repeated 1024 global channels are constant over the points), i.e. the
  concatenated / repeated input is never materialised and the contraction is a third of the reference's.

BatchNorm statistics, Dropout and every parameter keep torch semantics (the layers are the torch modules themselves; in
training mode BatchNorm and the activation after it run as one fused pass pair, mlsp_b200.bn / csrc/bn.cu).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import edgeconv, linear, ops, pool
from .bn import bn_act

K = 20  # PointDA/Models.py:13


def conv1x1(x: torch.Tensor, conv: nn.Module) -> torch.Tensor:
    """A 1x1 Conv1d / Conv2d as the GEMM it is, on the hand-written tcgen05 GEMM (mlsp_b200.linear -> mlsp_gemm_f32; fp32
    result up to summation order).  x (B,C,N) or (B,C,N,k) in channels-last strides; same result as conv(x)."""
    return linear.conv1x1(x, conv.weight, conv.bias)


def fc(x: torch.Tensor, lin: nn.Linear) -> torch.Tensor:
    """nn.Linear on a 2-D input through the same GEMM."""
    return linear.linear(x, lin.weight, lin.bias)


class Conv2dBlock(nn.Module):
    """conv_2d of PointDA/model_utils.py:45-63 (1x1 Conv2d + BatchNorm2d + activation), same attribute names."""

    def __init__(self, in_ch, out_ch, kernel=1, activation="leakyrelu", bias=False):
        super().__init__()
        act = nn.LeakyReLU(negative_slope=0.2, inplace=True) if activation == "leakyrelu" else nn.ReLU(inplace=True)
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=kernel, bias=bias), nn.BatchNorm2d(out_ch), act)

    def forward(self, x):
        slope = self.conv[2].negative_slope if isinstance(self.conv[2], nn.LeakyReLU) else 0.0
        return bn_act(conv1x1(x, self.conv[0]), self.conv[1], slope)      # BatchNorm + activation in one pass pair (bn.cu)


class FcBlock(nn.Module):
    """fc_layer of PointDA/model_utils.py:65-89 (Linear + BatchNorm1d + activation)."""

    def __init__(self, in_ch, out_ch, bn=True, activation="leakyrelu", bias=True):
        super().__init__()
        self.ac = nn.LeakyReLU(negative_slope=0.2, inplace=True) if activation == "leakyrelu" else nn.ReLU(inplace=True)
        if bn:
            self.fc = nn.Sequential(nn.Linear(in_ch, out_ch, bias=bias), nn.BatchNorm1d(out_ch), self.ac)
        else:
            self.fc = nn.Sequential(nn.Linear(in_ch, out_ch), self.ac)

    def forward(self, x):
        x = fc(x, self.fc[0])
        slope = self.ac.negative_slope if isinstance(self.ac, nn.LeakyReLU) else 0.0
        if len(self.fc) == 3:
            return bn_act(x, self.fc[1], slope)
        return self.ac(x)


class TransformNet(nn.Module):
    """transform_net of PointDA/model_utils.py:92-130 in its DGCNN configuration (leakyrelu, no conv bias)."""

    def __init__(self, in_ch=6, out=3):
        super().__init__()
        self.K = out
        self.conv2d1 = Conv2dBlock(in_ch, 64, bias=False)
        self.conv2d2 = Conv2dBlock(64, 128, bias=False)
        self.conv2d3 = Conv2dBlock(128, 1024, bias=False)
        self.fc1 = FcBlock(1024, 512, bias=False)
        self.fc2 = FcBlock(512, 256)
        self.fc3 = nn.Linear(256, out * out)

    def forward(self, x):                                   # x: the (B,6,N,k) edge tensor of the raw cloud
        x = self.conv2d2(self.conv2d1(x))
        x = pool.max_over_neighbours(x).unsqueeze(3)
        x = self.conv2d3(x)
        x = pool.max_over_points_cl(x).view(x.size(0), -1)
        x = fc(self.fc2(self.fc1(x)), self.fc3)
        x = x + torch.eye(self.K, device=x.device, dtype=x.dtype).view(1, self.K * self.K)
        return x.view(x.size(0), self.K, self.K)


class Classifier(nn.Module):
    """classifier of PointDA/model_utils.py:132-147 (dgcnn configuration)."""

    def __init__(self, num_class=10, dropout=0.5):
        super().__init__()
        self.mlp1 = FcBlock(1024, 512, bias=True)
        self.dp1 = nn.Dropout(p=dropout)
        self.mlp2 = FcBlock(512, 256, bias=True)
        self.dp2 = nn.Dropout(p=dropout)
        self.mlp3 = nn.Linear(256, num_class)

    def forward(self, x):
        return fc(self.dp2(self.mlp2(self.dp1(self.mlp1(x)))), self.mlp3)


class PointHead(nn.Module):
    """RegionReconstruction / Normal_prediction of PointDA/Models.py:165-227: 1x1 Conv1d stack input -> 256 -> 256 -> 128 -> 3."""

    of1, of2, of3 = 256, 256, 128

    def __init__(self, input_size, dropout=0.5):
        super().__init__()
        self.bn1 = nn.BatchNorm1d(self.of1)
        self.bn2 = nn.BatchNorm1d(self.of2)
        self.bn3 = nn.BatchNorm1d(self.of3)
        self.dp1 = nn.Dropout(p=dropout)
        self.dp2 = nn.Dropout(p=dropout)
        self.conv1 = nn.Conv1d(input_size, self.of1, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(self.of1, self.of2, kernel_size=1, bias=False)
        self.conv3 = nn.Conv1d(self.of2, self.of3, kernel_size=1, bias=False)
        self.conv4 = nn.Conv1d(self.of3, 3, kernel_size=1, bias=False)

    def tail(self, h1):
        """Everything after the first convolution; h1 = conv1(input) (B,256,N)."""
        x = self.dp1(bn_act(h1, self.bn1, 0.0))
        x = self.dp2(bn_act(conv1x1(x, self.conv2), self.bn2, 0.0))
        x = bn_act(conv1x1(x, self.conv3), self.bn3, 0.0)
        return conv1x1(x, self.conv4).permute(0, 2, 1)

    def forward(self, x):
        return self.tail(conv1x1(x, self.conv1))


class DensityHead(nn.Module):
    """Density_prediction of PointDA/Models.py:229-285 (dgcnn configuration): Conv1d input -> 512, then a per-point MLP to
    `num_class` soft cardinality classes and their expectation under the fixed ladder fc2 = pergroup * class index."""

    of1 = 512

    def __init__(self, input_size, num_class=16, pergroup=2, dropout=0.5):
        super().__init__()
        self.bn1 = nn.BatchNorm1d(self.of1)
        self.dp1 = nn.Dropout(p=dropout)
        self.conv1 = nn.Conv1d(input_size, self.of1, kernel_size=1, bias=False)
        self.num_class = num_class
        self.mlp1 = FcBlock(512, 256, bias=True)
        self.dp1 = nn.Dropout(p=dropout)
        self.mlp2 = FcBlock(256, 256, bias=True)
        self.dp2 = nn.Dropout(p=dropout)
        self.mlp3 = nn.Linear(256, num_class)
        self.fc2 = nn.Linear(num_class, 1, bias=False)
        with torch.no_grad():
            self.fc2.weight.copy_(pergroup * torch.arange(num_class, dtype=torch.float32).view(1, -1))
        self.fc2.weight.requires_grad = False

    def tail(self, h1):
        x = self.dp1(bn_act(h1, self.bn1, 0.0))
        x = x.permute(0, 2, 1).reshape(-1, self.of1)
        x = self.dp1(self.mlp1(x))
        p_vec = F.softmax(fc(self.dp2(self.mlp2(x)), self.mlp3), dim=1)
        return p_vec, self.fc2(p_vec)[:, 0]

    def forward(self, x):
        return self.tail(conv1x1(x, self.conv1))


class DGCNN(nn.Module):
    """DGCNN of PointDA/Models.py:82-162.  forward(x (B,3,N), activate_*=...) -> dict of logits, like the reference."""

    def __init__(self, num_class=10, density_num_class=16, pergroup=2, dropout=0.5, k=K):
        super().__init__()
        self.k = k
        self.input_transform_net = TransformNet(6, 3)
        self.conv1 = Conv2dBlock(6, 64, bias=False)
        self.conv2 = Conv2dBlock(64 * 2, 64, bias=False)
        self.conv3 = Conv2dBlock(64 * 2, 128, bias=False)
        self.conv4 = Conv2dBlock(128 * 2, 256, bias=False)
        num_f_prev = 64 + 64 + 128 + 256
        self.bn5 = nn.BatchNorm1d(1024)
        self.conv5 = nn.Conv1d(num_f_prev, 1024, kernel_size=1, bias=False)
        self.C = Classifier(num_class, dropout)
        self.DefRec = PointHead(num_f_prev + 1024, dropout)
        self.Norm_pred = PointHead(num_f_prev + 1024, dropout)
        self.Rec_scan = PointHead(num_f_prev + 1024, dropout)
        self.Density_cls = DensityHead(num_f_prev + 1024, density_num_class, pergroup, dropout)
        self.num_f_prev = num_f_prev
        # the EdgeConv layers share the parameters of conv1..conv4 (not registered twice: state_dict keys stay the reference's)
        object.__setattr__(self, "_edge", tuple(edgeconv.FusedEdgeConv.from_reference(getattr(self, n), k=k)
                                               for n in ("conv1", "conv2", "conv3", "conv4")))

    # ---- 8f rank 4: first layers of several heads as one contraction, without cat(x_cat, x5.repeat)
    def heads_first_layer(self, x_cat, x5, heads):
        """[head.conv1(cat(x_cat, x5[:, :, None].repeat(1, 1, N)))  for head in heads] as ONE GEMM over x_cat's 512 channels plus a
        per-cloud bias from the 1024 global channels (constant over the points).  x_cat (B,512,N), x5 (B,1024)."""
        C = self.num_f_prev
        W = torch.cat([h.conv1.weight.squeeze(-1) for h in heads], dim=0)        # (sum O, 1536)
        y = linear.conv1x1(x_cat, W[:, :C]) + linear.linear(x5, W[:, C:]).unsqueeze(2)
        return torch.split(y, [h.conv1.out_channels for h in heads], dim=1)

    def backbone(self, x):
        """x (B,3,N) -> (x_cat (B,512,N), x5 (B,1024)) : PointDA/Models.py:111-138."""
        B = x.size(0)
        x0 = ops.get_graph_feature(x, None, k=self.k)                            # fused knn + gather, (B,6,N,k)
        T = self.input_transform_net(x0)
        x = linear.apply_transform(T, x)
        feats = []
        h = x
        for layer in self._edge:                                                 # graph feature -> conv_2d -> max over k, x 4
            h = layer(h.contiguous())
            feats.append(h)
        x_cat = torch.cat(feats, dim=1)
        x5 = bn_act(conv1x1(x_cat, self.conv5), self.bn5, 0.2)
        x5 = pool.global_max_pool(x5).view(B, -1)
        return x_cat, x5

    def forward(self, x, visualization=False, activate_DefRec=False, activate_normal=False, activate_scan=False,
                activate_density=False, activate_density_normal_ondef=False):
        logits = {}
        x_cat, x5 = self.backbone(x)
        logits["cls"] = self.C(x5)
        if visualization:
            return x5
        want = []
        if activate_DefRec or activate_density_normal_ondef:
            want.append(("DefRec", self.DefRec))
        if activate_normal or activate_density_normal_ondef:
            want.append(("Normal", self.Norm_pred))
        if activate_scan:
            want.append(("Rec_scan", self.Rec_scan))
        if activate_density or activate_density_normal_ondef:
            want.append(("density", self.Density_cls))
        if want:
            firsts = self.heads_first_layer(x_cat, x5, [h for _, h in want])
            for (name, head), h1 in zip(want, firsts):
                if name == "density":
                    logits["density"], logits["density_mse"] = head.tail(h1)
                else:
                    logits[name] = head.tail(h1)
        return logits


# ---- the small elementwise losses of the target branch (a11: they stay torch) ----------------------------------------
def normal_loss(normal_pred, normal_gt, mask_cord, weight=1.0):
    """PointDA/trainer.py:551-557: -sum(|cos(pred, gt)| * w) / sum(w) on unit vectors (the sign of a normal is irrelevant)."""
    normal_pred = F.normalize(normal_pred, p=2, dim=-1)
    normal_gt = F.normalize(normal_gt, p=2, dim=-1)
    return -weight * torch.sum(torch.abs(torch.sum(normal_pred * normal_gt, dim=-1)) * mask_cord) / torch.sum(mask_cord)


def density_loss(logits, target, target_vec, mask=None, weight=1.0, lambda_1=0.05, lambda_2=1.0):
    """densityloss of MLSP/mlsp.py:430-454 -> (KL term, L1 term)."""
    p_vec, p_val = logits["density"], logits["density_mse"]
    tmp = torch.sum(target_vec * torch.log(p_vec + 1e-10), dim=1)
    l1 = F.l1_loss(p_val, target, reduction="none")
    if mask is not None:
        return (-weight * torch.sum(tmp * mask) / torch.sum(mask) * lambda_2,
                weight * torch.sum(l1 * mask) / torch.sum(mask) * lambda_1)
    return -weight * torch.mean(tmp) * lambda_2, weight * torch.mean(l1) * lambda_1


def target_branch_loss(model, trgt_batch, lookup, *, near=20, radius=0.13, density_num_class=16, pergroup=2, shift=0,
                       DefRec_weight=0.5, normal_pred_weight=1.0, Density_weight=1.0, DefRec_dist="volume_based_voxels",
                       defpart=False, pending=None):
    """The target-branch loss of one training step, PointDA/trainer.py:522-566 (Density_normal_viainput, Normal_ondef,
    Density_ondef): local-structure targets of the undeformed batch (one 3-D neighbourhood pass), deformation, forward with the
    three heads, position (Chamfer) + normal + cardinality losses.  trgt_batch (B,N,3) as the loader yields it.
    pending: the handle of `ops.deform_input_begin(trgt_batch.permute(0, 2, 1))` issued earlier in the step (before the source
    branch): the deformation then needs no stream synchronisation here (same results, same RNG stream positions)."""
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
    logits = model(trgt.contiguous(), activate_density_normal_ondef=True)
    loss = DefRec_weight * ops.reconstruction_loss(logits["DefRec"], trgt_orig, mask) * ops.DefRec_SCALER
    mask_cord = mask.permute(0, 2, 1)[:, :, 0]
    if not defpart:
        mask_cord = mask_cord * 26 + 1
    loss = loss + normal_loss(logits["Normal"], normal_gt, mask_cord, normal_pred_weight)
    kl, l1 = density_loss(logits, density_mse_label, density_label, mask_cord.reshape(-1), Density_weight)
    return loss + kl + l1 + 0.0 * logits["cls"].sum()      # the classifier takes part in every backward (DDP: no unused parameters)
