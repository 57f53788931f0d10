#!/usr/bin/env python
"""Per-tensor relative differences between mlsp_b200.dgcnn.DGCNN on the GPU and the reference-made fixture dgcnn_ondef.npz."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlsp_b200 import dgcnn
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dgcnn_ondef.npz"))
torch.manual_seed(int(g["seed"]))
model = dgcnn.DGCNN(num_class=10, density_num_class=16, pergroup=2, dropout=0.0).to(dev).train()
x = torch.from_numpy(g["x"]).to(dev).requires_grad_(True)
logits = model(x, activate_density_normal_ondef=True)
loss = (logits["DefRec"].square().mean() + logits["Normal"].square().mean() + logits["density_mse"].mean()
        + (logits["density"] * torch.arange(16.0, device=dev)).sum(1).mean() + logits["cls"].square().mean())
loss.backward()
print("loss", float(loss.detach()), float(g["loss"]))
def rel(a, ref):
    a, ref = a.detach().cpu().numpy().astype(np.float64), ref.astype(np.float64)
    return np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30), (np.abs(a - ref) > 1e-4 * np.abs(ref).max()).mean()
P = dict(model.named_parameters())
for name, t in [("cls", logits["cls"]), ("DefRec", logits["DefRec"]), ("Normal", logits["Normal"]), ("density", logits["density"]),
                ("density_mse", logits["density_mse"]), ("grad_x", x.grad), ("grad_conv1", P["conv1.conv.0.weight"].grad),
                ("grad_conv4", P["conv4.conv.0.weight"].grad), ("grad_fc3", P["input_transform_net.fc3.weight"].grad),
                ("bn5_running_mean", model.bn5.running_mean), ("conv2_bn_running_var", model.conv2.conv[1].running_var)]:
    print("%-22s max rel %.3e   frac > 1e-4: %.4f" % ((name,) + rel(t, g[name])))
