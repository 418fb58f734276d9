"""Prints the measured GPU-vs-reference errors behind the loss / gradient tolerances of tests/test_gpu_net.py
(run on a GPU box: python tests/micro/tolerance_probe.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden  # noqa: E402
from facet_graph_convolution_b200 import model as fm  # noqa: E402

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

g = golden("small_ops")
loss = fm.faceNormalsLoss(T(g["loss_fn"]), T(g["loss_gt"])).item()
print("small_ops loss abs err (deg): %.3e of %.4f" % (abs(loss - float(g["loss"])), float(g["loss"])))
nt = T(g["norm_in"]).requires_grad_(True)
fm.faceNormalsLoss(fm.normalizeTensor(nt), T(g["loss_gt"])).backward()
ref = g["loss_norm_grad"]
print("small_ops loss∘normalize grad rel err: %.3e" % (np.abs(nt.grad.cpu().numpy() - ref).max() / np.abs(ref).max()))

g = golden("net_train_small")
n = int(g["nparams"])
store = fm.VariableStore(dev, params=[g["p%02d" % i] for i in range(n)], requires_grad=True)
adjs = [T(g["adj0"]), T(g["adj1"]), T(g["adj2"])]
with fm.variable_store(store):
    y = fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0)
print("net_train_small y max-abs err: %.3e" % np.abs(y.detach().cpu().numpy() - g["y"]).max())
loss = fm.faceNormalsLoss(fm.normalizeTensor(y), T(g["gt"]))
print("net_train_small loss abs err (deg): %.3e of %.4f" % (abs(loss.item() - float(g["loss"])), float(g["loss"])))
loss.backward()
worst = 0.0
for i, t in enumerate(store.params):
    ref = g["g%02d" % i]
    scale = max(float(np.abs(ref).max()), 1e-3)
    e = np.abs(t.grad.cpu().numpy() - ref).max() / scale
    worst = max(worst, e)
    if e > 1e-4:
        print("   grad %2d %-28s rel err %.3e (|ref|max %.3e)" % (i, store.names[i], e, float(np.abs(ref).max())))
print("net_train_small worst grad rel err: %.3e" % worst)
