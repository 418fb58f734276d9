"""Times the fused regression head (32 -> 1024 -> 3) on `rows` rows.   python tests/micro/head_layer.py [rows [iters]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 561600
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(1, rows, 32, generator=g).to(dev)
W1 = (torch.randn(32, 1024, generator=g) * 0.05).to(dev)
b1 = (torch.randn(1024, generator=g) * 0.01).to(dev)
W2 = (torch.randn(1024, 3, generator=g) * 0.05).to(dev)
b2 = (torch.randn(3, generator=g) * 0.01).to(dev)
for _ in range(3):
    y = ops.mlp_head(x, W1, b1, W2, b2, 0.1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = ops.mlp_head(x, W1, b1, W2, b2, 0.1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("head rows=%d: %.4f ms per call (%.1f M rows/s)" % (rows, ms, rows / ms / 1e3))
