"""Times the training-path linear layers (custom_lin, reference Code/model.py:763-769) at the C4 size.
    python tests/micro/lin_layer.py [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for cin, cout in ((32, 1024), (1024, 3)):
    x = torch.randn(1, rows, cin, generator=g).to(dev)
    W = (torch.randn(cin, cout, generator=g) * 0.05).to(dev)
    b = torch.zeros(cout, device=dev)
    gy = torch.randn(1, rows, cout, generator=g).to(dev)
    print("lin %4d -> %4d rows=%d: fwd %.4f ms, bwd %.4f ms" % (cin, cout, rows, timed(lambda: ops.lin_fwd(x, W, b)),
                                                                 timed(lambda: ops.lin_bwd(gy, x, W))))
for cin, cout in ((32, 1024), (1024, 3)):
    x = torch.randn(1, rows, cin, generator=g).to(dev)
    W = (torch.randn(cin, cout, generator=g) * 0.05).to(dev)
    gy = torch.randn(1, rows, cout, generator=g).to(dev)
    print("lin %4d -> %4d: bwd without gx %.4f ms" % (cin, cout, timed(lambda: ops.lin_bwd(gy, x, W, need_gx=False))))
