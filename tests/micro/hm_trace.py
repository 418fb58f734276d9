"""Pipeline trace of the second-generation HMMA-aggregation kernel (FGC_HM_TRACE=1): clock64 stamps of CTA 0.
    FGC_HM_TRACE=1 python tests/micro/hm_trace.py [Cin Cout]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["FGC_HM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import mesh, ops

Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 530
_, F = mesh.grid_mesh(nq, nq, torus=True, morton=True)
adj_d, _ = ops.build_faces_adj(torch.from_numpy(F.astype(np.int32)).to(dev), K=16)
adj = adj_d[None].contiguous()
N = adj.shape[1]
g = torch.Generator(device="cpu").manual_seed(0)
M = 9
x = torch.randn(1, N, Cin, generator=g).to(dev)
W0 = (torch.randn(M, Cout, Cin, generator=g) * 0.05).to(dev)
b = (torch.randn(Cout, generator=g) * 0.01).to(dev)
u = (torch.randn(M, Cin, generator=g) * 0.05).to(dev)
v = (torch.randn(M, Cin, generator=g) * 0.05).to(dev)
c = (torch.randn(M, generator=g) * 0.05).to(dev)
for _ in range(3):
    y = ops.conv_fwd(x, adj, W0, b, u, v, c, act=1)
torch.cuda.synchronize()
out = (C.c_int64 * 1024)()
ops._lib.lib().fgc_debug_trace(out, 1024)
tr = np.array(list(out), dtype=np.int64).reshape(64, 16)
t0 = tr[tr > 0].min()
print("tile: agg0_arrive agg15_arrive | B3full_seen mma_issued | epi0: Dfull done | epi1: Dfull done | agg0: B3free wait begin end")
for t in range(4, 24):
    r = tr[t]
    f = lambda i: "%7d" % (r[i] - t0) if r[i] > 0 else "     -1"
    print("t%2d  %s %s | %s %s | %s %s | %s %s | %s %s" % (t, f(1), f(8), f(2), f(3), f(4), f(5), f(6), f(7), f(9), f(10)))
print("tile period (stage 2 issue): %.0f cycles" % np.diff(tr[4:40, 3]).mean())
print("(aggregator stamps -- columns 1, 2, 9, 10 -- need a build with -DFGC_HM_TRACE_AGG)")
