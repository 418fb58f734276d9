"""Times one facet-graph convolution layer of the network (M = 9) on a mesh adjacency.
    python tests/micro/hm_layer.py [Cin Cout [quads_x quads_y [K [iters]]]]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import mesh, ops

Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nx = int(sys.argv[3]) if len(sys.argv) > 3 else 530
ny = int(sys.argv[4]) if len(sys.argv) > 4 else 530
K = int(sys.argv[5]) if len(sys.argv) > 5 else 16
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
M = int(os.environ.get("HM_M", "9"))
t0 = time.time()
dev = torch.device("cuda:0")
_, F = mesh.grid_mesh(nx, ny, torus=True, morton=True)
adj_d, _ = ops.build_faces_adj(torch.from_numpy(F.astype(np.int32)).to(dev), K=K)
adj = adj_d[None].contiguous()
N = adj.shape[1]
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(1, N, Cin, generator=g).to(dev)
W0 = (torch.randn(M, Cout, Cin, generator=g) * 0.05).to(dev)
b = (torch.randn(Cout, generator=g) * 0.01).to(dev)
u = (torch.randn(M, Cin, generator=g) * 0.05).to(dev)
v = (torch.randn(M, Cin, generator=g) * 0.05).to(dev)
c = (torch.randn(M, generator=g) * 0.05).to(dev)
for _ in range(3):
    y = ops.conv_fwd(x, adj, W0, b, u, v, c, act=1)
torch.cuda.synchronize()
L = ops._lib.lib()
import ctypes as C
L.fgc_profile_begin(C.c_void_p(torch.cuda.current_stream().cuda_stream))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = ops.conv_fwd(x, adj, W0, b, u, v, c, act=1)
e1.record()
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
L.fgc_profile_end(buf, len(buf))
ms = e0.elapsed_time(e1) / iters
print("layer %d->%d M=%d K=%d rows=%d: %.4f ms per call, %.1f M facets/s (setup %.1fs)" % (Cin, Cout, M, K, N, ms, N / ms / 1e3, time.time() - t0))
for line in buf.value.decode().strip().split("\n"):
    nm, tot, n = line.split()
    print("   %-28s %8.4f ms x %d per call" % (nm, float(tot) / iters, int(n) // iters))
