"""Times the convolution backward (FFMA family for the network's M = 9 layers) per kernel.
    python tests/micro/bwd_layer.py [Cin Cout [B N [K]]]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import mesh, ops

Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 64
K = int(sys.argv[5]) if len(sys.argv) > 5 else 16
M = 9
dev = torch.device("cuda:0")
_, F = mesh.grid_mesh(nq, nq, torus=True, morton=True)
adj1 = mesh.faces_large_adj(F, K).astype(np.int32)
adj = torch.from_numpy(np.stack([adj1] * B)).to(dev)
N = adj.shape[1]
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(B, N, Cin, generator=g).to(dev)
gy = torch.randn(B, N, Cout, generator=g).to(dev)
W0 = (torch.randn(M, Cout, Cin, generator=g) * 0.05).to(dev)
u = (torch.randn(M, Cin, generator=g) * 0.05).to(dev)
v = (torch.randn(M, Cin, generator=g) * 0.05).to(dev)
c = (torch.randn(M, generator=g) * 0.05).to(dev)
rev = ops.ReverseAdjacency(adj)
for _ in range(3):
    ops.conv_bwd(gy, x, adj, rev, W0, u, v, c)
torch.cuda.synchronize()
L = ops._lib.lib()
iters = 10
L.fgc_profile_begin(C.c_void_p(torch.cuda.current_stream().cuda_stream))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.conv_bwd(gy, x, adj, rev, W0, u, v, c)
e1.record()
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
L.fgc_profile_end(buf, len(buf))
print("bwd %d->%d M=%d K=%d rows=%d: %.4f ms per call" % (Cin, Cout, M, K, B * N, e0.elapsed_time(e1) / iters))
for line in buf.value.decode().strip().split("\n"):
    nm, tot, n = line.split()
    print("   %-28s %8.4f ms x %d per call" % (nm, float(tot) / iters, int(n) // iters))
