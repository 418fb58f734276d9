"""Scratch: forward time of the network's 8 layer shapes (M=9, K=16) at patch size and batched."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import ops, mesh
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
M, K = 9, 16
layers = [(6, 32, 1), (32, 64, 4), (64, 128, 16), (128, 128, 16), (128, 64, 4), (128, 64, 4), (64, 32, 1), (64, 32, 1)]
for B in (1, 16):
    tot = 0.0
    for ci, co, div in layers:
        side = 100 // int(round(div ** 0.5))   # 100x100 quads at level 0
        _, F = mesh.grid_mesh(side, side, torus=True, morton=True)
        adj = mesh.dedup_adj(mesh.faces_large_adj(F, K))
        n = adj.shape[0]
        a = torch.from_numpy(np.broadcast_to(adj[None], (B, n, K)).copy()).to(dev)
        x = torch.randn(B, n, ci, device=dev)
        W0 = torch.randn(M, co, ci, device=dev) * 0.05; b = torch.randn(co, device=dev) * 0.01
        u = torch.randn(M, ci, device=dev) * 0.05; v = torch.randn(M, ci, device=dev) * 0.05; c = torch.randn(M, device=dev) * 0.05
        for _ in range(3): ops.conv_fwd(x, a, W0, b, u, v, c, act=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.conv_fwd(x, a, W0, b, u, v, c, act=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * n * (M * ci * co + 13 * M * ci + 2 * M * ci)
        print("B=%2d  %3d->%3d  rows %7d  %.4f ms  %.2f TFLOP/s" % (B, ci, co, B * n, ms, fl / ms / 1e9))
        tot += ms
    print("B=%d total %.3f ms -> %.1f M facets/s (20000 facets per patch)" % (B, tot, B * 20000 / tot / 1e3))
