import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from facet_graph_convolution_b200 import ops, _lib
dev = torch.device('cuda:0')
n = 1_000_000
adj = torch.from_numpy(bench.make_adjacency(n, 'mesh')).unsqueeze(0).to(dev)
W0, b, u, v, c = (torch.from_numpy(t).to(dev) for t in bench.make_params(1234))
x = torch.randn(1, n, 64, device=dev); gy = torch.randn(1, n, 64, device=dev)
rev = ops.ReverseAdjacency(adj); plan = ops.ConvPlan(adj, 8); rev.target_plan(8)
def step():
    s = ops.ConvSaved()
    y = ops.conv_fwd(x, adj, W0, b, u, v, c, plan=plan, save=s)
    return y, ops.conv_bwd(gy, x, adj, rev, W0, u, v, c, plan=plan, saved=s)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("CPU enqueue per step %.3f ms, total per step %.3f ms" % ((t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
