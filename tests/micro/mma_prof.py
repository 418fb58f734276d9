"""Scratch: planned forward at C2 scale, per-kernel times via the library profiler."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import ops, mesh, _lib
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rs = np.random.RandomState(0)
kind = sys.argv[1] if len(sys.argv) > 1 else "mesh"
n = 1_000_000
if kind == "random":
    adj = rs.randint(1, n + 1, size=(n, 16)).astype(np.int32); adj[:, 0] = np.arange(1, n + 1)
else:
    _, F = mesh.grid_mesh(1000, 500, torus=True, morton=True)
    adj = mesh.faces_large_adj(F, 16)
    if kind == "dedup": adj = mesh.dedup_adj(adj)
a = T(adj[None])
W0 = T((rs.randn(8, 64, 64) * 0.05).astype(np.float32)); b = T((rs.randn(64) * 0.01).astype(np.float32))
u = T((rs.randn(8, 64) * 0.05).astype(np.float32)); v = T((rs.randn(8, 64) * 0.05).astype(np.float32)); c = T((rs.randn(8) * 0.05).astype(np.float32))
x = torch.randn(1, n, 64, device=dev)
plan = ops.ConvPlan(a, 8)
R = plan.buf[256:256 + 4 * ((n + 15) // 16)].view(torch.int32)
print("distinct rows per tile: mean %.1f max %d  >64: %.2f%%" % (R.float().mean().item(), R.max().item(), 100 * (R > 64).float().mean().item()))
L = _lib.lib()
for _ in range(2): ops.conv_fwd(x, a, W0, b, u, v, c, plan=plan)
torch.cuda.synchronize()
L.fgc_profile_begin(C.c_void_p(torch.cuda.current_stream().cuda_stream))
for _ in range(3): ops.conv_fwd(x, a, W0, b, u, v, c, plan=plan)
buf = C.create_string_buffer(1 << 14); L.fgc_profile_end(buf, len(buf))
for ln in buf.value.decode().strip().splitlines():
    nm, tot, cnt = ln.split(); print("%-28s %.4f ms" % (nm, float(tot) / int(cnt)))
if os.environ.get("FGC_MMA_TRACE"):
    out = (C.c_int64 * 1024)()
    L.fgc_debug_trace(out, 1024)
    tr = np.array(list(out), dtype=np.int64).reshape(4, 32, 8)
    t0 = tr[tr > 0].min()
    names = ["drain: D1full cvt B3free B3full_arr | epilogue: D3full done", "q: ready Qfree Qfull_arr top Xfull staged_read", "loader: top Xfree issued arrived",
             "mma: top D1free Q+Xfull S1done B3wait B3full S3done Xfull"]
    for role in range(4):
        print(names[role])
        for t in range(2, 12):
            print("  t%2d " % t + " ".join("%7d" % (v - t0 if v > 0 else -1) for v in tr[role, t]))
