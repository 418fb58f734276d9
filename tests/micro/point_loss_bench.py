"""Times the point-set loss (fgc_point_set_loss) and the backward of the multi-scale vertex update on a B200:
    python tests/micro/point_loss_bench.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from facet_graph_convolution_b200 import mesh, ops  # noqa: E402

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def ev_ms(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


rs = np.random.RandomState(0)
for n0, n1, ns in ((10242, 10242, 500), (100000, 100000, 500), (20000, 20000, 0), (100000, 100000, 0)):
    p0, p1 = T(rs.rand(1, n0, 3).astype(np.float32)), T(rs.rand(1, n1, 3).astype(np.float32))
    i0 = T(rs.randint(0, n0, ns).astype(np.int32)) if ns else None
    i1 = T(rs.randint(0, n1, ns).astype(np.int32)) if ns else None
    mode = 1 if ns else 0
    ms = ev_ms(lambda: ops.point_set_loss(p0, p1, i0, i1, mode, need_grad=True))
    pairs = (ns * n1 + ns * n0) if ns else 2 * n0 * n1
    print("point_set_loss mode %d  n0 %6d n1 %6d samples %4d: %.3f ms  (%.1f G pair distances/s)"
          % (mode, n0, n1, ns, ms, pairs / ms / 1e6))

# backward of update_position_MS on an icosphere-5 mesh (20 480 faces, 10 242 vertices), finest scale, 20 sweeps
V, F = mesh.icosphere(5)
vf = mesh.vertex_faces(F, 25, V.shape[0]).astype(np.int32)
x, n = T(V.astype(np.float32)), T(mesh.face_normals(V, F).astype(np.float32))
faces, vft = T(F.astype(np.int32)), T(vf)
lists = ops.vertex_update_ms_lists(faces, vft, V.shape[0], 0, 2)
g = torch.randn_like(x)
t0 = time.perf_counter()
ms_f = ev_ms(lambda: ops.vertex_update_ms(x, n, faces, vft, 0, 2, 20))
ms_b = ev_ms(lambda: ops.vertex_update_ms_bwd(g, x, n, faces, vft, 0, 2, 20, lists))
print("update_position_MS scale 0, 20 sweeps, %d vertices: forward %.3f ms, backward %.3f ms" % (V.shape[0], ms_f, ms_b))
