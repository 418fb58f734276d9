"""Scratch GPU check of the planned (dense-assignment tcgen05) forward against the oracle."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from facet_graph_convolution_b200 import ops, mesh
from oracle import closed_form as cf

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def params(rs, M=8, Cin=64, Cout=64):
    return ((rs.randn(M, Cout, Cin) * 0.05).astype(np.float32), (rs.randn(Cout) * 0.01).astype(np.float32),
            (rs.randn(M, Cin) * 0.05).astype(np.float32), (rs.randn(M, Cin) * 0.05).astype(np.float32),
            (rs.randn(M) * 0.05).astype(np.float32))


def plan_ref(adj, M=8):
    B, N, K = adj.shape
    TF = 128 // M
    rows = B * N
    a = adj.reshape(rows, K)
    nt = (rows + TF - 1) // TF
    R = np.zeros(nt, np.int32); pair = np.zeros((rows, K), np.uint16); prow = []
    for t in range(nt):
        ids = {}
        r0, r1 = t * TF, min(rows, (t + 1) * TF)
        g = np.full((r1 - r0, K), -1, np.int64)
        for r in range(r0, r1):
            for k in range(K):
                i = a[r, k]
                if 0 < i <= N: g[r - r0, k] = (r // N) * N + i - 1
        d = np.unique(g[g >= 0]); R[t] = len(d); prow.append(d)
        for r in range(r0, r1):
            seen = set()
            for k in range(K):
                v = g[r - r0, k]
                if v < 0: continue
                li = int(np.searchsorted(d, v))
                mult = 0 if v in seen else int((g[r - r0] == v).sum())
                seen.add(v)
                pair[r, k] = li | (mult << 9) | 0x8000
    return R, pair, prow


def check(name, x, adj, P, bias_mask=True, act=0):
    W0, b, u, v, c = P
    ops.ConvPlan.MAX_MEAN_ROWS = 1e9
    plan = ops.ConvPlan(T(adj), W0.shape[0])
    assert plan.buf is not None
    torch.cuda.synchronize()
    y = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(u), T(v), T(c), bias_mask=bias_mask, act=act, plan=plan)
    torch.cuda.synchronize()
    y2 = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(u), T(v), T(c), bias_mask=bias_mask, act=act)
    ref = cf.conv_fwd(x, adj, W0, b, u, v, c, bias_mask=bias_mask)
    if act: ref = cf.lrelu(ref, 0.1)
    e = np.abs(y.cpu().numpy() - ref).max(); e2 = np.abs(y2.cpu().numpy() - ref).max()
    print("%-28s planned err %.3g   (old tc path err %.3g)  |ref|max %.3g" % (name, e, e2, np.abs(ref).max()), flush=True)
    return e


def main():
    rs = np.random.RandomState(0)
    # plan parity on a small random graph
    B, N, K = 2, 100, 16
    adj = rs.randint(0, N + 1, size=(B, N, K)).astype(np.int32); adj[:, :, 0] = np.arange(1, N + 1); adj[0, 7] = 0
    adj[1, 3, 5] = adj[1, 3, 2]
    ops.ConvPlan.MAX_MEAN_ROWS = 1e9
    plan = ops.ConvPlan(T(adj), 8); torch.cuda.synchronize()
    buf = plan.buf.cpu().numpy()
    rows = B * N; TF = 16; nt = (rows + TF - 1) // TF
    al = lambda v: (v + 255) // 256 * 256
    o = 256; offR = o; o = al(o + nt * 4); offI = o; o = al(o + rows * 4); offRow = o; o = al(o + nt * TF * K * 4); offP = o
    R = buf[offR:offR + nt * 4].view(np.int32)
    blk = buf[offP:offP + nt * (16 + TF * K * 2)].reshape(nt, 16 + TF * K * 2)
    assert np.array_equal(blk[:, :4].copy().view(np.int32).reshape(-1), R)
    pair = blk[:, 16:].copy().view(np.uint16).reshape(nt * TF, K)[:rows]
    prow = buf[offRow:offRow + nt * TF * K * 4].view(np.int32).reshape(nt, TF * K)
    inv = buf[offI:offI + rows * 4].view(np.float32)
    Rr, pr, prr = plan_ref(adj)
    assert np.array_equal(R, Rr), (R, Rr)
    assert np.array_equal(pair, pr)
    for t in range(nt): assert np.array_equal(prow[t, :R[t]], prr[t])
    cnt = (adj.reshape(rows, K) != 0).sum(1)
    assert np.array_equal(inv, np.where(cnt > 0, 1.0 / np.maximum(cnt, 1), 0).astype(np.float32))
    print("plan bit-exact", flush=True)
    errs = []
    x = rs.randn(B, N, 64).astype(np.float32)
    errs.append(check("random N=100 B=2 (multi-chunk)", x, adj, params(rs)))
    # mesh-like, single chunk
    _, F = mesh.grid_mesh(16, 8, torus=True, morton=True)
    a = mesh.faces_large_adj(F, 16)[None]
    x = rs.randn(1, a.shape[1], 64).astype(np.float32)
    errs.append(check("torus 16x8 mesh adj", x, a, params(rs)))
    errs.append(check("torus dedup + lrelu", x, mesh.dedup_adj(a[0])[None], params(rs), act=1))
    errs.append(check("torus nomask", x, a, params(rs), bias_mask=False))
    # ragged size, K=23
    N = 1237; K = 23
    adj = rs.randint(0, N + 1, size=(1, N, K)).astype(np.int32); adj[:, :, 0] = np.arange(1, N + 1); adj[0, 5] = 0
    adj[0, :, 12:] = np.where(rs.rand(N, K - 12) < 0.6, 0, adj[0, :, 12:])
    x = (rs.randn(1, N, 64) * 3).astype(np.float32)
    errs.append(check("random N=1237 K=23", x, adj, params(rs)))
    # timing at scale
    n = 1_000_000
    _, F = mesh.grid_mesh(1000, 500, torus=True, morton=True)
    a = T(mesh.faces_large_adj(F, 16)[None])
    P = [T(t) for t in params(rs)]
    W0, b, u, v, c = P
    x = torch.randn(1, n, 64, device=dev)
    t0 = time.time(); plan = ops.ConvPlan(a, 8); torch.cuda.synchronize(); print("plan build %.2f ms" % ((time.time() - t0) * 1e3))
    for use in (plan, None):
        for _ in range(3): y = ops.conv_fwd(x, a, W0, b, u, v, c, plan=use)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y = ops.conv_fwd(x, a, W0, b, u, v, c, plan=use)
        e1.record(); torch.cuda.synchronize()
        print("fwd (incl. pre-passes) %s: %.3f ms" % ("planned" if use is not None else "old tc ", e0.elapsed_time(e1) / 5), flush=True)
        if use is not None: yp = y
    print("planned vs old at scale: max diff %.3g" % (yp - y).abs().max().item())
    assert max(errs) < 1e-5, errs
    print("OK")


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def check_bwd(name, x, adj, P):
    W0, b, u, v, c = P
    rs = np.random.RandomState(7)
    gy = rs.randn(*x.shape[:2], W0.shape[1]).astype(np.float32)
    rev = ops.ReverseAdjacency(T(adj))
    ops.ConvPlan.MAX_MEAN_ROWS = 1e9
    tp = rev.target_plan(W0.shape[0])
    fp = ops.ConvPlan(T(adj), W0.shape[0])
    g1 = ops.conv_bwd(T(gy), T(x), T(adj), rev, T(W0), T(u), T(v), T(c), planned=True, plan=fp)
    g0 = ops.conv_bwd(T(gy), T(x), T(adj), rev, T(W0), T(u), T(v), T(c), planned=False)
    ref = cf.conv_bwd(gy, x, adj, W0, b, u, v, c)
    keys = ["gx", "gW0", "gb", "gu", "gv", "gc"]
    out = []
    for k, a1, a0 in zip(keys, g1, g0):
        r = ref[k]; sc = max(1.0, np.abs(r).max())
        out.append("%s %.2g/%.2g" % (k, np.abs(a1.cpu().numpy() - r).max() / sc, np.abs(a0.cpu().numpy() - r).max() / sc))
    print("%-26s Kr=%s planned/old rel err: %s" % (name, None if tp is None else tp[1], "  ".join(out)), flush=True)
    worst = max(np.abs(a1.cpu().numpy() - ref[k]).max() / max(1.0, np.abs(ref[k]).max()) for k, a1 in zip(keys, g1))
    assert tp is not None and worst < 2e-5, worst


def main_bwd():
    rs = np.random.RandomState(3)
    _, F = mesh.grid_mesh(16, 8, torus=True, morton=True)
    a = mesh.faces_large_adj(F, 16)[None]
    x = rs.randn(1, a.shape[1], 64).astype(np.float32)
    check_bwd("torus mesh adj", x, a, params(rs))
    check_bwd("torus dedup", x, mesh.dedup_adj(a[0])[None], params(rs))
    B, N, K = 2, 300, 12
    adj = rs.randint(0, N + 1, size=(B, N, K)).astype(np.int32); adj[:, :, 0] = np.arange(1, N + 1); adj[0, 7] = 0
    x = rs.randn(B, N, 64).astype(np.float32)
    check_bwd("random N=300 B=2 K=12", x, adj, params(rs))
    # timing at scale
    n = 1_000_000
    _, F = mesh.grid_mesh(1000, 500, torus=True, morton=True)
    a = T(mesh.faces_large_adj(F, 16)[None])
    P = [T(t) for t in params(rs)]
    W0, b, u, v, c = P
    x = torch.randn(1, n, 64, device=dev); gy = torch.randn(1, n, 64, device=dev)
    rev = ops.ReverseAdjacency(a)
    fp = ops.ConvPlan(a, 8)
    for planned in (True, False):
        for _ in range(2): ops.conv_bwd(gy, x, a, rev, W0, u, v, c, planned=planned, plan=fp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.conv_bwd(gy, x, a, rev, W0, u, v, c, planned=planned, plan=fp)
        e1.record(); torch.cuda.synchronize()
        print("bwd %s: %.3f ms" % ("planned" if planned else "old    ", e0.elapsed_time(e1) / 5), flush=True)
    print("BWD OK")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bwd":
    main_bwd()
