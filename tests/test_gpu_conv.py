"""Parity of the CUDA facet-graph convolution (through the Python facade -> C ABI) against the
golden vectors produced by the reference source and against the oracle.  Needs a B200."""
import numpy as np
import pytest
import torch

from conftest import CONV_CASES, golden
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu

# per-layer tolerance on O(1) activations: max-abs <= 1e-5 on y
TOL_Y = 1e-5
# gradients are sums over up to N*K terms of O(1) values: relative to the largest entry
TOL_G = 2e-5


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _mode(g):
    trans = bool(int(g["translation"]))
    v = -g["u"] if trans else g["v"]
    return trans, v


@pytest.mark.parametrize("name", CONV_CASES)
def test_gathered_neighbours_bit_exact(name):
    from facet_graph_convolution_b200 import ops
    g = golden(name)
    out = ops.gather_rows(T(g["x"]), T(g["adj"])).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), g["xg"].view(np.uint32))


@pytest.mark.parametrize("name", CONV_CASES)
def test_assignments(name):
    from facet_graph_convolution_b200 import ops
    g = golden(name)
    _, v = _mode(g)
    q = ops.assignments(T(g["x"]), T(g["adj"]), T(g["u"]), T(v), T(g["c"])).cpu().numpy()
    assert np.abs(q - g["q"]).max() < 2e-6
    assert np.abs(q.sum(-1) - 1).max() < 1e-5


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv_forward(name):
    from facet_graph_convolution_b200 import ops
    g = golden(name)
    _, v = _mode(g)
    y = ops.conv_fwd(T(g["x"]), T(g["adj"]), T(g["W0"]), T(g["b"]), T(g["u"]), T(v), T(g["c"]),
                     bias_mask=bool(int(g["bias_mask"]))).cpu().numpy()
    assert np.abs(y - g["y"]).max() < TOL_Y
    # fused leaky-ReLU epilogue
    ya = ops.conv_fwd(T(g["x"]), T(g["adj"]), T(g["W0"]), T(g["b"]), T(g["u"]), T(v), T(g["c"]),
                      bias_mask=bool(int(g["bias_mask"])), act=ops.ACT_LRELU, alpha=0.1).cpu().numpy()
    assert np.abs(ya - cf.lrelu(g["y"], np.float32(0.1))).max() < TOL_Y
    # int64 adjacency is accepted (the reference's preprocessing emits int64)
    y64 = ops.conv_fwd(T(g["x"]), T(g["adj"].astype(np.int64)), T(g["W0"]), T(g["b"]), T(g["u"]), T(v), T(g["c"]),
                       bias_mask=bool(int(g["bias_mask"]))).cpu().numpy()
    assert np.array_equal(y64, y)


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv_backward(name):
    from facet_graph_convolution_b200 import ops
    g = golden(name)
    trans, v = _mode(g)
    adj = T(g["adj"])
    rev = ops.ReverseAdjacency(adj)
    out = ops.conv_bwd(T(g["gy"]), T(g["x"]), adj, rev, T(g["W0"]), T(g["u"]), T(v), T(g["c"]),
                       bias_mask=bool(int(g["bias_mask"])))
    gx, gW0, gb, gu, gv, gc = (t.cpu().numpy() for t in out)
    if trans:  # v = -u  =>  d/du total = gu - gv
        gu = gu - gv
    for nm, got, ref in (("gx", gx, g["gx"]), ("gW0", gW0, g["gW0"]), ("gb", gb, g["gb"]), ("gu", gu, g["gu"]),
                         ("gc", gc, g["gc"])):
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(got - ref).max() / scale < TOL_G, nm
    if not trans:
        assert np.abs(gv - g["gv"]).max() / max(1.0, float(np.abs(g["gv"]).max())) < TOL_G


def test_reverse_adjacency_is_exact_and_sorted():
    from facet_graph_convolution_b200 import ops
    g = golden("conv_64_64_M8_K16_B2")
    adj = g["adj"]
    B, N, K = adj.shape
    rev = ops.ReverseAdjacency(T(adj))
    ptr = rev.ptr.cpu().numpy()
    edge = rev.edge.cpu().numpy()[: rev.nnz]
    flat = adj.reshape(-1)
    e = np.nonzero(flat)[0]
    tgt = (e // K // N) * N + flat[e] - 1
    order = np.lexsort((e, tgt))
    assert rev.nnz == e.size
    assert np.array_equal(edge, e[order].astype(np.int32))
    assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(np.bincount(tgt, minlength=B * N))]).astype(np.int32))


def test_backward_is_bit_reproducible():
    from facet_graph_convolution_b200 import ops
    g = golden("conv_64_64_M8_K16_B2")
    adj = T(g["adj"])
    rev = ops.ReverseAdjacency(adj)
    args = (T(g["gy"]), T(g["x"]), adj, rev, T(g["W0"]), T(g["u"]), T(g["v"]), T(g["c"]))
    a = [t.cpu().numpy() for t in ops.conv_bwd(*args)]
    b = [t.cpu().numpy() for t in ops.conv_bwd(*args)]
    for p, q in zip(a, b):
        assert np.array_equal(p.view(np.uint32), q.view(np.uint32))


def test_conv_variants_match_reference():
    from facet_graph_convolution_b200 import model as fm
    g = golden("conv_variants")
    x, adj = T(g["x"]), T(g["adj"])
    for t in (0, 1):
        tag = "posassign_t%d_" % t
        params = [g[tag + k] for k in (["W0", "b", "u", "c"] + ([] if t else ["vn"]))]
        with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=params)):
            y, _ = fm.custom_conv2d_pos_for_assignment(x, adj, 8, 4, translation_invariance=bool(t))
        assert np.abs(y.cpu().numpy() - g[tag + "y"]).max() < TOL_Y
        tag = "onlypos_t%d_" % t
        params = [g[tag + k] for k in (["W0", "b", "u", "c"] + ([] if t else ["v"]))]
        with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=params)):
            y, _ = fm.custom_conv2d_only_pos_for_assignment(x, adj, 8, 4, translation_invariance=bool(t))
        assert np.abs(y.cpu().numpy() - g[tag + "y"]).max() < TOL_Y


def test_custom_conv2d_signature_and_autograd():
    """custom_conv2d keeps the reference's signature/returns and its autograd matches the golden."""
    from facet_graph_convolution_b200 import model as fm
    g = golden("conv_6_32_M9_K23_B2")
    params = [g[k] for k in ("W0", "b", "u", "c", "v")]
    store = fm.VariableStore(dev(), params=params, requires_grad=True)
    x = T(g["x"]).requires_grad_(True)
    with fm.variable_store(store):
        y, aux = fm.custom_conv2d(x, T(g["adj"]), 32, 9)
    assert len(aux) == 3 and tuple(aux[0].shape) == (9, 32, 6)
    (y * T(g["gy"])).sum().backward()
    assert np.abs(y.detach().cpu().numpy() - g["y"]).max() < TOL_Y
    assert np.abs(x.grad.cpu().numpy() - g["gx"]).max() / max(1.0, np.abs(g["gx"]).max()) < TOL_G
    for t, k in zip(store.params, ("gW0", "gb", "gu", "gc", "gv")):
        assert np.abs(t.grad.cpu().numpy() - g[k]).max() / max(1.0, np.abs(g[k]).max()) < TOL_G, k


def test_error_behaviour():
    from facet_graph_convolution_b200 import _lib, ops
    x = torch.zeros(1, 8, 6, device=dev())
    adj = torch.zeros(1, 8, 40, dtype=torch.int32, device=dev())  # K > FGC_MAX_K
    W0 = torch.zeros(2, 5, 6, device=dev())
    with pytest.raises(_lib.FacetConvError):
        ops.conv_fwd(x, adj, W0, torch.zeros(5, device=dev()), torch.zeros(2, 6, device=dev()),
                     torch.zeros(2, 6, device=dev()), torch.zeros(2, device=dev()))
    with pytest.raises(_lib.FacetConvError):  # shape mismatch
        ops.conv_fwd(x, adj[:, :, :4], torch.zeros(2, 5, 7, device=dev()), torch.zeros(5, device=dev()),
                     torch.zeros(2, 6, device=dev()), torch.zeros(2, 6, device=dev()), torch.zeros(2, device=dev()))
    # out-of-range neighbour ids are clamped to padding, never read out of bounds
    adj2 = torch.full((1, 8, 4), 1000, dtype=torch.int32, device=dev())
    y = ops.conv_fwd(x + 1, adj2, W0 + 1, torch.ones(5, device=dev()), torch.zeros(2, 6, device=dev()),
                     torch.zeros(2, 6, device=dev()), torch.zeros(2, device=dev()))
    assert torch.isfinite(y).all()


def test_c2_shape_properties_at_scale():
    """Size-independent properties at a C2-like shape (N = 200k keeps the test short): linearity in W,
    fake rows give the bias, permutation of neighbour slots does not change y."""
    from facet_graph_convolution_b200 import ops
    torch.manual_seed(0)
    d = dev()
    N, K, M, C = 200_000, 16, 8, 64
    x = torch.randn(1, N, C, device=d)
    adj = torch.randint(1, N + 1, (1, N, K), device=d, dtype=torch.int32)
    adj[0, :, 0] = torch.arange(1, N + 1, device=d, dtype=torch.int32)
    adj[0, ::1000, 1:] = 0
    x[0, ::1000] = 0  # fake nodes: zero features, self-only adjacency
    W1 = torch.randn(M, C, C, device=d) * 0.05
    W2 = torch.randn(M, C, C, device=d) * 0.05
    b = torch.randn(C, device=d) * 0.01
    z = torch.zeros(C, device=d)
    u = torch.randn(M, C, device=d) * 0.05
    v = torch.randn(M, C, device=d) * 0.05
    c = torch.randn(M, device=d) * 0.05
    y1 = ops.conv_fwd(x, adj, W1, b, u, v, c)
    y2 = ops.conv_fwd(x, adj, W2, z, u, v, c)
    y12 = ops.conv_fwd(x, adj, W1 + W2, b, u, v, c)
    assert (y1 + y2 - y12).abs().max().item() < 2e-5
    # fake rows: s = q*0 = 0 => y = b exactly
    assert torch.equal(y1[0, ::1000], b.expand_as(y1[0, ::1000]))
    perm = torch.cat([torch.zeros(1, dtype=torch.long), 1 + torch.randperm(K - 1)]).to(d)
    yp = ops.conv_fwd(x, adj[:, :, perm].contiguous(), W1, b, u, v, c)
    assert (yp - y1).abs().max().item() < 2e-5
    # oracle spot check on a random subset of rows (the oracle gathers from the full x)
    rows = np.random.RandomState(1).choice(N, 64, replace=False)
    xs, adjs = x.cpu().numpy(), adj.cpu().numpy()
    xg = cf.gather_rows(xs, adjs[:, rows])
    un, vn, cn = (t.cpu().numpy().astype(np.float64) for t in (u, v, c))
    a = (xs[0, rows].astype(np.float64) @ un.T)[:, None, :] + np.einsum("nkc,mc->nkm", xg[0].astype(np.float64), vn) + cn
    e = np.exp(a - a.max(-1, keepdims=True))
    q = e / e.sum(-1, keepdims=True)
    s = np.einsum("nkm,nkc->nmc", q, xg[0].astype(np.float64))
    cnt = (adjs[0, rows] != 0).sum(-1)
    yref = np.einsum("moc,nmc->no", W1.cpu().numpy().astype(np.float64), s) / cnt[:, None] + b.cpu().numpy()
    assert np.abs(y1[0, rows].cpu().numpy() - yref).max() < TOL_Y


def _host_fwd_bwd(x, adj, gy, W0, b, u, v, c, bias_mask=True):
    """One call of the host-buffer C-ABI entry point (what bench.py's e2e arm times)."""
    import ctypes as C
    from facet_graph_convolution_b200 import _lib
    L = _lib.lib()
    B, N, Cin = x.shape
    M, Cout, Cw = W0.shape
    s = _lib.ConvShape(B, N, adj.shape[2], Cin, Cw, 0, u.shape[1], Cout, M)
    ins = [np.ascontiguousarray(a) for a in (x, adj.astype(np.int32), gy, W0, b, u, v, c)]
    outs = [np.empty((B, N, Cout), np.float32), np.empty_like(ins[0]), np.empty_like(ins[3]), np.empty_like(ins[4]),
            np.empty_like(ins[5]), np.empty_like(ins[6]), np.empty_like(ins[7])]
    P = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(L.fgc_conv_fwd_bwd_host(C.byref(s), *[P(a) for a in ins], *[P(a) for a in outs], int(bias_mask), 0),
               "fgc_conv_fwd_bwd_host")
    L.fgc_host_release()
    return outs


def test_host_entry_point_matches_oracle():
    """fgc_conv_fwd_bwd_host (host buffers in, host buffers out) on: a dense 64->64 M=8 layer over a mesh
    adjacency (planned tensor-core path, caches built inside the call), the same layer over a random
    adjacency (plan dropped: too many distinct rows per tile), and an M=9 layer (FFMA path)."""
    from facet_graph_convolution_b200 import mesh
    rs = np.random.RandomState(5)
    _, F = mesh.grid_mesh(20, 10, torus=True, morton=True)
    a_mesh = mesh.faces_large_adj(F, 16)[None]
    N = 300
    a_rand = rs.randint(0, N + 1, size=(1, N, 16)).astype(np.int32)
    a_rand[:, :, 0] = np.arange(1, N + 1)
    a_rand[0, 11] = 0
    for name, adj, Cin, Cout, M in (("mesh", a_mesh, 64, 64, 8), ("random", a_rand, 64, 64, 8),
                                    ("m9", a_rand, 32, 64, 9)):
        Nn = adj.shape[1]
        x = rs.randn(1, Nn, Cin).astype(np.float32)
        gy = rs.randn(1, Nn, Cout).astype(np.float32)
        W0 = (rs.randn(M, Cout, Cin) * 0.05).astype(np.float32)
        b = (rs.randn(Cout) * 0.01).astype(np.float32)
        u = (rs.randn(M, Cin) * 0.05).astype(np.float32)
        v = (rs.randn(M, Cin) * 0.05).astype(np.float32)
        c = (rs.randn(M) * 0.05).astype(np.float32)
        y, gx, gW0, gb, gu, gv, gc = _host_fwd_bwd(x, adj, gy, W0, b, u, v, c)
        assert np.abs(y - cf.conv_fwd(x, adj, W0, b, u, v, c)).max() < TOL_Y, name
        ref = cf.conv_bwd(gy, x, adj, W0, b, u, v, c)
        for k, got in (("gx", gx), ("gW0", gW0), ("gb", gb), ("gu", gu), ("gv", gv), ("gc", gc)):
            sc = max(1.0, float(np.abs(ref[k]).max()))
            assert np.abs(got - ref[k]).max() / sc < TOL_G, (name, k)
