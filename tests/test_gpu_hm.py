"""GPU parity of the HMMA-aggregation forward (conv_hm.cu) against the oracle's fp64 closed form for EVERY layer
shape of the reference network (Code/model.py:858-932: M = 9, 32->64, 64->128, 128->128, 128->64, 64->32) at
>= 4 096 rows, K in {16, 23}, plus the M = 8 benchmark layer; mesh and random adjacencies, ragged sizes, batches,
padding rows, repeated ids, fused upsampling.  Tolerance: max-abs <= 1e-5 on O(1) activations."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
TOL_Y = 1e-5


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def _params(rs, M, Cin, Cout):
    return ((rs.randn(M, Cout, Cin) * 0.05).astype(np.float32), (rs.randn(Cout) * 0.01).astype(np.float32),
            (rs.randn(M, Cin) * 0.05).astype(np.float32), (rs.randn(M, Cin) * 0.05).astype(np.float32),
            (rs.randn(M) * 0.05).astype(np.float32))


def _mesh_adj(nx, ny, K, dedup):
    from facet_graph_convolution_b200 import mesh
    _, F = mesh.grid_mesh(nx, ny, torus=True, morton=True)
    a = mesh.faces_large_adj(F, K)
    return (mesh.dedup_adj(a) if dedup else a)[None].astype(np.int32)


def _random_adj(rs, B, N, K):
    adj = rs.randint(0, N + 1, size=(B, N, K)).astype(np.int32)
    adj[:, :, 0] = np.arange(1, N + 1)
    adj[:, :, K // 2:] = np.where(rs.rand(B, N, K - K // 2) < 0.5, 0, adj[:, :, K // 2:])
    adj[0, 7] = 0                       # a facet without any neighbour (cnt = 0)
    adj[0, 12, 2] = adj[0, 12, 1]       # repeated id
    return adj


SHAPES = [(32, 64), (64, 128), (128, 128), (128, 64), (64, 32)]


@pytest.mark.parametrize("Cin,Cout", SHAPES)
@pytest.mark.parametrize("K", [16, 23])
def test_network_layer_shapes_at_size_mesh(Cin, Cout, K):
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(Cin * 7 + Cout + K)
    adj = _mesh_adj(52, 40, K, dedup=(K == 23))       # 4 160 facets: 130 tiles of 32
    N = adj.shape[1]
    assert N >= 4096
    x = rs.randn(1, N, Cin).astype(np.float32)
    W0, b, u, v, c = _params(rs, 9, Cin, Cout)
    y = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(u), T(v), T(c)).cpu().numpy()
    ref = cf.conv_fwd(x, adj, W0, b, u, v, c)
    assert np.abs(y - ref).max() < TOL_Y


@pytest.mark.parametrize("Cin,Cout", SHAPES + [(64, 64)])
def test_network_layer_shapes_at_size_random_ragged(Cin, Cout):
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(Cin + Cout)
    B, N, K = 2, 2077, 23                             # 4 154 rows, not a multiple of the 32-facet tile
    adj = _random_adj(rs, B, N, K)
    x = (rs.randn(B, N, Cin) * 2).astype(np.float32)
    W0, b, u, v, c = _params(rs, 9, Cin, Cout)
    for mask, act in ((True, 0), (False, 1)):
        y = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(u), T(v), T(c), bias_mask=mask, act=act, alpha=0.1).cpu().numpy()
        ref = cf.conv_fwd(x, adj, W0, b, u, v, c, bias_mask=mask)
        if act:
            ref = cf.lrelu(ref, 0.1)
        assert np.abs(y - ref).max() < TOL_Y * 2      # |x| ~ 2 here


def test_m8_layer_without_plan():
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(5)
    adj = _mesh_adj(52, 40, 16, dedup=False)
    N = adj.shape[1]
    x = rs.randn(1, N, 64).astype(np.float32)
    W0, b, u, v, c = _params(rs, 8, 64, 64)
    y = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(u), T(v), T(c)).cpu().numpy()
    assert np.abs(y - cf.conv_fwd(x, adj, W0, b, u, v, c)).max() < TOL_Y


@pytest.mark.parametrize("Cin,Cout", [(128, 64), (64, 32)])
def test_fused_upsampling_at_size(Cin, Cout):
    """custom_conv2d(custom_upsampling(h, 2), adj) (model.py:902-905, 923-926) with the repeat as an index shift."""
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(Cin)
    adj = _mesh_adj(52, 40, 16, dedup=True)
    N = adj.shape[1]
    xc = rs.randn(1, N // 4, Cin).astype(np.float32)
    W0, b, u, v, c = _params(rs, 9, Cin, Cout)
    y = ops.conv_fwd_up(T(xc), T(adj), T(W0), T(b), T(u), T(v), T(c), upshift=2, act=1, alpha=0.1)
    assert y is not None
    ref = cf.lrelu(cf.conv_fwd(cf.upsample(xc, 2), adj, W0, b, u, v, c), 0.1)
    assert np.abs(y.cpu().numpy() - ref).max() < TOL_Y


def test_extreme_logit_spread_recentres_the_softmax():
    """logits spread over hundreds of units (the pre-pass's row-wise shifts alone would let exp2 underflow for every
    weight at once): the kernel re-centres per (facet, neighbour) exactly like the reference's softmax."""
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(3)
    adj = _mesh_adj(40, 28, 16, dedup=True)
    N = adj.shape[1]
    W0, b, u, v, c = _params(rs, 9, 64, 32)
    u, v = (u * 60).astype(np.float32), (v * 60).astype(np.float32)
    x = rs.randn(1, N, 64).astype(np.float32)
    lg = np.einsum("nc,mc->nm", x[0].astype(np.float64), u.astype(np.float64))
    assert (lg.max(1) - lg.min(1)).max() > 100          # the case really is extreme
    y = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(u), T(v), T(c)).cpu().numpy()
    ref = cf.conv_fwd(x, adj, W0, b, u, v, c)
    assert np.isfinite(y).all()
    assert np.abs(y - ref).max() < 2e-5                 # logits of magnitude ~100 carry ~1e-5 of fp32 rounding themselves


def test_large_logit_spread_and_scale():
    """softmax with logits spread over +-60 and activations of magnitude 1e3 / 1e-3 (exact power-of-two scaling)."""
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(9)
    adj = _mesh_adj(40, 28, 16, dedup=True)
    N = adj.shape[1]
    W0, b, u, v, c = _params(rs, 9, 64, 32)
    u, v = u * 8, v * 8
    for scale in (1e3, 1e-3):
        x = (rs.randn(1, N, 64) * scale).astype(np.float32)
        uu, vv = (u / scale).astype(np.float32), (v / scale).astype(np.float32)
        y = ops.conv_fwd(T(x), T(adj), T(W0), T(b), T(uu), T(vv), T(c)).cpu().numpy()
        ref = cf.conv_fwd(x, adj, W0, b, uu, vv, c)
        assert np.abs(y - ref).max() < TOL_Y * max(scale, 1.0) * 4


def test_repeated_launches_are_bit_identical_at_size():
    """The aggregator warps hand a tile to the tensor core with an mbarrier arrive only (the generic->async proxy fence is
    executed by the issuing warp) and refill their row stages while the previous facet is still being contracted: any
    ordering hole in that pipeline shows up as run-to-run differences.  80 launches over 150 k rows (thousands of tiles per
    launch, every CTA with ~30 tiles in flight order) must agree bit for bit, and with the fp64 closed form."""
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(11)
    adj = _mesh_adj(300, 250, 16, dedup=False)          # 150 000 facets
    N = adj.shape[1]
    for Cin, Cout in ((64, 32), (128, 64)):
        x = rs.randn(1, N, Cin).astype(np.float32)
        W0, b, u, v, c = _params(rs, 9, Cin, Cout)
        tx, ta = T(x), T(adj)
        tp = [T(a) for a in (W0, b, u, v, c)]
        y0 = ops.conv_fwd(tx, ta, *tp, act=1, alpha=0.1)
        for _ in range(40):
            assert torch.equal(ops.conv_fwd(tx, ta, *tp, act=1, alpha=0.1), y0)
        ref = cf.lrelu(cf.conv_fwd(x[:, :4096], adj[:, :4096].clip(0, 4096), W0, b, u, v, c), 0.1)
        sub = ops.conv_fwd(T(x[:, :4096]), T(adj[:, :4096].clip(0, 4096)), *tp, act=1, alpha=0.1).cpu().numpy()
        assert np.abs(sub - ref).max() < TOL_Y
