"""Parity of pooling/unpooling/linear/normalisation/loss/vertex-update kernels and of the whole
multi-scale network against golden vectors produced by the reference source.  Needs a B200."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _params(g):
    return [g["p%02d" % i] for i in range(int(g["nparams"]))]


def test_small_ops():
    from facet_graph_convolution_b200 import model as fm
    from facet_graph_convolution_b200 import ops
    g = golden("small_ops")
    x, xz = T(g["x"]), T(g["xz"])
    assert np.array_equal(fm.custom_binary_tree_pooling(x, steps=2).cpu().numpy(), g["pool_max2"])
    assert np.array_equal(fm.custom_binary_tree_pooling(x, steps=1).cpu().numpy(), g["pool_max1"])
    aiz = fm.custom_binary_tree_pooling(xz, steps=2, pooltype="avg_ignore_zeros").cpu().numpy()
    assert np.abs(aiz - g["pool_aiz2"]).max() < 1e-7
    assert np.array_equal(fm.custom_upsampling(x, steps=2).cpu().numpy(), g["up2"])
    assert np.abs(fm.lrelu(x, 0.1).cpu().numpy() - g["lrelu"]).max() < 1e-7
    with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=[g["lin_W"], g["lin_b"]])):
        y = fm.custom_lin(x, 7)
    assert np.abs(y.cpu().numpy() - g["lin_y"]).max() < 2e-6
    n = fm.normalizeTensor(T(g["norm_in"])).cpu().numpy()
    assert np.abs(n - g["norm_out"]).max() < 2e-6
    loss = fm.faceNormalsLoss(T(g["loss_fn"]), T(g["loss_gt"])).item()
    assert abs(loss - float(g["loss"])) < 1e-4            # degrees; measured 2e-6 (tests/micro/tolerance_probe.py)
    # gradient of loss(normalizeTensor(n)) through both backward kernels
    nt = T(g["norm_in"]).requires_grad_(True)
    fm.faceNormalsLoss(fm.normalizeTensor(nt), T(g["loss_gt"])).backward()
    ref = g["loss_norm_grad"]
    assert np.abs(nt.grad.cpu().numpy() - ref).max() / np.abs(ref).max() < 1e-5     # measured 2e-7
    # concat / split / permutation gather are exact copies
    a, b = torch.randn(2, 10, 5, device=dev()), torch.randn(2, 10, 3, device=dev())
    cat = ops.concat2(a, b)
    assert torch.equal(cat, torch.cat([a, b], -1))
    ga, gb = ops.split2(cat, 5)
    assert torch.equal(ga, a) and torch.equal(gb, b)
    idx = torch.randperm(20, device=dev()).to(torch.int32)
    assert torch.equal(ops.gather_perm(a.reshape(20, 5), idx), a.reshape(20, 5)[idx.long()])


def test_pool_and_upsample_backward_match_autograd_semantics():
    from facet_graph_convolution_b200 import ops
    x = torch.randn(1, 32, 6, device=dev())
    x[0, 4:8] = 0.25  # ties: TF/torch split the gradient equally among tied maxima
    y = ops.pool_max(x, 4)
    gy = torch.randn_like(y)
    gx = ops.pool_max_bwd(gy, x, y, 4)
    xr = x.clone().requires_grad_(True)
    xr.reshape(1, 8, 4, 6).amax(dim=2).backward(gy)
    assert torch.allclose(gx, xr.grad, atol=1e-7)
    gu = torch.randn(1, 32, 6, device=dev())
    assert torch.allclose(ops.upsample_bwd(gu, 4), gu.reshape(1, 8, 4, 6).sum(2), atol=1e-6)


def test_lin_backward_and_fused_head():
    from facet_graph_convolution_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(1, 333, 32, device=dev())
    W1 = torch.randn(32, 1024, device=dev()) * 0.05
    b1 = torch.randn(1024, device=dev()) * 0.01
    W2 = torch.randn(1024, 3, device=dev()) * 0.05
    b2 = torch.randn(3, device=dev()) * 0.01
    h = ops.lin_fwd(x, W1, b1, ops.ACT_LRELU, 0.1)
    y_unfused = ops.lin_fwd(h, W2, b2)
    y_fused = ops.mlp_head(x, W1, b1, W2, b2, 0.1)
    ref = cf.lin(cf.lrelu(cf.lin(x.cpu().numpy().astype(np.float64), W1.cpu().numpy(), b1.cpu().numpy()), 0.1),
                 W2.cpu().numpy(), b2.cpu().numpy())
    assert np.abs(y_unfused.cpu().numpy() - ref).max() < 1e-5
    assert np.abs(y_fused.cpu().numpy() - ref).max() < 1e-5
    gy = torch.randn(1, 333, 1024, device=dev())
    gx, gW, gb = ops.lin_bwd(gy, x, W1)
    x64, g64, W64 = (t.cpu().double() for t in (x, gy, W1))
    assert (gx.cpu().double() - g64 @ W64.T).abs().max() < 1e-4
    assert (gW.cpu().double() - x64[0].T @ g64[0]).abs().max() < 2e-4
    assert (gb.cpu().double() - g64[0].sum(0)).abs().max() < 2e-4


def test_fused_head_on_tensor_cores_matches_float64():
    """The 32 -> 1024 -> 3 head above 2 048 rows runs on tcgen05 (per-row scaled fp16 hi/lo operands, two
    passes of 512 hidden units): y within 1e-5 of the float64 closed form, rows of very different magnitude,
    an all-zero row and a ragged last tile included."""
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(11)
    rows = 128 * 37 + 45
    x = (rs.randn(1, rows, 32) * 3).astype(np.float32)
    x[0, 5] = 0.0
    x[0, 6] *= 1e-4
    x[0, 7] *= 50.0
    W1 = (rs.randn(32, 1024) * 0.05).astype(np.float32)
    b1 = (rs.randn(1024) * 0.01).astype(np.float32)
    W2 = (rs.randn(1024, 3) * 0.05).astype(np.float32)
    b2 = (rs.randn(3) * 0.01).astype(np.float32)
    y = ops.mlp_head(T(x), T(W1), T(b1), T(W2), T(b2), 0.1).cpu().numpy()
    ref = cf.lin(cf.lrelu(cf.lin(x.astype(np.float64), W1, b1), 0.1), W2, b2)
    scale = np.maximum(1.0, np.abs(ref).max(axis=-1, keepdims=True))
    assert (np.abs(y - ref) / scale).max() < 1e-5
    y2 = ops.mlp_head(T(x), T(W1), T(b1), T(W2), T(b2), 0.1).cpu().numpy()
    assert np.array_equal(y, y2)


def test_network_single_scale_pipeline():
    """C1-shaped pipeline on the reference's own preprocessing of a noisy icosphere-3."""
    from facet_graph_convolution_b200 import model as fm
    from facet_graph_convolution_b200 import ops
    g = golden("net_icosphere3")
    adjs = [T(g["adj0"]), T(g["adj1"]), T(g["adj2"])]
    for fuse in (True, False):
        with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=_params(g))):
            y = fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0, fuse=fuse)
        assert np.abs(y.cpu().numpy() - g["y_raw"]).max() < 1e-5
    yn = fm.normalizeTensor(y)
    assert np.abs(yn.cpu().numpy() - g["y_norm"]).max() < 1e-4
    out = ops.gather_perm(yn.reshape(-1, 3), T(g["perm"]))[: int(g["nreal"])].cpu().numpy()
    pred = cf.host_normalize(out)  # float64 two-pass normalise stays on the host (train.py:136)
    assert np.abs(pred - g["pred_normals"]).max() < 1e-4
    ang = cf.angular_diff_vec(pred, g["pred_normals"])
    assert ang.mean() < 0.01 + np.degrees(np.arccos(0.999999))
    xo = fm.update_position2(T(g["verts_in"]), T(g["pred_normals"][None].astype(np.float32)), T(g["e_map"]),
                             T(g["v_e_map"]), iter_num=60, max_edges=20)
    assert np.abs(xo.cpu().numpy() - g["verts_out"]).max() < 1e-4


def test_network_multi_scale_pipeline():
    from facet_graph_convolution_b200 import model as fm
    g = golden("net_ms_icosphere2")
    adjs = [T(g["adj0"]), T(g["adj1"]), T(g["adj2"])]
    with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=_params(g))):
        ys = fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0, multiScale=True)
    for y, k in zip(ys, ("y0", "y1", "y2")):
        assert np.abs(y.cpu().numpy() - g[k]).max() < 1e-5
    ns = [fm.normalizeTensor(y) for y in ys]
    for n, k in zip(ns, ("n0", "n1", "n2")):
        assert np.abs(n.cpu().numpy() - g[k]).max() < 1e-4
    xo, dxl = fm.update_position_MS(T(g["verts_in"]), [T(g["n0"]), T(g["n1"]), T(g["n2"])], T(g["faces"]),
                                    T(g["v_faces"]), 2, iter_num_list=[int(i) for i in g["iters"]])
    assert len(dxl) == 3
    assert np.abs(xo.cpu().numpy() - g["verts_out"]).max() < 1e-4


def test_network_training_gradients():
    """d loss / d parameters through the whole network against torch autograd of the reference code."""
    from facet_graph_convolution_b200 import model as fm
    g = golden("net_train_small")
    n = int(g["nparams"])
    store = fm.VariableStore(dev(), params=[g["p%02d" % i] for i in range(n)], requires_grad=True)
    adjs = [T(g["adj0"]), T(g["adj1"]), T(g["adj2"])]
    with fm.variable_store(store):
        y = fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0)
    assert np.abs(y.detach().cpu().numpy() - g["y"]).max() < 1e-5
    loss = fm.faceNormalsLoss(fm.normalizeTensor(y), T(g["gt"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-4       # degrees; measured 8e-6 (tests/micro/tolerance_probe.py)
    loss.backward()
    for i, t in enumerate(store.params):
        ref = g["g%02d" % i]
        scale = max(float(np.abs(ref).max()), 1e-3)
        assert t.grad is not None, i
        assert np.abs(t.grad.cpu().numpy() - ref).max() / scale < 1e-4, (i, store.names[i])   # measured 3e-6


def _rand_adj(rs, B, N, K):
    adj = rs.randint(0, N + 1, size=(B, N, K)).astype(np.int32)
    adj[:, :, 0] = np.arange(1, N + 1, dtype=np.int32)[None]
    adj[:, :, K - 3:] *= (rs.rand(B, N, 3) < 0.5)  # ragged neighbourhoods
    return adj


@pytest.mark.parametrize("B,N,Cin,Cout", [(1, 8192, 128, 64), (3, 4096, 64, 32), (1, 4096 + 64, 64, 32)])
def test_fused_upsampling_is_bit_identical_to_the_repeated_input(B, N, Cin, Cout):
    """conv(custom_upsampling(x, 2)) with the repeat folded into the gather (row r reads coarse row r >> 2,
    model.py:817-825 then :427-504) == the same layer on the materialised tensor, bit for bit; batches and a
    ragged last tile included."""
    from facet_graph_convolution_b200 import ops
    rs = np.random.RandomState(5)
    M, K = 9, 16
    xc = rs.randn(B, N // 4, Cin).astype(np.float32)
    adj = _rand_adj(rs, B, N, K)
    W0 = (rs.randn(M, Cout, Cin) * 0.05).astype(np.float32)
    b = (rs.randn(Cout) * 0.05).astype(np.float32)
    u = (rs.randn(M, Cin) * 0.05).astype(np.float32)
    v = (rs.randn(M, Cin) * 0.05).astype(np.float32)
    c = (rs.randn(M) * 0.05).astype(np.float32)
    args = [T(a) for a in (W0, b, u, v, c)]
    y_up = ops.conv_fwd_up(T(xc), T(adj), *args, upshift=2)
    assert y_up is not None, "the M = 9 layers at >= 4096 rows must have the fused path"
    y_mat = ops.conv_fwd(ops.upsample(T(xc), 4), T(adj), *args)
    assert torch.equal(y_up, y_mat)
    ref = cf.conv_fwd(cf.upsample(xc[:1].astype(np.float64), 2), adj[:1], W0, b, u, v, c)
    assert np.abs(y_up[:1].cpu().numpy() - ref).max() < 1e-5
    # shapes without the path say so instead of computing something else (M = 5 has no tensor-core forward)
    a5 = [T(a) for a in (W0[:5], b, u[:5], v[:5], c[:5])]
    assert ops.conv_fwd_up(T(xc), T(adj), *a5, upshift=2) is None


def test_network_with_fused_upsampling_matches_unfused_at_size():
    """The whole network at a size where both upsampling layers take the fused path: identical to fuse=False
    within the per-layer tolerance (the unfused arm materialises every intermediate)."""
    from facet_graph_convolution_b200 import model as fm
    rs = np.random.RandomState(9)
    N0 = 4096 * 16
    adjs = [T(_rand_adj(rs, 1, N0 >> (2 * l), 16)) for l in range(3)]
    x = T(rs.randn(1, N0, 6).astype(np.float32))
    store = fm.VariableStore(dev(), seed=3)
    with torch.no_grad(), fm.variable_store(store):
        y0 = fm.get_model_reg_multi_scale(x, adjs, 1.0, fuse=False)
    with torch.no_grad(), fm.variable_store(store):
        y1 = fm.get_model_reg_multi_scale(x, adjs, 1.0, fuse=True)
    assert (y0 - y1).abs().max().item() < 1e-5


@pytest.mark.parametrize("gname,multi", [("net_icosphere3", False), ("net_ms_icosphere2", True)])
def test_saver_checkpoint_names_follow_the_creation_order_and_round_trip(tmp_path, gname, multi):
    """The variables the network creates (order, leaf names, shapes) are the ones `checkpoint.network_variables`
    names after the reference's Saver; the golden parameters written as a Saver file and loaded
    back by name give the golden outputs."""
    from facet_graph_convolution_b200 import checkpoint as ck
    from facet_graph_convolution_b200 import model as fm
    g = golden(gname)
    adjs = [T(g["adj0"]), T(g["adj1"]), T(g["adj2"])]
    spec = ck.network_variables(in_channels=g["x"].shape[2], multi_scale=multi)
    for fuse in (False, True):
        store = fm.VariableStore(dev(), seed=1)
        with torch.no_grad(), fm.variable_store(store):
            fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0, multiScale=multi, fuse=fuse)
        assert [tuple(p.shape) for p in store.params] == [s for _, s in spec]
        assert store.names == [n.rsplit("/", 1)[1].split("_")[0] for n, _ in spec]
    prefix = ck.save_network(str(tmp_path / "net"), _params(g), in_channels=g["x"].shape[2], multi_scale=multi,
                             global_step=7)
    params = ck.load_network(ck.latest_checkpoint(str(tmp_path)), in_channels=g["x"].shape[2], multi_scale=multi)
    assert prefix.endswith("net-7")
    with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=params)):
        ys = fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0, multiScale=multi)
    ys = ys if multi else (ys,)
    for y, k in zip(ys, ("y0", "y1", "y2") if multi else ("y_raw",)):
        assert np.abs(y.cpu().numpy() - g[k]).max() < 1e-5


def test_multi_scale_pipeline_from_raw_mesh():
    """Raw vertices and faces -> `coarsening.mesh_with_vertices` (seeded like the reference run) -> multi-scale
    network -> normals of the three levels -> `update_position_MS`, against what the reference's driver and
    graph functions produced for the same mesh (net_ms_icosphere2.npz)."""
    from facet_graph_convolution_b200 import coarsening
    from facet_graph_convolution_b200 import model as fm
    g = golden("net_ms_icosphere2")
    d = coarsening.mesh_with_vertices(g["V"], g["F"], g["adj0"].shape[2], rng=np.random.RandomState(1))
    adjs = [T(a.astype(np.int32)) for a in d["adjs"]]
    with torch.no_grad(), fm.variable_store(fm.VariableStore(dev(), params=_params(g))):
        ys = fm.get_model_reg_multi_scale(T(d["x"][None].astype(np.float32)), adjs, 1.0, multiScale=True)
    for y, k in zip(ys, ("y0", "y1", "y2")):
        assert np.abs(y.cpu().numpy() - g[k]).max() < 1e-5
    ns = [fm.normalizeTensor(y) for y in ys]
    for n, k in zip(ns, ("n0", "n1", "n2")):
        assert np.abs(n.cpu().numpy() - g[k]).max() < 1e-4
    xo, dxl = fm.update_position_MS(T(d["verts"][None].astype(np.float32)), ns, T(d["faces"][None].astype(np.int32)),
                                    T(d["v_faces"][None].astype(np.int32)), 2,
                                    iter_num_list=[int(i) for i in g["iters"]])
    assert len(dxl) == 3
    err = np.abs(xo.cpu().numpy() - g["verts_out"]).max()
    assert err < 1e-4, err      # north_star: vertex positions max-abs <= 1e-4
