"""The oracle (oracle/closed_form.py) against the golden vectors produced by executing the
reference sources themselves (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import CONV_CASES, golden
from oracle import closed_form as cf


@pytest.mark.parametrize("name", CONV_CASES)
def test_gather_bit_exact(name):
    g = golden(name)
    xg = cf.gather_rows(g["x"], g["adj"])
    assert xg.dtype == g["xg"].dtype
    assert np.array_equal(xg.view(np.uint32), g["xg"].view(np.uint32))


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv_forward_and_assignments(name):
    g = golden(name)
    mode = "translation" if int(g["translation"]) else "feature"
    v = None if mode == "translation" else g["v"]
    q = cf.assignments(g["x"].astype(np.float64), g["adj"], g["u"].astype(np.float64),
                       None if v is None else v.astype(np.float64), g["c"].astype(np.float64), mode)
    assert np.abs(q - g["q"]).max() < 2e-6
    y = cf.conv_fwd(g["x"], g["adj"], g["W0"], g["b"], g["u"], v, g["c"], bool(g["bias_mask"]), mode)
    assert np.abs(y - g["y"]).max() < 5e-6
    # fp32 evaluation of the restatement stays within the per-layer tolerance too
    y32 = cf.conv_fwd(g["x"], g["adj"], g["W0"], g["b"], g["u"], v, g["c"], bool(g["bias_mask"]), mode,
                      dtype=np.float32)
    assert np.abs(y32 - g["y"]).max() < 1e-5


@pytest.mark.parametrize("name", CONV_CASES)
def test_conv_backward(name):
    g = golden(name)
    mode = "translation" if int(g["translation"]) else "feature"
    v = None if mode == "translation" else g["v"]
    gr = cf.conv_bwd(g["gy"], g["x"], g["adj"], g["W0"], g["b"], g["u"], v, g["c"], bool(g["bias_mask"]), mode)
    for k, ref in (("gx", g["gx"]), ("gW0", g["gW0"]), ("gb", g["gb"]), ("gu", g["gu"]), ("gc", g["gc"])):
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(gr[k] - ref).max() / scale < 2e-5, k
    if v is not None:
        assert np.abs(gr["gv"] - g["gv"]).max() / max(1.0, np.abs(g["gv"]).max()) < 2e-5


def test_conv_variants():
    g = golden("conv_variants")
    for t in (0, 1):
        tag = "posassign_t%d_" % t
        y = cf.conv_pos_for_assignment_fwd(g["x"], g["adj"], g[tag + "W0"], g[tag + "b"], g[tag + "u"],
                                           None if t else g[tag + "vn"], g[tag + "c"], True, bool(t))
        assert np.abs(y - g[tag + "y"]).max() < 5e-6
        tag = "onlypos_t%d_" % t
        y = cf.conv_only_pos_for_assignment_fwd(g["x"], g["adj"], g[tag + "W0"], g[tag + "b"], g[tag + "u"],
                                                None if t else g[tag + "v"], g[tag + "c"], bool(t))
        assert np.abs(y - g[tag + "y"]).max() < 5e-6


def test_small_ops():
    g = golden("small_ops")
    assert np.array_equal(cf.pool_max(g["x"], 2), g["pool_max2"])
    assert np.array_equal(cf.pool_max(g["x"], 1), g["pool_max1"])
    assert np.abs(cf.pool_avg_ignore_zeros(g["xz"], 2) - g["pool_aiz2"]).max() < 1e-7
    assert np.array_equal(cf.upsample(g["x"], 2), g["up2"])
    assert np.abs(cf.lrelu(g["x"], np.float32(0.1)) - g["lrelu"]).max() < 1e-7
    assert np.abs(cf.lin(g["x"], g["lin_W"], g["lin_b"]) - g["lin_y"]).max() < 1e-6
    assert np.abs(cf.normalize_tensor(g["norm_in"]) - g["norm_out"]).max() < 2e-6
    assert abs(cf.face_normals_loss(g["loss_fn"], g["loss_gt"]) - float(g["loss"])) < 1e-3


def _params(g):
    return [g["p%02d" % i] for i in range(int(g["nparams"]))]


def test_network_single_scale_and_vertex_update():
    g = golden("net_icosphere3")
    p = cf.split_net_params(_params(g), multi_scale=False)
    y = cf.net_forward(g["x"], [g["adj0"], g["adj1"], g["adj2"]], p)
    assert np.abs(y - g["y_raw"]).max() < 2e-6
    yn = cf.normalize_tensor(y)
    assert np.abs(yn - g["y_norm"]).max() < 1e-4
    out = yn[0][g["perm"]][: int(g["nreal"])]
    pred = cf.host_normalize(out)
    assert np.abs(pred - g["pred_normals"]).max() < 1e-4
    xo = cf.update_position2(g["verts_in"][0], g["pred_normals"], g["e_map"][0], g["v_e_map"][0], 60)
    assert np.abs(xo - g["verts_out"][0]).max() < 1e-5


def test_network_multi_scale_and_ms_vertex_update():
    g = golden("net_ms_icosphere2")
    p = cf.split_net_params(_params(g), multi_scale=True)
    ys = cf.net_forward(g["x"], [g["adj0"], g["adj1"], g["adj2"]], p, multi_scale=True)
    for y, k in zip(ys, ("y0", "y1", "y2")):
        assert np.abs(y - g[k]).max() < 2e-6
    fc = cf.update_faces_center(g["verts_in"][0], g["faces"][0])
    for c, k in zip(fc, ("fc0", "fc1", "fc2")):
        assert np.abs(c - g[k]).max() < 1e-6
    xo, _ = cf.update_position_ms(g["verts_in"][0], [g["n0"], g["n1"], g["n2"]], g["faces"][0], g["v_faces"][0],
                                  2, [int(i) for i in g["iters"]])
    assert np.abs(xo - g["verts_out"][0]).max() < 1e-5


def test_index_layouts():
    from facet_graph_convolution_b200 import mesh
    g = golden("index_layouts")
    for tag in ("ico2", "torus", "open"):
        F = g[tag + "_F"]
        assert np.array_equal(mesh.faces_large_adj(F, 16), g[tag + "_adj16"])
        assert np.array_equal(mesh.faces_large_adj(F, 10), g[tag + "_adj10"])
        e, v = mesh.edge_maps(F, 20)
        assert np.array_equal(e, g[tag + "_emap"]) and np.array_equal(v, g[tag + "_vemap"])
        assert np.array_equal(mesh.vertex_faces(F, 25), g[tag + "_vf"])
    # the reference's only known-answer vector (Code/lib/coarsening.py:243-244)
    assert list(g["compute_perm_out"]) == [3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11]


@pytest.mark.parametrize("name", ["conv_64_64_M8_K16_B2", "conv_6_32_M9_K23_B2", "conv_32_64_M9_K16_nomask"])
def test_ref_order_port(name):
    """The reference-order torch-CPU port (timed as the CPU baseline) against the golden vectors."""
    import torch
    from oracle import ref_order
    g = golden(name)
    t = lambda k: torch.from_numpy(g[k])
    out = ref_order.conv_fwd_bwd(t("x"), t("adj"), t("gy"), t("W0"), t("b"), t("u"), t("v"), t("c"),
                                 bool(int(g["bias_mask"])))
    for got, key in zip(out, ("y", "gx", "gW0", "gb", "gu", "gv", "gc")):
        ref = g[key]
        assert np.abs(got.numpy() - ref).max() / max(1.0, np.abs(ref).max()) < 1e-5, key


def test_closed_form_network_matches_reference_on_c1_patch():
    """The oracle's network port against the reference run on the second patch of the two-patch C1 case
    (2 480 nodes of the noisy icosphere-5; the full-size patches are compared on the GPU)."""
    g = golden("c1_icosphere5_2patch")
    rs = np.random.RandomState(1234)
    params = [rs.normal(0.0, float(std), size=tuple(int(v) for v in shp if v > 0)).astype(np.float32)
              for shp, std in zip(g["pshape"], g["pstd"])]
    adjs = [g["adj1_%d" % l].astype(np.int32) for l in range(3)]
    y = cf.net_forward(g["x1"], adjs, cf.split_net_params(params))
    yn = cf.normalize_tensor(y)
    assert np.abs(yn[0] - g["y_norm1"]).max() < 2e-5


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_point_set_losses(tag):
    """oracle restatement of accuracyLoss / fullLoss / sampledAccuracyLoss against the reference's own outputs and
    autograd gradients (tests/golden/point_losses.npz, oracle/make_golden.py point_loss_cases)."""
    g = golden("point_losses")
    p0, p1, i0, i1 = (g[tag + "_" + k] for k in ("p0", "p1", "i0", "i1"))
    for nm, args in (("acc", (p0, p1, i0, None, "accuracy")), ("full", (p0, p1, i0, i1, "full")),
                     ("samp", (p0.reshape(1, -1, 3), p1.reshape(1, -1, 3), None, None, "accuracy"))):
        loss, grad = cf.point_set_loss(*args)
        ref_l, ref_g = float(g["%s_%s_loss" % (tag, nm)]), g["%s_%s_grad" % (tag, nm)]
        assert abs(loss - ref_l) <= 2e-6 * abs(ref_l), (tag, nm, loss, ref_l)
        assert np.abs(grad.reshape(ref_g.shape) - ref_g).max() <= 2e-5 * np.abs(ref_g).max(), (tag, nm)


def test_update_position_ms_backward():
    """oracle reverse-mode of update_position_MS against autograd through the reference (ms_train_icosphere2.npz:
    gradients of fullLoss(update_position_MS(...)) with respect to the vertices and the three heads)."""
    g = golden("ms_train_icosphere2")
    heads = [g["h0"], g["h1"], g["h2"]]
    iters = [int(i) for i in g["iters"]]
    xo, _ = cf.update_position_ms(g["verts_in"], heads, g["faces"], g["v_faces"], 2, iters)
    assert np.abs(xo - g["verts_out"].reshape(-1, 3)).max() < 1e-5
    loss, gx = cf.point_set_loss(xo[None], g["gt_verts"].astype(np.float64), g["ind0"], g["ind1"], "full")
    assert abs(loss - float(g["points_loss"])) < 2e-6 * abs(loss)
    gv, gn = cf.update_position_ms_bwd(gx[0], g["verts_in"], heads, g["faces"], g["v_faces"], 2, iters)
    # the golden is fp32 autograd through 120 sweeps (measured difference to this fp64 restatement: 2.2e-5 relative)
    assert np.abs(gv - g["gv_points"].reshape(-1, 3)).max() <= 1e-4 * np.abs(g["gv_points"]).max()
    for i in range(3):
        ref = g["gh%d_points" % i].reshape(-1, 3)
        assert np.abs(gn[i] - ref).max() <= 1e-4 * np.abs(ref).max(), i
