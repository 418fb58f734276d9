"""GPU parity of the vertex-space training objective (reference Code/train.py:741-781 trainAccuracyNet, :1060-1102
trainDoubleLossNet): the backward of update_position_MS and the gradient of every network parameter through
network -> normalizeTensor -> update_position_MS [80,20,20] -> fullLoss (+ faceNormalsLoss), against autograd through
the reference code (tests/golden/ms_train_icosphere2.npz, oracle/make_golden.py ms_train_case)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
GRAD_RTOL = 1e-4          # relative to the largest entry of the reference gradient


def dev():
    return torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _rel(a, ref):
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


def test_vertex_update_ms_backward_matches_reference():
    from facet_graph_convolution_b200 import model as fm
    g = golden("ms_train_icosphere2")
    vp = T(g["verts_in"]).requires_grad_(True)
    heads = [T(g["h%d" % i]).requires_grad_(True) for i in range(3)]
    xo, dxl = fm.update_position_MS(vp, heads, T(g["faces"]), T(g["v_faces"]), 2, iter_num_list=[int(i) for i in g["iters"]])
    assert np.abs(xo.detach().cpu().numpy() - g["verts_out"]).max() < 1e-4
    loss = fm.fullLoss(xo, T(g["gt_verts"]), T(g["ind0"]), T(g["ind1"]))
    assert abs(loss.item() - float(g["points_loss"])) <= 1e-5 * float(g["points_loss"])
    loss.backward()
    assert _rel(vp.grad.cpu().numpy(), g["gv_points"]) < GRAD_RTOL
    for i, h in enumerate(heads):
        assert _rel(h.grad.cpu().numpy(), g["gh%d_points" % i]) < GRAD_RTOL, i


def test_vertex_update_ms_backward_is_reproducible():
    from facet_graph_convolution_b200 import ops
    g = golden("ms_train_icosphere2")
    x, n1 = T(g["verts_in"]).reshape(-1, 3), T(g["h1"]).reshape(-1, 3)
    gout = torch.randn_like(x)
    a = ops.vertex_update_ms_bwd(gout, x, n1, T(g["faces"]), T(g["v_faces"]), 1, 2, 20)
    for _ in range(5):
        b = ops.vertex_update_ms_bwd(gout, x, n1, T(g["faces"]), T(g["v_faces"]), 1, 2, 20)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("double_loss", [False, True])
def test_vertex_trainer_parameter_gradients(double_loss):
    from facet_graph_convolution_b200 import model as fm
    from facet_graph_convolution_b200 import train as ftrain
    g, gp = golden("ms_train_icosphere2"), golden("net_ms_icosphere2")
    n = int(g["nparams"])
    net = fm.DenoisingNet(g["x"].shape[-1], multi_scale=True, device=dev(), params=[gp["p%02d" % i] for i in range(n)])
    patch = ftrain.VertexPatch(T(g["x"]), [T(g["adj%d" % i]) for i in range(3)], T(g["verts_in"]), T(g["gt_verts"]),
                               T(g["faces"]), T(g["v_faces"]), T(g["gt_normals"]))
    loss = ftrain.vertex_loss_on_patch(net, patch, np.random.RandomState(0), augment=False, double_loss=double_loss,
                                       iters=[int(i) for i in g["iters"]], sample_ids=(T(g["ind0"]), T(g["ind1"])))
    ref_loss = float(g["points_loss"]) + (float(g["normals_loss"]) if double_loss else 0.0)
    assert abs(loss.item() - ref_loss) <= 2e-5 * ref_loss
    loss.backward()
    plist = list(net.parameters())
    assert len(plist) == n
    checked = 0
    for i, p in enumerate(plist):
        key = ("gd%02d" if double_loss else "gp%02d") % i
        if key not in g:
            continue
        assert p.grad is not None, i
        assert _rel(p.grad.cpu().numpy(), g[key]) < GRAD_RTOL, (i, _rel(p.grad.cpu().numpy(), g[key]))
        checked += 1
    assert checked == (17 if double_loss else n)


def test_train_step_vertices_runs_and_updates():
    from facet_graph_convolution_b200 import model as fm
    from facet_graph_convolution_b200 import train as ftrain
    g, gp = golden("ms_train_icosphere2"), golden("net_ms_icosphere2")
    n = int(g["nparams"])
    net = fm.DenoisingNet(g["x"].shape[-1], multi_scale=True, device=dev(), params=[gp["p%02d" % i] for i in range(n)])
    patch = ftrain.VertexPatch(T(g["x"]), [T(g["adj%d" % i]) for i in range(3)], T(g["verts_in"]), T(g["gt_verts"]),
                               T(g["faces"]), T(g["v_faces"]), T(g["gt_normals"]))
    bucket = ftrain.GradBucket(list(net.parameters()))
    opt = ftrain.Adam(bucket)
    before = [p.detach().clone() for p in net.parameters()]
    rng = np.random.RandomState(3)
    losses = [ftrain.train_step_vertices(net, [patch], bucket, opt, rng, samples=60, double_loss=True) for _ in range(3)]
    assert all(np.isfinite(l) for l in losses)
    assert any(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
