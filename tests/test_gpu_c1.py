"""BASELINE config C1 at its real size against outputs of the reference source itself: noisy
icosphere-5 (20 480 faces), K = 16, the reference's own preprocessing, weights RandomState(1234) in creation
order.  Fixtures `c1_icosphere5_{1,2}patch.npz` were written by oracle/make_golden.py::c1_cases from
Code/model.py / Code/dataClasses.py / Code/train.py run unmodified; the 2-patch case exercises the overlap-sum +
float64 two-pass normalise of Code/train.py:117-136.  Tolerances are north_star's: normals and vertices max-abs
<= 1e-4, mean angular difference <= 0.01 degrees."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def c1_weights(g):
    """the 474 199 network weights of the fixture: RandomState(1234).normal(0, std, shape) in creation order"""
    rs = np.random.RandomState(1234)
    out = []
    for shp, std in zip(g["pshape"], g["pstd"]):
        shape = tuple(int(v) for v in shp if v > 0)
        out.append(rs.normal(0.0, float(std), size=shape).astype(np.float32))
    assert abs(sum(float(w.astype(np.float64).sum()) for w in out) - float(g["psum"])) < 1e-6
    assert sum(w.size for w in out) == 474199
    return out


def _run_patch(fm, ops, g, pi, params):
    adjs = [T(g["adj%d_%d" % (pi, l)].astype(np.int32)) for l in range(3)]
    with torch.no_grad(), fm.variable_store(fm.VariableStore("cuda:0", params=params)):
        y = fm.get_model_reg_multi_scale(T(g["x%d" % pi]), adjs, 1.0)
    yn = fm.normalizeTensor(y)
    err = np.abs(yn.cpu().numpy()[0] - g["y_norm%d" % pi]).max()
    out = ops.gather_perm(yn.reshape(-1, 3), T(g["perm%d" % pi]))[: int(g["nreal%d" % pi])].cpu().numpy()
    return out, err


@pytest.mark.parametrize("name", ["c1_icosphere5_1patch", "c1_icosphere5_2patch"])
def test_c1_inference_matches_reference(name):
    from facet_graph_convolution_b200 import model as fm
    from facet_graph_convolution_b200 import ops, patches
    g = golden(name)
    params = c1_weights(g)
    npatch = int(g["npatch"])
    assert g["F"].shape[0] == 20480 and g["V"].shape[0] == 10242
    contributions = []
    for pi in range(npatch):
        out, err = _run_patch(fm, ops, g, pi, params)
        assert err < 1e-4, (pi, err)                       # per-patch normalised network output
        ids = g["pidx%d" % pi] if npatch > 1 else np.arange(out.shape[0])
        contributions.append((ids, out))
    if npatch == 1:
        pred = cf.host_normalize(contributions[0][1])      # train.py:136
    else:
        pred = patches.merge(20480, contributions)         # train.py:126 + :136 (overlap sum, float64 normalise)
        covered = np.zeros(20480, np.int64)
        for ids, _ in contributions:
            np.add.at(covered, ids, 1)
        assert covered.min() >= 1 and covered.max() >= 2   # every facet covered, some by both patches
    assert np.abs(pred - g["pred_normals"]).max() < 1e-4
    ang = cf.angular_diff_vec(pred, g["pred_normals"])
    assert ang.mean() < 0.01 + np.degrees(np.arccos(0.999999))
    # vertex update from OUR normals (the whole C1 pipeline), 60 sweeps of update_position2
    faces = T(g["F"].astype(np.int32))
    e_map_d, v_e_d = ops.build_edge_maps(faces, max_edges=20, nv=g["V"].shape[0])
    xo = fm.update_position2(T(g["V"][None].astype(np.float32)), T(pred[None].astype(np.float32)), e_map_d[None],
                             v_e_d[None], iter_num=60, max_edges=20)
    assert np.abs(xo.cpu().numpy()[0] - g["verts_out"]).max() < 1e-4


def test_c4_training_gradient_matches_reference_at_patch_size():
    """BASELINE config C4's unit of work at its real size (one 8 192-node patch, K = 16): d loss / d parameter for
    all 474 199 parameters against torch autograd THROUGH THE REFERENCE SOURCE (c4_grad_8192.npz, written by
    oracle/make_golden.py::c4_grad_case).  Tolerance: 1e-4 of each tensor's largest gradient entry."""
    from facet_graph_convolution_b200 import model as fm
    g = golden("c4_grad_8192")
    rs = np.random.RandomState(99)
    params = [rs.normal(0.0, float(std), size=tuple(int(v) for v in shp if v > 0)).astype(np.float32)
              for shp, std in zip(g["pshape"], g["pstd"])]
    store = fm.VariableStore("cuda:0", params=params, requires_grad=True)
    adjs = [T(g["adj%d" % l].astype(np.int32)) for l in range(3)]
    with fm.variable_store(store):
        y = fm.get_model_reg_multi_scale(T(g["x"]), adjs, 1.0)
    assert np.abs(y.detach().cpu().numpy() - g["y"]).max() < 1e-5
    loss = fm.faceNormalsLoss(fm.normalizeTensor(y), T(g["gt"]))
    assert abs(loss.item() - float(g["loss"])) < 1e-3          # degrees
    loss.backward()
    worst = 0.0
    for i, t in enumerate(store.params):
        ref = g["g%02d" % i]
        scale = max(float(np.abs(ref).max()), 1e-6)
        assert t.grad is not None, i
        err = float(np.abs(t.grad.cpu().numpy() - ref).max()) / scale
        worst = max(worst, err)
        assert err < 1e-4, (i, store.names[i], err)
