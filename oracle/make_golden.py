"""TEST INFRASTRUCTURE ONLY.  Generates the golden fixtures under ``tests/golden/``.

Run in the build container (needs /root/reference; it cannot run on the GPU box):

    python -m oracle.make_golden

Every fixture is produced by executing the *reference source itself*, unmodified,
over ``oracle/tf_standin.py`` (torch-CPU eager; TensorFlow is not installable here):
  * ``Code/model.py``  custom_conv2d / variants / pooling / upsampling / lrelu /
    custom_lin / get_model_reg_multi_scale
  * ``Code/utils.py``  normalizeTensor, getFacesLargeAdj, getEdgeMap, getVerticesFaces
  * ``Code/train.py``  update_position2, update_position_MS, faceNormalsLoss
  * ``Code/dataClasses.py`` InferenceMesh preprocessing (real pyramids, permutations)
Gradients come from torch autograd through the same reference code.
All inputs, injected weights (RandomState draws in variable-creation order) and
outputs are stored as float32/int32 ``.npz`` so the oracle port and the CUDA path
can be checked against them anywhere.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_runner as rr  # noqa: E402
from facet_graph_convolution_b200 import mesh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t if dtype is None else t.to(dtype)


def random_adj(rs, B, N, K, p_zero_row=0.05, p_pad=0.3, dup=True):
    """1-indexed adjacency with self in column 0, random neighbours, zero padding at the tail,
    a few all-zero rows (cnt = 0) and duplicate neighbours."""
    adj = rs.randint(1, N + 1, size=(B, N, K)).astype(np.int32)
    adj[:, :, 0] = np.arange(1, N + 1)
    valid = rs.randint(1, K + 1, size=(B, N))
    mask = np.arange(K)[None, None, :] < valid[:, :, None]
    if p_pad > 0:
        adj = np.where(mask, adj, 0)
    if dup and K > 3:
        adj[:, ::7, 2] = adj[:, ::7, 1]
    zero_rows = rs.rand(B, N) < p_zero_row
    adj[zero_rows] = 0
    # a padding hole in the middle of a row must also work (count_nonzero semantics)
    if K > 4:
        adj[:, 5::11, 3] = 0
    return adj.astype(np.int32)


def conv_case(ref, name, B, N, Cin, Cout, M, K, seed, bias_mask=True, translation=False):
    rs = np.random.RandomState(seed)
    x = rs.randn(B, N, Cin).astype(np.float32)
    adj = random_adj(rs, B, N, K)
    gy = rs.randn(B, N, Cout).astype(np.float32)
    xt = T(x).requires_grad_(True)

    # weights as leaf tensors so autograd reaches them
    holder = []
    prov = rr.rng_provider(seed + 1000)

    def provider(shape, stddev, nm):
        t = T(prov(shape, stddev, nm)).requires_grad_(True)
        holder.append(t)
        return t

    ref.tf.variables.reset(provider)
    # Variable() clones the provided tensor (non-leaf but connected), gradients flow to holder
    with contextlib.redirect_stdout(io.StringIO()):
        y, _ = ref.model.custom_conv2d(xt, T(adj), Cout, M, biasMask=bias_mask,
                                        translation_invariance=translation)
    (y * T(gy)).sum().backward()
    names = ["W0", "b", "u", "c"] + ([] if translation else ["v"])
    d = dict(x=x, adj=adj, gy=gy, y=y.detach().numpy(), gx=xt.grad.numpy(),
             bias_mask=np.int32(bias_mask), translation=np.int32(translation))
    for nm, t in zip(names, holder):
        d[nm] = t.detach().numpy()
        d["g" + nm] = t.grad.numpy() if t.grad is not None else np.zeros_like(t.detach().numpy())
    # bit-exact gather fixture (reference get_patches on x itself)
    d["xg"] = ref.model.get_patches(T(x), T(adj)).numpy()
    d["q"] = (ref.model.get_weight_assigments_translation_invariance(T(x), T(adj), holder[2].detach(), holder[3].detach())
              if translation else
              ref.model.get_weight_assigments(T(x).permute(0, 2, 1), T(adj), holder[2].detach(), holder[4].detach(),
                                              holder[3].detach())).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, y.shape)


def variant_cases(ref):
    rs = np.random.RandomState(77)
    B, N, Cw, Cout, M, K = 1, 60, 5, 8, 4, 6
    x = rs.randn(B, N, Cw + 3).astype(np.float32)
    adj = random_adj(rs, B, N, K)
    d = dict(x=x, adj=adj)
    for trans in (False, True):
        (y, _), vs = rr.run(ref.model.custom_conv2d_pos_for_assignment, T(x), T(adj), Cout, M,
                            translation_invariance=trans, provider=rr.rng_provider(5 + trans))
        tag = "posassign_t%d_" % trans
        d[tag + "y"] = rr.to_np(y)
        for nm, v in zip(["W0", "b", "u", "c"] + ([] if trans else ["vn"]), vs):
            d[tag + nm] = v
        (y, _), vs = rr.run(ref.model.custom_conv2d_only_pos_for_assignment, T(x), T(adj), Cout, M,
                            translation_invariance=trans, provider=rr.rng_provider(9 + trans))
        tag = "onlypos_t%d_" % trans
        d[tag + "y"] = rr.to_np(y)
        for nm, v in zip(["W0", "b", "u", "c"] + ([] if trans else ["v"]), vs):
            d[tag + nm] = v
    np.savez_compressed(os.path.join(OUT, "conv_variants.npz"), **d)
    print("wrote conv_variants")


def small_ops(ref):
    rs = np.random.RandomState(3)
    x = rs.randn(2, 64, 5).astype(np.float32)
    xz = x.copy()
    xz[:, ::3] = 0  # all-zero rows for avg_ignore_zeros
    xz[:, 4:8] = 0
    d = dict(x=x, xz=xz)
    d["pool_max2"] = rr.to_np(ref.model.custom_binary_tree_pooling(T(x), steps=2, pooltype="max"))
    d["pool_max1"] = rr.to_np(ref.model.custom_binary_tree_pooling(T(x), steps=1, pooltype="max"))
    d["pool_aiz2"] = rr.to_np(ref.model.custom_binary_tree_pooling(T(xz), steps=2, pooltype="avg_ignore_zeros"))
    d["up2"] = rr.to_np(ref.model.custom_upsampling(T(x), steps=2))
    d["lrelu"] = rr.to_np(ref.model.lrelu(T(x), 0.1))
    y, vs = rr.run(ref.model.custom_lin, T(x), 7, provider=rr.rng_provider(4))
    d["lin_y"], d["lin_W"], d["lin_b"] = rr.to_np(y), vs[0], vs[1]
    n = rs.randn(1, 50, 3).astype(np.float32) * 0.02
    n[0, 7] = 0
    d["norm_in"] = n
    d["norm_out"] = rr.to_np(ref.utils.normalizeTensor(T(n)))
    gt = n.copy() + rs.randn(1, 50, 3).astype(np.float32) * 0.005
    gt /= np.maximum(np.linalg.norm(gt, axis=-1, keepdims=True), 1e-9)
    gt[0, 11] = 0
    gt = gt.astype(np.float32)
    fn = rr.to_np(ref.utils.normalizeTensor(T(n)))
    d["loss_fn"], d["loss_gt"] = fn, gt
    d["loss"] = np.float32(rr.to_np(ref.train.faceNormalsLoss(T(fn), T(gt))))
    # gradient of loss(normalizeTensor(n)) w.r.t. n through the reference code
    nt = T(n).requires_grad_(True)
    ref.train.faceNormalsLoss(ref.utils.normalizeTensor(nt), T(gt)).backward()
    d["loss_norm_grad"] = nt.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "small_ops.npz"), **d)
    print("wrote small_ops")


def preprocess(ref, V, F, K, multi=False, seed=0, max_patch=25000):
    """Runs the reference's own preprocessing (dataClasses.py) on arrays."""
    np.random.seed(seed)
    ref.dataClasses.K_faces = K
    im = ref.dataClasses.InferenceMesh(max_patch, 2, 3)
    with contextlib.redirect_stdout(io.StringIO()):
        if multi:
            im.fNum, im.vNum = F.shape[0], V.shape[0]
            ref.dataClasses.PreprocessedData.addMeshWithVertices(im, V, F)
        else:
            im.addMesh_TimeEfficient(V, F)
    return im


def net_and_vertex_cases(ref):
    # ---- C1-shaped single-scale pipeline on icosphere-3 (1280 faces), reference preprocessing
    V, F = mesh.icosphere(3)
    Vn = mesh.add_vertex_noise(V, F, 0.3, 0)
    K = 16
    im = preprocess(ref, Vn, F, K)
    x = im.in_list[0].astype(np.float32)
    adjs = [a.astype(np.int32) for a in im.adj_list[0]]
    perm = np.asarray(im.permutations[0]).astype(np.int32)
    nreal = int(im.num_faces[0])
    y, vs = rr.run(ref.model.get_model_reg_multi_scale, T(x), [T(a) for a in adjs], 1.0,
                   provider=rr.rng_provider(1234))
    yn = ref.utils.normalizeTensor(y)
    outN = rr.to_np(yn)[0][perm][:nreal]
    pred = ref.utils.normalize(outN)
    e_map = im.edge_map.astype(np.int32)
    v_e_map = im.v_e_map.astype(np.int32)
    verts = Vn[None].astype(np.float32)
    xo = ref.train.update_position2(T(verts), T(pred[None].astype(np.float32)), T(e_map), T(v_e_map),
                                    iter_num=60, max_edges=20)
    d = dict(V=Vn, F=F, x=x, adj0=adjs[0], adj1=adjs[1], adj2=adjs[2], perm=perm, nreal=np.int32(nreal),
             y_raw=rr.to_np(y), y_norm=rr.to_np(yn), pred_normals=pred, e_map=e_map, v_e_map=v_e_map,
             verts_in=verts, verts_out=rr.to_np(xo), nparams=np.int32(len(vs)))
    for i, v in enumerate(vs):
        d["p%02d" % i] = v
    np.savez_compressed(os.path.join(OUT, "net_icosphere3.npz"), **d)
    print("wrote net_icosphere3", x.shape, [a.shape for a in adjs])

    # gradient of the training loss through the whole reference network (small pyramid)
    rs = np.random.RandomState(21)
    N0, Kt = 96 * 16 // 16, 8
    N0 = 96
    xs = rs.randn(1, N0, 6).astype(np.float32)
    a0 = random_adj(rs, 1, N0, Kt, p_zero_row=0.0)
    a1 = random_adj(rs, 1, N0 // 4, Kt, p_zero_row=0.0)
    a2 = random_adj(rs, 1, N0 // 16, Kt, p_zero_row=0.0)
    gt = rs.randn(1, N0, 3).astype(np.float32)
    gt /= np.linalg.norm(gt, axis=-1, keepdims=True)
    gt[0, 5] = 0
    holder = []
    prov = rr.rng_provider(99)

    def provider(shape, stddev, nm):
        t = T(prov(shape, stddev, nm)).requires_grad_(True)
        holder.append(t)
        return t

    ref.tf.variables.reset(provider)
    with contextlib.redirect_stdout(io.StringIO()):
        yt = ref.model.get_model_reg_multi_scale(T(xs), [T(a0), T(a1), T(a2)], 1.0)
    loss = ref.train.faceNormalsLoss(ref.utils.normalizeTensor(yt), T(gt))
    loss.backward()
    d = dict(x=xs, adj0=a0, adj1=a1, adj2=a2, gt=gt, y=yt.detach().numpy(), loss=np.float32(loss.item()),
             nparams=np.int32(len(holder)))
    for i, t in enumerate(holder):
        d["p%02d" % i] = t.detach().numpy()
        d["g%02d" % i] = t.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "net_train_small.npz"), **d)
    print("wrote net_train_small loss", loss.item())

    # ---- multi-scale pipeline on icosphere-2 (320 faces) via addMeshWithVertices
    V, F = mesh.icosphere(2)
    Vn = mesh.add_vertex_noise(V, F, 0.3, 1)
    im = preprocess(ref, Vn, F, K, multi=True, seed=1)
    x = im.in_list[0].astype(np.float32)
    adjs = [a.astype(np.int32) for a in im.adj_list[0]]
    faces_p = np.asarray(im.faces_list[0]).astype(np.int32)
    v_faces = np.asarray(im.v_faces_list[0]).astype(np.int32)
    vpos = np.asarray(im.v_list[0]).astype(np.float32)
    ys, vs = rr.run(ref.model.get_model_reg_multi_scale, T(x), [T(a) for a in adjs], 1.0, multiScale=True,
                    provider=rr.rng_provider(4321))
    n0, n1, n2 = (ref.utils.normalizeTensor(t) for t in ys)
    xo, dxl = ref.train.update_position_MS(T(vpos), [n0, n1, n2], T(faces_p), T(v_faces), 2,
                                           iter_num_list=[8, 4, 4])
    d = dict(V=Vn, F=F, x=x, adj0=adjs[0], adj1=adjs[1], adj2=adjs[2], faces=faces_p, v_faces=v_faces,
             verts_in=vpos, y0=rr.to_np(ys[0]), y1=rr.to_np(ys[1]), y2=rr.to_np(ys[2]),
             n0=rr.to_np(n0), n1=rr.to_np(n1), n2=rr.to_np(n2), verts_out=rr.to_np(xo),
             iters=np.array([8, 4, 4], np.int32), nparams=np.int32(len(vs)))
    fc = ref.train.updateFacesCenter(T(vpos), T(faces_p), 2)
    d["fc0"], d["fc1"], d["fc2"] = (rr.to_np(t) for t in fc)
    for i, v in enumerate(vs):
        d["p%02d" % i] = v
    np.savez_compressed(os.path.join(OUT, "net_ms_icosphere2.npz"), **d)
    print("wrote net_ms_icosphere2", x.shape, faces_p.shape, v_faces.shape, vpos.shape)


def _pack_adj(a):
    a = np.asarray(a)
    return a.astype(np.uint16) if a.max() < 65536 else a.astype(np.int32)


def c1_cases(ref):
    """BASELINE config C1 exactly as SURVEY section 8(d) defines it: icosphere-5 (20 480 faces, 10 242 vertices), noise
    sigma = 0.3 x mean edge length (RandomState(0)), K = 16, reference preprocessing with np.random.seed(0), weights
    RandomState(1234) in creation order; once as a single patch (MAX_PATCH_SIZE 25 000) and once with the default
    20 000 (two patches: the overlap-sum + float64 two-pass normalise of Code/train.py:117-136), each followed by the
    60 sweeps of update_position2.  Inputs are stored (adjacency as uint16) so the GPU test needs no host pyramid."""
    V, F = mesh.icosphere(5)
    Vn = mesh.add_vertex_noise(V, F, 0.3, 0)
    K = 16
    for tag, max_patch in (("c1_icosphere5_1patch", 25000), ("c1_icosphere5_2patch", 20000)):
        im = preprocess(ref, Vn, F, K, seed=0, max_patch=max_patch)
        npatch = len(im.in_list)
        d = dict(V=Vn.astype(np.float32), F=F.astype(np.int32), npatch=np.int32(npatch), K=np.int32(K))
        num_faces = 0
        if npatch > 1:
            for pi in range(npatch):
                num_faces = max(num_faces, int(np.max(im.patch_indices[pi])) + 1)
        predicted = np.zeros([num_faces, 3]) if npatch > 1 else None
        vs = None
        for pi in range(npatch):
            x = im.in_list[pi].astype(np.float32)
            adjs = [a.astype(np.int32) for a in im.adj_list[pi]]
            perm = np.asarray(im.permutations[pi]).astype(np.int32)
            nreal = int(im.num_faces[pi])
            y, vs = rr.run(ref.model.get_model_reg_multi_scale, T(x), [T(a) for a in adjs], 1.0,
                           provider=rr.rng_provider(1234))
            yn = ref.utils.normalizeTensor(y)
            outN = rr.to_np(yn)[0][perm][:nreal]          # train.py:117-121
            if npatch == 1:
                predicted = outN
            else:
                pidx = np.asarray(im.patch_indices[pi])
                predicted[pidx] = predicted[pidx] + outN    # train.py:126
                d["pidx%d" % pi] = pidx.astype(np.int32)
            d["x%d" % pi] = x
            for l in range(3):
                d["adj%d_%d" % (pi, l)] = _pack_adj(adjs[l])
            d["perm%d" % pi] = perm
            d["nreal%d" % pi] = np.int32(nreal)
            d["y_norm%d" % pi] = rr.to_np(yn)[0].astype(np.float32)
            print(tag, "patch", pi, x.shape, [a.shape for a in adjs], "real", nreal)
        pred = ref.utils.normalize(predicted)               # train.py:136 (float64 two-pass normalise)
        e_map = im.edge_map.astype(np.int32)
        v_e_map = im.v_e_map.astype(np.int32)
        verts = Vn[None].astype(np.float32)
        xo = ref.train.update_position2(T(verts), T(pred[None].astype(np.float32)), T(e_map), T(v_e_map),
                                        iter_num=60, max_edges=20)
        # the 474 199 weights are not stored: RandomState(1234).normal(0, std, shape) in creation order reproduces
        # them (shapes / std-devs / a checksum are)
        shp = np.zeros((len(vs), 3), np.int32)
        for i, v in enumerate(vs):
            shp[i, :v.ndim] = v.shape
        std = np.array([0.01 if (v.ndim == 1 and i % 5 == 1 and i < 40) or (i >= 40 and v.ndim == 1) else 0.05
                        for i, v in enumerate(vs)], np.float64)
        chk = np.random.RandomState(1234)
        for i, v in enumerate(vs):
            assert np.array_equal(chk.normal(0.0, std[i], size=v.shape).astype(np.float32), v), i
        d.update(pred_normals=pred.astype(np.float64), verts_out=rr.to_np(xo)[0].astype(np.float32),
                 nparams=np.int32(len(vs)), pshape=shp, pstd=std,
                 psum=np.float64(sum(float(np.asarray(v, np.float64).sum()) for v in vs)))
        np.savez_compressed(os.path.join(OUT, tag + ".npz"), **d)
        print("wrote", tag, os.path.getsize(os.path.join(OUT, tag + ".npz")) // 1024, "KiB")


def c4_grad_case(ref):
    """BASELINE config C4's unit of work at its real size: ONE 8 192-node patch (2 048 / 512 nodes at the coarser
    levels, K = 16, Wang-et-al-shaped: a noisy height-field patch with its binary-tree pyramid), loss =
    faceNormalsLoss(normalizeTensor(net(x)), gt) as Code/train.py:493-517 builds it, gradient of every one of the
    474 199 parameters by torch autograd THROUGH THE REFERENCE SOURCE (Code/model.py / utils.py / train.py over the
    stand-in).  Weights RandomState(99) in creation order (not stored)."""
    from facet_graph_convolution_b200 import patches
    K = 16
    pl, _ = patches.grid_patches(64, 64, block=64, halo=0, K=K, seed=5)     # 64 x 64 quads = 8 192 facets, one patch
    p = pl[0]
    assert p.x.shape[0] == 8192
    x = p.x[None].astype(np.float32)
    adjs = [a[None].astype(np.int32) for a in p.adjs]
    rs = np.random.RandomState(17)
    gt = x[:, :, :3] + 0.2 * rs.randn(1, 8192, 3).astype(np.float32)          # "clean" normals near the noisy ones
    gt /= np.linalg.norm(gt, axis=-1, keepdims=True)
    gt[0, 100:110] = 0                                                        # rows the loss must ignore (fake nodes)
    holder = []
    prov = rr.rng_provider(99)

    def provider(shape, stddev, nm):
        t = T(prov(shape, stddev, nm)).requires_grad_(True)
        holder.append(t)
        return t

    ref.tf.variables.reset(provider)
    with contextlib.redirect_stdout(io.StringIO()):
        yt = ref.model.get_model_reg_multi_scale(T(x), [T(a) for a in adjs], 1.0)
    loss = ref.train.faceNormalsLoss(ref.utils.normalizeTensor(yt), T(gt))
    loss.backward()
    d = dict(x=x, gt=gt, loss=np.float32(loss.item()), y=yt.detach().numpy(), nparams=np.int32(len(holder)),
             pshape=np.array([list(t.shape) + [0] * (3 - t.dim()) for t in holder], np.int32),
             pstd=np.array([0.01 if (t.dim() == 1 and i % 5 == 1 and i < 40) or (i >= 40 and t.dim() == 1) else 0.05
                            for i, t in enumerate(holder)], np.float64))
    chk = np.random.RandomState(99)
    for i, t in enumerate(holder):
        assert np.array_equal(chk.normal(0.0, d["pstd"][i], size=tuple(t.shape)).astype(np.float32), t.detach().numpy()), i
        d["g%02d" % i] = t.grad.numpy()
    for l in range(3):
        d["adj%d" % l] = _pack_adj(adjs[l])
    np.savez_compressed(os.path.join(OUT, "c4_grad_8192.npz"), **d)
    print("wrote c4_grad_8192 loss", loss.item(), os.path.getsize(os.path.join(OUT, "c4_grad_8192.npz")) // 1024, "KiB")


def index_cases(ref):
    """Index layouts of the reference's host builders on small meshes (bit-exact targets)."""
    d = {}
    for tag, (V, F) in (("ico2", mesh.icosphere(2)), ("torus", mesh.grid_mesh(12, 10, True)),
                        ("open", mesh.grid_mesh(7, 5, False))):
        with contextlib.redirect_stdout(io.StringIO()):
            d[tag + "_F"] = F
            d[tag + "_adj16"] = ref.utils.getFacesLargeAdj(F, 16)
            d[tag + "_adj10"] = ref.utils.getFacesLargeAdj(F, 10)
            e, v = ref.utils.getEdgeMap(F, maxEdges=20)
            d[tag + "_emap"], d[tag + "_vemap"] = e, v
            d[tag + "_vf"] = ref.utils.getVerticesFaces(F, 25)
    # the reference's only known-answer test (Code/lib/coarsening.py:243-244)
    d["compute_perm_out"] = np.array(
        [np.asarray(p) for p in ref.coarsening.compute_perm([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]),
                                                            np.array([2, 1, 0, 1, 0])])], dtype=object)[0]
    np.savez_compressed(os.path.join(OUT, "index_layouts.npz"), **d)
    print("wrote index_layouts")


OBJ_TEXT = """# comment line
mtllib scene.mtl

v 0 0 0
v 1.5 0 0.25
v 1 1 -0.5
v 0 1 1e-3
vn 0 0 1
vt 0.5 0.5
v -1 0.5 2
v 0.5 -1 0.125 0.9 0.1 0.2
usemtl mat0
f 1 2 3
f 1/1/1 3/2/1 4/3/1
f 1//1 4//1 5//1 6//1
  f 2 6 1 5 3
f 6 5 4
"""


def obj_cases(ref):
    """The reference's OBJ reader and writer on a hand-written file (comments, materials, vn/vt records, a
    vertex with colour columns, v/vt/vn corners, a quad and a pentagon, an indented record) and on meshes
    with padding faces."""
    import tempfile
    d = {"obj_text": np.frombuffer(OBJ_TEXT.encode(), dtype=np.uint8)}
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "in.obj"), "w") as f:
            f.write(OBJ_TEXT)
        with contextlib.redirect_stdout(io.StringIO()):
            V, adj, free_ind, F, N = ref.utils.load_mesh(tmp, "in.obj", 0, False)
        assert adj == [] and free_ind == []
        d["V"], d["F"], d["N"] = V, F, N
        Vi, Fi = mesh.icosphere(1)
        rs = np.random.RandomState(3)
        Vn = (Vi + 0.05 * rs.randn(*Vi.shape)).astype(np.float32)
        Fp = np.concatenate([Fi[:5], -np.ones((2, 3), np.int32), Fi[5:], np.zeros((3, 3), np.int32), Fi[:2]]).astype(np.int32)
        ref.utils.write_mesh(Vn, Fp, os.path.join(tmp, "out.obj"))
        d["w_V"], d["w_F"] = Vn, Fp
        d["w_text"] = np.frombuffer(open(os.path.join(tmp, "out.obj"), "rb").read(), dtype=np.uint8)
        Vc = np.concatenate([Vn, rs.rand(Vn.shape[0], 3).astype(np.float32)], axis=1)
        ref.utils.write_mesh(Vc, Fi.astype(np.int32), os.path.join(tmp, "outc.obj"))
        d["wc_V"], d["wc_F"] = Vc, Fi.astype(np.int32)
        d["wc_text"] = np.frombuffer(open(os.path.join(tmp, "outc.obj"), "rb").read(), dtype=np.uint8)
        with contextlib.redirect_stdout(io.StringIO()):
            V2, _, _, F2, N2 = ref.utils.load_mesh(tmp, "out.obj", 0, False)
        d["r_V"], d["r_F"], d["r_N"] = V2, F2, N2
    np.savez_compressed(os.path.join(OUT, "obj_io.npz"), **d)
    print("wrote obj_io")


def coarsen_cases(ref):
    """The reference's weighted graph, coarsening, node orders and adjacency lists on noisy meshes whose
    positions are scaled so that the Gaussian edge weights really vary (at unit scale they all clamp to
    0.001), global NumPy generator seeded per case."""
    d = {}
    for tag, (V, F), scale, K, seed in (("ico3", mesh.icosphere(3), 0.004, 23, 4),
                                        ("open", mesh.grid_mesh(23, 17, False), 0.01, 16, 7),
                                        ("unit", mesh.grid_mesh(16, 16, True), 1.0, 16, 2)):
        V = mesh.add_vertex_noise(V, F, 0.3, seed=1)
        adj = mesh.faces_large_adj(F, K)
        feat = mesh.face_features(V, F).astype(np.float64)
        feat[:, 3:] *= scale
        with contextlib.redirect_stdout(io.StringIO()):
            coo = ref.utils.listToSparseWNormals(adj, feat[:, -3:], feat[:, :3])
            coo_in = (coo.row.copy(), coo.col.copy(), coo.data.copy())  # coarsen() adds a zero diagonal in place
            np.random.seed(seed)
            graphs, perm = ref.coarsening.coarsen(coo, 4)
            lists = [ref.utils.sparseToList(graphs[2 * l], K) for l in range(3)]
        d[tag + "_adj"], d[tag + "_feat"], d[tag + "_seed"], d[tag + "_K"] = adj, feat, np.int32(seed), np.int32(K)
        d[tag + "_coo_row"], d[tag + "_coo_col"], d[tag + "_coo_val"] = coo_in
        d[tag + "_perm"] = np.asarray(perm, dtype=np.int32)
        d[tag + "_inv_perm"] = np.asarray(ref.utils.inv_perm(perm), dtype=np.int32)
        for l, (la, sat) in enumerate(lists):
            d[tag + "_list%d" % l], d[tag + "_sat%d" % l] = la.astype(np.int32), np.bool_(sat)
        for l, G in enumerate(graphs):
            G = G.tocsr()
            G.sort_indices()
            d[tag + "_g%d_indptr" % l], d[tag + "_g%d_indices" % l], d[tag + "_g%d_data" % l] = G.indptr, G.indices, G.data
    np.savez_compressed(os.path.join(OUT, "coarsen_cases.npz"), **d)
    print("wrote coarsen_cases")


def patch_cases(ref):
    """The reference's patch growth (`getGraphPatch_wMask`) on a noisy icosphere-4 with a mask that fills up
    between calls, and its whole patch loop (`addMesh_TimeEfficient` above the size limit: seeds, growth,
    small-component rule, pyramid per patch) with the global generator seeded."""
    V, F = mesh.icosphere(4)
    V = mesh.add_vertex_noise(V, F, 0.3, seed=2)
    K = 16
    adj = mesh.faces_large_adj(F, K)
    d = {"V": V, "F": F, "K": np.int32(K)}
    mask = np.zeros(F.shape[0])
    rs = np.random.RandomState(0)
    for t in range(6):
        seed = int(rs.randint(F.shape[0]))
        nn = int(rs.choice([300, 1000, 2500]))
        mp = int(rs.choice([100, nn - 50, nn]))
        with contextlib.redirect_stdout(io.StringIO()):
            a, o, nxt = ref.utils.getGraphPatch_wMask(adj, nn, seed, mask, mp)
        d["g%d_args" % t] = np.array([nn, seed, mp, nxt], dtype=np.int64)
        d["g%d_mask" % t] = mask.astype(np.uint8)
        d["g%d_adj" % t], d["g%d_old" % t] = a.astype(np.int32), o.astype(np.int64)
        if t % 2 == 0 and mask[o].min() == 0:
            mask[o] = 1
    old_min = ref.dataClasses.MIN_PATCH_SIZE
    ref.dataClasses.MIN_PATCH_SIZE = 700
    try:
        im = preprocess(ref, V, F, K, seed=5, max_patch=1500)
    finally:
        ref.dataClasses.MIN_PATCH_SIZE = old_min
    d["drv_args"] = np.array([1500, 700, 5, len(im.in_list)], dtype=np.int64)  # patch size, min size, seed, patches
    for i in range(len(im.in_list)):
        d["drv%d_x" % i] = im.in_list[i][0].astype(np.float32)
        for l in range(3):
            d["drv%d_adj%d" % (i, l)] = im.adj_list[i][l][0].astype(np.int32)
        d["drv%d_ids" % i] = np.asarray(im.patch_indices[i], dtype=np.int64)
        d["drv%d_perm" % i] = np.asarray(im.permutations[i], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "patch_cases.npz"), **d)
    print("wrote patch_cases", len(im.in_list))


def _digest(a):
    import hashlib
    a = np.ascontiguousarray(a)
    return np.frombuffer(hashlib.sha1(a.tobytes()).digest(), dtype=np.uint8)


def patch_vertex_cases(ref):
    """`getMeshPatch` and the patch loop of `addMeshWithVertices` (mesh above the size limit) on the noisy
    icosphere-4 of patch_cases: three single patches in full, and for the eight patches of the driver run the
    shapes and SHA-1 digests of every output array (int64 / float64 / float32 as the reference leaves them)."""
    V, F = mesh.icosphere(4)
    V = mesh.add_vertex_noise(V, F, 0.3, seed=2)
    K = 16
    adj = mesh.faces_large_adj(F, K)
    d = {"K": np.int32(K)}
    for i, seed in enumerate((3, 700, 4000)):
        with contextlib.redirect_stdout(io.StringIO()):
            vO, fO, aO, vOld, fOld = ref.utils.getMeshPatch(V.astype(np.float32), F, adj, 400, seed)
        d["mp%d_seed" % i] = np.int64(seed)
        d["mp%d_v" % i], d["mp%d_f" % i], d["mp%d_adj" % i] = vO, fO.astype(np.int32), aO.astype(np.int32)
        d["mp%d_vold" % i], d["mp%d_fold" % i] = vOld.astype(np.int32), fOld.astype(np.int32)
    im = preprocess(ref, V, F, K, multi=True, seed=6, max_patch=1500)
    d["drv_args"] = np.array([1500, 6, len(im.in_list)], dtype=np.int64)
    for i in range(len(im.in_list)):
        arrs = dict(x=im.in_list[i][0], adj0=im.adj_list[i][0][0], adj1=im.adj_list[i][1][0], adj2=im.adj_list[i][2][0],
                    faces=im.faces_list[i][0], v_faces=im.v_faces_list[i][0], verts=im.v_list[i][0],
                    face_ids=im.fOldInd_list[i], vertex_ids=im.vOldInd_list[i], old_to_new=np.asarray(im.permutations[i]))
        for k, a in arrs.items():
            a = np.asarray(a)
            d["drv%d_%s_shape" % (i, k)] = np.array(a.shape, dtype=np.int64)
            d["drv%d_%s_dtype" % (i, k)] = np.frombuffer(a.dtype.str.encode(), dtype=np.uint8)
            d["drv%d_%s_sha1" % (i, k)] = _digest(a)
    np.savez_compressed(os.path.join(OUT, "patch_vertex_cases.npz"), **d)
    print("wrote patch_vertex_cases", len(im.in_list))


def point_loss_cases(ref):
    """accuracyLoss / fullLoss / sampledAccuracyLoss (Code/train.py:1332-1464) and their gradients with respect to the
    predicted points, from the reference functions (autograd through the stand-in)."""
    rs = np.random.RandomState(17)
    d = {}
    for tag, batch, n0, n1, ns, spread in (("a", 1, 700, 650, 120, 8.0), ("b", 2, 300, 340, 64, 3.0), ("c", 1, 1500, 1400, 500, 6000.0),
                                           ("d", 2, 420, 400, 100, 60000.0)):
        p0 = (rs.rand(batch, n0, 3) * spread).astype(np.float32)
        p1 = (rs.rand(batch, n1, 3) * spread).astype(np.float32)
        p1[:, : min(n0, n1) // 2] = p0[:, : min(n0, n1) // 2] + rs.randn(batch, min(n0, n1) // 2, 3).astype(np.float32) * 0.01 * spread
        i0 = rs.randint(0, n0, ns).astype(np.int32)          # with repetitions, as np.random.randint draws them
        i1 = rs.randint(0, n1, ns).astype(np.int32)
        d[tag + "_p0"], d[tag + "_p1"], d[tag + "_i0"], d[tag + "_i1"] = p0, p1, i0, i1
        for nm, fn in (("acc", lambda x: ref.train.accuracyLoss(x, T(p1), torch.from_numpy(i0.astype(np.int64)))),
                       ("full", lambda x: ref.train.fullLoss(x, T(p1), torch.from_numpy(i0.astype(np.int64)),
                                                               torch.from_numpy(i1.astype(np.int64)))),
                       ("samp", lambda x: ref.train.sampledAccuracyLoss(x, T(p1)))):
            x = T(p0).requires_grad_(True)
            loss = fn(x)
            loss.backward()
            d["%s_%s_loss" % (tag, nm)] = np.float32(rr.to_np(loss))
            d["%s_%s_grad" % (tag, nm)] = x.grad.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "point_losses.npz"), **d)
    print("wrote point_losses")


def ms_train_case(ref):
    """The objective of the vertex-space trainers on icosphere-2 (Code/train.py:760-781 trainAccuracyNet, :1080-1102
    trainDoubleLossNet): multi-scale network -> normalizeTensor of the fine head only -> update_position_MS [80,20,20]
    -> fullLoss (+ faceNormalsLoss), gradients of every parameter by autograd through the reference code; and the
    gradient of the vertex update alone with respect to its inputs."""
    from facet_graph_convolution_b200 import mesh
    K = 16
    V, F = mesh.icosphere(2)
    Vn = mesh.add_vertex_noise(V, F, 0.3, 1)
    im = preprocess(ref, Vn, F, K, multi=True, seed=1)
    x = im.in_list[0].astype(np.float32)
    adjs = [a.astype(np.int32) for a in im.adj_list[0]]
    faces_p = np.asarray(im.faces_list[0]).astype(np.int32)
    v_faces = np.asarray(im.v_faces_list[0]).astype(np.int32)
    vpos = np.asarray(im.v_list[0]).astype(np.float32)
    rs = np.random.RandomState(33)
    gtv = (vpos * 0.97 + 0.01 * rs.randn(*vpos.shape)).astype(np.float32)
    gtn = x[:, :, :3] + 0.1 * rs.randn(*x[:, :, :3].shape).astype(np.float32)
    nrm = np.linalg.norm(gtn, axis=-1, keepdims=True)
    gtn = np.where(np.abs(x[:, :, :3]).sum(-1, keepdims=True) > 0, gtn / np.maximum(nrm, 1e-9), 0.0).astype(np.float32)
    nv = vpos.reshape(-1, 3).shape[0]
    i0 = rs.randint(0, nv, 60).astype(np.int32)
    i1 = rs.randint(0, nv, 60).astype(np.int32)
    iters = [80, 20, 20]
    prov = rr.rng_provider(4321)
    holder = []

    def provider(shape, stddev, nm):
        t = T(prov(shape, stddev, nm)).requires_grad_(True)
        holder.append(t)
        return t

    ref.tf.variables.reset(provider)
    vp = T(vpos).requires_grad_(True)
    with contextlib.redirect_stdout(io.StringIO()):
        ys = ref.model.get_model_reg_multi_scale(T(x), [T(a) for a in adjs], 1.0, multiScale=True)
    n0 = ref.utils.normalizeTensor(ys[0])
    heads = [n0, ys[1], ys[2]]
    for t in heads:
        t.retain_grad()
    xo, _ = ref.train.update_position_MS(vp, heads, T(faces_p), T(v_faces), 2, iter_num_list=iters)
    L64 = lambda a: torch.from_numpy(a.astype(np.int64))
    points = ref.train.fullLoss(xo, T(gtv), L64(i0), L64(i1))
    normals = ref.train.faceNormalsLoss(n0, T(gtn))
    d = dict(x=x, adj0=adjs[0], adj1=adjs[1], adj2=adjs[2], faces=faces_p, v_faces=v_faces, verts_in=vpos, gt_verts=gtv,
             gt_normals=gtn, ind0=i0, ind1=i1, iters=np.array(iters, np.int32), nparams=np.int32(len(holder)),
             verts_out=rr.to_np(xo), points_loss=np.float32(points.item()), normals_loss=np.float32(normals.item()),
             h0=rr.to_np(heads[0]), h1=rr.to_np(heads[1]), h2=rr.to_np(heads[2]))
    points.backward(retain_graph=True)
    d["gv_points"] = vp.grad.numpy().copy()
    for i, t in enumerate(heads):
        d["gh%d_points" % i] = t.grad.numpy().copy()
    # the parameters are the draws of rng_provider(4321) in creation order: tests/golden/net_ms_icosphere2.npz holds them
    chk = np.load(os.path.join(OUT, "net_ms_icosphere2.npz"))
    for i, t in enumerate(holder):
        assert np.array_equal(chk["p%02d" % i], t.detach().numpy()), i
        d["gp%02d" % i] = t.grad.numpy().copy()
        t.grad = None
    (points + normals).backward()
    for i, t in enumerate(holder):          # the double loss: the first layer and the heads (the rest is the sum rule)
        if i < 5 or i >= len(holder) - 12:
            d["gd%02d" % i] = t.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ms_train_icosphere2.npz"), **d)
    print("wrote ms_train_icosphere2: points", points.item(), "normals", normals.item(), "params", len(holder))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = rr.load()
    torch.manual_seed(0)
    if len(sys.argv) > 1:          # python -m oracle.make_golden c1_cases [...]: only the named generators
        for nm in sys.argv[1:]:
            globals()[nm](ref)
        return
    conv_case(ref, "conv_6_32_M9_K23_B2", 2, 300, 6, 32, 9, 23, seed=1)
    conv_case(ref, "conv_64_32_M9_K23", 1, 200, 64, 32, 9, 23, seed=2)
    conv_case(ref, "conv_64_64_M8_K16_B2", 2, 256, 64, 64, 8, 16, seed=3)
    conv_case(ref, "conv_128_128_M9_K23", 1, 96, 128, 128, 9, 23, seed=4)
    conv_case(ref, "conv_32_64_M9_K16_nomask", 1, 130, 32, 64, 9, 16, seed=5, bias_mask=False)
    conv_case(ref, "conv_12_20_M5_K7_trans", 2, 90, 12, 20, 5, 7, seed=6, translation=True)
    variant_cases(ref)
    small_ops(ref)
    net_and_vertex_cases(ref)
    index_cases(ref)
    obj_cases(ref)
    coarsen_cases(ref)
    patch_cases(ref)
    patch_vertex_cases(ref)
    c1_cases(ref)
    c4_grad_case(ref)
    point_loss_cases(ref)
    ms_train_case(ref)


if __name__ == "__main__":
    main()
