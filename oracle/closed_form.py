"""TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never imported by the product
package ``facet_graph_convolution_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.

NumPy restatement (fp64 by default, any float dtype on request) of the reference's
facet-graph convolution hot path.  Each function cites the reference file:line it
follows (paths relative to /root/reference).

PARITY PINNING.  The reference ships no golden vectors for this path (its only
known-answer test is ``Code/lib/coarsening.py:243-244``, which pins compute_perm, not
the convolution).  The arithmetic lives in TensorFlow (un-pinned version, absent from
the reference tree and from this image).  This oracle is therefore pinned against
*outputs of the reference sources themselves*, executed unmodified in the build
container over ``oracle/tf_standin.py`` (torch-CPU eager): ``oracle/make_golden.py``
commits those outputs under ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks
every function below against them.  Gradients are pinned the same way (torch autograd
through the reference source).
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------- gather


def gather_rows(x, adj):
    """``concat([0-row, x])[adj]`` per batch element -- Code/model.py:380-399 (get_slices).

    x[B,N,C], adj[B,N,K] 1-indexed with 0 = padding -> [B,N,K,C].  Pure index work: the
    result must be bit-identical to the input rows.
    """
    x = np.asarray(x)
    adj = np.asarray(adj)
    B, N, C = x.shape
    xp = np.concatenate([np.zeros((B, 1, C), dtype=x.dtype), x], axis=1)
    out = np.empty(adj.shape + (C,), dtype=x.dtype)
    for b in range(B):
        out[b] = xp[b][adj[b]]
    return out


def neighbour_count(adj):
    """Number of non-zero adjacency entries per facet -- Code/model.py:436 (count_nonzero)."""
    return np.count_nonzero(np.asarray(adj), axis=2)


# ----------------------------------------------------------------------------- assignments


def _softmax_last(a):
    a = a - a.max(axis=-1, keepdims=True)
    e = np.exp(a)
    return e / e.sum(axis=-1, keepdims=True)


def assignment_logits(x, adj, u, v, c, mode="feature"):
    """Logits a[B,N,K,M] of the soft assignment.

    mode "feature"     : u_m.x_n + v_m.x_j + c_m, padding slot => x_j = 0
                         (Code/model.py:74-95, get_weight_assigments)
    mode "translation" : u_m.(x_n - x_j) + c_m, padding slot => x_j = 0
                         (Code/model.py:97-124, get_weight_assigments_translation_invariance)
    """
    xg = gather_rows(x, adj)  # [B,N,K,C]
    if mode == "feature":
        ux = np.einsum("mc,bnc->bnm", u, x)
        vx = np.einsum("mc,bnkc->bnkm", v, xg)
        return ux[:, :, None, :] + vx + c
    if mode == "translation":
        d = x[:, :, None, :] - xg
        return np.einsum("mc,bnkc->bnkm", u, d) + c
    raise ValueError(mode)


def assignments(x, adj, u, v, c, mode="feature"):
    """q[B,N,K,M] = softmax over M of the logits (Code/model.py:94, :123)."""
    return _softmax_last(assignment_logits(x, adj, u, v, c, mode))


# ----------------------------------------------------------------------------- layer forward


def conv_fwd(x, adj, W0, b, u, v, c, bias_mask=True, mode="feature", dtype=np.float64,
             return_aux=False):
    """FeaStNet facet-graph convolution -- Code/model.py:427-504 (custom_conv2d).

    x[B,N,Cin]; adj[B,N,K] int (1-indexed, 0 pad); W0[M,Cout,Cin]; b[Cout]; u,v[M,Cin]; c[M].
    y[n] = inv_cnt[n] * sum_m W0[m] @ (sum_k q[n,k,m] x_{j_k}) + [cnt>0 or not bias_mask] * b
    (mathematically equal to the reference's gather-of-Wx order).
    """
    x = np.asarray(x, dtype)
    W0, b, u, c = (np.asarray(t, dtype) for t in (W0, b, u, c))
    v = None if v is None else np.asarray(v, dtype)
    adj = np.asarray(adj)
    xg = gather_rows(x, adj)
    q = assignments(x, adj, u, v, c, mode)
    s = np.einsum("bnkm,bnkc->bnmc", q, xg)
    cnt = neighbour_count(adj)
    inv = np.where(cnt != 0, 1.0 / np.maximum(cnt, 1), 0.0).astype(dtype)
    y = np.einsum("moc,bnmc->bno", W0, s) * inv[..., None]
    if bias_mask:
        y = y + (cnt != 0)[..., None] * b
    else:
        y = y + b
    if return_aux:
        return y, dict(q=q, s=s, xg=xg, inv=inv, cnt=cnt)
    return y


def conv_pos_for_assignment_fwd(x, adj, W0, b, u, vn, c, bias_mask=True, translation=False,
                                dtype=np.float64):
    """Code/model.py:610-696 (custom_conv2d_pos_for_assignment).

    x = [features(Cw) | position(3)]; contraction uses the features only (W0[M,Cout,Cw]); the
    logits use u.x_n + v.x_j + c with v = [vn | -u_pos] (:656-658), vn free or -u_feat when
    ``translation`` (:637-640).
    """
    x = np.asarray(x, dtype)
    Cw = x.shape[2] - 3
    u = np.asarray(u, dtype)
    vn = -u[:, :Cw] if translation else np.asarray(vn, dtype)
    v = np.concatenate([vn, -u[:, Cw:]], axis=1)
    adj = np.asarray(adj)
    q = assignments(x, adj, u, v, np.asarray(c, dtype), "feature")
    xg = gather_rows(x[:, :, :Cw], adj)
    s = np.einsum("bnkm,bnkc->bnmc", q, xg)
    cnt = neighbour_count(adj)
    inv = np.where(cnt != 0, 1.0 / np.maximum(cnt, 1), 0.0).astype(dtype)
    y = np.einsum("moc,bnmc->bno", np.asarray(W0, dtype), s) * inv[..., None]
    bb = np.asarray(b, dtype)
    return y + ((cnt != 0)[..., None] * bb if bias_mask else bb)


def conv_only_pos_for_assignment_fwd(x, adj, W0, b, u, v, c, translation=False, dtype=np.float64):
    """Code/model.py:699-760 (custom_conv2d_only_pos_for_assignment).

    Logits from the trailing 3 position channels only (:708-714); bias added unmasked (:759).
    """
    x = np.asarray(x, dtype)
    Cw = x.shape[2] - 3
    adj = np.asarray(adj)
    xp = x[:, :, Cw:]
    q = assignments(xp, adj, np.asarray(u, dtype), None if translation else np.asarray(v, dtype),
                    np.asarray(c, dtype), "translation" if translation else "feature")
    xg = gather_rows(x[:, :, :Cw], adj)
    s = np.einsum("bnkm,bnkc->bnmc", q, xg)
    cnt = neighbour_count(adj)
    inv = np.where(cnt != 0, 1.0 / np.maximum(cnt, 1), 0.0).astype(dtype)
    y = np.einsum("moc,bnmc->bno", np.asarray(W0, dtype), s) * inv[..., None]
    return y + np.asarray(b, dtype)


# ----------------------------------------------------------------------------- layer backward


def conv_bwd(gy, x, adj, W0, b, u, v, c, bias_mask=True, mode="feature", dtype=np.float64):
    """Analytic gradients of conv_fwd (what TF autodiff of Code/model.py:427-504 computes).

    Returns dict(gx, gW0, gb, gu, gv, gc).  SURVEY.md App. A.3.
    """
    gy = np.asarray(gy, dtype)
    x = np.asarray(x, dtype)
    W0, u, c = (np.asarray(t, dtype) for t in (W0, u, c))
    v = None if v is None else np.asarray(v, dtype)
    adj = np.asarray(adj)
    B, N, Cin = x.shape
    y, aux = conv_fwd(x, adj, W0, b, u, v, c, bias_mask, mode, dtype, return_aux=True)
    q, s, xg, inv, cnt = aux["q"], aux["s"], aux["xg"], aux["inv"], aux["cnt"]
    gz = gy * inv[..., None]
    gb = (gy * (cnt != 0)[..., None]).sum(axis=(0, 1)) if bias_mask else gy.sum(axis=(0, 1))
    gW0 = np.einsum("bno,bnmc->moc", gz, s)
    ds = np.einsum("moc,bno->bnmc", W0, gz)
    dq = np.einsum("bnmc,bnkc->bnkm", ds, xg)
    da = q * (dq - (q * dq).sum(axis=-1, keepdims=True))
    gc = da.sum(axis=(0, 1, 2))
    # message every slot sends to the row it gathered
    msg = np.einsum("bnkm,bnmc->bnkc", q, ds)
    gx = np.zeros_like(x)
    if mode == "feature":
        A = da.sum(axis=2)  # [B,N,M]
        gu = np.einsum("bnm,bnc->mc", A, x)
        gv = np.einsum("bnkm,bnkc->mc", da, xg)
        gx += np.einsum("bnm,mc->bnc", A, u)
        msg = msg + np.einsum("bnkm,mc->bnkc", da, v)
    else:  # translation: a = u.(x_n - xg) + c
        d = x[:, :, None, :] - xg
        gu = np.einsum("bnkm,bnkc->mc", da, d)
        gv = None
        t = np.einsum("bnkm,mc->bnkc", da, u)
        gx += t.sum(axis=2)
        msg = msg - t
    for bb in range(B):
        idx = adj[bb].reshape(-1)
        valid = idx != 0
        np.add.at(gx[bb], idx[valid] - 1, msg[bb].reshape(-1, Cin)[valid])
    return dict(gx=gx, gW0=gW0, gb=gb, gu=gu, gv=gv, gc=gc)


# ----------------------------------------------------------------------------- small ops


def lrelu(x, alpha=0.1):
    """relu(x) - alpha*relu(-x) -- Code/model.py:828-830."""
    x = np.asarray(x)
    return np.maximum(x, 0) - alpha * np.maximum(-x, 0)


def lin(x, W, b):
    """x @ W + b with W[Cin,Cout] -- Code/model.py:763-769 (custom_lin)."""
    return np.asarray(x) @ np.asarray(W) + np.asarray(b)


def pool_max(x, steps=2):
    """Max over 2**steps consecutive rows -- Code/model.py:786-788."""
    x = np.asarray(x)
    B, N, C = x.shape
    g = 2 ** steps
    return x.reshape(B, N // g, g, C).max(axis=2)


def pool_avg_ignore_zeros(x, steps=2):
    """Pairwise mean that replaces an all-zero row by its sibling -- Code/model.py:792-814."""
    px = np.asarray(x)
    B, _, C = px.shape
    for _ in range(steps):
        px = px.reshape(B, -1, 2, C)
        l0, l1 = px[:, :, 0, :], px[:, :, 1, :]
        z0 = np.all(l0 == 0, axis=-1, keepdims=True)
        z1 = np.all(l1 == 0, axis=-1, keepdims=True)
        c0 = np.where(z0, l1, l0)
        c1 = np.where(z1, l0, l1)
        px = (c0 + c1) / 2
    return px


def upsample(x, steps=2):
    """Repeat every row 2**steps times -- Code/model.py:817-825."""
    return np.repeat(np.asarray(x), 2 ** steps, axis=1)


def normalize_tensor(x, dtype=np.float64):
    """Code/utils.py:1700-1715 (normalizeTensor): global mean-abs rescale, then per-row L2
    normalisation with three 1e-5 epsilons."""
    x = np.asarray(x, dtype)
    eps = np.asarray(np.float32(1e-5), dtype)  # the reference's epsilon is a float32 constant
    x = x / (np.abs(x).mean() + eps)
    nrm = np.sqrt(eps + (x * x).sum(axis=-1))
    inv = np.where(nrm > eps, 1.0 / (nrm + eps), 0.0)
    return x * inv[..., None]


def host_normalize(a):
    """Two-pass row normalisation with +1e-8, float64 -- Code/utils.py:26-35 (normalize)."""
    a = np.asarray(a, np.float64)
    for _ in range(2):
        n = np.sqrt((a * a).sum(axis=-1, keepdims=True)) + 0.00000001
        a = a * (1 / n)
    return a


def angular_diff_vec(n0, n1):
    """Per-row angle in degrees, acos(0.999999*dot) -- Code/utils.py:1217-1239."""
    d = (host_normalize(n0) * host_normalize(n1)).sum(axis=1)
    return np.arccos(0.999999 * d) * 180 / math.pi


def face_normals_loss(fn, gt, dtype=np.float64):
    """Code/train.py:1272-1294 (faceNormalsLoss): mean over real rows of acos(clamp(dot)) in
    degrees; a row is fake when sum|gt| <= 1e-3."""
    fn = np.asarray(fn, dtype)
    gt = np.asarray(gt, dtype)
    lim = np.asarray(np.float32(0.9999999), dtype)
    d = np.clip((fn * gt).sum(axis=-1), -lim, lim)
    fake = np.abs(gt).sum(axis=-1) <= 10e-4
    ang = np.where(fake, 0.0, 180 * np.arccos(d) / math.pi)
    return ang.sum() / (~fake).sum()


# ----------------------------------------------------------------------------- network

CONV_SPECS = (  # (name, Cin, Cout) in variable-creation order, M = 9 -- Code/model.py:853-932
    ("conv1", None, 32), ("conv2", 32, 64), ("conv3", 64, 128), ("dconv3", 128, 128),
    ("upconv2", 128, 64), ("dconv2", 128, 64), ("upconv1", 64, 32), ("dconv1", 64, 32),
)


def split_net_params(flat, multi_scale=False):
    """Groups the flat creation-order variable list (W0,b,u,c,v per conv; W,b per custom_lin)
    into a dict.  Order with multi_scale: conv1 conv2 conv3 dconv3 [fc2 out2] upconv2 dconv2
    [fc1 out1] upconv1 dconv1 fc0 out0 -- Code/model.py:853-941."""
    it = iter(flat)
    p = {}

    def conv(name):
        W0, b, u, c, v = (next(it) for _ in range(5))
        p[name] = dict(W0=W0, b=b, u=u, c=c, v=v)

    def linp(name):
        W, b = next(it), next(it)
        p[name] = dict(W=W, b=b)

    for n in ("conv1", "conv2", "conv3", "dconv3"):
        conv(n)
    if multi_scale:
        linp("fc2"), linp("out2")
    for n in ("upconv2", "dconv2"):
        conv(n)
    if multi_scale:
        linp("fc1"), linp("out1")
    for n in ("upconv1", "dconv1"):
        conv(n)
    linp("fc0"), linp("out0")
    rest = list(it)
    assert not rest, "unused parameters: %d" % len(rest)
    return p


def net_forward(x, adjs, params, multi_scale=False, alpha=0.1, steps=2, dtype=np.float64):
    """3-level U-Net of facet-graph convolutions -- Code/model.py:837-946
    (get_model_reg_multi_scale).  ``params`` is the dict from split_net_params."""

    def conv(name, h, adj):
        p = params[name]
        return conv_fwd(h, adj, p["W0"], p["b"], p["u"], p["v"], p["c"], True, "feature", dtype)

    def head(h, fc, out):
        h = lrelu(lin(h, np.asarray(params[fc]["W"], dtype), np.asarray(params[fc]["b"], dtype)), alpha)
        return lin(h, np.asarray(params[out]["W"], dtype), np.asarray(params[out]["b"], dtype))

    x = np.asarray(x, dtype)
    a0, a1, a2 = adjs
    h1 = lrelu(conv("conv1", x, a0), alpha)
    p1 = pool_max(h1, steps)
    h2 = lrelu(conv("conv2", p1, a1), alpha)
    p2 = pool_max(h2, steps)
    h3 = lrelu(conv("conv3", p2, a2), alpha)
    d3 = lrelu(conv("dconv3", h3, a2), alpha)
    y2 = head(d3, "fc2", "out2") if multi_scale else None
    up2 = upsample(d3, steps)
    uc2 = conv("upconv2", up2, a1)
    d2 = lrelu(conv("dconv2", np.concatenate([uc2, h2], axis=-1), a1), alpha)
    y1 = head(d2, "fc1", "out1") if multi_scale else None
    up1 = upsample(d2, steps)
    uc1 = conv("upconv1", up1, a0)
    d1 = lrelu(conv("dconv1", np.concatenate([uc1, h1], axis=-1), a0), alpha)
    y0 = head(d1, "fc0", "out0")
    return (y0, y1, y2) if multi_scale else y0


# ----------------------------------------------------------------------------- vertex updates


def update_position2(x, face_normals, edge_map, v_edges, iter_num=60, dtype=np.float64):
    """Edge-based Jacobi vertex update -- Code/train.py:1467-1557 (update_position2).

    x[V,3]; face_normals[F,3]; edge_map[E,4] = (v1,v2,f1,f2) with -1 = no face;
    v_edges[V,max_edges] edge ids with -1 padding.  Per iteration
      x_i += (1/18) * sum_{e in v_edges[i]} sum_{w in (v1,v2)} sum_{f in (f1,f2)} n_f (n_f.(x_w - x_i))
    Padded slots resolve to edge row 0 = (vertex 0, vertex 0, zero normal, zero normal).
    """
    x = np.asarray(x, dtype).reshape(-1, 3).copy()
    fn = np.concatenate([np.zeros((1, 3), dtype), np.asarray(face_normals, dtype).reshape(-1, 3)], 0)
    em = np.asarray(edge_map).reshape(-1, 4).astype(np.int64) + np.array([0, 0, 1, 1])
    em = np.concatenate([np.zeros((1, 4), np.int64), em], 0)
    ve = np.asarray(v_edges).reshape(x.shape[0], -1).astype(np.int64) + 1
    ne = em[ve]  # [V,E,4]
    nf = fn[ne[:, :, 2:]]  # [V,E,2,3]
    lmbd = np.asarray(1.0 / 18, dtype)  # Code/train.py:1469
    for _ in range(iter_num):
        d = x[ne[:, :, :2]] - x[:, None, None, :]  # [V,E,2(w),3]
        dp = np.einsum("vewc,vefc->vewf", d, nf)
        upd = np.einsum("vewf,vefc->vc", dp, nf)
        x = x + lmbd * upd
    return x


def update_faces_center(x, faces, steps=2, dtype=np.float64):
    """Face centres at the three graph levels -- Code/train.py:1768-1798 (updateFacesCenter).
    faces[N0,3] vertex ids in permuted node order, -1 rows for fake nodes."""
    xv = np.concatenate([np.zeros((1, 3), dtype), np.asarray(x, dtype).reshape(-1, 3)], 0)
    f = np.asarray(faces).reshape(-1, 3).astype(np.int64) + 1
    c0 = xv[f].mean(axis=1)[None]
    c1 = pool_avg_ignore_zeros(c0, steps)
    c2 = pool_avg_ignore_zeros(c1, steps)
    return [c0, c1, c2]


def update_position_ms(x, face_normals_list, faces, v_faces, steps=2, iter_num_list=(80, 20, 20),
                       dtype=np.float64):
    """Multi-scale vertex update -- Code/train.py:1668-1765 (update_position_MS).

    Scales are visited coarsest first; iter_num_list is indexed by the loop counter, so the
    coarsest scale gets iter_num_list[0] (:1686-1688,:1727).  Coarse face id = v_faces // 4**scale
    with *floor* division, so -1 stays padding (:1706-1715).  lambda_v = 1/#faces_v (:1679-1683).
    Returns (x[V,3], [dx per visited scale]).
    """
    x = np.asarray(x, dtype).reshape(-1, 3).copy()
    vf0 = np.asarray(v_faces).reshape(x.shape[0], -1).astype(np.int64)
    numf = (vf0 != -1).sum(axis=-1).astype(dtype)
    with np.errstate(divide="ignore"):
        lmbd = (1.0 / numf)[:, None]
    nscale = len(face_normals_list)
    dx_list = []
    for s in range(nscale):
        cur = nscale - 1 - s
        fn = np.concatenate([np.zeros((1, 3), dtype),
                             np.asarray(face_normals_list[cur], dtype).reshape(-1, 3)], 0)
        vf = np.floor_divide(vf0, (2 ** steps) ** cur) + 1
        vfn = fn[vf]  # [V,Kf,3]
        x_init = x
        for _ in range(iter_num_list[s]):
            cpos = update_faces_center(x, faces, steps, dtype)[cur].reshape(-1, 3)
            cpos = np.concatenate([np.zeros((1, 3), dtype), cpos], 0)
            e = cpos[vf] - x[:, None, :]
            w = (vfn * e).sum(axis=-1)
            x = x + lmbd * (w[..., None] * vfn).sum(axis=1)
        dx_list.append(x - x_init)
    return x, dx_list


def _pool_avg_ignore_zeros_bwd(x, g, steps=2):
    """Gradient of pool_avg_ignore_zeros with respect to x[1,N,3] given g[1,N/2**steps,3]: the tf.where masks of
    Code/model.py:799-809 route it (a zero row's gradient goes to its sibling, which then counts twice)."""
    levels = [np.asarray(x)]
    for _ in range(steps):
        levels.append(pool_avg_ignore_zeros(levels[-1], 1))
    for lvl in range(steps - 1, -1, -1):
        px = levels[lvl].reshape(1, -1, 2, 3)
        z0 = np.all(px[:, :, 0, :] == 0, axis=-1, keepdims=True)
        z1 = np.all(px[:, :, 1, :] == 0, axis=-1, keepdims=True)
        gh = g / 2
        g0 = np.where(z0, 0.0, gh) + np.where(z1, gh, 0.0)
        g1 = np.where(z0, gh, 0.0) + np.where(z1, 0.0, gh)
        g = np.stack([g0, g1], axis=2).reshape(1, -1, 3)
    return g


def update_position_ms_bwd(g_out, x, face_normals_list, faces, v_faces, steps=2, iter_num_list=(80, 20, 20),
                           dtype=np.float64):
    """Reverse-mode derivative of update_position_ms (what TensorFlow derives for Code/train.py:1724-1758 inside the
    vertex-space trainers, :771-781): given g_out = dL/dx_out returns (dL/dx_in[V,3], [dL/dnormals per scale, in the order
    of face_normals_list]).  One sweep is x'_v = x_v + lam_v sum_k (n_k.e_k) n_k with e_k = c_k(x) - x_v."""
    x = np.asarray(x, dtype).reshape(-1, 3).copy()
    V = x.shape[0]
    vf0 = np.asarray(v_faces).reshape(V, -1).astype(np.int64)
    numf = (vf0 != -1).sum(axis=-1).astype(dtype)
    with np.errstate(divide="ignore"):
        lmbd = (1.0 / numf)[:, None]
    f1 = np.asarray(faces).reshape(-1, 3).astype(np.int64) + 1
    nscale = len(face_normals_list)
    tape = []          # (scale, vf, vfn, x before the sweep)
    for s in range(nscale):
        cur = nscale - 1 - s
        fn = np.concatenate([np.zeros((1, 3), dtype), np.asarray(face_normals_list[cur], dtype).reshape(-1, 3)], 0)
        vf = np.floor_divide(vf0, (2 ** steps) ** cur) + 1
        vfn = fn[vf]
        for _ in range(iter_num_list[s]):
            tape.append((cur, vf, vfn, x))
            cpos = np.concatenate([np.zeros((1, 3), dtype), update_faces_center(x, faces, steps, dtype)[cur].reshape(-1, 3)], 0)
            e = cpos[vf] - x[:, None, :]
            x = x + lmbd * ((vfn * e).sum(-1)[..., None] * vfn).sum(axis=1)
    g = np.asarray(g_out, dtype).reshape(-1, 3).copy()
    gn = [np.zeros((np.asarray(t).reshape(-1, 3).shape[0] + 1, 3), dtype) for t in face_normals_list]
    for cur, vf, vfn, xt in reversed(tape):
        cents = update_faces_center(xt, faces, steps, dtype)
        cpos = np.concatenate([np.zeros((1, 3), dtype), cents[cur].reshape(-1, 3)], 0)
        e = cpos[vf] - xt[:, None, :]
        h = lmbd * g
        a = (vfn * h[:, None, :]).sum(-1)                 # n_k . h
        w = (vfn * e).sum(-1)                             # n_k . e_k
        np.add.at(gn[cur], vf, a[..., None] * e + w[..., None] * h[:, None, :])
        gc = np.zeros_like(cpos)
        np.add.at(gc, vf, a[..., None] * vfn)
        gx = g - (a[..., None] * vfn).sum(axis=1)
        g0 = gc[1:][None]
        for lvl in range(cur, 0, -1):                     # down the pooling pyramid to the fine face centres
            g0 = _pool_avg_ignore_zeros_bwd(cents[lvl - 1], g0, steps)
        gv = np.zeros((V + 1, 3), dtype)
        np.add.at(gv, f1, np.repeat(g0.reshape(-1, 1, 3) / 3.0, 3, axis=1))
        g = gx + gv[1:]
    return g, [t[1:] for t in gn]


# ----------------------------------------------------------------------------- point-set losses
def _nearest(a, c):
    """min_j |a_i - c_j| and its argmin per batch element -- the two reduce_min of Code/train.py:1355-1357 / :1408-1410
    without the [batch, n, m] tensor kept.  a[B,n,3], c[B,m,3] -> (dist[B,n], arg[B,n])."""
    B, n, _ = a.shape
    dist = np.empty((B, n), a.dtype)
    arg = np.empty((B, n), np.int64)
    for b in range(B):
        for i0 in range(0, n, 512):
            df = a[b, i0:i0 + 512, None, :] - c[b][None, :, :]
            d = np.sqrt((df * df).sum(-1))
            arg[b, i0:i0 + 512] = d.argmin(1)
            dist[b, i0:i0 + 512] = d.min(1)
    return dist, arg


def point_set_loss(P0, P1, ind0=None, ind1=None, mode="full", dtype=np.float64):
    """accuracyLoss (Code/train.py:1332-1370, mode "accuracy"), fullLoss (:1373-1424, mode "full") and, with both sets
    flattened over the batch and no sample, sampledAccuracyLoss (:1428-1464).  Returns (loss, d loss / d P0); the gradient
    of a minimum goes to its argmin, |x| differentiates to x / |x| (TensorFlow's reduce_min / norm gradients)."""
    P0, P1 = np.asarray(P0, dtype), np.asarray(P1, dtype)
    B, n0, _ = P0.shape
    i0 = np.arange(n0) if ind0 is None else np.asarray(ind0, np.int64)
    i1 = np.arange(P1.shape[1]) if ind1 is None else np.asarray(ind1, np.int64)
    g = np.zeros_like(P0)
    sP0 = P0[:, i0]
    # precision: the sampled predicted points against every ground-truth point
    thr_p = 5.0 if mode == "accuracy" else 5000.0
    dp, jp = _nearest(sP0, P1)
    keep = dp <= thr_p
    prec = np.where(keep, dp, 0.0)
    wp = 1000.0 / prec.size
    for b in range(B):
        diff = sP0[b] - P1[b][jp[b]]
        np.add.at(g[b], i0, (wp * keep[b] / np.where(dp[b] > 0, dp[b], 1.0))[:, None] * diff)
    if mode == "accuracy":      # every ground-truth point against the SAMPLED predicted points, no threshold
        dc, jc = _nearest(P1, sP0)
        comp, keepc, rows = dc, np.ones_like(dc, bool), i0[jc]
        q = P1
    else:                       # the sampled ground-truth points against every predicted point
        q = P1[:, i1]
        dc, jc = _nearest(q, P0)
        keepc = dc <= 5000.0
        comp, rows = np.where(keepc, dc, 0.0), jc
    wc = 1000.0 / comp.size
    for b in range(B):
        diff = P0[b][rows[b]] - q[b]
        np.add.at(g[b], rows[b], (wc * keepc[b] / np.where(dc[b] > 0, dc[b], 1.0))[:, None] * diff)
    return 1000.0 * (prec.mean() + comp.mean()), g
