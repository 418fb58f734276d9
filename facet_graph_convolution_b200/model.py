"""Host-side mirror of the reference's operator interface for the hot path.

Same function names, argument meaning and return values as reference ``Code/model.py``,
``Code/utils.py:normalizeTensor`` and the vertex updates / loss in ``Code/train.py`` so a
network definition written against the reference reads the same here; underneath every
call is the C ABI of libfacetconv_b200.so (CUDA, sm_100a).  No TensorFlow, no CPU path.

Parameters.  The reference creates TF variables *inside* each call, in the order
W0[M,Cout,Cin], b[Cout], u[M,Cin], c[M], v[M,Cin] (model.py:430-447) / W[Cin,Cout], b
(model.py:767-768), initialised N(0,0.05) / bias N(0,0.01) (model.py:17-18,31-44).  Here a
``VariableStore`` plays the role of the TF graph's variable collection: the first pass
creates (or replays a given list of) tensors in that order, later passes reuse them.
"""
from __future__ import annotations

import contextlib
import math
from typing import List, Optional, Sequence

import torch

from . import autograd as ag
from . import ops

std_dev = 0.05        # reference Code/model.py:17
std_dev_bias = 0.01   # reference Code/model.py:18

_ACTIVE_STORE: List["VariableStore"] = []


class VariableStore:
    """Creation-ordered parameter collection (the stand-in for TF's variable scope state)."""

    def __init__(self, device="cuda", params: Optional[Sequence] = None, seed: Optional[int] = None,
                 requires_grad: bool = False):
        self.device = torch.device(device)
        self.requires_grad = requires_grad
        self.params: List[torch.Tensor] = []
        self.names: List[str] = []
        self._given = None if params is None else list(params)
        self._cursor = 0
        self._frozen = False
        self._gen = None
        if seed is not None:
            self._gen = torch.Generator(device="cpu")
            self._gen.manual_seed(seed)

    def begin(self):
        """Start a new pass over the same variables (call before re-running a network function)."""
        if self.params:
            self._frozen = True
        self._cursor = 0
        return self

    def variable(self, shape, stddev, name):
        shape = tuple(int(s) for s in shape)
        if self._frozen:
            t = self.params[self._cursor]
            if tuple(t.shape) != shape:
                raise ValueError("variable %d (%s): expected shape %s, stored %s"
                                 % (self._cursor, name, shape, tuple(t.shape)))
            self._cursor += 1
            return t
        if self._given is not None:
            src = self._given[len(self.params)]
            t = torch.as_tensor(src, dtype=torch.float32).reshape(shape).to(self.device).contiguous()
        else:
            t = (torch.randn(shape, generator=self._gen, dtype=torch.float32) * stddev).to(self.device)
        if self.requires_grad:
            t = t.detach().requires_grad_(True)
        self.params.append(t)
        self.names.append(name)
        self._cursor += 1
        return t


@contextlib.contextmanager
def variable_store(store: VariableStore):
    _ACTIVE_STORE.append(store.begin())
    try:
        yield store
    finally:
        _ACTIVE_STORE.pop()


def _store_for(x) -> VariableStore:
    if _ACTIVE_STORE:
        return _ACTIVE_STORE[-1]
    raise RuntimeError("no active VariableStore: wrap the call in `with variable_store(store):` "
                       "(the reference creates TF variables inside each layer call)")


def weight_variable(shape):      # reference Code/model.py:31-34
    return _store_for(None).variable(shape, std_dev, "weight")


def bias_variable(shape):        # reference Code/model.py:36-39
    return _store_for(None).variable(shape, std_dev_bias, "bias")


def assignment_variable(shape):  # reference Code/model.py:41-44
    return _store_for(None).variable(shape, std_dev, "assignment")


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _conv(x, adj, W0, b, u, v, c, bias_mask, cw=None, ca0=0, ca=None, act=ops.ACT_NONE, alpha=0.1, rev=None):
    if _needs_grad(x, W0, b, u, v, c):
        y = ag.FacetConvFn.apply(x, adj, W0, b, u, v, c, bias_mask, cw, ca0, ca, rev)
        return ag.LReluFn.apply(y, alpha) if act == ops.ACT_LRELU else y
    plan = ag.conv_plan(adj, W0.shape[0]) if ops.planned_shape(x, W0, cw) else None
    return ops.conv_fwd(x, adj, W0, b, u, v, c, bias_mask, act, alpha, cw, ca0, ca, plan=plan)


# ----------------------------------------------------------------------------- operators
def custom_conv2d(x, adj, out_channels, M, biasMask=True, translation_invariance=False,
                  rotation_invariance=False, _act=ops.ACT_NONE, _alpha=0.1):
    """Drop-in for reference Code/model.py:427-504.  Returns (y[B,N,Cout], [W0,u,c])."""
    if rotation_invariance and not translation_invariance:
        raise NotImplementedError("rotation-invariant assignments are out of scope: never enabled by the "
                                  "reference network (model.py:841-842) and numerically broken there")
    Cin = x.shape[2]
    W0 = weight_variable([M, out_channels, Cin])
    b = bias_variable([out_channels])
    u = assignment_variable([M, Cin])
    c = assignment_variable([M])
    if translation_invariance:
        v = -u  # u.(x_n - x_j) + c  ==  u.x_n + (-u).x_j + c   (model.py:97-124)
    else:
        v = assignment_variable([M, Cin])
    y = _conv(x, adj, W0, b, u, v, c, bool(biasMask), act=_act, alpha=_alpha)
    return y, [W0, u, c]


def custom_conv2d_pos_for_assignment(x, adj, out_channels, M, biasMask=True, translation_invariance=False,
                                     rotation_invariance=False):
    """Drop-in for reference Code/model.py:610-696.  x = [features | position(3)].  Returns (y, W0)."""
    if rotation_invariance and not translation_invariance:
        raise NotImplementedError("reference raises NameError here (vn undefined, model.py:658)")
    Ca = x.shape[2]
    Cw = Ca - 3
    W0 = weight_variable([M, out_channels, Cw])
    b = bias_variable([out_channels])
    u = assignment_variable([M, Ca])
    c = assignment_variable([M])
    vn = -u[:, :Cw] if translation_invariance else assignment_variable([M, Cw])
    v = torch.cat([vn, -u[:, Cw:]], dim=1)
    y = _conv(x, adj, W0, b, u, v, c, bool(biasMask), cw=Cw, ca0=0, ca=Ca)
    return y, W0


def custom_conv2d_only_pos_for_assignment(x, adj, out_channels, M, translation_invariance=False,
                                          rotation_invariance=False):
    """Drop-in for reference Code/model.py:699-760 (logits from positions only, bias unmasked)."""
    Ca = x.shape[2]
    Cw = Ca - 3
    W0 = weight_variable([M, out_channels, Cw])
    b = bias_variable([out_channels])
    u = assignment_variable([M, 3])
    c = assignment_variable([M])
    v = -u if translation_invariance else assignment_variable([M, 3])
    y = _conv(x, adj, W0, b, u, v, c, False, cw=Cw, ca0=Cw, ca=3)
    return y, W0


def custom_lin(input, out_channels, _act=ops.ACT_NONE, _alpha=0.1):
    """Drop-in for reference Code/model.py:763-769: input @ W + b with W[Cin,Cout]."""
    Cin = input.shape[2]
    W = weight_variable([Cin, out_channels])
    b = bias_variable([out_channels])
    if _needs_grad(input, W, b):
        y = ag.LinFn.apply(input, W, b)
        return ag.LReluFn.apply(y, _alpha) if _act == ops.ACT_LRELU else y
    return ops.lin_fwd(input, W, b, _act, _alpha)


def custom_binary_tree_pooling(x, steps=1, pooltype="max"):
    """Drop-in for reference Code/model.py:779-815."""
    group = int(math.pow(2, steps))
    if pooltype == "max":
        return ag.PoolMaxFn.apply(x, group) if _needs_grad(x) else ops.pool_max(x, group)
    if pooltype == "avg_ignore_zeros":
        return ops.pool_avg_ignore_zeros(x, steps)
    raise NotImplementedError("pooltype %r (only 'max' and 'avg_ignore_zeros' are used by the reference)" % pooltype)


def custom_upsampling(x, steps=1):
    """Drop-in for reference Code/model.py:817-825."""
    group = int(math.pow(2, steps))
    return ag.UpsampleFn.apply(x, group) if _needs_grad(x) else ops.upsample(x, group)


def lrelu(x, alpha):
    """Drop-in for reference Code/model.py:828-830."""
    return ag.LReluFn.apply(x, alpha) if _needs_grad(x) else ops.lrelu(x, alpha)


def concat_channels(a, b):
    """tf.concat([a, b], axis=-1) at reference Code/model.py:909,929."""
    return ag.Concat2Fn.apply(a, b) if _needs_grad(a, b) else ops.concat2(a, b)


def normalizeTensor(x):
    """Drop-in for reference Code/utils.py:1700-1715 (one patch: batch dimension 1)."""
    if x.dim() == 3 and x.shape[0] != 1:
        raise ValueError("normalizeTensor: the global mean spans the whole tensor; pass one patch at a time")
    return ag.NormalizeFn.apply(x) if _needs_grad(x) else ops.normalize_rows(x)


def faceNormalsLoss(fn, gt_fn):
    """Drop-in for reference Code/train.py:1272-1294."""
    if _needs_grad(fn):
        return ag.FaceNormalsLossFn.apply(fn, gt_fn)
    return ops.face_normals_loss(fn, gt_fn)[0].reshape(())


def _point_set(P0, P1, ind0, ind1, mode):
    if _needs_grad(P0):
        return ag.PointSetLossFn.apply(P0, P1, ind0, ind1, mode)
    return ops.point_set_loss(P0, P1, ind0, ind1, mode)[0].reshape(())


def accuracyLoss(P0, P1, sample_ind):
    """Drop-in for reference Code/train.py:1332-1370: P0[batch, n0, 3] predicted, P1[batch, n1, 3] ground truth."""
    return _point_set(P0, P1, sample_ind, None, 0)


def fullLoss(P0, P1, sample_ind0, sample_ind1):
    """Drop-in for reference Code/train.py:1373-1424 (the loss of trainAccuracyNet / trainDoubleLossNet, :781, :1100)."""
    return _point_set(P0, P1, sample_ind0, sample_ind1, 1)


def sampledAccuracyLoss(P0, P1):
    """Drop-in for reference Code/train.py:1428-1464: both sets flattened over the batch (the reshape at :1441-1442)."""
    return _point_set(P0.reshape(1, -1, 3), P1.reshape(1, -1, 3), None, None, 0)


# ----------------------------------------------------------------------------- network
def _up_conv(h, adj, out_channels, M, steps, fused_ok):
    """custom_conv2d(custom_upsampling(h, steps), adj, out_channels, M)[0] (reference model.py:902-905, 923-926);
    variables are created in the reference's order either way."""
    if fused_ok:
        Cin = h.shape[2]
        W0 = weight_variable([M, out_channels, Cin])
        b = bias_variable([out_channels])
        u = assignment_variable([M, Cin])
        c = assignment_variable([M])
        v = assignment_variable([M, Cin])
        y = ops.conv_fwd_up(h, adj, W0, b, u, v, c, upshift=steps)
        if y is None:
            y = ops.conv_fwd(ops.upsample(h, 1 << steps), adj, W0, b, u, v, c)
        return y
    return custom_conv2d(custom_upsampling(h, steps=steps), adj, out_channels, M)[0]


def _net_variables(in_channels, multiScale=False):
    """The network's variables in the reference's creation order (model.py:853-941), created (or fetched from the
    active store) without running any layer."""
    M = 9
    out = []

    def conv(cin, cout):
        out.extend([weight_variable([M, cout, cin]), bias_variable([cout]), assignment_variable([M, cin]),
                    assignment_variable([M]), assignment_variable([M, cin])])

    def head(cin):
        out.extend([weight_variable([cin, 1024]), bias_variable([1024]), weight_variable([1024, 3]), bias_variable([3])])

    conv(in_channels, 32), conv(32, 64), conv(64, 128), conv(128, 128)
    if multiScale:
        head(128)
    conv(128, 64), conv(128, 64)
    if multiScale:
        head(64)
    conv(64, 32), conv(64, 32)
    head(32)
    return out


def _fused_forward_ok(x, adjs, multiScale):
    if multiScale or not x.is_cuda or x.dim() != 3 or x.shape[2] != 6 or len(adjs) != 3:
        return False
    B, N0, _ = x.shape
    K = adjs[0].shape[2]
    return (N0 % 16 == 0 and N0 >= 16 and K <= 32 and tuple(adjs[1].shape) == (B, N0 // 4, K)
            and tuple(adjs[2].shape) == (B, N0 // 16, K) and B * N0 >= 64)


def get_model_reg_multi_scale(x, adjs, keep_prob=1.0, coarsening_steps=2, multiScale=False, fuse=True):
    """Drop-in for reference Code/model.py:837-946: the 3-level U-Net of facet-graph convolutions.

    alpha = 0.1, coarsening_steps = 2, M = 9 and both invariance flags off are hard-coded by the
    reference (model.py:841-847,855,868,880); keep_prob is accepted and ignored there too.
    ``fuse`` lets inference fold the leaky ReLU into the producing kernel and use the fused
    regression head (never materialising the 1024-wide activation).
    """
    alpha = 0.1
    coarsening_steps = 2
    out_channels_reg = 3
    M = 9
    infer = fuse and not torch.is_grad_enabled()
    A = ops.ACT_LRELU
    if infer and _fused_forward_ok(x, adjs, multiScale):
        # the whole forward as one C-ABI call (fgc_net_fwd): pooling / up-sampling / concatenation fused into the
        # convolutions, weight images prepared once per set of parameters and kept on the variable store
        store = _store_for(x)
        params = _net_variables(6)
        prep = getattr(store, "_net_prepared", None)
        if prep is None or not prep.matches(params):
            prep = ops.NetPrepared(params)
            store._net_prepared = prep
        return ops.net_forward(x, adjs, prep)

    def conv_act(h, adj, cout):
        if infer:
            return custom_conv2d(h, adj, cout, M, _act=A, _alpha=alpha)[0]
        y, _ = custom_conv2d(h, adj, cout, M)
        return lrelu(y, alpha)

    def head(h):
        if infer:
            Cin = h.shape[2]
            W1, b1 = weight_variable([Cin, 1024]), bias_variable([1024])
            W2, b2 = weight_variable([1024, out_channels_reg]), bias_variable([out_channels_reg])
            return ops.mlp_head(h, W1, b1, W2, b2, alpha)
        h_fc = custom_lin(h, 1024, _act=A, _alpha=alpha)
        return custom_lin(h_fc, out_channels_reg)

    # Level0
    h_conv1_act = conv_act(x, adjs[0], 32)
    pool1 = custom_binary_tree_pooling(h_conv1_act, steps=coarsening_steps)
    # Level1
    h_conv2_act = conv_act(pool1, adjs[1], 64)
    pool2 = custom_binary_tree_pooling(h_conv2_act, steps=coarsening_steps)
    # Level2
    h_conv3_act = conv_act(pool2, adjs[2], 128)
    dconv3_act = conv_act(h_conv3_act, adjs[2], 128)
    y_conv2 = head(dconv3_act) if multiScale else None
    # Level1 (inference: the repeat x 4 is an index shift inside the layer when its shape has that path)
    upconv2 = _up_conv(dconv3_act, adjs[1], 64, M, coarsening_steps, infer)
    concat2 = concat_channels(upconv2, h_conv2_act)
    dconv2_act = conv_act(concat2, adjs[1], 64)
    y_conv1 = head(dconv2_act) if multiScale else None
    # Level0
    upconv1 = _up_conv(dconv2_act, adjs[0], 32, M, coarsening_steps, infer)
    concat1 = concat_channels(upconv1, h_conv1_act)
    dconv1_act = conv_act(concat1, adjs[0], 32)
    y_conv0 = head(dconv1_act)
    if multiScale:
        return y_conv0, y_conv1, y_conv2
    return y_conv0


# ----------------------------------------------------------------------------- vertex updates
def update_position2(x, face_normals, edge_map, v_edges, iter_num=20, max_edges=20):
    """Drop-in for reference Code/train.py:1467-1557.  x[1,V,3] -> [1,V,3]."""
    shape = x.shape
    out = ops.vertex_update_edges(x.reshape(-1, 3), face_normals.reshape(-1, 3), edge_map.reshape(-1, 4),
                                  v_edges.reshape(x.numel() // 3, -1), iters=iter_num, lam=1.0 / 18)
    return out.reshape(shape)


def update_position_MS(x, face_normals_list, faces, v_faces0, coarsening_steps, iter_num_list=(80, 20, 20),
                       index_lists=None):
    """Drop-in for reference Code/train.py:1668-1765.  Scales are visited coarsest first and
    iter_num_list is indexed by the loop counter (the coarsest scale gets iter_num_list[0]).
    Returns (x[1,V,3], [dx per visited scale]).  Differentiable with respect to x and the normals (the vertex-space
    trainers, train.py:771-781); `index_lists[scale]` = ops.vertex_update_ms_lists(...) lets a trainer build the
    backward's index lists once per mesh."""
    x = x.reshape(-1, 3)
    scale_num = len(face_normals_list)
    dx_list = []
    for s in range(scale_num):
        cur_scale = scale_num - 1 - s
        x_init = x
        nrm = face_normals_list[cur_scale].reshape(-1, 3)
        if _needs_grad(x) or _needs_grad(nrm):
            lists = None if index_lists is None else index_lists[cur_scale]
            x = ag.VertexUpdateMSFn.apply(x, nrm, faces.reshape(-1, 3), v_faces0.reshape(x.shape[0], -1), cur_scale,
                                          coarsening_steps, iter_num_list[s], lists)
        else:
            x = ops.vertex_update_ms(x, nrm, faces.reshape(-1, 3), v_faces0.reshape(x.shape[0], -1), cur_scale,
                                     coarsening_steps, iter_num_list[s])
        dx_list.append(x - x_init)
    return x.unsqueeze(0), dx_list


# ----------------------------------------------------------------------------- nn.Module front-ends
class FacetConv(torch.nn.Module):
    """Idiomatic module holding reference-layout parameters (W0[M,Cout,Cin], b, u, c, v).

    assign: "feature" (model.py:427-504), "translation" (:97-124), "feat+pos" (:610-696),
    "pos" (:699-760).  ``weight_layout="MIO"`` accepts BASELINE.json's W[M,Cin,Cout] at load time
    and transposes it once.
    """

    def __init__(self, in_channels, out_channels, M, bias_mask=True, assign="feature", device="cuda"):
        super().__init__()
        self.assign, self.bias_mask = assign, bias_mask
        Cw = in_channels - 3 if assign in ("feat+pos", "pos") else in_channels
        Ca = 3 if assign == "pos" else in_channels
        self.cw, self.ca0, self.ca = Cw, (Cw if assign == "pos" else 0), Ca
        P = torch.nn.Parameter
        self.W0 = P(torch.randn(M, out_channels, Cw, device=device) * std_dev)
        self.b = P(torch.randn(out_channels, device=device) * std_dev_bias)
        self.u = P(torch.randn(M, Ca, device=device) * std_dev)
        self.c = P(torch.randn(M, device=device) * std_dev)
        self.v = None
        if assign == "feature" or assign == "pos":
            self.v = P(torch.randn(M, Ca, device=device) * std_dev)
        elif assign == "feat+pos":
            self.v = P(torch.randn(M, Cw, device=device) * std_dev)

    def load_weight(self, W, weight_layout="MOI"):
        W = torch.as_tensor(W, dtype=torch.float32, device=self.W0.device)
        if weight_layout == "MIO":
            W = W.transpose(1, 2).contiguous()
        with torch.no_grad():
            self.W0.copy_(W)

    def forward(self, x, adj, rev=None):
        if self.assign == "translation":
            v = -self.u
        elif self.assign == "feat+pos":
            v = torch.cat([self.v, -self.u[:, self.cw:]], dim=1)
        else:
            v = self.v
        bm = False if self.assign == "pos" else self.bias_mask
        return _conv(x, adj, self.W0, self.b, self.u, v, self.c, bm, cw=self.cw, ca0=self.ca0, ca=self.ca, rev=rev)


class DenoisingNet(torch.nn.Module):
    """The reference network (model.py:837-946) as a module whose parameters are kept in the
    reference's variable-creation order (so a TF checkpoint could be mapped onto it).  The parameters
    exist as soon as the module does (their shapes depend only on in_channels / multi_scale / M = 9), so
    an optimizer, a GradBucket or load_state_dict set up before the first forward sees all of them."""

    def __init__(self, in_channels=6, multi_scale=False, device="cuda", params=None, seed=None):
        super().__init__()
        self.multi_scale = multi_scale
        self.in_channels = int(in_channels)
        store = VariableStore(device=device, params=params, seed=seed)
        with variable_store(store):
            _net_variables(self.in_channels, multi_scale)      # creation order W0,b,u,c,v per conv; W,b per linear
        self.plist = torch.nn.ParameterList([torch.nn.Parameter(t) for t in store.params])
        store.params = list(self.plist)
        self._store = store

    def forward(self, x, adjs):
        if x.shape[-1] != self.in_channels:
            raise ValueError("DenoisingNet built for %d input channels, got %d" % (self.in_channels, x.shape[-1]))
        with variable_store(self._store):
            return get_model_reg_multi_scale(x, adjs, 1.0, multiScale=self.multi_scale)
