"""torch.autograd glue: each Function's forward/backward is one or two C-ABI calls.

The reverse adjacency needed by the convolution backward is cached per adjacency tensor
(keyed by storage pointer + shape + version) so it is built once per patch, not per layer.
"""
from __future__ import annotations

import torch

from . import ops

import collections

# LRU, bounded by entries AND by the bytes the cached index tensors pin on the GPU (a training set cycles through far more
# patches than fit: callers that iterate over many adjacencies should own their ReverseAdjacency / ConvPlan instead)
_REV_CACHE = collections.OrderedDict()
_REV_CACHE_MAX = 64
_CACHE_MAX_BYTES = 2 << 30


def _entry_bytes(obj) -> int:
    n = 0
    for v in vars(obj).values():
        if isinstance(v, torch.Tensor):
            n += v.numel() * v.element_size()
    return n


def _evict(cache):
    while len(cache) > _REV_CACHE_MAX or (len(cache) > 1 and sum(e[2] for e in cache.values()) > _CACHE_MAX_BYTES):
        cache.popitem(last=False)


def reverse_adjacency(adj: torch.Tensor) -> ops.ReverseAdjacency:
    key = (adj.data_ptr(), tuple(adj.shape), adj._version, str(adj.device), adj.dtype)
    hit = _REV_CACHE.get(key)
    if hit is not None:
        _REV_CACHE.move_to_end(key)
        return hit[0]
    rev = ops.ReverseAdjacency(adj)
    _REV_CACHE[key] = (rev, adj, _entry_bytes(rev))  # keep adj alive so the pointer cannot be recycled
    _evict(_REV_CACHE)
    return rev


_PLAN_CACHE = collections.OrderedDict()


def conv_plan(adj: torch.Tensor, M: int):
    """Tile plan of an adjacency for the dense-assignment forward, cached like the reverse adjacency.
    Returns None when the shape has no planned path (the plan costs nothing to ask for then)."""
    if not ops.ConvPlan.supported(adj, M):
        return None
    key = (adj.data_ptr(), tuple(adj.shape), adj._version, str(adj.device), adj.dtype, int(M))
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        _PLAN_CACHE.move_to_end(key)
        return hit[0]
    plan = ops.ConvPlan(adj, M)
    _PLAN_CACHE[key] = (plan, adj, _entry_bytes(plan))
    _evict(_PLAN_CACHE)
    return plan


def clear_reverse_cache():
    _REV_CACHE.clear()
    _PLAN_CACHE.clear()


class FacetConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj, W0, b, u, v, c, bias_mask, cw, ca0, ca, rev):
        plan = conv_plan(adj, W0.shape[0]) if ops.planned_shape(x, W0, cw) else None
        # planned layers keep the forward's logits and x image for the backward (the parameters are
        # only updated after backward, so they are still the ones the logits were computed with)
        saved = ops.ConvSaved() if plan is not None else None
        y = ops.conv_fwd(x, adj, W0, b, u, v, c, bias_mask, ops.ACT_NONE, 0.0, cw, ca0, ca, plan=plan, save=saved)
        ctx.save_for_backward(x, adj, W0, u, v, c)
        ctx.cfg = (bias_mask, cw, ca0, ca, rev)
        ctx.fwd_saved = saved
        return y

    @staticmethod
    def backward(ctx, gy):
        x, adj, W0, u, v, c = ctx.saved_tensors
        bias_mask, cw, ca0, ca, rev = ctx.cfg
        if rev is None:
            rev = reverse_adjacency(adj)
        plan = conv_plan(adj, W0.shape[0]) if ops.planned_shape(x, W0, cw) else None
        gx, gW0, gb, gu, gv, gc = ops.conv_bwd(gy, x, adj, rev, W0, u, v, c, bias_mask, cw, ca0, ca, plan=plan,
                                               saved=ctx.fwd_saved)
        ctx.fwd_saved = None
        return gx, None, gW0, gb, gu, gv, gc, None, None, None, None, None


class PoolMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        y = ops.pool_max(x, group)
        ctx.save_for_backward(x, y)
        ctx.group = group
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y = ctx.saved_tensors
        return ops.pool_max_bwd(gy, x, y, ctx.group), None


class UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return ops.upsample(x, group)

    @staticmethod
    def backward(ctx, gy):
        return ops.upsample_bwd(gy, ctx.group), None


class LReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.save_for_backward(x)
        ctx.alpha = alpha
        return ops.lrelu(x, alpha)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return ops.lrelu_bwd(gy, x, ctx.alpha), None


class Concat2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[-1]
        return ops.concat2(a, b)

    @staticmethod
    def backward(ctx, gy):
        return ops.split2(gy, ctx.ca)


class LinFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return ops.lin_fwd(x, W, b)

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gx, gW, gb = ops.lin_bwd(gy, x, W, need_gx=ctx.needs_input_grad[0])
        return gx, gW, gb


class NormalizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.normalize_rows(x)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return ops.normalize_rows_bwd(gy, x)


class FaceNormalsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fn, gt):
        loss, gfn = ops.face_normals_loss(fn, gt, need_grad=True, gscale=1.0)
        ctx.save_for_backward(gfn)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (gfn,) = ctx.saved_tensors
        return gfn * g, None



class PointSetLossFn(torch.autograd.Function):
    """accuracyLoss / fullLoss (reference Code/train.py:1332-1424): gradient with respect to the predicted points."""

    @staticmethod
    def forward(ctx, p0, p1, ind0, ind1, mode):
        loss, gp0 = ops.point_set_loss(p0, p1, ind0, ind1, mode, need_grad=True)
        ctx.save_for_backward(gp0)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (gp0,) = ctx.saved_tensors
        return gp0 * g, None, None, None, None


class VertexUpdateMSFn(torch.autograd.Function):
    """One scale of update_position_MS (reference Code/train.py:1668-1765) with the gradient TensorFlow derives for it:
    with respect to the incoming vertices and to this scale's face normals."""

    @staticmethod
    def forward(ctx, x, normals, faces, v_faces, scale, steps, iters, lists):
        ctx.save_for_backward(x, normals, faces, v_faces)
        ctx.cfg = (int(scale), int(steps), int(iters), lists)
        return ops.vertex_update_ms(x, normals, faces, v_faces, scale, steps, iters)

    @staticmethod
    def backward(ctx, g):
        x, normals, faces, v_faces = ctx.saved_tensors
        scale, steps, iters, lists = ctx.cfg
        g_in, g_n = ops.vertex_update_ms_bwd(g.contiguous(), x, normals, faces, v_faces, scale, steps, iters, lists)
        return g_in, g_n.view_as(normals), None, None, None, None, None, None
