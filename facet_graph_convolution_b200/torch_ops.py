"""`torch.library` registration of the path's operators: the C-ABI calls of `ops.py` as
dispatcher-visible custom ops in the `fgc::` namespace, with shape functions (fake / meta tensors, so tracing and
`torch.compile` see through them without running a kernel) and the analytic backward of the convolution wired into
autograd.  Importing this module registers the ops; `facet_graph_convolution_b200.__init__` does so lazily through
`register()`.  The ops run the CUDA library only -- on CPU tensors they raise FacetConvError like `ops.py` does.

    y = torch.ops.fgc.conv_fwd(x, adj, W0, b, u, v, c, True, 0, 0.1)          # custom_conv2d (model.py:427-504)
    y = torch.ops.fgc.conv_fwd_up(x_coarse, adj, W0, b, u, v, c, 2, True, 0, 0.1)
    gx, gW0, gb, gu, gv, gc = torch.ops.fgc.conv_bwd(gy, x, adj, W0, u, v, c, True)
    p = torch.ops.fgc.pool_max(x, 4);  r = torch.ops.fgc.upsample(x, 4);  n = torch.ops.fgc.normalize_rows(y)
    y = torch.ops.fgc.mlp_head(x, W1, b1, W2, b2, 0.1)
    loss, _ = torch.ops.fgc.point_set_loss(p0, p1, ind0, ind1, 1)            # fullLoss (train.py:1373-1424), differentiable
    x1 = torch.ops.fgc.vertex_update_ms(x0, normals, faces, v_faces, scale, 2, 20)   # one scale of update_position_MS
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops

_REGISTERED = False


def register() -> None:
    """Idempotent: defines the `fgc::` ops once per process."""
    global _REGISTERED
    if _REGISTERED:
        return
    _REGISTERED = True
    from .autograd import reverse_adjacency

    @torch.library.custom_op("fgc::conv_fwd", mutates_args=())
    def conv_fwd(x: torch.Tensor, adj: torch.Tensor, W0: torch.Tensor, b: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                 c: torch.Tensor, bias_mask: bool, act: int, alpha: float) -> torch.Tensor:
        return ops.conv_fwd(x, adj, W0, b, u, v, c, bias_mask, act, alpha)

    @conv_fwd.register_fake
    def _(x, adj, W0, b, u, v, c, bias_mask, act, alpha):
        return x.new_empty((x.shape[0], x.shape[1], W0.shape[1]))

    @torch.library.custom_op("fgc::conv_bwd", mutates_args=())
    def conv_bwd(gy: torch.Tensor, x: torch.Tensor, adj: torch.Tensor, W0: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                 c: torch.Tensor, bias_mask: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                                                            torch.Tensor, torch.Tensor]:
        g = ops.conv_bwd(gy.contiguous(), x, adj, reverse_adjacency(adj), W0, u, v, c, bias_mask)
        return tuple(g[:6])

    @conv_bwd.register_fake
    def _(gy, x, adj, W0, u, v, c, bias_mask):
        return (torch.empty_like(x), torch.empty_like(W0), gy.new_empty((W0.shape[1],)), torch.empty_like(u),
                torch.empty_like(v), torch.empty_like(c))

    def _conv_setup(ctx, inputs, output):
        x, adj, W0, b, u, v, c, bias_mask, act, alpha = inputs
        if act != ops.ACT_NONE:
            raise RuntimeError("fgc::conv_fwd: autograd is defined for act = 0 (apply fgc::lrelu separately when training)")
        ctx.save_for_backward(x, adj, W0, u, v, c)
        ctx.bias_mask = bias_mask

    def _conv_backward(ctx, gy):
        x, adj, W0, u, v, c = ctx.saved_tensors
        gx, gW0, gb, gu, gv, gc = torch.ops.fgc.conv_bwd(gy, x, adj, W0, u, v, c, ctx.bias_mask)
        return gx, None, gW0, gb, gu, gv, gc, None, None, None

    conv_fwd.register_autograd(_conv_backward, setup_context=_conv_setup)

    @torch.library.custom_op("fgc::conv_fwd_up", mutates_args=())
    def conv_fwd_up(x_coarse: torch.Tensor, adj: torch.Tensor, W0: torch.Tensor, b: torch.Tensor, u: torch.Tensor,
                    v: torch.Tensor, c: torch.Tensor, upshift: int, bias_mask: bool, act: int, alpha: float) -> torch.Tensor:
        return ops.conv_fwd_up(x_coarse, adj, W0, b, u, v, c, upshift, bias_mask, act, alpha)

    @conv_fwd_up.register_fake
    def _(x_coarse, adj, W0, b, u, v, c, upshift, bias_mask, act, alpha):
        return x_coarse.new_empty((adj.shape[0], adj.shape[1], W0.shape[1]))

    @torch.library.custom_op("fgc::pool_max", mutates_args=())
    def pool_max(x: torch.Tensor, group: int) -> torch.Tensor:
        return ops.pool_max(x, group)

    @pool_max.register_fake
    def _(x, group):
        return x.new_empty((x.shape[0], x.shape[1] // group, x.shape[2]))

    @torch.library.custom_op("fgc::pool_max_bwd", mutates_args=())
    def pool_max_bwd(gy: torch.Tensor, x: torch.Tensor, y: torch.Tensor, group: int) -> torch.Tensor:
        return ops.pool_max_bwd(gy.contiguous(), x, y, group)

    @pool_max_bwd.register_fake
    def _(gy, x, y, group):
        return torch.empty_like(x)

    def _pool_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], output)
        ctx.group = inputs[1]

    def _pool_backward(ctx, gy):
        x, y = ctx.saved_tensors
        return torch.ops.fgc.pool_max_bwd(gy, x, y, ctx.group), None

    pool_max.register_autograd(_pool_backward, setup_context=_pool_setup)

    @torch.library.custom_op("fgc::upsample", mutates_args=())
    def upsample(x: torch.Tensor, group: int) -> torch.Tensor:
        return ops.upsample(x, group)

    @upsample.register_fake
    def _(x, group):
        return x.new_empty((x.shape[0], x.shape[1] * group, x.shape[2]))

    @torch.library.custom_op("fgc::upsample_bwd", mutates_args=())
    def upsample_bwd(gy: torch.Tensor, group: int) -> torch.Tensor:
        return ops.upsample_bwd(gy.contiguous(), group)

    @upsample_bwd.register_fake
    def _(gy, group):
        return gy.new_empty((gy.shape[0], gy.shape[1] // group, gy.shape[2]))

    def _up_setup(ctx, inputs, output):
        ctx.group = inputs[1]

    def _up_backward(ctx, gy):
        return torch.ops.fgc.upsample_bwd(gy, ctx.group), None

    upsample.register_autograd(_up_backward, setup_context=_up_setup)

    @torch.library.custom_op("fgc::normalize_rows", mutates_args=())
    def normalize_rows(x: torch.Tensor) -> torch.Tensor:
        return ops.normalize_rows(x)

    @normalize_rows.register_fake
    def _(x):
        return torch.empty_like(x)

    @torch.library.custom_op("fgc::normalize_rows_bwd", mutates_args=())
    def normalize_rows_bwd(gy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        return ops.normalize_rows_bwd(gy.contiguous(), x)

    @normalize_rows_bwd.register_fake
    def _(gy, x):
        return torch.empty_like(x)

    def _norm_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0])

    def _norm_backward(ctx, gy):
        return torch.ops.fgc.normalize_rows_bwd(gy, ctx.saved_tensors[0])

    normalize_rows.register_autograd(_norm_backward, setup_context=_norm_setup)

    @torch.library.custom_op("fgc::mlp_head", mutates_args=())
    def mlp_head(x: torch.Tensor, W1: torch.Tensor, b1: torch.Tensor, W2: torch.Tensor, b2: torch.Tensor,
                 alpha: float) -> torch.Tensor:
        return ops.mlp_head(x, W1, b1, W2, b2, alpha)

    @mlp_head.register_fake
    def _(x, W1, b1, W2, b2, alpha):
        return x.new_empty(tuple(x.shape[:-1]) + (W2.shape[1],))

    @torch.library.custom_op("fgc::gather_rows", mutates_args=())
    def gather_rows(x: torch.Tensor, adj: torch.Tensor) -> torch.Tensor:
        return ops.gather_rows(x, adj)

    @gather_rows.register_fake
    def _(x, adj):
        return x.new_empty((adj.shape[0], adj.shape[1], adj.shape[2], x.shape[2]))

    # ---- vertex-space training (reference Code/train.py:1332-1424 losses, :1668-1765 vertex update)
    @torch.library.custom_op("fgc::point_set_loss", mutates_args=())
    def point_set_loss(p0: torch.Tensor, p1: torch.Tensor, ind0: torch.Tensor, ind1: torch.Tensor,
                       mode: int) -> Tuple[torch.Tensor, torch.Tensor]:
        # empty index tensors = "every row"; returns (loss[1], d loss / d p0)
        loss, g = ops.point_set_loss(p0, p1, ind0 if ind0.numel() else None, ind1 if ind1.numel() else None, mode,
                                     need_grad=True)
        return loss, g

    @point_set_loss.register_fake
    def _(p0, p1, ind0, ind1, mode):
        return p0.new_empty((1,)), torch.empty_like(p0)

    def _psl_setup(ctx, inputs, output):
        ctx.save_for_backward(output[1])

    def _psl_backward(ctx, g_loss, g_grad):
        return ctx.saved_tensors[0] * g_loss, None, None, None, None

    point_set_loss.register_autograd(_psl_backward, setup_context=_psl_setup)

    @torch.library.custom_op("fgc::vertex_update_ms", mutates_args=())
    def vertex_update_ms(x: torch.Tensor, normals: torch.Tensor, faces: torch.Tensor, v_faces: torch.Tensor, scale: int,
                         steps: int, iters: int) -> torch.Tensor:
        return ops.vertex_update_ms(x, normals, faces, v_faces, scale, steps, iters)

    @vertex_update_ms.register_fake
    def _(x, normals, faces, v_faces, scale, steps, iters):
        return torch.empty_like(x)

    @torch.library.custom_op("fgc::vertex_update_ms_bwd", mutates_args=())
    def vertex_update_ms_bwd(g: torch.Tensor, x: torch.Tensor, normals: torch.Tensor, faces: torch.Tensor,
                             v_faces: torch.Tensor, scale: int, steps: int, iters: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return ops.vertex_update_ms_bwd(g.contiguous(), x, normals, faces, v_faces, scale, steps, iters)

    @vertex_update_ms_bwd.register_fake
    def _(g, x, normals, faces, v_faces, scale, steps, iters):
        return torch.empty_like(x), torch.empty_like(normals)

    def _vu_setup(ctx, inputs, output):
        x, normals, faces, v_faces, scale, steps, iters = inputs
        ctx.save_for_backward(x, normals, faces, v_faces)
        ctx.cfg = (scale, steps, iters)

    def _vu_backward(ctx, g):
        x, normals, faces, v_faces = ctx.saved_tensors
        gx, gn = torch.ops.fgc.vertex_update_ms_bwd(g, x, normals, faces, v_faces, *ctx.cfg)
        return gx, gn, None, None, None, None, None

    vertex_update_ms.register_autograd(_vu_backward, setup_context=_vu_setup)


OP_NAMES = ("conv_fwd", "conv_bwd", "conv_fwd_up", "pool_max", "pool_max_bwd", "upsample", "upsample_bwd", "normalize_rows",
            "normalize_rows_bwd", "mlp_head", "gather_rows", "point_set_loss", "vertex_update_ms", "vertex_update_ms_bwd")
