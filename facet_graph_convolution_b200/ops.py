"""Tensor-level wrappers over the C ABI (device tensors in, device tensors out).

PyTorch is used only for device memory, streams and autograd bookkeeping; all arithmetic
runs in libfacetconv_b200.so.  Every function requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import ConvShape, check

ACT_NONE, ACT_LRELU = 0, 1


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.FacetConvError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.FacetConvError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.int32:
        # the reference's preprocessing emits int64 and relies on placeholder casting
        # (reference Code/train.py:52-56)
        t = t.to(torch.int32)
    return t.contiguous()


def _ws(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=like.device)


def conv_shape(x, adj, W0, u, cw=None, ca0=0, ca=None) -> ConvShape:
    B, N, Cin = x.shape
    if adj.dim() != 3 or adj.shape[0] != B or adj.shape[1] != N:
        raise _lib.FacetConvError("adj must be [B,N,K] matching x[B,N,Cin]; got %s vs %s"
                                  % (tuple(adj.shape), tuple(x.shape)))
    M, Cout, Cw = W0.shape
    cw = Cin if cw is None else cw
    ca = (Cin - ca0) if ca is None else ca
    if Cw != cw:
        raise _lib.FacetConvError("W0 must be [M,Cout,%d], got %s" % (cw, tuple(W0.shape)))
    if tuple(u.shape) != (M, ca):
        raise _lib.FacetConvError("u must be [M,%d], got %s" % (ca, tuple(u.shape)))
    return ConvShape(B, N, adj.shape[2], Cin, cw, ca0, ca, Cout, M)


# ----------------------------------------------------------------------------- convolution
def planned_shape(x, W0, cw=None) -> bool:
    """Layer shapes served by the dense-assignment tensor-core kernels (conv_mma.cu)."""
    M, Cout, Cw = W0.shape
    return M == 8 and Cout == 64 and Cw == 64 and x.shape[2] % 4 == 0 and (cw is None or cw == 64)


class ConvPlan:
    """Caller-owned tile plan of an adjacency (built once, reused by every dense layer that runs on
    it): distinct neighbour rows per tile of 128/M facets + per-slot local indices.  ``nbytes == 0``
    means the shape has no planned path and conv_fwd ignores the plan."""

    MAX_MEAN_ROWS = 112.0

    @staticmethod
    def supported(adj: torch.Tensor, M: int) -> bool:
        B, N, K = adj.shape
        return adj.is_cuda and int(_lib.lib().fgc_conv_plan_bytes(B, N, K, int(M))) > 0

    def __init__(self, adj: torch.Tensor, M: int):
        L = _lib.lib()
        adj = _i32(adj, "adj")
        B, N, K = adj.shape
        self.shape, self.M = (B, N, K), int(M)
        self.nbytes = int(L.fgc_conv_plan_bytes(B, N, K, int(M)))
        self.buf = None
        self.mean_rows = None
        if self.nbytes:
            buf = torch.empty(self.nbytes, dtype=torch.uint8, device=adj.device)
            with torch.cuda.device(adj.device):
                check(L.fgc_build_conv_plan(_p(adj), B, N, K, int(M), _p(buf), self.nbytes, _stream(adj)),
                      "fgc_build_conv_plan")
            # plan header: sum of distinct rows over tiles, number of tiles (one-time host read)
            tot, nt = (int(v) for v in buf[:16].view(torch.int64).tolist())
            self.mean_rows = tot / max(nt, 1)
            # the dense-assignment path pays per distinct row: with little neighbour sharing between
            # the facets of a tile (random adjacency) the per-facet gather path is the faster one
            if self.mean_rows <= self.MAX_MEAN_ROWS:
                self.buf = buf


class ConvSaved:
    """What a planned forward leaves behind for its backward (an autograd context keeps it alive):
    the forward's workspace with the assignment logits and the fp16 image of x.  ``ws is None`` when
    the forward ran on a path that has nothing to reuse."""

    def __init__(self):
        self.ws = None


def conv_fwd(x, adj, W0, b, u, v, c, bias_mask=True, act=ACT_NONE, alpha=0.1, cw=None, ca0=0, ca=None,
             plan: Optional[ConvPlan] = None, save: Optional[ConvSaved] = None):
    """y[B,N,Cout] of the facet-graph convolution (reference Code/model.py:427-504)."""
    L = _lib.lib()
    x, W0, b, u, v, c = (_f32(t, n) for t, n in ((x, "x"), (W0, "W0"), (b, "b"), (u, "u"), (v, "v"), (c, "c")))
    adj = _i32(adj, "adj")
    s = conv_shape(x, adj, W0, u, cw, ca0, ca)
    y = torch.empty((s.B, s.N, s.Cout), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        nws = L.fgc_conv_fwd_workspace(C.byref(s))
        ws = _ws(nws, x)
        if plan is not None and plan.buf is not None:
            if plan.shape != (s.B, s.N, s.K) or plan.M != s.M:
                raise _lib.FacetConvError("conv plan built for %s/M=%d, layer is %s/M=%d"
                                          % (plan.shape, plan.M, (s.B, s.N, s.K), s.M))
            check(L.fgc_conv_fwd_planned(C.byref(s), _p(x), _p(adj), _p(plan.buf), _p(W0), _p(b), _p(u), _p(v),
                                         _p(c), _p(y), int(bool(bias_mask)), int(act), float(alpha), _p(ws),
                                         ws.numel(), _stream(x)), "fgc_conv_fwd_planned")
            if save is not None:
                save.ws = ws
        else:
            if save is not None:
                save.ws = None
            check(L.fgc_conv_fwd(C.byref(s), _p(x), _p(adj), _p(W0), _p(b), _p(u), _p(v), _p(c), _p(y),
                                 int(bool(bias_mask)), int(act), float(alpha), _p(ws), ws.numel(), _stream(x)),
                  "fgc_conv_fwd")
    return y


def conv_fwd_up(x_coarse, adj, W0, b, u, v, c, upshift, bias_mask=True, act=ACT_NONE, alpha=0.1):
    """conv_fwd(custom_upsampling(x_coarse, steps), adj, ...) without materialising the repeated tensor
    (2^upshift = 4^steps rows per coarse row).  Returns None when the shape has no fused path."""
    L = _lib.lib()
    x_coarse, W0, b, u, v, c = (_f32(t, n) for t, n in ((x_coarse, "x"), (W0, "W0"), (b, "b"), (u, "u"), (v, "v"), (c, "c")))
    adj = _i32(adj, "adj")
    B, Nc, Cin = x_coarse.shape
    N = adj.shape[1]
    if adj.shape[0] != B or (Nc << upshift) != N:
        raise _lib.FacetConvError("conv_fwd_up: adj has %d rows, x_coarse %d << %d" % (N, Nc, upshift))
    M, Cout, Cw = W0.shape
    s = ConvShape(B, N, adj.shape[2], Cin, Cw, 0, u.shape[1], Cout, M)
    if not L.fgc_conv_fwd_up_supported(C.byref(s), int(upshift)):
        return None
    y = torch.empty((B, N, Cout), dtype=torch.float32, device=x_coarse.device)
    with torch.cuda.device(x_coarse.device):
        ws = _ws(L.fgc_conv_fwd_workspace(C.byref(s)), x_coarse)
        check(L.fgc_conv_fwd_up(C.byref(s), _p(x_coarse), _p(adj), _p(W0), _p(b), _p(u), _p(v), _p(c), _p(y),
                                int(bool(bias_mask)), int(act), float(alpha), int(upshift), _p(ws), ws.numel(),
                                _stream(x_coarse)), "fgc_conv_fwd_up")
    return y


class ReverseAdjacency:
    """Caller-owned reverse adjacency (built once per adjacency tensor, reused by every backward
    through a layer that uses it).  Replaces the scatter of TF's gather gradient."""

    def __init__(self, adj: torch.Tensor):
        L = _lib.lib()
        adj = _i32(adj, "adj")
        B, N, K = adj.shape
        self.shape = (B, N, K)
        self.ptr = torch.empty(B * N + 1, dtype=torch.int32, device=adj.device)
        self.edge = torch.empty(B * N * K, dtype=torch.int32, device=adj.device)
        nnz = C.c_int64(0)
        with torch.cuda.device(adj.device):
            ws = _ws(L.fgc_reverse_adj_workspace(B, N, K), adj)
            check(L.fgc_build_reverse_adj(_p(adj), B, N, K, _p(self.ptr), _p(self.edge), C.byref(nnz),
                                          _p(ws), ws.numel(), _stream(adj)), "fgc_build_reverse_adj")
        self.nnz = int(nnz.value)
        self._adj_dev = adj.device
        self._tgt = {}   # M -> (radj, Kr, ConvPlan) or None

    def target_plan(self, M: int):
        """Reversed adjacency in forward layout + its tile plan (built on first use, cached): lets the
        gx pass of conv_bwd run on the dense-assignment kernel.  None when the in-degree exceeds the
        kernel's slot limit or the shape has no planned path."""
        if M in self._tgt:
            return self._tgt[M]
        L = _lib.lib()
        B, N, K = self.shape
        res = None
        deg = int((self.ptr[1:] - self.ptr[:-1]).max().item()) if B * N > 0 else 0
        Kr = max(8, (deg + 7) // 8 * 8)
        if Kr <= 32 and L.fgc_conv_plan_bytes(B, N, Kr, int(M)) > 0:
            radj = torch.empty((B, N, Kr), dtype=torch.int32, device=self._adj_dev)
            with torch.cuda.device(self._adj_dev):
                check(L.fgc_build_reverse_padded(_p(self.ptr), _p(self.edge), B, N, K, Kr, _p(radj), _stream(radj)),
                      "fgc_build_reverse_padded")
            plan = ConvPlan(radj, M)
            if plan.buf is not None:
                res = (radj, Kr, plan)
        self._tgt[M] = res
        return res


def conv_bwd(gy, x, adj, rev: ReverseAdjacency, W0, u, v, c, bias_mask=True, cw=None, ca0=0, ca=None,
             planned: bool = True, plan: Optional["ConvPlan"] = None, saved: Optional["ConvSaved"] = None):
    """(gx, gW0, gb, gu, gv, gc) -- deterministic backward of conv_fwd.  ``saved``: the ConvSaved a
    planned conv_fwd of the same x, u, v, c filled (its logits and x image are reused)."""
    L = _lib.lib()
    gy, x, W0, u, v, c = (_f32(t, n) for t, n in ((gy, "gy"), (x, "x"), (W0, "W0"), (u, "u"), (v, "v"), (c, "c")))
    adj = _i32(adj, "adj")
    s = conv_shape(x, adj, W0, u, cw, ca0, ca)
    if rev.shape != (s.B, s.N, s.K):
        raise _lib.FacetConvError("reverse adjacency built for %s, layer is %s" % (rev.shape, (s.B, s.N, s.K)))
    dev = x.device
    gx = torch.empty_like(x)
    gW0 = torch.empty_like(W0)
    gb = torch.empty(s.Cout, dtype=torch.float32, device=dev)
    gu = torch.empty_like(u)
    gv = torch.empty_like(v)
    gc = torch.empty_like(c)
    tp = rev.target_plan(s.M) if planned else None
    fplan = plan if (planned and plan is not None and plan.buf is not None) else None
    with torch.cuda.device(dev):
        ws = _ws(L.fgc_conv_bwd_workspace(C.byref(s)), x)
        if tp is not None or fplan is not None:
            radj, Kr, rplan = tp if tp is not None else (None, 0, None)
            fws = saved.ws if (saved is not None and fplan is not None) else None
            check(L.fgc_conv_bwd_planned(C.byref(s), _p(gy), _p(x), _p(adj), _p(fplan.buf) if fplan else None,
                                         _p(rev.ptr), _p(rev.edge), _p(radj), Kr,
                                         _p(rplan.buf) if rplan is not None else None, _p(W0), _p(u), _p(v), _p(c),
                                         _p(gx), _p(gW0), _p(gb), _p(gu), _p(gv), _p(gc), int(bool(bias_mask)),
                                         _p(fws) if fws is not None else None, fws.numel() if fws is not None else 0,
                                         _p(ws), ws.numel(), _stream(x)), "fgc_conv_bwd_planned")
        else:
            check(L.fgc_conv_bwd(C.byref(s), _p(gy), _p(x), _p(adj), _p(rev.ptr), _p(rev.edge), _p(W0), _p(u),
                                 _p(v), _p(c), _p(gx), _p(gW0), _p(gb), _p(gu), _p(gv), _p(gc),
                                 int(bool(bias_mask)), _p(ws), ws.numel(), _stream(x)), "fgc_conv_bwd")
    return gx, gW0, gb, gu, gv, gc


def gather_rows(x, adj):
    """concat([0],x)[adj] -> [B,N,K,C], bit-exact (reference Code/model.py:380-399)."""
    L = _lib.lib()
    x = _f32(x, "x")
    adj = _i32(adj, "adj")
    B, N, Cc = x.shape
    K = adj.shape[2]
    out = torch.empty((B, N, K, Cc), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.fgc_gather_rows(_p(x), _p(adj), _p(out), B, N, K, Cc, _stream(x)), "fgc_gather_rows")
    return out


def assignments(x, adj, u, v, c, ca0=0, ca=None):
    """q[B,N,K,M] (reference Code/model.py:74-95)."""
    L = _lib.lib()
    x, u, v, c = (_f32(t, n) for t, n in ((x, "x"), (u, "u"), (v, "v"), (c, "c")))
    adj = _i32(adj, "adj")
    B, N, Cin = x.shape
    M = u.shape[0]
    ca = (Cin - ca0) if ca is None else ca
    s = ConvShape(B, N, adj.shape[2], Cin, Cin, ca0, ca, 1, M)
    q = torch.empty((B, N, adj.shape[2], M), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        ws = _ws(B * N * 2 * M * 4 + 1024, x)
        check(L.fgc_assignments(C.byref(s), _p(x), _p(adj), _p(u), _p(v), _p(c), _p(q), _p(ws), ws.numel(),
                                _stream(x)), "fgc_assignments")
    return q


# ----------------------------------------------------------------------------- pooling / pointwise
def pool_max(x, group=4):
    L = _lib.lib()
    x = _f32(x, "x")
    B, N, Cc = x.shape
    if N % group:
        raise _lib.FacetConvError("pool_max: N=%d not divisible by %d" % (N, group))
    y = torch.empty((B, N // group, Cc), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.fgc_pool_max(_p(x), _p(y), B * (N // group), Cc, group, _stream(x)), "fgc_pool_max")
    return y


def pool_max_bwd(gy, x, y, group=4):
    L = _lib.lib()
    gy, x, y = _f32(gy, "gy"), _f32(x, "x"), _f32(y, "y")
    gx = torch.empty_like(x)
    B, No, Cc = y.shape
    with torch.cuda.device(x.device):
        check(L.fgc_pool_max_bwd(_p(gy), _p(x), _p(y), _p(gx), B * No, Cc, group, _stream(x)), "fgc_pool_max_bwd")
    return gx


def pool_avg_ignore_zeros(x, steps=2):
    L = _lib.lib()
    x = _f32(x, "x")
    B, N, Cc = x.shape
    y = torch.empty((B, N >> steps, Cc), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.fgc_pool_avg_ignore_zeros(_p(x), _p(y), B, N, Cc, steps, _stream(x)), "fgc_pool_avg_ignore_zeros")
    return y


def upsample(x, group=4):
    L = _lib.lib()
    x = _f32(x, "x")
    B, N, Cc = x.shape
    y = torch.empty((B, N * group, Cc), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.fgc_upsample(_p(x), _p(y), B * N, Cc, group, _stream(x)), "fgc_upsample")
    return y


def upsample_bwd(gy, group=4):
    L = _lib.lib()
    gy = _f32(gy, "gy")
    B, Ng, Cc = gy.shape
    gx = torch.empty((B, Ng // group, Cc), dtype=torch.float32, device=gy.device)
    with torch.cuda.device(gy.device):
        check(L.fgc_upsample_bwd(_p(gy), _p(gx), B * (Ng // group), Cc, group, _stream(gy)), "fgc_upsample_bwd")
    return gx


def lrelu(x, alpha=0.1):
    L = _lib.lib()
    x = _f32(x, "x")
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(L.fgc_lrelu(_p(x), _p(y), x.numel(), float(alpha), _stream(x)), "fgc_lrelu")
    return y


def lrelu_bwd(gy, x_pre, alpha=0.1):
    L = _lib.lib()
    gy, x_pre = _f32(gy, "gy"), _f32(x_pre, "x_pre")
    gx = torch.empty_like(x_pre)
    with torch.cuda.device(gy.device):
        check(L.fgc_lrelu_bwd(_p(gy), _p(x_pre), _p(gx), gy.numel(), float(alpha), _stream(gy)), "fgc_lrelu_bwd")
    return gx


def concat2(a, b):
    L = _lib.lib()
    a, b = _f32(a, "a"), _f32(b, "b")
    B, N, Ca = a.shape
    Cb = b.shape[2]
    y = torch.empty((B, N, Ca + Cb), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(L.fgc_concat2(_p(a), _p(b), _p(y), B * N, Ca, Cb, _stream(a)), "fgc_concat2")
    return y


def split2(gy, Ca):
    L = _lib.lib()
    gy = _f32(gy, "gy")
    B, N, Cc = gy.shape
    ga = torch.empty((B, N, Ca), dtype=torch.float32, device=gy.device)
    gb = torch.empty((B, N, Cc - Ca), dtype=torch.float32, device=gy.device)
    with torch.cuda.device(gy.device):
        check(L.fgc_split2(_p(gy), _p(ga), _p(gb), B * N, Ca, Cc - Ca, _stream(gy)), "fgc_split2")
    return ga, gb


def gather_perm(x, idx):
    """y[r] = x[idx[r]] on a [rows, C] tensor."""
    L = _lib.lib()
    x = _f32(x, "x")
    idx = _i32(idx, "idx")
    y = torch.empty((idx.numel(), x.shape[-1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.fgc_gather_perm(_p(x), _p(idx), _p(y), idx.numel(), x.shape[-1], _stream(x)), "fgc_gather_perm")
    return y


# ----------------------------------------------------------------------------- linear layers
def lin_fwd(x, W, b, act=ACT_NONE, alpha=0.1):
    L = _lib.lib()
    x, W, b = _f32(x, "x"), _f32(W, "W"), _f32(b, "b")
    Cin, Cout = W.shape
    rows = x.numel() // Cin
    y = torch.empty(tuple(x.shape[:-1]) + (Cout,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.fgc_lin_fwd(_p(x), _p(W), _p(b), _p(y), rows, Cin, Cout, int(act), float(alpha), _stream(x)),
              "fgc_lin_fwd")
    return y


def lin_bwd(gy, x, W, need_gx=True):
    L = _lib.lib()
    gy, x, W = _f32(gy, "gy"), _f32(x, "x"), _f32(W, "W")
    Cin, Cout = W.shape
    rows = x.numel() // Cin
    gx = torch.empty_like(x) if need_gx else None
    gW = torch.empty_like(W)
    gb = torch.empty(Cout, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_lin_bwd_workspace(rows, Cin, Cout), x)
        check(L.fgc_lin_bwd(_p(gy), _p(x), _p(W), _p(gx), _p(gW), _p(gb), rows, Cin, Cout, _p(ws), ws.numel(),
                            _stream(x)), "fgc_lin_bwd")
    return gx, gW, gb


def mlp_head(x, W1, b1, W2, b2, alpha=0.1):
    """lrelu(x@W1+b1)@W2+b2 without materialising the hidden activation (inference)."""
    L = _lib.lib()
    x, W1, b1, W2, b2 = (_f32(t, n) for t, n in ((x, "x"), (W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2")))
    Cin, H = W1.shape
    Cout = W2.shape[1]
    rows = x.numel() // Cin
    y = torch.empty(tuple(x.shape[:-1]) + (Cout,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_mlp_head_workspace(rows, Cin, H, Cout), x)
        check(L.fgc_mlp_head_fwd(_p(x), _p(W1), _p(b1), _p(W2), _p(b2), _p(y), rows, Cin, H, Cout, float(alpha),
                                 _p(ws), ws.numel(), _stream(x)), "fgc_mlp_head_fwd")
    return y


# ----------------------------------------------------------------------------- normalisation / loss
def normalize_rows(x):
    """reference Code/utils.py:1700-1715 (normalizeTensor) on ONE patch x[1,N,3] / [N,3]."""
    L = _lib.lib()
    x = _f32(x, "x")
    rows = x.numel() // 3
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_normalize_workspace(rows), x)
        check(L.fgc_normalize_rows(_p(x), _p(y), rows, _p(ws), ws.numel(), _stream(x)), "fgc_normalize_rows")
    return y


def normalize_rows_segmented(x, counts):
    """normalizeTensor per element of a batch of padded patches: x[B,Nmax,3], counts[B] int32 (device) = real
    node count of every element.  Each element gets its own global mean (utils.py:1700-1715); padding rows -> 0."""
    L = _lib.lib()
    x = _f32(x, "x")
    counts = _i32(counts, "counts")
    B, Nmax = x.shape[0], x.shape[1]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = _ws(B * 128 + 512, x)
        check(L.fgc_normalize_rows_segmented(_p(x), _p(y), B, Nmax, _p(counts), _p(ws), ws.numel(), _stream(x)),
              "fgc_normalize_rows_segmented")
    return y


def normalize_rows_bwd(gy, x):
    L = _lib.lib()
    gy, x = _f32(gy, "gy"), _f32(x, "x")
    rows = x.numel() // 3
    gx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_normalize_workspace(rows), x)
        check(L.fgc_normalize_rows_bwd(_p(gy), _p(x), _p(gx), rows, _p(ws), ws.numel(), _stream(x)),
              "fgc_normalize_rows_bwd")
    return gx


def face_normals_loss(fn, gt, need_grad=False, gscale=1.0):
    """reference Code/train.py:1272-1294.  Returns (loss[1], gfn or None)."""
    L = _lib.lib()
    fn, gt = _f32(fn, "fn"), _f32(gt, "gt")
    rows = fn.numel() // 3
    loss = torch.empty(1, dtype=torch.float32, device=fn.device)
    gfn = torch.empty_like(fn) if need_grad else None
    with torch.cuda.device(fn.device):
        ws = _ws(L.fgc_normalize_workspace(rows), fn)
        check(L.fgc_face_normals_loss(_p(fn), _p(gt), _p(loss), _p(gfn), rows, float(gscale), _p(ws), ws.numel(),
                                      _stream(fn)), "fgc_face_normals_loss")
    return loss, gfn


def point_set_loss(p0, p1, ind0=None, ind1=None, mode=1, need_grad=False):
    """reference Code/train.py:1332-1370 (accuracyLoss, mode 0) / :1373-1424 (fullLoss, mode 1).
    p0[batch, n0, 3], p1[batch, n1, 3]; ind0 / ind1 int32 sample rows or None.  Returns (loss[1], d loss / d p0 or None)."""
    L = _lib.lib()
    p0, p1 = _f32(p0, "p0"), _f32(p1, "p1")
    if p0.dim() != 3 or p1.dim() != 3 or p0.shape[2] != 3 or p1.shape[2] != 3 or p0.shape[0] != p1.shape[0]:
        raise ValueError("point_set_loss: p0[batch, n0, 3] and p1[batch, n1, 3] expected")
    batch, n0, n1 = p0.shape[0], p0.shape[1], p1.shape[1]
    ind0 = None if ind0 is None else _i32(ind0, "ind0")
    ind1 = None if ind1 is None else _i32(ind1, "ind1")
    for ind, n, nm in ((ind0, n0, "ind0"), (ind1, n1, "ind1")):
        if ind is not None and ind.numel() and bool(((ind < 0) | (ind >= n)).any()):      # one host sync per index set
            raise IndexError("point_set_loss: %s outside [0, %d)" % (nm, n))     # tf.gather raises on the CPU too
    loss = torch.empty(1, dtype=torch.float32, device=p0.device)
    gp0 = torch.empty_like(p0) if need_grad else None
    with torch.cuda.device(p0.device):
        ns0, ns1 = (0 if ind0 is None else ind0.numel()), (0 if ind1 is None else ind1.numel())
        ws = _ws(L.fgc_point_set_loss_workspace(batch, n0, n1, ns0, ns1), p0)
        check(L.fgc_point_set_loss(_p(p0), _p(p1), batch, n0, n1, _p(ind0), ns0, _p(ind1), ns1, int(mode), _p(loss), _p(gp0),
                                   _p(ws), ws.numel(),
                                   _stream(p0)), "fgc_point_set_loss")
    return loss, gp0


# ----------------------------------------------------------------------------- vertex updates
def vertex_update_edges(x, normals, edge_map, v_edges, iters=60, lam=1.0 / 18):
    """reference Code/train.py:1467-1557 (update_position2).  x[V,3] -> x[V,3]."""
    L = _lib.lib()
    x, normals = _f32(x, "x"), _f32(normals, "normals")
    edge_map, v_edges = _i32(edge_map, "edge_map"), _i32(v_edges, "v_edges")
    V = x.numel() // 3
    F = normals.numel() // 3
    E = edge_map.numel() // 4
    max_edges = v_edges.numel() // V
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_vertex_update_workspace(V), x)
        check(L.fgc_vertex_update_edges(_p(x), _p(out), _p(normals), _p(edge_map), _p(v_edges), V, F, E,
                                        max_edges, int(iters), float(lam), _p(ws), ws.numel(), _stream(x)),
              "fgc_vertex_update_edges")
    return out


def vertex_update_edges_range(x_in, x_out, normals, edge_map, v_edges, v_begin, v_end, lam=1.0 / 18):
    """ONE sweep of update_position2 (reference Code/train.py:1467-1557) over vertices [v_begin, v_end): reads all of
    x_in[V,3], writes rows v_begin..v_end-1 of x_out (in place on x_out; x_out may have padding rows past V)."""
    L = _lib.lib()
    x_in, normals = _f32(x_in, "x_in"), _f32(normals, "normals")
    edge_map, v_edges = _i32(edge_map, "edge_map"), _i32(v_edges, "v_edges")
    if x_out.dtype != torch.float32 or not x_out.is_contiguous() or x_out.device != x_in.device:
        raise _lib.FacetConvError("vertex_update_edges_range: x_out must be a contiguous float32 tensor on x_in's device")
    V = v_edges.shape[0]
    if x_in.numel() < 3 * V or x_out.numel() < 3 * V:
        raise _lib.FacetConvError("vertex_update_edges_range: x_in / x_out hold fewer than V = %d vertices" % V)
    with torch.cuda.device(x_in.device):
        check(L.fgc_vertex_update_edges_range(_p(x_in), _p(x_out), _p(normals), _p(edge_map), _p(v_edges), V,
                                              normals.numel() // 3, edge_map.numel() // 4, v_edges.numel() // V,
                                              int(v_begin), int(v_end), float(lam), _stream(x_in)),
              "fgc_vertex_update_edges_range")
    return x_out


def push_rows(src, dst_ptr: int, ids):
    """dst[ids[i]] = src[ids[i]] (rows of src.shape[-1] floats) where dst is given as a raw device pointer: another GPU's
    mapping of the same tensor (symmetric memory).  fgc_push_rows; runs on the current stream of src's device."""
    L = _lib.lib()
    src = _f32(src, "src")
    if ids.dtype != torch.int64 or not ids.is_contiguous() or ids.device != src.device:
        raise _lib.FacetConvError("push_rows: ids must be a contiguous int64 tensor on src's device")
    with torch.cuda.device(src.device):
        check(L.fgc_push_rows(_p(src), C.c_void_p(int(dst_ptr)), _p(ids), ids.numel(), int(src.shape[-1]), _stream(src)),
              "fgc_push_rows")


def vertex_update_ms(x, normals, faces, v_faces, scale, steps=2, iters=20):
    """One scale of reference Code/train.py:1668-1798 (update_position_MS)."""
    L = _lib.lib()
    x, normals = _f32(x, "x"), _f32(normals, "normals")
    faces, v_faces = _i32(faces, "faces"), _i32(v_faces, "v_faces")
    V = x.numel() // 3
    N0 = faces.numel() // 3
    max_faces = v_faces.numel() // V
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_vertex_update_ms_workspace(V, N0), x)
        check(L.fgc_vertex_update_ms(_p(x), _p(out), _p(normals), _p(faces), _p(v_faces), V, N0, max_faces,
                                     int(scale), int(steps), int(iters), _p(ws), ws.numel(), _stream(x)),
              "fgc_vertex_update_ms")
    return out


def _group_index(keys, valid, nbins):
    """CSR of the positions where `valid`, grouped by `keys` (ascending position inside a group): (ptr[nbins+1], ids)."""
    pos = torch.nonzero(valid.reshape(-1)).reshape(-1)
    k = keys.reshape(-1)[pos].to(torch.int64)
    order = torch.sort(k, stable=True).indices
    ptr = torch.zeros(nbins + 1, dtype=torch.int64, device=keys.device)
    ptr[1:] = torch.cumsum(torch.bincount(k, minlength=nbins), 0)
    return ptr.to(torch.int32), pos[order].to(torch.int32).contiguous()


def vertex_update_ms_lists(faces, v_faces, V, scale, steps=2):
    """Index lists of fgc_vertex_update_ms_bwd for one scale (depend on the mesh only: build once, reuse every step)."""
    faces, v_faces = _i32(faces, "faces").reshape(-1, 3), _i32(v_faces, "v_faces").reshape(V, -1)
    levels = int(scale) * int(steps)
    Fs = faces.shape[0] >> levels
    fc = torch.where(v_faces >= 0, v_faces >> levels, torch.full_like(v_faces, -1))
    slot_ptr, slot_id = _group_index(fc, (fc >= 0) & (fc < Fs), Fs)
    vert_ptr, vert_corner = _group_index(faces, faces >= 0, V)
    return slot_ptr, slot_id, vert_ptr, vert_corner


def vertex_update_ms_bwd(g_out, x, normals, faces, v_faces, scale, steps=2, iters=20, lists=None):
    """Backward of vertex_update_ms: (dL/dx_in[V,3], dL/dnormals[Fs,3])."""
    L = _lib.lib()
    x, normals, g_out = _f32(x, "x"), _f32(normals, "normals"), _f32(g_out, "g_out")
    faces, v_faces = _i32(faces, "faces"), _i32(v_faces, "v_faces")
    V = x.numel() // 3
    N0 = faces.numel() // 3
    max_faces = v_faces.numel() // V
    if lists is None:
        lists = vertex_update_ms_lists(faces, v_faces, V, scale, steps)
    slot_ptr, slot_id, vert_ptr, vert_corner = lists
    g_in, g_n = torch.empty_like(x), torch.empty_like(normals)
    with torch.cuda.device(x.device):
        ws = _ws(L.fgc_vertex_update_ms_bwd_workspace(V, N0, max_faces, int(iters)), x)
        check(L.fgc_vertex_update_ms_bwd(_p(x), _p(normals), _p(faces), _p(v_faces), V, N0, max_faces, int(scale), int(steps),
                                         int(iters), _p(slot_ptr), _p(slot_id), _p(vert_ptr), _p(vert_corner), _p(g_out),
                                         _p(g_in), _p(g_n), _p(ws), ws.numel(), _stream(x)), "fgc_vertex_update_ms_bwd")
    return g_in, g_n


# ----------------------------------------------------------------------------- whole-network inference forward
class NetPrepared:
    """Caller-owned tensor-core weight images of the reference network (fgc_net_prepare): depend only on the 44
    parameter tensors, so inference builds them once per checkpoint and reuses them for every patch batch."""

    def __init__(self, params):
        L = _lib.lib()
        if len(params) != int(L.fgc_net_param_count()):
            raise _lib.FacetConvError("NetPrepared: %d parameter tensors expected in creation order, got %d"
                                      % (int(L.fgc_net_param_count()), len(params)))
        self.params = [_f32(t, "param") for t in params]
        self.key = tuple((t.data_ptr(), t._version) for t in params)
        dev = self.params[0].device
        self.ptrs = (C.c_void_p * len(self.params))(*[t.data_ptr() for t in self.params])
        self.buf = torch.empty(int(L.fgc_net_prepared_bytes()), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(L.fgc_net_prepare(self.ptrs, len(self.params), _p(self.buf), self.buf.numel(), _stream(self.buf)),
                  "fgc_net_prepare")

    def matches(self, params) -> bool:
        return self.key == tuple((t.data_ptr(), t._version) for t in params)


def net_forward(x, adjs, prepared: NetPrepared, workspace: Optional[torch.Tensor] = None):
    """get_model_reg_multi_scale(x, adjs, multiScale=False) (reference Code/model.py:837-946) as one C-ABI call:
    x[B,N0,6], adjs = [adj0[B,N0,K], adj1[B,N0/4,K], adj2[B,N0/16,K]] -> y[B,N0,3] (before normalizeTensor)."""
    L = _lib.lib()
    x = _f32(x, "x")
    a0, a1, a2 = (_i32(a, "adj") for a in adjs)
    B, N0, Cin = x.shape
    K = a0.shape[2]
    if Cin != 6 or tuple(a0.shape) != (B, N0, K) or tuple(a1.shape) != (B, N0 // 4, K) or tuple(a2.shape) != (B, N0 // 16, K):
        raise _lib.FacetConvError("net_forward: x[B,N0,6], adj0[B,N0,K], adj1[B,N0/4,K], adj2[B,N0/16,K] expected, got %s %s %s %s"
                                  % (tuple(x.shape), tuple(a0.shape), tuple(a1.shape), tuple(a2.shape)))
    y = torch.empty((B, N0, 3), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        nws = int(L.fgc_net_fwd_workspace(B, N0, K))
        ws = workspace if (workspace is not None and workspace.numel() >= nws) else _ws(nws, x)
        check(L.fgc_net_fwd(B, N0, K, _p(x), _p(a0), _p(a1), _p(a2), prepared.ptrs, len(prepared.params), _p(prepared.buf),
                            _p(y), _p(ws), ws.numel(), _stream(x)), "fgc_net_fwd")
    return y


def launch_count() -> int:
    return int(_lib.lib().fgc_launch_count())


# ----------------------------------------------------------------------------- index builders
def build_faces_adj(faces, K=None, nv=None, kv=None):
    """GPU getFacesLargeAdj / getVerticesFaces (reference Code/utils.py:243-295, 370-395).
    faces[nf,3] int32 on the GPU -> (adj[nf,K] | None, v_faces[nv,kv] | None), bit-identical to the
    reference's host loops."""
    L = _lib.lib()
    faces = _i32(faces, "faces")
    if faces.dim() != 2 or faces.shape[1] != 3:
        raise _lib.FacetConvError("faces must be [nf,3], got %s" % (tuple(faces.shape),))
    nf = faces.shape[0]
    nv = int(faces.max().item()) + 1 if nv is None else int(nv)
    adj = torch.empty((nf, K), dtype=torch.int32, device=faces.device) if K else None
    vf = torch.empty((nv, kv), dtype=torch.int32, device=faces.device) if kv else None
    with torch.cuda.device(faces.device):
        ws = _ws(L.fgc_faces_adj_workspace(nf, nv), faces)
        check(L.fgc_build_faces_adj(_p(faces), nf, nv, int(K or 0), _p(adj) if adj is not None else None,
                                    _p(vf) if vf is not None else None, int(kv or 0), _p(ws), ws.numel(),
                                    _stream(faces)), "fgc_build_faces_adj")
    return adj, vf


def build_edge_maps(faces, max_edges=20, nv=None):
    """GPU getEdgeMap (reference Code/utils.py:91-183): (e_map[E,4], v_edges[nv,max_edges])."""
    L = _lib.lib()
    faces = _i32(faces, "faces")
    nf = faces.shape[0]
    nv = int(faces.max().item()) + 1 if nv is None else int(nv)
    e_map = torch.empty((3 * nf, 4), dtype=torch.int32, device=faces.device)
    v_e = torch.empty((nv, max_edges), dtype=torch.int32, device=faces.device)
    ne = C.c_int64(0)
    with torch.cuda.device(faces.device):
        ws = _ws(L.fgc_edge_maps_workspace(nf, nv), faces)
        check(L.fgc_build_edge_maps(_p(faces), nf, nv, int(max_edges), _p(e_map), C.byref(ne), _p(v_e), _p(ws),
                                    ws.numel(), _stream(faces)), "fgc_build_edge_maps")
    return e_map[: int(ne.value)], v_e


def face_features(verts, faces, normalize=True):
    """[unit normal | barycentre / bbox diagonal] per face (reference Code/utils.py:63-68, 1264-1294)."""
    L = _lib.lib()
    verts = _f32(verts, "verts")
    faces = _i32(faces, "faces")
    out = torch.empty((faces.shape[0], 6), dtype=torch.float32, device=verts.device)
    with torch.cuda.device(verts.device):
        ws = _ws(1024, verts)
        check(L.fgc_face_features(_p(verts), _p(faces), faces.shape[0], verts.shape[0], int(bool(normalize)), _p(out),
                                  _p(ws), ws.numel(), _stream(verts)), "fgc_face_features")
    return out
