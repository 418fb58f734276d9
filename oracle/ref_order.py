"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.

CPU port of the facet-graph convolution that follows the reference's *evaluation order*
(reference Code/model.py:427-504 with :74-95 and :380-405): W.x for every facet first, then
the materialised [B,N,K,M*Cout] gather, the transposes, the q-multiply and the two reductions.
This is what ``bench.py`` times as the CPU baseline (``cpu_baseline.kind = "port"`` and the
``--impl reference`` arm): TensorFlow is not installable in this image and the reference's
Python sources cannot travel to the GPU box, so the reference's own op sequence is restated on
torch-CPU tensors with every host thread available.  Backward = torch autograd of this
sequence (the analogue of TF autodiff).  It is validated against the golden vectors in
tests/test_oracle_golden.py::test_ref_order_port.
"""
from __future__ import annotations

import torch


def _pad_gather(t, adj):
    """concat([0-row, t])[adj] per batch element (model.py:380-399)."""
    B = t.shape[0]
    zeros = torch.zeros(B, 1, t.shape[2], dtype=t.dtype)
    tp = torch.cat([zeros, t], dim=1)
    return torch.stack([tp[b][adj[b].long()] for b in range(B)], dim=0)


def conv_fwd(x, adj, W0, b, u, v, c, bias_mask=True):
    """x[B,N,Cin] float32 CPU tensor, adj[B,N,K] int; returns y[B,N,Cout]."""
    B, N, Cin = x.shape
    M, Cout, _ = W0.shape
    K = adj.shape[2]
    cnt = (adj != 0).sum(dim=2)
    nz = cnt != 0
    inv = torch.where(nz, 1.0 / cnt.clamp(min=1).to(x.dtype), torch.zeros((), dtype=x.dtype))
    inv = inv.reshape(B, N, 1, 1)
    xt = x.transpose(1, 2)                                        # [B,Cin,N]      (:463)
    wx = torch.matmul(W0.reshape(M * Cout, Cin), xt).transpose(1, 2)   # [B,N,M*Cout]   (:464-468)
    patches = _pad_gather(wx, adj)                                # [B,N,K,M*Cout] (:470)
    ux = torch.matmul(u, xt)                                      # [B,M,N]        (:79)
    vx = torch.matmul(v, xt).transpose(1, 2)                      # [B,N,M]        (:80-82)
    lg = _pad_gather(vx, adj).permute(2, 0, 3, 1)                 # [K,B,M,N]      (:84-86)
    lg = (lg + ux).permute(0, 1, 3, 2) + c                        # [K,B,N,M]      (:88-91)
    q = torch.softmax(lg.permute(1, 2, 0, 3), dim=-1)             # [B,N,K,M]      (:93-94)
    p5 = patches.reshape(B, N, K, M, Cout).permute(4, 0, 1, 2, 3)  # [Cout,B,N,K,M] (:482-484)
    p5 = (q * p5).permute(1, 2, 3, 4, 0)                          # [B,N,K,M,Cout] (:485-486)
    y = (p5.sum(dim=2) * inv).sum(dim=2)                          # (:488-493)
    if bias_mask:
        y = torch.where(nz.unsqueeze(-1), y + b, y)               # (:496-498)
    else:
        y = y + b
    return y


def conv_fwd_bwd(x, adj, gy, W0, b, u, v, c, bias_mask=True):
    """One forward + backward pass; returns (y, gx, gW0, gb, gu, gv, gc)."""
    leaves = [t.detach().clone().requires_grad_(True) for t in (x, W0, b, u, v, c)]
    y = conv_fwd(leaves[0], adj, leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], bias_mask)
    grads = torch.autograd.grad(y, leaves, grad_outputs=gy)
    return (y.detach(),) + tuple(grads)
