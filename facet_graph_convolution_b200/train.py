"""Data-parallel training step of the denoising network.

Restates the *step* of reference ``Code/train.py:trainNet`` (:380-632) -- not its TF session
driver: random rotation of the inputs and of the ground-truth normals (:436-451, matrix from
``utils.py:2034-2074``), network forward, ``normalizeTensor`` (:503), 4000 facet ids sampled with
replacement (:561, :509-515), ``faceNormalsLoss`` (:517), Adam with TF defaults (:520).

Multi-GPU: replicas only.  Every rank steps on its own patches; the only exchange is ONE
all-reduce(sum) per step over a single flat fp32 gradient bucket (474 199 floats = 1.9 MB for the
default network; latency-bound on NVLink, so no bucketing/overlap machinery), then 1/world.
The reference has batch size 1 and no data parallelism (train.py:404-405); with B patches per
rank the loss is the mean over the patches of the step.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

COST_SAMPLES = 4000  # reference Code/train.py:415


def rand_rotation_matrix(rng: np.random.RandomState, deflection: float = 1.0) -> np.ndarray:
    """Uniform random rotation after Arvo's Graphics Gems III method, the construction used by
    reference Code/utils.py:2034-2074: a rotation about z followed by the Householder reflection
    (v v^T - I) with |v| = sqrt(2)."""
    theta, phi, z = rng.uniform(size=3)
    theta *= 2.0 * deflection * math.pi
    phi *= 2.0 * math.pi
    z *= 2.0 * deflection
    r = math.sqrt(z)
    v = np.array([math.sin(phi) * r, math.cos(phi) * r, math.sqrt(2.0 - z)])
    st, ct = math.sin(theta), math.cos(theta)
    Rz = np.array([[ct, st, 0.0], [-st, ct, 0.0], [0.0, 0.0, 1.0]])
    return (np.outer(v, v) - np.eye(3)).dot(Rz)


def rotate_features(x: torch.Tensor, R: torch.Tensor) -> torch.Tensor:
    """x[..., 3g] -> every consecutive xyz triple multiplied by R (train.py:444-451, channels % 3 == 0)."""
    C = x.shape[-1]
    if C % 3 != 0:
        raise ValueError("rotation augmentation needs a multiple of 3 channels (normal | position), got %d" % C)
    xr = x.reshape(*x.shape[:-1], C // 3, 3)
    return torch.matmul(xr, R.t().to(x)).reshape(x.shape)


class GradBucket:
    """One flat fp32 buffer holding every parameter gradient, in the reference's variable-creation
    order (W0,b,u,c,v per conv; W,b per linear), so a step needs exactly one all-reduce."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.params = list(params)
        if not self.params:
            raise ValueError("GradBucket: empty parameter list (a step over it would update nothing)")
        self.sizes = [int(p.numel()) for p in self.params]
        self.total = sum(self.sizes)
        dev = self.params[0].device
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        o = 0
        for p, n in zip(self.params, self.sizes):
            self.views.append(self.flat[o:o + n].view_as(p))
            o += n

    def pack(self):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        return self.flat

    def all_reduce_mean(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / dist.get_world_size(group))
        return self.flat

    def unpack(self):
        for p, v in zip(self.params, self.views):
            p.grad = v  # the optimizer reads straight from the bucket


class Adam:
    """tf.train.AdamOptimizer() defaults (lr 1e-3, beta 0.9/0.999, eps 1e-8; train.py:520) applied to
    the flat bucket in one fused pass over a flat copy of the parameters."""

    def __init__(self, bucket: GradBucket, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.b, self.lr, self.b1, self.b2, self.eps, self.t = bucket, lr, b1, b2, eps, 0
        self.m = torch.zeros_like(bucket.flat)
        self.v = torch.zeros_like(bucket.flat)

    @torch.no_grad()
    def step(self):
        self.t += 1
        g = self.b.flat
        self.m.mul_(self.b1).add_(g, alpha=1 - self.b1)
        self.v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
        # TF's formulation: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        upd = self.m / (self.v.sqrt() + self.eps)
        o = 0
        for p, n in zip(self.b.params, self.b.sizes):
            p.add_(upd[o:o + n].view_as(p), alpha=-lr_t)
            o += n


    def state_lists(self):
        """(first moments, second moments) per parameter, creation order -- the Saver's `/Adam`, `/Adam_1` slots."""
        ms, vs, o = [], [], 0
        for p, n in zip(self.b.params, self.b.sizes):
            ms.append(self.m[o:o + n].view_as(p))
            vs.append(self.v[o:o + n].view_as(p))
            o += n
        return ms, vs

    @torch.no_grad()
    def load_state(self, ms, vs, step: int):
        """Resume from per-parameter moments and the number of updates already done
        (`checkpoint.load_training_state`)."""
        o = 0
        for p, n, m, v in zip(self.b.params, self.b.sizes, ms, vs):
            self.m[o:o + n].copy_(torch.as_tensor(m, dtype=self.m.dtype).reshape(-1))
            self.v[o:o + n].copy_(torch.as_tensor(v, dtype=self.v.dtype).reshape(-1))
            o += n
        self.t = int(step)


def loss_on_patch(forward, x, adjs, gt, rng: np.random.RandomState, samples: int = COST_SAMPLES,
                  augment: bool = True):
    """One patch of the training objective.  `forward(x, adjs)` returns the raw network output
    [1,N0,3]; normalisation, sampling and the loss follow train.py:503-517."""
    from . import model as fm
    if augment:
        R = torch.from_numpy(rand_rotation_matrix(rng).astype(np.float32)).to(x.device)
        x = rotate_features(x, R)
        gt = rotate_features(gt, R)
    n = fm.normalizeTensor(forward(x, adjs))
    idx = torch.from_numpy(rng.randint(x.shape[1], size=samples)).to(x.device)
    return fm.faceNormalsLoss(n[:, idx, :].contiguous(), gt[:, idx, :].contiguous())


@torch.no_grad()
def sync_replicas(bucket: GradBucket, opt: Optional[Adam] = None, group=None, src: int = 0, check_only: bool = False):
    """Data-parallel replicas must start identical: train_step only averages GRADIENTS, so replicas whose initial
    parameters (unseeded random init per rank) or Adam state (a rank that resumed from a checkpoint) differ would
    drift apart silently while every step reports a finite loss.  Broadcasts parameters, both Adam moments and the
    step count from rank `src` (one flat buffer each); with check_only=True nothing is overwritten and a mismatch
    raises instead.  No-op without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.detach().reshape(-1) for p in bucket.params])
    bufs = [flat] + ([opt.m, opt.v, torch.tensor([float(opt.t)], device=flat.device)] if opt is not None else [])
    if check_only:
        for b in bufs:
            ref = b.clone()
            dist.broadcast(ref, src, group=group)
            bad = torch.tensor([0.0 if torch.equal(ref, b) else 1.0], device=flat.device)
            dist.all_reduce(bad, group=group)
            if float(bad.item()) > 0:
                raise RuntimeError("data-parallel replicas differ (parameters or optimizer state): call "
                                   "sync_replicas(bucket, opt) before the first train_step")
        return
    for b in bufs:
        dist.broadcast(b, src, group=group)
    o = 0
    for p, n in zip(bucket.params, bucket.sizes):
        p.copy_(flat[o:o + n].view_as(p))
        o += n
    if opt is not None:
        opt.t = int(bufs[3].item())


_STACK_CACHE = {}
_STACK_CACHE_MAX = 16


def _stacked_adjs(batch):
    """Adjacency pyramids of the batch concatenated along the batch axis.  Cached per batch composition
    (the caches of the reverse adjacency and of the tile plans are keyed by tensor identity, so a batch
    that comes back -- every epoch -- must come back as the same tensors)."""
    key = tuple((a.data_ptr(), a._version, tuple(a.shape)) for _, adjs, _ in batch for a in adjs)
    hit = _STACK_CACHE.get(key)
    if hit is not None:
        return hit[0]
    adjs = [torch.cat([b[1][lvl] for b in batch], dim=0) for lvl in range(len(batch[0][1]))]
    if len(_STACK_CACHE) >= _STACK_CACHE_MAX:
        _STACK_CACHE.pop(next(iter(_STACK_CACHE)))
    _STACK_CACHE[key] = (adjs, [b[1] for b in batch])   # keep the parts alive: their pointers are the key
    return adjs


def _stackable(batch) -> bool:
    x0, a0, g0 = batch[0]
    return len(batch) > 1 and all(x.shape == x0.shape and g.shape == g0.shape and len(a) == len(a0) and
                                  all(u.shape == v.shape for u, v in zip(a, a0)) for x, a, g in batch)


def loss_on_stacked_patches(forward, batch, rng: np.random.RandomState, samples: int = COST_SAMPLES,
                            augment: bool = True):
    """Mean over the patches of `batch` of the per-patch objective, with ONE network forward over the
    patches stacked along the batch axis (the layers treat batch elements independently; the
    per-patch normalizeTensor and sampling follow train.py:503-517 element by element).  Random
    numbers are drawn in the order of the patch-by-patch loop (rotation, then sample ids)."""
    from . import model as fm
    xs, gts, ids = [], [], []
    for x, _, gt in batch:
        if augment:
            R = torch.from_numpy(rand_rotation_matrix(rng).astype(np.float32)).to(x.device)
            x, gt = rotate_features(x, R), rotate_features(gt, R)
        xs.append(x)
        gts.append(gt)
        ids.append(torch.from_numpy(rng.randint(x.shape[1], size=samples)).to(x.device))
    y = forward(torch.cat(xs, dim=0), _stacked_adjs(batch))
    total = None
    for b, (gt, idx) in enumerate(zip(gts, ids)):
        n = fm.normalizeTensor(y[b:b + 1].contiguous())
        lb = fm.faceNormalsLoss(n[:, idx, :].contiguous(), gt[:, idx, :].contiguous()) / len(batch)
        total = lb if total is None else total + lb
    return total


def train_step(net, batch, bucket: GradBucket, opt: Adam, rng: np.random.RandomState, group=None,
               samples: int = COST_SAMPLES, augment: bool = True, stack: bool = True) -> float:
    """forward + backward over this rank's `batch` of (x[1,N0,Cin], adjs, gt[1,N0,3]) patches,
    one all-reduce of the flat gradient bucket, Adam.  Returns the rank-local mean loss.
    Equal-sized patches are stacked into one forward/backward (`stack`); ragged ones run one by one.
    Requires identical replicas (parameters and Adam state): call sync_replicas(bucket, opt) once after
    construction / resume."""
    for p in bucket.params:
        p.grad = None
    total = 0.0
    if stack and _stackable(batch):
        loss = loss_on_stacked_patches(net, batch, rng, samples, augment)
        loss.backward()
        total = float(loss.detach())
    else:
        for x, adjs, gt in batch:
            loss = loss_on_patch(net, x, adjs, gt, rng, samples, augment) / len(batch)
            loss.backward()
            total += float(loss.detach())
    bucket.pack()
    bucket.all_reduce_mean(group)
    opt.step()
    return total

# ----------------------------------------------------------------------------- vertex-space trainers
SAMP_NUM = 500               # reference Code/train.py:653, :935
MS_ITERS = (80, 20, 20)      # reference Code/train.py:771, :1089


class VertexPatch:
    """What one step of the vertex-space trainers feeds (Code/train.py:824-842): face features and adjacency pyramid,
    noisy and ground-truth vertices, the patch's faces and per-vertex face lists, optionally ground-truth face normals.
    The index lists of the vertex update's backward depend on the mesh only and are built on first use."""

    def __init__(self, x, adjs, verts, gt_verts, faces, v_faces, gt_normals=None, steps: int = 2):
        self.x, self.adjs, self.verts, self.gt_verts = x, adjs, verts, gt_verts
        self.faces, self.v_faces, self.gt_normals, self.steps = faces, v_faces, gt_normals, steps
        self._lists = None

    def index_lists(self):
        from . import ops
        if self._lists is None:
            V = self.verts.reshape(-1, 3).shape[0]
            self._lists = [ops.vertex_update_ms_lists(self.faces, self.v_faces, V, sc, self.steps) for sc in range(3)]
        return self._lists


def vertex_loss_on_patch(forward_ms, patch: VertexPatch, rng: np.random.RandomState, samples: int = SAMP_NUM,
                         augment: bool = True, double_loss: bool = False, iters=MS_ITERS, sample_ids=None):
    """The objective of trainAccuracyNet (Code/train.py:741-781) and, with `double_loss`, of trainDoubleLossNet
    (:1060-1102) for one patch: random rotation of the features and of both vertex sets, multi-scale network
    (`forward_ms(x, adjs)` returns the three heads), normalizeTensor of the FINE head only (:767), update_position_MS
    over the three heads, fullLoss between the refined and the ground-truth vertices (+ faceNormalsLoss of the fine head).
    Sample ids are drawn per step with replacement (:833-834) unless given."""
    from . import model as fm
    x, verts, gtv, gtn = patch.x, patch.verts, patch.gt_verts, patch.gt_normals
    if augment:
        R = torch.from_numpy(rand_rotation_matrix(rng).astype(np.float32)).to(x.device)
        x, verts, gtv = rotate_features(x, R), rotate_features(verts, R), rotate_features(gtv, R)
        if gtn is not None:
            gtn = rotate_features(gtn, R)
    if sample_ids is None:
        nv, ng = verts.reshape(-1, 3).shape[0], gtv.reshape(-1, 3).shape[0]
        sample_ids = (torch.from_numpy(rng.randint(nv, size=samples).astype(np.int32)).to(x.device),
                      torch.from_numpy(rng.randint(ng, size=samples).astype(np.int32)).to(x.device))
    n0, n1, n2 = forward_ms(x, patch.adjs)
    n0 = fm.normalizeTensor(n0)
    refined, _ = fm.update_position_MS(verts, [n0, n1, n2], patch.faces, patch.v_faces, patch.steps, iter_num_list=iters,
                                       index_lists=patch.index_lists())
    loss = fm.fullLoss(refined, gtv.reshape(1, -1, 3), sample_ids[0], sample_ids[1])
    if double_loss:
        if gtn is None:
            raise ValueError("double_loss needs ground-truth face normals")
        loss = loss + fm.faceNormalsLoss(n0, gtn)
    return loss


def train_step_vertices(net, batch: Sequence[VertexPatch], bucket: GradBucket, opt: Adam, rng: np.random.RandomState,
                        group=None, samples: int = SAMP_NUM, augment: bool = True, double_loss: bool = False,
                        iters=MS_ITERS) -> float:
    """One optimisation step of trainAccuracyNet / trainDoubleLossNet over this rank's patches (the reference steps on
    one patch at a time, :842; a batch averages the patch losses), then the same single all-reduce + Adam as train_step.
    `net` is a multi-scale DenoisingNet (three heads)."""
    for p in bucket.params:
        p.grad = None
    total = 0.0
    for patch in batch:
        loss = vertex_loss_on_patch(net, patch, rng, samples, augment, double_loss, iters) / len(batch)
        loss.backward()
        total += float(loss.detach())
    bucket.pack()
    bucket.all_reduce_mean(group)
    opt.step()
    return total
