"""Patch-sharded inference: the multi-GPU axis of the hot path.

The reference cuts a mesh into patches on the host and runs them *sequentially* through one
session (Code/train.py:100-126): every patch has its own adjacency pyramid, its own
``normalizeTensor`` global mean and no dependence on any other patch.  Here the same patch list
is dealt to one process per GPU; each rank runs whole patches with no collective on the data
path, and the per-facet results are merged exactly as the reference does on the host:

    out  = normals[oldToNew][:num_real]            train.py:117-121
    acc[patch_indices] += out                      train.py:126
    pred = normalize(acc)  (two passes, +1e-8, float64)   train.py:136, utils.py:26-35

Patches carry a ``core`` mask when they were cut with a halo (synthetic generators below): halo
facets give the network its receptive field, are computed redundantly, and are not written back.
Nothing here touches CUDA directly; the per-patch forward is a callable so the sharding / merge
logic is testable on CPU with ``gloo`` (tests/test_multi_rank_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import mesh


@dataclass
class Patch:
    x: np.ndarray                      # [N0, Cin] float32, node order of the pyramid (fake rows = 0)
    adjs: List[np.ndarray]             # [N_l, K] int32 per level, reference layout
    face_ids: np.ndarray               # [num_real] int64: global facet id of every real row (train.py:126)
    perm: Optional[np.ndarray] = None  # [N0] int32 oldToNew (train.py:117-121); None = identity
    core: Optional[np.ndarray] = None  # [num_real] bool: rows written back (halo rows dropped)

    @property
    def num_real(self) -> int:
        return int(self.face_ids.shape[0])

    @property
    def cost(self) -> int:
        return int(self.x.shape[0])


def partition(costs: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of patches to ranks; deterministic
    (ties broken by patch index) so every rank derives the same plan without communication."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    load = [0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        plan[r].append(i)
        load[r] += int(costs[i])
    for p in plan:
        p.sort()
    return plan


def patch_output(p: Patch, normals: np.ndarray):
    """Un-permute / trim one patch's network output (train.py:117-121); returns (ids, rows)."""
    out = np.asarray(normals).reshape(-1, normals.shape[-1])
    if p.perm is not None:
        out = out[p.perm]
    out = out[: p.num_real]
    ids = p.face_ids
    if p.core is not None:
        out, ids = out[p.core], ids[p.core]
    return ids, out


def host_normalize(a: np.ndarray) -> np.ndarray:
    """reference Code/utils.py:26-35 applied as train.py:136 does: two passes, +1e-8, float64."""
    a = np.asarray(a, np.float64)
    for _ in range(2):
        n = np.sqrt((a * a).sum(axis=1))[:, None] + 0.00000001
        a = a * (1 / n)
    return a


def merge(num_faces: int, contributions) -> np.ndarray:
    """Sum overlapping per-facet normals then renormalise (train.py:98,126,136)."""
    acc = np.zeros((num_faces, 3), np.float64)
    for ids, rows in contributions:
        np.add.at(acc, np.asarray(ids, np.int64), np.asarray(rows, np.float64))
    return host_normalize(acc)


def run_local(patches: Sequence[Patch], my: Sequence[int], forward: Callable[[Patch], np.ndarray]):
    """Runs this rank's patches; returns the concatenated (ids[int64], rows[float32]) it owns."""
    ids, rows = [np.zeros((0,), np.int64)], [np.zeros((0, 3), np.float32)]
    for i in my:
        a, b = patch_output(patches[i], forward(patches[i]))
        ids.append(np.asarray(a, np.int64))
        rows.append(np.asarray(b, np.float32))
    return np.concatenate(ids), np.concatenate(rows)


def gather_to_all(ids: np.ndarray, rows: np.ndarray, device=None):
    """all_gather of ragged per-rank results (no data-path collective happened before this point).
    Works with gloo (CPU tensors) and nccl (``device`` = the rank's CUDA device)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(ids, rows)]
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    n = torch.tensor([ids.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    cap = max(int(t.item()) for t in ns)
    pid = torch.zeros(cap, dtype=torch.int64, device=dev)
    prow = torch.zeros(cap, 3, dtype=torch.float32, device=dev)
    pid[: ids.shape[0]] = torch.from_numpy(ids).to(dev)
    prow[: rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(rows, np.float32)).to(dev)
    gid = [torch.zeros_like(pid) for _ in range(world)]
    grow = [torch.zeros_like(prow) for _ in range(world)]
    dist.all_gather(gid, pid)
    dist.all_gather(grow, prow)
    out = []
    for r in range(world):
        k = int(ns[r].item())
        out.append((gid[r][:k].cpu().numpy(), grow[r][:k].cpu().numpy()))
    return out


def infer_sharded(patches: Sequence[Patch], num_faces: int, forward: Callable[[Patch], np.ndarray],
                  rank: int = 0, world: int = 1, device=None) -> np.ndarray:
    """Patch-parallel inference: partition -> local forwards -> gather -> host merge."""
    plan = partition([p.cost for p in patches], world)
    ids, rows = run_local(patches, plan[rank], forward)
    return merge(num_faces, gather_to_all(ids, rows, device))


# ----------------------------------------------------------------------------- batching
def pad_patch(p: Patch, n0: int):
    """(x[n0,Cin], [adj_l[n0 >> 2l, K]]) of patch `p` grown to n0 level-0 nodes with fake nodes: zero
    features and a self-only adjacency row at every level -- the state the reference's coarsening
    leaves fake nodes in.  Fake rows never appear in a real row's adjacency,
    so the real rows of every layer are bit-identical to the unpadded run."""
    N = p.x.shape[0]
    assert n0 >= N and n0 % 16 == 0
    x = np.zeros((n0,) + p.x.shape[1:], p.x.dtype)
    x[:N] = p.x
    adjs = []
    for lvl, a in enumerate(p.adjs):
        nl, tl = a.shape[0], n0 >> (2 * lvl)
        out = np.zeros((tl, a.shape[1]), a.dtype)
        out[:nl] = a
        out[nl:, 0] = np.arange(nl + 1, tl + 1)
        adjs.append(out)
    return x, adjs


def batch_patches(patches: Sequence[Patch], idx: Sequence[int]):
    """Stacks the patches `idx` into one batch (x[B,n0,Cin], [adj_l[B,n_l,K]]): the layers treat batch
    elements independently (adjacency ids are per element), so one launch serves B patches.  The
    per-patch normalizeTensor (a global mean per patch) is applied to the rows [:N_b] of element b
    by the caller."""
    n0 = max(patches[i].x.shape[0] for i in idx)
    n0 = (n0 + 15) // 16 * 16
    xs, adjs = [], None
    for i in idx:
        x, a = pad_patch(patches[i], n0)
        xs.append(x)
        adjs = [[t] for t in a] if adjs is None else [u + [t] for u, t in zip(adjs, a)]
    return np.stack(xs), [np.stack(u) for u in adjs]


# ----------------------------------------------------------------------------- synthetic patches (C3 / C5)
def grid_patches(nx: int, ny: int, block: int = 100, halo: int = 3, K: int = 16, height=None, noise: float = 0.3,
                 seed: int = 0, only: Optional[Sequence[int]] = None):
    """Cuts an open nx x ny-quad height field (2*nx*ny triangles) into block x block-quad patches
    grown by `halo` quads on every side (3 levels x 2 convs of receptive field),
    each with a Morton-ordered facet list, reference-layout adjacency (getFacesLargeAdj semantics,
    deduplicated as the reference's sparse round trip leaves it) and the analytic 3-level
    binary-tree pyramid of mesh.build_pyramid.  Global facet id of quad (i,j), triangle t is
    2*(j*nx+i)+t.  Returns (patches, num_faces); `only` restricts generation to some patch indices
    (every rank can generate just its own share)."""
    bx = (nx + block - 1) // block
    by = (ny + block - 1) // block
    out = []
    rs = np.random.RandomState(seed)
    h = height or (lambda X, Y: 0.05 * np.sin(6.0 * X) * np.cos(4.0 * Y))
    todo = range(bx * by) if only is None else only
    for pi in todo:
        pj, pi_ = divmod(pi, bx)
        i0, i1 = pi_ * block, min(nx, (pi_ + 1) * block)
        j0, j1 = pj * block, min(ny, (pj + 1) * block)
        a0, a1 = max(0, i0 - halo), min(nx, i1 + halo)
        b0, b1 = max(0, j0 - halo), min(ny, j1 + halo)
        lx, ly = a1 - a0, b1 - b0
        jj, ii = np.meshgrid(np.arange(b0, b1 + 1), np.arange(a0, a1 + 1), indexing="ij")
        X, Y = ii / nx, jj / ny
        V = np.stack([X, Y, h(X, Y)], -1).reshape(-1, 3).astype(np.float64)
        qj, qi = np.meshgrid(np.arange(ly), np.arange(lx), indexing="ij")
        qi, qj = qi.reshape(-1), qj.reshape(-1)
        order = np.argsort(mesh._morton2(qi, qj), kind="stable")
        qi, qj = qi[order], qj[order]
        vx = lx + 1
        v00, v10, v01, v11 = qj * vx + qi, qj * vx + qi + 1, (qj + 1) * vx + qi, (qj + 1) * vx + qi + 1
        F = np.empty((qi.size * 2, 3), np.int32)
        F[0::2] = np.stack([v00, v10, v11], 1)
        F[1::2] = np.stack([v00, v11, v01], 1)
        if noise:
            e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], 0)
            ml = np.linalg.norm(V[e[:, 0]] - V[e[:, 1]], axis=1).mean()
            V = V + np.random.RandomState(seed * 7919 + pi).normal(0.0, noise * ml, size=V.shape)
        feat = mesh.face_features(V, F).astype(np.float32)
        adj = mesh.dedup_adj(mesh.faces_large_adj(F, K))
        gi, gj = qi + a0, qj + b0
        fid = np.empty(F.shape[0], np.int64)
        fid[0::2] = 2 * (gj * nx + gi)
        fid[1::2] = 2 * (gj * nx + gi) + 1
        core_q = (gi >= i0) & (gi < i1) & (gj >= j0) & (gj < j1)
        core = np.repeat(core_q, 2)
        featp, adjp = mesh.pad_to_multiple(feat, adj, 16)
        out.append(Patch(x=featp, adjs=mesh.build_pyramid(adjp, 3, K), face_ids=fid, perm=None, core=core))
    del rs
    return out, 2 * nx * ny


# ----------------------------------------------------------------------------- sharded vertex update (BASELINE config C5)
def vertex_ranges(V: int, world: int):
    """Equal contiguous vertex ranges, one per rank (the last ones may be short or empty): [(begin, end)] and the chunk."""
    chunk = (V + world - 1) // world
    return [(min(V, r * chunk), min(V, (r + 1) * chunk)) for r in range(world)], chunk


class VertexHalo:
    """Which vertex positions a rank of the sharded update must receive from / send to its peers every sweep: the
    vertices its own range reads through update_position2's index tensors (the two ends of every edge around an owned
    vertex) that another rank owns.  Built once per mesh from v_edges / edge_map (one exchange of id lists)."""

    def __init__(self, edge_map, v_edges, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        V = int(v_edges.shape[0])
        ranges, chunk = vertex_ranges(V, self.world)
        self.begin, self.end = ranges[self.rank]
        dev = v_edges.device
        # getEdgeMap (reference Code/utils.py:91-183) lists under a vertex the edges incident to it, so what an owned vertex
        # reads is the other end of each of its edges: one pass over the edge list finds the edges with exactly one owned
        # end (a scan of the range's v_edges rows would sort 20 ids per vertex to find the same thin band)
        a, bb = edge_map[:, 0].long(), edge_map[:, 1].long()
        own_a, own_b = (a >= self.begin) & (a < self.end), (bb >= self.begin) & (bb < self.end)
        need = torch.unique(torch.cat([bb[own_a & ~own_b], a[own_b & ~own_a]]))       # sorted => grouped by owner
        recv_counts = torch.bincount(need // chunk, minlength=self.world).to(torch.int64)
        send_counts = torch.empty_like(recv_counts)
        _all_to_all(send_counts, recv_counts, [1] * self.world, [1] * self.world, group)
        self.recv_splits = [int(c) for c in recv_counts.tolist()]
        self.send_splits = [int(c) for c in send_counts.tolist()]
        self.need = need
        self.send_ids = torch.empty(sum(self.send_splits), dtype=torch.int64, device=dev)
        _all_to_all(self.send_ids, need, self.send_splits, self.recv_splits, group)     # peers tell me what they need
        if self.send_ids.numel() and (int(self.send_ids.min()) < self.begin or int(self.send_ids.max()) >= self.end):
            raise RuntimeError("VertexHalo: a peer asked for a vertex this rank does not own")

    def exchange(self, x):
        """x[*,3]: owned rows are current; fills the halo rows from their owners (one all-to-all)."""
        import torch
        send = x[self.send_ids]
        recv = torch.empty((self.need.numel(), 3), dtype=x.dtype, device=x.device)
        _all_to_all(recv, send, [3 * c for c in self.recv_splits], [3 * c for c in self.send_splits], self.group, cols=3)
        x[self.need] = recv


def _all_to_all(out, inp, out_splits, in_splits, group, cols=1):
    """all_to_all_single with uneven splits (NCCL); gloo (CPU tests) goes through all_gather of the whole send buffers."""
    import torch
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(out.reshape(-1), inp.reshape(-1).contiguous(), out_splits, in_splits, group=group)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, (in_splits, inp.reshape(-1).cpu()), group=group)
    parts = []
    for src in range(world):
        spl, buf = sizes[src]
        off = sum(spl[:rank])
        parts.append(buf[off:off + spl[rank]])
    flat = torch.cat(parts) if parts else inp.reshape(-1)[:0]
    out.reshape(-1).copy_(flat.to(out.device))


def vertex_update_edges_sharded(x, normals, edge_map, v_edges, iters=60, lam=1.0 / 18, group=None, sweep=None,
                                exchange="halo"):
    """update_position2 (reference Code/train.py:1467-1557) with the vertices of ONE large mesh sharded over the ranks of
    `group`: every rank holds the whole index tensors, sweeps its own contiguous vertex range
    (fgc_vertex_update_edges_range) and after every Jacobi sweep receives the positions its next sweep reads from other
    ranks -- the only exchange step of the path.  exchange = "halo": one all-to-all of exactly those vertices (the ends of
    the edges around owned vertices; VertexHalo), and one all-gather of the ranges at the end; "allgather": all positions
    after every sweep; "p2p" (NCCL groups on one node): the ping-pong position buffers live in symmetric memory, after its
    sweep every rank stores the rows its peers read straight into THEIR buffers over NVLink (fgc_push_rows) and a
    device-side barrier closes the sweep -- no collective call, no host synchronisation inside the loop.  A Jacobi sweep reads nothing but the previous sweep's positions, so both are bit-identical to the
    single-device update.  `sweep(x_in, x_out, begin, end)` replaces the CUDA sweep in CPU tests.
    x[V,3] -> x[V,3] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    V = int(v_edges.shape[0])
    if sweep is None:
        from . import ops

        def sweep(x_in, x_out, b, e):
            ops.vertex_update_edges_range(x_in, x_out, normals, edge_map, v_edges, b, e, lam)
    ranges, chunk = vertex_ranges(V, world)
    b, e = ranges[rank]
    cur = torch.zeros((world * chunk, 3), dtype=x.dtype, device=x.device)   # padded to equal shards for the all-gather
    cur[:V] = x.reshape(V, 3)
    nxt = cur.clone()

    def gather_all(buf):
        mine = buf[rank * chunk:(rank + 1) * chunk]
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(buf, mine, group=group)            # in place: shard r lands at rows r * chunk
        else:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine.clone(), group=group)
            buf.copy_(torch.cat(parts, 0))

    if world > 1 and exchange == "p2p":
        return _vertex_update_p2p(x, V, b, e, chunk, edge_map, v_edges, int(iters), group, sweep, gather_all)
    halo = VertexHalo(edge_map, v_edges, group) if (world > 1 and exchange == "halo") else None
    for _ in range(int(iters)):
        sweep(cur, nxt, b, e)
        if halo is not None:
            halo.exchange(nxt)
        elif world > 1:
            gather_all(nxt)
        cur, nxt = nxt, cur
    if halo is not None:
        gather_all(cur)
    return cur[:V].clone()


_P2P_BUFFERS = {}


def _vertex_update_p2p(x, V, b, e, chunk, edge_map, v_edges, iters, group, sweep, gather_all):
    """exchange = "p2p" of vertex_update_edges_sharded: symmetric-memory ping-pong buffers, direct peer stores of the halo
    rows, one device-side barrier per sweep.  Hazards: sweep s reads buffer A and writes own rows of B; the pushes of
    sweep s write halo rows of the peers' B; the barrier orders them before sweep s+1 reads B and before the pushes of
    sweep s+1 touch A, whose last readers (sweep s) are then done on every rank."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    from . import ops
    grp = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(grp), dist.get_rank(grp)
    halo = VertexHalo(edge_map, v_edges, group)
    rows = world * chunk
    # the symmetric allocation and its rendezvous (an exchange of IPC handles, ~100 ms) are kept per (group, size): a
    # pipeline that updates scan after scan pays them once
    key = (id(grp), rows, str(x.device))
    if key not in _P2P_BUFFERS:
        buf = symm_mem.empty((2, rows, 3), dtype=torch.float32, device=x.device)
        _P2P_BUFFERS.clear()
        _P2P_BUFFERS[key] = (buf, symm_mem.rendezvous(buf, grp))
    buf, hdl = _P2P_BUFFERS[key]
    hdl.barrier()              # the previous call's readers are done on every rank
    buf.zero_()
    buf[0, :V] = x.reshape(V, 3)
    buf[1].copy_(buf[0])
    peers, off = [], 0
    for p_, n_ in enumerate(halo.send_splits):
        if n_:
            peers.append((p_, halo.send_ids[off:off + n_].contiguous()))
        off += n_
    plane = rows * 3 * 4
    hdl.barrier()
    for it in range(iters):
        ci = it & 1
        sweep(buf[ci], buf[1 - ci], b, e)
        for p_, ids in peers:
            ops.push_rows(buf[1 - ci], int(hdl.buffer_ptrs[p_]) + (1 - ci) * plane, ids)
        hdl.barrier()
    out = buf[iters & 1].clone()
    gather_all(out)
    return out[:V].clone()
