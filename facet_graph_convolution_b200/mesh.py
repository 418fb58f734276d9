"""Host-side synthetic mesh generators that emit the reference's index layouts directly.

The CUDA kernels consume index tensors whose layout is defined by the reference's
NumPy preprocessing.  The reference builds them with per-face
Python loops (hours at 20 M faces), so the bench/test workloads need vectorised
generators that produce *the same layouts*:

  * facet adjacency  adj[N,K] int32, 1-indexed, 0 = padding, column 0 = self, raw
    vertex-sharing list with duplicates        (reference Code/utils.py:243-295)
  * edge map e_map[E,4] = (v1,v2,f1,f2|-1) and v_e_map[V,max_edges]|-1
                                               (reference Code/utils.py:91-183)
  * vertex->faces v_f[V,k_v]|-1                (reference Code/utils.py:370-395)
  * face normals (two-pass normalise)          (reference Code/utils.py:63-68, 26-35)
  * barycentres divided by the bbox diagonal   (reference Code/utils.py:1264-1294)
  * a binary-tree pyramid (4 consecutive fine nodes -> 1 coarse node, fake nodes =
    zero features + self-only adjacency)       (layout of Code/dataClasses.py:112-146)

Everything here is NumPy on the host: it feeds the hot path, it is not the hot path.
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------- meshes


def icosphere(level: int):
    """Unit icosphere: 20*4**level faces.  Returns (V[nv,3] f64, F[nf,3] int32)."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    V = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(level):
        nv = V.shape[0]
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], 0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        uk, inv = np.unique(key, return_inverse=True)
        mid = V[uk // nv] + V[uk % nv]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        V = np.concatenate([V, mid], 0)
        nf = F.shape[0]
        a, b, c = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
        F = np.concatenate([np.stack([F[:, 0], a, c], 1), np.stack([F[:, 1], b, a], 1),
                            np.stack([F[:, 2], c, b], 1), np.stack([a, b, c], 1)], 0)
    return V, F.astype(np.int32)


def _morton2(ix, iy):
    def part(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFFFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v
    return part(ix) | (part(iy) << np.uint64(1))


def grid_mesh(nx: int, ny: int, torus: bool = True, morton: bool = True, height=None):
    """nx*ny quads split into 2 triangles each (2*nx*ny facets).

    torus=True wraps both directions (every vertex has valence 6).  Facets are
    Morton-ordered by quad when ``morton`` so that neighbouring facets are close in memory.
    Returns (V[nv,3] f64, F[nf,3] int32).
    """
    if torus:
        vx, vy = nx, ny
    else:
        vx, vy = nx + 1, ny + 1
    jj, ii = np.meshgrid(np.arange(vy), np.arange(vx), indexing="ij")
    if torus:
        R, r = 2.0, 0.7
        th = 2 * np.pi * ii / vx
        ph = 2 * np.pi * jj / vy
        V = np.stack([(R + r * np.cos(ph)) * np.cos(th), (R + r * np.cos(ph)) * np.sin(th),
                      r * np.sin(ph)], -1).reshape(-1, 3)
    else:
        X = ii / max(nx, 1)
        Y = jj / max(ny, 1)
        Z = np.zeros_like(X, dtype=np.float64) if height is None else height(X, Y)
        V = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float64)
    qj, qi = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    qi = qi.reshape(-1)
    qj = qj.reshape(-1)
    if morton:
        order = np.argsort(_morton2(qi, qj), kind="stable")
        qi, qj = qi[order], qj[order]
    i1 = (qi + 1) % vx if torus else qi + 1
    j1 = (qj + 1) % vy if torus else qj + 1
    v00 = qj * vx + qi
    v10 = qj * vx + i1
    v01 = j1 * vx + qi
    v11 = j1 * vx + i1
    F = np.empty((qi.size * 2, 3), np.int64)
    F[0::2] = np.stack([v00, v10, v11], 1)
    F[1::2] = np.stack([v00, v11, v01], 1)
    return V.astype(np.float64), F.astype(np.int32)


def add_vertex_noise(V, F, sigma_rel=0.3, seed=0):
    """Isotropic Gaussian vertex noise, sigma = sigma_rel * mean edge length."""
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], 0)
    mean_len = np.linalg.norm(V[e[:, 0]] - V[e[:, 1]], axis=1).mean()
    rs = np.random.RandomState(seed)
    return V + rs.normal(0.0, sigma_rel * mean_len, size=V.shape)


# ----------------------------------------------------------------------------- per-face features


def _normalize_once(a):
    n = np.sqrt((a * a).sum(1))[:, None] + 0.00000001
    return a * (1 / n)


def face_normals(V, F):
    """Unit face normals, two normalisation passes with +1e-8 (reference utils.py:63-68, 26-35)."""
    T = V[F]
    n = np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0])
    return _normalize_once(_normalize_once(n))


def face_barycenters(V, F, normalize=True):
    """Triangle centres of the vertices divided by the bbox diagonal (reference utils.py:1264-1294)."""
    if normalize:
        ext = V.max(axis=0) - V.min(axis=0)
        diag = math.sqrt(float(ext[0]) ** 2 + float(ext[1]) ** 2 + float(ext[2]) ** 2)
        V = V / diag
    return (V[F[:, 0]] + V[F[:, 1]] + V[F[:, 2]]) / 3


def face_features(V, F):
    """[normal(3) | barycentre(3)] rows, Cin = 6 (reference dataClasses.py:64)."""
    return np.concatenate([face_normals(V, F), face_barycenters(V, F)], axis=1)


# ----------------------------------------------------------------------------- adjacency


def _vertex_face_lists(F, nv):
    """CSR of faces around each vertex, faces in increasing order within a vertex."""
    nf = F.shape[0]
    vflat = F.reshape(-1).astype(np.int64)
    fflat = np.repeat(np.arange(nf, dtype=np.int64), 3)
    order = np.argsort(vflat * nf + fflat, kind="stable")
    vs, fs = vflat[order], fflat[order]
    cnt = np.bincount(vs, minlength=nv)
    ptr = np.concatenate([[0], np.cumsum(cnt)])
    return ptr, fs, cnt


def faces_large_adj(F, K):
    """Vertex-sharing facet adjacency in the reference's ``getFacesLargeAdj`` layout
    (utils.py:243-295): row f = [f+1, neighbours+1 ..., 0 padding]; edge-adjacent faces appear
    twice; appends happen vertex by vertex over ordered face pairs, and a row keeps only its
    first K-1 appends."""
    F = np.asarray(F)
    nf = F.shape[0]
    nv = int(F.max()) + 1
    ptr, fs, cnt = _vertex_face_lists(F, nv)
    tgt_chunks, val_chunks, key_chunks = [], [], []
    maxval = int(cnt.max())
    for d in np.unique(cnt):
        d = int(d)
        if d < 2:
            continue
        vsel = np.nonzero(cnt == d)[0]
        lists = fs[ptr[vsel][:, None] + np.arange(d)[None, :]]  # [nvd, d]
        a, b = np.triu_indices(d, 1)
        f1 = lists[:, a]
        f2 = lists[:, b]
        base = (vsel[:, None] * maxval + a[None, :]) * maxval + b[None, :]
        # event (v, a, b): first f1 <- f2, then f2 <- f1
        tgt_chunks += [f1.reshape(-1), f2.reshape(-1)]
        val_chunks += [f2.reshape(-1) + 1, f1.reshape(-1) + 1]
        key_chunks += [base.reshape(-1) * 2, base.reshape(-1) * 2 + 1]
    adj = np.zeros((nf, K), np.int32)
    adj[:, 0] = np.arange(1, nf + 1)
    if tgt_chunks:
        tgt = np.concatenate(tgt_chunks)
        val = np.concatenate(val_chunks)
        key = np.concatenate(key_chunks)
        order = np.lexsort((key, tgt))
        tgt, val = tgt[order], val[order]
        start = np.concatenate([[0], np.cumsum(np.bincount(tgt, minlength=nf))])[:-1]
        rank = np.arange(tgt.size) - start[tgt]
        keep = rank < K - 1
        adj[tgt[keep], rank[keep] + 1] = val[keep]
    return adj


def dedup_adj(adj):
    """Removes repeated neighbour ids inside each row (first occurrence kept, order preserved),
    re-padding with zeros -- the effect of the reference's sparse round trip on duplicates
    (SURVEY.md App. A.4 item 3), without its COO re-ordering."""
    adj = np.asarray(adj)
    N, K = adj.shape
    out = np.zeros_like(adj)
    # generic but simple: K is small
    for k in range(K):
        col = adj[:, k]
        dup = np.zeros(N, bool)
        for p in range(k):
            dup |= adj[:, p] == col
        keep = (~dup) & (col != 0)
        pos = (out != 0).sum(axis=1)
        out[np.nonzero(keep)[0], pos[keep]] = col[keep]
    return out


def edge_maps(F, max_edges=20):
    """(e_map[E,4], v_e_map[V,max_edges]) in the reference's ``getEdgeMap`` layout
    (utils.py:91-183): edges numbered by first appearance scanning faces, slots (v1v2, v1v3, v2v3);
    e_map row = (va, vb, first face, second face | -1); v_e_map lists edge ids per vertex in
    creation order, -1 padded."""
    F = np.asarray(F).astype(np.int64)
    nf = F.shape[0]
    nv = int(F.max()) + 1
    a = np.stack([F[:, 0], F[:, 0], F[:, 1]], 1).reshape(-1)
    b = np.stack([F[:, 1], F[:, 2], F[:, 2]], 1).reshape(-1)
    fid = np.repeat(np.arange(nf), 3)
    key = np.minimum(a, b) * nv + np.maximum(a, b)
    uk, first, inv = np.unique(key, return_index=True, return_inverse=True)
    eorder = np.argsort(first, kind="stable")  # unique-id -> rank by first appearance
    erank = np.empty_like(eorder)
    erank[eorder] = np.arange(eorder.size)
    eid = erank[inv]  # edge id of every half-edge
    E = uk.size
    e_map = np.full((E, 4), -1, np.int32)
    f_sorted = first[eorder]
    e_map[:, 0] = a[f_sorted]
    e_map[:, 1] = b[f_sorted]
    e_map[:, 2] = fid[f_sorted]
    later = np.ones(a.size, bool)
    later[first] = False
    # the reference overwrites column 3 with every later face, so the last one wins
    idx = np.nonzero(later)[0]
    e_map[eid[idx], 3] = fid[idx]
    v_e_map = np.full((nv, max_edges), -1, np.int32)
    vv = np.concatenate([e_map[:, 0], e_map[:, 1]]).astype(np.int64)
    ee = np.concatenate([np.arange(E), np.arange(E)])
    order = np.lexsort((ee, vv))
    vv, ee = vv[order], ee[order]
    start = np.concatenate([[0], np.cumsum(np.bincount(vv, minlength=nv))])[:-1]
    rank = np.arange(vv.size) - start[vv]
    if rank.max(initial=0) >= max_edges:
        raise ValueError("vertex with more than max_edges=%d edges" % max_edges)
    v_e_map[vv, rank] = ee
    return e_map, v_e_map


def vertex_faces(F, k_v=25, nv=0):
    """v_f[V,k_v]: faces around each vertex in increasing face order, -1 padded; rows of F equal to
    -1 (fake nodes) are skipped (reference utils.py:370-395, getVerticesFaces)."""
    F = np.asarray(F).astype(np.int64)
    if nv == 0:
        nv = int(F.max()) + 1
    real = F[:, 0] != -1
    fidx = np.nonzero(real)[0]
    vflat = F[real].reshape(-1)
    fflat = np.repeat(fidx, 3)
    order = np.lexsort((np.tile(np.arange(3), fidx.size), fflat, vflat))
    vs, fsort = vflat[order], fflat[order]
    start = np.concatenate([[0], np.cumsum(np.bincount(vs, minlength=nv))])[:-1]
    rank = np.arange(vs.size) - start[vs]
    if rank.max(initial=0) >= k_v:
        raise ValueError("vertex with more than k_v=%d faces" % k_v)
    v_f = np.full((nv, k_v), -1, np.int32)
    v_f[vs, rank] = fsort
    return v_f


# ----------------------------------------------------------------------------- pyramid


def pad_to_multiple(feat, adj, mult=16):
    """Appends fake nodes (zero features, self-only adjacency) so that N is a multiple of
    ``mult`` -- the state the reference's coarsening leaves fake nodes in."""
    N, K = adj.shape
    Np = (N + mult - 1) // mult * mult
    if Np == N:
        return feat, adj
    f2 = np.zeros((Np,) + feat.shape[1:], feat.dtype)
    f2[:N] = feat
    a2 = np.zeros((Np, K), adj.dtype)
    a2[:N] = adj
    a2[N:, 0] = np.arange(N + 1, Np + 1)
    return f2, a2


def coarsen_adj_by4(adj, K=None):
    """Adjacency of the graph whose node p is the group of fine nodes 4p..4p+3: the union of the
    children's neighbours mapped to parents, self first, duplicates removed, capped at K."""
    adj = np.asarray(adj).astype(np.int64)
    N, K0 = adj.shape
    K = K0 if K is None else K
    assert N % 4 == 0
    Nc = N // 4
    par = np.where(adj > 0, (adj - 1) // 4 + 1, 0).reshape(Nc, 4 * K0)
    self_id = np.arange(1, Nc + 1)[:, None]
    par = np.where(par == self_id, 0, par)
    par = np.sort(par, axis=1)
    dup = np.concatenate([np.zeros((Nc, 1), bool), par[:, 1:] == par[:, :-1]], axis=1)
    par = np.where(dup, 0, par)
    # stable compaction of the non-zero entries to the front
    order = np.argsort(par == 0, axis=1, kind="stable")
    par = np.take_along_axis(par, order, axis=1)
    out = np.zeros((Nc, K), np.int32)
    out[:, 0] = self_id[:, 0]
    w = min(K - 1, par.shape[1])
    out[:, 1:1 + w] = par[:, :w]
    # a group made only of fake nodes stays fake: self-only row already
    return out


def build_pyramid(adj0, levels=3, K=None):
    """[adj0, adj1, adj2]: binary-tree pyramid with 4 fine nodes per coarse node per level."""
    adjs = [np.asarray(adj0, np.int32)]
    for _ in range(levels - 1):
        adjs.append(coarsen_adj_by4(adjs[-1], K))
    return adjs
