"""Graph pyramid of a facet patch: weighted facet graph, heavy-edge coarsening, binary-tree node order and
the K-column adjacency lists of every level (host half).

Replaces, for the way `dataClasses.py:108-135` drives them, reference `utils.listToSparseWNormals`
(Code/utils.py:1753-1797), `lib/coarsening.coarsen` / `metis` / `compute_perm` / `perm_adjacency`
(Code/lib/coarsening.py:5-32, 35-131, 196-243, 269-296), `utils.sparseToList` (:1799-1828) and `utils.inv_perm`
(:1830-1836).  Host side: the pairing of a level visits nodes one after the other and every choice depends
on the earlier ones, so it is the native sequential routine `fgc_greedy_pairing` (O(nnz)); everything else is
vectorised NumPy / SciPy instead of per-node Python loops (the reference's `compute_perm` scans the whole
parent vector once per node, quadratic in the patch size).  The sparse assembly goes through the same SciPy
calls as the reference so that duplicate edges are summed in the same float32 order.

The reference draws its visiting orders from the global NumPy generator (coarsening.py:57, 96); `rng` here
defaults to that generator and is consumed in the same sequence, so a seeded run gives the reference's
pyramid exactly (tests/golden/coarsen_*.npz, produced by the reference functions in this container).  Ties
between equal edge scores are resolved by the column order of the sorted adjacency; this module fixes that
order (row-major, columns ascending), which is what the reference's unstable `np.argsort` yields in this
image -- on another NumPy build the reference itself may break ties differently.

`precision`: the reference scores an edge as `vv * (1.0/weights[tid] + 1.0/weights[nid])` on float32 arrays.
Under NumPy >= 2 (this image; the fixtures) Python floats are weak scalars and the whole expression, and the
running total, stay float32 -- `precision=32`, the default and the pinned behaviour.  Under the NumPy 1.x the
reference was written for, the same source promotes to float64 -- `precision=64` evaluates that; it has no
reference run to compare with here (only the structural test), so treat it as unpinned.
"""
import ctypes as C

import numpy as np
import scipy.sparse

from . import _lib

__all__ = ["list_to_sparse_w_normals", "coarsen", "compute_perm", "perm_adjacency", "sparse_to_list", "inv_perm",
           "greedy_pairing", "patch_pyramid", "get_graph_patch_w_mask", "extract_patches",
           "mesh_with_vertices", "get_mesh_patch", "extract_patches_with_vertices"]


def list_to_sparse_w_normals(adj, nodes_pos, nodes_normals):
    """Weighted facet graph of a 1-based, 0-padded adjacency list whose first column is the node itself:
    w(n, j) = max(<n_n, n_j> * exp(-|p_j - p_n|^2 / (2 * 0.001^2)), 0.001) as float32, a row ending at its
    first 0 (utils.py:1753-1797)."""
    adj = np.asarray(adj)
    N, K = adj.shape
    nb = adj[:, 1:].astype(np.int64) - 1
    keep = np.logical_and.accumulate(nb >= 0, axis=1)
    rows = np.broadcast_to(np.arange(N, dtype=np.int64)[:, None], nb.shape)[keep]
    cols = nb[keep]
    pos = np.asarray(nodes_pos)
    nrm = np.asarray(nodes_normals)
    dp = (nrm[rows] * nrm[cols]).sum(axis=-1)
    d = pos[cols] - pos[rows]
    dist = np.sqrt((d * d).sum(axis=-1))
    vals = np.maximum(dp * np.exp(-(dist ** 2) * (1.0 / (2 * 0.001 * 0.001))), 0.001).astype(np.float32)
    return scipy.sparse.coo_matrix((vals, (rows.astype(np.int32), cols.astype(np.int32))), shape=(N, N))


def greedy_pairing(rr, cc, vv, rid, weights, precision=32):
    """One level of pairing (coarsening.py:135-194) on row-sorted entries: (cluster_id[N], total score)."""
    rr = np.ascontiguousarray(rr, dtype=np.int32)
    cc = np.ascontiguousarray(cc, dtype=np.int32)
    vv = np.ascontiguousarray(vv, dtype=np.float32)
    rid = np.ascontiguousarray(rid, dtype=np.int64)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    if rr.size == 0:
        raise _lib.FacetConvError("greedy_pairing: graph without edges")
    n = int(rr[-1]) + 1
    if weights.size < n:
        raise _lib.FacetConvError("greedy_pairing: %d weights for %d rows" % (weights.size, n))
    cluster = np.empty(n, dtype=np.int32)
    total = C.c_double(0.0)
    ncl = C.c_int32(0)
    _lib.check(_lib.lib().fgc_greedy_pairing(
        rr.ctypes.data, cc.ctypes.data, vv.ctypes.data, rr.size, rid.ctypes.data, rid.size, weights.ctypes.data, n,
        int(precision), cluster.ctypes.data, C.byref(total), C.byref(ncl)), "fgc_greedy_pairing")
    return cluster, total.value


def _row_major_entries(W):
    """Non-zero entries of W (duplicates summed) ordered by row, then column."""
    if scipy.sparse.issparse(W) and W.format == "csr" and W.has_canonical_format:
        r = np.repeat(np.arange(W.shape[0], dtype=np.int32), np.diff(W.indptr))
        nz = W.data != 0
        return r[nz], W.indices[nz], W.data[nz]
    r, c, v = scipy.sparse.find(W)  # sums duplicates the way the reference's call does
    if r.size > 1:
        dr = np.diff(r)
        if np.all((dr > 0) | ((dr == 0) & (np.diff(c) > 0))):
            return r, c, v
    order = np.lexsort((c, r))
    return r[order], c[order], v[order]


def _pair_levels(W, levels, rng, precision, trials=3):
    N = W.shape[0]
    rid = rng.permutation(range(N))
    degree = W.sum(axis=0) - W.diagonal()
    graphs, parents = [W], []
    for _ in range(levels):
        weights = np.array(degree).squeeze()
        rr, cc, vv = _row_major_entries(W)
        best, best_score = None, 0.0
        for _t in range(trials):
            cluster, score = greedy_pairing(rr, cc, vv, rid, weights, precision)
            if score > best_score:
                best, best_score = cluster, score
            rid = rng.permutation(range(N))  # drawn after every trial, the last one unused (as the reference)
        if best is None:
            raise _lib.FacetConvError("coarsen: no edge with a positive score")
        parents.append(best)
        n_new = int(best.max()) + 1
        W = scipy.sparse.csr_matrix((vv, (best[rr], best[cc])), shape=(n_new, n_new))
        W.eliminate_zeros()
        graphs.append(W)
        N = n_new
        degree = W.sum(axis=0)  # self loops of merged pairs now count
        rid = np.argsort(np.array(degree).squeeze())
    return graphs, parents


def compute_perm(parents):
    """Node orders, finest level first, in which nodes 2i and 2i+1 of a level are the children of node i of
    the next one; singletons get a fake sibling, fake parents two fake children, fakes numbered after the real
    nodes in order of appearance (coarsening.py:196-240; known answer at :243-244).  Linear time."""
    if len(parents) == 0:
        return []
    layers = [np.arange(int(np.max(parents[-1])) + 1, dtype=np.int64)]
    for parent in parents[::-1]:
        parent = np.asarray(parent, dtype=np.int64)
        above = layers[-1]
        n_above = int(above.max()) + 1 if above.size else 0
        counts = np.bincount(parent, minlength=n_above)
        if counts.size > n_above or counts.max(initial=0) > 2:
            raise AssertionError("compute_perm: a cluster has more than two members or an unknown parent")
        by_parent = np.argsort(parent, kind="stable")
        first = np.concatenate([[0], np.cumsum(counts)[:-1]])
        cnt = counts[above]                                   # fake parents index past the real ones: count 0
        fakes_before = np.concatenate([[0], np.cumsum(2 - cnt)[:-1]]) + parent.size
        safe = np.minimum(first[above], max(parent.size - 1, 0))
        c0 = np.where(cnt >= 1, by_parent[safe], fakes_before)
        c1 = np.where(cnt == 2, by_parent[np.minimum(safe + 1, max(parent.size - 1, 0))],
                      fakes_before + (cnt == 0))
        layers.append(np.stack([c0, c1], axis=1).reshape(-1))
    return [l.tolist() for l in layers[::-1]]


def perm_adjacency(A, indices):
    """A with Mnew - M isolated nodes appended and node j moved to the position of j in `indices`
    (coarsening.py:269-296)."""
    if indices is None:
        return A
    M = A.shape[0]
    Mnew = len(indices)
    assert Mnew >= M
    A = A.tocoo()
    where = np.argsort(indices)
    return scipy.sparse.coo_matrix((A.data, (where[A.row], where[A.col])), shape=(Mnew, Mnew))


def coarsen(A, levels, self_connections=False, rng=None, precision=32):
    """`levels` rounds of pairing; returns (graphs as CSR with every level but the last in binary-tree order
    and padded with fake nodes, order of the finest level = new-to-old) -- coarsening.py:5-32."""
    rng = np.random if rng is None else rng
    graphs, parents = _pair_levels(A, levels, rng, precision)
    perms = compute_perm(parents)
    out = []
    for i, G in enumerate(graphs):
        G = G.tocoo(copy=True)  # (the reference zeroes the diagonal of the caller's matrix in place)
        if not self_connections:
            G.setdiag(0)
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = G.tocsr()
        G.eliminate_zeros()
        out.append(G)
    return out, (perms[0] if levels > 0 else None)


def sparse_to_list(A, K):
    """[N, K] 1-based adjacency list, column 0 the node itself, then the off-diagonal entries of the row in
    storage order, 0-padded; (list, saturated) with saturated = some row had more than K - 1 (utils.py:1799-1828)."""
    N = A.shape[0]
    cx = A.tocoo()
    off = cx.row != cx.col
    r, c = cx.row[off].astype(np.int64), cx.col[off].astype(np.int64)
    if r.size and np.any(np.diff(r) < 0):  # storage order is row-major for the CSR input; be safe otherwise
        o = np.argsort(r, kind="stable")
        r, c = r[o], c[o]
    counts = np.bincount(r, minlength=N)
    slot = np.arange(r.size) - np.repeat(np.concatenate([[0], np.cumsum(counts)[:-1]]), counts) + 1
    out = np.zeros((N, K), dtype=np.int32)
    out[:, 0] = np.arange(N) + 1
    ok = slot < K
    out[r[ok], slot[ok]] = c[ok] + 1
    return out, bool((~ok).any())


def inv_perm(perm):
    """inverse[p] = position of p in `perm` (utils.py:1830-1836)."""
    perm = np.asarray(perm, dtype=np.int64)
    inverse = np.zeros(max(perm.size, int(perm.max()) + 1 if perm.size else 0), dtype=np.int64)
    inverse[perm] = np.arange(perm.size)
    return inverse


def patch_pyramid(adj, features, K, level_num=3, step_num=2, rng=None, precision=32, max_retries=20):
    """The network inputs of one patch as `dataClasses.py:108-150` prepares them: the graph is coarsened
    (level_num - 1) * step_num times (again while some level saturates its K columns), the K-column lists
    of levels 0, step_num, 2 * step_num, ... are extracted and the feature rows are padded with zero rows for
    the fake nodes and put in tree order.  `features` = [N, 6] normals | positions.
    Returns (adjs [1, N_l, K] per level, features [N_0', C], new_to_old, old_to_new)."""
    features = np.asarray(features)
    coo = list_to_sparse_w_normals(adj, features[:, -3:], features[:, :3])
    for _ in range(max_retries):
        graphs, new_to_old = coarsen(coo, (level_num - 1) * step_num, rng=rng, precision=precision)
        adjs, saturated = [], False
        for lvl in range(level_num):
            a, sat = sparse_to_list(graphs[step_num * lvl], K)
            adjs.append(a[np.newaxis])
            saturated = saturated or sat
        if not saturated:
            break
    else:
        raise _lib.FacetConvError("patch_pyramid: a level still saturates K = %d after %d coarsenings" % (K, max_retries))
    new_to_old = np.asarray(new_to_old, dtype=np.int64)
    padded = np.concatenate([features, np.zeros((new_to_old.size - features.shape[0], features.shape[1]))], axis=0)
    return adjs, padded[new_to_old], new_to_old, inv_perm(new_to_old)


def get_graph_patch_w_mask(f_adj, nodes_num, seed, mask, min_patch_size):
    """One patch grown breadth-first from `seed` (utils.py:1508-1696): (patch adjacency [n_p, K] 1-based in
    patch-local ids, original id of every patch node, next seed or -1).  Native host routine `fgc_grow_patch`."""
    f_adj = np.ascontiguousarray(f_adj, dtype=np.int32)
    n, K = f_adj.shape
    taken = np.ascontiguousarray(np.asarray(mask) == 1, dtype=np.uint8)
    if taken.size != n:
        raise _lib.FacetConvError("get_graph_patch_w_mask: mask has %d entries for %d nodes" % (taken.size, n))
    cap = max(int(nodes_num), int(min_patch_size)) + K
    out = np.empty((cap, K), dtype=np.int32)
    old = np.empty(cap, dtype=np.int64)
    count, nxt = C.c_int64(0), C.c_int64(-1)
    _lib.check(_lib.lib().fgc_grow_patch(f_adj.ctypes.data, n, K, int(nodes_num), int(seed), taken.ctypes.data,
                                         int(min_patch_size), out.ctypes.data, cap, old.ctypes.data,
                                         C.byref(count), C.byref(nxt)), "fgc_grow_patch")
    return out[: count.value].astype(np.int64), old[: count.value].copy(), int(nxt.value)


def extract_patches(f_adj, features, patch_size, K, min_patch_size=2000, level_num=3, step_num=2, rng=None,
                    precision=32, min_component=100):
    """The patch loop of the reference's preprocessing (dataClasses.py:69-150) for a mesh above the size limit:
    seeds are drawn among the facets no patch owns yet (or taken from the previous patch's border), every
    patch is grown to `patch_size` (at least `min_patch_size` with context), components under 100 facets are
    dropped, and each patch gets its pyramid.  Returns `patches.Patch` objects (x, level lists, global facet
    ids, old-to-new permutation).  `rng` as in `coarsen`: the global NumPy generator by default, consumed in
    the reference's order (one `randint` per fresh seed, then the coarsening draws)."""
    from .patches import Patch
    rng = np.random if rng is None else rng
    f_adj = np.asarray(f_adj)
    features = np.asarray(features)
    n = f_adj.shape[0]
    owned = np.zeros(n)
    out = []
    next_seed = -1
    while np.any(owned == 0):
        if next_seed == -1:
            free = np.flatnonzero(owned == 0)
            seed = int(free[rng.randint(free.shape[0])])
        else:
            seed = next_seed
            if owned[seed] == 1:
                raise _lib.FacetConvError("extract_patches: the border seed %d is already owned" % seed)
        p_adj, old, next_seed = get_graph_patch_w_mask(f_adj, patch_size, seed, owned, min_patch_size)
        owned[old] = 1
        if old.shape[0] < min_component:
            continue
        feats = features[old]
        if level_num > 1:
            adjs, x, _, old_to_new = patch_pyramid(p_adj, feats, K, level_num, step_num, rng, precision)
            adjs = [a[0].astype(np.int32) for a in adjs]
        else:
            adjs, x, old_to_new = [p_adj.astype(np.int32)], feats, None
        out.append(Patch(x=x.astype(np.float32), adjs=adjs, face_ids=old.astype(np.int64),
                         perm=None if old_to_new is None else old_to_new.astype(np.int32)))
    return out


def mesh_with_vertices(V, F, K, level_num=3, step_num=2, rng=None, precision=32, kv=25, f_adj=None):
    """Inputs of the multi-scale network *with vertex update* for a mesh below the size limit, as the small-mesh
    branch of `PreprocessedData.addMeshWithVertices` prepares them (dataClasses.py:236-270, 377-443): vertices
    divided by their bounding-box diagonal (`normalizePointSets`, utils.py:2077-2107), facet features and
    pyramid in tree order, the face list padded with (-1,-1,-1) rows for the fake nodes and permuted the same
    way, and the vertex -> faces lists of that permuted list (`getVerticesFaces(faces, 25, V)`).
    `f_adj` may carry the adjacency already built on the GPU (`ops.build_faces_adj`).
    Returns dict(x [N0', 6], adjs [1, N_l, K] per level, faces [N0', 3], v_faces [V, kv], verts [V, 3],
    new_to_old, old_to_new, num_faces)."""
    from . import mesh
    V = np.asarray(V)
    F = np.asarray(F)
    fn = mesh.face_normals(V, F)
    adj = mesh.faces_large_adj(F, K) if f_adj is None else np.asarray(f_adj)
    pos = mesh.face_barycenters(V, F, normalize=True)
    feats = np.concatenate((fn, pos), axis=1)
    span = V.max(axis=0) - V.min(axis=0)
    verts = V / np.sqrt((span.astype(np.float64) ** 2).sum())
    coo = list_to_sparse_w_normals(adj, pos, fn)
    graphs, new_to_old = coarsen(coo, (level_num - 1) * step_num, rng=rng, precision=precision)
    new_to_old = np.asarray(new_to_old, dtype=np.int64)
    extra = new_to_old.size - F.shape[0]
    faces = np.concatenate((F.astype(np.int64), -np.ones((extra, 3), np.int64)), axis=0)[new_to_old]
    x = np.concatenate((feats, np.zeros((extra, feats.shape[1]))), axis=0)[new_to_old]
    adjs = [sparse_to_list(graphs[step_num * lvl], K)[0][np.newaxis] for lvl in range(level_num)]
    return dict(x=x, adjs=adjs, faces=faces, v_faces=mesh.vertex_faces(faces, kv, V.shape[0]), verts=verts,
                new_to_old=new_to_old, old_to_new=inv_perm(new_to_old), num_faces=F.shape[0])


def get_mesh_patch(v_in, f_in, f_adj_in, face_num, seed):
    """A patch of a triangle mesh grown breadth-first over the facet graph (utils.py:1298-1415): faces are
    numbered in discovery order exactly as `get_graph_patch_w_mask` does without a mask, vertices in order of
    first use by those faces.  Returns (vertices float32, faces in patch-local vertex ids, patch adjacency,
    original vertex ids, original face ids)."""
    v_in = np.asarray(v_in)
    f_in = np.asarray(f_in).astype(np.int64)
    adj_p, f_old, _ = get_graph_patch_w_mask(f_adj_in, face_num, seed, np.zeros(f_in.shape[0]), 0)
    corners = f_in[f_old].reshape(-1)
    uniq, first = np.unique(corners, return_index=True)
    v_old = uniq[np.argsort(first, kind="stable")]
    local = np.full(v_in.shape[0], -1, dtype=np.int64)
    local[v_old] = np.arange(v_old.size)
    return v_in[v_old].astype(np.float32), local[f_in[f_old]], adj_p, v_old, f_old


def extract_patches_with_vertices(V, F, patch_size, K, level_num=3, step_num=2, rng=None, precision=32, kv=25,
                                  f_adj=None, min_component=100):
    """The patch loop of `PreprocessedData.addMeshWithVertices` for a mesh above the size limit
    (dataClasses.py:236-376, inference case): every patch starts at a random face no patch covers yet and grows
    to `patch_size` faces whatever the earlier patches took; per patch the vertices (of the mesh divided by its
    bounding-box diagonal), the features / pyramid in tree order, the permuted face list with (-1,-1,-1)
    rows and its vertex -> faces lists.  Returns a list of dicts like `mesh_with_vertices` plus
    `face_ids` / `vertex_ids` (original ids in patch order)."""
    from . import mesh
    rng = np.random if rng is None else rng
    V = np.asarray(V)
    F = np.asarray(F)
    adj = mesh.faces_large_adj(F, K) if f_adj is None else np.asarray(f_adj)
    feats = np.concatenate((mesh.face_normals(V, F), mesh.face_barycenters(V, F, normalize=True)), axis=1)
    span = V.max(axis=0) - V.min(axis=0)
    Vn = V / np.sqrt((span.astype(np.float64) ** 2).sum())
    covered = np.zeros(F.shape[0])
    out = []
    while np.any(covered == 0):
        free = np.flatnonzero(covered == 0)
        seed = int(free[rng.randint(free.shape[0])])
        pv, pf, p_adj, v_old, f_old = get_mesh_patch(Vn, F, adj, patch_size, seed)
        covered[f_old] += 1
        if f_old.shape[0] < min_component:
            continue
        x0 = feats[f_old]
        graphs, new_to_old = coarsen(list_to_sparse_w_normals(p_adj, x0[:, -3:], x0[:, :3]),
                                     (level_num - 1) * step_num, rng=rng, precision=precision)
        new_to_old = np.asarray(new_to_old, dtype=np.int64)
        extra = new_to_old.size - f_old.shape[0]
        faces = np.concatenate((pf, -np.ones((extra, 3), np.int64)), axis=0)[new_to_old]
        x = np.concatenate((x0, np.zeros((extra, x0.shape[1]))), axis=0)[new_to_old]
        adjs = [sparse_to_list(graphs[step_num * lvl], K)[0][np.newaxis] for lvl in range(level_num)]
        out.append(dict(x=x, adjs=adjs, faces=faces, v_faces=mesh.vertex_faces(faces, kv, pv.shape[0]), verts=pv,
                        new_to_old=new_to_old, old_to_new=inv_perm(new_to_old), num_faces=f_old.shape[0],
                        face_ids=f_old, vertex_ids=v_old))
    return out
