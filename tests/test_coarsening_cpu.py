"""Graph pyramid of a patch (host half) against the reference's own functions: fixtures in
tests/golden/coarsen_cases.npz (oracle/make_golden.py:coarsen_cases) and the preprocessing outputs of the
reference driver already held by net_icosphere3.npz.  CPU only (the native pairing routine is host code)."""
import numpy as np
import pytest
import scipy.sparse

from conftest import golden
from facet_graph_convolution_b200 import _lib, coarsening as co, mesh


@pytest.fixture(scope="module", autouse=True)
def _library():
    from facet_graph_convolution_b200.build import build
    build()  # no-op when libfacetconv_b200.so is up to date; the pairing / growth routines live in it


def test_compute_perm_known_answer():
    # the reference's only known-answer test, Code/lib/coarsening.py:243-244
    assert co.compute_perm([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])]) == \
        [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]
    assert co.compute_perm([]) == []
    with pytest.raises(AssertionError):
        co.compute_perm([np.array([0, 0, 0])])


@pytest.mark.parametrize("tag", ["ico3", "open", "unit"])
def test_pyramid_equals_the_reference_functions(tag):
    g = golden("coarsen_cases")
    adj, feat, K = g[tag + "_adj"], g[tag + "_feat"], int(g[tag + "_K"])
    coo = co.list_to_sparse_w_normals(adj, feat[:, -3:], feat[:, :3])
    assert coo.data.dtype == np.float32
    assert np.array_equal(coo.row, g[tag + "_coo_row"]) and np.array_equal(coo.col, g[tag + "_coo_col"])
    assert np.array_equal(coo.data, g[tag + "_coo_val"])
    if tag != "unit":
        assert np.unique(coo.data).size > 1000  # the weights really vary in these cases
    graphs, perm = co.coarsen(coo, 4, rng=np.random.RandomState(int(g[tag + "_seed"])))
    assert np.array_equal(np.asarray(perm), g[tag + "_perm"])
    assert np.array_equal(co.inv_perm(perm), g[tag + "_inv_perm"])
    for l, G in enumerate(graphs):
        G = G.tocsr()
        G.sort_indices()
        assert np.array_equal(G.indptr, g[tag + "_g%d_indptr" % l]) and np.array_equal(G.indices, g[tag + "_g%d_indices" % l])
        assert np.array_equal(G.data, g[tag + "_g%d_data" % l])  # duplicate edges summed in the same order
    for l in range(3):
        lst, sat = co.sparse_to_list(graphs[2 * l], K)
        assert lst.dtype == np.int32 and np.array_equal(lst, g[tag + "_list%d" % l]) and sat == bool(g[tag + "_sat%d" % l])


def test_patch_pyramid_reproduces_the_reference_driver():
    """net_icosphere3.npz holds what `InferenceMesh.addMesh_TimeEfficient` (dataClasses.py) prepared with the
    global generator seeded to 0: level lists, permuted padded features, old-to-new permutation."""
    g = golden("net_icosphere3")
    K = g["adj0"].shape[2]
    adj = mesh.faces_large_adj(g["F"], K)
    np.random.seed(0)  # the default generator is the global one, like the reference
    adjs, x, new_to_old, old_to_new = co.patch_pyramid(adj, mesh.face_features(g["V"], g["F"]), K)
    for l in range(3):
        assert np.array_equal(adjs[l], g["adj%d" % l])
    assert np.array_equal(x[None].astype(np.float32), g["x"])
    assert np.array_equal(old_to_new, g["perm"]) and int(g["nreal"]) == adj.shape[0]
    assert np.array_equal(new_to_old[old_to_new[: adj.shape[0]]], np.arange(adj.shape[0]))


@pytest.mark.parametrize("precision", [32, 64])
def test_structure_at_size(precision):
    V, F = mesh.icosphere(5)
    V = mesh.add_vertex_noise(V, F, 0.3, seed=3)
    K = 16
    adj = mesh.faces_large_adj(F, K)
    feat = mesh.face_features(V, F).astype(np.float64)
    feat[:, 3:] *= 0.01
    coo = co.list_to_sparse_w_normals(adj, feat[:, -3:], feat[:, :3])
    rr, cc, vv = co._row_major_entries(coo)
    deg = np.array(coo.sum(axis=0) - coo.diagonal()).squeeze()
    rid = np.random.RandomState(0).permutation(coo.shape[0])
    cluster, score = co.greedy_pairing(rr, cc, vv, rid, deg, precision)
    sizes = np.bincount(cluster)
    assert sizes.min() >= 1 and sizes.max() <= 2 and cluster.max() + 1 == sizes.size and score > 0
    nbr = set(zip(rr.tolist(), cc.tolist()))
    order = np.argsort(cluster, kind="stable")
    pairs = order[np.repeat(np.cumsum(sizes) - sizes, 1)][sizes == 2], order[(np.cumsum(sizes) - 1)][sizes == 2]
    assert all((a, b) in nbr for a, b in zip(pairs[0].tolist(), pairs[1].tolist()))  # mates share an edge
    assert (sizes == 2).sum() > 0.8 * sizes.size  # a mesh graph pairs almost everything
    graphs, perm = co.coarsen(coo, 4, rng=np.random.RandomState(1), precision=precision)
    n = [G.shape[0] for G in graphs]
    assert all(n[i] == 2 * n[i + 1] for i in range(4)) and len(perm) == n[0] and sorted(perm) == list(range(n[0]))
    real = np.asarray(perm) < coo.shape[0]
    G0 = graphs[0]
    assert (G0 != G0.T).nnz == 0 and G0.diagonal().sum() == 0
    assert np.all(np.diff(G0.indptr)[~real] == 0)  # fake nodes are isolated
    lst, sat = co.sparse_to_list(G0, K)
    assert not sat and np.array_equal(lst[:, 0], np.arange(n[0]) + 1)
    # pooling by four consecutive nodes maps every fine edge into a coarse edge or inside one node
    G2 = graphs[2].tocoo()
    coarse = set(zip(G2.row.tolist(), G2.col.tolist()))
    c0 = G0.tocoo()
    assert all(a == b or (a, b) in coarse for a, b in zip((c0.row // 4).tolist(), (c0.col // 4).tolist()))


def test_sparse_to_list_saturation_and_native_argument_checks():
    A = scipy.sparse.csr_matrix(np.array([[5, 1, 1, 1], [1, 0, 0, 0], [1, 0, 0, 1], [1, 0, 1, 0]], np.float32))
    lst, sat = co.sparse_to_list(A, 3)
    assert sat and np.array_equal(lst, [[1, 2, 3], [2, 1, 0], [3, 1, 4], [4, 1, 3]])
    lst, sat = co.sparse_to_list(A, 4)
    assert not sat and np.array_equal(lst[0], [1, 2, 3, 4])
    r = np.array([0, 0, 1, 2], np.int32)
    c = np.array([1, 2, 0, 0], np.int32)
    v = np.ones(4, np.float32)
    w = np.ones(3, np.float32)
    cl, score = co.greedy_pairing(r, c, v, np.array([0, 1, 2]), w)
    assert cl.tolist() == [0, 0, 1] and score == 2.0  # first of two equal scores wins, node 2 stays single
    cl, _ = co.greedy_pairing(r, c, v, np.array([2, 1, 0]), w)
    assert cl.tolist() == [0, 1, 0]
    with pytest.raises(_lib.FacetConvError, match="sorted"):
        co.greedy_pairing(r[::-1], c, v, np.array([0, 1, 2]), w)
    with pytest.raises(_lib.FacetConvError, match="precision"):
        co.greedy_pairing(r, c, v, np.array([0, 1, 2]), w, precision=16)
    with pytest.raises(_lib.FacetConvError, match="order"):
        co.greedy_pairing(r, c, v, np.array([0, 1, 7]), w)
    with pytest.raises(_lib.FacetConvError):
        co.greedy_pairing(r[:0], c[:0], v[:0], np.array([0]), w)


def test_patch_growth_equals_the_reference_function():
    """tests/golden/patch_cases.npz: `getGraphPatch_wMask` (utils.py:1508-1696) called six times on a noisy
    icosphere-4 while the ownership mask fills up (context nodes, the minimum-size second phase, next seeds)."""
    g = golden("patch_cases")
    adj = mesh.faces_large_adj(g["F"], int(g["K"]))
    for t in range(6):
        nn, seed, mp, nxt = (int(v) for v in g["g%d_args" % t])
        a, old, s = co.get_graph_patch_w_mask(adj, nn, seed, g["g%d_mask" % t].astype(np.float64), mp)
        assert np.array_equal(a, g["g%d_adj" % t]) and np.array_equal(old, g["g%d_old" % t]) and s == nxt, t
    # structure: local ids are a bijection onto the patch, column 0 is the node, lists stay inside the patch
    a, old, _ = co.get_graph_patch_w_mask(adj, 800, 17, np.zeros(adj.shape[0]), 100)
    assert np.unique(old).size == old.size and np.array_equal(a[:, 0], np.arange(a.shape[0]) + 1)
    assert a.min() >= 0 and a.max() <= a.shape[0]
    inside = set(old.tolist())
    for i in (0, 5, a.shape[0] - 1):
        want = [j - 1 for j in adj[old[i], 1:] if j > 0 and j - 1 in inside]
        got = [int(old[j - 1]) for j in a[i, 1:] if j > 0]
        assert got == want
    with pytest.raises(_lib.FacetConvError, match="seed"):
        co.get_graph_patch_w_mask(adj, 100, adj.shape[0], np.zeros(adj.shape[0]), 50)
    with pytest.raises(_lib.FacetConvError, match="mask"):
        co.get_graph_patch_w_mask(adj, 100, 0, np.zeros(3), 50)
    bad = adj.copy()
    bad[3, 2] = adj.shape[0] + 5
    with pytest.raises(_lib.FacetConvError, match="outside"):
        co.get_graph_patch_w_mask(bad, 100, 3, np.zeros(adj.shape[0]), 50)


def test_patch_loop_equals_the_reference_driver():
    """The whole preprocessing of a mesh above the size limit (dataClasses.py:69-150) with the generator seeded
    like the reference run: same patches, same pyramids, same permutations."""
    g = golden("patch_cases")
    K = int(g["K"])
    patch_size, min_size, seed, count = (int(v) for v in g["drv_args"])
    adj = mesh.faces_large_adj(g["F"], K)
    feat = mesh.face_features(g["V"], g["F"])
    ps = co.extract_patches(adj, feat, patch_size, K, min_patch_size=min_size, rng=np.random.RandomState(seed))
    assert len(ps) == count
    covered = np.zeros(adj.shape[0], dtype=bool)
    for i, p in enumerate(ps):
        assert np.array_equal(p.x, g["drv%d_x" % i]) and np.array_equal(p.face_ids, g["drv%d_ids" % i])
        assert np.array_equal(p.perm, g["drv%d_perm" % i])
        for l in range(3):
            assert p.adjs[l].dtype == np.int32 and np.array_equal(p.adjs[l], g["drv%d_adj%d" % (i, l)])
        covered[p.face_ids] = True
    assert covered.all()


def test_mesh_with_vertices_reproduces_the_reference_driver():
    """net_ms_icosphere2.npz holds what `PreprocessedData.addMeshWithVertices` (small-mesh branch,
    dataClasses.py:236-270, 377-443) prepared with the global generator seeded to 1: normalised vertices,
    permuted features, level lists, the face list with (-1,-1,-1) rows for fake nodes, vertex -> faces lists."""
    g = golden("net_ms_icosphere2")
    K = g["adj0"].shape[2]
    d = co.mesh_with_vertices(g["V"], g["F"], K, rng=np.random.RandomState(1))
    for l in range(3):
        assert np.array_equal(d["adjs"][l], g["adj%d" % l])
    assert np.array_equal(d["x"][None].astype(np.float32), g["x"])
    assert np.array_equal(d["faces"][None], g["faces"]) and np.array_equal(d["v_faces"][None], g["v_faces"])
    assert np.array_equal(d["verts"][None].astype(np.float32), g["verts_in"])
    assert d["num_faces"] == g["F"].shape[0] and (d["faces"][d["new_to_old"] >= d["num_faces"]] == -1).all()
    assert np.array_equal(d["new_to_old"][d["old_to_new"][: d["num_faces"]]], np.arange(d["num_faces"]))
    # the adjacency may come from the GPU builder instead: same result
    d2 = co.mesh_with_vertices(g["V"], g["F"], K, rng=np.random.RandomState(1), f_adj=mesh.faces_large_adj(g["F"], K))
    assert np.array_equal(d2["x"], d["x"]) and np.array_equal(d2["v_faces"], d["v_faces"])


def test_mesh_patches_with_vertices_equal_the_reference():
    """tests/golden/patch_vertex_cases.npz: `getMeshPatch` (utils.py:1298-1415) in full for three seeds, and the
    patch loop of `addMeshWithVertices` (dataClasses.py:274-376) by shape, dtype and SHA-1 of every output."""
    import hashlib
    g = golden("patch_vertex_cases")
    V, F = mesh.icosphere(4)
    V = mesh.add_vertex_noise(V, F, 0.3, seed=2)
    K = int(g["K"])
    adj = mesh.faces_large_adj(F, K)
    for i in range(3):
        v, f, a, vold, fold = co.get_mesh_patch(V.astype(np.float32), F, adj, 400, int(g["mp%d_seed" % i]))
        assert v.dtype == np.float32 and np.array_equal(v, g["mp%d_v" % i]) and np.array_equal(f, g["mp%d_f" % i])
        assert np.array_equal(a, g["mp%d_adj" % i]) and np.array_equal(vold, g["mp%d_vold" % i])
        assert np.array_equal(fold, g["mp%d_fold" % i])
        assert np.array_equal(V.astype(np.float32)[vold][f], V.astype(np.float32)[F[fold]])  # same triangles
    patch_size, seed, count = (int(t) for t in g["drv_args"])
    ps = co.extract_patches_with_vertices(V, F, patch_size, K, rng=np.random.RandomState(seed))
    assert len(ps) == count
    for i, p in enumerate(ps):
        arrs = dict(x=p["x"], adj0=p["adjs"][0][0], adj1=p["adjs"][1][0], adj2=p["adjs"][2][0], faces=p["faces"],
                    v_faces=p["v_faces"], verts=p["verts"], face_ids=p["face_ids"], vertex_ids=p["vertex_ids"],
                    old_to_new=p["old_to_new"])
        for k, a in arrs.items():
            want_dtype = np.dtype(bytes(g["drv%d_%s_dtype" % (i, k)]).decode())
            a = np.ascontiguousarray(np.asarray(a).astype(want_dtype, copy=False))
            assert tuple(a.shape) == tuple(g["drv%d_%s_shape" % (i, k)]), (i, k)
            assert hashlib.sha1(a.tobytes()).digest() == bytes(g["drv%d_%s_sha1" % (i, k)]), (i, k)


def test_patch_pyramid_gives_up_when_k_is_too_small():
    """The reference coarsens again for as long as a level has more neighbours than K - 1 columns
    (dataClasses.py:116-129) -- forever when K is simply too small; here that is an error after 20 tries."""
    V, F = mesh.icosphere(2)
    adj = mesh.faces_large_adj(F, 12)
    with pytest.raises(_lib.FacetConvError, match="saturates"):
        co.patch_pyramid(adj, mesh.face_features(V, F), 3, rng=np.random.RandomState(0))
