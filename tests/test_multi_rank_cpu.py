"""N>1 host logic on CPU: gloo, world_size 2.  No CUDA kernels run here -- the
per-patch forward is a stand-in; what is tested is partitioning, the ragged gather, the
reference's overlap merge, the flat gradient bucket and its single all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facet_graph_convolution_b200 import patches as P
from facet_graph_convolution_b200 import train as T


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_forward(p):
    """Deterministic stand-in for the network: a function of the patch's own rows only."""
    x = p.x.astype(np.float64)
    n = x[:, :3] + 0.25 * np.roll(x[:, 3:6], 1, axis=1) + 0.01
    return n[None].astype(np.float32)


def _make_patches():
    rs = np.random.RandomState(3)
    num_faces = 500
    out = []
    for i, (lo, hi) in enumerate([(0, 200), (150, 360), (300, 500), (0, 60), (440, 500)]):
        ids = np.arange(lo, hi, dtype=np.int64)
        n0 = (len(ids) + 15) // 16 * 16
        x = np.zeros((n0, 6), np.float32)
        perm = rs.permutation(n0).astype(np.int32)           # oldToNew
        real = rs.randn(len(ids), 6).astype(np.float32)
        xo = np.zeros((n0, 6), np.float32)
        xo[: len(ids)] = real
        x[perm] = xo                                         # x[new] = padded[old]
        out.append(P.Patch(x=x, adjs=[], face_ids=ids, perm=perm, core=None))
    return out, num_faces


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        patches, nf = _make_patches()
        pred = P.infer_sharded(patches, nf, _fake_forward, rank, world)
        # gradient bucket: rank r contributes (r+1) * base
        torch.manual_seed(0)
        params = [torch.randn(4, 3), torch.randn(7), torch.randn(2, 2, 2)]
        for p in params:
            p.requires_grad_(True)
        b = T.GradBucket(params)
        opt = T.Adam(b)
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        b.pack()
        b.all_reduce_mean()
        opt.step()
        q.put((rank, pred, b.flat.clone().numpy(), [p.detach().clone().numpy() for p in params]))
    finally:
        dist.destroy_process_group()


def test_partition_is_deterministic_and_balanced():
    costs = [24512, 2016, 20000, 20000, 8192, 8192, 8192, 512]
    plan = P.partition(costs, 4)
    assert sorted(i for r in plan for i in r) == list(range(len(costs)))
    loads = [sum(costs[i] for i in r) for r in plan]
    assert max(loads) <= 1.35 * (sum(costs) / 4)
    assert plan == P.partition(list(costs), 4)
    assert P.partition([5, 5], 1) == [[0, 1]]
    assert P.partition([], 2) == [[], []]


def test_world2_inference_matches_single_process_and_reference_merge():
    patches, nf = _make_patches()
    single = P.infer_sharded(patches, nf, _fake_forward, 0, 1)
    # the reference's own merge (train.py:117-126,136), restated naively
    acc = np.zeros((nf, 3))
    for p in patches:
        out = _fake_forward(p)[0][p.perm][: p.num_real]
        acc[p.face_ids] += out
    for _ in range(2):
        acc = acc * (1 / (np.sqrt((acc * acc).sum(1))[:, None] + 1e-8))
    assert np.abs(single - acc).max() < 1e-12
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    for rank, pred, flat, params in res:
        assert np.abs(pred - single).max() < 1e-6
    # one all-reduce: mean of (1,2) * (i+1) = 1.5 * (i+1), identical on both ranks
    f0, f1 = res[0][2], res[1][2]
    assert np.array_equal(f0, f1)
    assert np.allclose(f0[:12], 1.5) and np.allclose(f0[12:19], 3.0) and np.allclose(f0[19:], 4.5)
    for a, b in zip(res[0][3], res[1][3]):
        assert np.array_equal(a, b)          # replicas stay bit-identical after the step
    # first Adam step moves every parameter by lr (TF formulation)
    torch.manual_seed(0)
    ref = [torch.randn(4, 3), torch.randn(7), torch.randn(2, 2, 2)]
    for a, r in zip(res[0][3], ref):
        assert np.allclose(a, r.numpy() - 1e-3, atol=1e-6)


def test_grid_patches_cover_every_facet_once_with_halo():
    pts, nf = P.grid_patches(20, 12, block=8, halo=3, K=16)
    assert nf == 480 and len(pts) == 3 * 2
    seen = np.zeros(nf, int)
    for p in pts:
        assert p.x.shape[0] % 16 == 0 and p.x.shape[1] == 6
        assert [a.shape[0] for a in p.adjs] == [p.x.shape[0], p.x.shape[0] // 4, p.x.shape[0] // 16]
        a0 = p.adjs[0]
        assert np.array_equal(a0[:, 0], np.arange(1, a0.shape[0] + 1))
        assert a0.min() >= 0 and a0.max() <= a0.shape[0]
        assert p.core.sum() < p.num_real          # a halo exists
        np.add.at(seen, p.face_ids[p.core], 1)
    assert np.array_equal(seen, np.ones(nf, int))
    # generating only a rank's share gives the same patches
    some, _ = P.grid_patches(20, 12, block=8, halo=3, K=16, only=[4])
    assert np.array_equal(some[0].x, pts[4].x) and np.array_equal(some[0].adjs[2], pts[4].adjs[2])


def test_batched_patches_pad_with_fake_nodes_only():
    """batch_patches: ragged patches stacked into one [B,n0,...] batch; the padding is fake nodes (zero
    features, self-only adjacency at every level), real rows are untouched and never reference padding."""
    pts, _ = P.grid_patches(20, 12, block=8, halo=3, K=16)
    idx = [0, 2, 5]
    xb, ab = P.batch_patches(pts, idx)
    n0 = xb.shape[1]
    assert n0 % 16 == 0 and n0 == max(pts[i].x.shape[0] for i in idx)
    assert [a.shape[1] for a in ab] == [n0, n0 // 4, n0 // 16]
    for b, i in enumerate(idx):
        p = pts[i]
        n = p.x.shape[0]
        assert np.array_equal(xb[b, :n], p.x) and not xb[b, n:].any()
        for lvl, a in enumerate(p.adjs):
            nl = a.shape[0]
            assert np.array_equal(ab[lvl][b, :nl], a)
            pad = ab[lvl][b, nl:]
            assert not pad[:, 1:].any() and np.array_equal(pad[:, 0], np.arange(nl + 1, ab[lvl].shape[1] + 1))
            assert a.max() <= nl


def test_rotation_matrix_and_feature_rotation():
    rng = np.random.RandomState(5)
    R = T.rand_rotation_matrix(rng)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-12 and abs(np.linalg.det(R) - 1) < 1e-12
    x = torch.randn(1, 10, 6)
    y = T.rotate_features(x, torch.from_numpy(R.astype(np.float32)))
    ref = np.concatenate([x[0, :, :3].numpy() @ R.T, x[0, :, 3:].numpy() @ R.T], axis=1)
    assert np.abs(y[0].numpy() - ref).max() < 1e-5


def _extracted_patches():
    from facet_graph_convolution_b200 import coarsening, mesh
    V, F = mesh.icosphere(4)
    V = mesh.add_vertex_noise(V, F, 0.3, seed=2)
    adj = mesh.faces_large_adj(F, 16)
    ps = coarsening.extract_patches(adj, mesh.face_features(V, F), 1500, 16, min_patch_size=700,
                                    rng=np.random.RandomState(5))
    return ps, F.shape[0]


def _worker_extracted(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        patches, nf = _extracted_patches()  # every rank cuts the same patches from the same seed
        q.put((rank, P.infer_sharded(patches, nf, _fake_forward, rank, world)))
    finally:
        dist.destroy_process_group()


def test_world2_inference_on_patches_cut_by_the_reference_patch_loop():
    """Patches produced by `coarsening.extract_patches` (overlapping context facets, tree-order permutation,
    fake rows) dealt to two ranks: same merged normals as one process and as the reference's merge."""
    patches, nf = _extracted_patches()
    assert len(patches) >= 3 and sum(p.num_real for p in patches) > nf  # context facets are computed twice
    single = P.infer_sharded(patches, nf, _fake_forward, 0, 1)
    acc = np.zeros((nf, 3))
    for p in patches:
        acc[p.face_ids] += _fake_forward(p)[0][p.perm][: p.num_real]
    for _ in range(2):
        acc = acc * (1 / (np.sqrt((acc * acc).sum(1))[:, None] + 1e-8))
    assert np.abs(single - acc).max() < 1e-12
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_extracted, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    for _, pred in res:
        assert np.abs(pred - single).max() < 1e-6


def _sync_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                     # unseeded-init stand-in: every rank draws its own weights
        params = [torch.randn(5, 3), torch.randn(4)]
        b = T.GradBucket(params)
        opt = T.Adam(b)
        opt.m.fill_(float(rank)), opt.v.fill_(float(rank) + 1)
        opt.t = 7 * rank
        caught = False
        try:
            T.sync_replicas(b, opt, check_only=True)      # replicas differ: must raise on every rank
        except RuntimeError:
            caught = True
        T.sync_replicas(b, opt)                           # broadcast from rank 0
        T.sync_replicas(b, opt, check_only=True)          # now identical: must not raise
        q.put((rank, caught, [p.clone().numpy() for p in params], opt.m.clone().numpy(), opt.v.clone().numpy(), opt.t))
    finally:
        dist.destroy_process_group()


def test_world2_replicas_are_synchronised_before_training():
    """train_step only averages gradients; differing initial parameters / Adam state must be caught
    (check_only) and repaired (broadcast from rank 0)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    assert res[0][1] and res[1][1]
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)
    torch.manual_seed(100)
    assert np.array_equal(res[1][2][0], torch.randn(5, 3).numpy())      # rank 0's draw won
    assert np.array_equal(res[0][3], res[1][3]) and np.array_equal(res[0][4], res[1][4])
    assert res[0][5] == res[1][5] == 0


def test_grad_bucket_rejects_an_empty_parameter_list_and_the_network_has_its_parameters_at_construction():
    with pytest.raises(ValueError):
        T.GradBucket([])
    from facet_graph_convolution_b200 import model as fm
    net = fm.DenoisingNet(device="cpu", seed=0)           # no forward has run
    ps = list(net.parameters())
    assert len(ps) == 44 and sum(p.numel() for p in ps) == 474199
    assert tuple(ps[0].shape) == (9, 32, 6) and tuple(ps[-2].shape) == (1024, 3)
    ms = fm.DenoisingNet(device="cpu", seed=0, multi_scale=True)
    assert sum(p.numel() for p in ms.parameters()) == 679005
    T.GradBucket(ps)                                      # an optimizer can be set up before the first forward


# ----------------------------------------------------------------------------- sharded vertex update (config C5)
def _vertex_case():
    from facet_graph_convolution_b200 import mesh
    V, F = mesh.grid_mesh(9, 7)
    rs = np.random.RandomState(4)
    V = (V + rs.randn(*V.shape) * 0.02).astype(np.float32)
    n = rs.randn(F.shape[0], 3).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    em, ve = mesh.edge_maps(F, 20)
    return V, n, em.astype(np.int32), ve.astype(np.int32)


def _oracle_sweep(normals, edge_map, v_edges):
    """one Jacobi sweep of update_position2 over a vertex range, from the oracle's closed form (fp64)"""
    from oracle import closed_form as cf

    def sweep(x_in, x_out, b, e):
        V = v_edges.shape[0]
        full = cf.update_position2(x_in[:V].numpy(), normals, edge_map, v_edges, iter_num=1)
        x_out[b:e] = torch.from_numpy(full[b:e])
    return sweep


def _vertex_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from facet_graph_convolution_b200 import patches
        V, n, em, ve = _vertex_case()
        outs = []
        for mode in ("halo", "allgather"):
            out = patches.vertex_update_edges_sharded(torch.from_numpy(V.astype(np.float64)), torch.from_numpy(n),
                                                      torch.from_numpy(em), torch.from_numpy(ve), iters=7,
                                                      sweep=_oracle_sweep(n, em, ve), exchange=mode)
            outs.append(out.numpy())
        halo = patches.VertexHalo(torch.from_numpy(em), torch.from_numpy(ve))
        q.put((rank, outs, int(halo.need.numel()), sum(halo.send_splits)))
    finally:
        dist.destroy_process_group()


def test_world2_sharded_vertex_update_equals_the_single_process_update():
    from oracle import closed_form as cf
    from facet_graph_convolution_b200 import patches
    V, n, em, ve = _vertex_case()
    ref = cf.update_position2(V.astype(np.float64), n, em, ve, iter_num=7)
    ranges, chunk = patches.vertex_ranges(V.shape[0], 3)
    assert ranges[0][0] == 0 and ranges[-1][1] == V.shape[0] and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert chunk * 3 >= V.shape[0]
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_vertex_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, outs, nneed, nsend = q.get(timeout=120)
        got[r] = (outs, nneed, nsend)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):                      # every rank ends with the whole mesh, equal to the unsharded update
        for out in got[r][0]:
            assert np.array_equal(out, ref), np.abs(out - ref).max()
    # the halo is a thin band, not the mesh: what one rank needs the other sends, and it is far fewer rows than V
    assert got[0][1] == got[1][2] and got[1][1] == got[0][2]
    assert 0 < got[0][1] < V.shape[0] // 3
