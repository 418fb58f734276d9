#!/usr/bin/env python
"""Network-level measurements of SURVEY.md section 8(d): C1, C3 and C4 (bench.py is C2, the headline).

    python benchmarks/net_bench.py --config c1          # single-mesh denoising: icosphere-5, 20 480 faces
    python benchmarks/net_bench.py --config c3          # 2M-facet mesh, ~100 patches with halo, sharded by patch
    python benchmarks/net_bench.py --config c4          # training step on 8 192-node patches (data parallel)
    python benchmarks/net_bench.py --config c5          # C3 at 3162x3162 quads (20M facets) + whole-mesh vertex update
    python benchmarks/net_bench.py --config index       # GPU index builders (adjacency, vertex-face, edge maps) at 2M faces
    python -m torch.distributed.run --nproc-per-node N ... benchmarks/net_bench.py --config c3|c4

Every config prints ONE JSON line (rank 0).  facets/s counts real input faces (C3: core faces, each
written back once); GPU time is CUDA-event time of the network forward (+ vertex update for C1) with
the patch tensors resident in HBM; `e2e` adds the pinned-host upload of the patch tensors and the
read-back of the normals per patch.  The CPU number beside it is the oracle's fp32 NumPy closed form
(`oracle/closed_form.py`, kind "port": TensorFlow is not installable and the reference sources cannot
travel to the GPU box) on a bounded sample, with the core count printed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def net_params(rs, multi_scale=False):
    """Random-init parameters in the reference's creation order (model.py:31-44 stddevs)."""
    M, cin = 9, 6
    shapes = []
    for ci, co in ((cin, 32), (32, 64), (64, 128), (128, 128), (128, 64), (128, 64), (64, 32), (64, 32)):
        shapes += [((M, co, ci), 0.05), ((co,), 0.01), ((M, ci), 0.05), ((M,), 0.05), ((M, ci), 0.05)]
    shapes += [((32, 1024), 0.05), ((1024,), 0.01), ((1024, 3), 0.05), ((3,), 0.01)]
    return [rs.normal(0, sd, sh).astype(np.float32) for sh, sd in shapes]


def icosphere_patch(level=5, K=16, seed=0):
    from facet_graph_convolution_b200 import mesh
    V, F = mesh.icosphere(level)
    Vn = mesh.add_vertex_noise(V, F, 0.3, seed)
    feat = mesh.face_features(Vn, F).astype(np.float32)
    adj = mesh.dedup_adj(mesh.faces_large_adj(F, K))
    featp, adjp = mesh.pad_to_multiple(feat, adj, 16)
    adjs = mesh.build_pyramid(adjp, 3, K)
    e_map, v_e = mesh.edge_maps(F, 20)
    return Vn.astype(np.float32), F, featp, adjs, e_map, v_e


def clocks_sampler(index):
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(index)
    s.start()
    return s


def cpu_net_forward(x, adjs, params, repeat=1):
    """CPU baseline leg: lives in bench.py (the one benchmark module that may execute oracle/)."""
    sys.path.insert(0, ROOT)
    import bench
    return bench.cpu_net_reference(x, adjs, params, repeat)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1", choices=["c1", "c3", "c4", "c5", "index", "mesh"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--grid", type=int, default=1000, help="c3: quads per side (1000 -> 2M facets)")
    ap.add_argument("--block", type=int, default=100)
    ap.add_argument("--batch", type=int, default=4, help="c4: patches per rank per step")
    ap.add_argument("--patch-batch", type=int, default=25, help="c3: patches per launch (1 = the reference's B=1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--graph", action="store_true", help="c1: also time the step replayed from a CUDA graph")
    ap.add_argument("--profile", action="store_true", help="per-kernel CUDA-event times of one forward (library profiler)")
    args = ap.parse_args()

    import torch
    from facet_graph_convolution_b200 import _lib, ops, patches, mesh
    from facet_graph_convolution_b200 import model as fm
    _lib.require_device()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = _lib.lib()
    rs = np.random.RandomState(1234)
    params = net_params(rs)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cores = os.cpu_count() or 1
    line = {"metric": "facets/sec", "unit": "facets/s", "n_gpus": world, "higher_is_better": True, "dtype": "f32",
            "data": "synthetic", "vs_baseline": None}

    if args.config == "c1":
        Vn, F, feat, adjs, e_map, v_e = icosphere_patch(5)
        nreal = F.shape[0]
        x_d, adjs_d = T(feat[None]), [T(a[None]) for a in adjs]
        v_d, em_d, ve_d = T(Vn[None]), T(e_map[None]), T(v_e[None])
        store = fm.VariableStore(dev, params=params)

        def step():
            with torch.no_grad(), fm.variable_store(store):
                y = fm.get_model_reg_multi_scale(x_d, adjs_d, 1.0)
                n = fm.normalizeTensor(y)
                xo = fm.update_position2(v_d, n[:, :nreal].contiguous(), em_d, ve_d, iter_num=60, max_edges=20)
            return n, xo

        def fwd_only():
            with torch.no_grad(), fm.variable_store(store):
                return fm.normalizeTensor(fm.get_model_reg_multi_scale(x_d, adjs_d, 1.0))

        for _ in range(args.warmup):
            step()
        barrier()
        if args.profile:
            import ctypes as C
            L.fgc_profile_begin(C.c_void_p(torch.cuda.current_stream().cuda_stream))
            fwd_only()
            buf = C.create_string_buffer(1 << 16)
            L.fgc_profile_end(buf, len(buf))
            line["kernels_ms"] = {ln.split()[0]: [round(float(ln.split()[1]), 4), int(ln.split()[2])]
                                  for ln in buf.value.decode().strip().splitlines()}
        smp = clocks_sampler(local_rank)
        n0 = L.fgc_launch_count()
        ts, tf = [], []
        for _ in range(args.steps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            fwd_only()
            e1.record()
            step()
            e2.record()
            torch.cuda.synchronize()
            tf.append(e0.elapsed_time(e1))
            ts.append(e1.elapsed_time(e2))
        launches = (L.fgc_launch_count() - n0) / args.steps
        clocks = smp.result()
        ms_net, ms_all = float(np.median(tf)), float(np.median(ts))
        if args.graph:
            # the whole step (all layer launches + 60 vertex sweeps) captured once and replayed: the kernels are
            # the same, only the launch gaps between ~150 small kernels go away
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                n_g, xo_g = step()
            n_e, xo_e = step()
            g.replay()
            torch.cuda.synchronize()
            assert torch.equal(n_g, n_e) and torch.equal(xo_g, xo_e), "graph replay differs from the eager step"
            tg = []
            for _ in range(args.steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                tg.append(e0.elapsed_time(e1))
            line["cuda_graph_ms_per_step"] = float(np.median(tg))
            line["cuda_graph_facets_per_s"] = nreal / (float(np.median(tg)) * 1e-3)
        cpu = None
        if not args.no_cpu:
            dt = cpu_net_forward(feat, adjs, params)
            cpu = {"value": nreal / dt, "unit": "facets/s", "cores": cores, "kind": "port",
                   "sample": "oracle/closed_form.py net_forward (fp32 NumPy), 1 pass over the same patch, %.2f s" % dt}
        line.update({"value": nreal / (ms_all * 1e-3), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all,
                     "scaling": "weak", "clocks": clocks, "gpu_launches": launches,
                     "config": {"workload": "C1 single-mesh denoising: noisy icosphere-5 (20480 faces, N0=%d nodes), K=16, M=9, "
                                            "network forward + normalizeTensor + update_position2 (60 sweeps)" % feat.shape[0]},
                     "network_forward_ms": ms_net, "network_forward_facets_per_s": nreal / (ms_net * 1e-3),
                     "cpu_baseline": cpu})

    elif args.config in ("c3", "c5"):
        if args.config == "c5" and args.grid == 1000:
            args.grid = 3162
        nx = ny = args.grid
        bx = (nx + args.block - 1) // args.block
        npatch = bx * bx
        costs = [1] * npatch   # equal-sized blocks: the partition only needs relative costs
        plan = patches.partition(costs, world)
        t0 = time.perf_counter()
        mine, num_faces = patches.grid_patches(nx, ny, block=args.block, halo=3, K=16, only=plan[rank])
        t_gen = time.perf_counter() - t0
        store = fm.VariableStore(dev, params=params)
        # patches are stacked `--patch-batch` at a time (padded with fake nodes to a common size): the
        # layers treat batch elements independently, so the real rows are those of the B=1 run
        PB = max(1, args.patch_batch)
        groups = [list(range(i, min(i + PB, len(mine)))) for i in range(0, len(mine), PB)]
        host = []
        for g in groups:
            xb, ab = patches.batch_patches(mine, g)
            host.append((torch.from_numpy(xb).pin_memory(), [torch.from_numpy(a).pin_memory() for a in ab],
                         [mine[i].x.shape[0] for i in g]))
        resident = [(x.to(dev), [a.to(dev) for a in adjs], ns) for x, adjs, ns in host]
        core = sum(int(p.core.sum()) for p in mine)

        def fwd(x, adjs, ns):
            with torch.no_grad(), fm.variable_store(store):
                y = fm.get_model_reg_multi_scale(x, adjs, 1.0)
                # normalizeTensor's mean is global per patch (utils.py:1700-1715): per element of the batch
                if len(ns) == 1:
                    return [fm.normalizeTensor(y[:, :ns[0]])]
                cnt = torch.tensor(ns, dtype=torch.int32).to(x.device, non_blocking=True)
                yn = ops.normalize_rows_segmented(y, cnt)
                return [yn[b:b + 1, :n] for b, n in enumerate(ns)]

        if PB > 1 and rank == 0:
            # the batched run reproduces the B=1 rows (bit for bit when both sizes select the same kernels --
            # tests/test_gpu_pipeline.py; small coarse levels run on the FFMA family, batched ones on tcgen05)
            p0 = mine[0]
            one = fwd(torch.from_numpy(p0.x[None]).to(dev), [torch.from_numpy(a[None]).to(dev) for a in p0.adjs],
                      [p0.x.shape[0]])[0]
            dmax = float((one - fwd(*resident[0])[0]).abs().max())
            assert dmax < 1e-5, "batched forward differs from the B=1 forward by %g" % dmax
        for x, adjs, ns in resident[: max(2, args.warmup)]:
            fwd(x, adjs, ns)
        barrier()
        if args.profile:
            import ctypes as C
            L.fgc_profile_begin(C.c_void_p(torch.cuda.current_stream().cuda_stream))
            fwd(*resident[0])
            buf = C.create_string_buffer(1 << 16)
            L.fgc_profile_end(buf, len(buf))
            line["kernels_ms_first_batch"] = {ln.split()[0]: [round(float(ln.split()[1]), 4), int(ln.split()[2])]
                                              for ln in buf.value.decode().strip().splitlines()}
        smp = clocks_sampler(local_rank)
        n0 = L.fgc_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for x, adjs, ns in resident:
            fwd(x, adjs, ns)
        e1.record()
        barrier()
        launches = L.fgc_launch_count() - n0
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = smp.result()
        # end to end: pinned-host patch tensors in, normals out, per batch of patches
        outs = [[torch.empty(1, n, 3).pin_memory() for n in ns] for _, _, ns in host]
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for (x, adjs, ns), os_ in zip(host, outs):
            xd, ad = x.to(dev, non_blocking=True), [a.to(dev, non_blocking=True) for a in adjs]
            for o, yn in zip(os_, fwd(xd, ad, ns)):
                o.copy_(yn, non_blocking=True)
                d2h += o.numel() * 4
            h2d += x.numel() * 4 + sum(a.numel() * 4 for a in adjs)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        total_core = core
        if world > 1:
            t = torch.tensor([core], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            total_core = int(t.item())
        cpu = None
        if rank == 0 and not args.no_cpu:
            p = mine[0]
            dtc = cpu_net_forward(p.x, p.adjs, params)
            cpu = {"value": int(p.core.sum()) / dtc, "unit": "facets/s", "cores": cores, "kind": "port",
                   "sample": "oracle/closed_form.py net_forward (fp32 NumPy) on 1 of %d patches (%d nodes), %.2f s; "
                             "per-patch time x patch count" % (npatch, p.x.shape[0], dtc)}
        vu = None
        if args.config == "c5" and rank == 0:
            # whole-mesh update_position2 (60 Jacobi sweeps) on one GPU: ~10M vertices, ~30M edges
            t0 = time.perf_counter()
            Vg, Fg = mesh.grid_mesh(nx, ny, torus=False, morton=False)
            e_map, v_e = mesh.edge_maps(Fg, 20)
            nrm = mesh.face_normals(Vg, Fg).astype(np.float32)
            t_idx = time.perf_counter() - t0
            vd, nd, ed, ved = T(Vg.astype(np.float32)[None]), T(nrm[None]), T(e_map[None]), T(v_e[None])
            fm.update_position2(vd, nd, ed, ved, iter_num=2, max_edges=20)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            fm.update_position2(vd, nd, ed, ved, iter_num=60, max_edges=20)
            g1.record()
            torch.cuda.synchronize()
            vms = g0.elapsed_time(g1)
            Vn_, En_, Fn_ = Vg.shape[0], e_map.shape[0], Fg.shape[0]
            sweep_bytes = 4 * (3 * Vn_ + 3 * Vn_) + 4 * (20 * Vn_ + 4 * En_) + 4 * 3 * Fn_   # SURVEY 8(d) byte model
            vu = {"ms_60_sweeps": vms, "vertices": Vn_, "edges": En_, "faces": Fn_, "host_index_build_s": t_idx,
                  "algorithmic_GBps": 60 * sweep_bytes / (vms * 1e-3) / 1e9}
        if world > 1:
            dist.barrier()
        line.update({"value": total_core / (ms * 1e-3), "steps": 1, "warmup": max(2, args.warmup), "ms_per_step": ms,
                     "scaling": "strong", "clocks": clocks, "gpu_launches": int(launches), "vertex_update": vu,
                     "config": {"workload": args.config.upper() + " multi-scale denoising net inference: %dx%d-quad height field = %d facets, %d patches "
                                            "of %dx%d quads + 3-quad halo, K=16, M=9, patches dealt to %d GPU(s), no collective, "
                                            "%d patches per launch"
                                            % (nx, ny, num_faces, npatch, args.block, args.block, world, PB),
                                "patches_this_rank": len(mine), "host_patch_generation_s": t_gen},
                     "e2e": {"value": total_core / dt, "unit": "facets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "ms_per_step": dt * 1e3},
                     "cpu_baseline": cpu})

    elif args.config == "index":
        # getFacesLargeAdj + getVerticesFaces + getEdgeMap on the GPU, host faces in,
        # index tensors resident on the device out; the CPU number is the vectorised NumPy restatement
        # (the reference's own Python loops take minutes at this size)
        nx = ny = args.grid
        _, Fh = mesh.grid_mesh(nx, ny, torus=False, morton=True)
        Fh = np.ascontiguousarray(Fh, np.int32)
        nf, nv = Fh.shape[0], int(Fh.max()) + 1
        Fp = torch.from_numpy(Fh).pin_memory()

        def build_all():
            Fd = Fp.to(dev, non_blocking=True)
            adj, vf = ops.build_faces_adj(Fd, K=16, kv=25, nv=nv)
            e_map, v_e = ops.build_edge_maps(Fd, 20, nv=nv)
            return adj, vf, e_map, v_e

        for _ in range(3):
            build_all()
        barrier()
        n0 = L.fgc_launch_count()
        ts = []
        for _ in range(max(5, args.steps // 2)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = build_all()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        launches = (L.fgc_launch_count() - n0) / len(ts)
        ms = float(np.median(ts))
        out_bytes = sum(t.numel() * 4 for t in out)
        cpu = None
        if not args.no_cpu:
            t0 = time.perf_counter()
            a_ref = mesh.faces_large_adj(Fh, 16)
            mesh.vertex_faces(Fh, 25)
            mesh.edge_maps(Fh, 20)
            dtc = time.perf_counter() - t0
            assert np.array_equal(out[0].cpu().numpy(), a_ref)
            cpu = {"value": nf / dtc, "unit": "facets/s", "cores": 1, "kind": "port",
                   "sample": "NumPy restatement (mesh.faces_large_adj + vertex_faces + edge_maps) of the reference's host "
                             "loops on the same %d faces, %.2f s" % (nf, dtc)}
        line.update({"value": nf / (ms * 1e-3), "steps": len(ts), "warmup": 3, "ms_per_step": ms, "scaling": "weak",
                     "gpu_launches": launches,
                     "config": {"workload": "index builders: %dx%d-quad grid = %d faces, %d vertices; adjacency K=16, vertex-face "
                                            "lists (25), edge map + vertex-edge lists (20); faces uploaded from pinned host memory "
                                            "inside the timed region" % (nx, ny, nf, nv)},
                     "output_bytes": out_bytes, "algorithmic_GBps": (nf * 12 + out_bytes) / (ms * 1e-3) / 1e9,
                     "cpu_baseline": cpu})

    elif args.config == "mesh":
        # The reference's `inferNet` from file to file on one GPU (Code/train.py:29-150): OBJ in, adjacency /
        # edge maps / features (GPU), patch pyramid (host, f-2), network from a Saver file (f-3),
        # normalizeTensor, un-permutation, host normalize, 60 vertex-update sweeps, OBJ out (f-4).  Wall-clock
        # per stage (file I/O and host preprocessing are CPU work, so CUDA events alone would miss them).
        import tempfile
        from facet_graph_convolution_b200 import checkpoint, coarsening, mesh_io
        K = 23  # settings.py:23 K_faces
        V0, F0 = mesh.icosphere(5)
        V0 = mesh.add_vertex_noise(V0, F0, 0.3, 0)
        tmp = tempfile.mkdtemp()
        mesh_io.write_mesh(V0, F0, os.path.join(tmp, "noisy.obj"))
        prefix = checkpoint.save_network(os.path.join(tmp, "net"), params, global_step=1)

        def run(seed):
            tm = {}
            t = time.perf_counter()
            Vh, _, _, Fh, _ = mesh_io.load_mesh(tmp, "noisy.obj", 0, False)
            tm["read_obj"] = time.perf_counter() - t
            t = time.perf_counter()
            Fd, Vd = T(Fh.astype(np.int32)), T(Vh)
            adj, _ = ops.build_faces_adj(Fd, K=K, nv=Vh.shape[0])
            e_map, v_e = ops.build_edge_maps(Fd, 20, nv=Vh.shape[0])
            feat = ops.face_features(Vd, Fd)
            adj_h, feat_h = adj.cpu().numpy(), feat.cpu().numpy()
            tm["gpu_index_and_features"] = time.perf_counter() - t
            t = time.perf_counter()
            adjs, x, _, old_to_new = coarsening.patch_pyramid(adj_h, feat_h, K, rng=np.random.RandomState(seed))
            tm["host_pyramid"] = time.perf_counter() - t
            t = time.perf_counter()
            store = fm.VariableStore(dev, params=checkpoint.load_network(checkpoint.latest_checkpoint(tmp), verify=True))
            tm["read_checkpoint"] = time.perf_counter() - t
            t = time.perf_counter()
            with torch.no_grad(), fm.variable_store(store):
                y = fm.get_model_reg_multi_scale(T(x[None].astype(np.float32)), [T(a) for a in adjs], 1.0)
                n = fm.normalizeTensor(y)
                out = ops.gather_perm(n.reshape(-1, 3), T(old_to_new.astype(np.int32)))[: Fh.shape[0]]
                pred = patches.host_normalize(out.cpu().numpy()).astype(np.float32)
                xo = fm.update_position2(Vd[None], T(pred[None]), e_map, v_e, iter_num=60, max_edges=20)
                Vout = xo[0].cpu().numpy()
            tm["network_and_vertex_update"] = time.perf_counter() - t
            t = time.perf_counter()
            mesh_io.write_mesh(Vout, Fh, os.path.join(tmp, "denoised.obj"))
            tm["write_obj"] = time.perf_counter() - t
            return tm, Fh.shape[0]

        for i in range(2):
            run(i)
        n0 = L.fgc_launch_count()
        runs = [run(10 + i) for i in range(5)]
        launches = (L.fgc_launch_count() - n0) / len(runs)
        nf = runs[0][1]
        stages = {k: float(np.median([r[0][k] for r in runs])) * 1e3 for k in runs[0][0]}
        ms = float(np.median([sum(r[0].values()) for r in runs])) * 1e3
        line.update({"value": nf / (ms * 1e-3), "steps": len(runs), "warmup": 2, "ms_per_step": ms, "scaling": "weak",
                     "gpu_launches": launches, "stages_ms": stages,
                     "config": {"workload": "file to file: noisy icosphere-5 OBJ (%d faces) -> adjacency K=%d, edge maps, "
                                            "features on the GPU -> host patch pyramid (4 coarsenings) -> network from a "
                                            "Saver checkpoint -> normals -> 60 vertex-update sweeps -> OBJ; wall clock, "
                                            "median of %d runs" % (nf, K, len(runs))},
                     "cpu_baseline": None})

    else:  # c4
        from facet_graph_convolution_b200 import train as ftrain
        nq = 64   # 64x64 quads = 8192 triangles per patch
        batch = []
        for bi in range(args.batch):
            P, _ = patches.grid_patches(nq, nq, block=nq, halo=0, K=16, seed=rank * 100 + bi)
            p = P[0]
            Vc, Fc = mesh.grid_mesh(nq, nq, torus=False, morton=True)
            gt = np.zeros((p.x.shape[0], 3), np.float32)
            gt[: p.num_real] = p.x[: p.num_real, :3]   # clean-ish target: the patch's own normals (synthetic)
            batch.append((T(p.x[None]), [T(a[None]) for a in p.adjs], T(gt[None])))
        net = fm.DenoisingNet(6, device=dev, params=params)
        net(batch[0][0], batch[0][1])
        plist = list(net.parameters())
        bucket = ftrain.GradBucket(plist)
        opt = ftrain.Adam(bucket)
        rng = np.random.RandomState(rank)
        group = dist.group.WORLD if world > 1 else None
        for _ in range(args.warmup):
            ftrain.train_step(net, batch, bucket, opt, rng, group)
        barrier()
        if args.profile:
            import ctypes as C
            L.fgc_profile_begin(C.c_void_p(torch.cuda.current_stream().cuda_stream))
            ftrain.train_step(net, batch, bucket, opt, rng, group)
            buf = C.create_string_buffer(1 << 16)
            L.fgc_profile_end(buf, len(buf))
            line["kernels_ms"] = {ln.split()[0]: [round(float(ln.split()[1]), 4), int(ln.split()[2])]
                                  for ln in buf.value.decode().strip().splitlines()}
        smp = clocks_sampler(local_rank)
        n0 = L.fgc_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ftrain.train_step(net, batch, bucket, opt, rng, group)
        e1.record()
        barrier()
        launches = (L.fgc_launch_count() - n0) / args.steps
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        clocks = smp.result()
        facets = args.batch * 8192 * world
        line.update({"value": facets / (ms * 1e-3), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                     "scaling": "weak", "clocks": clocks, "gpu_launches": launches,
                     "config": {"workload": "C4 training step: %d patches/GPU x 8192 level-0 nodes, Cin=6, K=16, M=9, random "
                                            "rotation, 4000 sampled facets, faceNormalsLoss, fwd+bwd, one all-reduce of the flat "
                                            "gradient bucket (%d floats), Adam" % (args.batch, bucket.flat.numel()),
                                "parallelism": "dp%d" % world}})

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
