"""BASELINE config C5, vertex part: update_position2 (reference Code/train.py:1467-1557, 60 Jacobi sweeps) over ONE large
mesh, single device against the vertex-sharded update (every rank sweeps its vertex range; per sweep one NCCL all-to-all
of the halo vertices, or one all-gather of all positions).

    python benchmarks/vertex_update_bench.py [--grid 3162] [--iters 60]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        benchmarks/vertex_update_bench.py --gpus N

Every rank builds the mesh and its index tensors itself (edge maps on the GPU, fgc_build_edge_maps); rank 0 checks that the
sharded result equals the single-device result bit for bit and prints one JSON line.  Times are CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from facet_graph_convolution_b200 import mesh, ops, patches  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--grid", type=int, default=3162, help="quads per side (3162 -> 20 M facets, 10 M vertices)")
    ap.add_argument("--iters", type=int, default=60)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    hf = None
    V, F = mesh.grid_mesh(args.grid, args.grid, torus=False, morton=False, height=hf)
    rs = np.random.RandomState(0)
    V = (V + rs.randn(*V.shape) * (0.1 / args.grid)).astype(np.float32)          # noisy positions, clean normals
    Fd = torch.from_numpy(F.astype(np.int32)).to(dev)
    Vd = torch.from_numpy(V).to(dev)
    tri = Vd[Fd.long()]
    n = torch.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0], dim=1)
    n = n / (n.norm(dim=1, keepdim=True) + 1e-8)
    e_map, v_e = ops.build_edge_maps(Fd, 20, nv=V.shape[0])
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    nv, nf, ne = V.shape[0], F.shape[0], int(e_map.shape[0])

    def timed(fn):
        fn(2)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn(args.iters)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return out, ms

    single, ms_single = timed(lambda it: ops.vertex_update_edges(Vd, n, e_map, v_e, iters=it))
    sharded, ms_sharded = timed(lambda it: patches.vertex_update_edges_sharded(Vd, n, e_map, v_e, iters=it, exchange="halo"))
    gathered, ms_gather = timed(lambda it: patches.vertex_update_edges_sharded(Vd, n, e_map, v_e, iters=it,
                                                                                exchange="allgather"))
    ms_p2p, p2p_err = None, None
    same_p2p = True
    if world > 1:
        try:
            pushed, ms_p2p = timed(lambda it: patches.vertex_update_edges_sharded(Vd, n, e_map, v_e, iters=it, exchange="p2p"))
            same_p2p = bool(torch.equal(single.reshape(-1, 3), pushed.reshape(-1, 3)))
        except Exception as ex:     # symmetric memory unavailable on this box / build: the NCCL variants stand
            p2p_err = repr(ex)[:300]
    same = bool(torch.equal(single.reshape(-1, 3), sharded.reshape(-1, 3)) and
                torch.equal(single.reshape(-1, 3), gathered.reshape(-1, 3)) and same_p2p)
    halo_rows = int(patches.VertexHalo(e_map, v_e).need.numel()) if world > 1 else 0
    if world > 1:
        t = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        same = bool(t.item())
    sweep_bytes = 4 * (3 * nv + 3 * nv) + 4 * (20 * nv + 4 * ne) + 4 * 3 * nf        # SURVEY 8(d) byte model of one sweep
    if rank == 0:
        print(json.dumps({
            "metric": "update_position2 sweeps/s over one mesh", "n_gpus": world, "vertices": nv, "faces": nf, "edges": ne,
            "iters": args.iters, "ms_single_device": ms_single, "ms_sharded": ms_sharded, "ms_sharded_allgather": ms_gather,
            "ms_sharded_p2p": ms_p2p, "p2p_error": p2p_err,
            "speedup": ms_single / ms_sharded, "bit_identical": same, "halo_rows_rank0": halo_rows,
            "halo_bytes_per_sweep_rank0": 12 * halo_rows, "allgather_bytes_per_sweep": 12 * nv, "algorithmic_GBps_single": args.iters * sweep_bytes / (ms_single * 1e-3) / 1e9,
            "algorithmic_GBps_sharded": args.iters * sweep_bytes / (ms_sharded * 1e-3) / 1e9, "setup_s": t_setup,
            "config": {"workload": "C5 vertex update: %dx%d-quad height field, %d sweeps; sharded = contiguous vertex ranges over "
                                   "%d rank(s), one NCCL all-to-all of the halo positions per sweep (ms_sharded; setup of the "
                                   "halo lists included) or one all-gather of all positions per sweep (ms_sharded_allgather)" % (args.grid, args.grid,
                                                                                                   args.iters, world)}}))
    if world > 1:
        dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
