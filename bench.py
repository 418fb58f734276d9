#!/usr/bin/env python
"""Headline benchmark: facets/s through one facet-graph convolution layer, forward + backward
(BASELINE.json configs[1], "C2": N = 1 000 000 facets, K = 16, M = 8, Cin = Cout = 64, fp32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--adjacency mesh|dedup|random] [--facets N]

One "step" = one forward + one backward pass of the layer over one batch of synthetic input
(mesh-like adjacency of a 1000x500-quad torus in the reference's getFacesLargeAdj layout).
Under torchrun every rank runs the same-sized shard (weak scaling); the only exchange is the
data-parallel all-reduce of the layer's parameter gradients (NCCL), as in training.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte model.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_NEIGH, M_W, C_IN, C_OUT = 16, 8, 64, 64


# ----------------------------------------------------------------------------- workload
def make_adjacency(n_facets: int, kind: str, seed: int = 0) -> np.ndarray:
    from facet_graph_convolution_b200 import mesh
    if kind == "random":
        rs = np.random.RandomState(seed)
        adj = rs.randint(1, n_facets + 1, size=(n_facets, K_NEIGH)).astype(np.int32)
        adj[:, 0] = np.arange(1, n_facets + 1)
        return adj
    ny = max(2, int(round((n_facets / 4) ** 0.5)))
    nx = max(2, n_facets // (2 * ny))
    assert 2 * nx * ny == n_facets, "facets must be 2*nx*ny (default 1000x500 quads)"
    _, F = mesh.grid_mesh(nx, ny, torus=True, morton=True)
    adj = mesh.faces_large_adj(F, K_NEIGH)
    if kind == "dedup":
        adj = mesh.dedup_adj(adj)
    return adj


def make_params(seed: int):
    rs = np.random.RandomState(seed)
    W0 = rs.normal(0, 0.05, (M_W, C_OUT, C_IN)).astype(np.float32)
    b = rs.normal(0, 0.01, (C_OUT,)).astype(np.float32)
    u = rs.normal(0, 0.05, (M_W, C_IN)).astype(np.float32)
    c = rs.normal(0, 0.05, (M_W,)).astype(np.float32)
    v = rs.normal(0, 0.05, (M_W, C_IN)).astype(np.float32)
    return W0, b, u, v, c


def bytes_fwd(n):
    return 4 * n * (C_IN + K_NEIGH + C_OUT) + 4 * (M_W * C_IN * C_OUT + 2 * M_W * C_IN + M_W + C_OUT)


def bytes_bwd(n):
    return 4 * n * (2 * C_IN + K_NEIGH + C_OUT) + 8 * (M_W * C_IN * C_OUT + 2 * M_W * C_IN + M_W + C_OUT)


# algorithmic bytes each kernel must move per launch: the layer tensors it reads/writes once
# (x, adj, gy, y, gx); workspace traffic (uvx, da_edge, partials) is implementation overhead
# and deliberately NOT counted.
def kernel_bytes(name, n):
    p = 4 * (M_W * C_IN * C_OUT)
    table = {
        "conv_fwd_kernel": bytes_fwd(n),
        "assign_logits_kernel": 4 * n * C_IN,
        "bwd_src_kernel": 4 * n * (C_IN + K_NEIGH + C_OUT) + p,
        "bwd_tgt_kernel": 4 * n * (C_OUT + C_IN) + p,
        "bwd_w_kernel": 4 * n * (C_IN + K_NEIGH + C_OUT) + p,
        "logits_bwd_kernel": 4 * n * 2 * C_IN,
        "conv_mma_kernel": bytes_fwd(n),
        "bwd_tgt_mma_kernel": 4 * n * (C_OUT + C_IN) + p,
        "bwd_src_mma_kernel": 4 * n * (C_IN + K_NEIGH + C_OUT) + p,
        "bwd_w_mma_kernel": 4 * n * (C_IN + K_NEIGH + C_OUT) + p,
        "prep_x_image_kernel": 4 * n * C_IN,
        "logits_bwd_x_kernel": 4 * n * 2 * C_IN,
        "logits_bwd_p_kernel": 4 * n * C_IN,
        "absmax_kernel": 4 * n * C_IN,
    }
    return table.get(name.replace("_tc_kernel", "_kernel"))


def measured_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json, written by profiles/summarize.py), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t["kernels"].get(name, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_reference_run(steps, warmup, sample_facets, adjacency_kind):
    """Times the reference-order CPU port (oracle/ref_order.py) on a bounded sample."""
    import torch
    from oracle import ref_order
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ny = max(2, int(round((sample_facets / 4) ** 0.5)))
    nx = max(2, sample_facets // (2 * ny))
    n = 2 * nx * ny
    adj = torch.from_numpy(make_adjacency(n, adjacency_kind)).unsqueeze(0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, n, C_IN, generator=g)
    gy = torch.randn(1, n, C_OUT, generator=g)
    W0, b, u, v, c = (torch.from_numpy(t) for t in make_params(1234))
    for _ in range(warmup):
        ref_order.conv_fwd_bwd(x, adj, gy, W0, b, u, v, c)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_order.conv_fwd_bwd(x, adj, gy, W0, b, u, v, c)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps * 1e3, n, cores


def cpu_net_reference(x, adjs, params, repeat=1):
    """cpu_baseline leg of the network-level runs (benchmarks/net_bench.py): seconds per forward of the
    oracle's closed form of the reference network (Code/model.py:837-946) in fp32 NumPy, all BLAS threads."""
    from oracle import closed_form as cf
    pd = cf.split_net_params(params)
    t0 = time.perf_counter()
    for _ in range(repeat):
        cf.net_forward(x[None].astype(np.float32), [a[None] for a in adjs], pd, dtype=np.float32)
    return (time.perf_counter() - t0) / repeat


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--adjacency", default="mesh", choices=["mesh", "dedup", "random"])
    ap.add_argument("--facets", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=20_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "C2 facet-graph conv layer fwd+bwd: N=%d facets/GPU, K=%d, M=%d, Cin=Cout=%d, fp32, %s adjacency" % (
        args.facets, K_NEIGH, M_W, C_IN, args.adjacency)
    config = {"workload": workload, "facets_per_gpu": args.facets, "K": K_NEIGH, "M": M_W, "Cin": C_IN,
              "Cout": C_OUT, "adjacency": args.adjacency, "parallelism": "dp%d" % world,
              "l2": "inputs (x 256 MB + adj 64 MB + gy 256 MB) larger than the 126 MB L2; no flush"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        warm = max(1, min(args.warmup, 1))
        val, ms, n, cores = cpu_reference_run(steps, warm, args.cpu_sample, args.adjacency)
        line = {"impl": "reference", "metric": "facets/sec", "value": val, "unit": "facets/s", "n_gpus": 0,
                "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "facets/s", "cores": cores, "kind": "port",
                                 "sample": "reference-order torch-CPU port (oracle/ref_order.py), fwd+bwd on %d "
                                           "facets per step (the reference materialises 32 KB/facet/pass; "
                                           "1M facets would need >100 GB); TensorFlow unavailable" % n},
                "e2e": {"value": val, "unit": "facets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    from facet_graph_convolution_b200 import _lib, ops
    _lib.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    n = args.facets
    adj_np = make_adjacency(n, args.adjacency, seed=rank)
    W0, b, u, v, c = (torch.from_numpy(t).to(dev) for t in make_params(1234))
    g = torch.Generator(device="cpu").manual_seed(rank)
    x_h = torch.randn(1, n, C_IN, generator=g).pin_memory()
    gy_h = torch.randn(1, n, C_OUT, generator=g).pin_memory()
    adj_h = torch.from_numpy(adj_np).unsqueeze(0).pin_memory()
    x, gy, adj = x_h.to(dev), gy_h.to(dev), adj_h.to(dev)
    # caller-owned caches, built once per adjacency outside the step (pure index work on adj):
    # reverse adjacency (+ its padded form / tile plan for the gx pass) and the forward tile plan
    rev = ops.ReverseAdjacency(adj)
    plan = ops.ConvPlan(adj, M_W)
    rev.target_plan(M_W)

    def step():
        # as in a training step, the backward reuses what the forward saved (logits, fp16 image of x)
        saved = ops.ConvSaved()
        y = ops.conv_fwd(x, adj, W0, b, u, v, c, plan=plan, save=saved)
        grads = ops.conv_bwd(gy, x, adj, rev, W0, u, v, c, plan=plan, saved=saved)
        if world > 1:
            # one all-reduce of the layer's parameter gradients, then the 1/world average on the compute stream
            # (as train.GradBucket.all_reduce_mean does): the next step's kernels are ordered behind the
            # collective, so an NCCL kernel still waiting for its peers never shares the SMs with the
            # persistent one-CTA-per-SM kernels (whose static tile split would wait for the displaced CTA)
            flat = torch.cat([t.reshape(-1) for t in grads[1:]])
            dist.all_reduce(flat)
            flat.mul_(1.0 / world)
        return y, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.fgc_launch_count()
    stream = torch.cuda.current_stream()
    profile = world == 1
    if profile:
        L.fgc_profile_begin(C.c_void_p(stream.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    prof_txt = ""
    if profile:
        buf = C.create_string_buffer(1 << 16)
        L.fgc_profile_end(buf, len(buf))
        prof_txt = buf.value.decode()
    launches = L.fgc_launch_count() - launches0
    clocks = sampler.result()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    if not profile:  # separate profiled pass (keeps the all-reduce out of the kernel brackets)
        L.fgc_profile_begin(C.c_void_p(stream.cuda_stream))
        for _ in range(args.steps):
            step_saved = ops.ConvSaved()
            ops.conv_fwd(x, adj, W0, b, u, v, c, plan=plan, save=step_saved)
            ops.conv_bwd(gy, x, adj, rev, W0, u, v, c, plan=plan, saved=step_saved)
        buf = C.create_string_buffer(1 << 16)
        L.fgc_profile_end(buf, len(buf))
        prof_txt = buf.value.decode()
        barrier()

    ms_per_step = ms_total / args.steps
    value = n * world * args.steps / (ms_total * 1e-3)

    # ---- per-kernel breakdown and roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernels = {}
    for ln in prof_txt.strip().splitlines():
        nm, tot, cnt = ln.split()
        kernels[nm] = {"ms_per_launch": float(tot) / int(cnt), "launches_per_step": int(cnt) / args.steps,
                       "ms_per_step": float(tot) / args.steps}
        kb = kernel_bytes(nm, n)
        if kb:
            gbs = kb / (kernels[nm]["ms_per_launch"] * 1e-3) / 1e9
            kernels[nm].update({"algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak_gbs})
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    roofline = None
    if dom:
        kb = kernel_bytes(dom, n)
        ach = (kb / (kernels[dom]["ms_per_launch"] * 1e-3)) / 1e9 if kb else None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                    "frac": (ach / peak_gbs) if ach else None, "traffic": measured_traffic(dom),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kb, "ms_per_launch": kernels[dom]["ms_per_launch"]}
    layer = {}
    if kernels:
        f_ms = kernels.get("conv_mma_kernel", kernels.get("conv_fwd_tc_kernel", kernels.get("conv_fwd_kernel", {}))).get("ms_per_step", 0.0)
        tot_ms = sum(v["ms_per_step"] for v in kernels.values())
        layer = {"fwd_main_kernel_ms": f_ms, "fwd_main_kernel_frac_of_hbm_roofline":
                 (bytes_fwd(n) / (f_ms * 1e-3) / 1e9 / peak_gbs) if f_ms else None,
                 "step_algorithmic_GBps": (bytes_fwd(n) + bytes_bwd(n)) / (ms_per_step * 1e-3) / 1e9,
                 "step_frac_of_hbm_roofline": (bytes_fwd(n) + bytes_bwd(n)) / (ms_per_step * 1e-3) / 1e9 / peak_gbs,
                 "kernel_ms_sum_per_step": tot_ms}

    # ---- end-to-end through the C ABI with HOST buffers (H2D + kernels + D2H inside the call)
    e2e = None
    if not args.no_e2e:
        s = _lib.ConvShape(1, n, K_NEIGH, C_IN, C_IN, 0, C_IN, C_OUT, M_W)
        hp = [t.cpu().contiguous() for t in (W0, b, u, v, c)]
        y_h = torch.empty(1, n, C_OUT).pin_memory()
        gx_h = torch.empty(1, n, C_IN).pin_memory()
        gouts = [torch.empty_like(t).pin_memory() for t in (hp[0], hp[1], hp[2], hp[3], hp[4])]
        P = lambda t: C.c_void_p(t.data_ptr())

        def e2e_step():
            _lib.check(L.fgc_conv_fwd_bwd_host(C.byref(s), P(x_h), P(adj_h), P(gy_h), P(hp[0]), P(hp[1]), P(hp[2]),
                                                P(hp[3]), P(hp[4]), P(y_h), P(gx_h), P(gouts[0]), P(gouts[1]),
                                                P(gouts[2]), P(gouts[3]), P(gouts[4]), 1, local_rank),
                       "fgc_conv_fwd_bwd_host")

        e2e_step()
        e2e_step()
        barrier()
        ksteps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()  # synchronous: returns after the D2H copies completed
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        npar = sum(t.numel() for t in hp)
        e2e = {"value": n * world * ksteps / dt, "unit": "facets/s",
               "h2d_bytes_per_step": 4 * (x_h.numel() + adj_h.numel() + gy_h.numel() + npar),
               "d2h_bytes_per_step": 4 * (y_h.numel() + gx_h.numel() + npar),
               "steps": ksteps, "ms_per_step": dt / ksteps * 1e3,
               "api": "fgc_conv_fwd_bwd_host (C ABI, pinned host buffers; the reverse adjacency and both tile "
                      "plans are rebuilt from adj inside every call)"}
        L.fgc_host_release()

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, ms, ns, cores = cpu_reference_run(3, 1, args.cpu_sample, args.adjacency)
        cpu_baseline = {"value": val, "unit": "facets/s", "cores": cores, "kind": "port",
                        "sample": "oracle/ref_order.py (reference evaluation order on torch-CPU, autograd backward), "
                                  "3 fwd+bwd steps on %d facets, %.0f ms/step" % (ns, ms)}

    if rank == 0:
        line = {"metric": "facets/sec", "value": value, "unit": "facets/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "kernels": kernels, "layer": layer}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
