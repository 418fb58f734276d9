#!/usr/bin/env python
"""Headline benchmark: facets/s of denoise inference -- BASELINE.json `metric`, quoted on configs[2] ("C3"): the full
multi-scale denoising network (reference Code/model.py:837-946 driven as Code/train.py:100-126 drives it, patch by
patch) on a synthetic 2 000 000-facet mesh cut into 100 halo patches, patches dealt to the ranks, NO collective.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3|c2]

One "step" = one pass of the network forward + normalizeTensor over ALL patches of the mesh (every rank runs its
share; strong scaling); `value` = real (core) facets x steps / max-over-ranks CUDA-event time with the patch tensors
resident in HBM; `e2e` = the same pass from pinned HOST patch tensors to pinned host normals (copies inside the timed
region).  `layers` / `roofline` give every layer's algorithmic GB/s (SURVEY section 8(d) byte model) against the
measured HBM peak.  `--config c2` (also summarised under `layer` of the default line) is BASELINE configs[1]: one
facet-graph convolution layer, N = 1 M facets, K = 16, M = 8, Cin = Cout = 64, forward + backward, weak scaling with
one NCCL all-reduce of the parameter gradients.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
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


def measured_traffic(name, rows_per_launch=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json, written by profiles/summarize.py), or None.
    An entry that records the rows of the captured launch is scaled to `rows_per_launch` of this run (the traffic of
    these kernels is proportional to the rows they process)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t["kernels"].get(name, {})
        v = e.get("dram_bytes_per_launch")
        if v is not None and rows_per_launch and e.get("rows_per_launch"):
            v = int(round(v * rows_per_launch / e["rows_per_launch"]))
        return v
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
            time.sleep(0.02)

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


# ----------------------------------------------------------------------------- C2: one layer, forward + backward
def run_c2(args, as_record=False):
    """BASELINE configs[1].  Returns the JSON line as a dict (rank 0; None elsewhere).  as_record: the short run whose
    summary the default (C3) line carries under `layer` -- no e2e, no CPU leg, no process group of its own."""
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
            return None
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
        return line

    # ------------------------------------------------------------------ B200 arm
    import torch
    from facet_graph_convolution_b200 import _lib, ops
    _lib.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if as_record:
        world = 1
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
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

    # ---- one-shot forward (inference sees an adjacency once: the tile plan is part of the cost) beside the amortised one
    one_shot = None
    try:
        def ev_ms(fn, reps=5):
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return float(np.median(ts))
        one_shot = {"plan_build_ms": ev_ms(lambda: ops.ConvPlan(adj, M_W)),
                    "forward_planned_ms": ev_ms(lambda: ops.conv_fwd(x, adj, W0, b, u, v, c, plan=plan)),
                    "forward_plan_free_ms": ev_ms(lambda: ops.conv_fwd(x, adj, W0, b, u, v, c)),
                    "note": "amortised (training: the plan is built once per adjacency) = forward_planned_ms; one-shot planned "
                            "= plan_build_ms + forward_planned_ms; the plan-free kernel needs no plan"}
    except Exception as e:  # noqa: BLE001
        one_shot = {"error": repr(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, ms, ns, cores = cpu_reference_run(3, 1, args.cpu_sample, args.adjacency)
        cpu_baseline = {"value": val, "unit": "facets/s", "cores": cores, "kind": "port",
                        "sample": "oracle/ref_order.py (reference evaluation order on torch-CPU, autograd backward), "
                                  "3 fwd+bwd steps on %d facets, %.0f ms/step" % (ns, ms)}

    line = None
    if rank == 0 or as_record:
        line = {"metric": "facets/sec", "value": value, "unit": "facets/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "kernels": kernels, "layer": layer, "one_shot": one_shot}
    return line


# ----------------------------------------------------------------------------- C3: the network on a patched mesh
NET_K, NET_M = 16, 9
# (name, Cin, Cout, level of the rows, pooled second output) in execution order -- Code/model.py:853-932
NET_LAYERS = [("conv1", 6, 32, 0, True), ("conv2", 32, 64, 1, True), ("conv3", 64, 128, 2, False),
              ("dconv3", 128, 128, 2, False), ("upconv2", 128, 64, 1, False), ("dconv2", 128, 64, 1, False),
              ("upconv1", 64, 32, 0, False), ("dconv1", 64, 32, 0, False)]


def net_params(seed=1234):
    """Random-init parameters in the reference's creation order (Code/model.py:31-44 std-devs)."""
    rs = np.random.RandomState(seed)
    shapes = []
    for _, ci, co, _, _ in NET_LAYERS:
        shapes += [((NET_M, co, ci), 0.05), ((co,), 0.01), ((NET_M, ci), 0.05), ((NET_M,), 0.05), ((NET_M, ci), 0.05)]
    shapes += [((32, 1024), 0.05), ((1024,), 0.01), ((1024, 3), 0.05), ((3,), 0.01)]
    return [rs.normal(0, sd, sh).astype(np.float32) for sh, sd in shapes]


def layer_bytes(name, rows0, K=NET_K, M=NET_M):
    """Algorithmic bytes of one layer call over rows0 level-0 rows: x read once, adj read once,
    y written once (+ the pooled copy when the pooling is fused), parameters once; head: 4 N (32 + 3); normalise 2 4 N 3."""
    if name == "head":
        return 4 * rows0 * (32 + 3) + 4 * (32 * 1024 + 1024 + 1024 * 3 + 3)
    if name == "normalize":
        return 2 * 4 * rows0 * 3
    for nm, ci, co, lvl, pooled in NET_LAYERS:
        if nm == name:
            n = rows0 >> (2 * lvl)
            b = 4 * n * (ci + K + co) + 4 * (M * ci * co + 2 * M * ci + M + co)
            return b + (4 * (n // 4) * co if pooled else 0)
    return None


# library profiler names -> layer (net_fwd.cu tags every launch of the fused forward with its layer)
def _layer_of(kernel_name):
    for nm, *_ in NET_LAYERS:
        if kernel_name in (nm, "prep_" + nm):
            return nm
    return {"conv_fwd_small_kernel": "conv1", "pool_max_kernel": "conv1", "absmax_kernel": "conv1",
            "mlp_head_tc_kernel": "head", "mlp_head_kernel": "head", "prep_head_w_kernel": "head",
            "seg_abs_sum_kernel": "normalize", "seg_normalize_kernel": "normalize", "abs_sum_kernel": "normalize",
            "normalize_rows_kernel": "normalize"}.get(kernel_name)


def make_c3_patches(grid, block, only):
    from facet_graph_convolution_b200 import patches
    return patches.grid_patches(grid, grid, block=block, halo=3, K=NET_K, only=only)


def cpu_net_run(steps, warmup, grid, block):
    """Reference arm / cpu_baseline leg: the oracle's closed form of the reference network (oracle/closed_form.py,
    fp32 NumPy with all BLAS threads) + normalizeTensor on ONE patch of the C3 mesh per step."""
    from oracle import closed_form as cf
    mine, _ = make_c3_patches(grid, block, [0])
    p = mine[0]
    pd = cf.split_net_params(net_params())
    x, adjs = p.x[None].astype(np.float32), [a[None] for a in p.adjs]
    run = lambda: cf.normalize_tensor(cf.net_forward(x, adjs, pd, dtype=np.float32), dtype=np.float32)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return int(p.core.sum()) / dt, dt * 1e3, p.x.shape[0], int(p.core.sum()), os.cpu_count() or 1


def run_c3(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    grid, block, PB = args.grid, args.block, max(1, args.patch_batch)
    npatch = ((grid + block - 1) // block) ** 2
    workload = ("C3 denoise inference: multi-scale network (8 facet-graph convs, M=%d, K=%d, 474199 params) + normalizeTensor "
                "on a %dx%d-quad height field = %d facets in %d patches of %dx%d quads + 3-quad halo, patches dealt to "
                "%d GPU(s), no collective, <= %d patches per launch" % (NET_M, NET_K, grid, grid, 2 * grid * grid, npatch,
                                                                      block, block, world, PB))
    config = {"workload": workload, "facets": 2 * grid * grid, "patches": npatch, "K": NET_K, "M": NET_M,
              "parallelism": "patch-sharded x%d (no data-path collective)" % world,
              "streams": "launch groups of a pass alternate over %d CUDA stream(s)" % max(1, args.streams),
              "l2": "patch tensors of one pass (x + 3 adjacency levels, ~240 MB per 2 M facets) exceed the 126 MB L2; "
                    "every pass re-reads them, no flush"}
    if args.impl == "reference":
        if rank != 0:
            return None
        steps, warm = max(1, min(args.steps, 10)), max(1, min(args.warmup, 1))
        val, ms, nodes, core, cores = cpu_net_run(steps, warm, grid, block)
        return {"impl": "reference", "metric": "facets/sec (denoise inference, fp32)", "value": val, "unit": "facets/s",
                "n_gpus": 0, "steps": steps,
                "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "facets/s", "cores": cores, "kind": "port",
                                 "sample": "oracle/closed_form.py net_forward + normalize_tensor (fp32 NumPy restatement of "
                                           "Code/model.py:837-946; TensorFlow is not installable and the Python reference "
                                           "cannot travel to the GPU box), one of the %d patches per step (%d nodes, %d core "
                                           "facets), whole mesh = per-patch time x %d" % (npatch, nodes, core, npatch)},
                "e2e": {"value": val, "unit": "facets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}

    import torch
    from facet_graph_convolution_b200 import _lib, ops, patches
    from facet_graph_convolution_b200 import model as fm
    _lib.require_device()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

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

    plan = patches.partition([1] * npatch, world)          # equal-sized blocks: every rank derives the same deal
    t0 = time.perf_counter()
    mine, num_faces = make_c3_patches(grid, block, plan[rank])
    t_gen = time.perf_counter() - t0
    store = fm.VariableStore(dev, params=net_params())
    def stack(groups):
        out = []
        for g in groups:
            xb, ab = patches.batch_patches(mine, g)
            out.append((torch.from_numpy(xb).pin_memory(), [torch.from_numpy(a).pin_memory() for a in ab],
                        torch.tensor([mine[i].x.shape[0] for i in g], dtype=torch.int32).pin_memory()))
        return out

    # launch groups of <= PB patches
    groups = [list(range(i, min(i + PB, len(mine)))) for i in range(0, len(mine), PB)]
    host = stack(groups)
    # the end-to-end pass starts computing as soon as a SMALL first group is on the device and uploads growing groups under
    # the forward of the previous one (about 1/10, 3/10, 6/10 of the rank's patches): upload(first) + compute instead of
    # upload(half) + compute
    n = len(mine)
    if n >= 6:
        # measured on one GPU (100 patches): 10,40 -> 10.93 ms, 5,25,60 -> 10.70, 8,30,65 -> 10.68, 5,30 -> 11.44; few
        # patches per rank (multi-GPU) keep three groups (2 GPUs, 50 patches each: 5.72 ms with three, 5.93 with four)
        split = args.e2e_split if args.e2e_split != "auto" else ("8,30,65" if n >= 80 else "10,40")
        cuts = sorted({min(n - 1, max(1, (int(pc) * n) // 100)) for pc in split.split(",") if pc})
        bounds = [0] + cuts + [n]
        host_e2e = stack([list(range(a, b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a])
    elif n >= 2:
        host_e2e = stack([list(range(0, n // 2)), list(range(n // 2, n))])
    else:
        host_e2e = host
    resident = [(x.to(dev), [a.to(dev) for a in adjs], ns.to(dev)) for x, adjs, ns in host]
    core = sum(int(p.core.sum()) for p in mine)
    rows0 = sum(int(x.shape[0] * x.shape[1]) for x, _, _ in host)     # level-0 rows launched per pass (halo + padding incl.)

    def fwd(x, adjs, cnt):
        """the public API: the reference's network function + normalizeTensor per patch (utils.py:1700-1715)"""
        with torch.no_grad(), fm.variable_store(store):
            y = fm.get_model_reg_multi_scale(x, adjs, 1.0)
            return ops.normalize_rows_segmented(y, cnt)

    # launch groups are independent: alternating them over two streams lets the ramp-up of one group's kernel fill the
    # tail of the other's (every kernel is one persistent CTA per SM)
    nstreams = max(1, min(args.streams, len(resident)))
    side = [torch.cuda.Stream() for _ in range(nstreams - 1)]

    def one_pass():
        if not side:
            for x, adjs, cnt in resident:
                fwd(x, adjs, cnt)
            return
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        for s_ in side:
            s_.wait_event(start)
        for i, (x, adjs, cnt) in enumerate(resident):
            k = i % nstreams
            if k == 0:
                fwd(x, adjs, cnt)
            else:
                with torch.cuda.stream(side[k - 1]):
                    fwd(x, adjs, cnt)
        for s_ in side:
            done = torch.cuda.Event()
            done.record(s_)
            main.wait_event(done)

    for _ in range(max(args.warmup, 3)):
        one_pass()
    barrier()
    # per-layer CUDA-event times of one pass (library profiler; outside the timed region)
    stream = torch.cuda.current_stream()
    L.fgc_profile_begin(C.c_void_p(stream.cuda_stream))
    for x, adjs, cnt in resident:      # single stream: the profiler brackets every launch with events on one stream
        fwd(x, adjs, cnt)
    buf = C.create_string_buffer(1 << 16)
    L.fgc_profile_end(buf, len(buf))
    prof = {ln.split()[0]: (float(ln.split()[1]), int(ln.split()[2])) for ln in buf.value.decode().strip().splitlines()}
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = L.fgc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    barrier()
    launches = (L.fgc_launch_count() - n0)
    clocks = sampler.result()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    total_core, total_rows0 = core, rows0
    if world > 1:
        t = torch.tensor([core, rows0, launches], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        total_core, total_rows0, launches = (int(v) for v in t.tolist())
    ms_per_step = ms_total / args.steps
    value = total_core * args.steps / (ms_total * 1e-3)

    # ---- per-layer roofline (this rank's pass): algorithmic bytes / (layer's kernels incl. its pre-pass)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    layers, other_ms = {}, 0.0
    for kn, (ms_k, cnt) in prof.items():
        ln = _layer_of(kn)
        if ln is None:
            other_ms += ms_k
            continue
        d = layers.setdefault(ln, {"ms_per_pass": 0.0, "launches_per_pass": 0, "kernels": {}})
        d["ms_per_pass"] += ms_k
        d["launches_per_pass"] += cnt
        d["kernels"][kn] = round(ms_k, 4)
    for ln, d in layers.items():
        kb = layer_bytes(ln, rows0)
        d["algorithmic_bytes_per_pass"] = kb
        d["algorithmic_GBps"] = kb / (d["ms_per_pass"] * 1e-3) / 1e9
        d["frac_of_hbm_peak"] = d["algorithmic_GBps"] / peak_gbs
    net_bytes = sum(d["algorithmic_bytes_per_pass"] for d in layers.values())
    prof_ms = sum(d["ms_per_pass"] for d in layers.values()) + other_ms
    dom = max(layers, key=lambda k: layers[k]["ms_per_pass"]) if layers else None
    roofline = None
    if dom:
        d = layers[dom]
        main_k = max(d["kernels"], key=lambda k: d["kernels"][k])
        roofline = {"bound": "hbm", "kernel": main_k, "layer": dom, "achieved": d["algorithmic_GBps"], "peak": peak_gbs,
                    "unit": "GB/s", "frac": d["frac_of_hbm_peak"], "traffic": measured_traffic(main_k, rows0 / max(1, len(groups))), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_pass"] / max(1, len(groups)),
                    "ms_per_launch": d["ms_per_pass"] / max(1, len(groups)),
                    "note": "dominant LAYER of the pass (its convolution launches + its pre-pass); one 'launch' = the layer over "
                            "one batch of %d patches; bytes per SURVEY 8(d): 4 B N (Cin + K + Cout) + parameters" % PB,
                    "network": {"algorithmic_bytes_per_pass": net_bytes, "algorithmic_GBps": net_bytes / (prof_ms * 1e-3) / 1e9,
                                "frac_of_hbm_peak": net_bytes / (prof_ms * 1e-3) / 1e9 / peak_gbs,
                                "bytes_per_level0_row": net_bytes / max(1, rows0)}}

    # ---- end to end: pinned HOST patch tensors in, pinned host normals out, uploads of batch i+1 under batch i
    e2e = None
    if not args.no_e2e:
        outs = [torch.empty(x.shape[0], x.shape[1], 3).pin_memory() for x, _, _ in host_e2e]
        copy_s = torch.cuda.Stream()
        down_s = torch.cuda.Stream()
        h2d = sum(x.numel() * 4 + sum(a.numel() * 4 for a in adjs) + ns.numel() * 4 for x, adjs, ns in host_e2e)
        d2h = sum(o.numel() * 4 for o in outs)

        def e2e_pass():
            cur = torch.cuda.current_stream()
            staged = None

            def upload(i):
                with torch.cuda.stream(copy_s):
                    x, adjs, ns = host_e2e[i]
                    t = (x.to(dev, non_blocking=True), [a.to(dev, non_blocking=True) for a in adjs], ns.to(dev, non_blocking=True))
                    ev = torch.cuda.Event()
                    ev.record(copy_s)
                return t, ev

            staged = upload(0)
            for i in range(len(host_e2e)):
                (xd, ad, nd), ev = staged
                if i + 1 < len(host_e2e):
                    staged = upload(i + 1)
                cur.wait_event(ev)
                yn = fwd(xd, ad, nd)
                for t in (xd, nd, *ad):
                    t.record_stream(cur)
                # the normals of this group leave on their own stream, under the forward of the next group
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(down_s):
                    down_s.wait_event(done)
                    outs[i].copy_(yn, non_blocking=True)
                yn.record_stream(down_s)
            torch.cuda.synchronize()

        e2e_pass()
        e2e_pass()
        barrier()
        ksteps = max(2, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_pass()
        dt = max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            h2d, d2h = (int(v) for v in t.tolist())
        e2e = {"value": total_core * ksteps / dt, "unit": "facets/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": ksteps, "ms_per_step": dt / ksteps * 1e3,
               "api": "model.get_model_reg_multi_scale + normalizeTensor per patch batch (fgc_net_fwd / "
                      "fgc_normalize_rows_segmented through the C ABI) from pinned host patch tensors (features + 3-level "
                      "adjacency pyramid) to pinned host normals; uploads of the next group and the download of the previous group's normals overlap the current forward"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, ms, nodes, pc, cores = cpu_net_run(2, 1, grid, block)
        cpu_baseline = {"value": val, "unit": "facets/s", "cores": cores, "kind": "port",
                        "sample": "oracle/closed_form.py net_forward + normalize_tensor (fp32 NumPy), 2 passes over 1 of the %d "
                                  "patches (%d nodes, %d core facets), %.0f ms per patch" % (npatch, nodes, pc, ms)}
    layer_rec = None
    if rank == 0 and world == 1 and not args.no_layer:
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup, sub.no_e2e, sub.no_cpu_baseline, sub.impl = 5, 3, True, True, "b200"
        r = run_c2(sub, as_record=True)
        layer_rec = {"workload": r["config"]["workload"], "facets_per_s": r["value"], "ms_per_step": r["ms_per_step"],
                     "roofline": r["roofline"], "layer": r["layer"], "one_shot": r.get("one_shot"),
                     "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in r["kernels"].items()}}
    # C4 (every rank: the step holds the all-reduce) and C1 (rank 0) beside the headline; a failure here must not
    # take the headline line with it
    train_rec, c1_rec = None, None
    if not args.no_extra:
        try:
            train_rec = train_record(rank, world, dev)
        except Exception as e:  # noqa: BLE001
            train_rec = {"error": repr(e)[:300]}
        if rank == 0:
            try:
                c1_rec = c1_record(dev, graph=(world == 1))   # no capture while other ranks wait in a collective
            except Exception as e:  # noqa: BLE001
                c1_rec = {"error": repr(e)[:300]}
    line = None
    if rank == 0:
        line = {"metric": "facets/sec (denoise inference, fp32)", "value": value, "unit": "facets/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "layers": layers,
                "cpu_baseline": cpu_baseline, "layer": layer_rec, "train": train_rec, "single_mesh": c1_rec,
                "rows": {"core_facets": total_core, "level0_rows_per_pass": total_rows0,
                         "host_patch_generation_s": t_gen, "patches_this_rank": len(mine)}}
    if world > 1:
        dist.barrier()
    return line



# ----------------------------------------------------------------------------- sub-records: C4 training step, C1 latency
def train_record(rank, world, dev, steps=10, warmup=3, batch_size=4):
    """C4 (BASELINE.json configs[3]): data-parallel training step of the multi-scale network on 8 192-facet patches
    (`Code/train.py:493-520, 619`): random rotation, forward, faceNormalsLoss on sampled facets, deterministic backward,
    ONE all-reduce of the flat gradient bucket over NCCL, Adam.  Weak scaling: `batch_size` patches per rank.  Every rank
    calls this (the all-reduce is a collective); CUDA-event time, max over ranks."""
    import torch
    import torch.distributed as dist
    from facet_graph_convolution_b200 import _lib, patches
    from facet_graph_convolution_b200 import model as fm
    from facet_graph_convolution_b200 import train as ftrain
    L = _lib.lib()
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nq = 64   # 64 x 64 quads = 8 192 triangles per patch
    batch = []
    for bi in range(batch_size):
        P, _ = patches.grid_patches(nq, nq, block=nq, halo=0, K=NET_K, seed=rank * 100 + bi)
        p = P[0]
        gt = np.zeros((p.x.shape[0], 3), np.float32)
        gt[: p.num_real] = p.x[: p.num_real, :3]
        batch.append((T(p.x[None]), [T(a[None]) for a in p.adjs], T(gt[None])))
    net = fm.DenoisingNet(6, device=dev, params=net_params())
    bucket = ftrain.GradBucket(list(net.parameters()))
    opt = ftrain.Adam(bucket)
    rng = np.random.RandomState(rank)
    group = dist.group.WORLD if world > 1 else None
    for _ in range(warmup):
        ftrain.train_step(net, batch, bucket, opt, rng, group)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = L.fgc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ftrain.train_step(net, batch, bucket, opt, rng, group)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    facets = batch_size * 2 * nq * nq * world
    return {"workload": "C4 training step: %d patches/GPU x 8192 level-0 facets, K=%d, M=%d, random rotation, faceNormalsLoss on "
                        "sampled facets, fwd + deterministic bwd, one NCCL all-reduce of the flat gradient bucket (%d floats = "
                        "%.2f MB), Adam" % (batch_size, NET_K, NET_M, bucket.flat.numel(), bucket.flat.numel() * 4 / 1e6),
            "facets_per_s": facets / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": warmup, "scaling": "weak",
            "parallelism": "dp%d" % world, "gpu_launches_per_step": (L.fgc_launch_count() - n0) / steps}


def c1_record(dev, steps=20, warmup=5, graph=True):
    """C1 (BASELINE.json configs[0]): one ~20k-facet mesh (icosphere-5, 20 480 facets, one patch, B = 1) through the
    network + normalizeTensor and the 60-sweep vertex update (`Code/train.py:100-136, 1467-1557`): latency per mesh."""
    import torch
    from facet_graph_convolution_b200 import mesh
    from facet_graph_convolution_b200 import model as fm
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    V, F = mesh.icosphere(5)
    Vn = mesh.add_vertex_noise(V, F, 0.3, 0).astype(np.float32)
    feat = mesh.face_features(Vn, F).astype(np.float32)
    feat, adj0 = mesh.pad_to_multiple(feat, mesh.dedup_adj(mesh.faces_large_adj(F, NET_K)), 16)
    adjs = mesh.build_pyramid(adj0, 3, NET_K)
    e_map, v_e = mesh.edge_maps(F, 20)
    nreal = F.shape[0]
    x_d, adjs_d = T(feat[None]), [T(a[None]) for a in adjs]
    v_d, em_d, ve_d = T(Vn[None]), T(e_map[None]), T(v_e[None])
    store = fm.VariableStore(dev, params=net_params())

    def fwd_only():
        with torch.no_grad(), fm.variable_store(store):
            return fm.normalizeTensor(fm.get_model_reg_multi_scale(x_d, adjs_d, 1.0))

    def vertex(n):
        with torch.no_grad():
            return fm.update_position2(v_d, n[:, :nreal].contiguous(), em_d, ve_d, iter_num=60, max_edges=20)

    for _ in range(warmup):
        vertex(fwd_only())
    torch.cuda.synchronize()
    tf, tv = [], []
    for _ in range(steps):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        n = fwd_only()
        e1.record()
        vertex(n)
        e2.record()
        torch.cuda.synchronize()
        tf.append(e0.elapsed_time(e1))
        tv.append(e1.elapsed_time(e2))
    ms_f, ms_v = float(np.median(tf)), float(np.median(tv))
    # the same step (about 150 launches: network, normalise, 60 sweeps) captured once and replayed from a CUDA graph
    graph_ms = None
    try:
        if not graph:
            raise RuntimeError("skipped")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            vertex(fwd_only())
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            x_g = vertex(fwd_only())
        x_e = vertex(fwd_only())
        g.replay()
        torch.cuda.synchronize()
        if torch.equal(x_g, x_e):
            tg = []
            for _ in range(steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                tg.append(e0.elapsed_time(e1))
            graph_ms = float(np.median(tg))
    except Exception:  # noqa: BLE001
        graph_ms = None
    rec_extra = {"ms_step_eager": ms_f + ms_v, "ms_step_cuda_graph": graph_ms}
    return {**rec_extra, "workload": "C1 single-mesh denoise: icosphere-5 (%d facets, one patch, B=1), network + normalizeTensor, then 60 "
                        "sweeps of update_position2" % nreal,
            "facets_per_s_forward": nreal / (ms_f * 1e-3), "ms_forward": ms_f, "ms_vertex_update_60_sweeps": ms_v,
            "steps": steps, "warmup": warmup}


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c2"])
    ap.add_argument("--grid", type=int, default=1000, help="c3: quads per side (1000 -> 2 M facets)")
    ap.add_argument("--block", type=int, default=100, help="c3: quads per side of a patch core")
    ap.add_argument("--patch-batch", type=int, default=100, help="c3: patches per launch (1 = the reference's B = 1)")
    ap.add_argument("--streams", type=int, default=2, help="c3: launch groups of a pass alternate over this many CUDA streams")
    ap.add_argument("--adjacency", default="mesh", choices=["mesh", "dedup", "random"], help="c2")
    ap.add_argument("--facets", type=int, default=1_000_000, help="c2")
    ap.add_argument("--cpu-sample", type=int, default=20_000, help="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-split", default="auto", help="c3: cumulative percentages at which the end-to-end pass cuts a "
                    "rank's patches into upload groups (10,40 -> groups of about 1/10, 3/10 and 6/10; auto: 8,30,65 from 80 "
                    "patches per rank, else 10,40)")
    ap.add_argument("--no-layer", action="store_true", help="c3: skip the C2 layer record")
    ap.add_argument("--no-extra", action="store_true", help="c3: skip the C4 training-step and C1 single-mesh records")
    args = ap.parse_args()
    line = run_c2(args) if args.config == "c2" else run_c3(args)
    if line is not None:
        print(json.dumps(line))
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
