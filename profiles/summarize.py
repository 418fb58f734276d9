#!/usr/bin/env python
"""Turns the raw ncu exports kept in this directory into the two files the repo reads:

    python profiles/summarize.py <tag>        # e.g. r1b

  <tag>_ncu_full_raw.csv  (ncu -i <rep> --page raw --csv of the `--set full` capture)
  <tag>_launches.csv      (ncu --metrics gpu__time_duration.sum launch list of the same command)
->
  traffic.json            per-kernel DRAM bytes per launch (bench.py's roofline.traffic)
  <tag>_summary.md        per-kernel table: time share, DRAM traffic, pipe utilisation, stalls
"""
import collections
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TIME_MS = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
SHORT = ["conv_fwd_tc_kernel", "bwd_src_tc_kernel", "bwd_tgt_tc_kernel", "bwd_w_tc_kernel", "conv_fwd_kernel",
         "bwd_src_kernel", "bwd_tgt_kernel", "bwd_w_kernel", "fconv_fwd_kernel", "fconv_tgt_kernel"]


MMA_ORDER = ["conv_mma_kernel", "bwd_tgt_mma_kernel", "bwd_w_mma_kernel"]   # launch order inside one fwd+bwd step
_mma_seen = [0]


def short_name(full):
    if "conv_mma_kernel" in full:
        # one kernel template serves the forward, the target-centric pass and the weight gradient
        # (MmaParams::mode); the launches of a step come in this fixed order
        nm = MMA_ORDER[_mma_seen[0] % 3]
        _mma_seen[0] += 1
        return nm
    if "bwd_src_mma_kernel" in full:
        return "bwd_src_mma_kernel"
    if "conv_fwd_tc_kernel" in full and full.rstrip().split(",")[-1].strip().startswith("1"):
        return "bwd_tgt_tc_kernel"      # conv_fwd_tc_kernel<M, COUT, MODE_TGT>
    if "conv_fwd_tc_kernel" in full and "(int)1>" in full:
        return "bwd_tgt_tc_kernel"
    for s in SHORT:
        if s in full:
            return s
    return full.split("(")[0].split("::")[-1]


def main():
    tag = sys.argv[1]
    rows = list(csv.reader(open(os.path.join(HERE, tag + "_ncu_full_raw.csv"))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale_bytes=False):
        if name not in idx:
            return None
        v = r[idx[name]].replace(",", "")
        try:
            f = float(v)
        except ValueError:
            return None
        if scale_bytes:
            f *= UNIT.get(units[idx[name]], 1.0)
        return f

    kernels = collections.OrderedDict()
    for r in rows[2:]:
        full = r[idx["Kernel Name"]]
        name = short_name(full if "(int)" in full else full + "," + r[idx["Kernel Name"]])
        # template MODE is not always printed in the raw page; disambiguate by order: 2nd conv_fwd_tc = TGT
        if name == "conv_fwd_tc_kernel" and name in kernels:
            name = "bwd_tgt_tc_kernel"
        kernels[name] = {
            "ncu_ms": (val(r, "gpu__time_duration.sum") or 0.0) * TIME_MS.get(units[idx["gpu__time_duration.sum"]], 1.0),
            "dram_bytes_per_launch": (val(r, "dram__bytes_read.sum", True) or 0) + (val(r, "dram__bytes_write.sum", True) or 0),
            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"),
            "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
            "regs": val(r, "launch__registers_per_thread"),
            "warp_insts": val(r, "smsp__inst_executed.sum"),
            "stall_long_scoreboard": val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        }
    shares = collections.OrderedDict()
    _mma_seen[0] = 0
    lpath = os.path.join(HERE, tag + "_launches.csv")
    if os.path.exists(lpath):
        lr = list(csv.reader(open(lpath)))
        h = next(i for i, r in enumerate(lr) if "Kernel Name" in r)
        li = {k: i for i, k in enumerate(lr[h])}
        for r in lr[h + 1:]:
            if len(r) < len(lr[h]) or r[li["Metric Name"]] != "gpu__time_duration.sum":
                continue
            full = r[li["Kernel Name"]]
            nm = "bwd_tgt_tc_kernel" if ("conv_fwd_tc_kernel" in full and full.split("(")[0].rstrip(">").endswith(" 1")) else short_name(full)
            v = float(r[li["Metric Value"]].replace(",", ""))
            v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(r[li["Metric Unit"]], 1e-6)
            a = shares.setdefault(nm, [0, 0.0])
            a[0] += 1
            a[1] += v
    json.dump({"source": tag + "_ncu_full_raw.csv (ncu --set full --clock-control none, one launch each, bench.py C2 workload)",
               "kernels": {k: {"dram_bytes_per_launch": v["dram_bytes_per_launch"]} for k, v in kernels.items()}},
              open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    tot = sum(a[1] for a in shares.values()) or 1.0
    with open(os.path.join(HERE, tag + "_summary.md"), "w") as f:
        f.write("# ncu summary `%s` (C2 workload: N=1M facets, K=16, M=8, 64->64, fwd+bwd)\n\n" % tag)
        f.write("`--set full` capture, one launch per kernel (cold-cache, serialised: compare shares, not absolutes).\n\n")
        f.write("| kernel | ncu ms | DRAM MB/launch | DRAM % | tensor pipe % | fma pipe % | issue active % | warps active % | L1 hit % | L2 hit % | regs | warp-insts | long-scoreboard stall |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k, v in kernels.items():
            f.write("| %s | %.3f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.0fM | %.2f |\n" % (
                k, v["ncu_ms"], v["dram_bytes_per_launch"] / 1e6, v["dram_pct"], v["tensor_pipe_pct"], v["fma_pipe_pct"],
                v["issue_active_pct"], v["warps_active_pct"], v["l1_hit_pct"], v["l2_hit_pct"], v["regs"],
                v["warp_insts"] / 1e6, v["stall_long_scoreboard"]))
        if shares:
            f.write("\nLaunch list (`%s_launches.csv`, `--metrics gpu__time_duration.sum`): share of the summed kernel time\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n" % tag)
            for k, a in shares.items():
                f.write("| %s | %d | %.3f | %.1f%% |\n" % (k, a[0], a[1], 100 * a[1] / tot))
    print(open(os.path.join(HERE, tag + "_summary.md")).read())


if __name__ == "__main__":
    main()
