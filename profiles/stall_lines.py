#!/usr/bin/env python
"""Per source line: warp-stall samples split by stall reason, from
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
Usage: python profiles/stall_lines.py src.csv [N] [lo-hi line range of conv file to total]"""
import csv
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rng = tuple(int(v) for v in sys.argv[3].split("-")) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
hdr = None
fileof = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fileof = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = agg.setdefault((fileof, ln), {"src": r[1][:70], "inst": 0, "samples": 0})
    def num(v):
        try:
            return int(v or 0)
        except ValueError:
            return 0
    d["inst"] += num(r[7])
    d["samples"] += num(r[4])
    d["wf_sh"] = d.get("wf_sh", 0) + num(r[19])
    d["tag_g"] = d.get("tag_g", 0) + num(r[16])
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and i < len(r):
            d[h] = d.get(h, 0) + num(r[i])
tot = sum(d["samples"] for d in agg.values()) or 1
keys = sorted({k for d in agg.values() for k in d if k.startswith("stall_")})
print("total samples", tot)
sums = {k: sum(d.get(k, 0) for d in agg.values()) for k in keys}
print("all lines:", {k[6:]: v for k, v in sorted(sums.items(), key=lambda kv: -kv[1]) if v})
if rng:
    sel = [d for (f, l), d in agg.items() if f and f.startswith("conv_hm") and rng[0] <= l <= rng[1]]
    s2 = {k: sum(d.get(k, 0) for d in sel) for k in keys}
    print("lines %d-%d: samples %d inst %d" % (rng[0], rng[1], sum(d["samples"] for d in sel), sum(d["inst"] for d in sel)),
          {k[6:]: v for k, v in sorted(s2.items(), key=lambda kv: -kv[1]) if v})
for (f, l), d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    top = sorted(((k[6:], v) for k, v in d.items() if k.startswith("stall_") and v), key=lambda kv: -kv[1])[:4]
    print("%5.1f%% inst=%9d wfsh=%8d tagg=%8d %s:%d %s | %s" % (100.0 * d["samples"] / tot, d["inst"], d.get("wf_sh", 0), d.get("tag_g", 0), f, l, d["src"], top))
print("shared wavefronts by line:")
for (f, l), d in sorted(agg.items(), key=lambda kv: -kv[1].get("wf_sh", 0))[:8]:
    print("   wfsh=%9d %s:%d %s" % (d.get("wf_sh", 0), f, l, d["src"]))
print("global tag requests by line:")
for (f, l), d in sorted(agg.items(), key=lambda kv: -kv[1].get("tag_g", 0))[:10]:
    print("   tagg=%9d %s:%d %s" % (d.get("tag_g", 0), f, l, d["src"]))
