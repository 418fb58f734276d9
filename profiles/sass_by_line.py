#!/usr/bin/env python
"""Joins an `ncu --page source --csv` SASS table with `nvdisasm -g` line info and prints executed
warp-instructions / stall samples per CUDA source line (top N).

    cuobjdump -xelf all libfacetconv_b200.so   # -> conv_fwd_tc.sm_100a.cubin ...
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > sass.csv
    python profiles/sass_by_line.py sass.csv conv_fwd_tc.sm_100a.cubin '<mangled-name-substring>' [N]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    sass_csv, cubin, func = sys.argv[1:4]
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # walk the function's text section
    lines, cur, infunc = [], None, False
    for ln in dis:
        if ln.startswith(".text.") or ln.strip().startswith(".section"):
            infunc = func in ln and ".text." in ln
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((cur, m.group(2).strip()))
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr)]
    if len(body) != len(lines):
        print("warning: %d ncu rows vs %d disassembled instructions" % (len(body), len(lines)))
    by = collections.defaultdict(lambda: [0, 0])
    tot = totS = 0
    for (loc, _), r in zip(lines, body):
        n = int(r[idx["Instructions Executed"]])
        s = int(r[idx["# Samples"]])
        by[loc][0] += n
        by[loc][1] += s
        tot += n
        totS += s
    print("total warp-instructions %d, samples %d" % (tot, totS))
    for loc, (n, s) in sorted(by.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("%-22s %12d %5.1f%%   samples %7d %5.1f%%" % ("%s:%s" % loc if loc else "?", n, 100.0 * n / tot, s,
                                                          100.0 * s / max(totS, 1)))


if __name__ == "__main__":
    main()
