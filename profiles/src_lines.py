#!/usr/bin/env python
"""Top source lines by warp-stall samples from an `ncu --page source --csv --print-source cuda,sass` export.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> --launch-count 1 > src.csv
    python profiles/src_lines.py src.csv [N] [file-filter]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    filt = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(path)))
    agg = collections.OrderedDict()
    fileof = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fileof = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] == "Line No" or r[0] == "":
            continue
        try:
            key = (fileof, int(r[0]), r[1][:100])
            agg[key] = [int(r[4] or 0), int(r[5] or 0), int(r[7] or 0)]
        except ValueError:
            continue
    tot = sum(v[0] for v in agg.values()) or 1
    print("total stall samples", tot)
    items = [(k, v) for k, v in agg.items() if filt is None or filt in k[0]]
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:topn]:
        print("%5.1f%% samples=%7d not-issued=%7d inst=%9d  %s:%d  %s" % (100 * v[0] / tot, v[0], v[1], v[2], k[0], k[1], k[2]))


if __name__ == "__main__":
    main()
