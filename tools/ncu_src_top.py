#!/usr/bin/env python
"""Top SASS instructions of an `ncu --page source --csv --print-source sass` export (tools/ncu_micro.sh / ncu_capture.sh)
by stall samples, with shared-memory wavefronts and executed counts.   python tools/ncu_src_top.py <src_sass.csv.gz> [N]"""
import csv
import gzip
import sys


def main():
    rows = list(csv.reader(gzip.open(sys.argv[1], "rt")))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    start = 0
    while start < len(rows):
        if rows[start][0] != "Kernel Name":
            start += 1
            continue
        hdr = rows[start + 1]
        idx = {h: i for i, h in enumerate(hdr)}
        end = start + 2
        while end < len(rows) and rows[end][0] != "Kernel Name":
            end += 1
        body = rows[start + 2:end]
        print("==", rows[start][1][:100])
        tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
        print("total samples", tot)
        stall_cols = [h for h in hdr if h.startswith("stall_")]
        order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:n]
        for i in sorted(order):
            r = body[i]
            st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
            print("%5d %6s smp %9s exe  shW %9s/%9s  %-60s %s" % (i, r[idx["# Samples"]], r[idx["Instructions Executed"]],
                                                                 r[idx["L1 Wavefronts Shared"]], r[idx["L1 Wavefronts Shared Ideal"]],
                                                                 r[idx["Source"]].strip()[:60], st))
        start = end


if __name__ == "__main__":
    main()
