#!/usr/bin/env python
"""Top stalled SASS instructions per kernel from `ncu --page source --csv --print-source sass` (gz or plain).
usage: ncu_top_sass.py file.csv.gz [kernel_index] [top_n]"""
import csv, gzip, sys
p = sys.argv[1]
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
f = gzip.open(p, "rt") if p.endswith(".gz") else open(p)
kernels, cur = [], None
for row in csv.reader(f):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
k = kernels[ki]
h = k["hdr"]
si = h.index("# Samples")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[si] or 0) for r in k["rows"])
print("kernel %d/%d %s  instructions %d  samples %d" % (ki, len(kernels), k["name"][:60], len(k["rows"]), tot))
agg = {}
for r in k["rows"]:
    for i in stall_cols:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print("stalls:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / max(tot, 1)) for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(k["rows"])), key=lambda j: -int(k["rows"][j][si] or 0))[:top]
for j in sorted(order):
    r = k["rows"][j]
    st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s %s" % (j, 100.0 * int(r[si] or 0) / max(tot, 1), r[1].strip()[:70], " ".join("%s:%d" % (n, v) for v, n in st if v)))
