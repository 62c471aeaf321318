#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv) per kernel: launches, ms, share of the step, DRAM GB,
GB/s, time-weighted tensor-pipe activity.  Optionally writes profiles/halo_traffic.json (the `roofline.traffic` figure
bench.py reports: DRAM bytes per launch of the dominant kernel).

    python tools/ncu_summary.py gpurun_out/launches.csv --title "..." --out profiles/r2_launch_summary.txt \
        --traffic-kernel conv_fwd_halo_kernel --traffic-out profiles/halo_traffic.json --round 2
"""
import argparse
import collections
import csv
import json
import re


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("pmfb::", "")
    return name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="")
    ap.add_argument("--out", default=None)
    ap.add_argument("--traffic-kernel", default=None)
    ap.add_argument("--traffic-out", default=None)
    ap.add_argument("--round", type=int, default=2)
    ap.add_argument("--workload", default="one training step, B=8, 480x640")
    a = ap.parse_args()
    rows = collections.OrderedDict()
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        k = rows.setdefault(r["ID"], {"name": r["Kernel Name"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("gpu__time_duration"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        k[m] = v
    agg = collections.OrderedDict()
    for k in rows.values():
        a_ = agg.setdefault(short(k["name"]), {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0, "tp": 0.0})
        ms = k.get("gpu__time_duration.sum", 0.0)
        a_["n"] += 1
        a_["ms"] += ms
        a_["rd"] += k.get("dram__bytes_read.sum", 0.0)
        a_["wr"] += k.get("dram__bytes_write.sum", 0.0)
        tp = [v for m, v in k.items() if m.startswith("sm__pipe_tensor")]
        a_["tp"] += (tp[0] if tp else 0.0) * ms
    total = sum(v["ms"] for v in agg.values())
    out = ["== %s" % a.title, "   source: %s; cold-cache serialised replays (compare SHARES, not absolutes).  total %.3f ms over %d launches"
           % (a.csv, total, sum(v["n"] for v in agg.values())),
           "kernel | launches | ms | share | DRAM read GB | DRAM write GB | GB/s | tensor pipe active % (time-weighted)"]
    for name, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if v["ms"] < 0.002 * total:
            continue
        out.append("%s | %d | %.3f | %.1f%% | %.2f | %.2f | %.0f | %.1f" % (
            name, v["n"], v["ms"], 100 * v["ms"] / total, v["rd"] / 1e9, v["wr"] / 1e9,
            (v["rd"] + v["wr"]) / 1e9 / max(v["ms"] * 1e-3, 1e-12), v["tp"] / max(v["ms"], 1e-12)))
    if a.traffic_kernel:
        sel = [v for n, v in agg.items() if a.traffic_kernel in n]
        n = sum(v["n"] for v in sel)
        b = sum(v["rd"] + v["wr"] for v in sel)
        ms = sum(v["ms"] for v in sel)
        out.append("%s (all variants): %d launches, %.2f GB DRAM traffic = %.1f MB per launch, %.3f ms (%.1f%% of the step)"
                   % (a.traffic_kernel, n, b / 1e9, b / max(n, 1) / 1e6, ms, 100 * ms / total))
        if a.traffic_out and n:
            json.dump({"kernel": "%s (all epilogue variants)" % a.traffic_kernel, "dram_bytes_per_launch": b / n, "launches": n,
                       "dram_gb_total": b / 1e9, "workload": a.workload,
                       "source": "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum, --clock-control none)" % (a.out or a.csv),
                       "round": a.round}, open(a.traffic_out, "w"))
    text = "\n".join(out) + "\n"
    if a.out:
        open(a.out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
