#!/usr/bin/env python
"""Condense `ncu --set full ... --page raw --csv` exports (tools/ncu_capture.sh) into the handful of metrics the roofline
discussion uses: duration, DRAM bytes, tensor-pipe activity, achieved occupancy, registers, L2 hit rate, shared-memory
wavefronts.

    python tools/ncu_full_summary.py gpurun_out/r2_halo_raw.csv gpurun_out/r2_wgrad_raw.csv > profiles/r2_ncu_full_summary.txt
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__shared_mem_per_block_dynamic", "smem KB")]


def main():
    print("ncu --set full --clock-control none --import-source on (tools/ncu_capture.sh), one eagerly launched training step, B=8, "
          "480x640, default f16 mode.  Per captured launch:")
    for f in sys.argv[1:]:
        rows = list(csv.reader(open(f)))
        hdr = rows[0]
        idx = {h: i for i, h in enumerate(hdr)}
        print("\n== %s" % f)
        print("kernel | " + " | ".join(n for _, n in COLS))
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
            vals = []
            for c, _ in COLS:
                v = r[idx[c]] if c in idx else ""
                try:
                    vals.append("%.1f" % float(v.replace(",", "")))
                except ValueError:
                    vals.append(v[:12])
            print("%s | %s" % (name, " | ".join(vals)))


if __name__ == "__main__":
    main()
