#!/usr/bin/env python
"""Condense `ncu -i <rep> --page raw --csv` into the per-launch summary table kept under profiles/.
Usage: python tools/ncu_summary.py raw.csv > summary.csv"""
import csv
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    idx = [names.index(c) for c in COLS if c in names]
    w = csv.writer(sys.stdout)
    w.writerow([names[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        out = [r[i] for i in idx]
        out[0] = out[0].split("(")[0].replace("rc::<unnamed>::", "").replace("void ", "")
        w.writerow(out)


if __name__ == "__main__":
    main(sys.argv[1])
