#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source sass,cuda` of one k_march launch by code region:
share of warp-instructions, average active lanes and stall samples.  Usage: python tools/ncu_regions.py src.csv"""
import csv
import sys


def region(f, l):
    if f == "rc_device.cuh":
        if 225 <= l <= 251:
            return "trace: node loop"
        if 140 <= l <= 175 or 252 <= l <= 257:
            return "trace: leaves / tri_test"
        if 190 <= l <= 224 or 258 <= l <= 264:
            return "trace: set-up"
        if l >= 268:
            return "shade_hit"
        return "vector helpers (rc_device.cuh < 140)"
    if f == "kernels.cu":
        return "kernels.cu:%d" % (l // 10 * 10)
    return str(f)


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [r for r in rows if "Instructions Executed" in r][0]
    ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    cur, line, seen = None, -1, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0].strip().isdigit():
            line = int(r[0])
            continue
        if len(r) <= ti or not r[2].startswith("0x") or r[2] in seen:
            continue
        try:
            seen[r[2]] = (int(r[ii]), int(r[ti]), int(r[si]), cur, line)
        except ValueError:
            pass
    tot_i = sum(v[0] for v in seen.values()); tot_t = sum(v[1] for v in seen.values()); tot_s = sum(v[2] for v in seen.values())
    print("warp-instructions %d, active lanes %.2f, samples %d" % (tot_i, tot_t / tot_i, tot_s))
    agg = {}
    for i, t, s, f, l in seen.values():
        a = agg.setdefault(region(f, l), [0, 0, 0]); a[0] += i; a[1] += t; a[2] += s
    for k, (i, t, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:28]:
        print("%-40s inst %5.1f%%  lanes %5.1f  samples %5.1f%%" % (k, 100 * i / tot_i, t / max(i, 1), 100 * s / tot_s))


if __name__ == "__main__":
    main(sys.argv[1])
