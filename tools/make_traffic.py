#!/usr/bin/env python
"""Per-launch DRAM traffic of the frame's kernels from `ncu -i <rep> --page raw --csv` exports (tools/profile_round.sh):
dram__bytes_read.sum + dram__bytes_write.sum, averaged over the k_march launches of one frame, and for k_gather.
Usage: python tools/make_traffic.py <tag> > profiles/<tag>_traffic.json     (reads gpurun_out/raw_<tag>_<workload>.csv)"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def per_kernel(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    k, rd, wr = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
    out = {}
    for r in rows[hdr + 2:]:
        if len(r) <= max(rd, wr):
            continue
        name = r[k].split("(")[0].split("::")[-1].split("<")[0]
        b = float(r[rd].replace(",", "")) * UNIT[units[rd]] + float(r[wr].replace(",", "")) * UNIT[units[wr]]
        out.setdefault(name, []).append(b)
    return out


def main(tag):
    res = {"source": f"profiles/{tag}_ncu_full_summary_<workload>.csv (ncu --set full --clock-control none, one frame, RC_GRAPH=0 for the capture)",
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the frame's k_march launches; at 1080p the cascade "
                   "stays in the 126 MB L2 between the kernels that write and read it, so DRAM traffic is below the algorithmic bytes"}
    for wl in ("teapot_1080p", "living_room_4k"):
        pk = per_kernel(f"gpurun_out/raw_{tag}_{wl}.csv")
        march = pk.get("k_march", [])
        gather = pk.get("k_gather_mma", pk.get("k_gather_pipe", pk.get("k_gather", [])))
        res[wl] = {"k_march_dram_bytes_per_launch": sum(march) / max(1, len(march)), "k_march_launches": len(march),
                   "k_gather_dram_bytes_per_launch": sum(gather) / max(1, len(gather)),
                   "k_gbuffer_dram_bytes_per_launch": sum(pk.get("k_gbuffer", [0])) / max(1, len(pk.get("k_gbuffer", [0])))}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
