#!/usr/bin/env python
"""Per-rank timeline of a tiled frame, measured on ONE GPU: renders each tile of an N-way partition of the frame as its own
context (exactly what rank r of a tiled multi-GPU run executes) and prints the CUDA-event stage times per tile.
  python tools/tile_times.py living_room_4k --grid 1x8 [--frames 6] [--set key=value ...]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--grid", default="1x8", help="NXxNY tiles")
    ap.add_argument("--frames", type=int, default=6)
    ap.add_argument("--set", action="append", default=[])
    ap.add_argument("--balanced", action="store_true", help="cost-balanced cuts (distributed.partition_balanced)")
    a = ap.parse_args()
    import torch
    import radiancecascade_b200 as rc
    from radiancecascade_b200 import distributed as rd
    name, W, H, lk = bench.WORKLOADS[a.workload]
    nx, ny = (int(v) for v in a.grid.lower().split("x"))
    tiles = rd.partition_grid(W, H, nx, ny)
    st = rc.AppState()
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for t in tiles:
        r = rc.DefaultRenderer.new(0, (W, H), st, rc.scenes.scene_path(name), rc.CascadeConfig(tile=t))
        for kv in a.set:
            k, v = kv.split("=")
            r.set_tuning(k, int(v))
        info = r.scene_info()
        acc = []
        for i in range(a.frames + 3):
            uc, pts = bench.frame_inputs(rc, info, W, H, i, lk)
            st.uniform_camera = uc
            st.light_position, st.extra_lights = pts[0], pts[1:]
            r.update(st)
            with torch.cuda.stream(stream):
                flush.zero_()
                r.render(stream.cuda_stream)
            stream.synchronize()
            if i >= 3:
                acc.append(r.stage_times())
        m = {k: round(float(np.mean([x[k] for x in acc])), 4) for k in acc[0]}
        marched = [x for x in r.rays_marched() if x]
        m["rays"] = int(sum(marched))
        m["tile"] = list(t)
        rows.append(m)
        del r
    out = {"workload": a.workload, "grid": a.grid, "set": a.set, "tiles": rows,
           "max_frame_ms": max(x["frame"] for x in rows), "sum_rays": sum(x["rays"] for x in rows)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
