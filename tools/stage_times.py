#!/usr/bin/env python
"""Quick A/B probe: mean CUDA-event stage times (ms) of a workload's frame along the orbit, L2 flushed between frames.
  python tools/stage_times.py living_room_4k [--frames 12] [--set key=value ...] [--levels]
Tuning knobs are passed with --set (rc_set_tuning) or through RC_* environment variables (read at rc_create)."""
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
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--set", action="append", default=[])
    ap.add_argument("--levels", action="store_true")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    import torch
    import radiancecascade_b200 as rc
    name, W, H, lk = bench.WORKLOADS[a.workload]
    st = rc.AppState()
    r = rc.DefaultRenderer.new(0, (W, H), st, rc.scenes.scene_path(name))
    for kv in a.set:
        k, v = kv.split("=")
        r.set_tuning(k, int(v))
    info = r.scene_info()
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows, frame_ms = [], []
    for i in range(a.frames + 3):
        uc, pts = bench.frame_inputs(rc, info, W, H, i, lk)
        st.uniform_camera = uc
        st.light_position, st.extra_lights = pts[0], pts[1:]
        r.update(st)
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r.render(stream.cuda_stream)
            e1.record(stream)
        stream.synchronize()
        if i >= 3:
            rows.append(r.stage_times())
            frame_ms.append(e0.elapsed_time(e1))
    out = {k: round(float(np.mean([x[k] for x in rows])), 4) for k in rows[0]}
    out["event_ms"] = round(float(np.mean(frame_ms)), 4)
    if a.levels:
        r.set_tuning("level_timing", 1)
        lv = []
        for i in range(4):
            with torch.cuda.stream(stream):
                flush.zero_()
                r.render(stream.cuda_stream)
            stream.synchronize()
            lv.append(r.level_times())
        out["level_ms"] = [round(float(np.mean([x[i] for x in lv[1:]])), 4) for i in range(len(lv[0]))]
    print(json.dumps({"workload": a.workload, "tag": a.tag, "set": a.set,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("RC_")}, **out}))


if __name__ == "__main__":
    main()
