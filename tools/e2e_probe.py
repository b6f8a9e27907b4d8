"""Where does the end-to-end frame time go?  Host-side timers around each call of bench.py's e2e loop."""
import os, sys, time
import ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import radiancecascade_b200 as rc
from radiancecascade_b200 import _ffi
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "teapot_1080p"
name, W, H, lk = bench.WORKLOADS[wl]
torch.cuda.set_device(0)
state = rc.AppState()
r = rc.DefaultRenderer.new(0, (W, H), state, rc.scenes.scene_path(name))
info = r.scene_info()
stream = torch.cuda.Stream(); sh = stream.cuda_stream
hosts = [torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True) for _ in range(2)]
nbytes = hosts[0].numel() * 2

def set_frame(i):
    uc, pts = bench.frame_inputs(rc, info, W, H, i, lk)
    state.uniform_camera = uc
    state.light_position, state.extra_lights = pts[0], pts[1:]
    r.update(state)

N = 40
side = torch.cuda.Stream()
junk = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
junk_host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
for mode in ("pipelined", "render_only_nosync", "pipelined_nowait"):
    T = dict(set=0.0, render=0.0, read=0.0, wait=0.0)
    for i in range(3):
        set_frame(i); r.render(sh)
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    prev = None
    for i in range(N):
        t0 = time.perf_counter(); 
        if mode != "copy_only": set_frame(i)
        t1 = time.perf_counter()
        if mode != "copy_only": r.render(sh)
        t2 = time.perf_counter()
        if mode in ("pipelined", "copy_only", "pipelined_nowait"):
            tk = r.read_irradiance_async(hosts[i & 1].data_ptr(), nbytes)
            t3 = time.perf_counter()
            if prev is not None and mode != "pipelined_nowait": r.read_wait(prev)
            prev = tk
        elif mode == "render_only_sync_each":
            t3 = time.perf_counter(); stream.synchronize()
        elif mode == "nosync_plus_unrelated_d2h":
            t3 = time.perf_counter()
            with torch.cuda.stream(side):
                junk_host.copy_(junk, non_blocking=True)
        elif mode == "nosync_plus_unrelated_h2d":
            t3 = time.perf_counter()
            with torch.cuda.stream(side):
                junk.copy_(junk_host, non_blocking=True)
        else:
            t3 = time.perf_counter()
        t4 = time.perf_counter()
        T["set"] += t1 - t0; T["render"] += t2 - t1; T["read"] += t3 - t2; T["wait"] += t4 - t3
    if prev is not None: r.read_wait(prev)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t_start) * 1e3 / N
    print(mode, "ms/frame %.4f" % tot, {k: round(v * 1e3 / N, 4) for k, v in T.items()}, "device stages", {k: round(v, 4) for k, v in r.stage_times().items()})
