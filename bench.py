#!/usr/bin/env python
"""bench.py — headline benchmark of the radiance-cascade GI hot path on B200.

A "step" is one GI frame (G-buffer, probe placement, per-level ray march fused with the
cascade merge, irradiance gather) of a bundled scene along the deterministic orbit camera
of SURVEY.md §8d.  Metric: G ray-samples/s (one ray sample = one (probe, direction, level)
interval query) with ms/frame in `ms_per_step`.

  python bench.py --gpus N --steps K --warmup W            # product (CUDA, C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the same spec

N = 1 workload: BASELINE.json configs[1] — teapot, 1920x1080, single light.  N > 1: every rank
renders its own frames of the orbit (multi-view batch, no data-path collective): weak scaling.
The reference itself has no GI path and cannot be built here, so the CPU
arm is the repo's C oracle (kind "port") on all host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gi_ray_samples_per_s"
UNIT = "Gray-samples/s"
WORKLOADS = {  # BASELINE.json configs
    "teapot_1080p": ("teapot", 1920, 1080, "bench"),
    "test_room_1080p": ("test_room", 1920, 1080, "room"),
    "living_room_4k": ("living_room", 3840, 2160, "bench"),
    "sonic_8k": ("sonic", 7680, 4320, "bench"),
    "cube_512": ("cube", 512, 512, "bench"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def frame_inputs(rc, info, W, H, frame, lights_kind):
    lo, hi = list(info.bbox_min), list(info.bbox_max)
    pos, tgt, zn, zf = rc.scenes.orbit_camera(lo, hi, frame)
    proj = rc.Projection.new(W, H, 45.0, zn, zf)
    uc = rc.UniformCamera.look_at(pos, tgt, proj)
    pts = rc.scenes.room_lights(lo, hi) if lights_kind == "room" else [rc.scenes.bench_light(lo, hi)]
    return uc, pts


def rays_per_frame(levels):
    return int(sum(l.texel_count for l in levels))


def algorithmic_bytes(levels, W, H):
    """every level written once, every level but the top read once by the level
    below, level 0 read by the gather, plus G-buffer (depth 4 + normal 4) read and irradiance (8) written."""
    T = [int(l.texel_count) for l in levels]
    casc = 8 * (sum(T) + sum(T[1:]) + T[0])
    return casc, casc + W * H * (8 + 8)


# ------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(workload, steps, warmup, sample_div):
    """Times the C oracle (all host threads) on a bounded sample: the same scene, camera path,
    lights and cascade parameters at 1/sample_div of the resolution per axis."""
    import radiancecascade_b200 as rc
    from oracle import gi_oracle as go
    name, W, H, lk = WORKLOADS[workload]
    w, h = W // sample_div, H // sample_div
    osc = go.OracleScene(rc.scenes.scene_path(name))
    go.set_num_threads(go.host_cores())     # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)

    class _Info:
        bbox_min, bbox_max = osc.bbox_min, osc.bbox_max
    p = osc.params(w, h, store_half=True)
    lv = osc.levels(p)
    rays = sum(l.gw * l.gh * l.D * l.D for l in lv)
    times = []
    for i in range(warmup + steps):
        uc, pts = frame_inputs(rc, _Info, w, h, i, lk)
        larr = np.array([[q[0], q[1], q[2], 1.0] for q in pts], dtype=np.float32)
        t = time.perf_counter()
        osc.render(p, uc.as_array(), larr)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return {"value": rays / (ms * 1e-3) / 1e9, "unit": UNIT, "cores": go.num_threads(), "kind": "port",
            "sample": f"{name} {w}x{h} (1/{sample_div} of {W}x{H} per axis), full cascade stack, "
                      f"{steps} frame(s), oracle/rc_oracle.c with OpenMP",
            "ms_per_step": ms, "rays_per_step": rays}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload or "teapot_1080p"
    name, W, H, _ = WORKLOADS[wl]
    cb = cpu_oracle_run(wl, args.steps, args.warmup, args.cpu_sample_div)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic camera path over the bundled scene",
            "config": {"workload": f"{name} {W}x{H} (bounded sample: {cb['sample']})", "levels": 6, "probe_spacing0": 4, "dir_res0": 4},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference repository has no GI path and cannot be built here (no Rust/Vulkan); this arm is the "
                    "repo's CPU oracle of the same specification"}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa(local):
    """Multi-rank runs: keep this process (and the pinned buffers it is about to allocate, first touch) on the CPUs
    NVML names as closest to its GPU, so that 8 ranks' read-backs do not all cross to one socket.  Best effort: any
    failure leaves the affinity untouched."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs nearest to the GPU (NVML affinity)"
    except Exception as e:      # noqa: BLE001
        return f"none ({type(e).__name__})"
    return "none"


# ------------------------------------------------------------------------- product arm
def run_product(args):
    import torch
    import radiancecascade_b200 as rc
    from radiancecascade_b200 import _ffi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    binding = bind_to_gpu_numa(local) if world > 1 else "not applied (single rank)"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    wl = args.workload or "teapot_1080p"
    name, W, H, lk = WORKLOADS[wl]
    state = rc.AppState()
    flags = _ffi.RC_CFG_SEPARATE_MERGE if args.separate_merge else 0
    r = rc.DefaultRenderer.new(local, (W, H), state, rc.scenes.scene_path(name), rc.CascadeConfig(flags=flags))
    info = r.scene_info()
    levels = r.levels()
    rays = rays_per_frame(levels)
    casc_bytes, frame_bytes = algorithmic_bytes(levels, W, H)

    stream = torch.cuda.Stream()
    sh = stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    host_out = torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True)
    host_ptr, host_bytes = host_out.data_ptr(), host_out.numel() * 2
    import ctypes as C

    def set_frame(i):
        uc, pts = frame_inputs(rc, info, W, H, i * world + rank, lk)   # rank r renders its own views
        state.uniform_camera = uc
        state.light_position, state.extra_lights = pts[0], pts[1:]
        r.update(state)

    # ---- device-timed loop: per-step events, L2 flushed between steps --------------------
    def device_loop(n, record):
        evs = []
        for i in range(n):
            set_frame(i)
            with torch.cuda.stream(stream):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                r.render(sh)
                e1.record(stream)
            evs.append((e0, e1))
            if record is not None:
                stream.synchronize()
                record.append(r.stage_times())
        stream.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    device_loop(max(args.warmup, 3), None)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stages = []
    ms_steps = device_loop(args.steps, stages)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    total_ms = float(sum(ms_steps))
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * rays / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public API: host camera in, host irradiance out ----------
    # The synthetic camera path (orbit positions -> UniformCamera / UniformLight structs) is prepared before the timed
    # region, like any pre-generated input batch; every timed step still hands its own camera and lights to rc_update
    # (the step's host -> device transfer), renders, and receives the irradiance in pinned host memory.
    def packed_inputs(n):
        out = []
        for i in range(n):
            uc, pts = frame_inputs(rc, info, W, H, i * world + rank, lk)
            state.uniform_camera = uc
            state.light_position, state.extra_lights = pts[0], pts[1:]
            out.append(r.pack_update(state))
        return out

    e2e_inputs = packed_inputs(max(args.steps, 2))

    def e2e_loop(n):
        t0 = time.perf_counter()
        for i in range(n):
            r.update_packed(e2e_inputs[i])     # host -> device: camera + lights (kernel parameter block)
            r.render(sh)
            st = r._lib.rc_read_target(r._h, _ffi.RC_TARGET_IRRADIANCE, C.c_void_p(host_ptr), host_bytes)   # D2H into pinned memory
            if st != 0:
                raise RuntimeError("rc_read_target failed")
        return (time.perf_counter() - t0) * 1e3

    # same, with the library's pipelined read-back (rc_read_target_async): the copy of frame i overlaps frame i+1;
    # every frame's irradiance still lands in pinned host memory inside the timed region (the last one is waited for)
    host_out2 = torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True)
    host_ptrs = (host_ptr, host_out2.data_ptr())

    def e2e_pipelined_loop(n):
        # Frame i is enqueued BEFORE the host waits for frame i-2 (the previous user of host buffer i & 1), so the GPU
        # always has the next frame queued while the host sleeps on a copy: with the wait placed first (as in the first
        # version of this loop) the ~0.3 ms the host needs to record and submit a frame showed up as an idle gap after
        # every frame (teapot 0.62 vs 0.54 ms device time).  The library orders the device side itself: frame i's gather
        # waits for the read-back of frame i-2, which used the same irradiance buffer.
        t0 = time.perf_counter()
        tickets = [None, None]
        for i in range(n):
            r.update_packed(e2e_inputs[i])
            r.render(sh)
            if tickets[i & 1] is not None:
                r.read_wait(tickets[i & 1])    # frame i-2 is now complete in host buffer i & 1
            tickets[i & 1] = r.read_irradiance_async(host_ptrs[i & 1], host_bytes)
        for k in (n, n + 1):                   # the last two frames, oldest first
            if tickets[k & 1] is not None:
                r.read_wait(tickets[k & 1])
        return (time.perf_counter() - t0) * 1e3

    def timed(loop):
        loop(2)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        ms = loop(args.steps)
        if dist:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    e2e_sync_ms = timed(e2e_loop)
    e2e_ms = timed(e2e_pipelined_loop)
    # direction culling: texels the last frame actually marched per level (None = every texel of the level)
    marched = [int(l.texel_count) if m is None else m for l, m in zip(levels, r.rays_marched())]
    launches_per_frame = r.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    # per-level breakdown: separate pass, because the per-level events switch off the PDL overlap of the level kernels
    lv_mean = None
    if rank == 0:
        r.set_tuning("level_timing", 1)
        lv = []
        for i in range(3):
            set_frame(i)
            with torch.cuda.stream(stream):
                flush.zero_()
                r.render(sh)
            stream.synchronize()
            lv.append(r.level_times())
        lv_mean = [float(np.mean([x[i] for x in lv])) for i in range(len(levels))]
        r.set_tuning("level_timing", 0)
    # the same frames with direction culling switched off: every texel of every level is marched (the exhaustive
    # evaluation the CPU arm performs).  Reported next to the headline so that both readings of "ray samples/s" are on
    # the line: `value` = nominal cascade texels / frame time of the product path, `all_rays.value` = marched == nominal.
    r.set_tuning("cull", 0)
    device_loop(3, None)
    torch.cuda.synchronize()
    nocull_ms = float(sum(device_loop(args.steps, None))) / args.steps
    if dist:
        t = torch.tensor([nocull_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nocull_ms = float(t.item())
    r.set_tuning("cull", 1)
    e2e_value = world * rays / (e2e_ms / args.steps * 1e-3) / 1e9
    n_lights = 1 + len(state.extra_lights)

    if rank == 0:
        peak, peak_src = measured_peaks()
        st_mean = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
        # dominant kernel: k_march (fused with the merge), one launch per materialised level -> average launch.
        # Algorithmic bytes (8 B written per texel + 8 B read per texel of the level above) are counted
        # for the texels that were actually marched, not for the nominal cascade.
        march_ms = st_mean["march"]
        n_march = max(1, sum(1 for m in marched if m > 0))
        avg_march_launch_ms = march_ms / n_march
        march_bytes_per_launch = 8 * (sum(marched) + sum(marched[1:])) / n_march
        traced = int(sum(marched))
        frame_bytes = 8 * (sum(marched) + sum(marched[1:]) + int(levels[0].texel_count)) + W * H * 16
        gather_bytes = 8 * int(levels[0].texel_count) + W * H * 16
        roof = {"bound": "hbm", "kernel": "k_march<fused> (per-level ray march + merge; latency/issue-bound, BVH in L2)",
                "achieved": march_bytes_per_launch / (avg_march_launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "peak_source": peak_src, "traffic": None, "algorithmic_bytes_per_launch": march_bytes_per_launch,
                "avg_launch_ms": avg_march_launch_ms}
        import glob
        tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
        tpath = tpaths[-1] if tpaths else ""
        if tpath:      # dram bytes per launch from the latest committed ncu --set full capture of this workload
            with open(tpath) as fh:
                tj = json.load(fh)
            if wl in tj:
                roof["traffic"] = tj[wl]["k_march_dram_bytes_per_launch"]
                roof["traffic_source"] = tj["source"].replace("<workload>", wl)
        roof["frac"] = roof["achieved"] / peak
        gather = {"kernel": "k_gather_pipe", "bound": "hbm", "achieved": gather_bytes / (st_mean["gather"] * 1e-3) / 1e9, "peak": peak,
                  "unit": "GB/s", "bytes": gather_bytes}
        gather["frac"] = gather["achieved"] / peak
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "ms_per_step_stats": {"median": float(np.median(ms_steps)), "p10": float(np.percentile(ms_steps, 10)),
                                  "p90": float(np.percentile(ms_steps, 90)), "note": "rank 0, per-step CUDA-event times along the orbit"},
            "value_marched": world * traced / (ms_per_step * 1e-3) / 1e9,
            "all_rays": {"value": world * rays / (nocull_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": nocull_ms,
                         "note": "direction culling off (rc_set_tuning cull 0): every texel of every level marched; irradiance bit-identical"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic orbit camera over the reference's bundled scene (scenes/%s.zip)" % name,
            "config": {"workload": f"{name} {W}x{H}, {n_lights} light(s), full cascade stack", "levels": len(levels),
                       "probe_spacing0": int(levels[0].spacing), "dir_res0": int(levels[0].dir_res),
                       "rays_per_frame": rays, "rays_marched_per_frame": traced, "marched_fraction": traced / rays,
                       "culling": "texels whose weight on the way to the irradiance is provably zero are not marched (irradiance bit-identical, "
                                  "tests/test_gpu_parity.py::test_direction_culling_*); `value` counts the nominal cascade, value_marched the marched texels",
                       "triangles": int(info.num_triangles),
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "parallelism": "1 GPU" if world == 1 else f"multi-view batch, one orbit view per rank x{world}",
                       "cpu_binding": binding,
                       "merge": "separate kernels" if args.separate_merge else "fused into march",
                       "gather": "k_gather_pipe: software-pipelined column of 32x8 tiles per block (cp.async prefetch)",
                       "submission": "one CUDA graph per frame (stream capture + cudaGraphExecUpdate)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 80 + 16 * n_lights, "d2h_bytes_per_step": host_bytes,
                    "readback": "rc_read_target_async: double-buffered; frame i is submitted before the host waits for the read-back of "
                                "frame i-2, so copies overlap the following frames; every frame is received in pinned host memory "
                                "inside the timed region",
                    "blocking_value": world * rays / (e2e_sync_ms / args.steps * 1e-3) / 1e9,
                    "blocking_ms_per_step": e2e_sync_ms / args.steps},
            "gpu_launches": launches_per_frame * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "roofline_gather": gather,
            "frame_hbm": {"algorithmic_bytes": frame_bytes, "achieved_gbs": frame_bytes / (ms_per_step * 1e-3) / 1e9,
                          "frac": frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak, "note": "whole frame vs HBM peak; march is not HBM-bound"},
            "stage_ms": st_mean, "level_ms": lv_mean,
            "level_ms_note": "separate pass with per-level events (PDL overlap between level kernels off)",
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_oracle_run(wl, 3, 1, args.cpu_sample_div)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------- tiled (strong scaling) mode
def run_tiled(args):
    """One frame cut into screen-space tiles, one per rank, halo recomputed locally, finished tiles exchanged
    with a single NCCL all-gather.  Strong scaling: the frame is fixed as N grows."""
    import torch
    import torch.distributed as dist
    import radiancecascade_b200 as rc
    from radiancecascade_b200 import distributed as rd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = args.workload or "living_room_4k"
    name, W, H, lk = WORKLOADS[wl]
    grid = None
    if args.grid:
        nx, ny = (int(x) for x in args.grid.lower().split("x"))
        assert nx * ny == world, "--grid must multiply to the world size"
        grid = (nx, ny)
    state = rc.AppState()
    tr = rd.TiledRenderer(rank, world, local, (W, H), state, rc.scenes.scene_path(name), grid=grid)
    r = tr.renderer
    info = r.scene_info()
    full_levels = None
    rays_local = rays_per_frame(r.levels())
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peer = args.exchange == "peer"
    if peer:
        tr.attach_peers()

    def step(i):
        uc, pts = frame_inputs(rc, info, W, H, i, lk)      # every rank renders the SAME view
        state.uniform_camera = uc
        state.light_position, state.extra_lights = pts[0], pts[1:]
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            tr.render(state, stream.cuda_stream)
            e1.record(stream)
            if peer:      # tiles were stored into every rank's frame by the gather kernel itself: only wait for the flags
                full = tr.gather_peer(stream.cuda_stream)
            else:
                full = rd.all_gather_tiles(rd.irradiance_tensor(r), tr.tiles, W, H)
            e2.record(stream)
        return e0, e1, e2, full

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize(); dist.barrier()
    evs = [step(i)[:3] for i in range(args.steps)]
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([sum(a.elapsed_time(c) for a, _, c in evs), sum(a.elapsed_time(b) for a, b, _ in evs), float(rays_local)],
                     device="cuda", dtype=torch.float64)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        # rays of the undivided frame (what a single GPU would march)
        full_rays = sum((-(-W // (4 << i))) * (-(-H // (4 << i))) * (4 << i) ** 2 for i in range(6))
        ms = float(tmax[0]) / args.steps
        print(json.dumps({
            "metric": METRIC, "mode": "tiled", "value": full_rays / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "render_ms_per_step": float(tmax[1]) / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic orbit camera",
            "config": {"workload": f"{name} {W}x{H} tiled over {world} GPU(s)", "tiles": tr.tiles,
                       "halo": "recomputed locally",
                       "collective": "none: k_gather stores each tile into all ranks' frames over NVLink peer memory" if peer
                       else "NCCL all_gather_into_tensor of RGBA16F tiles",
                       "redundant_rays": float(tsum[2]) / full_rays - 1.0, "l2": "flushed between timed steps"}}))
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--separate-merge", action="store_true")
    ap.add_argument("--mode", default="batch", choices=["batch", "tiled"], help="N>1: independent views per rank (default) or one tiled frame")
    ap.add_argument("--grid", default=None, help="tiled mode: NXxNY tile grid (default: horizontal strips)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="tiled mode: peer-memory stores fused into the gather kernel (default) or an NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-div", type=int, default=1,
                    help="CPU arm renders at 1/div of the resolution per axis (1 = the full workload: a 1080p frame is ~0.3 s on 16 cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "tiled":
        return run_tiled(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
