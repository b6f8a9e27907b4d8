#!/usr/bin/env python
"""bench.py — headline benchmark of the radiance-cascade GI hot path on B200.

A "step" is one GI frame (G-buffer, probe placement, per-level ray march fused with the cascade merge, irradiance
gather) of a bundled scene along the deterministic orbit camera of SURVEY.md §8d.

  python bench.py --gpus N --steps K --warmup W            # product (CUDA, C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the same spec, all host cores

Workload (every N): BASELINE.json's north-star configuration — living_room, 3840x2160 (configs[3]).
  N = 1   one context renders the whole frame.
  N > 1   ONE frame cut into horizontal strips, one per rank (strong scaling): upper-cascade halo recomputed locally,
          strip cuts re-balanced from measured frame times, finished strips exchanged by the gather kernel itself
          (stores into every rank's frame over NVLink peer memory).  `--mode batch` runs the multi-view batch instead
          (one orbit view per rank, no data-path exchange, weak scaling); a short batch run on sonic 7680x4320
          (configs[4]) is attached to the tiled line as `extra.batch_sonic_8k`.
Metric: G ray-samples/s, one ray sample = one (probe, direction, level) interval query ACTUALLY marched for the frame
  a single GPU would render (direction culling skips texels whose weight on the way to the irradiance is provably zero;
  redundant halo rays of the tiled mode are not counted).  `ms_per_step` is the anchor; the nominal-cascade rate is the
  secondary key `value_nominal_cascade`.
The reference has no GI path and cannot be built here: the CPU arm is the repo's C oracle
(`kind: "port"`), which marches every texel of the cascade; it never loads librc_b200.so.
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
DEFAULT_WORKLOAD = "living_room_4k"
WORKLOADS = {  # BASELINE.json configs
    "teapot_1080p": ("teapot", 1920, 1080, "bench"),
    "test_room_1080p": ("test_room", 1920, 1080, "room"),
    "living_room_4k": ("living_room", 3840, 2160, "bench"),
    "sonic_8k": ("sonic", 7680, 4320, "bench"),
    "cube_512": ("cube", 512, 512, "bench"),
}


def workload_string(wl):
    name, W, H, lk = WORKLOADS[wl]
    return f"{name} {W}x{H}, {'4 room lights' if lk == 'room' else '1 light'}, full cascade stack (P0 4, D0 4, 6 levels)"


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


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis and all(x.strip().isdigit() for x in vis.split(",")) and local < len(vis.split(",")):
        return int(vis.split(",")[local])
    return local


def frame_inputs(rc, info, W, H, frame, lights_kind):
    lo, hi = list(info.bbox_min), list(info.bbox_max)
    pos, tgt, zn, zf = rc.scenes.orbit_camera(lo, hi, frame)
    proj = rc.Projection.new(W, H, 45.0, zn, zf)
    uc = rc.UniformCamera.look_at(pos, tgt, proj)
    pts = rc.scenes.room_lights(lo, hi) if lights_kind == "room" else [rc.scenes.bench_light(lo, hi)]
    return uc, pts


def nominal_rays(levels):
    return int(sum(l.texel_count for l in levels))


def psnr(a, b, peak):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else float(10.0 * np.log10(peak * peak / mse))


# ------------------------------------------------------------------------- CPU arm (never touches librc_b200.so)
def cpu_oracle_frames(workload, steps, warmup, sample_div, keep_last=False, cam_fn=None):
    """Times the C oracle (all host threads) on `steps` frames of the workload's orbit at 1/sample_div of the resolution
    per axis (1 = the workload itself).  Scene ingest, camera and lights come from oracle/ref_ingest.py and the pure-Python
    scene helpers: the product library is not loaded."""
    import math
    from radiancecascade_b200 import scenes           # pure Python (paths, orbit); importing the package loads no .so
    from oracle import gi_oracle as go
    from oracle import ref_ingest as ri
    name, W, H, lk = WORKLOADS[workload]
    w, h = W // sample_div, H // sample_div
    osc = go.OracleScene(scenes.scene_path(name))
    go.set_num_threads(go.host_cores())     # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    p = osc.params(w, h, store_half=True)
    lv = osc.levels(p)
    rays = sum(l.gw * l.gh * l.D * l.D for l in lv)
    times, last = [], None
    for i in range(warmup + steps):
        pos, tgt, zn, zf = scenes.orbit_camera(osc.bbox_min, osc.bbox_max, i)
        if cam_fn is not None:     # the product arm's parity check hands the oracle the very camera bytes the GPU frame used
            cam = cam_fn(i, w, h)
        else:
            cam = ri.uniform_camera_look_at(pos, tgt, np.float32(math.radians(45.0)), np.float32(w) / np.float32(h), zn, zf)
        pts = scenes.room_lights(osc.bbox_min, osc.bbox_max) if lk == "room" else [scenes.bench_light(osc.bbox_min, osc.bbox_max)]
        larr = np.array([[q[0], q[1], q[2], 1.0] for q in pts], dtype=np.float32)
        t = time.perf_counter()
        out = osc.render(p, cam, larr)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
        if keep_last:
            last = out
    ms = 1e3 * float(np.mean(times))
    res = {"value": rays / (ms * 1e-3) / 1e9, "unit": UNIT, "cores": go.num_threads(), "kind": "port",
           "sample": (f"{name} {w}x{h}" + (f" (1/{sample_div} of {W}x{H} per axis)" if sample_div > 1 else " (the full workload)") +
                      f", full cascade stack, every texel marched, {steps} frame(s) of the orbit, oracle/rc_oracle.c with OpenMP"),
           "ms_per_step": ms, "rays_per_step": rays}
    return res, last


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload or DEFAULT_WORKLOAD
    cb, _ = cpu_oracle_frames(wl, args.steps, args.warmup, args.cpu_sample_div)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if (args.gpus > 1 and args.mode == "tiled") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic orbit camera over the reference's bundled scene",
            "config": {"workload": workload_string(wl), "levels": 6, "probe_spacing0": 4, "dir_res0": 4,
                       "sample": cb["sample"], "rays_per_frame": cb["rays_per_step"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference repository has no GI path and cannot be built here (no Rust / Vulkan); this arm is the repo's CPU "
                    "oracle of the same specification (kind: port).  It marches every texel of the cascade; the product marches "
                    "only the texels with a non-zero weight and reports those as its ray samples"}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa(local):
    """Multi-rank runs: keep this process (and the pinned buffers it allocates: first touch) on the CPUs NVML names as
    closest to its GPU.  Returns a description that includes what NVML reported, so the line shows whether the box has
    more than one NUMA node at all."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        try:
            node = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            node = None
        allowed = set(os.sched_getaffinity(0))
        cpus &= allowed
        if len(cpus) >= 2 and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} of {len(allowed)} CPUs (NVML affinity, NUMA node {node})"
        return f"not narrowed: NVML affinity covers all {len(allowed)} allowed CPUs (NUMA node {node})"
    except Exception as e:      # noqa: BLE001
        return f"none ({type(e).__name__})"


def roofline_block(marched_per_level, stage_ms, levels, W, H, wl):
    """`roofline` (dominant kernel, contract) + the gather's own and the whole frame's HBM figures."""
    peak, peak_src = measured_peaks()
    n_march = max(1, sum(1 for m in marched_per_level if m > 0))
    avg_launch_ms = stage_ms["march"] / n_march
    # SURVEY §8d per texel: 8 B written + 8 B of the level above read; counted for the texels actually marched
    bytes_per_launch = 8 * (sum(marched_per_level) + sum(marched_per_level[1:])) / n_march
    roof = {"bound": "hbm", "kernel": "k_march<fused> (per-level ray march + merge; issue / divergence-bound BVH traversal, scene in L2)",
            "achieved": bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "peak_source": peak_src,
            "traffic": None, "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_launch_ms}
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if tpaths:      # dram bytes per launch from the latest committed ncu --set full capture of this workload
        with open(tpaths[-1]) as fh:
            tj = json.load(fh)
        if wl in tj:
            roof["traffic"] = tj[wl]["k_march_dram_bytes_per_launch"]
            roof["traffic_source"] = tj["source"].replace("<workload>", wl) + " (committed capture, not measured in this run)"
    roof["frac"] = roof["achieved"] / peak
    gather_bytes = 8 * int(levels[0].texel_count) + W * H * 16
    gather = {"kernel": "k_gather_mma (tensor-core gather, TMA-staged probes)", "bound": "hbm",
              "achieved": gather_bytes / (stage_ms["gather"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "bytes": gather_bytes}
    gather["frac"] = gather["achieved"] / peak
    frame_bytes = 8 * (sum(marched_per_level) + sum(marched_per_level[1:]) + int(levels[0].texel_count)) + W * H * 16
    return roof, gather, frame_bytes, peak


# ------------------------------------------------------------------------- product arm, one frame per rank
def run_product(args):
    """N = 1 (the headline single-GPU line) and `--mode batch` for N > 1 (every rank renders its own views)."""
    import ctypes as C
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

    wl = args.workload or DEFAULT_WORKLOAD
    name, W, H, lk = WORKLOADS[wl]
    state = rc.AppState()
    flags = _ffi.RC_CFG_SEPARATE_MERGE if args.separate_merge else 0
    r = rc.DefaultRenderer.new(local, (W, H), state, rc.scenes.scene_path(name), rc.CascadeConfig(flags=flags))
    info = r.scene_info()
    levels = r.levels()
    nominal = nominal_rays(levels)

    stream = torch.cuda.Stream()
    sh = stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    host_out = torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True)
    host_ptr, host_bytes = host_out.data_ptr(), host_out.numel() * 2

    def set_frame(i):
        uc, pts = frame_inputs(rc, info, W, H, i * world + rank, lk)   # batch mode: rank r renders its own views
        state.uniform_camera = uc
        state.light_position, state.extra_lights = pts[0], pts[1:]
        r.update(state)

    # ---- device-timed loop: per-step events, L2 flushed between steps --------------------
    def device_loop(n, record, marched):
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
                marched.append([int(l.texel_count) if m is None else m for l, m in zip(levels, r.rays_marched())])
        stream.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    warm = max(args.warmup, 3)
    device_loop(warm, None, None)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(local))
    if rank == 0:
        sampler.start()
    stages, marched_steps = [], []
    ms_steps = device_loop(args.steps, stages, marched_steps)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    total_ms = float(sum(ms_steps))
    traced_total = float(sum(sum(m) for m in marched_steps))
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        t = torch.tensor([traced_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        traced_total = float(t.item())
    else:
        pass
    ms_per_step = total_ms / args.steps
    value = traced_total / (total_ms * 1e-3) / 1e9
    value_nominal = world * nominal / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public API: host camera in, host irradiance out ----------
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

    host_out2 = torch.empty((H, W, 4), dtype=torch.float16, pin_memory=True)
    host_ptrs = (host_ptr, host_out2.data_ptr())
    rgb48_bytes = W * H * 6
    e2e_fmt = {"rgb48": False}

    def e2e_pipelined_loop(n):
        # frame i is enqueued BEFORE the host waits for frame i-2 (the previous user of host buffer i & 1): the GPU always
        # has the next frame queued; the library orders the device side (frame i's gather waits for the read-back of i-2)
        t0 = time.perf_counter()
        tickets = [None, None]
        for i in range(n):
            r.update_packed(e2e_inputs[i])
            r.render(sh)
            if tickets[i & 1] is not None:
                r.read_wait(tickets[i & 1])
            if e2e_fmt["rgb48"]:
                tickets[i & 1] = r.read_irradiance_async(host_ptrs[i & 1], rgb48_bytes, rgb48=True)
            else:
                tickets[i & 1] = r.read_irradiance_async(host_ptrs[i & 1], host_bytes)
        for k in (n, n + 1):
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
    e2e_rgba_ms = timed(e2e_pipelined_loop)
    e2e_fmt["rgb48"] = True          # the 6-byte target: r, g, b bit-exact, coverage in the sign bit of r
    e2e_ms = timed(e2e_pipelined_loop)
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
    # the same frames with direction culling off: every texel of every level is marched (what the CPU arm does)
    r.set_tuning("cull", 0)
    device_loop(3, None, None)
    torch.cuda.synchronize()
    nocull_ms = float(sum(device_loop(args.steps, None, None))) / args.steps
    if dist:
        t = torch.tensor([nocull_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nocull_ms = float(t.item())
    r.set_tuning("cull", 1)
    traced_per_step = traced_total / args.steps
    e2e_value = traced_per_step / (e2e_ms / args.steps * 1e-3) / 1e9
    n_lights = 1 + len(state.extra_lights)

    if rank == 0:
        st_mean = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
        marched = [float(np.mean([m[i] for m in marched_steps])) for i in range(len(levels))]
        roof, gather, frame_bytes, peak = roofline_block(marched, st_mean, levels, W, H, wl)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step,
            "ms_per_step_stats": {"median": float(np.median(ms_steps)), "p10": float(np.percentile(ms_steps, 10)),
                                  "p90": float(np.percentile(ms_steps, 90)), "note": "rank 0, per-step CUDA-event times along the orbit"},
            "value_nominal_cascade": value_nominal,
            "all_rays": {"value": world * nominal / (nocull_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": nocull_ms,
                         "note": "direction culling off (rc_set_tuning cull 0): every texel of every level marched; irradiance bit-identical"},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic orbit camera over the reference's bundled scene (scenes/%s.zip)" % name,
            "config": {"workload": workload_string(wl), "levels": len(levels),
                       "probe_spacing0": int(levels[0].spacing), "dir_res0": int(levels[0].dir_res),
                       "rays_per_frame_nominal": nominal, "rays_marched_per_frame": traced_per_step / world,
                       "marched_fraction": traced_per_step / world / nominal,
                       "value_definition": "ray samples actually marched per second (texels whose weight on the way to the irradiance is "
                                           "provably zero are culled; irradiance bit-identical, tests/test_gpu_parity.py::test_direction_culling_*)",
                       "triangles": int(info.num_triangles),
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "parallelism": "1 GPU" if world == 1 else f"multi-view batch, one orbit view per rank x{world}",
                       "collective": "none (independent views)" if world > 1 else "none",
                       "cpu_binding": binding,
                       "merge": "separate kernels" if args.separate_merge else "fused into march",
                       "gather": "k_gather_mma: per-cell contraction on mma.sync, probes staged by TMA bulk copies",
                       "submission": "one CUDA graph per frame (stream capture + cudaGraphExecUpdate)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 80 + 16 * n_lights, "d2h_bytes_per_step": rgb48_bytes,
                    "readback": "rc_read_target_async(RC_TARGET_IRRADIANCE_RGB48): the irradiance as 3 x float16 per pixel (r, g, b bit-exact, "
                                "coverage in the sign bit of r) — alpha carries no other information; double-buffered, frame i is submitted before "
                                "the host waits for the read-back of frame i-2; every frame is received in pinned host memory inside the timed region",
                    "rgba16f_ms_per_step": e2e_rgba_ms / args.steps, "rgba16f_d2h_bytes_per_step": host_bytes,
                    "blocking_value": traced_per_step / (e2e_sync_ms / args.steps * 1e-3) / 1e9,
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
            # the CPU arm renders the same frames; its last frame is diffed against the GPU's frame of the same camera
            n_cpu = 2 if W * H > 4_000_000 else 3
            div = args.cpu_sample_div
            cb, last = cpu_oracle_frames(wl, n_cpu, 1, div, keep_last=div == 1,
                                         cam_fn=lambda i, w, h: frame_inputs(rc, info, w, h, i, lk)[0].as_array())
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["ms_per_step"] = cb["ms_per_step"]
            if last is not None:
                set_frame(n_cpu)                      # frames 0 .. n_cpu: warm-up + n_cpu timed -> the last one is index n_cpu
                r.render(sh)
                E = r.read_target(_ffi.RC_TARGET_IRRADIANCE).astype(np.float32)
                Eo = last["irradiance"]
                pk = float(Eo[..., :3].max())
                line["parity"] = {"frame": n_cpu, "max_abs": float(np.abs(E[..., :3] - Eo[..., :3]).max()), "peak": pk,
                                  "psnr_db": psnr(E[..., :3], Eo[..., :3], pk),
                                  "prim_equal": bool(np.array_equal(r.read_target(_ffi.RC_TARGET_PRIM), last["prim"])),
                                  "tolerance": "rc_spec.h S10: max_abs <= 1e-2 * peak, PSNR >= 50 dB, primary visibility bit-exact"}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------- tiled (strong scaling) mode
def run_tiled(args):
    """One frame cut into horizontal strips, one per rank: halo recomputed locally, cuts re-balanced from measured frame
    times (radiancecascade_b200.distributed.StripBalancer), finished strips exchanged inside the gather kernel over NVLink
    peer memory (or one NCCL all-gather with --exchange nccl).  Strong scaling: the frame is fixed as N grows."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import radiancecascade_b200 as rc
    from radiancecascade_b200 import _ffi, distributed as rd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    binding = bind_to_gpu_numa(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctl = dist.new_group(backend="gloo")          # host-side control plane (frame times for the balancer)
    wl = args.workload or DEFAULT_WORKLOAD
    name, W, H, lk = WORKLOADS[wl]
    grid = None
    if args.grid:
        nx, ny = (int(x) for x in args.grid.lower().split("x"))
        assert nx * ny == world, "--grid must multiply to the world size"
        grid = (nx, ny)
    state = rc.AppState()
    tr = rd.TiledRenderer(rank, world, local, (W, H), state, rc.scenes.scene_path(name), grid=grid, balance=not args.no_balance,
                          halo_exchange=args.halo == "exchange")
    r = tr.renderer
    info = r.scene_info()
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peer = args.exchange == "peer"
    if peer:
        tr.attach_peers()

    def set_frame(i):
        uc, pts = frame_inputs(rc, info, W, H, i, lk)      # every rank renders the SAME view
        state.uniform_camera = uc
        state.light_position, state.extra_lights = pts[0], pts[1:]

    def step(i):
        set_frame(i)
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            tr.render(state, sh)
            e1.record(stream)
            if peer:      # tiles were stored into every rank's frame by the gather kernel itself: only wait for the flags
                full = tr.gather_peer(sh)
            else:
                full = rd.all_gather_tiles(rd.irradiance_tensor(r), tr.tiles, W, H)
            e2.record(stream)
        return e0, e1, e2, full

    def rebalance_from(e0, e1):
        """times of the frame just rendered -> new cuts (every rank computes the same ones)"""
        stream.synchronize()
        moved = tr.rebalance(tr.all_gather_times(e0.elapsed_time(e1), group=ctl))
        dist.barrier(group=ctl)          # nobody starts the next frame while a rank is still re-tiling
        return moved

    # the frame a single GPU would render: its marched-ray count per step is the work unit of `value`
    useful = []
    if rank == 0:
        ref = rc.DefaultRenderer.new(local, (W, H), rc.AppState(), rc.scenes.scene_path(name))
        st0 = rc.AppState()
        for i in range(args.steps):
            uc, pts = frame_inputs(rc, info, W, H, i, lk)
            st0.uniform_camera = uc
            st0.light_position, st0.extra_lights = pts[0], pts[1:]
            ref.update(st0)
            ref.render()
            useful.append(sum(int(l.texel_count) if m is None else m for l, m in zip(ref.levels(), ref.rays_marched())))
        full_levels = ref.levels()
        nominal = nominal_rays(full_levels)
        del ref
        torch.cuda.empty_cache()

    warm = max(args.warmup, 3)
    rebalances = 0
    for it in range(args.calibrate):              # calibration on frame 0's camera: move the cuts until the ranks' times agree
        e0, e1, _, _ = step(0)
        if tr.balancer is None or not rebalance_from(e0, e1):
            if it >= 2:
                break
        else:
            rebalances += 1
    for i in range(warm):
        step(i)
    torch.cuda.synchronize(); dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local))
    if rank == 0:
        sampler.start()
    evs, stage_acc, rays_local, tiles_hist = [], [], [], []
    for i in range(args.steps):
        e0, e1, e2, _ = step(i)
        evs.append((e0, e1, e2))
        stream.synchronize()
        stage_acc.append(r.stage_times())
        rays_local.append(sum(m for m in r.rays_marched() if m))
        tiles_hist.append([t[3] for t in tr.tiles])
        # keep following the orbit: re-balance (outside the step's event bracket) every `rebalance_every` steps
        if tr.balancer is not None and args.rebalance_every > 0 and (i + 1) % args.rebalance_every == 0 and i + 1 < args.steps:
            if tr.rebalance(tr.all_gather_times(e0.elapsed_time(e1), group=ctl)):
                rebalances += 1
            dist.barrier(group=ctl)      # nobody starts the next frame while a rank is still re-tiling
    torch.cuda.synchronize(); dist.barrier()
    per_step = torch.tensor([[a.elapsed_time(c), a.elapsed_time(b)] for a, b, c in evs], device="cuda", dtype=torch.float64)
    mx = per_step.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)        # per step: the slowest rank
    mean = per_step.clone(); dist.all_reduce(mean, op=dist.ReduceOp.SUM); mean /= world
    rl = torch.tensor([float(sum(rays_local))], device="cuda", dtype=torch.float64)
    dist.all_reduce(rl, op=dist.ReduceOp.SUM)
    st_mean = {k: float(np.mean([s[k] for s in stage_acc])) for k in stage_acc[0]}
    st_all = [None] * world
    dist.all_gather_object(st_all, st_mean, group=ctl)

    tr_halo = tr.halo_exchange
    launches = torch.tensor([float(r.launch_count())], device="cuda", dtype=torch.float64)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)            # kernels per frame, all ranks

    # ---- end to end: host camera in, the whole frame out in (shared, pinned) host memory, every rank reading back its strip
    from multiprocessing import shared_memory
    shm_name = f"rcb200_frame_{os.environ.get('MASTER_PORT', '0')}"
    nbytes = W * H * 6                               # RC_TARGET_IRRADIANCE_RGB48: 3 x float16 per pixel
    if rank == 0:
        try:      # a crashed earlier run on the same port may have left the segment behind
            stale = shared_memory.SharedMemory(name=shm_name)
            stale.close(); stale.unlink()
        except FileNotFoundError:
            pass
    shm = shared_memory.SharedMemory(name=shm_name, create=True, size=2 * nbytes) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        shm = shared_memory.SharedMemory(name=shm_name)
        try:      # only the creator may unlink: keep this process's resource tracker from trying (and warning) at exit
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:       # noqa: BLE001
            pass
    host = np.ndarray((2, H, W, 3), dtype=np.uint16, buffer=shm.buf)
    base_ptr = host.ctypes.data
    cudart = torch.cuda.cudart()
    reg = cudart.cudaHostRegister(base_ptr, 2 * nbytes, 0)
    registered = int(reg) == 0 if not isinstance(reg, tuple) else int(reg[0]) == 0
    inputs = []
    for i in range(max(args.steps, 2)):
        set_frame(i)
        inputs.append(r.pack_update(state))

    def e2e_loop(n):
        t0 = time.perf_counter()
        tickets = [None, None]
        for i in range(n):
            r.update_packed(inputs[i])
            if tr.halo_exchange:
                with torch.cuda.stream(stream):
                    tr._render_exchange(sh)
            else:
                r.render(sh)
            if tickets[i & 1] is not None:
                r.read_wait(tickets[i & 1])
            x0, y0, w, h = tr.tiles[rank]
            tickets[i & 1] = r.read_irradiance_async(base_ptr + (i & 1) * nbytes + y0 * W * 6, w * h * 6, rgb48=True)   # my rows of the shared frame
        for k in (n, n + 1):
            if tickets[k & 1] is not None:
                r.read_wait(tickets[k & 1])
        return (time.perf_counter() - t0) * 1e3

    e2e_ms = None
    if not grid:                                   # strips are contiguous row blocks of the frame
        if peer:
            r.set_tuning("peer_stores", 0)         # the host frame is assembled from the ranks' own read-backs: no device exchange
        e2e_loop(2)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e2e_loop(args.steps)], device="cuda", dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        if peer:
            r.set_tuning("peer_stores", 1)
    if registered:
        cudart.cudaHostUnregister(base_ptr)
    del host
    shm.close()
    dist.barrier()
    if rank == 0:
        shm.unlink()

    # ---- secondary record: multi-view batch on sonic 7680x4320 (BASELINE.json configs[4]), one orbit view per rank
    extra = {}
    if not args.no_batch_extra:
        del tr, r
        torch.cuda.empty_cache()
        bname, BW, BH, blk = WORKLOADS["sonic_8k"]
        bst = rc.AppState()
        br = rc.DefaultRenderer.new(local, (BW, BH), bst, rc.scenes.scene_path(bname))
        binfo = br.scene_info()
        bev, bm = [], 0
        nb = 6
        for i in range(3 + nb):
            uc, pts = frame_inputs(rc, binfo, BW, BH, i * world + rank, blk)
            bst.uniform_camera = uc
            bst.light_position, bst.extra_lights = pts[0], pts[1:]
            br.update(bst)
            with torch.cuda.stream(stream):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); br.render(sh); b.record(stream)
            if i >= 3:
                bev.append((a, b))
                stream.synchronize()
                bm += sum(int(l.texel_count) if m is None else m for l, m in zip(br.levels(), br.rays_marched()))
        torch.cuda.synchronize()
        tb = torch.tensor([sum(a.elapsed_time(b) for a, b in bev), float(bm)], device="cuda", dtype=torch.float64)
        tmx = tb.clone(); dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        tsm = tb.clone(); dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
        extra["batch_sonic_8k"] = {"workload": workload_string("sonic_8k"), "scaling": "weak", "collective": "none (independent views)",
                                   "ms_per_step_per_rank": float(tmx[0]) / nb, "value": float(tsm[1]) / (float(tmx[0]) * 1e-3) / 1e9,
                                   "unit": UNIT, "steps": nb, "n_gpus": world}

    if rank == 0:
        clocks = sampler.stop()
        total_ms = float(mx[:, 0].sum())
        ms = total_ms / args.steps
        render_ms = float(mx[:, 1].sum()) / args.steps
        useful_total = float(sum(useful))
        roof = None                   # per-rank stage means (rank-ordered, `stage_ms_per_rank`) = the timeline of the strong-scaling limit
        try:
            peak, peak_src = measured_peaks()
            n_l = len(full_levels) - 1
            avg_launch_ms = st_all[0]["march"] / max(1, n_l)
            bpl = 16.0 * (float(rl.item()) / args.steps / world) / max(1, n_l)
            roof = {"bound": "hbm", "kernel": "k_march<fused> (per rank; issue / divergence-bound BVH traversal)", "achieved": bpl / (avg_launch_ms * 1e-3) / 1e9,
                    "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None, "algorithmic_bytes_per_launch": bpl,
                    "avg_launch_ms": avg_launch_ms}
            roof["frac"] = roof["achieved"] / peak
        except Exception:       # noqa: BLE001
            pass
        line = {
            "metric": METRIC, "mode": "tiled", "value": useful_total / (total_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "render_ms_per_step": render_ms,
            "mean_rank_ms_per_step": float(mean[:, 0].sum()) / args.steps,
            "value_nominal_cascade": nominal / (ms * 1e-3) / 1e9,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic orbit camera over the reference's bundled scene (scenes/%s.zip)" % name,
            "config": {"workload": workload_string(wl), "levels": 6, "probe_spacing0": 4, "dir_res0": 4,
                       "parallelism": f"one frame, {world} " + (f"tiles {args.grid}" if grid else "horizontal strips") + ", one per rank",
                       "tiles_rows_first_step": tiles_hist[0], "tiles_rows_last_step": tiles_hist[-1],
                       "balancing": ("off" if (args.no_balance or grid) else
                                     f"strip cuts from measured frame times: {rebalances} re-tilings (calibration + every {args.rebalance_every} steps, "
                                     "outside the timed brackets)"),
                       "halo": ("EXCHANGED (RC_CFG_HALO_EXCHANGE): levels >= 1 marched by the owning rank only; request masks once per frame and "
                                "child averages once per level moved between strip neighbours with NCCL send/recv" if tr_halo else
                                "upper-cascade halo recomputed locally (no data-path exchange between cascade levels)"),
                       "collective": (("NCCL send/recv (halo: request masks + child averages of border probe rows); final image: " if tr_halo else
                                       "none between cascade levels; final image: ") +
                                      ("k_gather_mma stores each finished strip into rank 0's frame over NVLink peer memory, flags instead of a collective"
                                       if peer else "NCCL all_gather_into_tensor of the RGBA16F strips")),
                       "rays_marched_single_gpu_per_frame": useful_total / args.steps,
                       "redundant_rays": float(rl.item()) / useful_total - 1.0, "l2": "flushed between timed steps (256 MiB memset)",
                       "value_definition": "ray samples a single GPU marches for the frame (redundant halo rays not counted) per second",
                       "cpu_binding": binding},
            "e2e": ({"value": useful_total / args.steps / (e2e_ms / args.steps * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                     "h2d_bytes_per_step": 96 * world, "d2h_bytes_per_step": nbytes,
                     "readback": "every rank reads its strip back (RC_TARGET_IRRADIANCE_RGB48: 3 x float16 per pixel, r g b bit-exact) into its rows of "
                                 "ONE process-shared pinned host frame (double-buffered, rc_read_target_async): the whole frame is in host memory "
                                 "when the step ends; N PCIe links in parallel",
                     "host_registered": registered} if e2e_ms else None),
            "gpu_launches": int(launches.item()) * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "stage_ms_per_rank": st_all,
            "extra": extra,
        }
        print(json.dumps(line))
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
    ap.add_argument("--mode", default="tiled", choices=["batch", "tiled"],
                    help="N>1: one tiled frame, strong scaling (default) or independent views per rank, weak scaling")
    ap.add_argument("--grid", default=None, help="tiled mode: NXxNY tile grid (default: balanced horizontal strips)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="tiled mode: peer-memory stores fused into the gather kernel (default) or an NCCL all-gather")
    ap.add_argument("--no-balance", action="store_true", help="tiled mode: equal-height strips")
    ap.add_argument("--halo", default="recompute", choices=["recompute", "exchange"],
                    help="tiled mode: upper-cascade halos recomputed by every rank (default) or exchanged between the ranks (NCCL send/recv)")
    ap.add_argument("--calibrate", type=int, default=10, help="tiled mode: at most this many balancing frames before the warm-up")
    ap.add_argument("--rebalance-every", type=int, default=4, help="tiled mode: re-balance every n timed steps (0: never)")
    ap.add_argument("--no-batch-extra", action="store_true", help="tiled mode: skip the sonic 8K batch record")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-div", type=int, default=1,
                    help="CPU arm renders at 1/div of the resolution per axis (1 = the full workload: a 4K frame is ~1.3 s on 16 cores)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args)
    if world > 1 and args.mode == "tiled":
        return run_tiled(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
