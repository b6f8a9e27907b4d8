#!/usr/bin/env python
"""Generate tests/golden/*: known-answer vectors produced by the CPU oracles (oracle/), run in
the build container.  The reference ships no golden vectors of its own, so these
pin the ORACLE (ingest restatement + GI spec), not the reference: "parity unpinned".

  python tools/make_golden.py

ingest_golden.json   per scene / model: counts, sha256 of the 17-float vertex stream and the index
                     buffer, UniformMaterial, enable_bit, texture sha256, bbox  (a1-a6)
camera_golden.json   UniformCamera bytes for the reference-default camera (with its quirks) and
                     orbit cameras  (a7)
fs_main_golden.npz   fs_main at 512 fixed surface points of the cube scene  (a10)
gi_cube64.npz        layout tables, directions, G-buffer prim/depth and irradiance of the cube at
                     64x64  (a12-a14; rc_spec.h)
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import radiancecascade_b200 as rc   # noqa: E402  (scene archives + camera helpers only)
from oracle import gi_oracle as go  # noqa: E402
from oracle import ref_ingest as ri  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ingest():
    out = {}
    for name in rc.scenes.SCENES:
        scenes, light = ri.ObjScene.load(rc.scenes.scene_path(name))
        models = []
        allpos = np.concatenate([s.vertices() for s in scenes])
        for s in scenes:
            mat = s.material()
            vs = s.vertex_stream()
            models.append({
                "name": s.name(), "vertices": int(len(vs)), "indices": int(len(s.indices())),
                "stream_sha256": sha(vs), "index_sha256": sha(s.indices().astype(np.uint32)),
                "nan_floats": int(np.isnan(vs).sum()),
                "uniform_material": [float(x) for x in ri.uniform_material(mat)],
                "enable_bit": ri.enable_bit(mat),
                "emission": [float(x) for x in (mat.emission if mat is not None else np.zeros(3))],
                "color_texture_sha256": sha(mat.color_texture) if (mat is not None and mat.color_texture is not None) else None,
                "normal_texture_sha256": sha(mat.normal_texture) if (mat is not None and mat.normal_texture is not None) else None,
            })
        out[name] = {"models": models, "light": None if light is None else [float(x) for x in light],
                     "bbox_min": [float(x) for x in allpos.min(0)], "bbox_max": [float(x) for x in allpos.max(0)],
                     "total_vertices": int(sum(m["vertices"] for m in models)),
                     "total_triangles": int(sum(m["indices"] for m in models) // 3)}
    with open(os.path.join(OUT, "ingest_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def camera():
    out = {}
    # reference default (src/app.rs:25-26): degrees used as radians, pitch clamped on the first frame
    pitch = float(min(max(np.float32(-20.0), -rc.renderer.SAFE_FRAC_PI_2), rc.renderer.SAFE_FRAC_PI_2))
    for label, (w, h) in {"default_1360x1360": (1360, 1360), "default_1360x768": (1360, 768)}.items():
        a = ri.uniform_camera((0.0, 5.0, 10.0), -90.0, pitch, np.float32(np.radians(np.float32(45.0))), np.float32(w) / np.float32(h), 0.1, 100.0)
        out[label] = {"hex": a.tobytes().hex(), "values": [float(x) for x in a]}
    for name in ("cube", "teapot"):
        sc, _ = ri.ObjScene.load(rc.scenes.scene_path(name))
        pos = np.concatenate([s.vertices() for s in sc])
        p, t, zn, zf = rc.scenes.orbit_camera(pos.min(0), pos.max(0), 5)
        a = ri.uniform_camera_look_at(p, t, np.float32(np.radians(np.float32(45.0))), np.float32(1920) / np.float32(1080), zn, zf)
        out[f"orbit5_{name}_1920x1080"] = {"hex": a.tobytes().hex(), "position": [float(x) for x in p], "target": [float(x) for x in t],
                                           "znear": zn, "zfar": zf}
    with open(os.path.join(OUT, "camera_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def fs_main():
    osc = go.OracleScene(rc.scenes.scene_path("cube"))
    rng = np.random.default_rng(7)
    n = 512
    prim = rng.integers(0, len(osc.tris), n).astype(np.uint32)
    u = (rng.random(n) * 0.9 + 0.05).astype(np.float32)
    v = (rng.random(n) * (1 - u) * 0.9).astype(np.float32)
    eye = (rng.normal(size=(n, 3)) * 4).astype(np.float32)
    pts = np.zeros((n, 8), np.float32)
    pts[:, 0] = prim.view(np.float32); pts[:, 1] = u; pts[:, 2] = v; pts[:, 4:7] = eye
    lights = np.array([[0.5, 2.5, 1.0, 1.0]], np.float32)
    out = osc.shade_points(pts, lights)
    np.savez_compressed(os.path.join(OUT, "fs_main_golden.npz"), points=pts, lights=lights, radiance=out)


def gi():
    name, W, H = "cube", 64, 64
    osc = go.OracleScene(rc.scenes.scene_path(name))
    pos, tgt, zn, zf = rc.scenes.orbit_camera(osc.bbox_min, osc.bbox_max, 5)
    cam = ri.uniform_camera_look_at(pos, tgt, np.float32(np.radians(np.float32(45.0))), np.float32(W) / np.float32(H), zn, zf)
    lights = np.array([[*rc.scenes.bench_light(osc.bbox_min, osc.bbox_max), 1.0]], np.float32)
    p = osc.params(W, H, store_half=True)
    out = osc.render(p, cam, lights)
    lv = out["levels"]
    np.savez_compressed(
        os.path.join(OUT, "gi_cube64.npz"), cam=cam, lights=lights,
        intervals=np.array([p.L0, p.t_far, p.offset], np.float32),
        levels=np.array([[l.P, l.D, l.gw, l.gh] for l in lv], np.int32),
        t_ranges=np.array([[l.t0, l.t1] for l in lv], np.float32),
        dirs0=out["dirs"][0], dirs1=out["dirs"][1],
        prim=out["prim"], depth=out["depth"], normal=out["normal"],
        irradiance=out["irradiance"].astype(np.float16),
        cascade0=out["cascades"][0].astype(np.float16))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ingest(); camera(); fs_main(); gi()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
