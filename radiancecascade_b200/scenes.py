"""Bundled scenes and synthetic camera paths.

The reference's assets live in scenes/*.zip (tools/pack_scenes.py).  `scene_path`
extracts an archive next to it (scenes/_extracted/, git-ignored) and returns the
.obj path to hand to rc_create.  librc_b200 decodes PNG and JPEG textures itself
(csrc/image.cpp, csrc/jpeg.cpp); `sidecars=True` additionally writes PIL-decoded
"<file>.rgba8" files, which the loader prefers when present (a way to inject
externally decoded pixels; not used by the tests or the benchmark).
"""
from __future__ import annotations

import math
import os
import struct
import zipfile
from typing import Dict, List, Tuple

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_DIR = os.path.join(_ROOT, "scenes")
EXTRACT_DIR = os.path.join(SCENE_DIR, "_extracted")

# scene name -> .obj path inside the archive (relative to the reference's resources/)
SCENES: Dict[str, str] = {
    "cube": "cube/cube.obj",
    "teapot": "teapot/teapot.obj",
    "test_room": "test_room/test_room.obj",
    "living_room": "living_room/living_room.obj",
    "sonic": "sonic.obj",
}


def resource_root() -> str:
    """≙ RESOURCE_PATH (src/primitives.rs:12) for the extracted archives."""
    return EXTRACT_DIR


def _write_sidecar(img_path: str) -> None:
    side = img_path + ".rgba8"
    if os.path.exists(side) and os.path.getmtime(side) >= os.path.getmtime(img_path):
        return
    from PIL import Image
    with Image.open(img_path) as im:
        rgba = np.asarray(im.convert("RGBA"), dtype=np.uint8)
    tmp = side + f".tmp{os.getpid()}"
    with open(tmp, "wb") as fh:
        fh.write(struct.pack("<II", rgba.shape[1], rgba.shape[0]))
        fh.write(rgba.tobytes())
    os.replace(tmp, side)


def scene_path(name: str, sidecars: bool = False) -> str:
    """Extract scenes/<name>.zip if needed; return the absolute .obj path."""
    if name not in SCENES:
        raise KeyError(f"unknown scene {name!r}; have {sorted(SCENES)}")
    obj = os.path.join(EXTRACT_DIR, SCENES[name])
    marker = os.path.join(EXTRACT_DIR, f".{name}.done")
    if not os.path.exists(marker):
        os.makedirs(EXTRACT_DIR, exist_ok=True)
        with zipfile.ZipFile(os.path.join(SCENE_DIR, name + ".zip")) as z:
            z.extractall(EXTRACT_DIR)
            members = z.namelist()
        for m in members:
            if m.lower().endswith((".jpg", ".jpeg")):
                if sidecars:
                    _write_sidecar(os.path.join(EXTRACT_DIR, m))
                elif os.path.exists(os.path.join(EXTRACT_DIR, m) + ".rgba8"):
                    os.remove(os.path.join(EXTRACT_DIR, m) + ".rgba8")
        with open(marker, "w") as fh:
            fh.write("ok\n")
    return obj


def orbit_camera(bbox_min, bbox_max, frame: int, n_frames: int = 64) -> Tuple[np.ndarray, np.ndarray, float, float]:
    """Deterministic orbit of SURVEY §8d: radius 0.75*diag about the bbox centre, height
    +0.25*diag, `n_frames` equal azimuth steps.  Returns (position, target, znear, zfar)."""
    lo = np.asarray(bbox_min, dtype=np.float64)
    hi = np.asarray(bbox_max, dtype=np.float64)
    c = 0.5 * (lo + hi)
    diag = float(np.linalg.norm(hi - lo))
    az = 2.0 * math.pi * (frame % n_frames) / n_frames
    pos = c + np.array([0.75 * diag * math.cos(az), 0.25 * diag, 0.75 * diag * math.sin(az)])
    return pos.astype(np.float32), c.astype(np.float32), 0.1, 4.0 * diag


def bench_light(bbox_min, bbox_max) -> Tuple[float, float, float]:
    """Bench light of SURVEY §8d: bbox centre + (0, 0.4*height, 0)."""
    lo = np.asarray(bbox_min, dtype=np.float64)
    hi = np.asarray(bbox_max, dtype=np.float64)
    c = 0.5 * (lo + hi)
    return float(c[0]), float(c[1] + 0.4 * (hi[1] - lo[1])), float(c[2])


def room_lights(bbox_min, bbox_max) -> List[Tuple[float, float, float]]:
    """Config c2: four lights at the upper quarter-points of the room bbox."""
    lo = np.asarray(bbox_min, dtype=np.float64)
    hi = np.asarray(bbox_max, dtype=np.float64)
    y = lo[1] + 0.75 * (hi[1] - lo[1])
    out = []
    for fx in (0.25, 0.75):
        for fz in (0.25, 0.75):
            out.append((float(lo[0] + fx * (hi[0] - lo[0])), float(y), float(lo[2] + fz * (hi[2] - lo[2]))))
    return out
