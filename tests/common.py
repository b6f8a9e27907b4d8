"""Shared helpers for the parity tests: build the same frame in the CUDA product
(through the C ABI) and in the CPU oracle."""
import numpy as np

import radiancecascade_b200 as rc
from radiancecascade_b200 import _ffi
from oracle import gi_oracle as go
from oracle import ref_ingest as ri

_oracle_cache = {}


def oracle_scene(name: str) -> go.OracleScene:
    if name not in _oracle_cache:
        _oracle_cache[name] = go.OracleScene(rc.scenes.scene_path(name))
    return _oracle_cache[name]


def frame_setup(name, W, H, frame=5, lights="bench"):
    """(AppState for the product, cam float32[20], lights float32[n][4]) for the SURVEY §8d orbit."""
    osc = oracle_scene(name)
    pos, tgt, zn, zf = rc.scenes.orbit_camera(osc.bbox_min, osc.bbox_max, frame)
    proj = rc.Projection.new(W, H, 45.0, zn, zf)
    uc = rc.UniformCamera.look_at(pos, tgt, proj)
    if lights == "bench":
        pts = [rc.scenes.bench_light(osc.bbox_min, osc.bbox_max)]
    elif lights == "room":
        pts = rc.scenes.room_lights(osc.bbox_min, osc.bbox_max)
    else:
        pts = [(0.0, 0.0, 0.0)]
    st = rc.AppState()
    st.uniform_camera = uc
    st.light_position = pts[0]
    st.extra_lights = pts[1:]
    larr = np.array([[p[0], p[1], p[2], 1.0] for p in pts], dtype=np.float32)
    return st, uc.as_array(), larr


def render_product(name, W, H, state, cascade=None, device=0):
    r = rc.DefaultRenderer.new(device, (W, H), state, rc.scenes.scene_path(name), cascade)
    r.update(state)
    r.render()
    return r


def half_to_f32(a):
    return np.asarray(a, dtype=np.float16).astype(np.float32)


def psnr(a, b, peak):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
