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


def _upper_pair(q, gmax):
    """rc_spec.h S1 (rc_device.cuh upper_pair): the two upper probes of probe index q along one axis, clamped."""
    base = np.where(q % 2 == 0, q // 2 - 1, (q - 1) // 2)
    return np.clip(base, 0, gmax - 1), np.clip(base + 1, 0, gmax - 1)


def predict_ray_lists(osc, p, out, gpu_depth, gpu_normal, n_lists):
    """numpy restatement of the direction-culling rule (kernels.cu k_gbuffer masks -> k_probes OR -> k_need push-up)
    for a FULL-frame context: sorted uint32 entries of every level's ray list.
      level 0 : entry = probe * D0^2 + texel;  requested iff some pixel the probe serves has dot(n, w_d) > 0
      level i : entry = probe * D_{i-1}^2 + quad;  R_1[k] |= R_0[p],  R_{i+1}[k] |= expand2x(R_i[p])  for every valid
                lower probe p and every VALID upper probe k among its four (clamped) bilinear neighbours
    `out` = OracleScene.render(...) (probe validity, level layout); the per-pixel masks use the product's own stored
    normal / depth (the oracle's encoded normal may differ by one snorm16 LSB, tests/test_gpu_parity.py::test_gbuffer)."""
    lv, rects = out["levels"], out["rects"]
    P, D0 = lv[0].P, lv[0].D
    H, W = gpu_depth.shape
    pm = osc.pixel_masks(p, dict(depth=gpu_depth, normal=gpu_normal))
    px0, py0, sw, sh = rects[0]
    assert (px0, py0) == (0, 0)
    need0 = np.zeros((sh, sw), np.uint32)
    for py in range(sh):
        ya, yb = max((py - 1) * P + P // 2, 0), min((py + 1) * P + P // 2, H)
        for px in range(sw):
            xa, xb = max((px - 1) * P + P // 2, 0), min((px + 1) * P + P // 2, W)
            if ya < yb and xa < xb:
                need0[py, px] = np.bitwise_or.reduce(pm[ya:yb, xa:xb], axis=None)
    bits = (need0[..., None] >> np.arange(D0 * D0, dtype=np.uint32)) & 1
    R = bits.astype(bool).reshape(sh * sw, D0 * D0)          # requests of level 0 at resolution D0
    lists = []
    for i in range(n_lists):
        _, _, sw, sh = rects[i]
        Dr = D0 if i == 0 else lv[i - 1].D
        valid = out["origins"][i][:, 3] != 0
        Ri = R & valid[:, None]
        probe, bit = np.nonzero(Ri)
        lists.append(np.sort((probe.astype(np.uint64) * (Dr * Dr) + bit).astype(np.uint32)))
        if i + 1 >= n_lists:
            break
        _, _, usw, ush = rects[i + 1]
        validu = out["origins"][i + 1][:, 3] != 0
        if i == 0:
            ex, UDr = Ri, Dr                                  # level 0 -> 1: same resolution
        else:
            ex = np.repeat(np.repeat(Ri.reshape(-1, Dr, Dr), 2, axis=1), 2, axis=2).reshape(-1, 4 * Dr * Dr)
            UDr = 2 * Dr
        py, px = np.divmod(np.arange(sw * sh), sw)
        x0, x1 = _upper_pair(px, usw)
        y0, y1 = _upper_pair(py, ush)
        Ru = np.zeros((usw * ush, UDr * UDr), bool)
        for u in (y0 * usw + x0, y0 * usw + x1, y1 * usw + x0, y1 * usw + x1):
            np.logical_or.at(Ru, u, ex)
        R = Ru & validu[:, None]
    return lists
