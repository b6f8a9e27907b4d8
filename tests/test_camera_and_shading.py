"""Camera / projection façade (a7), uniform layouts (a5, a8) and the two independent restatements of
fs_main (a10): numpy (oracle/ref_ingest.py) vs C (oracle/rc_oracle.c), pinned by golden vectors."""
import json
import math
import os

import numpy as np

import radiancecascade_b200 as rc
from oracle import gi_oracle as go
from oracle import ref_ingest as ri

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_reference_default_camera_golden():
    """AppState::new: Camera((0,5,10), -90, -20) used as radians, pitch clamped to -(pi/2 - 1e-4)
    (src/app.rs:25, src/camera.rs:194-198)."""
    gold = json.load(open(os.path.join(GOLD, "camera_golden.json")))
    st = rc.AppState()
    st.camera.clamp_pitch()
    assert abs(st.camera.pitch + (math.pi / 2 - 1e-4)) < 1e-6
    for label, (w, h) in {"default_1360x1360": (1360, 1360), "default_1360x768": (1360, 768)}.items():
        st.projection.resize(w, h)
        got = rc.UniformCamera.from_camera_project(st.camera, st.projection).as_array()
        assert got.tobytes().hex() == gold[label]["hex"]


def _close_ulps(a, b, ulps=8):
    """sin/cos differ by <= 1 ulp between libm's sinf (the C++ facade; what Rust's f32::sin_cos calls on
    Linux) and the oracle's double-rounded sin — a few ulps after the matrix product."""
    scale = np.maximum(np.abs(a), np.abs(b)).max()
    return bool(np.all(np.abs(a.astype(np.float64) - b.astype(np.float64)) <= ulps * 1.2e-7 * scale))


def test_camera_facade_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    for _ in range(50):
        pos = rng.normal(size=3) * 10
        yaw, pitch = rng.uniform(-3, 3), rng.uniform(-1.5, 1.5)
        proj = rc.Projection.new(int(rng.integers(64, 4000)), int(rng.integers(64, 3000)), float(rng.uniform(20, 90)), 0.1, float(rng.uniform(50, 1000)))
        a = rc.UniformCamera.from_camera_project(rc.Camera(tuple(pos), yaw, pitch), proj).as_array()
        b = ri.uniform_camera(pos, yaw, pitch, proj.fovy, proj.aspect, proj.znear, proj.zfar)
        assert _close_ulps(a, b)
        tgt = rng.normal(size=3)
        a = rc.UniformCamera.look_at(pos, tgt, proj).as_array()
        b = ri.uniform_camera_look_at(pos, tgt, proj.fovy, proj.aspect, proj.znear, proj.zfar)
        assert _close_ulps(a, b)


def test_orbit_camera_golden():
    gold = json.load(open(os.path.join(GOLD, "camera_golden.json")))
    for name in ("cube", "teapot"):
        g = gold[f"orbit5_{name}_1920x1080"]
        proj = rc.Projection.new(1920, 1080, 45.0, g["znear"], g["zfar"])
        got = rc.UniformCamera.look_at(g["position"], g["target"], proj).as_array()
        assert got.tobytes().hex() == g["hex"]


def test_perspective_conventions():
    """glam perspective_rh, depth 0..1: a point on the near plane maps to z=0, far plane to z=1, -z forward."""
    m = ri.projection_matrix(np.float32(math.radians(45)), 16 / 9, 0.1, 100.0).astype(np.float64)
    for z_view, z_ndc in ((-0.1, 0.0), (-100.0, 1.0)):
        clip = m @ np.array([0, 0, z_view, 1.0])
        assert abs(clip[2] / clip[3] - z_ndc) < 1e-5
    v = ri.look_to_rh((0, 0, 0), (0, 0, -1)).astype(np.float64)
    assert np.allclose(v, np.eye(4), atol=1e-7)


def test_fs_main_golden_and_two_restatements_agree():
    g = np.load(os.path.join(GOLD, "fs_main_golden.npz"))
    osc = go.OracleScene(rc.scenes.scene_path("cube"))
    pts, lights, want = g["points"], g["lights"], g["radiance"]
    got_c = osc.shade_points(pts, lights)
    assert np.allclose(got_c, want, rtol=0, atol=1e-6, equal_nan=True)
    # numpy restatement of src/shader.wgsl:76-100 at the same points
    prim = pts[:, 0].copy().view(np.uint32)
    u, v = pts[:, 1], pts[:, 2]
    w = (np.float32(1) - u) - v
    tri = osc.tris[prim]
    va, vb, vc = osc.verts[tri[:, 0]], osc.verts[tri[:, 1]], osc.verts[tri[:, 2]]
    at = va * w[:, None] + vb * u[:, None] + vc * v[:, None]
    mat = osc.materials[0]
    rad = ri.fs_main(at[:, 0:3], at[:, 3:6], at[:, 6:9], at[:, 9:12], at[:, 12:15], at[:, 15:17], ri.uniform_material(mat),
                     ri.enable_bit(mat), pts[:, 4:7], lights[:, :3], mat.color_texture, mat.normal_texture)
    rad = np.nan_to_num(np.clip(rad + mat.emission, 0, 65504), nan=0.0)
    ok = np.abs(rad - want[:, :3]) <= 5e-4 * np.maximum(1.0, np.abs(want[:, :3]))
    assert ok.mean() > 0.995      # texel-boundary flips of the nearest filter account for the rest


def test_fs_main_unlit_fallback_and_two_sided_normal():
    um = ri.uniform_material(None)     # sonic.obj: no material -> unlit -> output = albedo (white)
    out = ri.fs_main(np.zeros((1, 3)), np.ones((1, 3)), [[0, 0, 1]], [[1, 0, 0]], [[0, 1, 0]], np.zeros((1, 2)), um, 0,
                     np.array([0, 0, 5.0]), np.array([[0, 5.0, 0]]))
    assert np.allclose(out, 1.0)
    mat = ri.Material(np.float32([1, 1, 1]), np.float32([0.5, 0.5, 0.5]), np.float32([0, 0, 0]), np.float32(10), None, None, np.zeros(3, np.float32))
    um = ri.uniform_material(mat)
    front = ri.fs_main(np.zeros((1, 3)), np.ones((1, 3)), [[0, 0, 1]], [[1, 0, 0]], [[0, 1, 0]], np.zeros((1, 2)), um, 0,
                       np.array([0, 0, 5.0]), np.array([[0, 0, 3.0]]))
    back = ri.fs_main(np.zeros((1, 3)), np.ones((1, 3)), [[0, 0, -1]], [[1, 0, 0]], [[0, 1, 0]], np.zeros((1, 2)), um, 0,
                      np.array([0, 0, 5.0]), np.array([[0, 0, 3.0]]))
    assert np.allclose(front, back) and np.allclose(front, 0.05 + 0.7 * 0.5, atol=1e-6)   # ambient + diffuse, N flipped to the viewer


def test_srgb_helpers():
    t = ri.srgb_table()
    assert t[0] == 0 and abs(t[255] - 1) < 1e-7 and np.all(np.diff(t) > 0)
    assert np.array_equal(ri.srgb_encode_u8(t.astype(np.float64)), np.arange(256, dtype=np.uint8))
    idx = ri._mirror_index(np.array([-3, -1, 0, 3, 4, 7, 8]), 4)
    assert idx.tolist() == [2, 0, 0, 3, 3, 0, 0]      # MirrorRepeat (src/texture.rs:132-134)
