"""The C++ host side above the C ABI: include/rc_b200.hpp (the reference's type names — Camera, Projection,
UniformCamera, CameraController, AppState, RenderStage, DefaultRenderer, ObjScene; src/camera.rs, src/app.rs,
src/renderer.rs) and the headless frame driver radiancecascade_b200/rc_headless (≙ src/main.rs +
App::handle_redraw, src/window/app.rs:221-267).

CPU: the façade is compiled with -Wall -Wextra -Werror, run, and checked against the committed camera golden
vectors and a float32 restatement of CameraController::update_camera (src/camera.rs:170-199); the driver must
fail loudly without a GPU.  GPU: the driver's frames are bit-identical to the Python mirror's."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

import radiancecascade_b200 as rc
from radiancecascade_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "radiancecascade_b200")
DRIVER = os.path.join(PKG, "rc_headless")
F = np.float32


def _have_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="module")
def facade(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("facade") / "facade_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_check.cpp"), "-o", exe, "-L", PKG, "-lrc_b200",
                           "-Wl,-rpath," + PKG])
    out = subprocess.run([exe, rc.scenes.scene_path("cube")], check=True, capture_output=True, text=True).stdout
    return json.loads(out)


def test_facade_header_is_self_contained(tmp_path):
    src = tmp_path / "only_header.cpp"
    src.write_text('#include "rc_b200.hpp"\nint main() { return sizeof(rc::AppState) ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src)])


def test_facade_camera_matches_golden(facade):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "camera_golden.json")))
    assert abs(facade["default_pitch"] + (math.pi / 2 - 1e-4)) < 1e-6          # clamped by update_camera (src/camera.rs:194-198)
    for key in ("default_1360x1360", "default_1360x768", "orbit5_cube_1920x1080"):
        assert facade[key] == gold[key]["hex"], key
    # the façade and the Python mirror call the same library functions
    view = rc.Camera((1.0, 2.0, 3.0), 0.3, -0.2).calc_matrix()
    proj = rc.Projection.new(1920, 1080, 45.0, 0.1, 13.856406211853027).calc_matrix()
    assert facade["view"] == view.tobytes().hex() and facade["proj"] == proj.tobytes().hex()
    assert facade["light"] == np.array([1, 2, 3, 1], dtype=F).tobytes().hex()   # UniformLight::from(Vec3), src/primitives.rs:26-35


def _update_camera(pos, yaw, pitch, amounts, rot, scroll, speed, sens, dt):
    """float32 restatement of CameraController::update_camera (src/camera.rs:170-199)."""
    fwd_amt, right_amt, up_amt = (F(a) for a in amounts)
    speed, sens, dt = F(speed), F(sens), F(dt)

    def normalize(v):
        r = F(1.0) / np.sqrt(F(F(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]), dtype=F)
        return np.array([v[0] * r, v[1] * r, v[2] * r], dtype=F)

    ys, yc = F(math.sin(float(yaw))), F(math.cos(float(yaw)))        # libm sinf / cosf agree with the rounded double here
    fwd, right = normalize(np.array([yc, 0, ys], dtype=F)), normalize(np.array([-ys, 0, yc], dtype=F))
    pos = pos + ((fwd * fwd_amt) * speed) * dt
    pos = pos + ((right * right_amt) * speed) * dt
    ps, pc = F(math.sin(float(pitch))), F(math.cos(float(pitch)))
    toward = normalize(np.array([pc * yc, ps, pc * ys], dtype=F))
    pos = pos + (((toward * F(scroll)) * speed) * sens) * dt
    pos[1] = pos[1] + (up_amt * speed) * dt
    yaw = yaw + (F(rot[0]) * sens) * dt
    pitch = pitch + (-F(rot[1]) * sens) * dt
    lim = F(F(math.pi / 2) - F(0.0001))
    return pos.astype(F), F(yaw), F(min(max(pitch, -lim), lim))


def test_facade_camera_controller(facade):
    pos, yaw, pitch = np.array([0.5, 1.0, -2.0], dtype=F), F(0.7), F(0.1)
    for i in range(3):
        pos, yaw, pitch = _update_camera(pos, yaw, pitch, (1, 1, 1), (12.0, -7.0), -100.0 if i == 0 else 0.0, 4.0, 0.4, 1.0 / 60.0)
    want = np.array([pos[0], pos[1], pos[2], yaw, pitch], dtype=np.float64)
    got = np.array(facade["walk"], dtype=np.float64)
    assert np.all(np.abs(got - want) <= 4e-7 * np.maximum(1.0, np.abs(want))), (got, want)   # sinf vs rounded double sin: <= 1 ulp
    assert abs(facade["clamped_pitch"] - (math.pi / 2 - 1e-4)) < 1e-6


def test_facade_scene_ingest(facade):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ingest_golden.json")))["cube"]
    sc = rc.ObjScene.load(rc.scenes.scene_path("cube"))
    info = sc.info()
    v, i = sc.model_stream(0)
    assert (facade["models"], facade["vertices"], facade["triangles"]) == (info.num_models, info.num_vertices, info.num_triangles)
    assert facade["stream_floats"] == v.size and facade["indices"] == i.size
    assert facade["enable_bit"] == sc.model_material(0)[1] and facade["name"] == sc.model_name(0)
    assert facade["missing_scene_throws"] is True
    assert (facade["vertices"], facade["triangles"]) == (gold["total_vertices"], gold["total_triangles"])
    assert facade["name"] == gold["models"][0]["name"] and facade["enable_bit"] == gold["models"][0]["enable_bit"]


def test_facade_renderer_fails_loudly_without_gpu(facade):
    if _have_gpu():
        assert facade["renderer_status"] == _ffi.RC_OK and facade["launches"] > 0
    else:
        assert facade["renderer_status"] == _ffi.RC_ERR_NO_DEVICE and facade["launches"] == 0


def test_headless_driver_cli():
    assert os.path.exists(DRIVER), "run __graft_entry__.build() first"
    out = subprocess.run([DRIVER, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "--camera orbit" in out.stdout
    bad = subprocess.run([DRIVER, "--size", "64x64"], capture_output=True, text=True)
    assert bad.returncode == 2 and "--scene is required" in bad.stderr
    if not _have_gpu():    # no CPU fallback: rc::Error(RC_ERR_NO_DEVICE) -> exit code 3
        run = subprocess.run([DRIVER, "--scene", rc.scenes.scene_path("cube"), "--size", "64x64"], capture_output=True, text=True)
        assert run.returncode == 3 and "no usable CUDA device" in run.stderr and "rc_status 4" in run.stderr


@pytest.mark.gpu
def test_headless_driver_frames_equal_python_mirror(tmp_path):
    """Orbit frames 5 and 6 of the cube through rc_headless (C++ façade) and through the Python mirror: same
    library, same camera / light arithmetic -> bit-identical irradiance; PFM and PPM outputs decode to the same pixels."""
    W, H = 96, 64
    path = rc.scenes.scene_path("cube")
    pre = str(tmp_path / "f")
    run = subprocess.run([DRIVER, "--scene", path, "--size", f"{W}x{H}", "--frames", "2", "--first-frame", "5", "--warmup", "1",
                          "--out-raw", pre, "--out-image", pre], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    lines = [json.loads(l) for l in run.stdout.strip().splitlines()]
    assert [l.get("frame") for l in lines[:-1]] == [5, 6] and lines[-1]["summary"] and lines[-1]["kernel_launches"] > 0
    assert lines[-1]["mean_stage_ms"]["frame"] > 0

    st = rc.AppState()
    r = rc.DefaultRenderer.new(0, (W, H), st, path)
    info = r.scene_info()
    lo, hi = list(info.bbox_min), list(info.bbox_max)
    for frame in (5, 6):
        pos, tgt, zn, zf = rc.scenes.orbit_camera(lo, hi, frame)
        assert np.array_equal(np.array(lines[frame - 5]["eye"], dtype=np.float32), pos)   # %.9g round-trips a float32
        st.uniform_camera = rc.UniformCamera.look_at(pos, tgt, rc.Projection.new(W, H, 45.0, zn, zf))
        st.light_position = rc.scenes.bench_light(lo, hi)
        r.update(st)
        r.render()
        E = r.read_target(_ffi.RC_TARGET_IRRADIANCE)
        raw = np.fromfile(f"{pre}_{frame:04d}.bin", dtype=np.float16).reshape(H, W, 4)
        assert np.array_equal(raw.view(np.uint16), E.view(np.uint16))
        assert float(E[..., :3].max()) > 0
        with open(f"{pre}_{frame:04d}.pfm", "rb") as fh:
            assert fh.readline() == b"PF\n" and fh.readline() == f"{W} {H}\n".encode() and fh.readline() == b"-1.0\n"
            pfm = np.frombuffer(fh.read(), dtype="<f4").reshape(H, W, 3)[::-1]
        assert np.array_equal(pfm, E[..., :3].astype(np.float32))

    # the reference's own start-up state: AppState::new camera, light at the origin, composite target as PPM
    run = subprocess.run([DRIVER, "--scene", path, "--size", "64x64", "--camera", "default", "--light", "origin", "--walk", "1,0,0,3,0",
                          "--frames", "2", "--read", "composite", "--out-image", pre + "c", "--quiet"], capture_output=True, text=True)
    assert run.returncode == 0, run.stderr
    with open(pre + "c_0001.ppm", "rb") as fh:
        assert fh.readline() == b"P6\n" and fh.readline() == b"64 64\n" and fh.readline() == b"255\n"
        assert len(fh.read()) == 64 * 64 * 3
