"""Host-side mirror of the reference's renderer interface over the C ABI.

Names and argument meaning follow the reference so that tests read like tests
of the reference would: `Camera`, `Projection`, `UniformCamera`
(src/camera.rs:9-80), `AppState` (src/app.rs:9-37), and `DefaultRenderer` with
the `RenderStage` methods `update` / `resize` / `render`
(src/app.rs:3-7, src/renderer.rs:156-632).  All arithmetic and all rendering
happen inside librc_b200.so; nothing here computes pixels.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import RcError

SAFE_FRAC_PI_2 = np.float32(np.float32(math.pi / 2) - np.float32(0.0001))  # src/camera.rs:25


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@dataclass
class Camera:
    """src/camera.rs:27-53.  yaw / pitch are used as radians exactly as stored
    (the reference's AppState passes degrees: SURVEY Appendix B quirk 1)."""
    position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    yaw: float = 0.0
    pitch: float = 0.0

    def calc_matrix(self) -> np.ndarray:
        """Column-major 4x4 (glam Mat4::look_to_rh), as float32[16]."""
        out = np.zeros(16, dtype=np.float32)
        pos = np.asarray(self.position, dtype=np.float32)
        _ffi.load().rc_camera_view_matrix(_fp(pos), np.float32(self.yaw), np.float32(self.pitch), _fp(out))
        return out

    def clamp_pitch(self) -> None:
        """The clamp CameraController::update_camera applies before every frame (src/camera.rs:194-198)."""
        self.pitch = float(min(max(np.float32(self.pitch), -SAFE_FRAC_PI_2), SAFE_FRAC_PI_2))


@dataclass
class Projection:
    """src/camera.rs:55-80; `fovy` is given in DEGREES to new(), stored in radians."""
    aspect: float = 1.0
    fovy: float = math.radians(45.0)
    znear: float = 0.1
    zfar: float = 100.0

    @staticmethod
    def new(width: int, height: int, fovy_deg: float, znear: float, zfar: float) -> "Projection":
        return Projection(float(np.float32(width) / np.float32(height)), float(np.float32(np.radians(np.float32(fovy_deg)))),
                          znear, zfar)

    def resize(self, width: int, height: int) -> None:
        self.aspect = float(np.float32(width) / np.float32(height))

    def calc_matrix(self) -> np.ndarray:
        out = np.zeros(16, dtype=np.float32)
        _ffi.load().rc_projection_matrix(np.float32(self.fovy), np.float32(self.aspect), np.float32(self.znear),
                                         np.float32(self.zfar), _fp(out))
        return out


class UniformCamera:
    """src/camera.rs:9-23: 80 bytes = proj*view (column-major) + (eye, 1)."""

    def __init__(self, raw: _ffi.rc_camera):
        self.raw = raw

    @staticmethod
    def from_camera_project(camera: Camera, projection: Projection) -> "UniformCamera":
        raw = _ffi.rc_camera()
        pos = np.asarray(camera.position, dtype=np.float32)
        _ffi.load().rc_uniform_camera(_fp(pos), np.float32(camera.yaw), np.float32(camera.pitch),
                                      np.float32(projection.fovy), np.float32(projection.aspect),
                                      np.float32(projection.znear), np.float32(projection.zfar), C.byref(raw))
        return UniformCamera(raw)

    @staticmethod
    def look_at(position, target, projection: Projection) -> "UniformCamera":
        """Synthetic-path helper: same look_to_rh arithmetic with dir = normalize(target - position)."""
        raw = _ffi.rc_camera()
        pos = np.asarray(position, dtype=np.float32)
        tgt = np.asarray(target, dtype=np.float32)
        _ffi.load().rc_uniform_camera_look_at(_fp(pos), _fp(tgt), np.float32(projection.fovy), np.float32(projection.aspect),
                                              np.float32(projection.znear), np.float32(projection.zfar), C.byref(raw))
        return UniformCamera(raw)

    @staticmethod
    def from_array(a: Sequence[float]) -> "UniformCamera":
        raw = _ffi.rc_camera()
        a = np.asarray(a, dtype=np.float32).reshape(20)
        C.memmove(C.byref(raw), a.ctypes.data, 80)
        return UniformCamera(raw)

    def as_array(self) -> np.ndarray:
        return np.frombuffer(bytes(self.raw), dtype=np.float32).copy()


@dataclass
class AppState:
    """src/app.rs:9-37 minus the interactive controller / egui fields (out of scope)."""
    camera: Camera = field(default_factory=lambda: Camera((0.0, 5.0, 10.0), -90.0, -20.0))   # src/app.rs:25
    projection: Projection = field(default_factory=lambda: Projection.new(1, 1, 45.0, 0.1, 100.0))  # src/app.rs:26
    enable_normal_map: bool = True
    normal_map_changed: bool = False
    given_light_position: bool = False
    light_position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    extra_lights: List[Tuple[float, float, float]] = field(default_factory=list)  # beyond the reference's single light
    uniform_camera: Optional[UniformCamera] = None  # overrides camera/projection when set (look-at paths)


@dataclass
class CascadeConfig:
    """include/rc_spec.h parameters; zeros mean the defaults."""
    probe_spacing0: int = 0
    dir_res0: int = 0
    num_levels: int = 0
    interval0: float = 0.0
    t_far: float = 0.0
    normal_offset: float = 0.0
    sky: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    flags: int = 0
    tile: Optional[Tuple[int, int, int, int]] = None  # x0, y0, w, h


class DefaultRenderer:
    """≙ DefaultRenderer (src/renderer.rs:156-632) behind rc_create/rc_update/rc_resize/rc_render."""

    def __init__(self, handle, width: int, height: int, path: str):
        self._h = handle
        self.width, self.height, self.path = width, height, path
        self._lib = _ffi.load()

    # DefaultRenderer::new(device, config, queue, state, path)   src/renderer.rs:168-174
    @staticmethod
    def new(device: int, config: Tuple[int, int], state: AppState, path: str,
            cascade: Optional[CascadeConfig] = None, resource_root: Optional[str] = None) -> "DefaultRenderer":
        lib = _ffi.load()
        cc = cascade or CascadeConfig()
        cfg = _ffi.rc_config()
        cfg.struct_size = C.sizeof(_ffi.rc_config)
        cfg.width, cfg.height = int(config[0]), int(config[1])
        cfg.device = int(device)
        cfg.scene_path = path.encode()
        cfg.resource_root = resource_root.encode() if resource_root else None
        cfg.probe_spacing0, cfg.dir_res0, cfg.num_levels = cc.probe_spacing0, cc.dir_res0, cc.num_levels
        cfg.interval0, cfg.t_far, cfg.normal_offset = cc.interval0, cc.t_far, cc.normal_offset
        cfg.sky = (C.c_float * 3)(*cc.sky)
        cfg.flags = cc.flags
        if cc.tile:
            cfg.tile_x0, cfg.tile_y0, cfg.tile_w, cfg.tile_h = cc.tile
        h = C.c_void_p()
        st = lib.rc_create(C.byref(cfg), C.byref(h))
        if st != _ffi.RC_OK:
            raise RcError(st, (lib.rc_last_error(None) or b"").decode())
        r = DefaultRenderer(h, cfg.width, cfg.height, path)
        info = r.scene_info()
        state.given_light_position = bool(info.light_from_obj)   # src/renderer.rs:177
        return r

    def _check(self, st: int) -> None:
        if st != _ffi.RC_OK:
            raise RcError(st, (self._lib.rc_last_error(self._h) or b"").decode())

    def close(self) -> None:
        if self._h:
            self._lib.rc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # RenderStage::update + the two queue.write_buffer calls of AppInternal::update (src/window/app.rs:112-131)
    def update(self, state: AppState) -> None:
        uc = state.uniform_camera or UniformCamera.from_camera_project(state.camera, state.projection)
        pts = [state.light_position] + list(state.extra_lights)
        lights = (_ffi.rc_light * len(pts))()
        for i, p in enumerate(pts):
            lights[i].position = (C.c_float * 4)(p[0], p[1], p[2], 1.0)   # UniformLight::from(Vec3)
        flags = _ffi.RC_UPD_ENABLE_NORMAL_MAP if state.enable_normal_map else 0
        self._check(self._lib.rc_update(self._h, C.byref(uc.raw), lights, len(pts), flags))
        state.normal_map_changed = False

    def pack_update(self, state: AppState):
        """The arguments of rc_update for `state`, built ahead of time: (rc_camera, rc_light[n], n, flags)."""
        uc = state.uniform_camera or UniformCamera.from_camera_project(state.camera, state.projection)
        pts = [state.light_position] + list(state.extra_lights)
        lights = (_ffi.rc_light * len(pts))()
        for i, p in enumerate(pts):
            lights[i].position = (C.c_float * 4)(p[0], p[1], p[2], 1.0)
        return uc.raw, lights, len(pts), (_ffi.RC_UPD_ENABLE_NORMAL_MAP if state.enable_normal_map else 0)

    def update_packed(self, packed) -> None:
        """rc_update with arguments from pack_update (the per-frame host -> device write, nothing else)."""
        cam, lights, n, flags = packed
        self._check(self._lib.rc_update(self._h, C.byref(cam), lights, n, flags))

    # RenderStage::resize (src/renderer.rs:615-618)
    def resize(self, width: int, height: int) -> None:
        self._check(self._lib.rc_resize(self._h, width, height))
        self.width, self.height = width, height

    def set_tile(self, tile: Optional[Tuple[int, int, int, int]]) -> None:
        """Move this context's screen-space tile (x0, y0, w, h) inside the frame; None = the full frame (rc_set_tile)."""
        x0, y0, w, h = tile if tile else (0, 0, 0, 0)
        self._check(self._lib.rc_set_tile(self._h, x0, y0, w, h))

    # RenderStage::render (src/renderer.rs:559-613): enqueue only
    def render(self, stream: Optional[int] = None) -> None:
        self._check(self._lib.rc_render(self._h, C.c_void_p(stream) if stream else None))

    def render_begin(self, stream=None) -> None:
        self._check(self._lib.rc_render_begin(self._h, C.c_void_p(stream) if stream else None))

    def render_lists(self, stream=None) -> None:
        """Halo exchange: build the ray lists of the owned probes from the (completed) request masks."""
        self._check(self._lib.rc_render_lists(self._h, C.c_void_p(stream) if stream else None))

    def exchange_level_info(self, level: int) -> "_ffi.rc_exchange_info":
        out = _ffi.rc_exchange_info()
        self._check(self._lib.rc_exchange_level_info(self._h, level, C.byref(out)))
        return out

    def render_level(self, level: int, stream=None) -> None:
        self._check(self._lib.rc_render_level(self._h, level, C.c_void_p(stream) if stream else None))

    def render_end(self, stream=None) -> None:
        self._check(self._lib.rc_render_end(self._h, C.c_void_p(stream) if stream else None))

    def synchronize(self) -> None:
        self._check(self._lib.rc_synchronize(self._h))

    # ---- read-back ------------------------------------------------------
    _DTYPES = {_ffi.RC_TARGET_IRRADIANCE: (np.float16, 4), _ffi.RC_TARGET_DIRECT: (np.float16, 4),
               _ffi.RC_TARGET_ALBEDO: (np.float16, 4), _ffi.RC_TARGET_DEPTH: (np.float32, 1),
               _ffi.RC_TARGET_NORMAL: (np.uint32, 1), _ffi.RC_TARGET_PRIM: (np.uint32, 1),
               _ffi.RC_TARGET_COMPOSITE: (np.uint8, 4), _ffi.RC_TARGET_DIRECT_SRGB8: (np.uint8, 4),
               _ffi.RC_TARGET_IRRADIANCE_RGB48: (np.uint16, 3)}

    def target_bytes(self, which: int) -> int:
        n = C.c_size_t()
        self._check(self._lib.rc_target_bytes(self._h, which, C.byref(n)))
        return n.value

    def read_target(self, which: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        nbytes = self.target_bytes(which)
        if which >= _ffi.RC_TARGET_CASCADE0:
            dtype, ch = np.float16, 4
            shape = (nbytes // 8, 4)
        else:
            dtype, ch = self._DTYPES[which]
            tw, th = self.tile_size()
            shape = (th, tw, ch) if ch > 1 else (th, tw)
        if out is None:
            out = np.empty(shape, dtype=dtype)
        assert out.nbytes >= nbytes and out.flags["C_CONTIGUOUS"]
        self._check(self._lib.rc_read_target(self._h, which, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def read_irradiance_async(self, host_ptr: int, nbytes: int, rgb48: bool = False) -> int:
        """Pipelined read-back into page-locked memory at `host_ptr`; returns a ticket for read_wait().
        rgb48: the 6-byte-per-pixel target RC_TARGET_IRRADIANCE_RGB48 instead of RGBA16F."""
        t = C.c_uint32()
        which = _ffi.RC_TARGET_IRRADIANCE_RGB48 if rgb48 else _ffi.RC_TARGET_IRRADIANCE
        self._check(self._lib.rc_read_target_async(self._h, which, C.c_void_p(host_ptr), nbytes, C.byref(t)))
        return t.value

    @staticmethod
    def unpack_rgb48(a: np.ndarray) -> np.ndarray:
        """RC_TARGET_IRRADIANCE_RGB48 (uint16 [h][w][3]) -> float16 RGBA [h][w][4], bit-exact."""
        a = np.asarray(a, dtype=np.uint16)
        out = np.zeros(a.shape[:-1] + (4,), dtype=np.uint16)
        covered = (a[..., 0] & 0x8000) == 0
        out[..., 0] = a[..., 0] & 0x7FFF
        out[..., 1:3] = a[..., 1:3]
        out[..., 3] = np.where(covered, np.uint16(0x3C00), np.uint16(0))
        return out.view(np.float16)

    def read_wait(self, ticket: int) -> None:
        self._check(self._lib.rc_read_wait(self._h, ticket))

    def read_cascade(self, level: int) -> np.ndarray:
        return self.read_target(_ffi.RC_TARGET_CASCADE0 + level)

    def tile(self) -> Tuple[int, int, int, int]:
        t = (C.c_uint32 * 4)()
        self._check(self._lib.rc_get_tile(self._h, t))
        return int(t[0]), int(t[1]), int(t[2]), int(t[3])

    def tile_size(self) -> Tuple[int, int]:
        t = self.tile()
        return t[2], t[3]

    def intervals(self) -> Tuple[float, float, float]:
        """(L0, t_far, probe normal offset) in use — rc_spec.h S2 / S6."""
        t = (C.c_float * 3)()
        self._check(self._lib.rc_get_intervals(self._h, t))
        return float(t[0]), float(t[1]), float(t[2])

    def stage_times(self) -> Dict[str, float]:
        ms = (C.c_float * len(_ffi.STAGES))()
        self._check(self._lib.rc_stage_times(self._h, ms, len(_ffi.STAGES)))
        return dict(zip(_ffi.STAGES, [float(x) for x in ms]))

    def level_times(self) -> List[float]:
        n = len(self.levels())
        ms = (C.c_float * n)()
        self._check(self._lib.rc_level_times(self._h, ms, n))
        return [float(x) for x in ms]

    def set_tuning(self, key: str, value: int) -> None:
        self._check(self._lib.rc_set_tuning(self._h, key.encode(), int(value)))

    # ---- tiled multi-GPU: final-image exchange through NVLink peer memory (include/rc_b200.h rc_peer_*)
    def peer_export(self) -> bytes:
        h = C.create_string_buffer(64)
        self._check(self._lib.rc_peer_export(self._h, h, 64))
        return h.raw

    def peer_attach(self, handles: List[bytes], rank: int) -> None:
        blob = C.create_string_buffer(b"".join(handles), 64 * len(handles))
        self._check(self._lib.rc_peer_attach(self._h, blob, len(handles), rank))

    def peer_wait(self, stream: Optional[int] = None) -> None:
        self._check(self._lib.rc_peer_wait(self._h, C.c_void_p(stream) if stream else None))

    def peer_frame(self, check: bool = False) -> Tuple[int, int, int]:
        """(device pointer, bytes, timed-out waits) of the assembled full frame."""
        p, n, t = C.c_void_p(), C.c_size_t(), C.c_uint32(0)
        self._check(self._lib.rc_peer_frame(self._h, C.byref(p), C.byref(n), C.byref(t) if check else None))
        return int(p.value), int(n.value), int(t.value)

    def rays_marched(self) -> List[Optional[int]]:
        """Texels each level of the last frame actually marched (None: the level was not culled)."""
        n = len(self.levels())
        v = (C.c_uint32 * n)()
        self._check(self._lib.rc_rays_marched(self._h, v, n))
        return [None if int(x) == 0xFFFFFFFF else int(x) for x in v]

    def ray_list(self, level: int) -> np.ndarray:
        """The level's ray list of the last culled frame (uint32 entries, unordered; include/rc_b200.h rc_get_ray_list)."""
        n = C.c_uint32()
        self._check(self._lib.rc_get_ray_list(self._h, level, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            self._check(self._lib.rc_get_ray_list(self._h, level, out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(n)))
        return out

    def split_list(self, level: int) -> Tuple[np.ndarray, np.ndarray]:
        """(entries that are traversed, entries classified as certain misses) of the level's list in the last frame
        (include/rc_b200.h rc_get_split_list; RcError RC_ERR_STATE when the frame did not classify the level)."""
        ne, nm = C.c_uint32(), C.c_uint32()
        self._check(self._lib.rc_get_split_list(self._h, level, None, 0, C.byref(ne), C.byref(nm)))
        out = np.zeros(ne.value + nm.value, dtype=np.uint32)
        if out.size:
            self._check(self._lib.rc_get_split_list(self._h, level, out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(ne), C.byref(nm)))
        return out[:ne.value], out[ne.value:]

    def launch_count(self) -> int:
        n = C.c_uint32()
        self._check(self._lib.rc_launch_count(self._h, C.byref(n)))
        return n.value

    def levels(self) -> List[_ffi.rc_level_info]:
        n = C.c_uint32()
        arr = (_ffi.rc_level_info * 16)()
        self._check(self._lib.rc_get_levels(self._h, arr, 16, C.byref(n)))
        return [arr[i] for i in range(n.value)]

    def scene_info(self) -> _ffi.rc_scene_info:
        info = _ffi.rc_scene_info()
        self._check(self._lib.rc_get_scene_info(self._h, C.byref(info)))
        return info

    def directions(self, level: int) -> np.ndarray:
        D = self.levels()[level].dir_res
        out = np.zeros((D * D, 3), dtype=np.float32)
        self._check(self._lib.rc_get_directions(self._h, level, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def model_stream(self, model: int) -> Tuple[np.ndarray, np.ndarray]:
        """(vertex stream float32[nv][17], index buffer uint32[ni]) — src/renderer.rs:371-420."""
        nv, ni = C.c_uint32(), C.c_uint32()
        self._check(self._lib.rc_get_model_stream(self._h, model, None, 0, None, 0, C.byref(nv), C.byref(ni)))
        v = np.zeros((nv.value, 17), dtype=np.float32)
        i = np.zeros(ni.value, dtype=np.uint32)
        self._check(self._lib.rc_get_model_stream(self._h, model, v.ctypes.data_as(C.c_void_p), v.nbytes,
                                                  i.ctypes.data_as(C.c_void_p), i.nbytes, C.byref(nv), C.byref(ni)))
        return v, i

    def model_material(self, model: int) -> Tuple[np.ndarray, int, np.ndarray]:
        """(UniformMaterial float32[16], enable_bit, Ke float32[3])."""
        raw = np.zeros(20, dtype=np.float32)
        self._check(self._lib.rc_get_model_material(self._h, model, raw.ctypes.data_as(C.c_void_p), raw.nbytes))
        return raw[:16].copy(), int(raw[16:17].view(np.uint32)[0]), raw[17:20].copy()

    def trace_rays(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 4), dtype=np.float32)
        self._check(self._lib.rc_trace_rays(self._h, rays.ctypes.data_as(C.c_void_p), len(rays), hits.ctypes.data_as(C.c_void_p)))
        return hits

    def shade_points(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 8)
        out = np.zeros((len(pts), 4), dtype=np.float32)
        self._check(self._lib.rc_shade_points(self._h, pts.ctypes.data_as(C.c_void_p), len(pts), out.ctypes.data_as(C.c_void_p)))
        return out

    def cascade_device_ptr(self, level: int) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.rc_cascade_device_ptr(self._h, level, C.byref(p), C.byref(n)))
        return p.value, n.value

    def irradiance_device_ptr(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.rc_irradiance_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value


class ObjScene:
    """Device-free view of the scene ingest (≙ ObjScene::load + the per-model preparation
    of DefaultRenderer::new, src/primitives.rs:122-175, src/renderer.rs:370-497)."""

    def __init__(self, handle):
        self._h = handle
        self._lib = _ffi.load()

    @staticmethod
    def load(path: str, no_textures: bool = False) -> "ObjScene":
        lib = _ffi.load()
        h = C.c_void_p()
        st = lib.rc_scene_load(path.encode(), _ffi.RC_CFG_NO_TEXTURES if no_textures else 0, C.byref(h))
        if st != _ffi.RC_OK:
            raise RcError(st, (lib.rc_last_error(None) or b"").decode())
        return ObjScene(h)

    def close(self) -> None:
        if self._h:
            self._lib.rc_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> _ffi.rc_scene_info:
        info = _ffi.rc_scene_info()
        st = self._lib.rc_scene_get_info(self._h, C.byref(info))
        if st != _ffi.RC_OK:
            raise RcError(st, "rc_scene_get_info")
        return info

    def model_stream(self, model: int) -> Tuple[np.ndarray, np.ndarray]:
        nv, ni = C.c_uint32(), C.c_uint32()
        self._lib.rc_scene_model_stream(self._h, model, None, 0, None, 0, C.byref(nv), C.byref(ni))
        v = np.zeros((nv.value, 17), dtype=np.float32)
        i = np.zeros(ni.value, dtype=np.uint32)
        st = self._lib.rc_scene_model_stream(self._h, model, v.ctypes.data_as(C.c_void_p), v.nbytes,
                                             i.ctypes.data_as(C.c_void_p), i.nbytes, C.byref(nv), C.byref(ni))
        if st != _ffi.RC_OK:
            raise RcError(st, "rc_scene_model_stream")
        return v, i

    def model_material(self, model: int) -> Tuple[np.ndarray, int, np.ndarray]:
        raw = np.zeros(20, dtype=np.float32)
        st = self._lib.rc_scene_model_material(self._h, model, raw.ctypes.data_as(C.c_void_p), raw.nbytes)
        if st != _ffi.RC_OK:
            raise RcError(st, "rc_scene_model_material")
        return raw[:16].copy(), int(raw[16:17].view(np.uint32)[0]), raw[17:20].copy()

    def model_name(self, model: int) -> str:
        buf = C.create_string_buffer(512)
        self._lib.rc_scene_model_name(self._h, model, buf, 512)
        return buf.value.decode(errors="replace")

    def model_texture(self, model: int, which: int) -> Optional[np.ndarray]:
        w, h = C.c_uint32(), C.c_uint32()
        self._lib.rc_scene_model_texture(self._h, model, which, None, 0, C.byref(w), C.byref(h))
        if w.value == 0:
            return None
        out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
        self._lib.rc_scene_model_texture(self._h, model, which, out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(w), C.byref(h))
        return out


def decode_image_file(path: str) -> np.ndarray:
    """RGBA8 pixels of a PNG / JPEG file through librc_b200's built-in decoders (csrc/image.cpp, csrc/jpeg.cpp)."""
    lib = _ffi.load()
    w, h = C.c_uint32(), C.c_uint32()
    st = lib.rc_decode_image_file(path.encode(), None, 0, C.byref(w), C.byref(h))
    if st != _ffi.RC_OK:
        raise RcError(st, (lib.rc_last_error(None) or b"").decode())
    out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
    st = lib.rc_decode_image_file(path.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(w), C.byref(h))
    if st != _ffi.RC_OK:
        raise RcError(st, "rc_decode_image_file")
    return out
