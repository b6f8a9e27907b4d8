"""ctypes binding of librc_b200.so (include/rc_b200.h).

The library is the product; this module only declares its entry points.  If the
shared object is missing the import fails loudly — there is no Python or CPU
fallback for any compute call.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RC_B200_LIB") or os.path.join(_HERE, "librc_b200.so")   # override: A/B builds of the same ABI

RC_OK = 0
RC_ERR_INVALID_ARG, RC_ERR_SCENE_LOAD, RC_ERR_CUDA, RC_ERR_NO_DEVICE, RC_ERR_BUFFER_SIZE, RC_ERR_STATE = 1, 2, 3, 4, 5, 6
STATUS_NAMES = {0: "RC_OK", 1: "RC_ERR_INVALID_ARG", 2: "RC_ERR_SCENE_LOAD", 3: "RC_ERR_CUDA",
                4: "RC_ERR_NO_DEVICE", 5: "RC_ERR_BUFFER_SIZE", 6: "RC_ERR_STATE"}

RC_CFG_SEPARATE_MERGE = 0x1
RC_CFG_NO_TEXTURES = 0x2
RC_CFG_HALO_EXCHANGE = 0x4
RC_CFG_RASTER_CLIP = 0x8
RC_CFG_FLOATING_PROBES = 0x10
RC_UPD_ENABLE_NORMAL_MAP = 0x1

(RC_TARGET_IRRADIANCE, RC_TARGET_DIRECT, RC_TARGET_DEPTH, RC_TARGET_NORMAL, RC_TARGET_ALBEDO, RC_TARGET_PRIM,
 RC_TARGET_COMPOSITE, RC_TARGET_DIRECT_SRGB8) = range(8)
RC_TARGET_IRRADIANCE_RGB48 = 8
RC_TARGET_CASCADE0 = 16

STAGES = ("gbuffer", "probes", "march", "merge", "gather", "frame")


class rc_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("device", C.c_int32),
                ("scene_path", C.c_char_p), ("resource_root", C.c_char_p),
                ("probe_spacing0", C.c_uint32), ("dir_res0", C.c_uint32), ("num_levels", C.c_uint32),
                ("interval0", C.c_float), ("t_far", C.c_float), ("normal_offset", C.c_float), ("sky", C.c_float * 3),
                ("flags", C.c_uint32),
                ("tile_x0", C.c_uint32), ("tile_y0", C.c_uint32), ("tile_w", C.c_uint32), ("tile_h", C.c_uint32)]


class rc_camera(C.Structure):
    _fields_ = [("view_proj", C.c_float * 16), ("eye", C.c_float * 4)]


class rc_light(C.Structure):
    _fields_ = [("position", C.c_float * 4)]


class rc_level_info(C.Structure):
    _fields_ = [("spacing", C.c_uint32), ("dir_res", C.c_uint32), ("grid_w", C.c_uint32), ("grid_h", C.c_uint32),
                ("px0", C.c_int32), ("py0", C.c_int32), ("sub_w", C.c_uint32), ("sub_h", C.c_uint32),
                ("texel_offset", C.c_uint64), ("texel_count", C.c_uint64), ("t_begin", C.c_float), ("t_end", C.c_float)]


class rc_exchange_info(C.Structure):
    _fields_ = [("need_ptr", C.c_void_p), ("avg_ptr", C.c_void_p), ("need_words_per_probe", C.c_uint32), ("avg_float4_per_probe", C.c_uint32),
                ("px0", C.c_int32), ("py0", C.c_int32), ("sub_w", C.c_uint32), ("sub_h", C.c_uint32),
                ("own_x0", C.c_int32), ("own_y0", C.c_int32), ("own_x1", C.c_int32), ("own_y1", C.c_int32),
                ("exchanged", C.c_uint32), ("pad", C.c_uint32)]


class rc_scene_info(C.Structure):
    _fields_ = [("num_models", C.c_uint32), ("num_vertices", C.c_uint32), ("num_triangles", C.c_uint32),
                ("num_materials", C.c_uint32), ("num_textures", C.c_uint32), ("bvh_nodes", C.c_uint32),
                ("light_from_obj", C.c_uint32), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("obj_light", C.c_float * 3)]


# every symbol include/rc_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "rc_abi_version": (C.c_uint32, []),
    "rc_last_error": (C.c_char_p, [_P]),
    "rc_create": (C.c_int32, [C.POINTER(rc_config), C.POINTER(_P)]),
    "rc_destroy": (None, [_P]),
    "rc_update": (C.c_int32, [_P, C.POINTER(rc_camera), C.POINTER(rc_light), C.c_uint32, C.c_uint32]),
    "rc_resize": (C.c_int32, [_P, C.c_uint32, C.c_uint32]),
    "rc_set_tile": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rc_render_lists": (C.c_int32, [_P, _P]),
    "rc_exchange_level_info": (C.c_int32, [_P, C.c_uint32, C.POINTER(rc_exchange_info)]),
    "rc_render": (C.c_int32, [_P, _P]),
    "rc_render_begin": (C.c_int32, [_P, _P]),
    "rc_render_level": (C.c_int32, [_P, C.c_uint32, _P]),
    "rc_render_end": (C.c_int32, [_P, _P]),
    "rc_synchronize": (C.c_int32, [_P]),
    "rc_read_target": (C.c_int32, [_P, C.c_int, _P, C.c_size_t]),
    "rc_read_target_async": (C.c_int32, [_P, C.c_int, _P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "rc_read_wait": (C.c_int32, [_P, C.c_uint32]),
    "rc_target_bytes": (C.c_int32, [_P, C.c_int, C.POINTER(C.c_size_t)]),
    "rc_stage_times": (C.c_int32, [_P, C.POINTER(C.c_float), C.c_uint32]),
    "rc_level_times": (C.c_int32, [_P, C.POINTER(C.c_float), C.c_uint32]),
    "rc_set_tuning": (C.c_int32, [_P, C.c_char_p, C.c_int]),
    "rc_launch_count": (C.c_int32, [_P, C.POINTER(C.c_uint32)]),
    "rc_rays_marched": (C.c_int32, [_P, C.POINTER(C.c_uint32), C.c_uint32]),
    "rc_get_ray_list": (C.c_int32, [_P, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]),
    "rc_get_split_list": (C.c_int32, [_P, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rc_peer_export": (C.c_int32, [_P, C.c_void_p, C.c_size_t]),
    "rc_peer_attach": (C.c_int32, [_P, C.c_void_p, C.c_uint32, C.c_uint32]),
    "rc_peer_wait": (C.c_int32, [_P, C.c_void_p]),
    "rc_peer_frame": (C.c_int32, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]),
    "rc_get_levels": (C.c_int32, [_P, C.POINTER(rc_level_info), C.c_uint32, C.POINTER(C.c_uint32)]),
    "rc_get_scene_info": (C.c_int32, [_P, C.POINTER(rc_scene_info)]),
    "rc_get_tile": (C.c_int32, [_P, C.POINTER(C.c_uint32)]),
    "rc_get_intervals": (C.c_int32, [_P, C.POINTER(C.c_float)]),
    "rc_get_directions": (C.c_int32, [_P, C.c_uint32, _P, C.c_size_t]),
    "rc_get_model_stream": (C.c_int32, [_P, C.c_uint32, _P, C.c_size_t, _P, C.c_size_t,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rc_get_model_material": (C.c_int32, [_P, C.c_uint32, _P, C.c_size_t]),
    "rc_scene_load": (C.c_int32, [C.c_char_p, C.c_uint32, C.POINTER(_P)]),
    "rc_scene_free": (None, [_P]),
    "rc_scene_get_info": (C.c_int32, [_P, C.POINTER(rc_scene_info)]),
    "rc_scene_model_stream": (C.c_int32, [_P, C.c_uint32, _P, C.c_size_t, _P, C.c_size_t,
                                          C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rc_scene_model_material": (C.c_int32, [_P, C.c_uint32, _P, C.c_size_t]),
    "rc_scene_model_name": (C.c_int32, [_P, C.c_uint32, C.c_char_p, C.c_size_t]),
    "rc_scene_model_texture": (C.c_int32, [_P, C.c_uint32, C.c_uint32, _P, C.c_size_t,
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rc_decode_image_file": (C.c_int32, [C.c_char_p, _P, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rc_trace_rays": (C.c_int32, [_P, _P, C.c_uint32, _P]),
    "rc_shade_points": (C.c_int32, [_P, _P, C.c_uint32, _P]),
    "rc_cascade_device_ptr": (C.c_int32, [_P, C.c_uint32, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "rc_irradiance_device_ptr": (C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "rc_camera_view_matrix": (None, [C.POINTER(C.c_float), C.c_float, C.c_float, C.POINTER(C.c_float)]),
    "rc_projection_matrix": (None, [C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]),
    "rc_uniform_camera": (None, [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.POINTER(rc_camera)]),
    "rc_uniform_camera_look_at": (None, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float,
                                         C.c_float, C.POINTER(rc_camera)]),
}

_lib = None


def load() -> C.CDLL:
    """Load librc_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C radiancecascade_b200/csrc` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class RcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
