"""ctypes driver of the C GI oracle (oracle/rc_oracle.c) fed by the Python
ingest oracle (oracle/ref_ingest.py).

TEST INFRASTRUCTURE ONLY — see the headers of those two files.  PARITY UNPINNED
(no reference GI path exists): this checks the CUDA product against
an independent restatement of include/rc_spec.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import ref_ingest as ri

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.rco_scene_create.restype = C.c_void_p
        _LIB.rco_num_threads.restype = C.c_int
    return _LIB


class Params(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("P0", C.c_int), ("D0", C.c_int), ("N", C.c_int),
                ("L0", C.c_float), ("t_far", C.c_float), ("offset", C.c_float), ("sky", C.c_float * 3),
                ("tile_x0", C.c_int), ("tile_y0", C.c_int), ("tile_w", C.c_int), ("tile_h", C.c_int),
                ("store_half", C.c_int), ("clip", C.c_int), ("floating", C.c_int)]


class Level(C.Structure):
    _fields_ = [("P", C.c_int), ("D", C.c_int), ("gw", C.c_int), ("gh", C.c_int), ("t0", C.c_float), ("t1", C.c_float)]


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def default_intervals(bbox_min, bbox_max) -> Tuple[np.float32, np.float32, np.float32, np.float32]:
    """rc_spec.h S2/S6 defaults in float32: diag, L0 = diag/256, t_far = 4*diag, offset = L0/16."""
    F = np.float32
    d = (np.asarray(bbox_max, dtype=F) - np.asarray(bbox_min, dtype=F)).astype(F)
    diag = F(np.sqrt(F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))))
    L0 = F(diag / F(256.0))
    return diag, L0, F(F(4.0) * diag), F(L0 / F(16.0))


def level_rects(W: int, H: int, P0: int, N: int, tile: Tuple[int, int, int, int]) -> List[Tuple[int, int, int, int]]:
    """Probe sub-grid (px0, py0, sw, sh) each level must hold so that the tile's
    pixels can be gathered and every level merged (rc_spec.h S1 footprints)."""
    x0, y0, tw, th = tile

    def gather_range(a0, a1, g):   # pixels [a0, a1] -> level-0 probes
        lo = (a0 - P0 // 2) // P0
        hi = (a1 - P0 // 2) // P0 + 1
        return max(0, min(lo, g - 1)), max(0, min(hi, g - 1))

    def up_range(lo, hi, g):       # probes [lo, hi] at level i -> level i+1
        def pair(q):
            base = q // 2 - 1 if q % 2 == 0 else (q - 1) // 2
            return max(0, min(base, g - 1)), max(0, min(base + 1, g - 1))
        return pair(lo)[0], pair(hi)[1]

    rects = []
    P = P0
    gw, gh = -(-W // P), -(-H // P)
    xr = gather_range(x0, x0 + tw - 1, gw)
    yr = gather_range(y0, y0 + th - 1, gh)
    rects.append((xr[0], yr[0], xr[1] - xr[0] + 1, yr[1] - yr[0] + 1))
    for i in range(1, N):
        P = P0 << i
        gw, gh = -(-W // P), -(-H // P)
        xr = up_range(xr[0], xr[1], gw)
        yr = up_range(yr[0], yr[1], gh)
        rects.append((xr[0], yr[0], xr[1] - xr[0] + 1, yr[1] - yr[0] + 1))
    return rects


class OracleScene:
    """Flattened scene arrays + the C oracle's scene handle."""

    def __init__(self, obj_path: str, use_textures: bool = True):
        self.models, self.obj_light = ri.ObjScene.load(obj_path)
        verts, tris, tri_model, umat, ebit, ke, tex_id = [], [], [], [], [], [], []
        tex_blobs: List[np.ndarray] = []
        tex_wh: List[Tuple[int, int]] = []
        tex_cache: Dict[int, int] = {}
        self.streams, self.index_buffers, self.materials = [], [], []
        voff = 0
        for m, sc in enumerate(self.models):
            vs = sc.vertex_stream()
            ix = sc.indices()
            self.streams.append(vs)
            self.index_buffers.append(ix)
            verts.append(vs)
            tris.append(ix.reshape(-1, 3).astype(np.uint32) + np.uint32(voff))
            tri_model.append(np.full(len(ix) // 3, m, dtype=np.uint32))
            voff += len(vs)
            mat = sc.material(decode_textures=use_textures)
            self.materials.append(mat)
            umat.append(ri.uniform_material(mat))
            ebit.append(ri.enable_bit(mat))
            ke.append(mat.emission if mat is not None else np.zeros(3, np.float32))
            ids = [-1, -1]
            if mat is not None:
                for j, t in enumerate((mat.color_texture, mat.normal_texture)):
                    if t is not None:
                        ids[j] = len(tex_blobs)
                        tex_blobs.append(np.ascontiguousarray(t))
                        tex_wh.append((t.shape[1], t.shape[0]))
            tex_id.append(ids)
        self.verts = np.ascontiguousarray(np.concatenate(verts), dtype=np.float32)
        self.tris = np.ascontiguousarray(np.concatenate(tris), dtype=np.uint32)
        self.tri_model = np.ascontiguousarray(np.concatenate(tri_model), dtype=np.uint32)
        self.umat = np.ascontiguousarray(np.stack(umat), dtype=np.float32)
        self.ebit = np.asarray(ebit, dtype=np.uint32)
        self.ke = np.ascontiguousarray(np.stack(ke), dtype=np.float32)
        self.tex_id = np.asarray(tex_id, dtype=np.int32)
        offs, blob, o = [], [], 0
        for t in tex_blobs:
            offs.append(o)
            blob.append(t.reshape(-1))
            o += t.size
        self.tex_data = np.concatenate(blob).astype(np.uint8) if blob else np.zeros(4, np.uint8)
        self.tex_off = np.asarray(offs if offs else [0], dtype=np.uint64)
        self.tex_wh = np.asarray(tex_wh if tex_wh else [(1, 1)], dtype=np.uint32)
        self.n_tex = len(tex_blobs)
        pos = self.verts[:, :3]
        self.bbox_min = pos.min(0)
        self.bbox_max = pos.max(0)
        self.handle = C.c_void_p(lib().rco_scene_create(
            C.c_int(len(self.verts)), _p(self.verts), C.c_int(len(self.tris)), _p(self.tris), _p(self.tri_model),
            C.c_int(len(self.models)), _p(self.umat), _p(self.ebit), _p(self.ke), _p(self.tex_id),
            C.c_int(self.n_tex), _p(self.tex_data), _p(self.tex_off), _p(self.tex_wh)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().rco_scene_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- queries -----------------------------------------------------------
    def trace(self, rays: np.ndarray, brute: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros((len(rays), 4), dtype=np.float32)
        lib().rco_trace(self.handle, _p(rays), C.c_int(len(rays)), _p(hits), C.c_int(int(brute)))
        return hits

    def shade_points(self, pts: np.ndarray, lights: np.ndarray, flags: int = 1) -> np.ndarray:
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 8)
        lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1, 4)
        out = np.zeros((len(pts), 4), dtype=np.float32)
        lib().rco_shade_points(self.handle, _p(pts), C.c_int(len(pts)), _p(lights), C.c_int(len(lights)),
                               C.c_uint32(flags), _p(out))
        return out

    def params(self, W, H, P0=4, D0=4, N=6, L0=0.0, t_far=0.0, offset=0.0, sky=(0, 0, 0),
               tile=None, store_half=True, clip=False, floating=False) -> Params:
        diag, dL0, dfar, _ = default_intervals(self.bbox_min, self.bbox_max)
        L0 = np.float32(L0) if L0 > 0 else dL0
        t_far = np.float32(t_far) if t_far > 0 else dfar
        offset = np.float32(offset) if offset > 0 else np.float32(L0 / np.float32(16.0))
        tile = tile or (0, 0, W, H)
        return Params(W, H, P0, D0, N, L0, t_far, offset, (C.c_float * 3)(*sky), *tile, int(store_half), int(clip), int(floating))

    @staticmethod
    def levels(p: Params) -> List[Level]:
        out = []
        for i in range(p.N):
            L = Level()
            lib().rco_level_layout(C.byref(p), C.c_int(i), C.byref(L))
            out.append(L)
        return out

    @staticmethod
    def directions(D: int) -> np.ndarray:
        out = np.zeros((D * D, 3), dtype=np.float32)
        lib().rco_directions(C.c_int(D), _p(out))
        return out

    @staticmethod
    def primary_basis(cam: np.ndarray) -> np.ndarray:
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        out = np.zeros(9, dtype=np.float32)
        lib().rco_primary_basis(_p(cam), _p(out))
        return out

    def gbuffer(self, p: Params, cam, lights, flags=1):
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1, 4)
        n = p.tile_w * p.tile_h
        depth = np.zeros(n, np.float32); prim = np.zeros(n, np.uint32); normal = np.zeros(n, np.uint32)
        albedo = np.zeros((n, 4), np.float32); direct = np.zeros((n, 4), np.float32)
        lib().rco_gbuffer(self.handle, C.byref(p), _p(cam), _p(lights), C.c_int(len(lights)), C.c_uint32(flags),
                          _p(depth), _p(prim), _p(normal), _p(albedo), _p(direct))
        sh = (p.tile_h, p.tile_w)
        return dict(depth=depth.reshape(sh), prim=prim.reshape(sh), normal=normal.reshape(sh),
                    albedo=albedo.reshape(sh + (4,)), direct=direct.reshape(sh + (4,)))

    def raster(self, p: Params, cam):
        """The reference's render pass restated as a rasteriser (rc_oracle.c rco_raster): near / far clipping, depth test Less,
        draw order.  Returns prim [h][w], zndc [h][w] (1.0 = clear), bary [h][w][2]."""
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        n = p.tile_w * p.tile_h
        prim = np.zeros(n, np.uint32); z = np.zeros(n, np.float32); bary = np.zeros((n, 2), np.float32)
        lib().rco_raster(self.handle, C.byref(p), _p(cam), _p(prim), _p(z), _p(bary))
        sh = (p.tile_h, p.tile_w)
        return dict(prim=prim.reshape(sh), zndc=z.reshape(sh), bary=bary.reshape(sh + (2,)))

    def pixel_masks(self, p: Params, gb) -> np.ndarray:
        """uint32[tile_h][tile_w]: level-0 directions with a positive cosine at each pixel (direction culling, S9)."""
        lv0 = self.levels(p)[0]
        d0 = self.directions(lv0.D)
        depth = np.ascontiguousarray(gb["depth"], np.float32).reshape(-1)
        normal = np.ascontiguousarray(gb["normal"], np.uint32).reshape(-1)
        out = np.zeros(p.tile_w * p.tile_h, np.uint32)
        lib().rco_pixel_masks(C.byref(p), _p(d0), _p(depth), _p(normal), _p(out))
        return out.reshape(p.tile_h, p.tile_w)

    def render(self, p: Params, cam, lights, flags=1, keep_raw=False, want_hits=False):
        """Full frame per rc_spec.h: G-buffer, probes, march, top-down merge, gather."""
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1, 4)
        L = lib()
        gb = self.gbuffer(p, cam, lights, flags)
        rects = level_rects(p.W, p.H, p.P0, p.N, (p.tile_x0, p.tile_y0, p.tile_w, p.tile_h))
        lv = self.levels(p)
        origins, nrms, casc, dirs, raws, hits = [], [], [], [], [], []
        for i in range(p.N):
            px0, py0, sw, sh = rects[i]
            og = np.zeros((sw * sh, 4), np.float32); nr = np.zeros((sw * sh, 4), np.float32)
            L.rco_probes(self.handle, C.byref(p), _p(cam), C.c_int(i), C.c_int(px0), C.c_int(py0), C.c_int(sw), C.c_int(sh), _p(og), _p(nr))
            d = self.directions(lv[i].D)
            tex = np.zeros((sw * sh * lv[i].D * lv[i].D, 4), np.float32)
            ht = np.zeros(len(tex), np.float32) if want_hits else None
            hp = np.zeros(len(tex), np.uint32) if want_hits else None
            L.rco_march(self.handle, C.byref(p), C.c_int(i), C.c_int(sw), C.c_int(sh), _p(og), _p(d), _p(lights),
                        C.c_int(len(lights)), C.c_uint32(flags), _p(tex),
                        _p(ht) if want_hits else None, _p(hp) if want_hits else None)
            origins.append(og); nrms.append(nr); casc.append(tex); dirs.append(d)
            if keep_raw:
                raws.append(tex.copy())
            if want_hits:
                hits.append((ht, hp))
        for i in range(p.N - 2, -1, -1):
            lx, ly, lw, lh = rects[i]
            ux, uy, uw, uh = rects[i + 1]
            L.rco_merge(C.byref(p), C.c_int(i), C.c_int(lx), C.c_int(ly), C.c_int(lw), C.c_int(lh),
                        _p(origins[i]), _p(nrms[i]), _p(casc[i]),
                        C.c_int(ux), C.c_int(uy), C.c_int(uw), C.c_int(uh), _p(origins[i + 1]), _p(casc[i + 1]))
        px0, py0, sw, sh = rects[0]
        E = np.zeros((p.tile_h, p.tile_w, 4), np.float32)
        L.rco_gather(C.byref(p), _p(cam), C.c_int(px0), C.c_int(py0), C.c_int(sw), C.c_int(sh), _p(origins[0]),
                     _p(casc[0]), _p(dirs[0]), _p(gb["depth"]), _p(gb["normal"]), _p(E))
        out = dict(gb)
        out.update(irradiance=E, cascades=casc, origins=origins, probe_normals=nrms, rects=rects, levels=lv, dirs=dirs)
        if keep_raw:
            out["raw"] = raws
        if want_hits:
            out["hits"] = hits
        return out


def num_threads() -> int:
    return int(lib().rco_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP threads of the oracle's loops (overrides OMP_NUM_THREADS, which torchrun sets to 1 for its workers)."""
    lib().rco_set_num_threads(C.c_int(int(n)))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
