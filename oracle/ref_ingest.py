"""CPU restatement of the reference's scene ingest, camera math and direct lighting.

TEST INFRASTRUCTURE ONLY.  Nothing under radiancecascade_b200/ may import this
module; it is the checker for tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg.

PARITY UNPINNED: the reference (jw910731/RadianceCascade) ships no tests, golden
vectors or fixtures and cannot be built here (no Rust, no Vulkan;
SURVEY.md Appendix D).  Every function below restates reference source by
file:line; third-party arithmetic that is not vendored in the reference (tobj
4.0.2, glam 0.29.2, image 0.25.5 — Cargo.lock) is restated from the crates'
published behaviour and anchored on the reference's call sites.

All arithmetic is IEEE binary32 through numpy float32 scalars/arrays (numpy never
fuses a*b+c), mirroring Rust's non-contracting f32 semantics.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

F = np.float32
MISSING = -1


# --------------------------------------------------------------------------
# tobj 4.0.2 restatement (un-vendored; call site src/primitives.rs:104-113 with
# LoadOptions { triangulate: true, single_index: true, ..Default })
# --------------------------------------------------------------------------
@dataclass
class TobjMaterial:
    name: str = ""
    ambient: Optional[np.ndarray] = None      # Ka
    diffuse: Optional[np.ndarray] = None      # Kd
    specular: Optional[np.ndarray] = None     # Ks
    shininess: Optional[np.float32] = None    # Ns
    diffuse_texture: Optional[str] = None     # map_Kd
    normal_texture: Optional[str] = None      # map_Bump | map_bump | bump
    unknown_param: Dict[str, str] = field(default_factory=dict)  # Ke lands here


@dataclass
class TobjModel:
    name: str
    positions: np.ndarray    # float32 [3*nv]
    normals: np.ndarray      # float32 [3*k]  (compacted: only vertices that carry vn)
    texcoords: np.ndarray    # float32 [2*k]
    vertex_color: np.ndarray # float32 [3*k]
    indices: np.ndarray      # uint32  [3*nt]
    material_id: Optional[int]


def _parse_f32(tok: str) -> np.float32:
    # Rust str::parse::<f32>() is correctly rounded from the decimal string.
    # float() -> float32 double-rounds; for the <= 9 significant digit literals
    # in the bundled scenes the two agree (checked against the C++ strtof loader
    # in tests/test_ingest_parity.py).
    return np.float32(float(tok))


def load_mtl(path: str) -> Tuple[List[TobjMaterial], Dict[str, int]]:
    """tobj::load_mtl_buf: one material per `newmtl`; unknown keys kept as strings."""
    mats: List[TobjMaterial] = []
    name_map: Dict[str, int] = {}
    cur: Optional[TobjMaterial] = None
    with open(path, "r", errors="replace") as fh:
        for raw in fh:
            line = raw.strip()
            words = line.split()
            if not words or words[0].startswith("#"):
                continue
            key = words[0]
            if key == "newmtl":
                if cur is not None:
                    name_map[cur.name] = len(mats)
                    mats.append(cur)
                cur = TobjMaterial(name=line[6:].strip())
                continue
            if cur is None:
                continue
            rest = line[len(key):].strip()
            if key == "Ka":
                cur.ambient = np.array([_parse_f32(w) for w in words[1:4]], dtype=F)
            elif key == "Kd":
                cur.diffuse = np.array([_parse_f32(w) for w in words[1:4]], dtype=F)
            elif key == "Ks":
                cur.specular = np.array([_parse_f32(w) for w in words[1:4]], dtype=F)
            elif key == "Ns":
                cur.shininess = _parse_f32(words[1])
            elif key == "map_Kd":
                cur.diffuse_texture = rest
            elif key in ("map_Bump", "map_bump", "bump"):
                cur.normal_texture = rest
            elif key in ("Ni", "d", "illum", "map_Ka", "map_Ks", "map_Ns", "map_d"):
                pass  # parsed by tobj, unused by the reference
            else:
                cur.unknown_param[key] = rest
    if cur is not None:
        name_map[cur.name] = len(mats)
        mats.append(cur)
    return mats, name_map


def _parse_vertex_indices(word: str, npos: int, ntex: int, nnorm: int) -> Tuple[int, int, int]:
    """tobj VertexIndices::parse: v[/vt][/vn]; negative = relative to the arrays so far."""
    out = [MISSING, MISSING, MISSING]
    sizes = (npos, ntex, nnorm)
    for i, part in enumerate(word.split("/")[:3]):
        if part == "":
            continue
        x = int(part)
        out[i] = sizes[i] + x if x < 0 else x - 1
    return out[0], out[1], out[2]


def _export_faces(pos, vcol, tex, nrm, faces, mat_id, name) -> TobjModel:
    """tobj export_faces (single_index path): one output vertex per distinct
    (v,vt,vn) tuple in order of first appearance within the model; polygons
    fan-triangulated; points/lines blown up to zero-area triangles."""
    index_map: Dict[Tuple[int, int, int], int] = {}
    positions: List[float] = []
    normals: List[float] = []
    texcoords: List[float] = []
    colors: List[float] = []
    indices: List[int] = []

    def add(vi):
        got = index_map.get(vi)
        if got is not None:
            indices.append(got)
            return
        v, vt, vn = vi
        positions.extend(pos[3 * v:3 * v + 3])
        if len(tex) and vt != MISSING:
            texcoords.extend(tex[2 * vt:2 * vt + 2])
        if len(nrm) and vn != MISSING:
            normals.extend(nrm[3 * vn:3 * vn + 3])
        if len(vcol):
            colors.extend(vcol[3 * v:3 * v + 3])
        nxt = len(index_map)
        indices.append(nxt)
        index_map[vi] = nxt

    for f in faces:
        n = len(f)
        if n == 1:
            add(f[0]); add(f[0]); add(f[0])
        elif n == 2:
            add(f[0]); add(f[1]); add(f[1])
        elif n == 3:
            add(f[0]); add(f[1]); add(f[2])
        elif n == 4:
            add(f[0]); add(f[1]); add(f[2])
            add(f[0]); add(f[2]); add(f[3])
        else:
            a = f[0]
            b = f[1]
            for c in f[2:]:
                add(a); add(b); add(c)
                b = c
    return TobjModel(
        name=name,
        positions=np.array(positions, dtype=F),
        normals=np.array(normals, dtype=F),
        texcoords=np.array(texcoords, dtype=F),
        vertex_color=np.array(colors, dtype=F),
        indices=np.array(indices, dtype=np.uint32),
        material_id=mat_id,
    )


def load_obj(path: str) -> Tuple[List[TobjModel], List[TobjMaterial]]:
    """tobj::load_obj(path, {triangulate, single_index}).  A new model starts at
    each `o`/`g` and whenever `usemtl` changes after faces were emitted; a
    missing .mtl file is an error (materials? at src/primitives.rs:132)."""
    pos: List[np.float32] = []
    vcol: List[np.float32] = []
    tex: List[np.float32] = []
    nrm: List[np.float32] = []
    faces: List[List[Tuple[int, int, int]]] = []
    models: List[TobjModel] = []
    materials: List[TobjMaterial] = []
    mat_map: Dict[str, int] = {}
    name = "unnamed_object"
    mat_id: Optional[int] = None
    base = os.path.dirname(path)
    with open(path, "r", errors="replace") as fh:
        for raw in fh:
            line = raw.strip()
            words = line.split()
            if not words:
                continue
            key = words[0]
            if key == "v":
                pos.extend(_parse_f32(w) for w in words[1:4])
                if len(words) >= 7:
                    vcol.extend(_parse_f32(w) for w in words[4:7])
            elif key == "vt":
                tex.extend(_parse_f32(w) for w in words[1:3])
            elif key == "vn":
                nrm.extend(_parse_f32(w) for w in words[1:4])
            elif key in ("f", "l"):
                faces.append([_parse_vertex_indices(w, len(pos) // 3, len(tex) // 2, len(nrm) // 3)
                              for w in words[1:]])
            elif key in ("o", "g"):
                if faces:
                    models.append(_export_faces(pos, vcol, tex, nrm, faces, mat_id, name))
                    faces = []
                name = line[1:].strip() or "unnamed_object"
            elif key == "mtllib":
                mtl_name = line[len("mtllib"):].strip()
                mats, nm = load_mtl(os.path.join(base, mtl_name))  # raises if missing
                off = len(materials)
                materials.extend(mats)
                for k, v in nm.items():
                    mat_map[k] = v + off
            elif key == "usemtl":
                mat_name = line[len("usemtl"):].strip()
                new_mat = mat_map.get(mat_name)
                if mat_id != new_mat and faces:
                    models.append(_export_faces(pos, vcol, tex, nrm, faces, mat_id, name))
                    faces = []
                mat_id = new_mat
    models.append(_export_faces(pos, vcol, tex, nrm, faces, mat_id, name))
    return models, materials


# --------------------------------------------------------------------------
# ObjScene (src/primitives.rs:115-416)
# --------------------------------------------------------------------------
@dataclass
class Material:
    """src/primitives.rs:75-83 (+ Ke, which the reference ignores; GI emission)."""
    ambient: Optional[np.ndarray]
    diffuse: Optional[np.ndarray]
    specular: Optional[np.ndarray]
    shininess: Optional[np.float32]
    color_texture: Optional[np.ndarray]   # uint8 [h][w][4]
    normal_texture: Optional[np.ndarray]  # uint8 [h][w][4]
    emission: np.ndarray


class ObjScene:
    def __init__(self, model: TobjModel, obj_dir: str, material: Optional[TobjMaterial]):
        self.model = model
        self.obj_dir = obj_dir
        self.materials = material

    # src/primitives.rs:122-175
    @staticmethod
    def load(path: str, light_predicate: Callable[[TobjMaterial], bool] = lambda m: m.name == "Light"):
        models, materials = load_obj(path)
        light = None
        for md in models:
            if md.material_id is None or md.material_id >= len(materials):
                continue
            if not light_predicate(materials[md.material_id]):
                continue
            p = md.positions.reshape(-1, 3)
            s = np.zeros(3, dtype=F)
            for row in p:            # Iterator::sum::<Vec3>() — sequential f32 adds
                s = (s + row).astype(F)
            light = (s / F(len(p))).astype(F)
            break                    # .take(1)
        obj_dir = os.path.dirname(path)
        scenes = [ObjScene(m, obj_dir,
                           materials[m.material_id] if (m.material_id is not None and m.material_id < len(materials)) else None)
                  for m in models]
        return scenes, light

    def name(self) -> str:
        return self.model.name

    def vertices(self) -> np.ndarray:        # :218-225
        return self.model.positions.reshape(-1, 3)

    def vertex_colors(self) -> np.ndarray:   # :227-234
        return self.model.vertex_color.reshape(-1, 3)

    def normals(self) -> np.ndarray:         # :236-243
        return self.model.normals.reshape(-1, 3)

    def texcoords(self) -> np.ndarray:       # :356-367
        if len(self.model.positions) // 3 == len(self.model.texcoords) // 2:
            return self.model.texcoords.reshape(-1, 2)
        return np.zeros((0, 2), dtype=F)

    def indices(self) -> np.ndarray:         # :369-376  (a,b,c) -> (c,b,a)
        return self.model.indices.reshape(-1, 3)[:, ::-1].reshape(-1).copy()

    def vertex_count(self) -> int:           # :378-380 — the INDEX count
        return len(self.model.indices)

    def tbn(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """src/primitives.rs:245-354 with glam 0.29.2 arithmetic."""
        pos = self.vertices()
        nv = len(pos)
        uv = self.texcoords()
        if len(uv) != nv:
            uv = np.zeros((nv, 2), dtype=F)
        idx = self.indices().reshape(-1, 3)
        p0, p1, p2 = pos[idx[:, 0]], pos[idx[:, 1]], pos[idx[:, 2]]
        u0, u1, u2 = uv[idx[:, 0]], uv[idx[:, 1]], uv[idx[:, 2]]
        with np.errstate(all="ignore"):
            dp1 = (p1 - p0).astype(F)
            dp2 = (p2 - p0).astype(F)
            s = F(2048.0)                       # 2.0f32.powi(11)
            duv1 = ((u1 - u0) * s).astype(F)
            duv2 = ((u2 - u0) * s).astype(F)
            a, b = duv1[:, 0], duv1[:, 1]       # mat2 x_axis
            c, d = duv2[:, 0], duv2[:, 1]       # mat2 y_axis
            det = (a * d - b * c).astype(F)     # (a*d) - (b*c), two rounded products
            inv = (F(1.0) / det).astype(F)
            r00 = (d * inv).astype(F)           # r.col(0).x
            r01 = (b * (-inv)).astype(F)        # r.col(0).y
            r10 = (c * (-inv)).astype(F)        # r.col(1).x
            r11 = (a * inv).astype(F)           # r.col(1).y
            tangent = (r00[:, None] * dp1 - r01[:, None] * dp2).astype(F)
            bitangent = ((-r10)[:, None] * dp1 + r11[:, None] * dp2).astype(F)
            normal = _normalize(_cross(bitangent, tangent))
            ok = ~(np.isnan(tangent).any(1) | np.isnan(bitangent).any(1) | np.isnan(normal).any(1))
        T = np.zeros((nv, 3), dtype=F)
        B = np.zeros((nv, 3), dtype=F)
        N = np.zeros((nv, 3), dtype=F)
        cnt = np.zeros(nv, dtype=np.int64)
        with np.errstate(all="ignore"):
            # sequential accumulation in triangle order, vertices c[0], c[1], c[2]
            tri = np.nonzero(ok)[0]
            flat = idx[tri].reshape(-1)
            np.add.at(T, flat, np.repeat(tangent[tri], 3, axis=0))
            np.add.at(B, flat, np.repeat(bitangent[tri], 3, axis=0))
            np.add.at(N, flat, np.repeat(normal[tri], 3, axis=0))
            np.add.at(cnt, flat, 1)
            has = cnt > 0
            cf = np.maximum(cnt, 1).astype(F)[:, None]
            Tn = _normalize((T / cf).astype(F))
            Bn = _normalize((B / cf).astype(F))
            Nn = _normalize((N / cf).astype(F))
        Tn[~has] = (1, 0, 0)
        Bn[~has] = (0, 1, 0)
        Nn[~has] = (0, 0, 1)
        return Tn, Bn, Nn

    def vertex_stream(self) -> np.ndarray:
        """src/renderer.rs:371-410: [pos3|colour3|normal3|tangent3|bitangent3|uv2] f32,
        with the iterator-chain fallbacks (colour -> 1, normal -> OBJ then tbn then Z,
        tangent -> X, bitangent -> Y, uv -> 0)."""
        pos = self.vertices()
        nv = len(pos)
        T, B, N = self.tbn()
        out = np.zeros((nv, 17), dtype=F)
        out[:, 0:3] = pos
        col = self.vertex_colors()
        out[:, 3:6] = 1.0
        out[:min(len(col), nv), 3:6] = col[:nv]
        # zip_longest(normals, tbn normal): Both -> OBJ normal, Left -> OBJ, Right -> tbn
        nr = self.normals()
        chain = np.concatenate([nr, N[len(nr):]]) if len(nr) < len(N) else nr
        out[:, 6:9] = (0, 0, 1)
        k = min(len(chain), nv)
        out[:k, 6:9] = chain[:k]
        out[:, 9:12] = T
        out[:, 12:15] = B
        uv = self.texcoords()
        out[:min(len(uv), nv), 15:17] = uv[:nv]
        return out

    def material(self, decode_textures: bool = True) -> Optional[Material]:
        """src/primitives.rs:386-415.  A texture that cannot be opened/decoded is None."""
        e = self.materials
        if e is None:
            return None
        ke = np.zeros(3, dtype=F)
        if "Ke" in e.unknown_param:
            w = e.unknown_param["Ke"].split()
            if len(w) >= 3:
                ke = np.array([_parse_f32(x) for x in w[:3]], dtype=F)
        return Material(
            ambient=e.ambient, diffuse=e.diffuse, specular=e.specular, shininess=e.shininess,
            color_texture=_load_rgba8(os.path.join(self.obj_dir, e.diffuse_texture)) if (e.diffuse_texture and decode_textures) else None,
            normal_texture=_load_rgba8(os.path.join(self.obj_dir, e.normal_texture)) if (e.normal_texture and decode_textures) else None,
            emission=ke)


def _load_rgba8(path: str) -> Optional[np.ndarray]:
    """image::ImageReader::open(p).decode().to_rgba8() (src/primitives.rs:391-404,
    src/texture.rs:92).  PIL stands in for image 0.25.5; JPEG IDCT may differ by
    +-1 LSB from zune-jpeg (SURVEY §7 hard parts)."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            return np.asarray(im.convert("RGBA"), dtype=np.uint8).copy()
    except Exception:
        return None


def uniform_material(mat: Optional[Material]) -> np.ndarray:
    """UniformMaterial (src/primitives.rs:37-73): 3 x vec4 (.w = present ? 1 : 0),
    shininess (default 1.0), 3 x u32 padding -> 16 float32 (64 bytes)."""
    out = np.zeros(16, dtype=F)
    if mat is None:
        out[12] = 1.0          # Material::default(): everything None, Ns -> 1.0
        return out
    for i, v in enumerate((mat.ambient, mat.diffuse, mat.specular)):
        if v is not None:
            out[4 * i:4 * i + 3] = v
            out[4 * i + 3] = 1.0
    out[12] = mat.shininess if mat.shininess is not None else F(1.0)
    return out


def enable_bit(mat: Optional[Material]) -> int:
    """src/renderer.rs:422-423,455-456: colour | normal << 1."""
    if mat is None:
        return 0
    return int(mat.color_texture is not None) | (int(mat.normal_texture is not None) << 1)


# --------------------------------------------------------------------------
# glam 0.29.2 helpers (scalar-equivalent arithmetic)
# --------------------------------------------------------------------------
def _cross(a, b):
    a = np.asarray(a, dtype=F); b = np.asarray(b, dtype=F)
    x = (a[..., 1] * b[..., 2]).astype(F) - (b[..., 1] * a[..., 2]).astype(F)
    y = (a[..., 2] * b[..., 0]).astype(F) - (b[..., 2] * a[..., 0]).astype(F)
    z = (a[..., 0] * b[..., 1]).astype(F) - (b[..., 0] * a[..., 1]).astype(F)
    return np.stack([x, y, z], axis=-1).astype(F)


def _dot(a, b):
    a = np.asarray(a, dtype=F); b = np.asarray(b, dtype=F)
    return (((a[..., 0] * b[..., 0]).astype(F) + (a[..., 1] * b[..., 1]).astype(F)).astype(F)
            + (a[..., 2] * b[..., 2]).astype(F)).astype(F)


def _normalize(v):
    v = np.asarray(v, dtype=F)
    with np.errstate(all="ignore"):
        rl = (F(1.0) / np.sqrt(_dot(v, v)).astype(F)).astype(F)   # length().recip()
        return (v * rl[..., None]).astype(F)


# --------------------------------------------------------------------------
# Camera / Projection / UniformCamera (src/camera.rs:9-80)
# --------------------------------------------------------------------------
def look_to_rh(eye, direction, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """glam Mat4::look_to_rh; returns a 4x4 array indexed [row][col]."""
    eye = np.asarray(eye, dtype=F)
    f = _normalize(np.asarray(direction, dtype=F))
    s = _normalize(_cross(f, np.asarray(up, dtype=F)))
    u = _cross(s, f)
    m = np.zeros((4, 4), dtype=F)
    m[0, :3] = s
    m[1, :3] = u
    m[2, :3] = -f
    m[0, 3] = -_dot(eye, s)
    m[1, 3] = -_dot(eye, u)
    m[2, 3] = _dot(eye, f)
    m[3, 3] = 1.0
    return m


def camera_matrix(position, yaw: float, pitch: float) -> np.ndarray:
    """Camera::calc_matrix (src/camera.rs:43-52).  yaw/pitch are used as radians
    exactly as stored (the reference passes degrees — SURVEY Appendix B quirk 1)."""
    sp, cp = F(math.sin(F(pitch))), F(math.cos(F(pitch)))
    sy, cy = F(math.sin(F(yaw))), F(math.cos(F(yaw)))
    d = np.array([cp * cy, sp, cp * sy], dtype=F)
    return look_to_rh(position, _normalize(d))


def projection_matrix(fovy: float, aspect: float, znear: float, zfar: float) -> np.ndarray:
    """glam Mat4::perspective_rh (depth 0..1), via Projection::calc_matrix (src/camera.rs:77-79)."""
    half = F(F(0.5) * F(fovy))
    sin_f, cos_f = F(math.sin(half)), F(math.cos(half))
    h = F(cos_f / sin_f)
    w = F(h / F(aspect))
    r = F(F(zfar) / F(F(znear) - F(zfar)))
    m = np.zeros((4, 4), dtype=F)
    m[0, 0] = w
    m[1, 1] = h
    m[2, 2] = r
    m[3, 2] = -1.0
    m[2, 3] = F(r * F(znear))
    return m


def mat4_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """glam Mat4 * Mat4: column j = ((A0*b0j + A1*b1j) + A2*b2j) + A3*b3j, no fma."""
    out = np.zeros((4, 4), dtype=F)
    for j in range(4):
        acc = (a[:, 0] * b[0, j]).astype(F)
        for k in range(1, 4):
            acc = (acc + (a[:, k] * b[k, j]).astype(F)).astype(F)
        out[:, j] = acc
    return out


def uniform_camera(position, yaw, pitch, fovy, aspect, znear, zfar) -> np.ndarray:
    """UniformCamera::from_camera_project (src/camera.rs:17-22): 16 floats
    column-major proj*view then (eye,1) -> float32[20] (80 bytes)."""
    m = mat4_mul(projection_matrix(fovy, aspect, znear, zfar), camera_matrix(position, yaw, pitch))
    return np.concatenate([m.T.reshape(-1), np.asarray(position, dtype=F), [F(1.0)]]).astype(F)


def uniform_camera_look_at(position, target, fovy, aspect, znear, zfar) -> np.ndarray:
    d = (np.asarray(target, dtype=F) - np.asarray(position, dtype=F)).astype(F)
    m = mat4_mul(projection_matrix(fovy, aspect, znear, zfar), look_to_rh(position, _normalize(d)))
    return np.concatenate([m.T.reshape(-1), np.asarray(position, dtype=F), [F(1.0)]]).astype(F)


# --------------------------------------------------------------------------
# Texture sampling (src/texture.rs:85-147) and fs_main (src/shader.wgsl:76-100)
# --------------------------------------------------------------------------
def srgb_table() -> np.ndarray:
    c = np.arange(256, dtype=np.float64) / 255.0
    lin = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
    return lin.astype(F)


_SRGB = srgb_table()


def _mirror_index(i: np.ndarray, size: int) -> np.ndarray:
    """Vulkan MirrorRepeat: (size-1) - mirror((i mod 2size) - size), mirror(n) = n>=0 ? n : -(1+n)."""
    m = np.mod(i, 2 * size) - size
    m = np.where(m >= 0, m, -(1 + m))
    return (size - 1) - m


def sample_nearest(tex: Optional[np.ndarray], uv: np.ndarray, srgb: bool) -> np.ndarray:
    """Nearest texel, MirrorRepeat (the reference's min filter).  A missing
    texture is the 1x1 (0,0,0,0) texel of Texture::empty (src/texture.rs:12-64)."""
    uv = np.asarray(uv, dtype=F)
    if tex is None:
        return np.zeros(uv.shape[:-1] + (3,), dtype=F)
    h, w = tex.shape[:2]
    with np.errstate(all="ignore"):
        ix = np.floor((uv[..., 0] * F(w)).astype(F)).astype(np.int64)
        iy = np.floor((uv[..., 1] * F(h)).astype(F)).astype(np.int64)
    ix = _mirror_index(ix, w)
    iy = _mirror_index(iy, h)
    t = tex[iy, ix, :3]
    return _SRGB[t] if srgb else (t.astype(F) / F(255.0)).astype(F)


def fs_main(P, color, Nv, Tv, Bv, uv, umat, ebit, eye, lights, color_tex=None, normal_tex=None) -> np.ndarray:
    """src/shader.wgsl:76-100 restated, vectorised over points.
    umat: float32[16] UniformMaterial; lights: [n][3] (the reference has one;
    diffuse and specular are summed over lights).  `eye` may be per-point."""
    P = np.asarray(P, dtype=F); Nv = np.asarray(Nv, dtype=F)
    Tv = np.asarray(Tv, dtype=F); Bv = np.asarray(Bv, dtype=F)
    uv = np.asarray(uv, dtype=F); color = np.asarray(color, dtype=F)
    eye = np.asarray(eye, dtype=F)
    b0 = ebit & 1
    b1 = (ebit >> 1) & 1
    Ka, Kd, Ks, Ns = umat[0:4], umat[4:8], umat[8:12], umat[12]
    with np.errstate(all="ignore"):
        uvp = np.stack([uv[..., 0], (F(1.0) - uv[..., 1]).astype(F)], axis=-1)          # :78
        albedo = sample_nearest(color_tex, uvp, srgb=True) if b0 else color            # :80
        L = np.broadcast_to((Ka[:3] * F(0.05) * Ka[3]).astype(F), P.shape).copy()       # :82-83
        if b1:
            c = (sample_nearest(normal_tex, uvp, srgb=False) * F(2.0) - F(1.0)).astype(F)  # :85
            raw = _normalize((c[..., 0:1] * _normalize(Tv) + c[..., 1:2] * _normalize(Bv)).astype(F)
                             + (c[..., 2:3] * Nv).astype(F))                            # :86 (Nv not normalised)
        else:
            raw = _normalize(Nv)
        V = _normalize((eye - P).astype(F))                                             # :87
        ndv = _dot(V, raw)                                                              # :88
        N = np.where(ndv[..., None] < 0, -raw, raw)                                     # :89
        for lp in np.asarray(lights, dtype=F).reshape(-1, 3):
            Ld = _normalize((lp - P).astype(F))                                         # :91
            ndl = np.fmax(_dot(Ld, N), F(0.0))      # :92  max() with C fmaxf semantics: NaN -> 0 (rc_spec.h S7)
            L = (L + ((Kd[:3] * F(0.7)).astype(F) * ndl[..., None] * Kd[3]).astype(F)).astype(F)  # :93
            Hd = _normalize((V + Ld).astype(F))                                         # :95
            s = np.power(np.fmax(_dot(N, Hd), F(0.0)), F(Ns)).astype(F)                 # :96
            gate = (ndv > F(1e-6)).astype(F)
            L = (L + (Ks[:3] * s[..., None] * Ks[3] * gate[..., None]).astype(F)).astype(F)  # :97
        pred = ((Ka[:3] - F(1e-5)) + (Kd[:3] - F(1e-5)) + (Ks[:3] - F(1e-5))).astype(F)  # :99
        unlit = F(1.0) if F(pred[0] + pred[1] + pred[2]) <= 0 else F(0.0)
        return ((L + unlit) * albedo).astype(F)                                         # :100


def srgb_encode_u8(lin: np.ndarray) -> np.ndarray:
    """Bgra8UnormSrgb store: sRGB_encode(clamp(x,0,1)) rounded to 8 bits."""
    x = np.clip(np.asarray(lin, dtype=np.float64), 0.0, 1.0)
    e = np.where(x <= 0.0031308, x * 12.92, 1.055 * np.power(x, 1.0 / 2.4) - 0.055)
    return np.floor(e * 255.0 + 0.5).astype(np.uint8)
