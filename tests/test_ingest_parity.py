"""Scene ingest: the product's C++ loader (rc_scene_*, device-free) against
the numpy restatement of the reference (oracle/ref_ingest.py) and against the committed golden
hashes (tests/golden/ingest_golden.json, tools/make_golden.py).  Bit-exact."""
import hashlib
import json
import os

import numpy as np
import pytest

import radiancecascade_b200 as rc
from oracle import ref_ingest as ri

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ingest_golden.json")))
SCENES = sorted(rc.scenes.SCENES)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same_bits(a, b):
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize("name", SCENES)
def test_product_loader_matches_golden_and_oracle(name):
    hs = rc.ObjScene.load(rc.scenes.scene_path(name))
    info = hs.info()
    g = GOLD[name]
    assert info.num_models == len(g["models"])
    assert info.num_vertices == g["total_vertices"] and info.num_triangles == g["total_triangles"]
    assert [float(x) for x in info.bbox_min] == g["bbox_min"] and [float(x) for x in info.bbox_max] == g["bbox_max"]
    assert bool(info.light_from_obj) == (g["light"] is not None)    # no bundled scene has a material named "Light"
    osc, _ = ri.ObjScene.load(rc.scenes.scene_path(name))
    for m, gm in enumerate(g["models"]):
        v, i = hs.model_stream(m)
        assert hs.model_name(m) == gm["name"]
        assert (len(v), len(i)) == (gm["vertices"], gm["indices"])
        assert sha(v) == gm["stream_sha256"], (name, m, "vertex stream")      # src/renderer.rs:371-410
        assert sha(i) == gm["index_sha256"], (name, m, "index buffer")        # src/primitives.rs:369-376
        assert same_bits(v, osc[m].vertex_stream())
        um, eb, ke = hs.model_material(m)
        assert [float(x) for x in um] == gm["uniform_material"] and eb == gm["enable_bit"]   # src/primitives.rs:37-73
        assert [float(x) for x in ke] == gm["emission"]
        for which, key in ((0, "color_texture_sha256"), (1, "normal_texture_sha256")):
            t = hs.model_texture(m, which)
            assert (None if t is None else sha(t)) == gm[key]                 # src/texture.rs:85-147 (to_rgba8)


def test_survey_counts():
    """Counts of SURVEY Appendix B; living_room additionally carries the 5,869 `l` elements that tobj's
    triangulate turns into zero-area triangles (a,b,b) — the survey counted `f` lines only."""
    want = {"cube": (277, 428, 1), "teapot": (8334, 15704, 2), "test_room": (248, 152, 2), "sonic": (10280, 17298, 13),
            "living_room": (46167, 28656 + 5869, 41)}
    for name, (nv, nt, nm) in want.items():
        g = GOLD[name]
        assert (g["total_vertices"], g["total_triangles"], len(g["models"])) == (nv, nt, nm)


def test_winding_is_reversed_and_tbn_fallbacks():
    osc, _ = ri.ObjScene.load(rc.scenes.scene_path("sonic"))     # no texcoords usable for most models, no mtllib
    m = osc[0]
    raw = m.model.indices.reshape(-1, 3)
    assert np.array_equal(m.indices().reshape(-1, 3), raw[:, ::-1])
    assert m.materials is None and ri.enable_bit(m.material()) == 0
    um = ri.uniform_material(None)
    assert um[12] == 1.0 and not um[:12].any()                  # Material::default(): Ns -> 1.0, all K absent
    # teapot has no `vn`: the stream's normal column is the tbn normal (zip_longest Right branch)
    tp, _ = ri.ObjScene.load(rc.scenes.scene_path("teapot"))
    T, B, N = tp[0].tbn()
    assert len(tp[0].normals()) == 0 and same_bits(tp[0].vertex_stream()[:, 6:9], N)


def test_line_elements_make_degenerate_triangles():
    lr, _ = ri.ObjScene.load(rc.scenes.scene_path("living_room"))
    last = lr[-1]
    ix = last.model.indices.reshape(-1, 3)
    degenerate = (ix[:, 1] == ix[:, 2])
    assert degenerate.sum() == 5869
    assert len(last.texcoords()) == 0          # lengths disagree -> texcoords() returns [] (src/primitives.rs:356-367)


def test_missing_mtl_is_an_error(tmp_path):
    p = tmp_path / "a.obj"
    p.write_text("mtllib nothere.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    with pytest.raises(rc.RcError):
        rc.ObjScene.load(str(p))
    with pytest.raises(FileNotFoundError):
        ri.load_obj(str(p))


def test_small_obj_semantics(tmp_path):
    """Fan triangulation, negative indices, per-model re-indexing, usemtl split — product vs oracle."""
    (tmp_path / "m.mtl").write_text("newmtl A\nKd 1 0 0\nnewmtl Light\nKa 1 1 1\nKe 2 2 2\n")
    p = tmp_path / "s.obj"
    p.write_text("mtllib m.mtl\no first\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 1.5 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 1\n"
                 "usemtl A\nf 1/1 2/2 3/3 4/4 5/5\nusemtl Light\nf -5/1 -4/2 -3/3\no second\nf 1/1 3/3 4/4\n")
    hs = rc.ObjScene.load(str(p))
    osc, light = ri.ObjScene.load(str(p))
    assert hs.info().num_models == len(osc) == 3
    assert hs.info().light_from_obj == 1 and light is not None
    assert np.allclose(list(hs.info().obj_light), light)
    for m in range(3):
        v, i = hs.model_stream(m)
        assert same_bits(v, osc[m].vertex_stream()) and np.array_equal(i, osc[m].indices())
    assert len(osc[0].indices()) == 9          # pentagon -> 3 fan triangles
    assert hs.model_material(1)[2].tolist() == [2.0, 2.0, 2.0]   # Ke


def _pil_rgba(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("RGBA"), dtype=np.uint8)


def test_bundled_textures_decode_like_pil():
    """Every texture file of the bundled scenes: built-in PNG / JPEG decoders (baseline and progressive)
    against PIL (libjpeg-turbo / libpng) — bit-exact."""
    from radiancecascade_b200.renderer import decode_image_file
    n = 0
    for name in SCENES:
        root = os.path.dirname(rc.scenes.scene_path(name))
        for d, _, files in os.walk(root):
            for f in files:
                if f.lower().endswith((".jpg", ".jpeg", ".png")):
                    p = os.path.join(d, f)
                    assert np.array_equal(decode_image_file(p), _pil_rgba(p)), p
                    n += 1
    assert n >= 20


@pytest.mark.parametrize("mode,subsampling,progressive,size", [
    ("RGB", 0, False, (67, 45)), ("RGB", 1, False, (67, 45)), ("RGB", 2, False, (67, 45)), ("RGB", 2, True, (130, 71)),
    ("RGB", 1, True, (33, 90)), ("L", 0, False, (50, 50)), ("L", 0, True, (41, 23)), ("RGB", 2, False, (16, 16)), ("RGB", 2, False, (1, 1))])
def test_synthetic_jpeg_variants(tmp_path, mode, subsampling, progressive, size):
    """4:4:4 / 4:2:2 / 4:2:0, grayscale, progressive scans, odd sizes: code paths the bundled assets do not all reach."""
    from PIL import Image
    from radiancecascade_b200.renderer import decode_image_file
    rng = np.random.default_rng(size[0] * 131 + size[1])
    w, h = size
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 7) % 256], -1).astype(np.int32)
    img = np.clip(base + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)
    im = Image.fromarray(img if mode == "RGB" else img[..., 0], mode)
    p = str(tmp_path / "t.jpg")
    kw = dict(quality=87, progressive=progressive)
    if mode == "RGB":
        kw["subsampling"] = subsampling
    im.save(p, "JPEG", **kw)
    assert np.array_equal(decode_image_file(p), _pil_rgba(p))


def test_hostile_jpeg_headers_are_rejected(tmp_path):
    """Textures come from untrusted MTL paths: scans with out-of-range spectral selection / successive approximation,
    truncated DQT / SOF / DRI segments and impossible coefficient categories must fail cleanly (RcError), never write
    outside the 64-coefficient block (heap overwrite through Se > 63 in a refinement scan)."""
    from PIL import Image
    from radiancecascade_b200.renderer import decode_image_file
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    good = str(tmp_path / "good.jpg")
    Image.fromarray(img, "RGB").save(good, "JPEG", quality=80, progressive=True)
    data = bytearray(open(good, "rb").read())
    assert decode_image_file(good).shape == (48, 64, 4)

    def segments(buf):
        i = 2
        while i + 4 <= len(buf):
            assert buf[i] == 0xFF
            m = buf[i + 1]
            ln = (buf[i + 2] << 8) | buf[i + 3]
            yield i, m, ln
            if m == 0xDA:
                j = i + 2 + ln          # skip the entropy-coded data
                while j + 1 < len(buf) and not (buf[j] == 0xFF and buf[j + 1] not in (0, 0xFF) and not 0xD0 <= buf[j + 1] <= 0xD7):
                    j += 1
                i = j
            else:
                i += 2 + ln

    sos = [(i, ln) for i, m, ln in segments(data) if m == 0xDA]
    assert len(sos) >= 4
    bad = []
    for k, (i, ln) in enumerate(sos):
        ns = data[i + 4]
        ss_at = i + 5 + 2 * ns
        if data[ss_at] > 0 and data[ss_at + 2] >> 4:       # an AC refinement scan: Se -> 255
            b = bytearray(data); b[ss_at + 1] = 255; bad.append(("se255_refine", b)); break
    for k, (i, ln) in enumerate(sos):
        ns = data[i + 4]
        ss_at = i + 5 + 2 * ns
        if data[ss_at] > 0:                                 # first AC scan: Se -> 200, and Ss > Se
            b = bytearray(data); b[ss_at + 1] = 200; bad.append(("se200", b))
            b = bytearray(data); b[ss_at] = 40; b[ss_at + 1] = 10; bad.append(("ss_gt_se", b))
            b = bytearray(data); b[ss_at + 2] = 0x0F; bad.append(("al15", b))
            break
    i, ln = sos[0]
    b = bytearray(data); b[i + 5 + 2 * data[i + 4] + 1] = 5; bad.append(("dc_scan_se5", b))
    for i, m, ln in segments(data):
        if m == 0xDB:
            b = bytearray(data); b[i + 2:i + 4] = (10).to_bytes(2, "big"); del b[i + 12:i + 2 + ln]; bad.append(("short_dqt", b)); break
    for i, m, ln in segments(data):
        if m == 0xC2:
            b = bytearray(data); b[i + 2:i + 4] = (4).to_bytes(2, "big"); del b[i + 6:i + 2 + ln]; bad.append(("short_sof", b))
            b = bytearray(data); b[i + 5:i + 9] = bytes([0xFF, 0xFF, 0xFF, 0xFF]); bad.append(("huge_sof", b)); break
    assert len(bad) >= 7
    for name, b in bad:
        q = str(tmp_path / f"{name}.jpg")
        open(q, "wb").write(bytes(b))
        with pytest.raises(rc.RcError):
            decode_image_file(q)
    # truncation anywhere must never crash: either a clean error or a (partially grey) image
    for cut in range(20, len(data), max(1, len(data) // 97)):
        q = str(tmp_path / "cut.jpg")
        open(q, "wb").write(bytes(data[:cut]))
        try:
            decode_image_file(q)
        except rc.RcError:
            pass
    # random byte corruption of the entropy-coded segments / tables: same requirement
    for t in range(60):
        b = bytearray(data)
        for _ in range(4):
            b[int(rng.integers(2, len(b)))] = int(rng.integers(0, 256))
        q = str(tmp_path / "fuzz.jpg")
        open(q, "wb").write(bytes(b))
        try:
            decode_image_file(q)
        except rc.RcError:
            pass


def test_png_variants(tmp_path):
    from PIL import Image
    from radiancecascade_b200.renderer import decode_image_file
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    for mode, arr in (("RGBA", a), ("RGB", a[..., :3]), ("L", a[..., 0]), ("LA", a[..., :2])):
        p = str(tmp_path / f"{mode}.png")
        Image.fromarray(arr, mode).save(p)
        assert np.array_equal(decode_image_file(p), _pil_rgba(p)), mode
    pal = Image.fromarray(a[..., 0] % 7, "P")
    pal.putpalette([i * 9 % 256 for i in range(768)])
    p = str(tmp_path / "P.png")
    pal.save(p, bits=4)
    assert np.array_equal(decode_image_file(p), _pil_rgba(p))
    with pytest.raises(rc.RcError):
        decode_image_file(str(tmp_path / "missing.png"))


def test_empty_and_degenerate_scenes(tmp_path):
    """Edge cases of the loader: no faces at all, only a line element, a face with a missing vertex."""
    p = tmp_path / "nofaces.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\n")
    hs = rc.ObjScene.load(str(p))             # tobj emits one (empty) model for the trailing object
    osc, _ = ri.ObjScene.load(str(p))
    assert hs.info().num_models == len(osc) == 1 and hs.info().num_triangles == 0 and hs.info().num_vertices == 0
    p = tmp_path / "line.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nl 1 2\nf 1 2 3\n")
    hs = rc.ObjScene.load(str(p))
    osc, _ = ri.ObjScene.load(str(p))
    v, i = hs.model_stream(0)
    assert np.array_equal(i, osc[0].indices()) and len(i) == 6       # (a,b,b) reversed -> (b,b,a), then the triangle
    assert same_bits(v, osc[0].vertex_stream())
    p = tmp_path / "oob.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nf 1 2 7\n")
    with pytest.raises(rc.RcError):
        rc.ObjScene.load(str(p))              # tobj: FaceVertexOutOfBounds -> the reference's unwrap() panics
