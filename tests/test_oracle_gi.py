"""The GI oracle (oracle/rc_oracle.c) against its golden fixture and against properties of the
specification (include/rc_spec.h) that do not depend on any implementation."""
import os
import sys

import numpy as np
import pytest

import radiancecascade_b200 as rc
from oracle import gi_oracle as go
from oracle import ref_ingest as ri

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "gi_cube64.npz"))


@pytest.fixture(scope="module")
def cube():
    return go.OracleScene(rc.scenes.scene_path("cube"))


def test_golden_frame(cube):
    p = cube.params(64, 64, store_half=True)
    assert np.array_equal(np.array([p.L0, p.t_far, p.offset], np.float32), GOLD["intervals"])
    lv = cube.levels(p)
    assert np.array_equal(np.array([[l.P, l.D, l.gw, l.gh] for l in lv], np.int32), GOLD["levels"])
    assert np.array_equal(np.array([[l.t0, l.t1] for l in lv], np.float32), GOLD["t_ranges"])
    assert np.array_equal(cube.directions(4), GOLD["dirs0"]) and np.array_equal(cube.directions(8), GOLD["dirs1"])
    out = cube.render(p, GOLD["cam"], GOLD["lights"])
    assert np.array_equal(out["prim"], GOLD["prim"]) and np.array_equal(out["depth"], GOLD["depth"])
    assert np.array_equal(out["normal"], GOLD["normal"])
    assert np.array_equal(out["irradiance"].astype(np.float16), GOLD["irradiance"])
    assert np.array_equal(out["cascades"][0].astype(np.float16), GOLD["cascade0"])


def test_intervals_are_contiguous_and_scale_by_four(cube):
    p = cube.params(256, 128)
    lv = cube.levels(p)
    assert lv[0].t0 == 0.0 and lv[-1].t1 == p.t_far
    for a, b in zip(lv[:-1], lv[1:]):
        assert a.t1 == b.t0                          # no gap, no overlap (S2)
    lens = [l.t1 - l.t0 for l in lv[:-1]]
    assert np.allclose(np.array(lens[1:]) / np.array(lens[:-1]), 4.0, rtol=1e-5)
    for i, l in enumerate(lv):
        assert (l.P, l.D) == (4 << i, 4 << i) and (l.gw, l.gh) == (-(-256 // l.P), -(-128 // l.P))


@pytest.mark.parametrize("D", [4, 8, 16, 32])
def test_directions_equal_area_and_nested(D):
    d = go.OracleScene.directions(D).astype(np.float64).reshape(D, D, 3)
    assert np.allclose(np.linalg.norm(d, axis=-1), 1.0, atol=1e-6)
    assert np.abs(d.sum((0, 1))).max() < 1e-4                      # symmetric over the sphere
    # equal area: the midpoint rule with the uniform weight 4*pi/D^2 integrates the cosine lobe about +z to
    # pi*(1 - 4/D^2) — a whole ring of texel centres sits on the equator.  At D0 = 4 that is 0.75*pi, which is
    # why S9 normalises the cosine weights instead of using 4*pi/D0^2.
    cosw = np.maximum(d[..., 2], 0).sum() * 4 * np.pi / (D * D)
    assert abs(cosw - np.pi * (1 - 4.0 / (D * D))) < 1e-5
    c = go.OracleScene.directions(2 * D).astype(np.float64).reshape(2 * D, 2 * D, 3)
    kids = c.reshape(D, 2, D, 2, 3).mean((1, 3))
    kids /= np.linalg.norm(kids, axis=-1, keepdims=True)
    assert (kids * d).sum(-1).min() > np.cos(1.5 * np.sqrt(4 * np.pi / (D * D)))   # children straddle their parent


def test_bvh_equals_brute_force_including_ties(cube):
    rng = np.random.default_rng(5)
    n = 30000
    o = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # rays aimed exactly at mesh vertices hit several triangles at the same t: the id tie-break must hold
    verts = cube.verts[:, :3]
    tgt = verts[rng.integers(0, len(verts), 4000)]
    d[:4000] = tgt - o[:4000]
    d[:4000] /= np.linalg.norm(d[:4000], axis=1, keepdims=True)
    rays = np.concatenate([o, np.zeros((n, 1), np.float32), d, np.full((n, 1), 3e38, np.float32)], 1).astype(np.float32)
    a, b = cube.trace(rays), cube.trace(rays, brute=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    hit = a[:, 0] >= 0
    assert 0.2 < hit.mean() < 0.9 and np.all(a[hit, 1] >= 0) and np.all(a[hit, 1] + a[hit, 2] <= 1.0 + 1e-6)


def test_interval_restriction(cube):
    """A ray split into [0,t) and [t,inf) finds its hit in exactly one of the two pieces (S5: tmin <= t < tmax)."""
    rng = np.random.default_rng(6)
    n = 5000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    d = -o / np.linalg.norm(o, axis=1, keepdims=True)
    full = cube.trace(np.concatenate([o, np.zeros((n, 1)), d, np.full((n, 1), 3e38)], 1).astype(np.float32))
    hit = full[:, 0] >= 0
    assert hit.mean() > 0.5
    t = full[:, 0].copy()
    lo = cube.trace(np.concatenate([o, np.zeros((n, 1)), d, t[:, None]], 1).astype(np.float32))      # [0, t): excludes the hit
    hi = cube.trace(np.concatenate([o, t[:, None], d, np.full((n, 1), 3e38)], 1).astype(np.float32))  # [t, inf): includes it
    assert np.all(lo[hit, 0] < 0)
    assert np.array_equal(hi[hit].view(np.uint32), full[hit].view(np.uint32))


def test_merge_and_gather_are_linear_in_radiance(cube):
    """S8/S9 are linear maps of the raw radiance: doubling the light term doubles E (ambient-free material not needed:
    compare E(lights) - E(no light) scaling through two light positions superposed)."""
    W = H = 48
    p = cube.params(W, H, store_half=False)
    pos, tgt, zn, zf = rc.scenes.orbit_camera(cube.bbox_min, cube.bbox_max, 9)
    cam = ri.uniform_camera_look_at(pos, tgt, np.float32(np.radians(np.float32(45.0))), np.float32(1), zn, zf)
    la = np.array([[0.0, 3.0, 0.5, 1.0]], np.float32)
    lb = np.array([[2.5, -1.0, 2.0, 1.0]], np.float32)
    Ea = cube.render(p, cam, la)["irradiance"][..., :3]
    Eb = cube.render(p, cam, lb)["irradiance"][..., :3]
    Eab = cube.render(p, cam, np.concatenate([la, lb]))["irradiance"][..., :3]
    E0 = cube.render(p, cam, np.zeros((0, 4), np.float32))["irradiance"][..., :3]
    # ambient is counted once; diffuse + specular add over lights (S7)
    assert np.abs((Ea - E0) + (Eb - E0) - (Eab - E0)).max() <= 2e-5 * max(1.0, Eab.max())


def test_uniform_environment_gathers_pi(cube):
    """Gather of a constant radiance field: E = L * sum_d max(n.w_d, 0) * 4*pi/D0^2 -> ~pi*L (S9 quadrature)."""
    W = H = 32
    p = cube.params(W, H, store_half=False)
    cam = GOLD["cam"]
    gb = cube.gbuffer(p, cam, GOLD["lights"])
    rect = go.level_rects(W, H, 4, 6, (0, 0, W, H))[0]
    og = np.zeros((rect[2] * rect[3], 4), np.float32); nr = np.zeros_like(og)
    import ctypes as C
    go.lib().rco_probes(cube.handle, C.byref(p), go._p(np.ascontiguousarray(cam, np.float32)), C.c_int(0), C.c_int(0), C.c_int(0),
                        C.c_int(rect[2]), C.c_int(rect[3]), go._p(og), go._p(nr))
    d0 = cube.directions(4)
    c0 = np.zeros((rect[2] * rect[3] * 16, 4), np.float32); c0[:, :3] = 2.0
    E = np.zeros((H, W, 4), np.float32)
    go.lib().rco_gather(C.byref(p), go._p(np.ascontiguousarray(cam, np.float32)), C.c_int(0), C.c_int(0), C.c_int(rect[2]), C.c_int(rect[3]),
                        go._p(og), go._p(c0), go._p(d0), go._p(gb["depth"]), go._p(gb["normal"]), go._p(E))
    geo = gb["depth"] >= 0
    assert geo.mean() > 0.5
    assert np.all(np.abs(E[geo][:, :3] / 2.0 - np.pi) < 0.6)      # 16-direction quadrature of the cosine lobe
    assert np.all(E[~geo] == 0)


def test_tile_rect_recursion_covers_the_merge_footprint():
    for tile in [(0, 0, 192, 108), (64, 32, 128, 64), (150, 70, 42, 38), (0, 100, 192, 8)]:
        rects = go.level_rects(192, 108, 4, 6, tile)
        for i in range(5):
            lx, ly, lw, lh = rects[i]
            ux, uy, uw, uh = rects[i + 1]
            gw, gh = -(-192 // (4 << (i + 1))), -(-108 // (4 << (i + 1)))
            for q, (u0, un, g) in ((lx, (ux, uw, gw)), (lx + lw - 1, (ux, uw, gw)), (ly, (uy, uh, gh)), (ly + lh - 1, (uy, uh, gh))):
                base = q // 2 - 1 if q % 2 == 0 else (q - 1) // 2
                for k in (base, base + 1):
                    k = max(0, min(k, g - 1))
                    assert u0 <= k < u0 + un


def _clip_camera(name, W, H, near, far, frame=5):
    import math
    import radiancecascade_b200 as rc
    from oracle import ref_ingest as ri
    osc = go.OracleScene(rc.scenes.scene_path(name))
    pos, tgt, _, _ = rc.scenes.orbit_camera(osc.bbox_min, osc.bbox_max, frame)
    cam = ri.uniform_camera_look_at(pos, tgt, np.float32(math.radians(45.0)), np.float32(W) / np.float32(H), near, far)
    return osc, cam


@pytest.mark.parametrize("name,W,H,near,far", [("teapot", 320, 180, 0.1, 100.0), ("living_room", 320, 180, 0.1, 100.0),
                                               ("cube", 256, 256, 2.0, 4.0), ("test_room", 320, 180, 3.0, 9.0),
                                               ("sonic", 200, 260, 0.1, 100.0)])
def test_clipped_ray_cast_equals_the_raster_restatement(name, W, H, near, far):
    """rc_spec.h S4b against the independent restatement of the reference's render pass (rco_raster: clip-space polygon
    clipping at the near / far planes, scan conversion at pixel centres, depth test Less in draw order,
    src/renderer.rs:332-360): same triangle in every pixel except a few edge pixels; where they agree, the perspective-correct
    barycentrics and the depth agree to float tolerance.  near 0.1 / far 100 are the reference's own values (src/app.rs:26):
    most of the teapot (extent 154) lies beyond its far plane."""
    osc, cam = _clip_camera(name, W, H, near, far)
    p = osc.params(W, H, clip=True)
    lights = np.array([[0, 0, 0, 1]], np.float32)
    gb = osc.gbuffer(p, cam, lights)
    ra = osc.raster(p, cam)
    same = gb["prim"] == ra["prim"]
    assert same.mean() >= 0.998, same.mean()
    hit = same & (gb["prim"] != 0xFFFFFFFF)
    assert hit.sum() > 100
    # depth: z_ndc of the ray's hit point through the same matrix
    from oracle.gi_oracle import OracleScene
    b = OracleScene.primary_basis(cam).reshape(3, 3).astype(np.float64)      # Dx, Dy, Dc
    ys, xs = np.nonzero(hit)
    nx = (2 * xs + 1) / W - 1.0
    ny = 1.0 - (2 * ys + 1) / H
    q = nx[:, None] * b[0] + ny[:, None] * b[1] + b[2]
    d = q / np.linalg.norm(q, axis=1, keepdims=True)
    P = cam[16:19].astype(np.float64) + gb["depth"][ys, xs, None].astype(np.float64) * d
    M = cam[:16].reshape(4, 4).T.astype(np.float64)                         # column-major -> rows
    clip = P @ M[:, :3].T + M[:, 3]
    z = clip[:, 2] / clip[:, 3]
    assert np.abs(z - ra["zndc"][ys, xs]).max() < 2e-4
    assert np.all((z > -1e-5) & (z < 1 + 1e-5))
    gbu = osc.render(p, cam, lights, want_hits=False)        # same G-buffer through the full path (probes use the range too)
    assert np.array_equal(gbu["prim"], gb["prim"])
    unc = osc.gbuffer(osc.params(W, H, clip=False), cam, lights)
    if name in ("teapot", "cube", "test_room"):
        assert (unc["prim"] != gb["prim"]).mean() > 0.05     # the planes really cut something away


def test_floating_probes_leave_no_lower_probe_without_an_upper_probe():
    """rc_spec.h S6 with RC_CFG_FLOATING_PROBES: a probe whose anchor pixel sees no geometry floats to the first finer-level
    anchor inside its cell that does.  On the teapot's silhouette (78 % of the frame is background) that makes every valid probe of level i have at least
    one valid probe among its four upper probes (S1), so S8 never falls back to "far field = sky" next to geometry; probes whose
    whole cell is empty stay invalid, and a probe whose own anchor hits is exactly where it was before the rule existed."""
    name, W, H = "teapot", 480, 270
    osc = go.OracleScene(rc.scenes.scene_path(name))
    pos, tgt, zn, zf = rc.scenes.orbit_camera(osc.bbox_min, osc.bbox_max, 5)
    import math
    cam = ri.uniform_camera_look_at(pos, tgt, np.float32(math.radians(45.0)), np.float32(W) / np.float32(H), zn, zf)
    p = osc.params(W, H, floating=True)
    out = osc.render(p, cam, np.array([[0, 0, 0, 1]], np.float32))
    lv, rects = out["levels"], out["rects"]
    depth = out["depth"]
    # without the flag some valid lower probes have no valid upper probe (the limitation the flag removes)
    plain = osc.render(osc.params(W, H), cam, np.array([[0, 0, 0, 1]], np.float32))
    orphans = 0
    for i in range(p.N - 1):
        _, _, sw, sh = rects[i]
        _, _, usw, ush = rects[i + 1]
        v = (plain["origins"][i][:, 3] != 0).reshape(sh, sw)
        uv = (plain["origins"][i + 1][:, 3] != 0).reshape(ush, usw)
        ys, xs = np.nonzero(v)
        from common import _upper_pair
        x0, x1 = _upper_pair(xs, usw)
        y0, y1 = _upper_pair(ys, ush)
        orphans += int((~(uv[y0, x0] | uv[y0, x1] | uv[y1, x0] | uv[y1, x1])).sum())
    assert orphans > 0
    floated = 0
    for i in range(p.N):
        _, _, sw, sh = rects[i]
        valid = (out["origins"][i][:, 3] != 0).reshape(sh, sw)
        P = lv[i].P
        ay = np.minimum(np.arange(sh) * P + P // 2, H - 1)
        ax = np.minimum(np.arange(sw) * P + P // 2, W - 1)
        own = depth[np.ix_(ay, ax)] >= 0
        assert np.all(valid[own])                                  # an anchor that hits always gives a valid probe
        floated += int((valid & ~own).sum())
        # a cell without any geometry at its level-0 anchors stays invalid
        if i + 1 < p.N:
            _, _, usw, ush = rects[i + 1]
            uvalid = (out["origins"][i + 1][:, 3] != 0).reshape(ush, usw)
            ys, xs = np.nonzero(valid)
            from common import _upper_pair
            x0, x1 = _upper_pair(xs, usw)
            y0, y1 = _upper_pair(ys, ush)
            has = uvalid[y0, x0] | uvalid[y0, x1] | uvalid[y1, x0] | uvalid[y1, x1]
            assert np.all(has), (i, int((~has).sum()))
    assert floated > 20


@pytest.mark.parametrize("name", ["living_room", "test_room"])
def test_rays_that_miss_the_padded_scene_box_hit_nothing(name):
    """The principle behind the product's split ray lists (k_split), checked on CPU with the oracle alone: a cascade
    ray whose float32 slab test against the scene's bounding box — padded by 1e-4 of its diagonal like the product's BVH boxes,
    clipped to the level's interval [t0, t1) — fails has no hit in the brute-force closest-hit search (rc_spec.h S5).  In the
    orbit's open views that is a large share of the rays of every level (the reason the split pays)."""
    W, H = 320, 180
    osc = go.OracleScene(rc.scenes.scene_path(name))
    pos, tgt, zn, zf = rc.scenes.orbit_camera(osc.bbox_min, osc.bbox_max, 5)
    import math
    cam = ri.uniform_camera_look_at(pos, tgt, np.float32(math.radians(45.0)), np.float32(W) / np.float32(H), zn, zf)
    out = osc.render(osc.params(W, H), cam, np.array([[0, 0, 0, 1]], np.float32))
    f32 = np.float32
    lo, hi = osc.bbox_min.astype(f32), osc.bbox_max.astype(f32)
    pad = f32(1e-4) * f32(np.linalg.norm((hi - lo).astype(np.float64)))
    lo, hi = lo - pad, hi + pad
    rng = np.random.default_rng(3)
    shares = []
    for i, lv in enumerate(out["levels"][:5]):
        og = out["origins"][i]
        probes = np.nonzero(og[:, 3] != 0)[0]
        dirs = out["dirs"][i].astype(f32)
        n = 4000
        pi = probes[rng.integers(0, len(probes), n)]
        di = rng.integers(0, len(dirs), n)
        o, d = og[pi, :3].astype(f32), dirs[di]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = (f32(1.0) / np.where(np.abs(d) > f32(1e-20), d, np.copysign(f32(1e-20), d))).astype(f32)
            ta, tb = ((lo - o) * inv).astype(f32), ((hi - o) * inv).astype(f32)
        tn = np.maximum(np.minimum(ta, tb).max(1), f32(lv.t0))
        tf = np.minimum(np.maximum(ta, tb).min(1), f32(lv.t1))
        miss = ~(tn <= tf)
        shares.append(float(miss.mean()))
        if not miss.any():
            continue
        rays = np.concatenate([o[miss], np.full((int(miss.sum()), 1), lv.t0, f32), d[miss], np.full((int(miss.sum()), 1), lv.t1, f32)], 1)
        hits = osc.trace(rays, brute=True)
        assert np.all(hits[:, 3].view(np.uint32) == 0xFFFFFFFF), f"level {i}: a ray that misses the padded box hits a triangle"
    assert max(shares) > 0.2      # random directions of valid probes: a good part leaves the scene at once
