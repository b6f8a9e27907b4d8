"""Analysis script (CPU, oracle only; not collected by pytest): how many upper-level texels would visibility-driven
culling remove beyond direction culling?  A level-i+1 quad is then requested only by level-i texels that are requested
AND miss inside their own interval.  Result on the bundled scenes (480x270): 3-6 % fewer rays — not pursued.
Usage: python tests/estimate_visibility_culling.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from common import frame_setup, oracle_scene   # noqa: E402

def upper_pair(q, g):
    base = np.where(q % 2 == 0, q // 2 - 1, (q - 1) // 2)
    return np.clip(base, 0, g - 1), np.clip(base + 1, 0, g - 1)

def run(name, W, H, frame=5):
    osc = oracle_scene(name)
    st, cam, lights = frame_setup(name, W, H, frame=frame)
    p = osc.params(W, H, store_half=True)
    out = osc.render(p, cam, lights, keep_raw=True)
    lv = out["levels"]; N = len(lv)
    rects = out["rects"]
    # level-0 request: directions with positive cosine for the probe's own anchor normal (proxy for the pixel OR)
    need = []
    D0 = lv[0].D
    sw0, sh0 = rects[0][2], rects[0][3]
    n0 = out["probe_normals"][0][:, :3]; valid0 = out["origins"][0][:, 3] != 0
    cos = n0 @ out["dirs"][0].T
    R = (cos > -0.2) & valid0[:, None]            # generous: neighbouring pixels' normals differ a little
    Rd = R.copy(); Rv = R.copy()
    tot = []
    for i in range(N):
        D = lv[i].D; sw, sh = rects[i][2], rects[i][3]
        raw = out["raw"][i].reshape(sw * sh, D * D, 4)
        miss = raw[..., 3] != 0
        valid = out["origins"][i][:, 3] != 0
        nd, nv = int(Rd.sum()), int(Rv.sum())
        tot.append((i, sw * sh * D * D, nd, nv, float((miss & Rd).sum()) / max(nd, 1)))
        if i == N - 1: break
        # push up: requested lower texel (p, d) -> quad of d at 4 upper probes
        usw, ush = rects[i + 1][2], rects[i + 1][3]
        UD = lv[i + 1].D
        validu = out["origins"][i + 1][:, 3] != 0
        py, px = np.divmod(np.arange(sw * sh), sw)
        x0, x1 = upper_pair(px, usw); y0, y1 = upper_pair(py, ush)
        ups = [y0 * usw + x0, y0 * usw + x1, y1 * usw + x0, y1 * usw + x1]
        def push(Rl):
            # Rl [probes, D*D] bool -> upper [uprobes, UD*UD]
            Ru = np.zeros((usw * ush, UD, UD), bool)
            Rl3 = Rl.reshape(sw * sh, D, D)
            ex = np.repeat(np.repeat(Rl3, 2, axis=1), 2, axis=2)   # [probes, UD, UD]
            for u in ups:
                np.logical_or.at(Ru, u, ex)
            Ru &= validu[:, None, None]
            return Ru.reshape(usw * ush, UD * UD)
        Rd_next = push(Rd & valid[:, None])
        Rv_next = push(Rv & valid[:, None] & miss)
        Rd, Rv = Rd_next, Rv_next
    print(name, W, H)
    sd = sv = 0
    for i, n, nd, nv, mf in tot:
        print(f"  L{i}: texels {n:9d}  dir-culled {nd:9d} ({nd/n:.3f})  +visibility {nv:9d} ({nv/n:.3f})  miss-frac-of-requested {mf:.2f}")
        sd += nd; sv += nv
    print("  total marched: dir", sd, "vis", sv, "ratio", sv / sd)

run("living_room", 480, 270)
run("teapot", 480, 270)
run("test_room", 480, 270)
