/*
 * rc_spec.h — the frozen GI specification (builder-owned; NOT in the reference).
 *
 * The reference repository contains no radiance-cascade code, so
 * the cascade layout, ray march, merge and gather are defined HERE, following
 * SURVEY.md Appendix C.  The CUDA kernels (radiancecascade_b200/csrc) and the
 * CPU oracle (oracle/) each restate these formulas independently; tests check
 * integer tables bit-exact and radiometric outputs within the tolerances at
 * the end of this file.  Every floating-point formula below fixes its
 * operation order; `fma(a,b,c)` means one correctly rounded a*b+c, every other
 * operation is a correctly rounded IEEE-754 binary32 operation (the CUDA side
 * is compiled with --fmad=false, the oracle with -ffp-contract=off, so neither
 * compiler invents or removes an fma).
 *
 * Notation: W,H full-frame size in pixels; level i in [0,N).
 *
 * ---- S1. layout (integer; bit-exact) --------------------------------------
 *   P_i = P0 << i                     probe spacing in pixels
 *   D_i = D0 << i                     directions per axis (D_i^2 directions)
 *   G_i = (ceil(W/P_i), ceil(H/P_i))  probe grid
 *   anchor(px,py) = ( min(px*P_i + P_i/2, W-1), min(py*P_i + P_i/2, H-1) )   integer pixel
 *   storage, PROBE-MAJOR:  texel(px,py,dx,dy) = ((py*G_i.x + px)*D_i + dy)*D_i + dx
 *     (one probe's directions are contiguous: a level-0 probe is one 128-byte
 *     line; SURVEY C.1 proposed direction-major and asked to evaluate
 *     probe-major — probe-major is what makes a warp's rays share an origin)
 *   texel format: 4 x float16 (radiance r,g,b, transmittance a), 8 bytes
 *   children of direction (dx,dy) at level i: (2dx+{0,1}, 2dy+{0,1}) at level i+1
 *   upper probes of probe p at level i (same rule for x and y):
 *       p even: base = p/2 - 1, weights (0.25, 0.75);  p odd: base = (p-1)/2, weights (0.75, 0.25)
 *       indices base, base+1 clamped to [0, G_{i+1}-1]
 *   level-0 probes of pixel x:  s = x - P0/2;  base = floor_div(s, P0);  f = (s - base*P0)/P0
 *       weights (1-f, f); indices clamped to [0, G_0-1]
 *
 * ---- S2. intervals ---------------------------------------------------------
 *   t_i = L0 * (4^i - 1)/3 (computed in double, rounded to float); level i
 *   covers [t_i, t_{i+1}); the top level covers [t_{N-1}, t_far).
 *   defaults: L0 = bbox_diag/256, t_far = 4*bbox_diag, bbox_diag in float.
 *
 * ---- S3. directions (equal-area octahedral, Clarberg 2008) -----------------
 *   u = (2dx+1)/D - 1, v = (2dy+1)/D - 1   (double)
 *   d = 1-(|u|+|v|); r = 1-|d|; phi = r==0 ? 0 : (pi/4)*((|v|-|u|)/r + 1);
 *   f = r*sqrt(2-r*r);  dir = ( copysign(f*cos(phi),u), copysign(f*sin(phi),v), copysign(1-r*r,d) )
 *   evaluated in double on the host, rounded to float; every texel subtends
 *   exactly 4*pi/D^2 sr and the four children tile their parent.
 *
 * ---- S4. primary rays / G-buffer ------------------------------------------
 *   From the 80-byte camera uniform (view_proj column-major M, eye):
 *   Minv = M^-1 in double; A = col0(Minv), B = col1(Minv), C = col2(Minv)+col3(Minv);
 *   Dx = A.xyz - eye*A.w, Dy = B.xyz - eye*B.w, Dc = C.xyz - eye*C.w, all three
 *   scaled by sign(C.w)/|Dc| and rounded to float.
 *   pixel (x,y): nx = float(2x+1)/float(W) - 1;  ny = 1 - float(2y+1)/float(H);
 *   q = fma(nx, Dx, fma(ny, Dy, Dc)) per component;  dir = q * (1/sqrt(dot(q,q)))
 *   dot(a,b) = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x)).
 *   closest hit over t in [0, FLT_MAX).  depth target = t (ray distance), -1 on miss.
 *   albedo / direct targets = fs_main at the hit with V = -dir and the reference's sampler
 *   (src/texture.rs:132-140: MirrorRepeat, mag Linear, min Nearest, one mip): the texture footprint
 *   is the difference between the texcoords of the +x / +y neighbour pixels' rays on the hit
 *   triangle's plane (S5 barycentrics without the inside tests) and this pixel's;
 *   rho = max(|d(uv)/dx * size|, |d(uv)/dy * size|); rho <= 1 -> bilinear after the sRGB decode
 *   (texel centres at integer + 0.5), otherwise, or when a neighbour ray is parallel to the plane, nearest.
 *
 * ---- S4b. the reference's clip volume (optional: RC_CFG_RASTER_CLIP) -------------
 *   The reference's render pass keeps, per pixel, the closest fragment between the near and far planes of its projection
 *   (perspective_rh, depth 0..1: src/camera.rs:77-79; Depth32Float / Less / clear 1.0: src/renderer.rs:354-360, 585-592).
 *   With the flag every primary ray (pixels and probe anchors) is limited to that range, derived from view_proj alone:
 *   r2 = (M[2], M[6], M[10], M[14]), r3 = (M[3], M[7], M[11], M[15])   rows 2, 3 of the column-major M
 *   z0 = fma(r2.z,e.z, fma(r2.y,e.y, fma(r2.x,e.x, r2.w)));  w0 likewise with r3             (e = eye)
 *   zd = dot(r2.xyz, dir);  wd = dot(r3.xyz, dir)
 *   tmin = zd > 0 ? max(-z0/zd, 0) : 0;      g = zd - wd;  tmax = g > 0 ? (w0 - z0)/g : FLT_MAX
 *   closest hit over [tmin, tmax).  Ties in t go to the lower triangle id = the triangle drawn first, which is what the
 *   depth test Less keeps.  oracle/rc_oracle.c rco_raster restates the pass independently as a clipping scan-converter;
 *   the two agree on the triangle of every pixel except edge pixels (tests/test_oracle_gi.py, tests/test_gpu_parity.py).
 *
 * ---- S5. ray / triangle (two-sided Moller-Trumbore; cull_mode None,
 *          src/renderer.rs:332-343) ------------------------------------------
 *   triangle = (v0, e1 = v1-v0, e2 = v2-v0) with (v0,v1,v2) in the REVERSED
 *   winding the reference draws (src/primitives.rs:369-376)
 *   cross(a,b) = ( fma(a.y,b.z, -(a.z*b.y)), fma(a.z,b.x, -(a.x*b.z)), fma(a.x,b.y, -(a.y*b.x)) )
 *   p = cross(dir,e2); det = dot(e1,p); reject if det == 0 or det is NaN; inv = 1/det
 *   s = o - v0; u = dot(s,p)*inv; reject unless 0 <= u <= 1
 *   q = cross(s,e1); v = dot(dir,q)*inv; reject unless v >= 0 and u+v <= 1
 *   t = dot(e2,q)*inv; accept iff tmin <= t < tmax
 *   a triangle is skipped entirely when any two of its three vertex positions are
 *   bitwise-equal floats (the zero-area triangles tobj makes from `l` / `p` elements)
 *   closest hit = minimum (t, global triangle id) lexicographically — independent
 *   of any acceleration structure.  Global triangle id = triangles of model 0,
 *   then model 1, ... in index-buffer order.
 *
 * ---- S6. probes -------------------------------------------------------------
 *   probe(px,py) of level i: primary ray through anchor(px,py); invalid on a miss.
 *   Optional (RC_CFG_FLOATING_PROBES): on a miss the probe FLOATS instead: the candidates are the anchors
 *   of the probes of the finer levels l = i-1, ..., 0 that lie inside its cell — (qx,qy) with px*2^(i-l) <= qx < (px+1)*2^(i-l)
 *   (likewise y), inside the level-l grid — taken level by level downwards and row-major (qy, then qx) within a level; the
 *   first candidate whose primary ray hits places the probe (hit point, direction and triangle of THAT ray below).  Invalid
 *   only if none hits.  Consequence: a valid probe of level i always has a valid upper probe (the one whose cell contains its
 *   own cell lists its anchor and all of its candidates), so the far field of S8 is never dropped at a silhouette.  Without
 *   the flag a lower probe none of whose four upper probes is valid takes the sky as its whole far field (S8).  The flag costs
 *   4-8 % of a frame on one GPU and ~10 % per rank of a tiled frame (halo probes must trace their candidates): off by default.
 *   hit point h = fma(t, dir, eye);  ng = normalize(cross(e1,e2)), negated if dot(ng,dir) > 0
 *   origin = fma(offset, ng, h);  offset default L0/16.
 *   normalize(x) = x * (1/sqrt(dot(x,x)))   (glam 0.29 form)
 *
 * ---- S7. march --------------------------------------------------------------
 *   texel (probe, direction w): closest hit of origin + t*w, t in [t_i, t_{i+1}).
 *   hit:  (rgb, a) = (Ke + shade(hit, V = -w), 0)
 *   miss: (0,0,0, 1);  top-level miss: (sky, 1);  invalid probe: (0,0,0,1)
 *   shade = fs_main of the reference (src/shader.wgsl:76-100) with:
 *   attributes interpolated as fma(a2, v, fma(a1, u, a0*((1-u)-v))); world
 *   position fma(t, w, origin); textures sampled at the NEAREST texel with
 *   MirrorRepeat addressing (the reference's min filter; a ray has no
 *   derivatives); sRGB decode through a 256-entry table computed in double;
 *   pow(x, Ns) with C powf semantics (pow(x,0) = 1); max(x,0) with C fmaxf semantics
 *   (a NaN operand yields 0: WGSL leaves max(NaN,0) to the implementation; the cube's
 *   NaN tangents exercise this); every radiance component is finally clamped to
 *   [0, 65504] with NaN -> 0; diffuse and specular summed over lights; no shadow
 *   rays (the reference has none).
 *
 * ---- S8. merge (level i+1 -> i, i = N-2 .. 0) ------------------------------
 *   for the 4 upper probes k (S1) of probe p:
 *     delta = o_k - o_p; l2 = dot(delta,delta); h = dot(n_p, delta)
 *     g_k = l2 > 0 ? 1/(1 + RC_PLANE_K * (h*h)/l2) : 1
 *     w_k = bilinear_k * g_k * valid_k          (bilinear_k = wx*wy)
 *   S = ((w_0 + w_1) + w_2) + w_3;  if S <= 0 (no valid upper probe): far = (sky, 0)  else
 *   far = sum_k (w_k/S) * 0.25*(((c_k0 + c_k1) + c_k2) + c_k3)   accumulated k = 0..3 with fma,
 *   c_kj = level i+1 texel of probe k, child j (j = 2*(cy) + cx), read back from float16
 *   merged.rgb = fma(raw.a, far.rgb, raw.rgb);  merged.a = raw.a * far.a;  stored as float16 (RN)
 *   k order: (x0,y0), (x1,y0), (x0,y1), (x1,y1).
 *
 * ---- S9. gather --------------------------------------------------------------
 *   pixel (x,y) with geometry: shading normal n = decode of the stored snorm16
 *   octahedral normal; h = hit point; the 4 level-0 probes k with weights as in S8
 *   using n_p := n, o_p := h.
 *   cs_d = max(dot(n, w_d), 0);  C = sum_d cs_d (d in storage order, plain adds);  q = C > 0 ? pi / C : 0
 *   E = sum_k ((w_k/S) * q) * sum_d cs_d * c_k,d.rgb      (d in storage order, fma accumulation; k = 0..3, fma)
 *   The cosine lobe is integrated with NORMALISED weights (q instead of 4*pi/D0^2): with D0 = 4
 *   half of the upper-hemisphere texel centres lie on the equator and the plain midpoint rule
 *   returns 0.75*pi for a uniform environment; normalising makes a constant radiance field L
 *   gather to exactly pi*L for every normal (tests/test_oracle_gi.py).
 *   output float16 (E.rgb, 1); pixels without geometry: (0,0,0,0).
 *
 * ---- S10. tolerances ---------------------------------------------------------
 *   layout tables, prim-id buffer, hit/miss classification: bit-exact (hit t equal as floats)
 *   direct / albedo / radiance:   |gpu - oracle| <= 2e-3 * max(1,|oracle|)   (float16 storage; powf ulp differences)
 *   irradiance:                   max-abs <= 1e-2 * max(E) and PSNR >= 50 dB
 */
#ifndef RC_SPEC_H
#define RC_SPEC_H

#define RC_DEFAULT_P0 4u
#define RC_DEFAULT_D0 4u
#define RC_DEFAULT_LEVELS 6u
#define RC_MAX_LEVELS 10u
#define RC_INTERVAL0_DIVISOR 256.0f   /* L0 = bbox_diag / 256 */
#define RC_TFAR_FACTOR 4.0f           /* t_far = 4 * bbox_diag */
#define RC_OFFSET_DIVISOR 16.0f       /* probe lift = L0 / 16 */
#define RC_PLANE_K 16.0f              /* plane-distance rejection strength in S8/S9 */
#define RC_MAX_LIGHTS 8u
#define RC_PI_F 3.14159274101257324219f /* float(pi) */

/* Lighting constants of the reference's fs_main (src/shader.wgsl:83,93,97,99). */
#define RC_AMBIENT_COEF 0.05f
#define RC_DIFFUSE_COEF 0.7f
#define RC_SPECULAR_COEF 1.0f
#define RC_UNLIT_EPS 1e-5f
#define RC_NDOTV_EPS 1e-6f

#endif /* RC_SPEC_H */
