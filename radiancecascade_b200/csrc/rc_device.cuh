// rc_device.cuh — device-side data layout, spec arithmetic (rc_spec.h S4-S7), BVH
// traversal and fs_main restatement shared by the sm_100a kernels.
// Compiled with --fmad=false: every fma below is explicit (rc_spec.h preamble).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rc_spec.h"

namespace rc {

// ------------------------------------------------------------------ layout in HBM
struct DMaterial {          // UniformMaterial (src/primitives.rs:37-46) + GI extras + per-material constants of fs_main
    float ka[4], kd[4], ks[4];
    float ns;
    uint32_t ebit;          // enable_bit (src/renderer.rs:422-423)
    int32_t tex_c, tex_n;   // texture ids or -1 (Texture::empty)
    float ke[4];
    // terms of fs_main that depend on the material only, evaluated once on the host with the very expressions
    // shade_hit used to evaluate per hit (f32, no contraction): ambient Ka*0.05*Ka.w (:82-83), the "unlit" selector
    // (:99-100) and whether the specular chain can contribute at all
    float amb[3];
    float unlit;            // 1.0 or 0.0
    uint32_t spec;          // (Ks present and non-zero) or Ns < 0 / NaN
    uint32_t pad[3];
};

struct DTexture { uint64_t offset; uint32_t w, h; };

struct DScene {
    const float4* nodes;        // 4 x float4 per BVH node (bvh.h)
    const float4* tri_geom;     // leaf order, 3 x float4 per triangle: (v0, id bits), (e1, 0), (e2, 0)
    const float4* tri_eg;       // global-id order, 2 x float4: e1, e2 (geometric normal for probes)
    const uint32_t* tris;       // [NT][3] global vertex ids, reversed winding (src/primitives.rs:369-376)
    const uint32_t* tri_model;  // [NT]
    const float* verts;         // [NV][17] interleaved stream (src/renderer.rs:371-410)
    const DMaterial* mats;      // [models]
    const DTexture* tex;
    const uint8_t* tex_data;    // RGBA8 texels
    const float* srgb;          // 256-entry sRGB decode table
};

struct DLights { float pos[RC_MAX_LIGHTS][4]; int n; uint32_t flags; };

struct DCamera {                // rc_spec.h S4
    float3 eye, dx, dy, dc;
    int W, H;
    // S4b (RC_CFG_RASTER_CLIP): rows 2 and 3 of view_proj, .w = the row applied to (eye, 1); clip == 0: rays see [0, FLT_MAX)
    float4 clip_z, clip_w;
    int clip;
    // rows 0, 1, 3 of view_proj (x_clip, y_clip, w_clip of a world point): the triangle binning of the primary-visibility pass
    float4 row_x, row_y, row_w;
    // conservative screen-space rectangle of the scene's bounding box (pixels, inclusive; the whole frame when a corner of the
    // box is behind the eye): no primary ray outside it hits anything — k_probes' S6 search skips cells that miss it
    int sb_x0, sb_y0, sb_x1, sb_y1;
};

struct DLevel {
    int P, D, gw, gh;           // spacing, direction res, full-frame grid
    int px0, py0, sw, sh;       // sub-grid held by this context
    float t0, t1;
    unsigned long long texel_offset;  // texels before this level in the cascade buffer
    unsigned int probe_offset;        // probes before this level in the probe arrays
};

struct DLevelSet { DLevel lv[RC_MAX_LEVELS]; int n; };

// Host side of the per-material constants (called once per material at scene load; plain f32, no contraction —
// the expressions are the ones fs_main spells out, src/shader.wgsl:82-83, 97-100).
inline void fill_material_constants(DMaterial& m)
{
    for (int k = 0; k < 3; k++) m.amb[k] = m.ka[k] * 0.05f * m.ka[3];
    const float pred = ((m.ka[0] - 1e-5f) + (m.kd[0] - 1e-5f) + (m.ks[0] - 1e-5f))
                     + ((m.ka[1] - 1e-5f) + (m.kd[1] - 1e-5f) + (m.ks[1] - 1e-5f))
                     + ((m.ka[2] - 1e-5f) + (m.kd[2] - 1e-5f) + (m.ks[2] - 1e-5f));                   // :99
    m.unlit = pred <= 0.0f ? 1.0f : 0.0f;
    m.spec = ((m.ks[3] != 0.0f && (m.ks[0] != 0.0f || m.ks[1] != 0.0f || m.ks[2] != 0.0f)) || !(m.ns >= 0.0f)) ? 1u : 0u;
}

// ------------------------------------------------------------------ spec arithmetic
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 vsub(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 vadd(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 vscale(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 vneg(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float vdot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ float3 vcross(float3 a, float3 b)
{
    return f3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ float3 vnormalize(float3 a) { float r = 1.0f / sqrtf(vdot(a, a)); return vscale(a, r); }
__device__ __forceinline__ float3 vfma(float s, float3 d, float3 o) { return f3(fmaf(s, d.x, o.x), fmaf(s, d.y, o.y), fmaf(s, d.z, o.z)); }
__device__ __forceinline__ float3 xyz(float4 a) { return f3(a.x, a.y, a.z); }

// S4 primary ray through full-frame pixel (x, y)
__device__ __forceinline__ float3 primary_dir(const DCamera& c, int x, int y)
{
    float nx = (float)(2 * x + 1) / (float)c.W - 1.0f;
    float ny = 1.0f - (float)(2 * y + 1) / (float)c.H;
    float3 q = f3(fmaf(nx, c.dx.x, fmaf(ny, c.dy.x, c.dc.x)), fmaf(nx, c.dx.y, fmaf(ny, c.dy.y, c.dc.y)),
                  fmaf(nx, c.dx.z, fmaf(ny, c.dy.z, c.dc.z)));
    return vnormalize(q);
}

// S4b: the part of a primary ray between the near and far planes of view_proj (the reference's clip volume, depth 0..1:
// src/camera.rs:77-79) — z_clip(t) = z0 + t*zd >= 0 and z_clip(t) <= w_clip(t) = w0 + t*wd
__device__ __forceinline__ void primary_range(const DCamera& c, float3 d, float& tmin, float& tmax)
{
    tmin = 0.0f; tmax = 3.402823466e+38f;
    if (!c.clip) return;
    const float zd = vdot(xyz(c.clip_z), d), wd = vdot(xyz(c.clip_w), d);
    if (zd > 0.0f) tmin = fmaxf(-c.clip_z.w / zd, 0.0f);
    const float g = zd - wd;
    if (g > 0.0f) tmax = (c.clip_w.w - c.clip_z.w) / g;
}

// ------------------------------------------------------------------ closest hit (S5)
struct Hit { float t, u, v; uint32_t prim; };

__device__ __forceinline__ float safe_inv(float d)
{
    // |d| below 1e-20 would make 1/d overflow; a slab this parallel can only be crossed beyond any tmax
    return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d));
}

// S5 on a triangle already in registers / shared memory: a = (v0, id bits), b = (e1, -), c = (e2, -).
// Bit-identical to tri_test; rejects the two common miss cases before the IEEE division:
//   u < 0:  u = fl(a * fl(1/det)) with a = dot(s, p).  Opposite signs of a and det make the exact product negative, and with
//           |a| > 1e-30 and |det| < 1e30 its magnitude is far above the underflow threshold, so fl() cannot return -0 (which
//           S5's `u >= 0` would accept).
//   u > 1:  |a| > 1.0001 |det| puts the exact quotient above 1 by 1e-4, a thousand ulps beyond what the two roundings move.
__device__ __forceinline__ void tri_test_v(float4 a, float4 b, float4 c, float3 o, float3 d, float tmin, float tmax, Hit& h)
{
    const float3 e1 = xyz(b), e2 = xyz(c);
    const float3 p = vcross(d, e2);
    const float det = vdot(e1, p);
    if (!(det != 0.0f)) return;
    const float3 s = vsub(o, xyz(a));
    const float un = vdot(s, p);
    const float ad = fabsf(det), au = fabsf(un);
    if (ad < 1e30f && ad > 1e-30f) {
        if (((un < 0.0f) != (det < 0.0f)) && au > 1e-30f) return;
        if (au > 1.0001f * ad && au < 1e30f) return;
    }
    const float inv = 1.0f / det;
    const float u = un * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return;
    const float3 q = vcross(s, e1);
    const float v = vdot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return;
    const float t = vdot(e2, q) * inv;
    const uint32_t id = __float_as_uint(a.w);
    if (t >= tmin && t < tmax && (t < h.t || (t == h.t && id < h.prim))) { h.t = t; h.u = u; h.v = v; h.prim = id; }
}

__device__ __forceinline__ void tri_test(const float4* __restrict__ g, float3 o, float3 d, float tmin, float tmax, Hit& h)
{
    const float4 a = __ldg(g), b = __ldg(g + 1), c = __ldg(g + 2);
    const float3 e1 = xyz(b), e2 = xyz(c);
    const float3 p = vcross(d, e2);
    const float det = vdot(e1, p);
    if (!(det != 0.0f)) return;
    const float inv = 1.0f / det;
    const float3 s = vsub(o, xyz(a));
    const float u = vdot(s, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return;
    const float3 q = vcross(s, e1);
    const float v = vdot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return;
    const float t = vdot(e2, q) * inv;
    const uint32_t id = __float_as_uint(a.w);
    if (t >= tmin && t < tmax && (t < h.t || (t == h.t && id < h.prim))) { h.t = t; h.u = u; h.v = v; h.prim = id; }
}

// While-while traversal of the 2-wide BVH with a per-thread stack.  h.t doubles as the
// current far bound (initialised to tmax); on return h.prim == ~0u means miss.
//
// `entry` (optional): the probe's entry frontier (k_entry) — up to RC_ENTRY_SLOTS node / leaf links, valid ones
// first, padded with kDoneLink — a set of subtrees that together hold every triangle within reach of the
// probe's interval.  The traversal then starts with those links on its stack instead of at the root, skipping
// the upper levels of the tree that every ray of the probe would walk through identically.
#ifndef RC_DIST_STACK
#define RC_DIST_STACK 0   // 1: the traversal stack also keeps each deferred subtree's entry distance (A/B build: make EXTRA=-DRC_DIST_STACK=1)
#endif
constexpr int RC_ENTRY_SLOTS = 8;
constexpr int kDoneLinkC = (int)0x80000000;

// trace_inv: `inv` = (safe_inv(d.x), safe_inv(d.y), safe_inv(d.z)) supplied by the caller (k_march reads it from the
// level's direction table); trace computes it.
__device__ __forceinline__ Hit trace_inv(const DScene& s, float3 o, float3 d, float3 inv, float tmin, float tmax,
                                         const int4* __restrict__ entry = nullptr)
{
    Hit h; h.t = tmax; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
    const float3 noi = f3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
    // True while-while (Aila & Laine): every lane descends inner nodes until it holds a leaf (or is
    // done); the warp reconverges at the end of the node loop and tests triangles together.  The
    // earlier "one node, then drain leaves" shape ran 57% of the kernel's warp-instructions — the
    // triangle tests — with 1.5-3 active lanes (profiles/r1_a_*).
    constexpr int kDoneLink = kDoneLinkC;   // never a real leaf code (first < 2^28)
    int stack[48];
#if RC_DIST_STACK
    // entry distance of every deferred subtree: when it is popped after a closer hit has been found it is dropped
    // without visiting its node (its children could only fail the same n <= min(far, h.t) test: child boxes lie
    // inside the parent's, and the slab arithmetic is monotonic) — same hits, fewer node visits
    float dstack[48];
#define RC_PUSH(link, dist) do { stack[sp] = (link); dstack[sp] = (dist); sp++; } while (0)
#define RC_POP() do { cur = kDoneLink; while (sp) { --sp; if (!(dstack[sp] > h.t)) { cur = stack[sp]; break; } } } while (0)
#else
#define RC_PUSH(link, dist) do { stack[sp++] = (link); } while (0)
#define RC_POP() do { cur = sp ? stack[--sp] : kDoneLink; } while (0)
#endif
    int sp = 0;
    int cur = 0;
    if (entry) {
        const int4 ea = __ldg(entry), eb = __ldg(entry + 1);
        if (eb.w != kDoneLink) RC_PUSH(eb.w, 0.0f);
        if (eb.z != kDoneLink) RC_PUSH(eb.z, 0.0f);
        if (eb.y != kDoneLink) RC_PUSH(eb.y, 0.0f);
        if (eb.x != kDoneLink) RC_PUSH(eb.x, 0.0f);
        if (ea.w != kDoneLink) RC_PUSH(ea.w, 0.0f);
        if (ea.z != kDoneLink) RC_PUSH(ea.z, 0.0f);
        if (ea.y != kDoneLink) RC_PUSH(ea.y, 0.0f);
        cur = ea.x;
    }
    while (cur != kDoneLink) {
        while (cur >= 0) {
            const float4* n = s.nodes + 4 * (size_t)cur;
            const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
            // child 0 slabs
            float ax = fmaf(q0.x, inv.x, noi.x), bx = fmaf(q0.w, inv.x, noi.x);
            float ay = fmaf(q0.y, inv.y, noi.y), by = fmaf(q1.x, inv.y, noi.y);
            float az = fmaf(q0.z, inv.z, noi.z), bz = fmaf(q1.y, inv.z, noi.z);
            const float n0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
            const float f0 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));
            // child 1 slabs
            ax = fmaf(q1.z, inv.x, noi.x); bx = fmaf(q2.y, inv.x, noi.x);
            ay = fmaf(q1.w, inv.y, noi.y); by = fmaf(q2.z, inv.y, noi.y);
            az = fmaf(q2.x, inv.z, noi.z); bz = fmaf(q2.w, inv.z, noi.z);
            const float n1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
            const float f1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));
            const bool hit0 = n0 <= f0, hit1 = n1 <= f1;   // <=: equal-t candidates stay reachable for the id tie-break
            const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            // select-based step: one predicated push, one (rare) pop branch.  (prefetch.global.L1 of the deferred
            // child was measured: +-1 %, not kept)
            const bool first1 = hit1 && (!hit0 || n1 < n0);
            const int nearc = first1 ? c1 : c0, farc = first1 ? c0 : c1;
            if (hit0 && hit1) RC_PUSH(farc, first1 ? n0 : n1);
            if (hit0 || hit1) cur = nearc;
            else RC_POP();
        }
        while (cur < 0 && cur != kDoneLink) {
            const uint32_t leaf = (uint32_t)~cur;
            const uint32_t first = leaf >> 3, cnt = leaf & 7u;
            for (uint32_t i = 0; i < cnt; i++) tri_test(s.tri_geom + 3 * (size_t)(first + i), o, d, tmin, tmax, h);
            RC_POP();
        }
    }
#undef RC_PUSH
#undef RC_POP
    if (h.prim == 0xffffffffu) h.t = -1.0f;
    return h;
}

__device__ __forceinline__ Hit trace(const DScene& s, float3 o, float3 d, float tmin, float tmax, const int4* __restrict__ entry = nullptr)
{
    return trace_inv(s, o, d, f3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z)), tmin, tmax, entry);
}

// ------------------------------------------------------------------ shading (S7; src/shader.wgsl:76-100)
__device__ __forceinline__ int mirror_idx(long long i, long long size)   // Vulkan MirrorRepeat
{
    long long m = i % (2 * size);
    if (m < 0) m += 2 * size;
    m -= size;
    if (m < 0) m = -(1 + m);
    return (int)((size - 1) - m);
}

__device__ __forceinline__ float3 sample_nearest(const DScene& s, int tex, float u, float v, bool srgb)
{
    if (tex < 0) return f3(0.f, 0.f, 0.f);   // Texture::empty (src/texture.rs:12-64)
    const DTexture t = s.tex[tex];
    float fu = floorf(u * (float)t.w), fv = floorf(v * (float)t.h);
    if (!(fabsf(fu) < 1e9f)) fu = 0.0f;
    if (!(fabsf(fv) < 1e9f)) fv = 0.0f;
    const int ix = mirror_idx((long long)fu, t.w), iy = mirror_idx((long long)fv, t.h);
    const uchar4 p = *reinterpret_cast<const uchar4*>(s.tex_data + t.offset + 4 * ((size_t)iy * t.w + ix));
    if (srgb) return f3(__ldg(s.srgb + p.x), __ldg(s.srgb + p.y), __ldg(s.srgb + p.z));
    return f3(p.x / 255.0f, p.y / 255.0f, p.z / 255.0f);
}

__device__ __forceinline__ float3 fetch_texel(const DScene& s, const DTexture& t, long long ix, long long iy, bool srgb)
{
    const int x = mirror_idx(ix, t.w), y = mirror_idx(iy, t.h);
    const uchar4 p = *reinterpret_cast<const uchar4*>(s.tex_data + t.offset + 4 * ((size_t)y * t.w + x));
    if (srgb) return f3(__ldg(s.srgb + p.x), __ldg(s.srgb + p.y), __ldg(s.srgb + p.z));
    return f3(p.x / 255.0f, p.y / 255.0f, p.z / 255.0f);
}

// The reference's sampler (src/texture.rs:132-140): MirrorRepeat, mag Linear, min Nearest, one mip.  `duv` holds the
// texture-coordinate differences to the pixel's +x and +y neighbours (du/dx, dv/dx, du/dy, dv/dy); the filter is
// chosen as the API specifies: rho = max(|d(uv)/dx * size|, |d(uv)/dy * size|), magnification iff log2(rho) <= 0.
// Linear filtering happens after the sRGB decode.  duv == nullptr (GI hit: a ray has no derivatives) -> nearest (S7).
__device__ __forceinline__ float3 sample_tex(const DScene& s, int tex, float u, float v, bool srgb, const float* duv)
{
    if (tex < 0 || duv == nullptr) return sample_nearest(s, tex, u, v, srgb);
    const DTexture t = s.tex[tex];
    const float w = (float)t.w, h = (float)t.h;
    const float ax = duv[0] * w, ay = duv[1] * h, bx = duv[2] * w, by = duv[3] * h;
    const float rho = fmaxf(sqrtf(ax * ax + ay * ay), sqrtf(bx * bx + by * by));
    if (!(rho <= 1.0f)) return sample_nearest(s, tex, u, v, srgb);      // minification (or undefined derivatives)
    float fu = u * w - 0.5f, fv = v * h - 0.5f;
    if (!(fabsf(fu) < 1e9f)) fu = 0.0f;
    if (!(fabsf(fv) < 1e9f)) fv = 0.0f;
    const float iu = floorf(fu), iv = floorf(fv);
    const float a = fu - iu, b = fv - iv;
    const long long x0 = (long long)iu, y0 = (long long)iv;
    const float3 c00 = fetch_texel(s, t, x0, y0, srgb), c10 = fetch_texel(s, t, x0 + 1, y0, srgb);
    const float3 c01 = fetch_texel(s, t, x0, y0 + 1, srgb), c11 = fetch_texel(s, t, x0 + 1, y0 + 1, srgb);
    const float ia = 1.0f - a, ib = 1.0f - b;
    const float3 top = f3(c00.x * ia + c10.x * a, c00.y * ia + c10.y * a, c00.z * ia + c10.z * a);
    const float3 bot = f3(c01.x * ia + c11.x * a, c01.y * ia + c11.y * a, c01.z * ia + c11.z * a);
    return f3(top.x * ib + bot.x * b, top.y * ib + bot.y * b, top.z * ib + bot.z * b);
}

// barycentrics of the point where ray (o, d) meets the plane of triangle (v0, e1, e2) — S5 arithmetic without the
// inside tests; used for the texture footprint of primary rays
__device__ __forceinline__ bool plane_bary(float3 v0, float3 e1, float3 e2, float3 o, float3 d, float& u, float& v)
{
    const float3 p = vcross(d, e2);
    const float det = vdot(e1, p);
    if (!(det != 0.0f)) return false;
    const float inv = 1.0f / det;
    const float3 sv = vsub(o, v0);
    u = vdot(sv, p) * inv;
    v = vdot(d, vcross(sv, e1)) * inv;
    return true;
}

__device__ __forceinline__ float lerp3(float a0, float a1, float a2, float u, float v)
{
    const float w = (1.0f - u) - v;
    return fmaf(a2, v, fmaf(a1, u, a0 * w));
}

__device__ __forceinline__ float clamp_rad(float x) { return fminf(fmaxf(x, 0.0f), 65504.0f); }   // NaN -> 0

struct Shade { float3 rad, n, albedo, direct; };

// fs_main at a hit with view vector Vd (= -ray direction); rad = Ke + (L + unlit) * albedo.
// fp (optional): barycentrics (u,v) of the +x and +y neighbour pixels' rays on this triangle's plane -> texture
// footprint for the sampler's mag/min decision (G-buffer pass only).
// NORMAL_ONLY: stop after the two-sided shading normal (all the per-frame G-buffer pass needs).
template <bool NORMAL_ONLY = false>
__device__ __forceinline__ Shade shade_hit(const DScene& s, const DLights& L, uint32_t prim, float u, float v,
                                           float3 P, float3 Vd, const float* fp = nullptr)
{
    const uint32_t i0 = s.tris[3 * (size_t)prim], i1 = s.tris[3 * (size_t)prim + 1], i2 = s.tris[3 * (size_t)prim + 2];
    const float* a = s.verts + 17 * (size_t)i0;
    const float* b = s.verts + 17 * (size_t)i1;
    const float* c = s.verts + 17 * (size_t)i2;
    // fs_main evaluates everything and masks with 0/1 factors; here an attribute is fetched and interpolated
    // only when its factor is not 0 (the result is bit-identical: x*0 contributes exactly 0 for finite x, and
    // the skipped terms are never NaN-producing for the cases gated below).
    const DMaterial& m = s.mats[s.tri_model[prim]];
    uint32_t eb = m.ebit;
    if (!(L.flags & 1u)) eb &= 1u;                                   // src/renderer.rs:623
    const bool b0 = eb & 1u, b1 = (eb >> 1) & 1u;
#define RC_ATTR(k) lerp3(__ldg(a + (k)), __ldg(b + (k)), __ldg(c + (k)), u, v)
    float tu = 0.f, tv = 0.f;
    float duv_store[4];
    const float* duv = nullptr;
    if ((b0 && !NORMAL_ONLY) || b1) {
        tu = RC_ATTR(15); tv = 1.0f - RC_ATTR(16);                   // :78
        if (fp) {
            const float ux = lerp3(__ldg(a + 15), __ldg(b + 15), __ldg(c + 15), fp[0], fp[1]);
            const float vx = 1.0f - lerp3(__ldg(a + 16), __ldg(b + 16), __ldg(c + 16), fp[0], fp[1]);
            const float uy = lerp3(__ldg(a + 15), __ldg(b + 15), __ldg(c + 15), fp[2], fp[3]);
            const float vy = 1.0f - lerp3(__ldg(a + 16), __ldg(b + 16), __ldg(c + 16), fp[2], fp[3]);
            duv_store[0] = ux - tu; duv_store[1] = vx - tv; duv_store[2] = uy - tu; duv_store[3] = vy - tv;
            duv = duv_store;
        }
    }
    float3 albedo = f3(0.f, 0.f, 0.f);
    if (!NORMAL_ONLY) albedo = b0 ? sample_tex(s, m.tex_c, tu, tv, true, duv) : f3(RC_ATTR(3), RC_ATTR(4), RC_ATTR(5));   // :80
    float3 Lc = f3(m.amb[0], m.amb[1], m.amb[2]);                     // :82-83, Ka * 0.05 * Ka.w (fill_material_constants)
    const float3 Nv = f3(RC_ATTR(6), RC_ATTR(7), RC_ATTR(8));
    float3 raw;
    if (b1) {
        const float3 cs = sample_tex(s, m.tex_n, tu, tv, false, duv);
        const float3 cf = f3(cs.x * 2.0f - 1.0f, cs.y * 2.0f - 1.0f, cs.z * 2.0f - 1.0f);             // :85
        const float3 T = vnormalize(f3(RC_ATTR(9), RC_ATTR(10), RC_ATTR(11))), B = vnormalize(f3(RC_ATTR(12), RC_ATTR(13), RC_ATTR(14)));
        raw = vnormalize(vadd(vadd(vscale(T, cf.x), vscale(B, cf.y)), vscale(Nv, cf.z)));             // :86
    } else {
        raw = vnormalize(Nv);
    }
#undef RC_ATTR
    const float ndv = vdot(Vd, raw);                                 // :88
    const float3 N = ndv < 0.0f ? vneg(raw) : raw;                   // :89
    if (NORMAL_ONLY) { Shade r; r.rad = r.albedo = r.direct = f3(0.f, 0.f, 0.f); r.n = N; return r; }
    // specular chain (normalize, powf) only when Ks can contribute: Ks present and non-zero, or Ns < 0
    // (pow(0, Ns<0) = inf must still poison the result exactly as the plain formula does)
    const bool spec = m.spec != 0u;
    for (int li = 0; li < L.n; li++) {
        const float3 lp = f3(L.pos[li][0], L.pos[li][1], L.pos[li][2]);
        const float3 Ld = vnormalize(vsub(lp, P));                   // :91
        const float ndl = fmaxf(vdot(Ld, N), 0.0f);                  // :92
        const float kd = 0.7f * ndl * m.kd[3];
        Lc = f3(fmaf(m.kd[0], kd, Lc.x), fmaf(m.kd[1], kd, Lc.y), fmaf(m.kd[2], kd, Lc.z));           // :93
        if (spec) {
            const float3 Hd = vnormalize(vadd(Vd, Ld));                  // :95
            const float st = powf(fmaxf(vdot(N, Hd), 0.0f), m.ns);       // :96
            const float ks = st * m.ks[3] * (ndv > 1e-6f ? 1.0f : 0.0f); // :97
            Lc = f3(fmaf(m.ks[0], ks, Lc.x), fmaf(m.ks[1], ks, Lc.y), fmaf(m.ks[2], ks, Lc.z));
        }
    }
    const float unlit = m.unlit;                                     // :99-100 (fill_material_constants)
    Shade r;
    r.direct = f3((Lc.x + unlit) * albedo.x, (Lc.y + unlit) * albedo.y, (Lc.z + unlit) * albedo.z);   // :100
    r.rad = f3(clamp_rad(m.ke[0] + r.direct.x), clamp_rad(m.ke[1] + r.direct.y), clamp_rad(m.ke[2] + r.direct.z));
    r.n = N;
    r.albedo = albedo;
    return r;
}

// ------------------------------------------------------------------ packing
__device__ __forceinline__ uint2 pack_half4(float r, float g, float b, float a)
{
    __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    return o;
}
__device__ __forceinline__ float4 unpack_half4(uint2 p)
{
    const float2 lo = __half22float2(*reinterpret_cast<__half2*>(&p.x));
    const float2 hi = __half22float2(*reinterpret_cast<__half2*>(&p.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// standard octahedral encode of the shading normal, 2 x snorm16 (G-buffer only)
__device__ __forceinline__ uint32_t oct_encode(float3 n)
{
    const float s = fabsf(n.x) + fabsf(n.y) + fabsf(n.z);
    float px = n.x / s, py = n.y / s;
    if (n.z < 0.0f) {
        const float qx = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
        const float qy = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
        px = qx; py = qy;
    }
    const int ix = __float2int_rn(fminf(fmaxf(px, -1.0f), 1.0f) * 32767.0f);
    const int iy = __float2int_rn(fminf(fmaxf(py, -1.0f), 1.0f) * 32767.0f);
    return ((uint32_t)(uint16_t)(int16_t)ix) | (((uint32_t)(uint16_t)(int16_t)iy) << 16);
}
// a / 32767 for an integer-valued |a| <= 32768, correctly rounded: the division's own Newton sequence (reciprocal refined
// once, quotient corrected once with the exact remainder) without the generic operand check (FCHK) and its out-of-line
// slow path, which every zero numerator took — i.e. every axis-aligned normal (ncu: half of the living-room pixels).
// Bit-equal to a / 32767.0f for all 65536 inputs: tests/cpp/div32767_check.c (run by tests/test_abi.py).
__host__ __device__ __forceinline__ float div32767(float a)
{
    const float r0 = 3.0518509447574615479e-05f;                 // 0x38000100
    const float r = fmaf(fmaf(r0, -32767.0f, 1.0f), r0, r0);
    const float q = a * r;
    return fmaf(r, fmaf(q, -32767.0f, a), q);
}

__device__ __forceinline__ float3 oct_decode(uint32_t e)
{
    const float px = div32767((float)(int16_t)(e & 0xffffu)), py = div32767((float)(int16_t)(e >> 16));
    const float z = 1.0f - fabsf(px) - fabsf(py);
    float x = px, y = py;
    if (z < 0.0f) {
        x = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
        y = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
    }
    return vnormalize(f3(x, y, z));
}

// S8 plane-distance weight of probe k (origin ok.xyz, ok.w = valid) seen from (op, np)
__device__ __forceinline__ float plane_weight(float3 np, float3 op, float4 ok)
{
    if (ok.w == 0.0f) return 0.0f;
    const float3 delta = vsub(xyz(ok), op);
    const float l2 = vdot(delta, delta), hh = vdot(np, delta);
    return l2 > 0.0f ? 1.0f / (1.0f + RC_PLANE_K * (hh * hh) / l2) : 1.0f;
}

// S1: the two upper probes (clamped) and bilinear weights of probe index q along one axis
__device__ __forceinline__ void upper_pair(int q, int gmax, int& i0, int& i1, float& w0, float& w1)
{
    int base;
    if ((q & 1) == 0) { base = q / 2 - 1; w0 = 0.25f; w1 = 0.75f; }
    else { base = (q - 1) / 2; w0 = 0.75f; w1 = 0.25f; }
    i0 = min(max(base, 0), gmax - 1);
    i1 = min(max(base + 1, 0), gmax - 1);
}

}  // namespace rc
