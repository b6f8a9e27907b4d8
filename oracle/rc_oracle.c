/*
 * rc_oracle.c — CPU oracle of the radiance-cascade GI path (plain C + OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product
 * (radiancecascade_b200/) never links, imports or calls this file.
 *
 * PARITY UNPINNED: the reference (jw910731/RadianceCascade) has no ray march,
 * cascade merge or irradiance gather and no tests or golden
 * vectors.  This file is an independent restatement of the
 * builder-owned GI specification include/rc_spec.h (sections S1..S9 are cited
 * per function) plus the reference's direct-lighting arithmetic
 * (src/shader.wgsl:76-100, cited in rco_shade).  Compile with
 *   gcc -O2 -fopenmp -mfma -mf16c -ffp-contract=off
 * so that fmaf() is one rounding and nothing else is contracted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* S4: dot(a,b) = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x)) */
static inline float vdot(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
/* S5: cross with one fma per component */
static inline v3 vcross(v3 a, v3 b)
{
    return V(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
/* S6: normalize(x) = x * (1/sqrt(dot(x,x))) */
static inline v3 vnormalize(v3 a) { float r = 1.0f / sqrtf(vdot(a, a)); return vscale(a, r); }
/* fma(s, d, o) per component */
static inline v3 vfma(float s, v3 d, v3 o) { return V(fmaf(s, d.x, o.x), fmaf(s, d.y, o.y), fmaf(s, d.z, o.z)); }

static inline float half_round(float x) { return (float)(_Float16)x; }

/* ------------------------------------------------------------------ scene */
typedef struct {
    int n_verts, n_tris, n_models;
    const float* verts;       /* [n_verts][17]  (src/renderer.rs:371-410 layout) */
    const uint32_t* tris;     /* [n_tris][3] global vertex ids, reversed winding */
    const uint32_t* tri_model;/* [n_tris] */
    const float* umat;        /* [n_models][16] UniformMaterial */
    const uint32_t* ebit;     /* [n_models] */
    const float* ke;          /* [n_models][3] */
    const int32_t* tex_id;    /* [n_models][2] colour, normal; -1 none */
    int n_tex;
    const uint8_t* tex_data;  /* RGBA8 blob */
    const uint64_t* tex_off;  /* [n_tex] */
    const uint32_t* tex_wh;   /* [n_tex][2] */
    /* derived */
    v3 *v0, *e1, *e2;
    uint8_t* skip;            /* S5: two bitwise-equal vertex positions */
    /* bvh */
    int n_nodes;
    struct onode* nodes;
    uint32_t* order;          /* triangle ids in leaf order */
    float srgb[256];
    v3 bbmin, bbmax;
} rco_scene;

struct onode { v3 lo, hi; int left, right; int first, count; };

static void tri_bounds(const rco_scene* s, uint32_t t, v3* lo, v3* hi)
{
    const uint32_t* ix = s->tris + 3 * (size_t)t;
    *lo = V(FLT_MAX, FLT_MAX, FLT_MAX); *hi = V(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int k = 0; k < 3; k++) {
        const float* p = s->verts + 17 * (size_t)ix[k];
        if (p[0] < lo->x) lo->x = p[0]; if (p[0] > hi->x) hi->x = p[0];
        if (p[1] < lo->y) lo->y = p[1]; if (p[1] > hi->y) hi->y = p[1];
        if (p[2] < lo->z) lo->z = p[2]; if (p[2] > hi->z) hi->z = p[2];
    }
}

static float* g_cent; static int g_axis;
static int cmp_cent(const void* a, const void* b)
{
    float ca = g_cent[3 * (size_t)(*(const uint32_t*)a) + g_axis], cb = g_cent[3 * (size_t)(*(const uint32_t*)b) + g_axis];
    if (ca < cb) return -1; if (ca > cb) return 1;
    uint32_t ia = *(const uint32_t*)a, ib = *(const uint32_t*)b;
    return ia < ib ? -1 : (ia > ib);
}

static int build_rec(rco_scene* s, int first, int count, float pad)
{
    int id = s->n_nodes++;
    struct onode* n = &s->nodes[id];
    v3 lo = V(FLT_MAX, FLT_MAX, FLT_MAX), hi = V(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    v3 clo = lo, chi = hi;
    for (int i = 0; i < count; i++) {
        v3 a, b; uint32_t t = s->order[first + i];
        tri_bounds(s, t, &a, &b);
        lo = V(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z));
        hi = V(fmaxf(hi.x, b.x), fmaxf(hi.y, b.y), fmaxf(hi.z, b.z));
        const float* c = g_cent + 3 * (size_t)t;
        clo = V(fminf(clo.x, c[0]), fminf(clo.y, c[1]), fminf(clo.z, c[2]));
        chi = V(fmaxf(chi.x, c[0]), fmaxf(chi.y, c[1]), fmaxf(chi.z, c[2]));
    }
    n->lo = V(lo.x - pad, lo.y - pad, lo.z - pad);
    n->hi = V(hi.x + pad, hi.y + pad, hi.z + pad);
    n->first = first; n->count = count; n->left = n->right = -1;
    if (count <= 4) return id;
    float ex = chi.x - clo.x, ey = chi.y - clo.y, ez = chi.z - clo.z;
    g_axis = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
    qsort(s->order + first, count, sizeof(uint32_t), cmp_cent);
    int half = count / 2;
    n->count = 0;
    int l = build_rec(s, first, half, pad);
    int r = build_rec(s, first + half, count - half, pad);
    s->nodes[id].left = l; s->nodes[id].right = r;
    return id;
}

rco_scene* rco_scene_create(int n_verts, const float* verts, int n_tris, const uint32_t* tris,
                            const uint32_t* tri_model, int n_models, const float* umat,
                            const uint32_t* ebit, const float* ke, const int32_t* tex_id,
                            int n_tex, const uint8_t* tex_data, const uint64_t* tex_off, const uint32_t* tex_wh)
{
    rco_scene* s = (rco_scene*)calloc(1, sizeof(rco_scene));
    s->n_verts = n_verts; s->n_tris = n_tris; s->n_models = n_models;
    s->verts = verts; s->tris = tris; s->tri_model = tri_model; s->umat = umat; s->ebit = ebit; s->ke = ke;
    s->tex_id = tex_id; s->n_tex = n_tex; s->tex_data = tex_data; s->tex_off = tex_off; s->tex_wh = tex_wh;
    s->v0 = (v3*)malloc(sizeof(v3) * (size_t)n_tris);
    s->e1 = (v3*)malloc(sizeof(v3) * (size_t)n_tris);
    s->e2 = (v3*)malloc(sizeof(v3) * (size_t)n_tris);
    g_cent = (float*)malloc(sizeof(float) * 3 * (size_t)n_tris);
    s->order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_tris);
    s->skip = (uint8_t*)calloc((size_t)n_tris + 1, 1);
    s->bbmin = V(FLT_MAX, FLT_MAX, FLT_MAX); s->bbmax = V(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int i = 0; i < n_verts; i++) {
        const float* p = verts + 17 * (size_t)i;
        s->bbmin = V(fminf(s->bbmin.x, p[0]), fminf(s->bbmin.y, p[1]), fminf(s->bbmin.z, p[2]));
        s->bbmax = V(fmaxf(s->bbmax.x, p[0]), fmaxf(s->bbmax.y, p[1]), fmaxf(s->bbmax.z, p[2]));
    }
    for (int t = 0; t < n_tris; t++) {
        const float* a = verts + 17 * (size_t)tris[3 * t + 0];
        const float* b = verts + 17 * (size_t)tris[3 * t + 1];
        const float* c = verts + 17 * (size_t)tris[3 * t + 2];
        s->v0[t] = V(a[0], a[1], a[2]);
        s->e1[t] = V(b[0] - a[0], b[1] - a[1], b[2] - a[2]);   /* S5: e1 = v1 - v0 */
        s->e2[t] = V(c[0] - a[0], c[1] - a[1], c[2] - a[2]);
        s->skip[t] = (memcmp(a, b, 12) == 0) || (memcmp(a, c, 12) == 0) || (memcmp(b, c, 12) == 0);
        v3 lo, hi; tri_bounds(s, (uint32_t)t, &lo, &hi);
        g_cent[3 * t + 0] = 0.5f * (lo.x + hi.x); g_cent[3 * t + 1] = 0.5f * (lo.y + hi.y); g_cent[3 * t + 2] = 0.5f * (lo.z + hi.z);
        s->order[t] = (uint32_t)t;
    }
    v3 d = vsub(s->bbmax, s->bbmin);
    float diag = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    s->nodes = (struct onode*)malloc(sizeof(struct onode) * (size_t)(2 * n_tris + 2));
    s->n_nodes = 0;
    if (n_tris > 0) build_rec(s, 0, n_tris, 1e-4f * diag);
    free(g_cent); g_cent = NULL;
    for (int i = 0; i < 256; i++) {   /* S7: sRGB decode table in double */
        double c = i / 255.0;
        s->srgb[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
    }
    return s;
}

void rco_scene_destroy(rco_scene* s)
{
    if (!s) return;
    free(s->v0); free(s->e1); free(s->e2); free(s->skip); free(s->nodes); free(s->order); free(s);
}

/* S5: two-sided Moller-Trumbore with the specified operation order. */
static inline int ray_tri(const rco_scene* s, uint32_t t, v3 o, v3 d, float tmin, float tmax,
                          float* tt, float* uu, float* vv)
{
    if (s->skip[t]) return 0;
    v3 e1 = s->e1[t], e2 = s->e2[t];
    v3 p = vcross(d, e2);
    float det = vdot(e1, p);
    if (!(det != 0.0f)) return 0;            /* rejects 0 and NaN */
    float inv = 1.0f / det;
    v3 sv = vsub(o, s->v0[t]);
    float u = vdot(sv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return 0;
    v3 q = vcross(sv, e1);
    float v = vdot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return 0;
    float th = vdot(e2, q) * inv;
    if (!(th >= tmin && th < tmax)) return 0;
    *tt = th; *uu = u; *vv = v;
    return 1;
}

typedef struct { float t, u, v; uint32_t prim; } rco_hit;

static inline void consider(const rco_scene* s, uint32_t t, v3 o, v3 d, float tmin, float tmax, rco_hit* h)
{
    float tt, uu, vv;
    /* candidates with t == best t are still evaluated: ties go to the lower id (S5) */
    float lim = (h->prim == 0xffffffffu) ? tmax : nextafterf(h->t, FLT_MAX);
    if (lim > tmax) lim = tmax;
    if (ray_tri(s, t, o, d, tmin, lim, &tt, &uu, &vv)) {
        if (h->prim == 0xffffffffu || tt < h->t || (tt == h->t && t < h->prim)) {
            h->t = tt; h->u = uu; h->v = vv; h->prim = t;
        }
    }
}

static rco_hit trace_brute(const rco_scene* s, v3 o, v3 d, float tmin, float tmax)
{
    rco_hit h = { -1.0f, 0, 0, 0xffffffffu };
    for (int t = 0; t < s->n_tris; t++) consider(s, (uint32_t)t, o, d, tmin, tmax, &h);
    return h;
}

/* conservative slab test: never rejects a box that a S5-accepted hit lies in */
static inline int box_overlap(const struct onode* n, v3 o, v3 d, float tmin, float tmax)
{
    float t0 = tmin, t1 = tmax;
    const float oo[3] = { o.x, o.y, o.z }, dd[3] = { d.x, d.y, d.z };
    const float lo[3] = { n->lo.x, n->lo.y, n->lo.z }, hi[3] = { n->hi.x, n->hi.y, n->hi.z };
    for (int a = 0; a < 3; a++) {
        if (dd[a] == 0.0f) {
            if (oo[a] < lo[a] || oo[a] > hi[a]) return 0;
        } else {
            float inv = 1.0f / dd[a];
            float ta = (lo[a] - oo[a]) * inv, tb = (hi[a] - oo[a]) * inv;
            if (ta > tb) { float x = ta; ta = tb; tb = x; }
            ta -= fabsf(ta) * 4e-7f; tb += fabsf(tb) * 4e-7f;
            if (ta > t0) t0 = ta;
            if (tb < t1) t1 = tb;
            if (t0 > t1) return 0;
        }
    }
    return 1;
}

static rco_hit trace_bvh(const rco_scene* s, v3 o, v3 d, float tmin, float tmax)
{
    rco_hit h = { -1.0f, 0, 0, 0xffffffffu };
    if (s->n_nodes == 0) return h;
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const struct onode* n = &s->nodes[stack[--sp]];
        float lim = (h.prim == 0xffffffffu) ? tmax : h.t;
        /* <= lim: equal-t candidates must still be visited for the id tie-break */
        if (!box_overlap(n, o, d, tmin, nextafterf(lim, FLT_MAX))) continue;
        if (n->left < 0) {
            for (int i = 0; i < n->count; i++) consider(s, s->order[n->first + i], o, d, tmin, tmax, &h);
        } else {
            stack[sp++] = n->left; stack[sp++] = n->right;
        }
    }
    return h;
}

/* rays: [n][8] = o.xyz, tmin, d.xyz, tmax;  hits: [n][4] = t, u, v, prim bits */
void rco_trace(const rco_scene* s, const float* rays, int n, float* hits, int brute)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        const float* r = rays + 8 * (size_t)i;
        v3 o = V(r[0], r[1], r[2]), d = V(r[4], r[5], r[6]);
        rco_hit h = brute ? trace_brute(s, o, d, r[3], r[7]) : trace_bvh(s, o, d, r[3], r[7]);
        float* out = hits + 4 * (size_t)i;
        out[0] = h.t; out[1] = h.u; out[2] = h.v; memcpy(out + 3, &h.prim, 4);
    }
}

/* ---------------------------------------------------------------- shading */
static inline int mirror_idx(long i, long size)   /* Vulkan MirrorRepeat */
{
    long m = i % (2 * size); if (m < 0) m += 2 * size;
    m -= size;
    if (m < 0) m = -(1 + m);
    return (int)((size - 1) - m);
}

static v3 sample_nearest(const rco_scene* s, int tex, float u, float v, int srgb)
{
    if (tex < 0) return V(0, 0, 0);               /* Texture::empty: (0,0,0,0) (src/texture.rs:12-64) */
    long w = s->tex_wh[2 * tex], h = s->tex_wh[2 * tex + 1];
    float fu = floorf(u * (float)w), fv = floorf(v * (float)h);
    if (!(fabsf(fu) < 1e9f)) fu = 0.0f;           /* NaN / huge uv: texel 0 */
    if (!(fabsf(fv) < 1e9f)) fv = 0.0f;
    int ix = mirror_idx((long)fu, w), iy = mirror_idx((long)fv, h);
    const uint8_t* p = s->tex_data + s->tex_off[tex] + 4 * ((size_t)iy * w + ix);
    if (srgb) return V(s->srgb[p[0]], s->srgb[p[1]], s->srgb[p[2]]);
    return V(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f);
}

/* The reference's sampler (src/texture.rs:132-140): MirrorRepeat, mag Linear, min Nearest, one mip level.
 * duv = texture-coordinate differences to the +x / +y neighbour pixels; rho = max(|duv_x * size|, |duv_y * size|);
 * magnification (linear, after the sRGB decode) iff rho <= 1, else — and whenever duv is NULL (a GI ray) — nearest. */
static v3 texel_at(const rco_scene* s, int tex, long x, long y, int srgb)
{
    long w = s->tex_wh[2 * tex], h = s->tex_wh[2 * tex + 1];
    const uint8_t* p = s->tex_data + s->tex_off[tex] + 4 * ((size_t)mirror_idx(y, h) * w + mirror_idx(x, w));
    if (srgb) return V(s->srgb[p[0]], s->srgb[p[1]], s->srgb[p[2]]);
    return V(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f);
}

static v3 sample_tex(const rco_scene* s, int tex, float u, float v, int srgb, const float* duv)
{
    if (tex < 0 || !duv) return sample_nearest(s, tex, u, v, srgb);
    float w = (float)s->tex_wh[2 * tex], h = (float)s->tex_wh[2 * tex + 1];
    float ax = duv[0] * w, ay = duv[1] * h, bx = duv[2] * w, by = duv[3] * h;
    float rho = fmaxf(sqrtf(ax * ax + ay * ay), sqrtf(bx * bx + by * by));
    if (!(rho <= 1.0f)) return sample_nearest(s, tex, u, v, srgb);
    float fu = u * w - 0.5f, fv = v * h - 0.5f;
    if (!(fabsf(fu) < 1e9f)) fu = 0.0f;
    if (!(fabsf(fv) < 1e9f)) fv = 0.0f;
    float iu = floorf(fu), iv = floorf(fv), a = fu - iu, b = fv - iv;
    long x0 = (long)iu, y0 = (long)iv;
    v3 c00 = texel_at(s, tex, x0, y0, srgb), c10 = texel_at(s, tex, x0 + 1, y0, srgb);
    v3 c01 = texel_at(s, tex, x0, y0 + 1, srgb), c11 = texel_at(s, tex, x0 + 1, y0 + 1, srgb);
    float ia = 1.0f - a, ib = 1.0f - b;
    v3 top = V(c00.x * ia + c10.x * a, c00.y * ia + c10.y * a, c00.z * ia + c10.z * a);
    v3 bot = V(c01.x * ia + c11.x * a, c01.y * ia + c11.y * a, c01.z * ia + c11.z * a);
    return V(top.x * ib + bot.x * b, top.y * ib + bot.y * b, top.z * ib + bot.z * b);
}

/* S5 arithmetic without the inside tests: barycentrics of ray (o,d) on the plane of triangle t */
static int plane_bary(const rco_scene* s, uint32_t t, v3 o, v3 d, float* u, float* v)
{
    v3 e1 = s->e1[t], e2 = s->e2[t];
    v3 p = vcross(d, e2);
    float det = vdot(e1, p);
    if (!(det != 0.0f)) return 0;
    float inv = 1.0f / det;
    v3 sv = vsub(o, s->v0[t]);
    *u = vdot(sv, p) * inv;
    *v = vdot(d, vcross(sv, e1)) * inv;
    return 1;
}

typedef struct { int n; float pos[8][3]; uint32_t flags; } rco_lights;

/* S7 attribute interpolation: fma(a2, v, fma(a1, u, a0*((1-u)-v))) */
static inline float lerp3(float a0, float a1, float a2, float u, float v)
{
    float w = (1.0f - u) - v;
    return fmaf(a2, v, fmaf(a1, u, a0 * w));
}

static inline float clamp_rad(float x) { return fminf(fmaxf(x, 0.0f), 65504.0f); } /* NaN -> 0 */

/* fs_main (src/shader.wgsl:76-100) at a hit, view vector Vd = -ray dir.
 * Returns Ke + (L + unlit)*albedo in rad, and the shading normal / albedo / lit colour. */
static void rco_shade_fp(const rco_scene* s, uint32_t prim, float u, float v, v3 P, v3 Vd,
                         const rco_lights* lights, v3* rad, v3* nshade, v3* albedo_out, v3* direct_out, const float* fp)
{
    const uint32_t* ix = s->tris + 3 * (size_t)prim;
    const float *a = s->verts + 17 * (size_t)ix[0], *b = s->verts + 17 * (size_t)ix[1], *c = s->verts + 17 * (size_t)ix[2];
    float at[17];
    for (int k = 3; k < 17; k++) at[k] = lerp3(a[k], b[k], c[k], u, v);
    uint32_t m = s->tri_model[prim];
    const float* um = s->umat + 16 * (size_t)m;
    uint32_t eb = s->ebit[m];
    if (!(lights->flags & 1u)) eb &= 1u;          /* normal-map toggle (src/renderer.rs:623) */
    int b0 = eb & 1, b1 = (eb >> 1) & 1;
    float tu = at[15], tv = 1.0f - at[16];                                  /* :78 */
    float duv_store[4];
    const float* duv = NULL;
    if (fp) {   /* texture footprint of a primary ray: neighbours' texcoords minus this pixel's */
        duv_store[0] = lerp3(a[15], b[15], c[15], fp[0], fp[1]) - tu;
        duv_store[1] = (1.0f - lerp3(a[16], b[16], c[16], fp[0], fp[1])) - tv;
        duv_store[2] = lerp3(a[15], b[15], c[15], fp[2], fp[3]) - tu;
        duv_store[3] = (1.0f - lerp3(a[16], b[16], c[16], fp[2], fp[3])) - tv;
        duv = duv_store;
    }
    v3 color = V(at[3], at[4], at[5]);
    v3 albedo = b0 ? sample_tex(s, s->tex_id[2 * m], tu, tv, 1, duv) : color; /* :80 */
    v3 L = V(um[0] * 0.05f * um[3], um[1] * 0.05f * um[3], um[2] * 0.05f * um[3]);   /* :82-83 */
    v3 Nv = V(at[6], at[7], at[8]);
    v3 raw;
    if (b1) {
        v3 cs = sample_tex(s, s->tex_id[2 * m + 1], tu, tv, 0, duv);
        v3 cf = V(cs.x * 2.0f - 1.0f, cs.y * 2.0f - 1.0f, cs.z * 2.0f - 1.0f);      /* :85 */
        v3 T = vnormalize(V(at[9], at[10], at[11])), B = vnormalize(V(at[12], at[13], at[14]));
        v3 mix = vadd(vadd(vscale(T, cf.x), vscale(B, cf.y)), vscale(Nv, cf.z));  /* :86, Nv not normalised */
        raw = vnormalize(mix);
    } else {
        raw = vnormalize(Nv);
    }
    float ndv = vdot(Vd, raw);                                              /* :88 */
    v3 N = ndv < 0.0f ? vneg(raw) : raw;                                    /* :89 */
    for (int li = 0; li < lights->n; li++) {
        v3 lp = V(lights->pos[li][0], lights->pos[li][1], lights->pos[li][2]);
        v3 Ld = vnormalize(vsub(lp, P));                                    /* :91 */
        float ndl = fmaxf(vdot(Ld, N), 0.0f);                               /* :92 */
        float kd = 0.7f * ndl * um[7];
        L = V(fmaf(um[4], kd, L.x), fmaf(um[5], kd, L.y), fmaf(um[6], kd, L.z));   /* :93 */
        v3 Hd = vnormalize(vadd(Vd, Ld));                                   /* :95 */
        float st = powf(fmaxf(vdot(N, Hd), 0.0f), um[12]);                  /* :96 */
        float ks = st * um[11] * (ndv > 1e-6f ? 1.0f : 0.0f);               /* :97 */
        L = V(fmaf(um[8], ks, L.x), fmaf(um[9], ks, L.y), fmaf(um[10], ks, L.z));
    }
    float pred = ((um[0] - 1e-5f) + (um[4] - 1e-5f) + (um[8] - 1e-5f))
               + ((um[1] - 1e-5f) + (um[5] - 1e-5f) + (um[9] - 1e-5f))
               + ((um[2] - 1e-5f) + (um[6] - 1e-5f) + (um[10] - 1e-5f));     /* :99 */
    float unlit = pred <= 0.0f ? 1.0f : 0.0f;
    v3 lit = V((L.x + unlit) * albedo.x, (L.y + unlit) * albedo.y, (L.z + unlit) * albedo.z);  /* :100 */
    const float* ke = s->ke + 3 * (size_t)m;
    *rad = V(clamp_rad(ke[0] + lit.x), clamp_rad(ke[1] + lit.y), clamp_rad(ke[2] + lit.z));
    if (nshade) *nshade = N;
    if (albedo_out) *albedo_out = albedo;
    if (direct_out) *direct_out = lit;
}

static void rco_shade(const rco_scene* s, uint32_t prim, float u, float v, v3 P, v3 Vd,
                      const rco_lights* lights, v3* rad, v3* nshade, v3* albedo_out, v3* direct_out)
{
    rco_shade_fp(s, prim, u, v, P, Vd, lights, rad, nshade, albedo_out, direct_out, NULL);   /* GI hit: nearest filter (S7) */
}

/* in: [n][8] = prim bits, u, v, pad, view origin xyz, pad -> out [n][4] rgb of Ke + fs_main, a = 0 */
void rco_shade_points(const rco_scene* s, const float* in, int n, const float* light_pos, int n_lights,
                      uint32_t flags, float* out)
{
    rco_lights L; L.n = n_lights; L.flags = flags;
    for (int i = 0; i < n_lights; i++) memcpy(L.pos[i], light_pos + 4 * i, 12);
    for (int i = 0; i < n; i++) {
        const float* r = in + 8 * (size_t)i;
        uint32_t prim; memcpy(&prim, r, 4);
        float u = r[1], v = r[2];
        v3 eye = V(r[4], r[5], r[6]);
        const uint32_t* ix = s->tris + 3 * (size_t)prim;
        const float *a = s->verts + 17 * (size_t)ix[0], *b = s->verts + 17 * (size_t)ix[1], *c = s->verts + 17 * (size_t)ix[2];
        v3 P = V(lerp3(a[0], b[0], c[0], u, v), lerp3(a[1], b[1], c[1], u, v), lerp3(a[2], b[2], c[2], u, v));
        v3 Vd = vnormalize(vsub(eye, P));
        v3 rad;
        rco_shade(s, prim, u, v, P, Vd, &L, &rad, NULL, NULL, NULL);
        out[4 * i] = rad.x; out[4 * i + 1] = rad.y; out[4 * i + 2] = rad.z; out[4 * i + 3] = 0.0f;
    }
}

/* ---------------------------------------------------------------- frame */
typedef struct {
    int W, H, P0, D0, N;
    float L0, t_far, offset;
    float sky[3];
    int tile_x0, tile_y0, tile_w, tile_h;   /* pixels produced by gather/gbuffer */
    int store_half;                          /* round stored cascade texels / outputs through float16 */
    int clip;                                /* S4b: primary rays limited to the near / far planes of view_proj */
    int floating;                            /* S6: probes with an empty anchor float to a finer-level anchor (optional) */
} rco_params;

typedef struct {
    int P, D, gw, gh;
    float t0, t1;
} rco_level;

/* S1 + S2 */
void rco_level_layout(const rco_params* p, int i, rco_level* L)
{
    L->P = p->P0 << i; L->D = p->D0 << i;
    L->gw = (p->W + L->P - 1) / L->P; L->gh = (p->H + L->P - 1) / L->P;
    double a = (pow(4.0, i) - 1.0) / 3.0, b = (pow(4.0, i + 1) - 1.0) / 3.0;
    L->t0 = (float)((double)p->L0 * a);
    L->t1 = (i == p->N - 1) ? p->t_far : (float)((double)p->L0 * b);
}

/* S3: equal-area octahedral directions, double -> float; out [D*D][3] */
void rco_directions(int D, float* out)
{
    const double PI = 3.14159265358979323846;
    for (int dy = 0; dy < D; dy++) for (int dx = 0; dx < D; dx++) {
        double u = (2.0 * dx + 1.0) / D - 1.0, v = (2.0 * dy + 1.0) / D - 1.0;
        double d = 1.0 - (fabs(u) + fabs(v)), r = 1.0 - fabs(d);
        double phi = (r == 0.0) ? 0.0 : (PI / 4.0) * ((fabs(v) - fabs(u)) / r + 1.0);
        double f = r * sqrt(2.0 - r * r);
        float* o = out + 3 * ((size_t)dy * D + dx);
        o[0] = (float)copysign(f * cos(phi), u);
        o[1] = (float)copysign(f * sin(phi), v);
        o[2] = (float)copysign(1.0 - r * r, d);
    }
}

/* S4: primary-ray basis from the 80-byte camera uniform.  out: Dx, Dy, Dc (9 floats) */
static int invert4(const double m[16], double inv[16])
{
    double a[4][8];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { a[r][c] = m[c * 4 + r]; a[r][c + 4] = (r == c); }
    for (int c = 0; c < 4; c++) {
        int piv = c; for (int r = c + 1; r < 4; r++) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0.0) return 0;
        if (piv != c) for (int k = 0; k < 8; k++) { double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
        double d = a[c][c]; for (int k = 0; k < 8; k++) a[c][k] /= d;
        for (int r = 0; r < 4; r++) if (r != c) { double f = a[r][c]; for (int k = 0; k < 8; k++) a[r][k] -= f * a[c][k]; }
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) inv[c * 4 + r] = a[r][c + 4];
    return 1;
}

int rco_primary_basis(const float cam[20], float out9[9])
{
    double m[16], inv[16];
    for (int i = 0; i < 16; i++) m[i] = cam[i];
    if (!invert4(m, inv)) return 0;
    double e[3] = { cam[16], cam[17], cam[18] };
    double A[4], B[4], C[4];
    for (int r = 0; r < 4; r++) { A[r] = inv[0 * 4 + r]; B[r] = inv[1 * 4 + r]; C[r] = inv[2 * 4 + r] + inv[3 * 4 + r]; }
    double dx[3], dy[3], dc[3];
    for (int k = 0; k < 3; k++) { dx[k] = A[k] - e[k] * A[3]; dy[k] = B[k] - e[k] * B[3]; dc[k] = C[k] - e[k] * C[3]; }
    double sc = (C[3] < 0 ? -1.0 : 1.0) / sqrt(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]);
    for (int k = 0; k < 3; k++) { out9[k] = (float)(dx[k] * sc); out9[3 + k] = (float)(dy[k] * sc); out9[6 + k] = (float)(dc[k] * sc); }
    return 1;
}

static inline v3 primary_dir(const rco_params* p, const float b[9], int x, int y)
{
    float nx = (float)(2 * x + 1) / (float)p->W - 1.0f;
    float ny = 1.0f - (float)(2 * y + 1) / (float)p->H;
    v3 q = V(fmaf(nx, b[0], fmaf(ny, b[3], b[6])), fmaf(nx, b[1], fmaf(ny, b[4], b[7])), fmaf(nx, b[2], fmaf(ny, b[5], b[8])));
    return vnormalize(q);
}

/* standard (non-equal-area) octahedral encode of the shading normal, 2 x snorm16 */
static uint32_t oct_encode(v3 n)
{
    float s = fabsf(n.x) + fabsf(n.y) + fabsf(n.z);
    float px = n.x / s, py = n.y / s;
    if (n.z < 0.0f) {
        float qx = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
        float qy = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
        px = qx; py = qy;
    }
    int ix = (int)lrintf(fminf(fmaxf(px, -1.0f), 1.0f) * 32767.0f);
    int iy = (int)lrintf(fminf(fmaxf(py, -1.0f), 1.0f) * 32767.0f);
    return ((uint32_t)(uint16_t)(int16_t)ix) | (((uint32_t)(uint16_t)(int16_t)iy) << 16);
}

static v3 oct_decode(uint32_t e)
{
    float px = (float)(int16_t)(e & 0xffffu) / 32767.0f, py = (float)(int16_t)(e >> 16) / 32767.0f;
    float z = 1.0f - fabsf(px) - fabsf(py);
    float x = px, y = py;
    if (z < 0.0f) {
        x = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
        y = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
    }
    return vnormalize(V(x, y, z));
}

/* G-buffer over the tile (S4).  depth[th][tw], prim[th][tw], normal[th][tw] (oct u32),
 * albedo[th][tw][4], direct[th][tw][4] (float32; rounded through half if store_half). */
/* S4b: range of a primary ray between the near and far planes of the column-major view_proj in cam[0..15]
 * (src/camera.rs:77-79 perspective_rh, depth 0..1): z_clip(t) = z0 + t*zd >= 0 and z_clip(t) <= w_clip(t) = w0 + t*wd */
static inline void primary_range(const rco_params* p, const float cam[20], v3 eye, v3 d, float* tmin, float* tmax)
{
    *tmin = 0.0f; *tmax = FLT_MAX;
    if (!p->clip) return;
    const float z0 = fmaf(cam[10], eye.z, fmaf(cam[6], eye.y, fmaf(cam[2], eye.x, cam[14])));
    const float w0 = fmaf(cam[11], eye.z, fmaf(cam[7], eye.y, fmaf(cam[3], eye.x, cam[15])));
    const float zd = vdot(V(cam[2], cam[6], cam[10]), d), wd = vdot(V(cam[3], cam[7], cam[11]), d);
    if (zd > 0.0f) *tmin = fmaxf(-z0 / zd, 0.0f);
    const float g = zd - wd;
    if (g > 0.0f) *tmax = (w0 - z0) / g;
}

void rco_gbuffer(const rco_scene* s, const rco_params* p, const float cam[20], const float* light_pos, int n_lights,
                 uint32_t flags, float* depth, uint32_t* prim, uint32_t* normal, float* albedo, float* direct)
{
    float b[9]; rco_primary_basis(cam, b);
    v3 eye = V(cam[16], cam[17], cam[18]);
    rco_lights L; L.n = n_lights; L.flags = flags;
    for (int i = 0; i < n_lights; i++) memcpy(L.pos[i], light_pos + 4 * i, 12);
#pragma omp parallel for schedule(dynamic, 4)
    for (int ty = 0; ty < p->tile_h; ty++) for (int tx = 0; tx < p->tile_w; tx++) {
        int x = p->tile_x0 + tx, y = p->tile_y0 + ty;
        size_t o = (size_t)ty * p->tile_w + tx;
        v3 d = primary_dir(p, b, x, y);
        float tmin, tmax;
        primary_range(p, cam, eye, d, &tmin, &tmax);
        rco_hit h = trace_bvh(s, eye, d, tmin, tmax);
        if (h.prim == 0xffffffffu) {
            depth[o] = -1.0f; prim[o] = 0xffffffffu; normal[o] = 0;
            for (int k = 0; k < 4; k++) { albedo[4 * o + k] = 0; direct[4 * o + k] = 0; }
            continue;
        }
        v3 P = vfma(h.t, d, eye), rad, ns, al, di;
        float fp[4];
        const float* fpp = NULL;
        if (s->ebit[s->tri_model[h.prim]] != 0u &&
            plane_bary(s, h.prim, eye, primary_dir(p, b, x + 1, y), &fp[0], &fp[1]) &&
            plane_bary(s, h.prim, eye, primary_dir(p, b, x, y + 1), &fp[2], &fp[3])) fpp = fp;
        rco_shade_fp(s, h.prim, h.u, h.v, P, vneg(d), &L, &rad, &ns, &al, &di, fpp);
        depth[o] = h.t; prim[o] = h.prim; normal[o] = oct_encode(ns);
        float av[3] = { al.x, al.y, al.z }, dv[3] = { di.x, di.y, di.z };
        for (int k = 0; k < 3; k++) {
            float A = clamp_rad(av[k]), D = clamp_rad(dv[k]);
            albedo[4 * o + k] = p->store_half ? half_round(A) : A;
            direct[4 * o + k] = p->store_half ? half_round(D) : D;
        }
        albedo[4 * o + 3] = 1.0f; direct[4 * o + 3] = 1.0f;
    }
}

/* S6: probes of level i over the sub-grid [px0,px0+sw) x [py0,py0+sh).
 * origin[n][4] = xyz, valid;  nrm[n][4] = ng xyz, 0 */
void rco_probes(const rco_scene* s, const rco_params* p, const float cam[20], int level,
                int px0, int py0, int sw, int sh, float* origin, float* nrm)
{
    rco_level L; rco_level_layout(p, level, &L);
    rco_level FL[16];
    for (int l = 0; l <= level && l < 16; l++) rco_level_layout(p, l, &FL[l]);
    float b[9]; rco_primary_basis(cam, b);
    v3 eye = V(cam[16], cam[17], cam[18]);
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < sh; j++) for (int i = 0; i < sw; i++) {
        int px = px0 + i, py = py0 + j;
        size_t o = (size_t)j * sw + i;
        /* S6: the probe's own anchor first; when it sees no geometry the probe FLOATS to the first anchor of the finer
         * levels' probes inside its cell that does (level by level downwards, row-major within a level) */
        rco_hit h; h.prim = 0xffffffffu; h.t = -1.0f; h.u = h.v = 0.0f;
        v3 d = V(0, 0, 0);
        for (int l = level; l >= (p->floating ? 0 : level) && h.prim == 0xffffffffu; l--) {
            const rco_level F = FL[l];
            const int sc = 1 << (level - l);
            for (int qy = py * sc; qy < (py + 1) * sc && qy < F.gh && h.prim == 0xffffffffu; qy++)
                for (int qx = px * sc; qx < (px + 1) * sc && qx < F.gw; qx++) {
                    int ax = qx * F.P + F.P / 2, ay = qy * F.P + F.P / 2;
                    if (ax > p->W - 1) ax = p->W - 1; if (ay > p->H - 1) ay = p->H - 1;
                    d = primary_dir(p, b, ax, ay);
                    float tmin, tmax;
                    primary_range(p, cam, eye, d, &tmin, &tmax);
                    h = trace_bvh(s, eye, d, tmin, tmax);
                    if (h.prim != 0xffffffffu) break;
                }
        }
        if (h.prim == 0xffffffffu) {
            for (int k = 0; k < 4; k++) { origin[4 * o + k] = 0; nrm[4 * o + k] = 0; }
            continue;
        }
        v3 hp = vfma(h.t, d, eye);
        v3 ng = vnormalize(vcross(s->e1[h.prim], s->e2[h.prim]));
        if (vdot(ng, d) > 0.0f) ng = vneg(ng);
        v3 og = vfma(p->offset, ng, hp);
        origin[4 * o] = og.x; origin[4 * o + 1] = og.y; origin[4 * o + 2] = og.z; origin[4 * o + 3] = 1.0f;
        nrm[4 * o] = ng.x; nrm[4 * o + 1] = ng.y; nrm[4 * o + 2] = ng.z; nrm[4 * o + 3] = 0.0f;
    }
}

/* S7: raw interval radiance of level i; out [sh*sw*D*D][4] float32, probe-major (S1).
 * hit_t (optional) [texels]: hit distance or -1; hit_prim (optional). */
void rco_march(const rco_scene* s, const rco_params* p, int level, int sw, int sh,
               const float* origin, const float* dirs, const float* light_pos, int n_lights, uint32_t flags,
               float* out, float* hit_t, uint32_t* hit_prim)
{
    rco_level L; rco_level_layout(p, level, &L);
    rco_lights Ls; Ls.n = n_lights; Ls.flags = flags;
    for (int i = 0; i < n_lights; i++) memcpy(Ls.pos[i], light_pos + 4 * i, 12);
    int DD = L.D * L.D;
    long n = (long)sw * sh;
    int top = (level == p->N - 1);
#pragma omp parallel for schedule(dynamic, 1)
    for (long pr = 0; pr < n; pr++) {
        const float* og = origin + 4 * pr;
        float* o4 = out + 4 * (size_t)pr * DD;
        for (int d = 0; d < DD; d++) {
            float* t4 = o4 + 4 * d;
            size_t ti = (size_t)pr * DD + d;
            if (og[3] == 0.0f) {
                t4[0] = t4[1] = t4[2] = 0.0f; t4[3] = 1.0f;
                if (hit_t) hit_t[ti] = -1.0f; if (hit_prim) hit_prim[ti] = 0xffffffffu;
                continue;
            }
            v3 o = V(og[0], og[1], og[2]), w = V(dirs[3 * d], dirs[3 * d + 1], dirs[3 * d + 2]);
            rco_hit h = trace_bvh(s, o, w, L.t0, L.t1);
            if (hit_t) hit_t[ti] = h.t; if (hit_prim) hit_prim[ti] = h.prim;
            if (h.prim == 0xffffffffu) {
                t4[0] = top ? p->sky[0] : 0.0f; t4[1] = top ? p->sky[1] : 0.0f; t4[2] = top ? p->sky[2] : 0.0f; t4[3] = 1.0f;
            } else {
                v3 P = vfma(h.t, w, o), rad;
                rco_shade(s, h.prim, h.u, h.v, P, vneg(w), &Ls, &rad, NULL, NULL, NULL);
                t4[0] = rad.x; t4[1] = rad.y; t4[2] = rad.z; t4[3] = 0.0f;
            }
            if (p->store_half) for (int k = 0; k < 4; k++) t4[k] = half_round(t4[k]);
        }
    }
}

/* S1: the two upper probes and weights of probe index q along one axis */
static inline void upper_pair(int q, int gmax, int* i0, int* i1, float* w0, float* w1)
{
    int base;
    if ((q & 1) == 0) { base = q / 2 - 1; *w0 = 0.25f; *w1 = 0.75f; }
    else { base = (q - 1) / 2; *w0 = 0.75f; *w1 = 0.25f; }
    int a = base, b = base + 1;
    if (a < 0) a = 0; if (a > gmax - 1) a = gmax - 1;
    if (b < 0) b = 0; if (b > gmax - 1) b = gmax - 1;
    *i0 = a; *i1 = b;
}

/* S8 plane-distance weight */
static inline float plane_weight(v3 np_, v3 op, const float* ok)
{
    if (ok[3] == 0.0f) return 0.0f;
    v3 delta = vsub(V(ok[0], ok[1], ok[2]), op);
    float l2 = vdot(delta, delta), h = vdot(np_, delta);
    return l2 > 0.0f ? 1.0f / (1.0f + 16.0f * (h * h) / l2) : 1.0f;
}

/* S8: merge level i (lo, in place) with merged level i+1 (up).  Sub-grids:
 * lo covers probes [lpx0, lpx0+lsw) x [lpy0, ...), up covers [upx0, ...). Upper
 * probes outside the up sub-grid must not be referenced (caller sizes the halo). */
void rco_merge(const rco_params* p, int level,
               int lpx0, int lpy0, int lsw, int lsh, const float* lo_origin, const float* lo_nrm, float* lo,
               int upx0, int upy0, int usw, int ush, const float* up_origin, const float* up)
{
    rco_level L, U; rco_level_layout(p, level, &L); rco_level_layout(p, level + 1, &U);
    int D = L.D, UD = U.D;
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < lsh; j++) for (int i = 0; i < lsw; i++) {
        size_t pr = (size_t)j * lsw + i;
        const float* og = lo_origin + 4 * pr;
        float* base = lo + 4 * pr * D * D;
        if (og[3] == 0.0f) continue;                 /* invalid probe stays (0,0,0,1) */
        int x0, x1, y0, y1; float wx0, wx1, wy0, wy1;
        upper_pair(lpx0 + i, U.gw, &x0, &x1, &wx0, &wx1);
        upper_pair(lpy0 + j, U.gh, &y0, &y1, &wy0, &wy1);
        int ux[4] = { x0, x1, x0, x1 }, uy[4] = { y0, y0, y1, y1 };
        float bw[4] = { wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1 };
        v3 op = V(og[0], og[1], og[2]), np_ = V(lo_nrm[4 * pr], lo_nrm[4 * pr + 1], lo_nrm[4 * pr + 2]);
        float w[4]; size_t upr[4];
        for (int k = 0; k < 4; k++) {
            upr[k] = (size_t)(uy[k] - upy0) * usw + (ux[k] - upx0);
            w[k] = bw[k] * plane_weight(np_, op, up_origin + 4 * upr[k]);
        }
        float S = ((w[0] + w[1]) + w[2]) + w[3];
        for (int dy = 0; dy < D; dy++) for (int dx = 0; dx < D; dx++) {
            float* t4 = base + 4 * ((size_t)dy * D + dx);
            float far[4] = { p->sky[0], p->sky[1], p->sky[2], 0.0f };   /* S8: no valid upper probe -> the sky */
            if (S > 0.0f) {
                far[3] = 0.0f;
                for (int k = 0; k < 4; k++) {
                    float wk = w[k] / S;
                    const float* ub = up + 4 * upr[k] * UD * UD;
                    const float* c0 = ub + 4 * ((size_t)(2 * dy) * UD + 2 * dx);
                    const float* c1 = c0 + 4;
                    const float* c2 = ub + 4 * ((size_t)(2 * dy + 1) * UD + 2 * dx);
                    const float* c3 = c2 + 4;
                    for (int ch = 0; ch < 4; ch++) {
                        float avg = 0.25f * (((c0[ch] + c1[ch]) + c2[ch]) + c3[ch]);
                        far[ch] = fmaf(wk, avg, far[ch]);
                    }
                }
            }
            float a = t4[3];
            t4[0] = fmaf(a, far[0], t4[0]); t4[1] = fmaf(a, far[1], t4[1]); t4[2] = fmaf(a, far[2], t4[2]);
            t4[3] = a * far[3];
            if (p->store_half) for (int k = 0; k < 4; k++) t4[k] = half_round(fminf(t4[k], 65504.0f));
        }
    }
}

/* S9: gather over the tile.  c0 = merged level 0 over sub-grid [px0..) x [py0..);
 * depth/normal = G-buffer of the tile.  out [th][tw][4] */
void rco_gather(const rco_params* p, const float cam[20],
                int px0, int py0, int sw, int sh, const float* origin0, const float* c0, const float* dirs0,
                const float* depth, const uint32_t* normal, float* out)
{
    rco_level L; rco_level_layout(p, 0, &L);
    float b[9]; rco_primary_basis(cam, b);
    v3 eye = V(cam[16], cam[17], cam[18]);
    int DD = L.D * L.D;
#pragma omp parallel for schedule(dynamic, 4)
    for (int ty = 0; ty < p->tile_h; ty++) for (int tx = 0; tx < p->tile_w; tx++) {
        int x = p->tile_x0 + tx, y = p->tile_y0 + ty;
        size_t o = (size_t)ty * p->tile_w + tx;
        float* e = out + 4 * o;
        if (depth[o] < 0.0f) { e[0] = e[1] = e[2] = e[3] = 0.0f; continue; }
        v3 d = primary_dir(p, b, x, y);
        v3 hp = vfma(depth[o], d, eye);
        v3 n = oct_decode(normal[o]);
        int idx[2][2]; float wt[2][2];
        int coord[2] = { x, y }, gmax[2] = { L.gw, L.gh };
        for (int a = 0; a < 2; a++) {
            int s_ = coord[a] - L.P / 2;
            int base = (s_ >= 0) ? s_ / L.P : -((-s_ + L.P - 1) / L.P);
            float f = (float)(s_ - base * L.P) / (float)L.P;
            int i0 = base, i1 = base + 1;
            if (i0 < 0) i0 = 0; if (i0 > gmax[a] - 1) i0 = gmax[a] - 1;
            if (i1 < 0) i1 = 0; if (i1 > gmax[a] - 1) i1 = gmax[a] - 1;
            idx[a][0] = i0; idx[a][1] = i1; wt[a][0] = 1.0f - f; wt[a][1] = f;
        }
        int ux[4] = { idx[0][0], idx[0][1], idx[0][0], idx[0][1] }, uy[4] = { idx[1][0], idx[1][0], idx[1][1], idx[1][1] };
        float bw[4] = { wt[0][0] * wt[1][0], wt[0][1] * wt[1][0], wt[0][0] * wt[1][1], wt[0][1] * wt[1][1] };
        float w[4]; size_t pr[4];
        for (int k = 0; k < 4; k++) {
            pr[k] = (size_t)(uy[k] - py0) * sw + (ux[k] - px0);
            w[k] = bw[k] * plane_weight(n, hp, origin0 + 4 * pr[k]);
        }
        float S = ((w[0] + w[1]) + w[2]) + w[3];
        float E[3] = { 0, 0, 0 };
        /* S9: normalised cosine quadrature, q = pi / sum_d max(n.w_d, 0) */
        float csum = 0.0f;
        for (int di = 0; di < DD; di++)
            csum = csum + fmaxf(vdot(n, V(dirs0[3 * di], dirs0[3 * di + 1], dirs0[3 * di + 2])), 0.0f);
        const float dw = csum > 0.0f ? 3.14159274101257324219f / csum : 0.0f;
        if (S > 0.0f) {
            for (int k = 0; k < 4; k++) {
                float wk = w[k] / S;
                float acc[3] = { 0, 0, 0 };
                const float* cb = c0 + 4 * pr[k] * DD;
                for (int di = 0; di < DD; di++) {
                    float cs = fmaxf(vdot(n, V(dirs0[3 * di], dirs0[3 * di + 1], dirs0[3 * di + 2])), 0.0f);
                    acc[0] = fmaf(cs, cb[4 * di], acc[0]); acc[1] = fmaf(cs, cb[4 * di + 1], acc[1]); acc[2] = fmaf(cs, cb[4 * di + 2], acc[2]);
                }
                float wd = wk * dw;
                E[0] = fmaf(wd, acc[0], E[0]); E[1] = fmaf(wd, acc[1], E[1]); E[2] = fmaf(wd, acc[2], E[2]);
            }
        }
        for (int k = 0; k < 3; k++) e[k] = p->store_half ? half_round(fminf(E[k], 65504.0f)) : E[k];
        e[3] = 1.0f;
    }
}

/* Direction culling (kernels.cu k_gbuffer), restated for the tests: per pixel the mask of level-0 directions the gather
 * above weights with cs_d = max(dot(n, w_d), 0) > 0 — bit d set iff dot(n, w_d) > 0 with the decoded stored normal. */
void rco_pixel_masks(const rco_params* p, const float* dirs0, const float* depth, const uint32_t* normal, uint32_t* mask)
{
    rco_level L; rco_level_layout(p, 0, &L);
    int DD = L.D * L.D;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < p->tile_w * p->tile_h; i++) {
        uint32_t m = 0u;
        if (!(depth[i] < 0.0f) && DD <= 32) {
            v3 n = oct_decode(normal[i]);
            for (int di = 0; di < DD; di++)
                if (vdot(n, V(dirs0[3 * di], dirs0[3 * di + 1], dirs0[3 * di + 2])) > 0.0f) m |= 1u << di;
        }
        mask[i] = m;
    }
}

/* ---- Independent restatement of the reference's render pass as a RASTERISER (src/renderer.rs:332-360, 565-593;
 * src/shader.wgsl:30-43): every triangle is transformed by the column-major view_proj (vs_main), clipped against the
 * near (z_clip >= 0) and far (z_clip <= w_clip) planes, projected, and scan-converted at pixel centres with
 * perspective-correct barycentrics; depth test Less against a buffer cleared to 1.0, cull_mode None, triangles in draw
 * order (model by model, index order) so that the first of two equal depths wins.  It shares no code with the ray caster
 * above: tests/test_gpu_parity.py compares the product's clipped G-buffer (S4b) with it — same triangle per pixel except
 * on edge pixels, same barycentrics and depth within float tolerance.
 * prim [th][tw] (0xffffffff = clear colour), zndc [th][tw] (1.0 = clear), bary [th][tw][2] = (u, v) of S5. */
typedef struct { double x, y, z, w, b[3]; } rvert;

static int clip_poly(const rvert* in, int n, rvert* out, int plane /* 0: z >= 0, 1: w - z >= 0 */)
{
    int m = 0;
    for (int i = 0; i < n; i++) {
        const rvert *a = &in[i], *c = &in[(i + 1) % n];
        const double da = plane ? a->w - a->z : a->z, dc = plane ? c->w - c->z : c->z;
        if (da >= 0.0) out[m++] = *a;
        if ((da >= 0.0) != (dc >= 0.0)) {
            const double t = da / (da - dc);
            rvert v;
            v.x = a->x + t * (c->x - a->x); v.y = a->y + t * (c->y - a->y); v.z = a->z + t * (c->z - a->z); v.w = a->w + t * (c->w - a->w);
            for (int k = 0; k < 3; k++) v.b[k] = a->b[k] + t * (c->b[k] - a->b[k]);
            out[m++] = v;
        }
    }
    return m;
}

void rco_raster(const rco_scene* s, const rco_params* p, const float cam[20], uint32_t* prim, float* zndc, float* bary)
{
    const int tw = p->tile_w, th = p->tile_h;
    for (size_t i = 0; i < (size_t)tw * th; i++) { prim[i] = 0xffffffffu; zndc[i] = 1.0f; bary[2 * i] = bary[2 * i + 1] = 0.0f; }
    const int band = 16, nb = (th + band - 1) / band;
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < nb; bi++) {
        const int ya = p->tile_y0 + bi * band, yb = (ya + band < p->tile_y0 + th) ? ya + band : p->tile_y0 + th;
        for (int t = 0; t < s->n_tris; t++) {
            if (s->skip[t]) continue;
            rvert poly[8], tmp[8];
            for (int k = 0; k < 3; k++) {
                const float* v = s->verts + 17 * (size_t)s->tris[3 * (size_t)t + k];
                rvert r;
                r.x = (double)cam[0] * v[0] + (double)cam[4] * v[1] + (double)cam[8] * v[2] + cam[12];
                r.y = (double)cam[1] * v[0] + (double)cam[5] * v[1] + (double)cam[9] * v[2] + cam[13];
                r.z = (double)cam[2] * v[0] + (double)cam[6] * v[1] + (double)cam[10] * v[2] + cam[14];
                r.w = (double)cam[3] * v[0] + (double)cam[7] * v[1] + (double)cam[11] * v[2] + cam[15];
                r.b[0] = r.b[1] = r.b[2] = 0.0; r.b[k] = 1.0;
                poly[k] = r;
            }
            int n = clip_poly(poly, 3, tmp, 0);
            if (n < 3) continue;
            n = clip_poly(tmp, n, poly, 1);
            if (n < 3) continue;
            double sx[8], sy[8], sz[8], iw[8];
            for (int k = 0; k < n; k++) {
                if (!(poly[k].w > 0.0)) { n = 0; break; }
                iw[k] = 1.0 / poly[k].w;
                sx[k] = (poly[k].x * iw[k] * 0.5 + 0.5) * p->W;
                sy[k] = (1.0 - (poly[k].y * iw[k] * 0.5 + 0.5)) * p->H;
                sz[k] = poly[k].z * iw[k];
            }
            for (int f = 1; f + 1 < n; f++) {
                const int id[3] = { 0, f, f + 1 };
                const double x0 = sx[id[0]], y0 = sy[id[0]], x1 = sx[id[1]], y1 = sy[id[1]], x2 = sx[id[2]], y2 = sy[id[2]];
                double area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
                if (area == 0.0 || area != area) continue;
                const double sg = area < 0.0 ? -1.0 : 1.0;
                area *= sg;
                double fx0 = fmin(x0, fmin(x1, x2)), fx1 = fmax(x0, fmax(x1, x2)), fy0 = fmin(y0, fmin(y1, y2)), fy1 = fmax(y0, fmax(y1, y2));
                int ix0 = (int)floor(fx0 - 0.5), ix1 = (int)ceil(fx1 - 0.5), iy0 = (int)floor(fy0 - 0.5), iy1 = (int)ceil(fy1 - 0.5);
                if (ix0 < p->tile_x0) ix0 = p->tile_x0;
                if (ix1 > p->tile_x0 + tw - 1) ix1 = p->tile_x0 + tw - 1;
                if (iy0 < ya) iy0 = ya;
                if (iy1 > yb - 1) iy1 = yb - 1;
                for (int y = iy0; y <= iy1; y++) for (int x = ix0; x <= ix1; x++) {
                    const double cx = x + 0.5, cy = y + 0.5;
                    const double e0 = sg * ((x2 - x1) * (cy - y1) - (y2 - y1) * (cx - x1));
                    const double e1 = sg * ((x0 - x2) * (cy - y2) - (y0 - y2) * (cx - x2));
                    const double e2 = sg * ((x1 - x0) * (cy - y0) - (y1 - y0) * (cx - x0));
                    if (e0 < 0.0 || e1 < 0.0 || e2 < 0.0) continue;
                    const double l0 = e0 / area, l1 = e1 / area, l2 = e2 / area;
                    const float z = (float)(l0 * sz[id[0]] + l1 * sz[id[1]] + l2 * sz[id[2]]);
                    const size_t o = (size_t)(y - p->tile_y0) * tw + (x - p->tile_x0);
                    if (!(z >= 0.0f && z <= 1.0f) || !(z < zndc[o])) continue;      /* viewport depth range, depth test Less */
                    const double q0 = l0 * iw[id[0]], q1 = l1 * iw[id[1]], q2 = l2 * iw[id[2]], qs = q0 + q1 + q2;
                    double bb[3];
                    for (int k = 0; k < 3; k++) bb[k] = (q0 * poly[id[0]].b[k] + q1 * poly[id[1]].b[k] + q2 * poly[id[2]].b[k]) / qs;
                    zndc[o] = z; prim[o] = (uint32_t)t;
                    bary[2 * o] = (float)bb[1]; bary[2 * o + 1] = (float)bb[2];
                }
            }
        }
    }
}

int rco_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the OpenMP loops (the timing arm of bench.py uses every host core even when a launcher such as
 * torchrun exported OMP_NUM_THREADS=1 for its worker processes). */
void rco_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void rco_scene_bbox(const rco_scene* s, float out6[6])
{
    out6[0] = s->bbmin.x; out6[1] = s->bbmin.y; out6[2] = s->bbmin.z;
    out6[3] = s->bbmax.x; out6[4] = s->bbmax.y; out6[5] = s->bbmax.z;
}
