// kernels.cu — hand-written CUDA kernels of the radiance-cascade GI path for sm_100a.
//
//   k_gbuffer    primary visibility + shading normal per pixel (+ per-pixel direction masks)   (rc_spec.h S4, src/shader.wgsl:76-100)
//   k_direct     deferred fs_main (albedo / direct colour), on demand                           (src/shader.wgsl:76-100)
//   k_probes     probe placement for all cascade levels (+ level-0 request masks)              (S6)
//   k_link_entry per-probe upper-probe slots and bilateral weights (S1, S8) + optional per-probe BVH entry frontiers
//   k_need       direction culling: request masks pushed up the cascade, one ray list per level
//   k_split      split ray lists: entries whose rays all miss the BVH root's two child boxes go to the back of a copy
//                of the level's list; k_march skips their traversal
//   k_march      per-level interval ray march fused with the merge from level i+1 and with the child
//                averages the level below will read (S7, S8); list-driven in culled frames
//   k_merge / k_child_avg / k_fill_top / k_march_all / k_march_persist / k_march_compact / k_march_quad / k_march_pool /
//   k_need_chain / k_bin + k_gbuffer_binned   A/B variants, measured-and-rejected schedules and the every-texel path
//                (rc_set_tuning; all bit-identical to the default)
//   k_gather_mma final irradiance gather on mma.sync with TMA-staged probes (+ peer-memory stores of the tile in tiled
//                multi-GPU mode) (S9);  k_gather / k_gather_pipe: the scalar forms in S9's exact summation order
//   k_peer_*     flag handshake of the peer-memory frame exchange
//
// Data layout (all in HBM, sized at rc_create): cascade levels are probe-major RGBA16F texels (8 B): one warp
// marches directions of ONE probe (shared origin -> coherent BVH traversal); the merge reads, per lower texel,
// 4 upper probes x one 16-byte child average.  The one dense contraction of the path is the gather's 16x16x8 product
// per 4x4-pixel cell (mma.sync; far too small for tcgen05 / TMEM); the BVH + triangles (< 4 MB) live in L2 (126 MB)
// and are read through the read-only path.
#include "kernels.cuh"

namespace rc {

namespace {

constexpr int kBlock = 256;

// Plain (coherent, L1-allocating) 128-bit load: the merged upper level is written by the previous
// kernel of the PDL chain while this kernel may already be running, so the non-coherent .nc path is not used for it.
__device__ __forceinline__ uint4 ld_u4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float4 ld_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ------------------------------------------------------------------ G-buffer
// Per frame the GI path needs primary visibility only: distance, triangle id, barycentrics and the two-sided
// shading normal.  fs_main's colour outputs (the reference's whole frame, src/shader.wgsl:76-100) are not an input
// of any GI stage, so they are evaluated on demand by k_direct (deferred shading) when a DIRECT / ALBEDO /
// COMPOSITE target is read.
__device__ __forceinline__ bool pixel_footprint(const DScene& s, const DCamera& cam, uint32_t prim, int x, int y, float* fp)
{
    const float* v0p = s.verts + 17 * (size_t)s.tris[3 * (size_t)prim];
    const float3 v0 = f3(__ldg(v0p), __ldg(v0p + 1), __ldg(v0p + 2));
    const float3 e1 = xyz(__ldg(s.tri_eg + 2 * (size_t)prim)), e2 = xyz(__ldg(s.tri_eg + 2 * (size_t)prim + 1));
    const float3 dx = primary_dir(cam, x + 1, y), dy = primary_dir(cam, x, y + 1);
    return plane_bary(v0, e1, e2, cam.eye, dx, fp[0], fp[1]) && plane_bary(v0, e1, e2, cam.eye, dy, fp[2], fp[3]);
}

// S1: the two level-0 probes (clamped) and bilinear weights of pixel coordinate `coord` along one axis
__device__ __forceinline__ void gather_axis(int coord, int P, int gmax, int& i0, int& i1, float& w0, float& w1)
{
    const int s = coord - P / 2;
    int base, rem;
    float f;
    if ((P & (P - 1)) == 0) {           // power-of-two spacing: arithmetic shift == floor division
        const int lp = 31 - __clz(P);
        base = s >> lp;
        rem = s & (P - 1);
        f = (float)rem * __int_as_float((127 - lp) << 23);   // rem * 2^-lp == rem / P exactly: no IEEE division
    } else {
        base = (s >= 0) ? s / P : -((-s + P - 1) / P);
        rem = s - base * P;
        f = (float)rem / (float)P;
    }
    i0 = min(max(base, 0), gmax - 1);
    i1 = min(max(base + 1, 0), gmax - 1);
    w0 = 1.0f - f; w1 = f;
}

// Direction culling, level 0 (`pixmask` != null).  The gather (S9) weights texel (probe, d) by
// cs_d = max(dot(n, w_d), 0) of the pixels that interpolate the probe: a level-0 direction that is at or below
// the horizon of EVERY such pixel is multiplied by zero and need not be traced.  Each covered pixel evaluates
// the very expression the gather will evaluate (decoded normal, same dot product) and stores its mask of
// directions with cs_d > 0 (D0^2 <= 16 bits); k_probes ORs the masks of the pixels each level-0 probe serves
// and k_need carries the result up the cascade.
// everything k_gbuffer stores for one pixel once its closest hit is known (shared by the BVH and the binned variant)
__device__ __forceinline__ void gbuffer_store(const DScene& s, const DCamera& cam, const DLights& L, const TileRect& tile, const GBufferOut& out,
                                              int DD0, const float* s_dirs, uint16_t* __restrict__ pixmask, int tx, int ty, float3 d, const Hit& h)
{
    const size_t o = (size_t)ty * tile.w + tx;
    if (out.occ) {   // (floating probes only) a warp covers 8x4 pixels, i.e. lies inside one 32x8 occupancy cell: one stamp per warp that saw geometry
        const unsigned act = __activemask();
        const unsigned hits = __ballot_sync(act, h.prim != 0xffffffffu);
        if (hits && (threadIdx.x & 31u) == (unsigned)(__ffs((int)hits) - 1)) out.occ[(size_t)(ty >> 3) * out.ow + (tx >> 5)] = out.frame;
    }
    if (h.prim == 0xffffffffu) {
        out.depth[o] = -1.0f; out.prim[o] = 0xffffffffu; out.normal[o] = 0u; out.bary[o] = make_float2(0.f, 0.f);
        if (pixmask) pixmask[o] = 0;
        return;
    }
    const float3 P = vfma(h.t, d, cam.eye);
    // a normal-mapped material needs the texture footprint for the sampler's mag/min decision (S4)
    float fp[4];
    const float* fpp = nullptr;
    if ((s.mats[s.tri_model[h.prim]].ebit & 2u) && (L.flags & 1u) && pixel_footprint(s, cam, h.prim, tile.x0 + tx, tile.y0 + ty, fp)) fpp = fp;
    const Shade sh = shade_hit<true>(s, L, h.prim, h.u, h.v, P, vneg(d), fpp);
    const uint32_t enc = oct_encode(sh.n);
    out.depth[o] = h.t;
    out.prim[o] = h.prim;
    out.normal[o] = enc;
    out.bary[o] = make_float2(h.u, h.v);
    if (pixmask) {
        // S9: n = decoded stored normal, cs_d = max(dot(n, w_d), 0) > 0  <=>  dot(n, w_d) > 0
        const float3 n = oct_decode(enc);
        uint32_t m = 0u;
        if (DD0 == 16) {
#pragma unroll
            for (int di = 0; di < 16; di++)
                if (vdot(n, f3(s_dirs[3 * di], s_dirs[3 * di + 1], s_dirs[3 * di + 2])) > 0.0f) m |= 1u << di;
        } else {
            for (int di = 0; di < DD0; di++)
                if (vdot(n, f3(s_dirs[3 * di], s_dirs[3 * di + 1], s_dirs[3 * di + 2])) > 0.0f) m |= 1u << di;
        }
        pixmask[o] = (uint16_t)m;
    }
}

#ifndef RC_GBUF_MINB
#define RC_GBUF_MINB 4   // resident 256-thread blocks per SM the register allocation must allow (A/B: make EXTRA=-DRC_GBUF_MINB=n)
#endif
__global__ void __launch_bounds__(kBlock, RC_GBUF_MINB) k_gbuffer(DScene s, DCamera cam, DLights L, TileRect tile, GBufferOut out,
                                                    int DD0, const float* __restrict__ dirs0, uint16_t* __restrict__ pixmask)
{
    __shared__ float s_dirs[3 * 16];
    if (pixmask) {
        if (threadIdx.x < 3 * DD0) s_dirs[threadIdx.x] = dirs0[threadIdx.x];
        __syncthreads();
    }
    // a block covers 32x8 pixels; each warp an 8x4 pixel tile (compact frustum -> coherent traversal; four
    // 32-byte row segments per store)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int ty = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (tx >= tile.w || ty >= tile.h) return;
    const float3 d = primary_dir(cam, tile.x0 + tx, tile.y0 + ty);
    float tmin, tmax;
    primary_range(cam, d, tmin, tmax);
    const Hit h = trace(s, cam.eye, d, tmin, tmax);
    gbuffer_store(s, cam, L, tile, out, DD0, s_dirs, pixmask, tx, ty, d, h);
}

// ------------------------------------------------------------------ primary visibility by triangle binning
// Primary rays share their origin, so "which triangles can a pixel's ray hit" is a 2-D question: the triangles whose
// projection covers the pixel centre.  k_bin projects every triangle with the frame's view_proj (vs_main,
// src/shader.wgsl:30-43), and appends it to the candidate list of every 16x16-pixel tile its projection (grown by one
// pixel) touches; k_gbuffer_binned then runs S5 — the same tri_test arithmetic as the BVH path — for every pixel of a
// tile over the tile's staged candidates and keeps min (t, id).  Closest hits are defined over ALL triangles (S5), and
// a triangle S5 accepts for a pixel projects onto that pixel's centre to within float rounding (sub-pixel), so the
// conservative candidate set gives bit-identical hits.  Per pixel this costs ~10 cheaply rejected candidates instead of
// ~11 two-box node visits + 5.5 triangle tests (ncu r1: 1410 thread-instructions per pixel, 69 % of them traversal).
// Lists have a fixed capacity (kBinCap per tile, no count / scan / fill passes): a tile with more candidates — distant,
// finely tessellated geometry — traces its pixels through the BVH instead, where the BVH is the better structure anyway.
constexpr int kBinTile = 16, kBinCap = 96;

struct BinTri {            // projection of one triangle: screen-space bounds (pixel-centre units) and edge functions
    int x0, x1, y0, y1;    // inclusive pixel range, already grown by one pixel and clamped to the tile rect; x1 < x0: nothing
    float ea[3], eb[3], ec[3];   // edge k: ea*x + eb*y + ec >= 0 inside (valid when `edges`)
    int edges;
};

__device__ __forceinline__ BinTri bin_project(const DCamera& cam, const TileRect& tile, float4 a, float4 e1, float4 e2)
{
    BinTri r;
    r.x0 = r.y0 = 0; r.x1 = r.y1 = -1; r.edges = 0;
    const float3 v[3] = {xyz(a), vadd(xyz(a), xyz(e1)), vadd(xyz(a), xyz(e2))};
    float cx[3], cy[3], cw[3];
    float wmax = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        cx[k] = fmaf(cam.row_x.z, v[k].z, fmaf(cam.row_x.y, v[k].y, fmaf(cam.row_x.x, v[k].x, cam.row_x.w)));
        cy[k] = fmaf(cam.row_y.z, v[k].z, fmaf(cam.row_y.y, v[k].y, fmaf(cam.row_y.x, v[k].x, cam.row_y.w)));
        cw[k] = fmaf(cam.row_w.z, v[k].z, fmaf(cam.row_w.y, v[k].y, fmaf(cam.row_w.x, v[k].x, cam.row_w.w)));
        wmax = fmaxf(wmax, fabsf(cw[k]));
    }
    // points in front of the eye plane: w >= weps.  A ray only ever hits points with w > 0; the slab between 0 and weps
    // is covered by treating every triangle that reaches into it as touching the whole screen.
    const float weps = 1e-4f * wmax + 1e-30f;
    const bool in[3] = {cw[0] > weps, cw[1] > weps, cw[2] > weps};
    if (!(in[0] || in[1] || in[2])) {
        if (cw[0] > 0.0f || cw[1] > 0.0f || cw[2] > 0.0f || !(wmax == wmax)) { r.x0 = 0; r.y0 = 0; r.x1 = tile.w - 1; r.y1 = tile.h - 1; }
        return r;
    }
    const float W = (float)cam.W, H = (float)cam.H;
    float sx[3], sy[3];
    float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
    bool clipped = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (in[k]) {
            const float iw = 1.0f / cw[k];
            sx[k] = (cx[k] * iw * 0.5f + 0.5f) * W - 0.5f;            // pixel-centre units: pixel px has its centre at px
            sy[k] = (1.0f - (cy[k] * iw * 0.5f + 0.5f)) * H - 0.5f;
            xmin = fminf(xmin, sx[k]); xmax = fmaxf(xmax, sx[k]); ymin = fminf(ymin, sy[k]); ymax = fmaxf(ymax, sy[k]);
        }
        // an edge that crosses the plane w = weps contributes its crossing point (the part in front of the plane is the convex
        // hull of the kept vertices and the crossing points, so their bounding box bounds its projection)
        const int k1 = (k + 1) % 3;
        if (in[k] != in[k1]) {
            clipped = true;
            const float tt = (weps - cw[k]) / (cw[k1] - cw[k]);
            const float px = fmaf(tt, cx[k1] - cx[k], cx[k]), py = fmaf(tt, cy[k1] - cy[k], cy[k]);
            const float iw = 1.0f / weps;
            const float qx = (px * iw * 0.5f + 0.5f) * W - 0.5f, qy = (1.0f - (py * iw * 0.5f + 0.5f)) * H - 0.5f;
            xmin = fminf(xmin, qx); xmax = fmaxf(xmax, qx); ymin = fminf(ymin, qy); ymax = fmaxf(ymax, qy);
        }
    }
    if (!(xmin == xmin) || !(xmax == xmax) || !(ymin == ymin) || !(ymax == ymax)) {     // NaN anywhere: the whole tile rect
        r.x0 = 0; r.y0 = 0; r.x1 = tile.w - 1; r.y1 = tile.h - 1;
        return r;
    }
    // clamp before the int conversion (projections near the eye plane reach 1e30 pixels); grown by 4 pixels when clipped
    // (the crossing points are rounded at a much larger scale than the kept vertices)
    const float grow = clipped ? 4.0f : 0.0f;
    xmin = fminf(fmaxf(xmin - grow, -1.0e6f), 1.0e6f); xmax = fminf(fmaxf(xmax + grow, -1.0e6f), 1.0e6f);
    ymin = fminf(fmaxf(ymin - grow, -1.0e6f), 1.0e6f); ymax = fminf(fmaxf(ymax + grow, -1.0e6f), 1.0e6f);
    const int x0 = (int)floorf(xmin) - 1 - tile.x0, x1 = (int)ceilf(xmax) + 1 - tile.x0;
    const int y0 = (int)floorf(ymin) - 1 - tile.y0, y1 = (int)ceilf(ymax) + 1 - tile.y0;
    if (x1 < 0 || y1 < 0 || x0 >= tile.w || y0 >= tile.h) return r;
    r.x0 = max(x0, 0); r.y0 = max(y0, 0); r.x1 = min(x1, tile.w - 1); r.y1 = min(y1, tile.h - 1);
    // edge functions in tile-relative pixel coordinates, oriented so that the inside is >= 0
    const float area = clipped ? 0.0f : (sx[1] - sx[0]) * (sy[2] - sy[0]) - (sx[2] - sx[0]) * (sy[1] - sy[0]);
    if (fabsf(area) > 1e-3f && fabsf(xmin) < 3.0e4f && fabsf(xmax) < 3.0e4f && fabsf(ymin) < 3.0e4f && fabsf(ymax) < 3.0e4f) {
        const float sg = area < 0.0f ? -1.0f : 1.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int i0 = (k + 1) % 3, i1 = (k + 2) % 3;
            const float A = sg * (sy[i0] - sy[i1]), B = sg * (sx[i1] - sx[i0]);
            // E(x, y) = A*(x - sx[i0]) + B*(y - sy[i0]) with (x, y) in frame pixels; shift to tile-relative pixels
            r.ea[k] = A; r.eb[k] = B;
            r.ec[k] = -(A * (sx[i0] - (float)tile.x0) + B * (sy[i0] - (float)tile.y0));
        }
        r.edges = 1;
    }
    return r;
}

// does the projected triangle (grown by ~1.5 pixels) reach the 16x16 tile (bx, by)?  Conservative: false only when the
// whole tile lies outside one edge
__device__ __forceinline__ bool bin_touches(const BinTri& t, int bx, int by)
{
    if (!t.edges) return true;
    const float xa = (float)(bx * kBinTile) - 1.5f, xb = (float)(bx * kBinTile + kBinTile - 1) + 1.5f;
    const float ya = (float)(by * kBinTile) - 1.5f, yb = (float)(by * kBinTile + kBinTile - 1) + 1.5f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float A = t.ea[k], B = t.eb[k];
        const float emax = A * (A >= 0.0f ? xb : xa) + B * (B >= 0.0f ? yb : ya) + t.ec[k];
        // slack for the rounding of the edge function itself (coordinates up to ~1e4 pixels)
        if (emax < -1e-3f * (fabsf(A) + fabsf(B)) * 16.0f - 1e-4f * fabsf(t.ec[k])) return false;
    }
    return true;
}

__device__ __forceinline__ void bin_append(unsigned int* __restrict__ count, uint32_t* __restrict__ lists, int tile_id, uint32_t tri)
{
    const unsigned slot = atomicAdd(count + tile_id, 1u);
    if (slot < (unsigned)kBinCap) lists[(size_t)tile_id * kBinCap + slot] = tri;
}

// A triangle that touches more than kBinHuge tiles (a wall seen from inside the room covers thousands) is not binned at
// all: it goes to the frame's "huge" list (its tile range + edge functions, 64 bytes), and every tile's block tests the
// whole list against its own tile — 256 threads, a few hundred entries, three edge evaluations each.  Walking such ranges
// in k_bin serialised a handful of warps for a millisecond (first version: living_room 4K 1.09 ms, test_room 0.54 ms).
constexpr int kBinHuge = 512, kBinBlock = 128;
struct HugeTri { int bx0, bx1, by0, by1; float ea[3], eb[3], ec[3]; uint32_t tri; int edges; int pad; };
static_assert(sizeof(HugeTri) == 64, "HugeTri is fetched as four 16-byte vectors");

// One thread per leaf-ordered triangle projects it; the (triangle, tile) pairs of a block's triangles are then flattened
// (block-wide prefix sum of the per-triangle tile counts) and dealt out to the threads pair by pair, so a triangle with 400
// tiles and one with 1 tile cost their block the same per pair, and the appends (one returning atomic each) of different
// threads are in flight together.  (First version: per-thread / per-warp loops — 122 us at 4K, 6 % of the warp slots busy.)
__global__ void __launch_bounds__(kBinBlock) k_bin(DScene s, DCamera cam, TileRect tile, uint32_t n_tris, int ntx, unsigned int* __restrict__ count,
                                                   uint32_t* __restrict__ lists, unsigned int* __restrict__ huge_count, HugeTri* __restrict__ huge)
{
    __shared__ BinTri s_t[kBinBlock];
    __shared__ int s_pre[kBinBlock + 1];
    __shared__ int s_wsum[kBinBlock / 32];
    const uint32_t i = blockIdx.x * kBinBlock + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    BinTri t;
    t.x0 = t.y0 = 0; t.x1 = t.y1 = -1; t.edges = 0;
    if (i < n_tris) t = bin_project(cam, tile, __ldg(s.tri_geom + 3 * (size_t)i), __ldg(s.tri_geom + 3 * (size_t)i + 1), __ldg(s.tri_geom + 3 * (size_t)i + 2));
    const bool any = t.x1 >= t.x0 && t.y1 >= t.y0;
    // from here on the ranges are in TILES
    t.x0 /= kBinTile; t.x1 /= kBinTile; t.y0 /= kBinTile; t.y1 /= kBinTile;
    int ntl = any ? (t.x1 - t.x0 + 1) * (t.y1 - t.y0 + 1) : 0;
    if (ntl > kBinHuge) {
        HugeTri h;
        h.bx0 = t.x0; h.bx1 = t.x1; h.by0 = t.y0; h.by1 = t.y1; h.tri = i; h.edges = t.edges; h.pad = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) { h.ea[k] = t.ea[k]; h.eb[k] = t.eb[k]; h.ec[k] = t.ec[k]; }
        huge[atomicAdd(huge_count, 1u)] = h;
        ntl = 0;
    }
    s_t[threadIdx.x] = t;
    // block-wide exclusive prefix sum of ntl
    int pre = ntl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, o); if ((int)lane >= o) pre += v; }
    if (lane == 31) s_wsum[wid] = pre;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int k = 0; k < kBinBlock / 32; k++) if (k < (int)wid) base += s_wsum[k];
    s_pre[threadIdx.x] = base + pre - ntl;
    if (threadIdx.x == kBinBlock - 1) s_pre[kBinBlock] = base + pre;
    __syncthreads();
    const int total = s_pre[kBinBlock];
    for (int pidx = threadIdx.x; pidx < total; pidx += kBinBlock) {
        int lo = 0, hi = kBinBlock - 1;          // the last m with s_pre[m] <= pidx (entries with ntl = 0 repeat a prefix: take the last)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_pre[mid] <= pidx) lo = mid; else hi = mid - 1;
        }
        const BinTri& b = s_t[lo];
        const int k = pidx - s_pre[lo], wx = b.x1 - b.x0 + 1;
        const int bx = b.x0 + k % wx, by = b.y0 + k / wx;
        if (bin_touches(b, bx, by)) bin_append(count, lists, by * ntx + bx, blockIdx.x * kBinBlock + (uint32_t)lo);
    }
}

// one block per 16x16-pixel tile; warps cover 8x4 pixels each.  While a candidate is staged its projection is evaluated once
// more against the eight 8x4 sub-tiles: a warp only runs S5 for the candidates that can reach its own pixels.
__global__ void __launch_bounds__(kBlock) k_gbuffer_binned(DScene s, DCamera cam, DLights L, TileRect tile, GBufferOut out, int DD0,
                                                           const float* __restrict__ dirs0, uint16_t* __restrict__ pixmask, int ntx,
                                                           unsigned int* __restrict__ count, const uint32_t* __restrict__ lists,
                                                           const unsigned int* __restrict__ huge_count, const HugeTri* __restrict__ huge)
{
    __shared__ float s_dirs[3 * 16];
    __shared__ float4 s_tri[3 * kBinCap];
    __shared__ uint32_t s_idx[kBinCap];
    __shared__ uint32_t s_mask[kBinCap];
    __shared__ unsigned s_n;
    const int tile_id = blockIdx.y * ntx + blockIdx.x;
    if (threadIdx.x == 0) { s_n = count[tile_id]; count[tile_id] = 0u; }      // consumed: the next frame's k_bin starts from empty lists
    if (pixmask && threadIdx.x < 3 * DD0) s_dirs[threadIdx.x] = dirs0[threadIdx.x];
    __syncthreads();
    const unsigned n_small = s_n;
    __syncthreads();
    if (n_small <= (unsigned)kBinCap) {
        for (unsigned k = threadIdx.x; k < n_small; k += kBlock) s_idx[k] = __ldg(lists + (size_t)tile_id * kBinCap + k);
        // the frame's huge triangles, each tested against this tile
        const unsigned nh = __ldg(huge_count);
        const uint4* hv = reinterpret_cast<const uint4*>(huge);
        for (unsigned k = threadIdx.x; k < nh; k += kBlock) {
            const uint4 q0 = __ldg(hv + 4 * (size_t)k);
            if ((int)blockIdx.x < (int)q0.x || (int)blockIdx.x > (int)q0.y || (int)blockIdx.y < (int)q0.z || (int)blockIdx.y > (int)q0.w) continue;
            const uint4 q1 = __ldg(hv + 4 * (size_t)k + 1), q2 = __ldg(hv + 4 * (size_t)k + 2), q3 = __ldg(hv + 4 * (size_t)k + 3);
            BinTri b;
            b.ea[0] = __uint_as_float(q1.x); b.ea[1] = __uint_as_float(q1.y); b.ea[2] = __uint_as_float(q1.z);
            b.eb[0] = __uint_as_float(q1.w); b.eb[1] = __uint_as_float(q2.x); b.eb[2] = __uint_as_float(q2.y);
            b.ec[0] = __uint_as_float(q2.z); b.ec[1] = __uint_as_float(q2.w); b.ec[2] = __uint_as_float(q3.x);
            b.edges = (int)q3.z;
            if (!bin_touches(b, (int)blockIdx.x, (int)blockIdx.y)) continue;
            const unsigned slot = atomicAdd(&s_n, 1u);
            if (slot < (unsigned)kBinCap) s_idx[slot] = q3.y;
        }
    }
    __syncthreads();
    const unsigned n = s_n;
    const bool binned = n <= (unsigned)kBinCap;
    if (binned) {
        for (unsigned k = threadIdx.x; k < n; k += kBlock) {
            const uint32_t tri = s_idx[k];
            const float4 a = __ldg(s.tri_geom + 3 * (size_t)tri), e1 = __ldg(s.tri_geom + 3 * (size_t)tri + 1), e2 = __ldg(s.tri_geom + 3 * (size_t)tri + 2);
            s_tri[3 * k] = a; s_tri[3 * k + 1] = e1; s_tri[3 * k + 2] = e2;
            // which of the tile's eight 8x4-pixel sub-tiles (= warps) the projection, grown by 1.5 pixels, can reach
            const BinTri b = bin_project(cam, tile, a, e1, e2);
            uint32_t m = 0xffu;
            if (b.edges) {
                m = 0u;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const float xa = (float)(blockIdx.x * kBinTile + (w & 1) * 8) - 1.5f, xb = xa + 7.0f + 3.0f;
                    const float ya = (float)(blockIdx.y * kBinTile + (w >> 1) * 4) - 1.5f, yb = ya + 3.0f + 3.0f;
                    bool in = true;
#pragma unroll
                    for (int e = 0; e < 3; e++) {
                        const float A = b.ea[e], B = b.eb[e];
                        const float emax = A * (A >= 0.0f ? xb : xa) + B * (B >= 0.0f ? yb : ya) + b.ec[e];
                        if (emax < -1e-3f * (fabsf(A) + fabsf(B)) * 16.0f - 1e-4f * fabsf(b.ec[e])) in = false;
                    }
                    if (in) m |= 1u << w;
                }
            }
            s_mask[k] = m;
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x * kBinTile + (warp & 1) * 8 + (lane & 7);
    const int ty = blockIdx.y * kBinTile + (warp >> 1) * 4 + (lane >> 3);
    if (tx >= tile.w || ty >= tile.h) return;
    const float3 d = primary_dir(cam, tile.x0 + tx, tile.y0 + ty);
    float tmin, tmax;
    primary_range(cam, d, tmin, tmax);
    Hit h;
    if (binned) {
        h.t = tmax; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
        for (unsigned k = 0; k < n; k++)
            if ((s_mask[k] >> warp) & 1u) tri_test_v(s_tri[3 * k], s_tri[3 * k + 1], s_tri[3 * k + 2], cam.eye, d, tmin, tmax, h);
        if (h.prim == 0xffffffffu) h.t = -1.0f;
    } else {
        h = trace(s, cam.eye, d, tmin, tmax);
    }
    gbuffer_store(s, cam, L, tile, out, DD0, s_dirs, pixmask, tx, ty, d, h);
}

// fs_main for every covered pixel from the stored visibility (on demand; not part of the per-frame GI path)
__global__ void __launch_bounds__(kBlock) k_direct(DScene s, DCamera cam, DLights L, TileRect tile, const float* __restrict__ depth,
                                                   const uint32_t* __restrict__ prim, const float2* __restrict__ bary,
                                                   uint2* __restrict__ albedo, uint2* __restrict__ direct)
{
    const int tx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (tx >= tile.w || ty >= tile.h) return;
    const size_t o = (size_t)ty * tile.w + tx;
    const uint32_t id = prim[o];
    if (id == 0xffffffffu) { albedo[o] = make_uint2(0u, 0u); direct[o] = make_uint2(0u, 0u); return; }
    const float3 d = primary_dir(cam, tile.x0 + tx, tile.y0 + ty);
    const float3 P = vfma(depth[o], d, cam.eye);
    const float2 uv = bary[o];
    float fp[4];
    const float* fpp = nullptr;
    if (s.mats[s.tri_model[id]].ebit != 0u && pixel_footprint(s, cam, id, tile.x0 + tx, tile.y0 + ty, fp)) fpp = fp;
    const Shade sh = shade_hit<false>(s, L, id, uv.x, uv.y, P, vneg(d), fpp);
    albedo[o] = pack_half4(clamp_rad(sh.albedo.x), clamp_rad(sh.albedo.y), clamp_rad(sh.albedo.z), 1.0f);
    direct[o] = pack_half4(clamp_rad(sh.direct.x), clamp_rad(sh.direct.y), clamp_rad(sh.direct.z), 1.0f);
}

// ------------------------------------------------------------------ probes (S6), all levels in one launch
// A probe whose anchor pixel lies inside the tile reuses the G-buffer's primary hit (same ray,
// same S4/S5 arithmetic -> bit-identical); only halo probes of a multi-GPU tile trace their own ray.
__device__ __forceinline__ int level_of(const DLevelSet& ls, unsigned i)
{
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < ls.n; k++) if (i >= ls.lv[k].probe_offset) l = k;
    return l;
}

// S6 shortcut: true when the whole cell of probe (px, py) lies inside the tile and none of the 32x8-pixel occupancy cells it
// touches was stamped in this frame — then no candidate anchor can hit.  The flags are dealt out to `n` cooperating lanes
// (k-th lane takes flags k, k+n, ...); the caller combines the lanes' answers with a vote.  A cell that leaves the tile is
// never "empty" here: its candidates are traced.
__device__ __forceinline__ bool cell_empty(const DLevel& lv, const DCamera& cam, const TileRect& tile, int px, int py,
                                           const uint32_t* __restrict__ occ, uint32_t frame, int ow, int k, int n)
{
    {   // a cell that misses the screen-space rectangle of the scene's bounding box is empty wherever it lies (halo probes of a
        // multi-GPU strip would otherwise TRACE every candidate anchor of every sky cell: probes stage 0.065 -> 0.13 ms at 8 ranks)
        const int fx0 = px * lv.P, fy0 = py * lv.P, fx1 = min((px + 1) * lv.P, cam.W) - 1, fy1 = min((py + 1) * lv.P, cam.H) - 1;
        if (fx1 < cam.sb_x0 || fx0 > cam.sb_x1 || fy1 < cam.sb_y0 || fy0 > cam.sb_y1) return true;
    }
    const int x0 = px * lv.P - tile.x0, y0 = py * lv.P - tile.y0;
    const int x1 = min((px + 1) * lv.P, cam.W) - 1 - tile.x0, y1 = min((py + 1) * lv.P, cam.H) - 1 - tile.y0;
    if (x0 < 0 || y0 < 0 || x1 >= tile.w || y1 >= tile.h) return false;
    const int ox0 = x0 >> 5, oy0 = y0 >> 3, nx = (x1 >> 5) - ox0 + 1, ny = (y1 >> 3) - oy0 + 1;
    for (int i = k; i < nx * ny; i += n)
        if (occ[(size_t)(oy0 + i / nx) * ow + ox0 + i % nx] == frame) return false;
    return true;
}

// the traced branch of anchor_hit, out of line: k_probes inlines anchor_hit three times, and three copies of the BVH traversal
// cost it 64 registers (42 % occupancy) for a branch only halo probes of a multi-GPU tile ever take
__device__ __noinline__ Hit trace_primary(const DScene& s, const DCamera& cam, float3 d)
{
    float tmin, tmax;
    primary_range(cam, d, tmin, tmax);
    return trace(s, cam.eye, d, tmin, tmax);
}

// primary hit through the anchor pixel of probe (qx, qy) of level fl: the G-buffer's when the anchor lies in the tile (same
// ray, same S4/S5 arithmetic -> bit-identical), traced otherwise (halo probes of a multi-GPU tile)
template <bool OUTLINE>
__device__ __forceinline__ bool anchor_hit(const DScene& s, const DCamera& cam, const DLevel& fl, const TileRect& tile, int qx, int qy,
                                           const float* __restrict__ depth, const uint32_t* __restrict__ prim, float3& d, float& t, uint32_t& id)
{
    const int ax = min(qx * fl.P + fl.P / 2, cam.W - 1), ay = min(qy * fl.P + fl.P / 2, cam.H - 1);
    const int tx = ax - tile.x0, ty = ay - tile.y0;
    if (tx >= 0 && tx < tile.w && ty >= 0 && ty < tile.h) {
        const size_t o = (size_t)ty * tile.w + tx;
        id = prim[o];
        if (id == 0xffffffffu) return false;
        t = depth[o];
        d = primary_dir(cam, ax, ay);
        return true;
    }
    d = primary_dir(cam, ax, ay);
    Hit h;
    if (OUTLINE) {
        h = trace_primary(s, cam, d);
    } else {
        float tmin, tmax;
        primary_range(cam, d, tmin, tmax);
        h = trace(s, cam.eye, d, tmin, tmax);
    }
    t = h.t;
    id = h.prim;
    return id != 0xffffffffu;
}

// S6.  Probes of the levels below `warp_level` take one thread each (their cells hold at most 1 + 4 + 16 candidate anchors);
// from `warp_level` up a probe takes a warp, which checks 32 candidates of a floating probe at a time — an empty cell of the
// top level has 1365 of them, and one thread walking them would outlast the whole frame.
template <bool FLOATING>      // compiled twice: without the S6 candidate search k_probes is the lean round-1 kernel (45 registers, no stack)
__global__ void __launch_bounds__(kBlock) k_probes(DScene s, DCamera cam, DLevelSet ls, unsigned total, unsigned thread_probes, TileRect tile,
                                                   float offset, const float* __restrict__ depth, const uint32_t* __restrict__ prim,
                                                   float4* __restrict__ origin, float4* __restrict__ normal,
                                                   const uint16_t* __restrict__ pixmask, uint32_t* __restrict__ need0,
                                                   const uint32_t* __restrict__ occ, uint32_t frame, int ow)
{
    constexpr bool floating = FLOATING;
    const unsigned g = blockIdx.x * kBlock + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const bool warp_mode = FLOATING && g >= thread_probes;          // thread_probes is a multiple of 32: whole warps are in one mode
    const unsigned gi = warp_mode ? thread_probes + ((g - thread_probes) >> 5) : g;
    if (gi >= total) return;
    const int level = level_of(ls, gi);
    const DLevel& lv = ls.lv[level];
    const int i = (int)(gi - lv.probe_offset);

    const int px = lv.px0 + i % lv.sw, py = lv.py0 + i / lv.sw;
    if (pixmask && level == 0) {
        // direction culling: the pixels whose gather (S1/S9) can touch this probe are x in [(px-1)P + P/2, (px+1)P + P/2)
        // (clamped indices at the frame border stay inside that range), likewise y; only the tile's pixels exist here
        const int xa = max((px - 1) * lv.P + lv.P / 2, tile.x0), xb = min((px + 1) * lv.P + lv.P / 2, tile.x0 + tile.w);
        const int ya = max((py - 1) * lv.P + lv.P / 2, tile.y0), yb = min((py + 1) * lv.P + lv.P / 2, tile.y0 + tile.h);
        uint32_t m = 0u;
        for (int y = ya; y < yb; y++) {
            const uint16_t* row = pixmask + (size_t)(y - tile.y0) * tile.w + (xa - tile.x0);
            int x = 0;
            const int n = xb - xa;
            if ((reinterpret_cast<uintptr_t>(row) & 3u) == 0)   // two pixels per load
                for (; x + 1 < n; x += 2) m |= *reinterpret_cast<const uint32_t*>(row + x);
            for (; x < n; x++) m |= row[x];
        }
        need0[i] = (m | (m >> 16)) & 0xffffu;
    }
    // S6: the probe's own anchor first; when it sees no geometry the probe floats to the first anchor of the finer levels'
    // probes inside its cell that does (level by level downwards, row-major within a level)
    float3 d = f3(0.f, 0.f, 0.f);
    float t = -1.0f;
    uint32_t id = 0xffffffffu;
    if (!warp_mode) {
        // (thread mode: at most 1 + 4 + 16 candidates.)  All candidate anchors that lie inside the tile are looked up FIRST, with
        // independent loads — walking them one by one made every empty cell a chain of up to 21 dependent L2 round trips
        // (k_probes 25 -> 55 us at 4K) — then the candidates are resolved in S6's order; only anchors outside the tile trace.
        if (!anchor_hit<FLOATING>(s, cam, lv, tile, px, py, depth, prim, d, t, id) && floating && level >= 1 && !cell_empty(lv, cam, tile, px, py, occ, frame, ow, 0, 1)) {
            uint32_t hitm = 0u, unkm = 0u;       // candidate k (S6 order, own anchor excluded): G-buffer says hit / not in the tile
            int k = 0;
            for (int l = level - 1; l >= 0; l--) {
                const DLevel& fl = ls.lv[l];
                const int sc = 1 << (level - l);
                const int qy1 = min((py + 1) * sc, fl.gh), qx1 = min((px + 1) * sc, fl.gw);
                for (int qy = py * sc; qy < qy1; qy++)
                    for (int qx = px * sc; qx < qx1; qx++, k++) {
                        const int tx = min(qx * fl.P + fl.P / 2, cam.W - 1) - tile.x0, ty = min(qy * fl.P + fl.P / 2, cam.H - 1) - tile.y0;
                        if (tx >= 0 && tx < tile.w && ty >= 0 && ty < tile.h) { if (prim[(size_t)ty * tile.w + tx] != 0xffffffffu) hitm |= 1u << k; }
                        else unkm |= 1u << k;
                    }
            }
            if (hitm | unkm) {
                k = 0;
                for (int l = level - 1; l >= 0 && id == 0xffffffffu; l--) {
                    const DLevel& fl = ls.lv[l];
                    const int sc = 1 << (level - l);
                    const int qy1 = min((py + 1) * sc, fl.gh), qx1 = min((px + 1) * sc, fl.gw);
                    for (int qy = py * sc; qy < qy1 && id == 0xffffffffu; qy++)
                        for (int qx = px * sc; qx < qx1; qx++, k++)
                            if ((((hitm | unkm) >> k) & 1u) && anchor_hit<FLOATING>(s, cam, fl, tile, qx, qy, depth, prim, d, t, id)) break;
                }
            }
        }
    } else {
        // own anchor by lane 0; an empty cell (all 32 lanes look at its occupancy stamps together) ends the search at once
        bool own = false;
        if (lane == 0) own = anchor_hit<FLOATING>(s, cam, lv, tile, px, py, depth, prim, d, t, id);
        own = __shfl_sync(0xffffffffu, own ? 1 : 0, 0) != 0;
        const bool skip = own || !floating || level == 0 || __all_sync(0xffffffffu, cell_empty(lv, cam, tile, px, py, occ, frame, ow, (int)lane, 32));
        for (int l = skip ? -1 : level - 1; l >= 0 && id == 0xffffffffu; l--) {     // id is warp-uniform at every loop test
            const DLevel& fl = ls.lv[l];
            const int sc = 1 << (level - l);
            const int qx0 = px * sc, qy0 = py * sc;
            const int wdt = min((px + 1) * sc, fl.gw) - qx0, hgt = min((py + 1) * sc, fl.gh) - qy0;
            const int n = wdt > 0 && hgt > 0 ? wdt * hgt : 0;
            for (int base = 0; base < n; base += 32) {
                const int c = base + (int)lane;
                float3 dc = f3(0.f, 0.f, 0.f);
                float tc = -1.0f;
                uint32_t ic = 0xffffffffu;
                const bool hit = c < n && anchor_hit<FLOATING>(s, cam, fl, tile, qx0 + c % wdt, qy0 + c / wdt, depth, prim, dc, tc, ic);
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {                                            // the first candidate in row-major order
                    const int src = __ffs((int)m) - 1;
                    id = __shfl_sync(0xffffffffu, ic, src);
                    t = __shfl_sync(0xffffffffu, tc, src);
                    d = f3(__shfl_sync(0xffffffffu, dc.x, src), __shfl_sync(0xffffffffu, dc.y, src), __shfl_sync(0xffffffffu, dc.z, src));
                    break;
                }
            }
        }
        if (lane != 0) return;
    }
    if (id == 0xffffffffu) {
        origin[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
        normal[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float3 hp = vfma(t, d, cam.eye);
    const float3 e1 = xyz(__ldg(s.tri_eg + 2 * (size_t)id)), e2 = xyz(__ldg(s.tri_eg + 2 * (size_t)id + 1));
    float3 ng = vnormalize(vcross(e1, e2));
    if (vdot(ng, d) > 0.0f) ng = vneg(ng);
    const float3 og = vfma(offset, ng, hp);
    origin[gi] = make_float4(og.x, og.y, og.z, 1.0f);
    normal[gi] = make_float4(ng.x, ng.y, ng.z, 0.0f);
}

// ------------------------------------------------------------------ link (S1 + S8 weights), levels 0..N-2 in one launch
__device__ __forceinline__ void link_one(const DLevelSet& ls, unsigned gi, unsigned total, const float4* __restrict__ origin,
                                         const float4* __restrict__ normal, uint4* __restrict__ link_idx,
                                         float4* __restrict__ link_w)
{
    if (gi >= total) return;
    const int l = level_of(ls, gi);
    const DLevel &lo = ls.lv[l], &up = ls.lv[l + 1];
    const int i = (int)(gi - lo.probe_offset);
    const int px = lo.px0 + i % lo.sw, py = lo.py0 + i / lo.sw;
    int x0, x1, y0, y1;
    float wx0, wx1, wy0, wy1;
    upper_pair(px, up.gw, x0, x1, wx0, wx1);
    upper_pair(py, up.gh, y0, y1, wy0, wy1);
    const uint32_t k0 = (uint32_t)((y0 - up.py0) * up.sw + (x0 - up.px0));
    const uint32_t k1 = (uint32_t)((y0 - up.py0) * up.sw + (x1 - up.px0));
    const uint32_t k2 = (uint32_t)((y1 - up.py0) * up.sw + (x0 - up.px0));
    const uint32_t k3 = (uint32_t)((y1 - up.py0) * up.sw + (x1 - up.px0));
    const float4* up_origin = origin + up.probe_offset;
    const float4 og = origin[gi];
    float4 w = make_float4(-1.f, 0.f, 0.f, 0.f);
    if (og.w != 0.0f) {
        const float3 op = xyz(og), np = xyz(normal[gi]);
        const float w0 = (wx0 * wy0) * plane_weight(np, op, __ldg(up_origin + k0));
        const float w1 = (wx1 * wy0) * plane_weight(np, op, __ldg(up_origin + k1));
        const float w2 = (wx0 * wy1) * plane_weight(np, op, __ldg(up_origin + k2));
        const float w3 = (wx1 * wy1) * plane_weight(np, op, __ldg(up_origin + k3));
        const float S = ((w0 + w1) + w2) + w3;
        if (S > 0.0f) w = make_float4(w0 / S, w1 / S, w2 / S, w3 / S);
    }
    link_idx[gi] = make_uint4(k0, k1, k2, k3);
    link_w[gi] = w;
}

// ------------------------------------------------------------------ entry frontier (per probe, levels < n_levels)
// All D^2 rays of a probe share their origin and their interval [t0, t1): every triangle any of them can hit
// lies within the ball B(origin, t1).  One thread per probe walks the BVH with that ball and records a small
// frontier of subtrees (<= RC_ENTRY_SLOTS links) that covers every leaf whose box meets the ball; the probe's
// rays start their traversal there (trace(..., entry)) instead of at the root.  Closest hits are defined as
// min (t, triangle id) over ALL triangles (S5), so a conservative superset leaves results bit-identical.
// Expansion rule (expected node visits per ray, p = chance a ray meets a child's box):
//   replacing node X by its ball-meeting children saves the visit of X and costs (1 - p) per child; a child
//   that contains the origin has p = 1, a child that does not has p < 1/2.  Hence: expand when at most one
//   child meets the ball, or when one of two contains the origin; otherwise keep X.
__device__ __forceinline__ float box_dist2(float3 lo, float3 hi, float3 o)
{
    const float dx = fmaxf(fmaxf(lo.x - o.x, o.x - hi.x), 0.0f);
    const float dy = fmaxf(fmaxf(lo.y - o.y, o.y - hi.y), 0.0f);
    const float dz = fmaxf(fmaxf(lo.z - o.z, o.z - hi.z), 0.0f);
    return dx * dx + dy * dy + dz * dz;   // NaN / 1e30 "empty" boxes compare false against any radius
}

// One thread per GROUP of g x g neighbouring probes of a level (g = 2 at the dense lower levels, where the
// probes of a group lie well within one interval length of each other): the ball is grown to hold every
// member's ball and the frontier is written to each member's slot, so the march kernel indexes it by probe.
// Depth-first with the origin-holding child first; the frontier (emitted + pending) never exceeds
// RC_ENTRY_SLOTS, so one 8-entry array serves as output list (from the bottom) and stack (from the top).
__device__ __forceinline__ void entry_one(const DScene& s, const DLevelSet& ls, const EntryPlan& plan, unsigned gi,
                                          const float4* __restrict__ origin, int4* __restrict__ entry)
{
    if (gi >= plan.group_offset[plan.n]) return;
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < plan.n; k++) if (gi >= plan.group_offset[k]) l = k;
    const DLevel& lv = ls.lv[l];
    const int g = plan.g[l], ngx = (lv.sw + g - 1) / g;
    const int grp = (int)(gi - plan.group_offset[l]);
    const int bx = (grp % ngx) * g, by = (grp / ngx) * g;
    // bounding ball of the members' balls (g <= 2: the members' origins stay in registers)
    float3 c = f3(0.f, 0.f, 0.f);
    int nvalid = 0;
    float4 mo[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = k & 1, j = k >> 1;
        mo[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < g && j < g && bx + i < lv.sw && by + j < lv.sh) mo[k] = __ldg(origin + lv.probe_offset + (size_t)(by + j) * lv.sw + (bx + i));
    }
#pragma unroll
    for (int k = 0; k < 4; k++) if (mo[k].w != 0.0f) { c = vadd(c, xyz(mo[k])); nvalid++; }
    constexpr int F = RC_ENTRY_SLOTS;
    int a[F];
    int nout = 0, npend = 0;
    if (nvalid) {
        c = vscale(c, 1.0f / (float)nvalid);
        float spread2 = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; k++) if (mo[k].w != 0.0f) { const float3 dl = vsub(xyz(mo[k]), c); spread2 = fmaxf(spread2, vdot(dl, dl)); }
        const float r = (lv.t1 + sqrtf(spread2)) * 1.001f;   // |w| = 1 +- 2e-7; boxes carry their own 1e-4*diag pad
        const float r2 = r * r;
        a[F - 1] = 0; npend = 1;                             // the root
        while (npend) {
            const int X = a[F - npend]; npend--;
            const float4* nd = s.nodes + 4 * (size_t)X;
            const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
            const float d0 = box_dist2(f3(q0.x, q0.y, q0.z), f3(q0.w, q1.x, q1.y), c);
            const float d1 = box_dist2(f3(q1.z, q1.w, q2.x), f3(q2.y, q2.z, q2.w), c);
            const bool m0 = d0 <= r2, m1 = d1 <= r2;
            const bool in0 = m0 && d0 == 0.0f, in1 = m1 && d1 == 0.0f;
            if (m0 && m1 && (!(in0 || in1) || nout + npend + 2 > F)) { a[nout++] = X; continue; }   // keep X
            const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
            // leaves are final; inner children go on the stack, the origin-holding one on top
            if (m0 && c0 < 0) a[nout++] = c0;
            if (m1 && c1 < 0) a[nout++] = c1;
            const bool p0 = m0 && c0 >= 0, p1 = m1 && c1 >= 0;
            if (p0 && p1) {
                const bool first1 = in1 && !in0;
                npend++; a[F - npend] = first1 ? c0 : c1;
                // the deferred sibling is popped many iterations later: have its node in L1 by then
                asm volatile("prefetch.global.L1 [%0];" ::"l"(s.nodes + 4 * (size_t)(first1 ? c0 : c1)));
                npend++; a[F - npend] = first1 ? c1 : c0;
            } else if (p0) { npend++; a[F - npend] = c0; }
            else if (p1) { npend++; a[F - npend] = c1; }
        }
    }
    for (int k = nout; k < F; k++) a[k] = kDoneLinkC;
    const int4 ea = make_int4(a[0], a[1], a[2], a[3]), eb = make_int4(a[4], a[5], a[6], a[7]);
    for (int j = 0; j < g; j++)
        for (int i = 0; i < g; i++) {
            if (bx + i >= lv.sw || by + j >= lv.sh) continue;
            const size_t pi = lv.probe_offset + (size_t)(by + j) * lv.sw + (bx + i);
            entry[2 * pi] = ea;
            entry[2 * pi + 1] = eb;
        }
}

// k_link and k_entry are independent (both read only the probe origins) and each is a single short, latency-bound
// wave: one launch runs them side by side (blocks [0, link_blocks) link, the rest build entry frontiers).
__global__ void __launch_bounds__(kBlock) k_link_entry(DScene s, DLevelSet ls, EntryPlan plan, unsigned link_total, unsigned link_blocks,
                                                       const float4* __restrict__ origin, const float4* __restrict__ normal,
                                                       uint4* __restrict__ link_idx, float4* __restrict__ link_w, int4* __restrict__ entry)
{
    if (blockIdx.x < link_blocks) link_one(ls, blockIdx.x * kBlock + threadIdx.x, link_total, origin, normal, link_idx, link_w);
    else entry_one(s, ls, plan, (blockIdx.x - link_blocks) * kBlock + threadIdx.x, origin, entry);
}

// ------------------------------------------------------------------ direction culling: request masks and ray lists
// Request mask R_i of a probe of level i, one bit per REQUEST, row-major at resolution Dr_i:
//   level 0:   a request is a texel;  Dr_0 = D_0;      bit set by k_gbuffer (a pixel weights it with cs_d > 0)
//   level i>0: a request is a 2x2 quad of texels = one direction of level i-1;  Dr_i = D_{i-1}
// A lower texel that misses takes its far field from the quad below its own direction in each of its (up to
// four) upper probes (S8), so  R_1[k] |= R_0[p]  and  R_{i+1}[k] |= expand2x(R_i[p])  for every lower probe p
// and every upper probe k that p links to with non-zero weight (k_link's own tables: exactly the set far_field
// reads).  Texels outside the masks can only ever be multiplied by zero on the way to the irradiance.
// One launch per level i, bottom-up, one thread per (probe, mask word): (a) append the word's requests to the
// level's ray list (warp-aggregated append; entry = probe * Dr^2 + bit), (b) push the mask up to level i+1.
// 16 bits -> 32 bits, every bit doubled: bit i -> bits 2i, 2i+1
__device__ __forceinline__ uint32_t dup_spread16(uint32_t x)
{
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x | (x << 1);
}

// trace_inv's first iteration (node 0 = the root, far bound = tmax) with the slab reciprocals `inv` of a direction from the level's
// table: true when both child boxes are missed, i.e. the traversal would pop its empty stack and return "no hit" at once
__device__ __forceinline__ bool root_miss(float4 q0, float4 q1, float4 q2, float3 o, float3 inv, float tmin, float tmax)
{
    const float3 noi = f3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
    float ax = fmaf(q0.x, inv.x, noi.x), bx = fmaf(q0.w, inv.x, noi.x);
    float ay = fmaf(q0.y, inv.y, noi.y), by = fmaf(q1.x, inv.y, noi.y);
    float az = fmaf(q0.z, inv.z, noi.z), bz = fmaf(q1.y, inv.z, noi.z);
    const float n0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    const float f0 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    ax = fmaf(q1.z, inv.x, noi.x); bx = fmaf(q2.y, inv.x, noi.x);
    ay = fmaf(q1.w, inv.y, noi.y); by = fmaf(q2.z, inv.y, noi.y);
    az = fmaf(q2.x, inv.z, noi.z); bz = fmaf(q2.w, inv.z, noi.z);
    const float n1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    const float f1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    return !(n0 <= f0 || n1 <= f1);
}

template <int BLK>
__device__ __forceinline__ void need_body(const DLevel& lv, int Dr, int has_upper, int up_words, const float4* __restrict__ origin,
                                          const uint4* __restrict__ link_idx, const float4* __restrict__ link_w,
                                          uint32_t* __restrict__ need, uint32_t* __restrict__ need_up,
                                          uint32_t* __restrict__ list, unsigned int* __restrict__ count, int clear, int trigger, int dir_major,
                                          int tile_order, int append, int4 own, size_t gi, unsigned* s_warp, unsigned* s_base_p, bool grid_dep)
{
    // The per-level launches form a chain of short, latency-bound waves.  Launched with programmatic stream
    // serialization (launch_need pdl) level i+1 is set up while level i still runs: it may read what kernels before
    // level i produced (origins, link tables) at once and waits for level i — whose atomicOr's fill its masks — only
    // before it reads them.  `trigger` is set only when the NEXT launch in the stream is such a k_need: the march kernel
    // that follows the last level reads the list length at its very top and must not start early.
    if (trigger) cudaTriggerProgrammaticLaunchCompletion();
    const int bits = Dr * Dr, words = (bits + 31) >> 5;
    const size_t total = (size_t)lv.sw * lv.sh * words;
    uint32_t r = 0u, probe = 0u;
    int w = 0;
    float4 lw = make_float4(-1.f, 0.f, 0.f, 0.f);
    uint4 li = make_uint4(0u, 0u, 0u, 0u);
    if (gi < total) {
        probe = (uint32_t)(gi / words);
        w = (int)(gi - (size_t)probe * words);
        // independent loads first: the kernel is one short latency-bound wave
        const float valid = __ldg(origin + probe).w;
        if (has_upper) { lw = __ldg(link_w + probe); li = __ldg(link_idx + probe); }
        if (grid_dep) cudaGridDependencySynchronize();
        r = need[gi];
        if (valid == 0.0f) r = 0u;
        if (append && own.z >= 0) {
            // halo exchange: this rank marches only the probes it OWNS (own = x0, y0, x1, y1 inclusive, sub-grid coordinates);
            // the requests of the others were sent to their owners before this launch and are dropped here
            const int ppx = (int)(probe % (uint32_t)lv.sw), ppy = (int)(probe / (uint32_t)lv.sw);
            if (ppx < own.x || ppx > own.z || ppy < own.y || ppy > own.w) r = 0u;
        }
        if (w == words - 1 && (bits & 31)) r &= (1u << (bits & 31)) - 1u;
        // consume and clear: the masks of levels >= 1 are accumulated by atomicOr, so they must be empty when the next
        // frame starts (level 0 is overwritten by k_probes).  A stale bit could only ever add a ray, never lose one.
        if (clear) need[gi] = 0u;
    }
    // (a) ray list: block-aggregated append (one atomicAdd per block; all blocks hit the same counter)
    unsigned& s_base = *s_base_p;
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    // tile-ordered append (tile_order != 0, levels whose requests are quads): permuted request masks
    int tiled = 0;
    uint32_t pm = 0u;
    unsigned long long pm64 = 0ull;
    if (tile_order && !dir_major) {
        if (Dr == 16) {          // word = rows 2k, 2k+1 of 16: interleave the nibbles of the two rows
            uint32_t lo = r & 0xffffu, hi = r >> 16;
            lo = (lo | (lo << 8)) & 0x00ff00ffu; lo = (lo | (lo << 4)) & 0x0f0f0f0fu;
            hi = (hi | (hi << 8)) & 0x00ff00ffu; hi = (hi | (hi << 4)) & 0x0f0f0f0fu;
            pm = lo | (hi << 4);
            tiled = 1;
        } else if (Dr == 8) {    // word = four rows of 8 = nibbles (row, half): swap nibbles 1 <-> 2 and 5 <-> 6
            pm = (r & 0xf00ff00fu) | ((r & 0x00f000f0u) << 4) | ((r & 0x0f000f00u) >> 4);
            tiled = 1;
        } else if (Dr == 32) {   // word = one row of 32: rows (2k, 2k+1) are adjacent lanes (32 words per probe = one warp)
            const uint32_t partner = __shfl_down_sync(0xffffffffu, r, 1);
            if (!(w & 1)) {
                unsigned long long a = r, b = partner;
                a = (a | (a << 16)) & 0x0000ffff0000ffffull; a = (a | (a << 8)) & 0x00ff00ff00ff00ffull; a = (a | (a << 4)) & 0x0f0f0f0f0f0f0f0full;
                b = (b | (b << 16)) & 0x0000ffff0000ffffull; b = (b | (b << 8)) & 0x00ff00ff00ff00ffull; b = (b | (b << 4)) & 0x0f0f0f0f0f0f0f0full;
                pm64 = a | (b << 4);
            }
            tiled = 2;
        }
    }
    int n = tiled == 2 ? __popcll(pm64) : __popc(r);
    if (!append) { n = 0; pm = 0u; pm64 = 0ull; }      // masks only (halo exchange, first pass): nothing is listed yet
    int pre = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if ((int)lane >= o) pre += t; }
    if (lane == 31) s_warp[wid] = (unsigned)pre;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int k = 0; k < BLK / 32; k++) { const unsigned t = s_warp[k]; s_warp[k] = tot; tot += t; }
        s_base = tot ? atomicAdd(count, tot) : 0u;
    }
    __syncthreads();
    if (dir_major && append) {
        // Direction-major order inside the warp's share of the list: the warp's threads are 32 / words neighbouring
        // probes of a row; their requests are appended request by request (same direction across the probes) instead
        // of probe by probe.  A warp of k_march then marches (nearly) parallel rays from adjacent origins — same
        // nodes, same hit / miss outcome — instead of one probe's whole fan.  The list is a set: its order cannot
        // change any texel.
        unsigned wbase = s_base + s_warp[wid];
        const unsigned lt = (1u << lane) - 1u;
        for (int ww = 0; ww < words; ww++) {
            for (uint32_t any = __reduce_or_sync(0xffffffffu, w == ww ? r : 0u); any; any &= any - 1u) {
                const int d = __ffs((int)any) - 1;
                const bool on = w == ww && ((r >> d) & 1u);
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) list[wbase + (unsigned)__popc(m & lt)] = probe * (uint32_t)bits + (uint32_t)(32 * ww + d);
                wbase += (unsigned)__popc(m);
            }
        }
    } else if (tiled == 2) {
        // Dr = 32: the even-row thread appends its own and its odd-row partner's requests, 4x2-tile by 4x2-tile
        // (see below); position p of the 64-bit permuted mask = tile * 8 + row * 4 + column-in-tile
        unsigned base = s_base + s_warp[wid] + (unsigned)(pre - n);
        for (unsigned long long m = pm64; m; m &= m - 1ull) {
            const int p = __ffsll((long long)m) - 1;
            list[base++] = probe * (uint32_t)bits + (uint32_t)(32 * (w + ((p >> 2) & 1)) + (p >> 3) * 4 + (p & 3));
        }
    } else if (tiled == 1) {
        // Tile order: a k_march warp takes 8 consecutive entries (8 quads of 2x2 texels).  In row-major order those are a
        // 16 x 2 strip of directions — half the width of the octahedral map; ordered 4x2-tile by 4x2-tile they are an 8 x 4
        // block of directions, a compact cone whose rays visit the same nodes and end together (ncu r2e: the node loop of
        // level 3 ran 7.1 iterations per warp with 14 of 32 lanes active).  The list is a set: order cannot change a texel.
        unsigned base = s_base + s_warp[wid] + (unsigned)(pre - n);
        for (uint32_t m = pm; m; m &= m - 1u) {
            const int p = __ffs((int)m) - 1;
            int ob;
            if (Dr == 16) ob = ((p >> 2) & 1) * 16 + (p >> 3) * 4 + (p & 3);
            else { const int np = p >> 2; ob = ((np & 4) | ((np & 1) << 1) | ((np >> 1) & 1)) * 4 + (p & 3); }   // Dr == 8
            list[base++] = probe * (uint32_t)bits + (uint32_t)(32 * w + ob);
        }
    } else if (append) {
        unsigned base = s_base + s_warp[wid] + (unsigned)(pre - n);
        for (uint32_t m = r; m; m &= m - 1u) list[base++] = probe * (uint32_t)bits + (uint32_t)(32 * w + __ffs((int)m) - 1);
    }
    // (b) the upper level's requests
    if (!has_upper || !r) return;
    if (lw.x < 0.0f) return;                       // no valid upper probe: the far field is the sky (S8)
    uint32_t tw_idx[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, tw_val[4] = {0u, 0u, 0u, 0u};
    if (has_upper == 1) {                          // level 0 -> 1: same resolution, no expansion
        tw_idx[0] = (uint32_t)w; tw_val[0] = r;
    } else if ((Dr & (Dr - 1)) == 0 && Dr >= 2) {  // expand2x by bit spreading (power-of-two Dr: every default level)
        if (Dr >= 32) {        // the word is 32 requests of row y: rows 2y and 2y+1 of the target get the same two words
            const int wpr = Dr >> 5, y = w / wpr, xw = w - y * wpr;
            const uint32_t a = dup_spread16(r & 0xffffu), b = dup_spread16(r >> 16);
            const uint32_t row0 = (uint32_t)((2 * y) * (2 * wpr) + 2 * xw), row1 = row0 + (uint32_t)(2 * wpr);
            tw_idx[0] = row0; tw_val[0] = a; tw_idx[1] = row0 + 1; tw_val[1] = b;
            tw_idx[2] = row1; tw_val[2] = a; tw_idx[3] = row1 + 1; tw_val[3] = b;
        } else if (Dr == 16) { // two rows of 16 -> four target rows of 32
            const uint32_t a = dup_spread16(r & 0xffffu), b = dup_spread16(r >> 16);
            tw_idx[0] = 4u * w; tw_val[0] = a; tw_idx[1] = 4u * w + 1; tw_val[1] = a;
            tw_idx[2] = 4u * w + 2; tw_val[2] = b; tw_idx[3] = 4u * w + 3; tw_val[3] = b;
        } else if (Dr == 8) {  // four rows of 8 -> eight target rows of 16, two per word
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t t = dup_spread16((r >> (8 * q)) & 0xffu);
                tw_idx[q] = 4u * w + q; tw_val[q] = t | (t << 16);
            }
        } else if (Dr == 4) {  // four rows of 4 -> eight target rows of 8, four per word
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t t0 = dup_spread16((r >> (8 * q)) & 0xfu), t1 = dup_spread16((r >> (8 * q + 4)) & 0xfu);
                tw_idx[q] = (uint32_t)q; tw_val[q] = t0 | (t0 << 8) | (t1 << 16) | (t1 << 24);
            }
        } else {               // Dr == 2: two rows of 2 -> four target rows of 4 in one word
            const uint32_t t0 = dup_spread16(r & 3u), t1 = dup_spread16((r >> 2) & 3u);
            tw_idx[0] = 0u; tw_val[0] = t0 | (t0 << 4) | (t1 << 8) | (t1 << 12);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) if (!tw_val[q]) tw_idx[q] = 0xffffffffu;
    } else {                                       // expand2x: request (x, y) -> requests (2x..2x+1, 2y..2y+1) at 2*Dr
        const int D2 = 2 * Dr;
        for (uint32_t m = r; m; m &= m - 1u) {
            const int sbit = 32 * w + __ffs((int)m) - 1, x = sbit % Dr, y = sbit / Dr;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t tbit = (uint32_t)((2 * y + j) * D2 + 2 * x);   // even: bits tbit, tbit+1 share a word
                const uint32_t ti = tbit >> 5, tv = 3u << (tbit & 31u);
                int slot = -1;
#pragma unroll
                for (int q = 0; q < 4; q++) if (slot < 0 && (tw_idx[q] == ti || tw_idx[q] == 0xffffffffu)) slot = q;
                if (slot < 0) {   // more than four target words (non-power-of-two Dr): flush one
                    const uint32_t up[4] = {li.x, li.y, li.z, li.w};
                    const float wk[4] = {lw.x, lw.y, lw.z, lw.w};
                    for (int k = 0; k < 4; k++) if (wk[k] > 0.0f) atomicOr(need_up + (size_t)up[k] * up_words + tw_idx[0], tw_val[0]);
                    slot = 0; tw_idx[0] = 0xffffffffu; tw_val[0] = 0u;
                }
                tw_idx[slot] = ti; tw_val[slot] |= tv;
            }
        }
    }
    const uint32_t up[4] = {li.x, li.y, li.z, li.w};
    const float wk[4] = {lw.x, lw.y, lw.z, lw.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!(wk[k] > 0.0f)) continue;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (tw_idx[q] != 0xffffffffu) atomicOr(need_up + (size_t)up[k] * up_words + tw_idx[q], tw_val[q]);
    }
}


__global__ void __launch_bounds__(kBlock) k_need(DLevel lv, int Dr, int has_upper, int up_words, const float4* __restrict__ origin,
                                                 const uint4* __restrict__ link_idx, const float4* __restrict__ link_w,
                                                 uint32_t* __restrict__ need, uint32_t* __restrict__ need_up,
                                                 uint32_t* __restrict__ list, unsigned int* __restrict__ count, int clear, int trigger, int dir_major,
                                                 int tile_order, int append, int4 own)
{
    __shared__ unsigned s_warp[kBlock / 32], s_base;
    need_body<kBlock>(lv, Dr, has_upper, up_words, origin, link_idx, link_w, need, need_up, list, count, clear, trigger, dir_major, tile_order,
                      append, own, (size_t)blockIdx.x * kBlock + threadIdx.x, s_warp, &s_base, true);
}

// Levels >= 1 of the request chain in ONE launch: a cluster of eight 1024-thread blocks walks the levels bottom-up with a
// cluster barrier (release / acquire at cluster scope: the atomicOr pushes of level i are visible to every block before level
// i+1 reads its masks) instead of a kernel boundary between them.  Each of those levels is a fraction of a wave of work; as
// separate launches they cost 12-14 us apiece at 4K (7-10 us at 1080p and in every strip of a tiled frame) whatever their size.
constexpr int kChainBlock = 1024, kChainCtas = 8;

__global__ void __launch_bounds__(kChainBlock) k_need_chain(NeedChain c, const float4* __restrict__ origin, const uint4* __restrict__ link_idx,
                                                            const float4* __restrict__ link_w, uint32_t* __restrict__ need_all,
                                                            uint32_t* __restrict__ list_all, unsigned int* __restrict__ counts)
{
    __shared__ unsigned s_warp[kChainBlock / 32], s_base;
    for (int i = c.first; i <= c.last; i++) {
        const DLevel& lv = c.lv[i];
        const int words = (c.Dr[i] * c.Dr[i] + 31) >> 5;
        const size_t total = (size_t)lv.sw * lv.sh * words;
        for (size_t base = 0; base < total; base += (size_t)kChainBlock * kChainCtas) {
            need_body<kChainBlock>(lv, c.Dr[i], c.has_upper[i], c.up_words[i], origin + lv.probe_offset, link_idx + lv.probe_offset,
                                   link_w + lv.probe_offset, need_all + c.need_off[i], need_all + c.need_up_off[i], list_all + c.list_off[i],
                                   counts + i, c.clear[i], 0, c.dir_major[i], c.tile_order[i], c.append, c.own[i],
                                   base + (size_t)blockIdx.x * kChainBlock + threadIdx.x, s_warp, &s_base, false);
            __syncthreads();      // s_warp / s_base are reused by the next chunk
        }
        if (i < c.last) {
            __threadfence();
            asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
}

// ------------------------------------------------------------------ split ray lists
// In the open scenes half of a level's requests or more are rays that leave the scene's bounds at once; in k_march they idle
// next to lanes that traverse for hundreds of instructions (ncu r2e: 14 of 32 lanes active in level 3's node loop).  k_split
// runs once per frame between the request chain and the march, one thread per list entry of every level chosen by the host:
// an entry whose rays ALL fail the first step of the traversal — root_miss, the very expressions of trace_inv's first iteration
// — is a certain miss.  The level's list is copied into a second one, entries that enter the tree at the front (in their old
// order: the 4x2-tile order of k_need survives), certain misses at the back; k_march gives the latter threads that skip
// the traversal.  Same rays, same outcome for each: every texel is unchanged.
constexpr int kSplitBlock = 256, kSplitChunk = 2048;

// true when every ray of list entry e (levels >= 1: the 2x2 children of a quad; level 0: one texel) misses both root boxes
__device__ __forceinline__ bool entry_misses(const SplitJob& jb, float4 q0, float4 q1, float4 q2, bool quad, int ld, uint32_t DD, uint32_t e)
{
    const DLevel& lv = jb.lv;
    if (quad) {
        // consecutive entries of a list are mostly consecutive quads of one probe: one broadcast origin and three coalesced
        // 16-byte loads per lane (through the per-direction table the four children cost eight loads on 64-byte strides,
        // and the kernel was bound by L1 wavefronts, not by its arithmetic)
        const uint32_t qd = e & ((DD >> 2) - 1u), probe = e >> (2 * ld - 2);
        const float3 o = xyz(__ldg(jb.origin + probe));
        const float4 a = __ldg(jb.qinv + 3 * (size_t)qd), b = __ldg(jb.qinv + 3 * (size_t)qd + 1), c = __ldg(jb.qinv + 3 * (size_t)qd + 2);
        // one child that enters the tree settles the quad: neighbouring entries mostly agree, so whole warps leave after the first test
        if (!root_miss(q0, q1, q2, o, f3(a.x, a.y, a.z), lv.t0, lv.t1)) return false;
        if (!root_miss(q0, q1, q2, o, f3(c.y, c.z, c.w), lv.t0, lv.t1)) return false;      // (the diagonally opposite child next)
        return root_miss(q0, q1, q2, o, f3(a.w, b.x, b.y), lv.t0, lv.t1) & root_miss(q0, q1, q2, o, f3(b.z, b.w, c.x), lv.t0, lv.t1);
    }
    const uint32_t probe = e >> (2 * ld), d = e & (DD - 1u);
    const float3 o = xyz(__ldg(jb.origin + probe));
    const float4 a0 = __ldg(jb.dirq + 2 * (size_t)d), a1 = __ldg(jb.dirq + 2 * (size_t)d + 1);
    return root_miss(q0, q1, q2, o, f3(a0.w, a1.x, a1.y), lv.t0, lv.t1);
}

// Block `blk` of `nblk` working on one level's list: chunks of kSplitChunk consecutive entries, kSplitBlock consecutive entries per
// step (one per thread, coalesced), all steps classified before the chunk is appended with ONE atomicAdd per class — a block's
// critical path holds one atomic round trip, not one per step, and the chunk lands in the new list as a whole and in order.
constexpr int kSplitSteps = kSplitChunk / kSplitBlock;

__device__ __forceinline__ void split_body(const SplitJob& jb, unsigned blk, unsigned nblk, unsigned* s_cnt, unsigned* s_base)
{
    const unsigned n = jb.counts[jb.level];
    unsigned int* n_triv = jb.counts + RC_MAX_LEVELS + jb.level;      // bit 31: "this level was classified in this frame" (rc_api: adaptive choice)
    unsigned int* n_real = jb.counts + 2 * RC_MAX_LEVELS + jb.level;
    const bool quad = jb.level >= 1;
    const int ld = 31 - __clz(jb.lv.D);
    const uint32_t DD = (uint32_t)jb.lv.D * (uint32_t)jb.lv.D;
    if (blk == 0 && threadIdx.x == 0) atomicOr(n_triv, 0x80000000u);
    const float4 q0 = __ldg(jb.root), q1 = __ldg(jb.root + 1), q2 = __ldg(jb.root + 2);
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    constexpr int kWarps = kSplitBlock / 32;
    static_assert(kSplitSteps * kWarps == 64, "the chunk's (step, warp) cells are scanned by one warp, two per lane");
    for (size_t chunk = (size_t)blk * kSplitChunk; chunk < n; chunk += (size_t)nblk * kSplitChunk) {
        uint32_t e[kSplitSteps];
        unsigned have = 0u, miss = 0u;
#pragma unroll
        for (int k = 0; k < kSplitSteps; k++) {
            const size_t j = chunk + (size_t)k * kSplitBlock + threadIdx.x;
            e[k] = 0u;
            if (j < n) { e[k] = __ldg(jb.list_a + j); have |= 1u << k; }
        }
#pragma unroll
        for (int k = 0; k < kSplitSteps; k++)
            if (((have >> k) & 1u) && entry_misses(jb, q0, q1, q2, quad, ld, DD, e[k])) miss |= 1u << k;
        // stable partition of the chunk: (entering, certain miss) counts per (step, warp) cell packed 16 : 16, cells in list order
#pragma unroll
        for (int k = 0; k < kSplitSteps; k++) {
            const unsigned br = __ballot_sync(0xffffffffu, ((have & ~miss) >> k) & 1u), bt = __ballot_sync(0xffffffffu, (miss >> k) & 1u);
            if (lane == 0) s_cnt[k * kWarps + wid] = (unsigned)__popc(br) | ((unsigned)__popc(bt) << 16);
        }
        __syncthreads();
        if (wid == 0) {
            const unsigned a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
            unsigned incl = a + b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
            s_cnt[2 * lane] = incl - (a + b);
            s_cnt[2 * lane + 1] = incl - b;
            if (lane == 31) {
                s_base[0] = (incl & 0xffffu) ? atomicAdd(n_real, incl & 0xffffu) : 0u;
                s_base[1] = (incl >> 16) ? (atomicAdd(n_triv, incl >> 16) & 0x7fffffffu) : 0u;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSplitSteps; k++) {
            const bool r = ((have & ~miss) >> k) & 1u, t = (miss >> k) & 1u;
            const unsigned br = __ballot_sync(0xffffffffu, r), bt = __ballot_sync(0xffffffffu, t);
            const unsigned cell = s_cnt[k * kWarps + wid];
            if (r) jb.list_b[s_base[0] + (cell & 0xffffu) + (unsigned)__popc(br & lt)] = e[k];
            else if (t) jb.list_b[jb.cap - 1u - (s_base[1] + (cell >> 16) + (unsigned)__popc(bt & lt))] = e[k];
        }
        __syncthreads();      // s_cnt / s_base are reused by the next chunk
    }
}

__global__ void __launch_bounds__(kSplitBlock) k_split(SplitPlan plan)
{
    __shared__ unsigned s_cnt[kSplitSteps * (kSplitBlock / 32)], s_base[2];
    int k = 0;
    while (k + 1 < plan.n && blockIdx.x >= plan.block_off[k + 1]) k++;
    split_body(plan.job[k], blockIdx.x - plan.block_off[k], plan.block_off[k + 1] - plan.block_off[k], s_cnt, s_base);
}

// far-field radiance of lower texel (dx, dy) from the merged upper level (S8).
// `up_avg` holds, per upper probe and LOWER direction, a_k = 0.25*(((c0 + c1) + c2) + c3) of the four child texels
// (float16 values summed in float32, exactly the S8 expression) — written once by the kernel that finalised
// level i+1 (k_march's warp-shuffle epilogue / k_fill_top / k_child_avg).  Each upper texel is needed by 16 lower
// probes; averaging at the producer replaces 8 x 16-byte loads, 32 half2 conversions and 16 adds per lower texel
// by 4 x 16-byte loads here.
// up_const: the upper level is a top level that cannot hit anything (k_fill_top's case): every texel of a valid
// upper probe is (sky, 1) rounded to float16, of an invalid one (0,0,0,1), and so are their child averages —
// `up_avg` then points at the upper probes' ORIGINS (w = valid) and nothing needs to be filled or stored.
__device__ __forceinline__ float4 far_field(const float4* __restrict__ up_avg, int D, uint4 li, float4 lw, int dx, int dy, float3 sky,
                                            bool up_const = false)
{
    if (lw.x < 0.0f) return make_float4(sky.x, sky.y, sky.z, 0.f);   // S8: no valid upper probe -> the sky
    float4 far = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t idx[4] = {li.x, li.y, li.z, li.w};
    const float wk[4] = {lw.x, lw.y, lw.z, lw.w};
    const size_t DD = (size_t)D * D, off = (size_t)dy * D + dx;
    const float4 skyh = unpack_half4(pack_half4(sky.x, sky.y, sky.z, 1.0f));
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float4 a;
        if (up_const) a = __ldg(up_avg + idx[k]).w != 0.0f ? skyh : make_float4(0.f, 0.f, 0.f, 1.f);
        else a = ld_f4(up_avg + idx[k] * DD + off);   // coherent: written by the PDL predecessor (see ld_u4)
        far.x = fmaf(wk[k], a.x, far.x);
        far.y = fmaf(wk[k], a.y, far.y);
        far.z = fmaf(wk[k], a.z, far.z);
        far.w = fmaf(wk[k], a.w, far.w);
    }
    return far;
}

// Child average of one finalised texel quad, computed where the texels are produced: the 2x2 children of a
// lower direction sit in lanes l, l^1, l^ystep, l^(1|ystep) of the warp (ystep = 8 for the 8x4 direction tile
// and for LINEAR with D = 8; = D for LINEAR with D <= 16).  Every lane must call this (full-mask shuffles);
// the lane with even dx and even dy returns true and owns the result.
__device__ __forceinline__ bool child_avg_shfl(uint2 t, int ystep, int dx, int dy, float4& avg)
{
    constexpr unsigned FULL = 0xffffffffu;
    const uint2 tx = make_uint2(__shfl_xor_sync(FULL, t.x, 1), __shfl_xor_sync(FULL, t.y, 1));
    const uint2 ty = make_uint2(__shfl_xor_sync(FULL, t.x, ystep), __shfl_xor_sync(FULL, t.y, ystep));
    const uint2 txy = make_uint2(__shfl_xor_sync(FULL, t.x, ystep | 1), __shfl_xor_sync(FULL, t.y, ystep | 1));
    if ((dx | dy) & 1) return false;
    const float4 c0 = unpack_half4(t), c1 = unpack_half4(tx), c2 = unpack_half4(ty), c3 = unpack_half4(txy);
    avg = make_float4(0.25f * (((c0.x + c1.x) + c2.x) + c3.x), 0.25f * (((c0.y + c1.y) + c2.y) + c3.y),
                      0.25f * (((c0.z + c1.z) + c2.z) + c3.z), 0.25f * (((c0.w + c1.w) + c2.w) + c3.w));
    return true;
}

// ------------------------------------------------------------------ march (+ fused merge)
// One thread per texel in storage order: a warp = 32 consecutive directions of one probe
// (level 0 with D0 = 4: two probes), so origins are shared and stores are one 256-byte segment.
// Thread -> texel mapping (the storage layout is always probe-major):
//   MAP_LINEAR     storage order: a warp = 32 consecutive directions of one probe
//   MAP_DIR_TILE   a warp = an 8x4 tile of directions of one probe (compact cone from a shared origin)
//   MAP_PROBE_TILE a warp = one direction of an 8x4 tile of neighbouring probes (parallel rays, nearby origins)
enum { MAP_LINEAR = 0, MAP_DIR_TILE = 1, MAP_PROBE_TILE = 2 };

__device__ __forceinline__ bool decode_texel(const DLevel& lv, int map, size_t g, uint32_t& probe, uint32_t& d)
{
    const uint32_t DD = (uint32_t)(lv.D * lv.D);
    if (map == MAP_PROBE_TILE) {
        const uint32_t ntx = (uint32_t)(lv.sw + 7) >> 3, nty = (uint32_t)(lv.sh + 3) >> 2;
        const size_t warp = g >> 5;
        const uint32_t lane = (uint32_t)g & 31u;
        const uint32_t ptile = (uint32_t)(warp / DD);
        d = (uint32_t)(warp - (size_t)ptile * DD);
        if (ptile >= ntx * nty) return false;
        const uint32_t px = (ptile % ntx) * 8 + (lane & 7u), py = (ptile / ntx) * 4 + (lane >> 3);
        if (px >= (uint32_t)lv.sw || py >= (uint32_t)lv.sh) return false;
        probe = py * (uint32_t)lv.sw + px;
        return true;
    }
    if (g >= (size_t)lv.sw * lv.sh * DD) return false;
    if ((lv.D & (lv.D - 1)) == 0) {
        // D is a power of two (every default level): shifts instead of two 64/32-bit divisions (~60 SASS
        // instructions, a third of the set-up cost of a ray that leaves the scene immediately)
        const int ld = 31 - __clz(lv.D);
        probe = (uint32_t)(g >> (2 * ld));
        const uint32_t j = (uint32_t)g & (DD - 1u);
        if (map == MAP_DIR_TILE) {
            const uint32_t tile = j >> 5, lane = j & 31u, tsh = (uint32_t)(ld - 3), tmask = ((uint32_t)lv.D >> 3) - 1u;
            d = (((tile >> tsh) * 4 + (lane >> 3)) << ld) + (tile & tmask) * 8 + (lane & 7u);
        } else {
            d = j;
        }
        return true;
    }
    probe = (uint32_t)(g / DD);
    const uint32_t j = (uint32_t)(g - (size_t)probe * DD);
    if (map == MAP_DIR_TILE) {
        const uint32_t tile = j >> 5, lane = j & 31u, tpr = (uint32_t)lv.D >> 3;
        d = ((tile / tpr) * 4 + (lane >> 3)) * (uint32_t)lv.D + (tile % tpr) * 8 + (lane & 7u);
    } else {
        d = j;
    }
    return true;
}

// S7 (+ S8 when fused): radiance of one texel from its closest hit, packed RGBA16F
template <bool FUSED>
__device__ __forceinline__ uint2 finalize_texel(const DScene& s, const DLights& L, const DLevel& lv, int UD, int top, float3 sky,
                                                uint32_t probe, uint32_t d, float3 o, float3 w, const Hit& h,
                                                const float4* __restrict__ up_avg, const uint4* __restrict__ link_idx,
                                                const float4* __restrict__ link_w)
{
    float4 c;
    if (h.prim != 0xffffffffu) {
        const float3 P = vfma(h.t, w, o);
        const Shade sh = shade_hit(s, L, h.prim, h.u, h.v, P, vneg(w));
        c = make_float4(sh.rad.x, sh.rad.y, sh.rad.z, 0.0f);
    } else if (top) {
        c = make_float4(sky.x, sky.y, sky.z, 1.0f);
    } else {
        c = make_float4(0.f, 0.f, 0.f, 1.0f);
    }
    if (FUSED && !top) {
        // Programmatic dependent launch: this level's kernel may have started while level i+1's tail was
        // still running; everything above (the whole ray march) is independent of it.  Wait here, right
        // before the first read of level i+1's merged texels (no-op when launched without the PDL attribute).
        cudaGridDependencySynchronize();
        // S8 applied to the float16-rounded raw texel (the same value the unfused path reads back)
        const float4 raw = unpack_half4(pack_half4(c.x, c.y, c.z, c.w));
        if (raw.w != 0.0f) {
            int dx, dy;
            if ((lv.D & (lv.D - 1)) == 0) { dx = (int)(d & (uint32_t)(lv.D - 1)); dy = (int)(d >> (31 - __clz(lv.D))); }
            else { dx = (int)(d % (uint32_t)lv.D); dy = (int)(d / (uint32_t)lv.D); }
            const float4 far = far_field(up_avg, lv.D, __ldg(link_idx + probe), __ldg(link_w + probe), dx, dy, sky, UD < 0);
            c = make_float4(fmaf(raw.w, far.x, raw.x), fmaf(raw.w, far.y, raw.y), fmaf(raw.w, far.z, raw.z), raw.w * far.w);
            c.x = fminf(c.x, 65504.0f); c.y = fminf(c.y, 65504.0f); c.z = fminf(c.z, 65504.0f);
        } else {
            c = raw;   // a = 0: fma(0, far, raw) = raw exactly, the far field cannot contribute
        }
    }
    return pack_half4(c.x, c.y, c.z, c.w);
}

// MINB = minimum resident 128-thread blocks per SM the register allocation must allow (8 -> 64 regs,
// 12 -> 40 regs, 16 -> 32 regs): the kernel is latency-bound (ncu: ~30 % warps active, ~7 dependent node
// loads per warp), so occupancy is traded against spills and measured.
template <bool FUSED, int MINB>
__global__ void __launch_bounds__(128, MINB) k_march(DScene s, DLights L, DLevel lv, int UD, int top, float3 sky, int map, size_t total,
                                                  const float4* __restrict__ origin, const float4* __restrict__ dirq,
                                                  uint2* __restrict__ texels, const float4* __restrict__ up_avg,
                                                  const uint4* __restrict__ link_idx, const float4* __restrict__ link_w,
                                                  const int4* __restrict__ entry, float4* __restrict__ avg_out, int ystep,
                                                  const uint32_t* __restrict__ list, const unsigned int* __restrict__ count, int quad,
                                                  const unsigned int* __restrict__ count_triv, unsigned list_cap)
{
    cudaTriggerProgrammaticLaunchCompletion();   // the next level's kernel may begin once every block got here
    const size_t DD = (size_t)lv.D * lv.D;
    // culled mode (list != null): thread j marches the j-th requested texel of the level's ray list (k_need) —
    // for levels >= 1 the list holds 2x2 quads, four consecutive lanes each, so the child average is two shuffles.
    // Entries [0, n_real) are read from the front of the list; the n_triv entries after them from its back: certain
    // misses (k_split), which get the merge and the store but no traversal.
    unsigned n_real = 0u;
    if (list) {
        n_real = __ldg(count);
        total = ((size_t)n_real + (count_triv ? (__ldg(count_triv) & 0x7fffffffu) : 0u)) * (quad ? 4u : 1u);
    }
    // grid-stride over whole warps (the child-average epilogue shuffles with a full mask): with a full-size grid
    // this is one iteration; with a resident grid (rc_set_tuning "march_waves") each warp walks packets g, g + G, ...
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31u;
    for (size_t g0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); g0 < total; g0 += stride) {
        uint32_t probe = 0, d = 0;
        bool in_range, certain_miss = false;
        if (list) {
            const size_t j = g0 + lane;
            in_range = j < total;
            if (in_range) {
                const int ld = 31 - __clz(lv.D);   // culled mode requires power-of-two D (rc_api)
                const size_t q = quad ? j >> 2 : j;
                certain_miss = q >= n_real;
                const uint32_t e = __ldg(list + (certain_miss ? (size_t)list_cap - 1u - (q - n_real) : q));
                if (quad) {
                    const uint32_t q = e & (uint32_t)((DD >> 2) - 1);
                    const uint32_t x = q & (uint32_t)((lv.D >> 1) - 1), y = q >> (ld - 1);
                    probe = e >> (2 * ld - 2);
                    d = ((2u * y + (((uint32_t)j >> 1) & 1u)) << ld) + 2u * x + ((uint32_t)j & 1u);
                } else {
                    probe = e >> (2 * ld);
                    d = e & (uint32_t)(DD - 1);
                }
            }
        } else {
            in_range = decode_texel(lv, map, g0 + lane, probe, d);
        }
        uint2 t = pack_half4(0.f, 0.f, 0.f, 1.f);   // invalid probe (S7)
        if (in_range) {
            const float4 og = __ldg(origin + probe);
            if (og.w != 0.0f) {
                // direction and its slab reciprocals from the level's table (divided once on the host, same IEEE expression)
                const float4 qa = __ldg(dirq + 2 * (size_t)d), qb = __ldg(dirq + 2 * (size_t)d + 1);
                const float3 w = xyz(qa);
                const float3 o = xyz(og);
                Hit h; h.t = -1.0f; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
                if (!certain_miss) h = trace_inv(s, o, w, f3(qa.w, qb.x, qb.y), lv.t0, lv.t1, entry ? entry + 2 * (size_t)probe : nullptr);
                t = finalize_texel<FUSED>(s, L, lv, UD, top, sky, probe, d, o, w, h, up_avg, link_idx, link_w);
            }
            texels[(size_t)probe * DD + d] = t;
        }
        if (avg_out) {   // this kernel finalises the level: leave the child averages for the level below (far_field)
            const int ld = 31 - __clz(lv.D);   // launch_march only passes avg_out for power-of-two D
            const int dx = (int)(d & (uint32_t)(lv.D - 1)), dy = (int)(d >> ld);
            float4 avg;
            if (child_avg_shfl(t, ystep, dx, dy, avg) && in_range)
                avg_out[(size_t)probe * (DD >> 2) + (size_t)(dy >> 1) * (lv.D >> 1) + (dx >> 1)] = avg;
        }
    }
}

// ------------------------------------------------------------------ march with a block-local ray pool (culled levels >= 1)
// ncu (r2s, level 3 of the 4K frame, split lists): two thirds of k_march's warp-instructions are traversal, executed with 12 of
// 32 lanes — a warp marches 32 fixed rays and idles behind its longest one.  Here a block owns a pool of kPoolRays consecutive
// rays (256 list quads) and works in two phases.  (A) Traversal: a lane that has finished its ray takes the next one of the
// pool (one shared-memory atomicAdd per warp and refill; the warp refills when fewer than `thresh` lanes are busy, Aila &
// Laine's "replace terminated rays") and leaves the closest hit — 16 bytes — in shared memory.  (B) After a barrier the block
// walks the pool in list order exactly like k_march: four consecutive lanes per quad shade, merge, store and form the child
// average.  Which lane traverses a ray cannot change its closest hit (min (t, triangle id) over the same candidate set,
// rc_spec.h S5), so every texel is unchanged.
constexpr int kPoolRays = 1024;

template <bool FUSED, int MINB>
__global__ void __launch_bounds__(128, MINB) k_march_pool(DScene s, DLights L, DLevel lv, int UD, float3 sky, const float4* __restrict__ origin,
                                                          const float4* __restrict__ dirq, uint2* __restrict__ texels,
                                                          const float4* __restrict__ up_avg, const uint4* __restrict__ link_idx,
                                                          const float4* __restrict__ link_w, const int4* __restrict__ entry,
                                                          float4* __restrict__ avg_out, const uint32_t* __restrict__ list,
                                                          const unsigned int* __restrict__ count, const unsigned int* __restrict__ count_triv,
                                                          unsigned list_cap, int thresh)
{
    cudaTriggerProgrammaticLaunchCompletion();
    __shared__ float4 s_hit[kPoolRays];
    __shared__ unsigned s_next;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kDone = kDoneLinkC;
    const unsigned n_real = __ldg(count), n_triv = count_triv ? (__ldg(count_triv) & 0x7fffffffu) : 0u;
    const size_t real_rays = 4 * (size_t)n_real, total = 4 * ((size_t)n_real + n_triv);
    const size_t DD = (size_t)lv.D * lv.D;
    const int ld = 31 - __clz(lv.D);
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    for (size_t b0 = (size_t)blockIdx.x * kPoolRays; b0 < total; b0 += (size_t)gridDim.x * kPoolRays) {
        const unsigned pool = real_rays > b0 ? (unsigned)(real_rays - b0 < (size_t)kPoolRays ? real_rays - b0 : (size_t)kPoolRays) : 0u;
        if (threadIdx.x == 0) s_next = 0u;
        __syncthreads();
        if (pool) {
            // ---- (A) traversal with refill
            int ray = -1, cur = kDone, sp = 0;
            int stack[48];
            float3 o = f3(0.f, 0.f, 0.f), w = o, inv = o, noi = o;
            Hit h; h.t = 0.f; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
            bool exhausted = false;
            for (;;) {
                const unsigned freem = __ballot_sync(FULL, ray < 0);
                if (freem && !exhausted) {
                    const int leader = __ffs((int)freem) - 1;
                    const unsigned want = (unsigned)__popc(freem);
                    unsigned base = 0u;
                    if ((int)lane == leader) base = atomicAdd(&s_next, want);
                    base = __shfl_sync(FULL, base, leader);
                    if (ray < 0) {
                        const unsigned r = base + (unsigned)__popc(freem & lt);
                        if (r < pool) {
                            const size_t g = b0 + r;
                            const uint32_t e = __ldg(list + (g >> 2));
                            const uint32_t q = e & (uint32_t)((DD >> 2) - 1), probe = e >> (2 * ld - 2);
                            const uint32_t x = q & (uint32_t)((lv.D >> 1) - 1), y = q >> (ld - 1);
                            const uint32_t d = ((2u * y + (((uint32_t)g >> 1) & 1u)) << ld) + 2u * x + ((uint32_t)g & 1u);
                            const float4 og = __ldg(origin + probe);
                            const float4 qa = __ldg(dirq + 2 * (size_t)d), qb = __ldg(dirq + 2 * (size_t)d + 1);
                            o = xyz(og); w = xyz(qa); inv = f3(qa.w, qb.x, qb.y);
                            noi = f3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
                            h.t = lv.t1; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
                            sp = 0; cur = 0; ray = (int)r;
                            if (entry) {      // as trace_inv: start at the probe's entry frontier
                                const int4 ea = __ldg(entry + 2 * (size_t)probe), eb = __ldg(entry + 2 * (size_t)probe + 1);
                                if (eb.w != kDone) stack[sp++] = eb.w;
                                if (eb.z != kDone) stack[sp++] = eb.z;
                                if (eb.y != kDone) stack[sp++] = eb.y;
                                if (eb.x != kDone) stack[sp++] = eb.x;
                                if (ea.w != kDone) stack[sp++] = ea.w;
                                if (ea.z != kDone) stack[sp++] = ea.z;
                                if (ea.y != kDone) stack[sp++] = ea.y;
                                cur = ea.x;
                            }
                            if (og.w == 0.0f) cur = kDone;      // (lists hold valid probes only; phase B writes the S7 texel of an invalid one)
                        }
                    }
                    if (base + want >= pool) exhausted = true;
                }
                unsigned act = __ballot_sync(FULL, ray >= 0);
                if (!act) break;
                const int lim = exhausted ? 1 : thresh;
                do {
                    while (cur >= 0) {      // trace_inv's node step
                        const float4* n = s.nodes + 4 * (size_t)cur;
                        const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
                        float ax = fmaf(q0.x, inv.x, noi.x), bx = fmaf(q0.w, inv.x, noi.x);
                        float ay = fmaf(q0.y, inv.y, noi.y), by = fmaf(q1.x, inv.y, noi.y);
                        float az = fmaf(q0.z, inv.z, noi.z), bz = fmaf(q1.y, inv.z, noi.z);
                        const float n0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), lv.t0));
                        const float f0 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));
                        ax = fmaf(q1.z, inv.x, noi.x); bx = fmaf(q2.y, inv.x, noi.x);
                        ay = fmaf(q1.w, inv.y, noi.y); by = fmaf(q2.z, inv.y, noi.y);
                        az = fmaf(q2.x, inv.z, noi.z); bz = fmaf(q2.w, inv.z, noi.z);
                        const float n1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), lv.t0));
                        const float f1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));
                        const bool hit0 = n0 <= f0, hit1 = n1 <= f1;
                        const int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
                        const bool first1 = hit1 && (!hit0 || n1 < n0);
                        const int nearc = first1 ? c1 : c0, farc = first1 ? c0 : c1;
                        if (hit0 && hit1) stack[sp++] = farc;
                        if (hit0 || hit1) cur = nearc;
                        else cur = sp ? stack[--sp] : kDone;
                    }
                    while (cur < 0 && cur != kDone) {
                        const uint32_t leaf = (uint32_t)~cur;
                        const uint32_t first = leaf >> 3, cnt = leaf & 7u;
                        for (uint32_t i = 0; i < cnt; i++) tri_test(s.tri_geom + 3 * (size_t)(first + i), o, w, lv.t0, lv.t1, h);
                        cur = sp ? stack[--sp] : kDone;
                    }
                    if (ray >= 0 && cur == kDone) {
                        s_hit[ray] = make_float4(h.prim == 0xffffffffu ? -1.0f : h.t, h.u, h.v, __uint_as_float(h.prim));
                        ray = -1;
                    }
                    act = __ballot_sync(FULL, ray >= 0);
                } while (__popc(act) >= lim);
            }
        }
        __syncthreads();
        // ---- (B) shade / merge / store / child average, in list order (k_march's epilogue)
#pragma unroll 1
        for (int k = 0; k < kPoolRays / 128; k++) {
            const unsigned r = (unsigned)k * 128u + threadIdx.x;
            const size_t j = b0 + r;
            if (b0 + (size_t)k * 128u >= total) break;      // (uniform over the block)
            const bool in_range = j < total;
            uint32_t probe = 0, d = 0;
            bool certain_miss = false;
            if (in_range) {
                const size_t qi = j >> 2;
                certain_miss = qi >= n_real;
                const uint32_t e = __ldg(list + (certain_miss ? (size_t)list_cap - 1u - (qi - n_real) : qi));
                const uint32_t q = e & (uint32_t)((DD >> 2) - 1);
                const uint32_t x = q & (uint32_t)((lv.D >> 1) - 1), y = q >> (ld - 1);
                probe = e >> (2 * ld - 2);
                d = ((2u * y + (((uint32_t)j >> 1) & 1u)) << ld) + 2u * x + ((uint32_t)j & 1u);
            }
            uint2 t = pack_half4(0.f, 0.f, 0.f, 1.f);   // invalid probe (S7)
            if (in_range) {
                const float4 og = __ldg(origin + probe);
                if (og.w != 0.0f) {
                    const float4 qa = __ldg(dirq + 2 * (size_t)d);
                    const float3 w = xyz(qa), o = xyz(og);
                    Hit h; h.t = -1.0f; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
                    if (!certain_miss) { const float4 hr = s_hit[r]; h.t = hr.x; h.u = hr.y; h.v = hr.z; h.prim = __float_as_uint(hr.w); }
                    t = finalize_texel<FUSED>(s, L, lv, UD, 0, sky, probe, d, o, w, h, up_avg, link_idx, link_w);
                }
                texels[(size_t)probe * DD + d] = t;
            }
            if (avg_out) {
                const int dx = (int)(d & (uint32_t)(lv.D - 1)), dy = (int)(d >> ld);
                float4 avg;
                if (child_avg_shfl(t, 2, dx, dy, avg) && in_range)
                    avg_out[(size_t)probe * (DD >> 2) + (size_t)(dy >> 1) * (lv.D >> 1) + (dx >> 1)] = avg;
            }
        }
        __syncthreads();      // s_hit / s_next are reused by the next pool
    }
}

// ------------------------------------------------------------------ march, one 2x2 quad of texels per thread (culled levels >= 1)
// The ray lists of levels >= 1 hold quads (the four children of one direction of the level below).  k_march gives each
// ray its own thread, so every ray pays the list decode, the probe / link fetches, the merge set-up and three shuffles for
// the child average — ncu (r2e, level 3 of the 4K frame): ~310 of the ~550 thread-instructions a ray costs are such
// per-ray overhead, only ~240 are traversal.  Here a thread owns the whole quad: it decodes once, fetches the probe and
// its link record once, marches the four rays one after the other (same origin, neighbouring directions), stores the
// four texels as two 16-byte vectors and forms the child average in registers.  Same rays, same arithmetic per ray:
// bit-identical texels and averages.
template <bool FUSED, int MINB>
__global__ void __launch_bounds__(64, MINB) k_march_quad(DScene s, DLights L, DLevel lv, int UD, float3 sky, const float4* __restrict__ origin,
                                                        const float4* __restrict__ dirq, uint2* __restrict__ texels, const float4* __restrict__ up_avg,
                                                        const uint4* __restrict__ link_idx, const float4* __restrict__ link_w,
                                                        float4* __restrict__ avg_out, const uint32_t* __restrict__ list,
                                                        const unsigned int* __restrict__ count)
{
    cudaTriggerProgrammaticLaunchCompletion();   // the next level's kernel may begin once every block got here
    const unsigned total = __ldg(count);
    const int ld = 31 - __clz(lv.D);             // culled mode requires power-of-two D
    const uint32_t DD = (uint32_t)lv.D * (uint32_t)lv.D, HD = (uint32_t)lv.D >> 1;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
        const uint32_t e = __ldg(list + j);
        const uint32_t q = e & ((DD >> 2) - 1u), probe = e >> (2 * ld - 2);
        const uint32_t x = q & (HD - 1u), y = q >> (ld - 1);
        const float4 og = __ldg(origin + probe);
        uint2 t[4];
#pragma unroll
        for (int c = 0; c < 4; c++) t[c] = pack_half4(0.f, 0.f, 0.f, 1.f);   // invalid probe (S7)
        if (og.w != 0.0f) {
            const float3 o = xyz(og);
            Hit h[4];
            float3 w[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t d = ((2u * y + (uint32_t)(c >> 1)) << ld) + 2u * x + (uint32_t)(c & 1);
                const float4 qa = __ldg(dirq + 2 * (size_t)d), qb = __ldg(dirq + 2 * (size_t)d + 1);
                w[c] = xyz(qa);
                h[c] = trace_inv(s, o, w[c], f3(qa.w, qb.x, qb.y), lv.t0, lv.t1);
            }
            // S7: radiance of the hits (the float16-rounded raw texel is what S8 reads)
            float4 raw[4];
            bool any_miss = false;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float4 v;
                if (h[c].prim != 0xffffffffu) {
                    const Shade sh = shade_hit(s, L, h[c].prim, h[c].u, h[c].v, vfma(h[c].t, w[c], o), vneg(w[c]));
                    v = make_float4(sh.rad.x, sh.rad.y, sh.rad.z, 0.0f);
                } else {
                    v = make_float4(0.f, 0.f, 0.f, 1.0f);
                    any_miss = true;
                }
                t[c] = pack_half4(v.x, v.y, v.z, v.w);
                raw[c] = unpack_half4(t[c]);
            }
            if (FUSED && any_miss) {
                // S8, once per quad: the link record, then per missed ray the four upper averages of its own direction
                cudaGridDependencySynchronize();
                const uint4 li = __ldg(link_idx + probe);
                const float4 lw = __ldg(link_w + probe);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (raw[c].w == 0.0f) continue;     // a = 0: fma(0, far, raw) = raw exactly
                    const float4 far = far_field(up_avg, lv.D, li, lw, (int)(2u * x) + (c & 1), (int)(2u * y) + (c >> 1), sky, UD < 0);
                    float4 m = make_float4(fmaf(raw[c].w, far.x, raw[c].x), fmaf(raw[c].w, far.y, raw[c].y), fmaf(raw[c].w, far.z, raw[c].z), raw[c].w * far.w);
                    t[c] = pack_half4(fminf(m.x, 65504.0f), fminf(m.y, 65504.0f), fminf(m.z, 65504.0f), m.w);
                }
            }
        }
        uint2* base = texels + (size_t)probe * DD + ((size_t)(2u * y) << ld) + 2u * x;
        *reinterpret_cast<uint4*>(base) = make_uint4(t[0].x, t[0].y, t[1].x, t[1].y);
        *reinterpret_cast<uint4*>(base + lv.D) = make_uint4(t[2].x, t[2].y, t[3].x, t[3].y);
        if (avg_out) {
            const float4 c0 = unpack_half4(t[0]), c1 = unpack_half4(t[1]), c2 = unpack_half4(t[2]), c3 = unpack_half4(t[3]);
            avg_out[(size_t)probe * (DD >> 2) + (size_t)y * HD + x] =
                make_float4(0.25f * (((c0.x + c1.x) + c2.x) + c3.x), 0.25f * (((c0.y + c1.y) + c2.y) + c3.y),
                            0.25f * (((c0.z + c1.z) + c2.z) + c3.z), 0.25f * (((c0.w + c1.w) + c2.w) + c3.w));
        }
    }
}

// ------------------------------------------------------------------ all levels in ONE launch (small frames)
// The ray marches of the N levels are independent of each other — only the merge runs top-down.  On a small
// frame (1080p, open scene: ~0.5 M real rays per level = 2-3 waves of blocks) each per-level kernel spends
// most of its time in its own tail (ncu: 27-39 % of warp slots active against 62.5 % allocated).  Marching
// every level's rays in one grid leaves a single tail; the merges then run as cheap streaming kernels
// (k_merge; fused == separate bit-for-bit, tests/test_gpu_parity.py::test_fused_equals_separate).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_march_all(DScene s, DLights L, DLevelSet ls, MarchPlan plan, float3 sky,
                                                        const float4* __restrict__ origin, const float* __restrict__ dirs,
                                                        uint2* __restrict__ cascade, const int4* __restrict__ entry)
{
    int k = 0;
#pragma unroll 1
    for (int j = 1; j < plan.n; j++) if (blockIdx.x >= plan.block_offset[j]) k = j;
    const int level = plan.level[k];
    const DLevel& lv = ls.lv[level];
    const size_t DD = (size_t)lv.D * lv.D;
    const size_t g = (size_t)(blockIdx.x - plan.block_offset[k]) * blockDim.x + threadIdx.x;
    uint32_t probe, d;
    if (!decode_texel(lv, plan.map[k], g, probe, d)) return;
    uint2* texels = cascade + lv.texel_offset;
    const size_t i = (size_t)probe * DD + d;
    const float4 og = __ldg(origin + lv.probe_offset + probe);
    if (og.w == 0.0f) { texels[i] = pack_half4(0.f, 0.f, 0.f, 1.f); return; }   // invalid probe (S7)
    const float* dl = dirs + plan.dir_offset[k];
    const float3 w = f3(__ldg(dl + 3 * d), __ldg(dl + 3 * d + 1), __ldg(dl + 3 * d + 2));
    const float3 o = xyz(og);
    const Hit h = trace(s, o, w, lv.t0, lv.t1, plan.use_entry[k] ? entry + 2 * ((size_t)lv.probe_offset + probe) : nullptr);
    texels[i] = finalize_texel<false>(s, L, lv, 0, plan.top[k], sky, probe, d, o, w, h, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------ march with block-level ray compaction
// In open scenes most rays of the upper levels leave the scene bounds inside their interval: they need one
// box test and the (uniform) merge, but in k_march they share warps with rays that traverse for hundreds of
// instructions, so traversal code runs with ~10 of 32 lanes active.  Here every thread first tests its ray
// against the root node's two child boxes; the misses finish immediately (merge + store, all such lanes
// together), the survivors are packed — in order, via warp ballots and a block prefix — into the block's
// first warps, which traverse with (nearly) full warps while the other warps have already exited.
__device__ __forceinline__ bool root_overlap(const DScene& s, float3 o, float3 d, float tmin, float tmax)
{
    const float3 inv = f3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    const float3 noi = f3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
    const float4 q0 = __ldg(s.nodes), q1 = __ldg(s.nodes + 1), q2 = __ldg(s.nodes + 2);
    float ax = fmaf(q0.x, inv.x, noi.x), bx = fmaf(q0.w, inv.x, noi.x);
    float ay = fmaf(q0.y, inv.y, noi.y), by = fmaf(q1.x, inv.y, noi.y);
    float az = fmaf(q0.z, inv.z, noi.z), bz = fmaf(q1.y, inv.z, noi.z);
    const float n0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    const float f0 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    ax = fmaf(q1.z, inv.x, noi.x); bx = fmaf(q2.y, inv.x, noi.x);
    ay = fmaf(q1.w, inv.y, noi.y); by = fmaf(q2.z, inv.y, noi.y);
    az = fmaf(q2.x, inv.z, noi.z); bz = fmaf(q2.w, inv.z, noi.z);
    const float n1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    const float f1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    return n0 <= f0 || n1 <= f1;
}

template <bool FUSED, int MINB>
__global__ void __launch_bounds__(128, MINB) k_march_compact(DScene s, DLights L, DLevel lv, int UD, int top, float3 sky, int map,
                                                             const float4* __restrict__ origin, const float* __restrict__ dirs,
                                                             uint2* __restrict__ texels, const float4* __restrict__ up_avg,
                                                             const uint4* __restrict__ link_idx, const float4* __restrict__ link_w)
{
    cudaTriggerProgrammaticLaunchCompletion();
    __shared__ uint2 s_ray[128];      // (probe, direction) of the surviving rays, in thread order
    __shared__ int s_warp[4];
    const size_t DD = (size_t)lv.D * lv.D;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t probe = 0, d = 0;
    bool survive = false;
    Hit miss; miss.t = -1.f; miss.u = 0.f; miss.v = 0.f; miss.prim = 0xffffffffu;
    if (decode_texel(lv, map, (size_t)blockIdx.x * 128 + threadIdx.x, probe, d)) {
        const size_t i = (size_t)probe * DD + d;
        const float4 og = __ldg(origin + probe);
        if (og.w == 0.0f) {
            texels[i] = pack_half4(0.f, 0.f, 0.f, 1.f);   // invalid probe (S7)
        } else {
            const float3 w = f3(__ldg(dirs + 3 * d), __ldg(dirs + 3 * d + 1), __ldg(dirs + 3 * d + 2));
            const float3 o = xyz(og);
            survive = root_overlap(s, o, w, lv.t0, lv.t1);
            if (!survive) texels[i] = finalize_texel<FUSED>(s, L, lv, UD, top, sky, probe, d, o, w, miss, up_avg, link_idx, link_w);
        }
    }
    // ordered compaction of the survivors
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int c = s_warp[k]; if (k < (int)warp) base += c; total += c; }
    if (survive) s_ray[base + __popc(m & ((1u << lane) - 1u))] = make_uint2(probe, d);
    __syncthreads();
    if ((int)threadIdx.x >= total) return;
    const uint2 r = s_ray[threadIdx.x];
    const float4 og = __ldg(origin + r.x);
    const float3 o = xyz(og);
    const float3 w = f3(__ldg(dirs + 3 * r.y), __ldg(dirs + 3 * r.y + 1), __ldg(dirs + 3 * r.y + 2));
    const Hit h = trace(s, o, w, lv.t0, lv.t1);
    texels[(size_t)r.x * DD + r.y] = finalize_texel<FUSED>(s, L, lv, UD, top, sky, r.x, r.y, o, w, h, up_avg, link_idx, link_w);
}

// ------------------------------------------------------------------ persistent march with ray replacement
// One launch per level with a resident grid (148 SMs x blocks/SM).  Lanes fetch rays from a global
// counter; when fewer than `thresh` lanes of a warp are still traversing, the finished lanes shade /
// merge / store together and pull new rays (Aila & Laine's dynamic fetch with lane replacement): a
// warp no longer idles 3/4 of its lanes behind its longest ray (ncu: 7.7-20.7 active lanes of 32 in
// the one-ray-per-thread kernel).  Levels are chained with programmatic dependent launch: level i-1
// starts filling SMs that level i's tail has vacated and only waits (cudaGridDependencySynchronize)
// right before its first read of level i's merged texels.
constexpr int kDone = (int)0x80000000;
// Rays a warp takes from the global counter at a time.  A 1080p level is only ~440 rays per resident
// warp, so the chunk must stay small for the tail to balance (512 made one chunk the critical path).
constexpr unsigned long long kChunk = 64;

template <bool FUSED>
__global__ void __launch_bounds__(128, 8) k_march_persist(DScene s, DLights L, DLevel lv, int UD, int top, float3 sky, int map,
                                                          unsigned long long total, int thresh, unsigned int* __restrict__ counter,
                                                          const float4* __restrict__ origin, const float* __restrict__ dirs,
                                                          uint2* __restrict__ texels, const float4* __restrict__ up_avg,
                                                          const uint4* __restrict__ link_idx, const float4* __restrict__ link_w)
{
    cudaTriggerProgrammaticLaunchCompletion();
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const size_t DD = (size_t)lv.D * lv.D;
    bool have = false, pend = false, exhausted = false, global_done = false, synced = !FUSED;
    unsigned long long wnext = 0, wend = 0;   // warp-uniform chunk of ray indices
    uint32_t probe = 0, d = 0;
    float3 o = f3(0.f, 0.f, 0.f), w = o, inv = o, noi = o;
    Hit h; h.t = 0.f; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
    int sp = 0, cur = kDone;
    int stack[48];
    for (;;) {
        // (1) rays that finished in the last round: shade / merge / store together
        if (__any_sync(FULL, pend)) {
            if (!synced) { cudaGridDependencySynchronize(); synced = true; }   // level i+1 is complete and visible
            if (pend) {
                texels[(size_t)probe * DD + d] = finalize_texel<FUSED>(s, L, lv, UD, top, sky, probe, d, o, w, h, up_avg, link_idx, link_w);
                pend = false;
            }
        }
        // (2) refill free lanes: the warp owns a chunk [wnext, wend) of consecutive rays and takes a new
        //     chunk from the global counter only when it runs dry (one same-address atomic per kChunk rays)
        if (!exhausted) {
            const unsigned freem = __ballot_sync(FULL, !have);
            if (freem) {
                if (wnext >= wend && !global_done) {
                    unsigned long long base = 0;
                    if (lane == 0) base = (unsigned long long)atomicAdd(counter, (unsigned)kChunk) ;
                    base = __shfl_sync(FULL, base, 0);
                    wnext = base;
                    wend = base + kChunk < total ? base + kChunk : total;
                    if (base >= total) { global_done = true; wnext = wend = 0; }
                }
                const unsigned rem = (unsigned)(wend - wnext);
                const unsigned rank = (unsigned)__popc(freem & ((1u << lane) - 1u));
                if (!have && rank < rem) {
                    const unsigned long long g = wnext + rank;
                    if (decode_texel(lv, map, (size_t)g, probe, d)) {
                        const float4 og = __ldg(origin + probe);
                        if (og.w == 0.0f) {
                            texels[(size_t)probe * DD + d] = pack_half4(0.f, 0.f, 0.f, 1.f);   // invalid probe (S7)
                        } else {
                            o = xyz(og);
                            w = f3(__ldg(dirs + 3 * d), __ldg(dirs + 3 * d + 1), __ldg(dirs + 3 * d + 2));
                            inv = f3(safe_inv(w.x), safe_inv(w.y), safe_inv(w.z));
                            noi = f3(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
                            h.t = lv.t1; h.u = 0.f; h.v = 0.f; h.prim = 0xffffffffu;
                            sp = 0; cur = 0; have = true;
                        }
                    }
                }
                const unsigned want = (unsigned)__popc(freem);
                wnext += want < rem ? want : rem;
                if (global_done) exhausted = true;
            }
        }
        // (3) traverse until too few lanes are left busy
        unsigned act = __ballot_sync(FULL, have);
        if (act == 0u) { if (exhausted) break; continue; }
        const int lim = exhausted ? 1 : thresh;
        do {
            if (have) {
                if (cur >= 0) {
                    const float4* n = s.nodes + 4 * (size_t)cur;
                    const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
                    float ax = fmaf(q0.x, inv.x, noi.x), bx = fmaf(q0.w, inv.x, noi.x);
                    float ay = fmaf(q0.y, inv.y, noi.y), by = fmaf(q1.x, inv.y, noi.y);
                    float az = fmaf(q0.z, inv.z, noi.z), bz = fmaf(q1.y, inv.z, noi.z);
                    const float n0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), lv.t0));
                    const float f0 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));
                    ax = fmaf(q1.z, inv.x, noi.x); bx = fmaf(q2.y, inv.x, noi.x);
                    ay = fmaf(q1.w, inv.y, noi.y); by = fmaf(q2.z, inv.y, noi.y);
                    az = fmaf(q2.x, inv.z, noi.z); bz = fmaf(q2.w, inv.z, noi.z);
                    const float n1 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), lv.t0));
                    const float f1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), h.t));
                    const bool hit0 = n0 <= f0, hit1 = n1 <= f1;
                    int c0 = __float_as_int(q3.x), c1 = __float_as_int(q3.y);
                    if (hit0 && hit1) {
                        if (n1 < n0) { const int t = c0; c0 = c1; c1 = t; }
                        stack[sp++] = c1;
                        cur = c0;
                    } else if (hit0) cur = c0;
                    else if (hit1) cur = c1;
                    else cur = sp ? stack[--sp] : kDone;
                }
                if (cur < 0 && cur != kDone) {
                    const uint32_t leaf = (uint32_t)~cur;
                    const uint32_t first = leaf >> 3, cnt = leaf & 7u;
                    for (uint32_t k = 0; k < cnt; k++) tri_test(s.tri_geom + 3 * (size_t)(first + k), o, w, lv.t0, lv.t1, h);
                    cur = sp ? stack[--sp] : kDone;
                }
                if (cur == kDone) { have = false; pend = true; }
            }
            act = __ballot_sync(FULL, have);
        } while (__popc(act) >= lim);
    }
}

// ------------------------------------------------------------------ top level without geometry in range
// With the default intervals the top level starts at 1.33 x the scene diagonal: from a probe on a surface every
// point of such a ray lies outside the scene's bounding box, so the march is a guaranteed miss and the level is
// filled with (sky, 1) — (0,0,0,1) for invalid probes — at streaming speed (128-bit stores, two texels per thread).
__global__ void __launch_bounds__(kBlock) k_fill_top(DLevel lv, float3 sky, const float4* __restrict__ origin, uint4* __restrict__ texels2,
                                                     float4* __restrict__ avg_out)
{
    const size_t half = ((size_t)lv.D * lv.D) >> 1;
    const size_t n2 = (size_t)lv.sw * lv.sh * half;
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n2) return;
    const bool valid = __ldg(origin + (i / half)).w != 0.0f;
    const uint2 t = valid ? pack_half4(sky.x, sky.y, sky.z, 1.0f) : pack_half4(0.f, 0.f, 0.f, 1.0f);
    texels2[i] = make_uint4(t.x, t.y, t.x, t.y);
    // child averages of four equal texels: 0.25*(((c + c) + c) + c) == c exactly (c <= 65504); one per two threads
    if (avg_out && !(i & 1)) avg_out[i >> 1] = unpack_half4(t);
}

// Child averages of a finalised level from its texels — for the paths whose march kernel cannot produce them in
// its epilogue (separate-merge mode, the persistent / compacting / batched variants, probe-tile mapping, odd D).
__global__ void __launch_bounds__(kBlock) k_child_avg(DLevel lv, const uint2* __restrict__ texels, float4* __restrict__ avg_out)
{
    const int HD = lv.D >> 1;
    const size_t per_probe = (size_t)HD * HD;
    const size_t n = (size_t)lv.sw * lv.sh * per_probe;
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const size_t probe = i / per_probe;
    const int r = (int)(i - probe * per_probe), hy = r / HD, hx = r - hy * HD;
    const uint2* base = texels + probe * (size_t)lv.D * lv.D + (size_t)(2 * hy) * lv.D + 2 * hx;
    const uint4 r0 = ld_u4(base), r1 = ld_u4(base + lv.D);
    const float4 c0 = unpack_half4(make_uint2(r0.x, r0.y)), c1 = unpack_half4(make_uint2(r0.z, r0.w));
    const float4 c2 = unpack_half4(make_uint2(r1.x, r1.y)), c3 = unpack_half4(make_uint2(r1.z, r1.w));
    avg_out[i] = make_float4(0.25f * (((c0.x + c1.x) + c2.x) + c3.x), 0.25f * (((c0.y + c1.y) + c2.y) + c3.y),
                             0.25f * (((c0.z + c1.z) + c2.z) + c3.z), 0.25f * (((c0.w + c1.w) + c2.w) + c3.w));
}

// ------------------------------------------------------------------ stand-alone merge (in place)
__global__ void __launch_bounds__(kBlock) k_merge(DLevel lv, int UD, float3 sky, const float4* __restrict__ origin,
                                                  uint2* __restrict__ texels, const float4* __restrict__ up_avg,
                                                  const uint4* __restrict__ link_idx, const float4* __restrict__ link_w)
{
    const size_t DD = (size_t)lv.D * lv.D;
    const size_t n = (size_t)lv.sw * lv.sh * DD;
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const uint32_t probe = (uint32_t)(i / DD), d = (uint32_t)(i - (size_t)probe * DD);
    if (__ldg(origin + probe).w == 0.0f) return;
    const float4 raw = unpack_half4(texels[i]);
    const int dx = (int)(d % (uint32_t)lv.D), dy = (int)(d / (uint32_t)lv.D);
    const float4 far = far_field(up_avg, lv.D, __ldg(link_idx + probe), __ldg(link_w + probe), dx, dy, sky);
    float4 c = make_float4(fmaf(raw.w, far.x, raw.x), fmaf(raw.w, far.y, raw.y), fmaf(raw.w, far.z, raw.z), raw.w * far.w);
    texels[i] = pack_half4(fminf(c.x, 65504.0f), fminf(c.y, 65504.0f), fminf(c.z, 65504.0f), c.w);
}

// ------------------------------------------------------------------ gather (S9)
// S9 for one pixel of the tile from the staged level-0 probes (s_tex / s_org hold probe rows rmin.., columns cmin..,
// `ncols` per row, `stride` uint4 per probe).  Shared by k_gather and k_gather_pipe: the same statements, so the two
// kernels are bit-identical.
template <int DDT>
__device__ __forceinline__ uint2 gather_pixel(const DCamera& cam, const DLevel& l0, const TileRect& tile, int tx, int ty, bool inside, size_t o,
                                              const float* __restrict__ depth, const uint32_t* __restrict__ normal,
                                              const uint4* s_tex, const float4* s_org, const float* s_dirs, int cmin, int rmin, int ncols)
{
    const int DD = DDT ? DDT : l0.D * l0.D, H2 = DD >> 1, stride = H2 + 1;
    uint2 v = make_uint2(0u, 0u);            // pixels without geometry: (0,0,0,0)
    const float dep = inside ? depth[o] : -1.0f;
    if (dep >= 0.0f) {
        const int x = tile.x0 + tx, y = tile.y0 + ty;
        const float3 d = primary_dir(cam, x, y);
        const float3 hp = vfma(dep, d, cam.eye);
        const float3 n = oct_decode(normal[o]);
        int x0, x1, y0, y1;
        float wx0, wx1, wy0, wy1;
        gather_axis(x, l0.P, l0.gw, x0, x1, wx0, wx1);
        gather_axis(y, l0.P, l0.gh, y0, y1, wy0, wy1);
        const int lk[4] = {(y0 - rmin) * ncols + (x0 - cmin), (y0 - rmin) * ncols + (x1 - cmin),
                           (y1 - rmin) * ncols + (x0 - cmin), (y1 - rmin) * ncols + (x1 - cmin)};
        float w[4];
        w[0] = (wx0 * wy0) * plane_weight(n, hp, s_org[lk[0]]);
        w[1] = (wx1 * wy0) * plane_weight(n, hp, s_org[lk[1]]);
        w[2] = (wx0 * wy1) * plane_weight(n, hp, s_org[lk[2]]);
        w[3] = (wx1 * wy1) * plane_weight(n, hp, s_org[lk[3]]);
        const float S = ((w[0] + w[1]) + w[2]) + w[3];
        // S9: one cosine per direction, shared by the four probes; normalised quadrature q = pi / sum(cos)
        float3 acc[4] = {f3(0.f, 0.f, 0.f), f3(0.f, 0.f, 0.f), f3(0.f, 0.f, 0.f), f3(0.f, 0.f, 0.f)};
        float csum = 0.0f;
    #pragma unroll
        for (int di = 0; di < DD; di += 2) {
            const float ca = fmaxf(vdot(n, f3(s_dirs[3 * di], s_dirs[3 * di + 1], s_dirs[3 * di + 2])), 0.0f);
            const float cb = fmaxf(vdot(n, f3(s_dirs[3 * di + 3], s_dirs[3 * di + 4], s_dirs[3 * di + 5])), 0.0f);
            csum = csum + ca;
            csum = csum + cb;
            // both cosines zero (about half of the directions: the lower hemisphere): fma(0, texel, acc) == acc for the
            // finite, non-negative texels a cascade holds, so the pair is skipped — lanes of a warp are neighbouring
            // pixels with similar normals, the branch is mostly uniform
            if (!(ca > 0.0f || cb > 0.0f)) continue;
    #pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint4 r = s_tex[lk[k] * stride + (di >> 1)];
                const float4 a = unpack_half4(make_uint2(r.x, r.y)), b = unpack_half4(make_uint2(r.z, r.w));
                acc[k] = f3(fmaf(ca, a.x, acc[k].x), fmaf(ca, a.y, acc[k].y), fmaf(ca, a.z, acc[k].z));
                acc[k] = f3(fmaf(cb, b.x, acc[k].x), fmaf(cb, b.y, acc[k].y), fmaf(cb, b.z, acc[k].z));
            }
        }
        float3 E = f3(0.f, 0.f, 0.f);
        if (S > 0.0f) {
            const float q = csum > 0.0f ? RC_PI_F / csum : 0.0f;
    #pragma unroll
            for (int k = 0; k < 4; k++) {
                const float wd = (w[k] / S) * q;
                E = f3(fmaf(wd, acc[k].x, E.x), fmaf(wd, acc[k].y, E.y), fmaf(wd, acc[k].z, E.z));
            }
        }
        v = pack_half4(fminf(E.x, 65504.0f), fminf(E.y, 65504.0f), fminf(E.z, 65504.0f), 1.0f);
    }
    return v;
}


// Shared-memory staged gather.  A block owns 32x8 pixels; the <= (32/P0+2) x (8/P0+2) level-0 probes its
// pixels interpolate between are copied once, coalesced (one 128-byte line per probe at D0 = 4), into
// shared memory with a 16-byte pad per probe (so that the 8-9 probes a warp touches fall into different
// banks), then every pixel reads its 4 probes from there.  The direct-from-L1 version spent ~9 L1
// wavefronts per 128-bit load (lanes of a warp touch 9 different lines) and ran at 15 % (1080p) / 7.5 %
// (4K) of the HBM roofline (profiles/r1_a_*).
// DDT = D0*D0 when known at compile time (16 for the default D0 = 4: loops unroll, divisions become shifts), 0 = generic.
template <int DDT>
__global__ void __launch_bounds__(kBlock) k_gather(DCamera cam, DLevel l0, TileRect tile, const float4* __restrict__ origin0,
                                                   const uint2* __restrict__ texels0, const float* __restrict__ dirs0,
                                                   const float* __restrict__ depth, const uint32_t* __restrict__ normal,
                                                   uint2* __restrict__ out, int max_probes,
                                                   unsigned int* __restrict__ counts_in, unsigned int* __restrict__ counts_out,
                                                   PeerOut peer)
{
    extern __shared__ uint4 s_mem[];
    // last kernel of the frame: publish the ray-list lengths to (mapped, pinned) host memory — posted writes,
    // nothing waits for them; the host sizes the next frames' march grids from whatever has arrived
    if (counts_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 3 * RC_MAX_LEVELS) {
        if (threadIdx.x < 2 * RC_MAX_LEVELS) counts_out[threadIdx.x] = counts_in[threadIdx.x];   // (lengths, certain misses; k_split)
        counts_in[threadIdx.x] = 0u;   // the next frame's lists start empty
    }
    const int DD = DDT ? DDT : l0.D * l0.D, H2 = DD >> 1, stride = H2 + 1;
    uint4* s_tex = s_mem;                                                   // [max_probes][stride]
    float4* s_org = reinterpret_cast<float4*>(s_mem + (size_t)max_probes * stride);   // [max_probes]
    float* s_dirs = reinterpret_cast<float*>(s_org + max_probes);          // [DD][3]

    const int X0 = tile.x0 + blockIdx.x * 32, Y0 = tile.y0 + blockIdx.y * 8;
    const int X1 = min(X0 + 31, tile.x0 + tile.w - 1), Y1 = min(Y0 + 7, tile.y0 + tile.h - 1);
    int cmin, cmax, rmin, rmax, unused;
    float fu0, fu1;
    gather_axis(X0, l0.P, l0.gw, cmin, unused, fu0, fu1);
    gather_axis(X1, l0.P, l0.gw, unused, cmax, fu0, fu1);
    gather_axis(Y0, l0.P, l0.gh, rmin, unused, fu0, fu1);
    gather_axis(Y1, l0.P, l0.gh, unused, rmax, fu0, fu1);
    const int ncols = cmax - cmin + 1, nrows = rmax - rmin + 1;
    const uint4* tex4 = reinterpret_cast<const uint4*>(texels0);
    // one warp per probe row: consecutive lanes copy consecutive 16-byte pieces of consecutive probes
    for (int pr = threadIdx.x >> 5; pr < nrows; pr += kBlock / 32) {
        const size_t rowbase = (size_t)(rmin + pr - l0.py0) * l0.sw + (cmin - l0.px0);
        for (int idx = threadIdx.x & 31; idx < ncols * H2; idx += 32) {
            const int pc = idx / H2, j = idx - pc * H2;     // H2 is a compile-time power of two when DDT != 0
            s_tex[(pr * ncols + pc) * stride + j] = ld_u4(tex4 + (rowbase + pc) * H2 + j);
        }
        for (int pc = threadIdx.x & 31; pc < ncols; pc += 32) s_org[pr * ncols + pc] = origin0[rowbase + pc];
    }
    for (int k = threadIdx.x; k < 3 * DD; k += kBlock) s_dirs[k] = dirs0[k];
    __syncthreads();

    const int tx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool inside = tx < tile.w && ty < tile.h;
    const size_t o = inside ? (size_t)ty * tile.w + tx : 0;
    const uint2 v = gather_pixel<DDT>(cam, l0, tile, tx, ty, inside, o, depth, normal, s_tex, s_org, s_dirs, cmin, rmin, ncols);
    if (inside) out[o] = v;
    if (peer.world) {
        // Fused final-image all-gather: the tile is stored straight into every rank's full-frame
        // buffer through NVLink peer mappings, as it is produced (plain stores, nothing waits for them here).
        // k_peer_publish, the next kernel in the stream, raises the "delivered" flags once this grid has retired.
        if (inside) {
            const size_t fo = (size_t)(tile.y0 + ty) * peer.W + (tile.x0 + tx);
            for (int d = 0; d < peer.world; d++) if ((peer.dst_mask >> d) & 1u) peer.frame[d][fo] = v;
        }
    }
}

// ------------------------------------------------------------------ gather, software-pipelined (default D0 = 4)
// k_gather's blocks run load -> barrier -> compute, and the ncu source view of the 4K frame puts 38 % of the kernel's
// stall samples on that load phase (the probes of a tile come from L2 / HBM; four resident blocks per SM do not cover
// it: issue slots are busy 65 % of the time).  Here a block walks a column of `tiles` vertically adjacent 32x8 pixel
// tiles with two shared-memory buffers: the probes of tile j+1 are fetched with cp.async (16 bytes per request,
// no registers, no waiting) while tile j is being computed.  The arithmetic is gather_pixel's: bit-identical output.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct GatherWindow { int cmin, rmin, ncols, nrows; };   // level-0 probes a 32x8 pixel tile interpolates between

__device__ __forceinline__ GatherWindow gather_window(const DLevel& l0, const TileRect& tile, int X0, int Y0)
{
    const int X1 = min(X0 + 31, tile.x0 + tile.w - 1), Y1 = min(Y0 + 7, tile.y0 + tile.h - 1);
    int cmin, cmax, rmin, rmax, unused;
    float fu0, fu1;
    gather_axis(X0, l0.P, l0.gw, cmin, unused, fu0, fu1);
    gather_axis(X1, l0.P, l0.gw, unused, cmax, fu0, fu1);
    gather_axis(Y0, l0.P, l0.gh, rmin, unused, fu0, fu1);
    gather_axis(Y1, l0.P, l0.gh, unused, rmax, fu0, fu1);
    GatherWindow g;
    g.cmin = cmin; g.rmin = rmin; g.ncols = cmax - cmin + 1; g.nrows = rmax - rmin + 1;
    return g;
}

__global__ void __launch_bounds__(kBlock) k_gather_pipe(DCamera cam, DLevel l0, TileRect tile, const float4* __restrict__ origin0,
                                                        const uint2* __restrict__ texels0, const float* __restrict__ dirs0,
                                                        const float* __restrict__ depth, const uint32_t* __restrict__ normal,
                                                        uint2* __restrict__ out, int max_probes, int tiles,
                                                        unsigned int* __restrict__ counts_in, unsigned int* __restrict__ counts_out,
                                                        PeerOut peer)
{
    constexpr int DD = 16, H2 = DD >> 1, stride = H2 + 1;
    extern __shared__ uint4 s_mem[];
    if (counts_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 3 * RC_MAX_LEVELS) {   // see k_gather
        if (threadIdx.x < 2 * RC_MAX_LEVELS) counts_out[threadIdx.x] = counts_in[threadIdx.x];
        counts_in[threadIdx.x] = 0u;
    }
    const size_t buf_u4 = (size_t)max_probes * (stride + 1);            // texels + one float4 origin per probe
    float* s_dirs = reinterpret_cast<float*>(s_mem + 2 * buf_u4);      // [DD][3]
    const uint4* tex4 = reinterpret_cast<const uint4*>(texels0);
    const int X0 = tile.x0 + blockIdx.x * 32;
    const int row0 = blockIdx.y * tiles;                                // first tile row of this block
    const int ntiles = min(tiles, (tile.h + 7) / 8 - row0);             // block-uniform

    // asynchronous copy of tile j's probe window into buffer b (one warp per probe row, as in k_gather)
    auto stage = [&](int j, int b) {
        const GatherWindow g = gather_window(l0, tile, X0, tile.y0 + (row0 + j) * 8);
        uint4* s_tex = s_mem + b * buf_u4;
        float4* s_org = reinterpret_cast<float4*>(s_tex + (size_t)max_probes * stride);
        for (int pr = threadIdx.x >> 5; pr < g.nrows; pr += kBlock / 32) {
            const size_t rowbase = (size_t)(g.rmin + pr - l0.py0) * l0.sw + (g.cmin - l0.px0);
            for (int idx = threadIdx.x & 31; idx < g.ncols * H2; idx += 32) {
                const int pc = idx / H2, k = idx - pc * H2;
                cp_async16(s_tex + (pr * g.ncols + pc) * stride + k, tex4 + (rowbase + pc) * H2 + k);
            }
            for (int pc = threadIdx.x & 31; pc < g.ncols; pc += 32) cp_async16(s_org + pr * g.ncols + pc, origin0 + rowbase + pc);
        }
        cp_async_commit();
    };

    stage(0, 0);
    for (int k = threadIdx.x; k < 3 * DD; k += kBlock) s_dirs[k] = dirs0[k];
    for (int j = 0; j < ntiles; j++) {
        if (j + 1 < ntiles) { stage(j + 1, (j + 1) & 1); cp_async_wait<1>(); }   // tile j has landed, tile j+1 is in flight
        else cp_async_wait<0>();
        __syncthreads();
        const GatherWindow g = gather_window(l0, tile, X0, tile.y0 + (row0 + j) * 8);
        const uint4* s_tex = s_mem + (j & 1) * buf_u4;
        const float4* s_org = reinterpret_cast<const float4*>(s_tex + (size_t)max_probes * stride);
        const int tx = blockIdx.x * 32 + (threadIdx.x & 31);
        const int ty = (row0 + j) * 8 + (threadIdx.x >> 5);
        const bool inside = tx < tile.w && ty < tile.h;
        const size_t o = inside ? (size_t)ty * tile.w + tx : 0;
        const uint2 v = gather_pixel<16>(cam, l0, tile, tx, ty, inside, o, depth, normal, s_tex, s_org, s_dirs, g.cmin, g.rmin, g.ncols);
        if (inside) out[o] = v;
        if (peer.world && inside) {   // fused final-image exchange, see k_gather
            const size_t fo = (size_t)(tile.y0 + ty) * peer.W + (tile.x0 + tx);
            for (int d = 0; d < peer.world; d++) if ((peer.dst_mask >> d) & 1u) peer.frame[d][fo] = v;
        }
        __syncthreads();   // everyone is done with buffer j & 1 before tile j+2 is staged into it
    }
}

// ------------------------------------------------------------------ gather on the tensor cores (default D0 = 4, P0 = 4)
// The 16 pixels x in [4bx+2, 4bx+6), y in [4by+2, 4by+6) of a "cell" interpolate between the same four level-0 probes
// (S1), so S9 for a cell is a small dense contraction:  X_k[pixel][ch] = sum_d cs_d[pixel] * c_k,d[ch]  for each of the
// four probes k, then E[pixel] = sum_k a_k[pixel] * X_k[pixel]  with  a_k = (w_k / S) * (pi / C).  The cascade texels
// already ARE float16, so the inner sum is mma.sync.m16n8k16 (f16 x f16 -> f32) with the probe's 128 bytes used as the
// B operand exactly as they lie in memory:
//   B  rows k' = 0..7:  the probe's 16-byte rows (texel pair 2k', 2k'+1; 8 halves = r g b a r g b a), read with one
//      ldmatrix.x4.trans per cell (the four probes = the four 8x8 matrices);  rows 8..15: the same rows with the halves
//      rotated by four (the value the lane 16 away holds -> one shuffle), so the ODD texel of a pair lands in columns 0..3
//   A  cols 0..7 = cs of the even texels, cols 8..15 = cs of the odd texels, rounded to float16 (the only rounding that
//      differs from the scalar kernels; S10's tolerance covers it: |dE| <= 2^-11 * E)
//   C  columns 0..3 = X_k (r, g, b, a); the thread pair (lane&3) in {0, 1} owns them and finishes E in float32.
// 8 warp-level MMAs + ~40 other instructions replace the 192 FFMA + 128 half->float conversions + 32 LDS.128 that every
// pixel of k_gather spends on the same sums (ncu: 821 warp-instructions per 32 pixels, issue-bound).
// Per-pixel terms (hit point, decoded normal, the four plane weights, the 16 cosines) are computed by the lane that owns
// the pixel (two cells = 32 pixels per warp) and handed to the fragment owners through shared memory.  The per-pixel
// reciprocals (1/(1 + K h^2/l2), 1/S, pi/C) and the ray normalisation use the SFU approximations (2 ulp): none of them
// feeds an index table, and S10 bounds the irradiance error.  The decoded normal and the cosines use the exact expressions
// of S9 — the culling masks (k_gbuffer) are derived from the very same values.
// Staging: the probe window of a block's 8x2 cells (<= 9 x 3 probes, rows contiguous in the probe-major cascade) is
// fetched by ONE thread with cp.async.bulk (TMA, 1-D) into shared memory, completion on an mbarrier, double-buffered
// along the column of tiles the block walks — no per-thread copy instructions, no registers.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RC_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RC_MBAR_DONE;\n"
        "bra RC_MBAR_WAIT;\n"
        "RC_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi)
{
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct Dirs16 { float v[48]; };   // level-0 directions, by value in the kernel parameter block (constant bank)

constexpr int kGmCellsX = 8, kGmCellsY = 2;                        // cells per block tile: 32 x 8 pixels, 8 warps x 2 cells
constexpr int kGmCols = kGmCellsX + 1, kGmRows = kGmCellsY + 1;    // probe window of a tile

struct GatherMmaSmem {
    uint4 tex[2][kGmRows * kGmCols * 8];      // [buffer][probe][8 rows of 16 bytes]
    float4 org[2][kGmRows * kGmCols];
    // A operands, one 16-byte record per (warp, cell, fragment row g, lane-in-quad j), stored at slot j ^ ((g >> 1) & 3):
    //   .x = cs halves (4j, 4j+2) of pixel row g, .y = the same of row g + 8, .z / .w = halves (4j+1, 4j+3) of rows g / g + 8
    // — exactly the four registers mma.m16n8k16 wants, so the fragment owner fetches them with one LDS.128
    uint4 afrag[8 * 2 * 8 * 4];
    float4 a[256];                            // a_k = (w_k / S) * (pi / C) per pixel
    float alpha[256];
    uint2 outt[8 * 32];                       // the step's 32 x 8 pixel tile of results, row-major, for coalesced 16-byte stores
    unsigned long long bar[2];
};

template <bool SYM>
__global__ void __launch_bounds__(256) k_gather_mma(DCamera cam, DLevel l0, TileRect tile, Dirs16 dirs, const float* __restrict__ axis_nx,
                                                    const float* __restrict__ axis_ny, const float4* __restrict__ origin0,
                                                    const uint2* __restrict__ texels0, const float* __restrict__ depth,
                                                    const uint32_t* __restrict__ normal, uint2* __restrict__ out, int cbx0, int cby0,
                                                    int nsteps_total, int steps_per_block, unsigned int* __restrict__ counts_in,
                                                    unsigned int* __restrict__ counts_out, PeerOut peer)
{
    extern __shared__ __align__(128) unsigned char gm_raw[];
    GatherMmaSmem& sm = *reinterpret_cast<GatherMmaSmem*>(gm_raw);
    if (counts_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 3 * RC_MAX_LEVELS) {   // see k_gather
        if (threadIdx.x < 2 * RC_MAX_LEVELS) counts_out[threadIdx.x] = counts_in[threadIdx.x];
        counts_in[threadIdx.x] = 0u;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bxB = cbx0 + blockIdx.x * kGmCellsX;                           // first cell column of this block
    const int step0 = blockIdx.y * steps_per_block;
    const int nsteps = min(steps_per_block, nsteps_total - step0);          // block-uniform
    // probes the context holds (a multi-GPU tile holds a sub-grid; every probe a pixel of the tile needs is inside it)
    const int cxlo = max(0, l0.px0), cxhi = min(l0.gw, l0.px0 + l0.sw) - 1;
    const int cylo = max(0, l0.py0), cyhi = min(l0.gh, l0.py0 + l0.sh) - 1;
    auto clampx = [&](int v) { return min(max(v, cxlo), cxhi); };
    auto clampy = [&](int v) { return min(max(v, cylo), cyhi); };
    const int pc_lo = clampx(bxB), ncols = clampx(bxB + kGmCellsX) - pc_lo + 1;

    auto issue = [&](int s) {   // one thread: TMA bulk copies of step s's probe rows into buffer s & 1
        const int byS = cby0 + (step0 + s) * kGmCellsY;
        const int pr_lo = clampy(byS), nrows = clampy(byS + kGmCellsY) - pr_lo + 1;
        const int b = s & 1;
        mbar_expect_tx(&sm.bar[b], (unsigned)(nrows * ncols * 144));
        for (int r = 0; r < nrows; r++) {
            const size_t p0 = (size_t)(pr_lo + r - l0.py0) * l0.sw + (pc_lo - l0.px0);
            bulk_g2s(&sm.tex[b][r * kGmCols * 8], texels0 + p0 * 16, (unsigned)(ncols * 128), &sm.bar[b]);
            bulk_g2s(&sm.org[b][r * kGmCols], origin0 + p0, (unsigned)(ncols * 16), &sm.bar[b]);
        }
    };
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && nsteps > 0) issue(0);

    // ---- per-thread invariants
    // phase 1: this lane owns pixel (mx, my) of cell c of the warp's cell pair
    const int c = lane >> 4, m = lane & 15, mx = m & 3, my = m >> 2;
    const int bx_own = bxB + 2 * (warp & 3) + c;
    const int x = 4 * bx_own + 2 + mx;
    const bool x_in = x >= tile.x0 && x < tile.x0 + tile.w;
    const float nx = x_in ? __ldg(axis_nx + x) : 0.0f;
    const float fx = (float)mx * 0.25f, fy = (float)my * 0.25f;     // S1: f = ((x - P/2) mod P) / P, exact
    const float bil[4] = {(1.0f - fx) * (1.0f - fy), fx * (1.0f - fy), (1.0f - fx) * fy, fx * fy};
    const int lx0 = clampx(bx_own) - pc_lo, lx1 = clampx(bx_own + 1) - pc_lo;
    const int wrow = warp >> 2;                                      // cell row of this warp inside the tile
    // the A records this lane WRITES: pixel row (gw = m & 7, hw = m >> 3) of cell c; slot j' = j ^ ((gw >> 1) & 3)
    const int gw = m & 7, hw = m >> 3;
    uint32_t* const afw = reinterpret_cast<uint32_t*>(&sm.afrag[((warp * 2 + c) * 8 + gw) * 4]) + hw;
    const int swz_w = (gw >> 1) & 3;
    // phase 2: fragment row g, quad lane j
    const int g = lane >> 2, j = lane & 3, q = lane >> 3;
    const int rec = warp * 32;
    const uint4* const afr0 = &sm.afrag[((warp * 2 + 0) * 8 + g) * 4 + (j ^ ((g >> 1) & 3))];
    const uint4* const afr1 = &sm.afrag[((warp * 2 + 1) * 8 + g) * 4 + (j ^ ((g >> 1) & 3))];
    const int lxq0 = clampx(bxB + 2 * (warp & 3) + (q & 1)) - pc_lo, lxq1 = clampx(bxB + 2 * (warp & 3) + 1 + (q & 1)) - pc_lo;

    // prefetch of the G-buffer of the lane's pixel, one step ahead
    auto gload = [&](int s, float& dep, uint32_t& nrm, int& y) {
        y = 4 * (cby0 + (step0 + s) * kGmCellsY + wrow) + 2 + my;
        dep = -1.0f; nrm = 0u;
        if (x_in && y >= tile.y0 && y < tile.y0 + tile.h) {
            const size_t o = (size_t)(y - tile.y0) * tile.w + (x - tile.x0);
            dep = __ldg(depth + o);
            nrm = __ldg(normal + o);
        }
    };
    float dep_n = -1.0f; uint32_t nrm_n = 0u; int y_n = 0;
    if (nsteps > 0) gload(0, dep_n, nrm_n, y_n);

    for (int s = 0; s < nsteps; s++) {
        if (threadIdx.x == 0 && s + 1 < nsteps) issue(s + 1);       // buffer (s+1)&1 was released by the barrier that ended step s-1
        const int b = s & 1;
        const int byS = cby0 + (step0 + s) * kGmCellsY, by = byS + wrow;
        const int pr_lo = clampy(byS);
        const int ly0 = clampy(by) - pr_lo, ly1 = clampy(by + 1) - pr_lo;
        const float dep = dep_n; const uint32_t nrm = nrm_n; const int y = y_n;
        if (s + 1 < nsteps) gload(s + 1, dep_n, nrm_n, y_n);
        // A warp whose 32 pixels all lack geometry (31-41 % of the pixels of living_room's orbit, ~78 % of the teapot's and of sonic's,
        // nearly all of them in such warps) has nothing to compute: S9 gives (0,0,0,0), which is also what the contraction would
        // produce from a_k = 0.  It leaves its 8 x 4 pixels of the output tile empty and waits for the others.
        const bool warp_has_geometry = __any_sync(0xffffffffu, dep >= 0.0f);
        // (warp 0 always waits: its thread 0 re-arms this barrier two steps later and must have seen this phase complete)
        if (warp == 0 || warp_has_geometry) mbar_wait(&sm.bar[b], (unsigned)((s >> 1) & 1));
        if (!warp_has_geometry) {
            sm.outt[(4 * wrow + (lane >> 3)) * 32 + 8 * (warp & 3) + (lane & 7)] = make_uint2(0u, 0u);
        } else {
        const uint4* s_tex = sm.tex[b];
        const float4* s_org = sm.org[b];

        // ---- phase 1: per-pixel terms, one lane per pixel
        {
            float cs[16];
            float4 ak = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int d = 0; d < 16; d++) cs[d] = 0.0f;
            if (dep >= 0.0f) {
                const float ny = __ldg(axis_ny + y);
                const float3 qd = f3(fmaf(nx, cam.dx.x, fmaf(ny, cam.dy.x, cam.dc.x)), fmaf(nx, cam.dx.y, fmaf(ny, cam.dy.y, cam.dc.y)),
                                     fmaf(nx, cam.dx.z, fmaf(ny, cam.dy.z, cam.dc.z)));
                const float rl = rsqrtf(vdot(qd, qd));
                const float3 hp = vfma(dep * rl, qd, cam.eye);
                const float3 n = oct_decode(nrm);
                const int lk[4] = {ly0 * kGmCols + lx0, ly0 * kGmCols + lx1, ly1 * kGmCols + lx0, ly1 * kGmCols + lx1};
                float w[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float4 ok = s_org[lk[k]];
                    const float3 delta = vsub(xyz(ok), hp);
                    const float l2 = vdot(delta, delta), hh = vdot(n, delta);
                    // S8: 1 / (1 + K h^2 / l2) == l2 / (l2 + K h^2)
                    const float gk = l2 > 0.0f ? l2 * fast_rcp(fmaf(RC_PLANE_K * hh, hh, l2)) : 1.0f;
                    w[k] = ok.w != 0.0f ? bil[k] * gk : 0.0f;
                }
                const float S = ((w[0] + w[1]) + w[2]) + w[3];
                if (SYM) {
                    // the S3 table for D = 4 is point-symmetric: w_d' = -w_d for the pairs below (checked on the host), and
                    // dot(n, -w) == -dot(n, w) exactly, so 8 dot products give all 16 cosines
#define RC_CS_PAIR(A, B)                                                                                          \
    {                                                                                                             \
        const float dt = vdot(n, f3(dirs.v[3 * (A)], dirs.v[3 * (A) + 1], dirs.v[3 * (A) + 2]));                   \
        cs[A] = fmaxf(dt, 0.0f);                                                                                  \
        cs[B] = fmaxf(-dt, 0.0f);                                                                                 \
    }
                    RC_CS_PAIR(5, 15) RC_CS_PAIR(6, 12) RC_CS_PAIR(9, 3) RC_CS_PAIR(10, 0)
                    RC_CS_PAIR(1, 14) RC_CS_PAIR(2, 13) RC_CS_PAIR(4, 11) RC_CS_PAIR(7, 8)
#undef RC_CS_PAIR
                } else {
#pragma unroll
                    for (int d = 0; d < 16; d++) cs[d] = fmaxf(vdot(n, f3(dirs.v[3 * d], dirs.v[3 * d + 1], dirs.v[3 * d + 2])), 0.0f);
                }
                float C = 0.0f;
#pragma unroll
                for (int d = 0; d < 16; d++) C = C + cs[d];
                const float SC = S * C;
                const float qs = SC > 0.0f ? RC_PI_F * fast_rcp(SC) : 0.0f;      // (1/S) * (pi/C)
                ak = make_float4(w[0] * qs, w[1] * qs, w[2] * qs, w[3] * qs);
            }
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                uint32_t* dst = afw + 4 * (jj ^ swz_w);
                dst[0] = pack_h2(cs[4 * jj], cs[4 * jj + 2]);           // even texels of the pairs 2jj, 2jj+1
                dst[2] = pack_h2(cs[4 * jj + 1], cs[4 * jj + 3]);       // odd texels
            }
            sm.a[rec + lane] = ak;
            sm.alpha[rec + lane] = dep >= 0.0f ? 1.0f : 0.0f;
        }
        __syncwarp();

        // ---- phase 2: two cells per warp, four MMAs each
        const int lyq = (q >> 1 ? ly1 : ly0);
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const uint4 A = cc ? *afr1 : *afr0;
            const int r0 = rec + cc * 16 + g;                        // pixel records of fragment rows g (and g + 8)
            const float4 a0 = sm.a[r0], a1 = sm.a[r0 + 8];
            // ldmatrix: lane l supplies the address of row (l & 7) of matrix (l >> 3) = probe (l >> 3) of the cell
            const unsigned addr = smem_u32(&s_tex[(lyq * kGmCols + (cc ? lxq1 : lxq0)) * 8 + (lane & 7)]);
            uint32_t bq[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(bq[0]), "=r"(bq[1]), "=r"(bq[2]), "=r"(bq[3]) : "r"(addr));
            float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
            const float wa[4] = {a0.x, a0.y, a0.z, a0.w}, wb[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t brot = __shfl_xor_sync(0xffffffffu, bq[k], 16);   // the same rows, halves rotated by four
                float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
                             : "r"(A.x), "r"(A.y), "r"(A.z), "r"(A.w), "r"(bq[k]), "r"(brot));
                e0 = fmaf(wa[k], c0, e0); e1 = fmaf(wa[k], c1, e1);
                e2 = fmaf(wb[k], c2, e2); e3 = fmaf(wb[k], c3, e3);
            }
            // ---- phase 3: lanes j = 0 (r, g) and j = 1 (b, alpha) put fragment rows g and g + 8 into the block's output tile
            if (j < 2) {
                const int tx = 8 * (warp & 3) + 4 * cc + (g & 3);       // pixel column inside the 32 x 8 tile
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float va = h ? e2 : e0, vb = h ? e3 : e1;
                    // pixels without geometry have a_k = 0 -> E = 0 and alpha 0: (0,0,0,0) as S9 demands
                    const float second = j == 0 ? fminf(vb, 65504.0f) : sm.alpha[r0 + 8 * h];
                    const int ty = 4 * wrow + (g >> 2) + 2 * h;
                    reinterpret_cast<uint32_t*>(&sm.outt[ty * 32 + tx])[j] = pack_h2(fminf(va, 65504.0f), second);
                }
            }
        }
        }      // warp_has_geometry
        __syncthreads();
        // ---- phase 4: the finished 32 x 8 tile leaves the block as 16-byte vectors (two pixels), row segments contiguous:
        // full sectors for the local store and for the NVLink stores of the fused final-image exchange (4-byte stores from
        // the fragment owners made the peer path slower than the whole single-GPU gather: 0.20 vs 0.157 ms per half frame)
        if (threadIdx.x < 128) {
            const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 2;
            const int px = 4 * bxB + 2 + tx - tile.x0, py = 4 * byS + 2 + ty - tile.y0;      // tile coordinates of the pixel pair
            if (py >= 0 && py < tile.h && px + 1 >= 0 && px < tile.w) {
                const uint2 v0 = sm.outt[ty * 32 + tx], v1 = sm.outt[ty * 32 + tx + 1];
                const bool in0 = px >= 0, in1 = px + 1 < tile.w;
                const size_t o = (size_t)py * tile.w + (in0 ? px : px + 1);
                // pixel pairs start at even tile columns + 2: 16-byte alignment holds when the row pitch and tile.x0 are even
                const bool vec = in0 && in1 && !(tile.w & 1) && !((px + tile.x0) & 1) && !(tile.x0 & 1);
                if (vec) *reinterpret_cast<uint4*>(out + o) = make_uint4(v0.x, v0.y, v1.x, v1.y);
                else { if (in0) out[o] = v0; if (in1) out[o + (in0 ? 1 : 0)] = v1; }
                if (peer.world) {   // fused final-image exchange, see k_gather
                    const size_t fo = (size_t)(py + tile.y0) * peer.W + ((in0 ? px : px + 1) + tile.x0);
                    const bool pvec = in0 && in1 && !(peer.W & 1) && !((px + tile.x0) & 1);
                    for (int d = 0; d < peer.world; d++) {
                        if (!((peer.dst_mask >> d) & 1u)) continue;
                        if (pvec) *reinterpret_cast<uint4*>(peer.frame[d] + fo) = make_uint4(v0.x, v0.y, v1.x, v1.y);
                        else { if (in0) peer.frame[d][fo] = v0; if (in1) peer.frame[d][fo + (in0 ? 1 : 0)] = v1; }
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with buffer s & 1 (step s + 2 overwrites it) and with its pixel records
    }
}

// ------------------------------------------------------------------ peer-memory frame exchange (tiled multi-GPU)
// Control block of a rank (uint32 words, in its IPC-shared allocation): [kPeerArrived + r] = last frame rank r has
// delivered into THIS rank's frame buffers; [kPeerReleased + r] = last frame rank r has finished consuming (so that
// its buffer slot may be overwritten); [kPeerError] = number of timed-out waits.  Frames alternate between two
// buffer slots, so a producer may run at most one frame ahead of the slowest consumer.
__device__ __forceinline__ bool peer_spin(volatile uint32_t* flag, uint32_t want)
{
    const long long t0 = clock64();
    while ((int)(*flag - want) < 0) {
        if (clock64() - t0 > 4000000000ll) return false;   // ~2 s: never hang the GPU on a lost peer
        __nanosleep(200);
    }
    return true;
}

// before this rank's gather of frame `seq`: tell every rank that this rank is done reading frame seq-1 (stream
// order guarantees it), then wait until every rank has released frame seq-2, whose slot frame `seq` overwrites
__global__ void k_peer_begin(PeerOut peer, uint32_t* my_ctrl)
{
    const int r = threadIdx.x;
    if (r >= peer.world) return;
    __threadfence_system();
    *((volatile uint32_t*)(peer.ctrl[r] + kPeerReleased + peer.rank)) = peer.seq - 1u;
    // only the ranks that receive frames hold (and release) buffer slots
    if (((peer.dst_mask >> r) & 1u) && peer.seq >= 2u && !peer_spin(my_ctrl + kPeerReleased + r, peer.seq - 2u)) atomicAdd(my_ctrl + kPeerError, 1u);
}

// after this rank's gather of frame `seq` (a kernel boundary: all of its peer stores have been performed): tell
// every rank that this rank's tile has been delivered
__global__ void k_peer_publish(PeerOut peer)
{
    const int r = threadIdx.x;
    if (r >= peer.world) return;
    __threadfence_system();
    if ((peer.dst_mask >> r) & 1u) *((volatile uint32_t*)(peer.ctrl[r] + kPeerArrived + peer.rank)) = peer.seq;
}

// consumer side: every rank's tile of frame `seq` has landed in this rank's frame buffer
__global__ void k_peer_wait(int world, uint32_t seq, uint32_t* my_ctrl)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    if (!peer_spin(my_ctrl + kPeerArrived + r, seq)) atomicAdd(my_ctrl + kPeerError, 1u);
    __threadfence_system();
}

// ------------------------------------------------------------------ read-back by SMs (rc_read_target_async, optional)
// Streams a finished target into page-locked host memory with plain 16-byte stores (posted PCIe writes) from a
// few resident blocks, as an alternative to the copy engine (rc_set_tuning "copy_blocks").
__global__ void __launch_bounds__(kBlock) k_copy_to_host(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n16; i += (size_t)gridDim.x * kBlock) dst[i] = ld_u4(src + i);
}

// ------------------------------------------------------------------ 6-byte read-back format of the irradiance
// RGBA16F spends a quarter of the frame's PCIe bytes on alpha, which is only the coverage flag (1 where geometry, else 0) —
// and at 4K the 66 MB read-back (1.3 ms) is what bounds the end-to-end frame rate, not the 1.3 ms of rendering.  RGB48:
// three float16 per pixel, bit-exact r, g, b (E >= 0, so the sign bit of r is free) with the sign bit of r SET where the
// pixel has no geometry.  Eight pixels (64 bytes in, 48 bytes out) per thread, 16-byte vectors both ways.
__global__ void __launch_bounds__(kBlock) k_pack_rgb48(size_t n8, size_t n, const uint4* irr2, uint4* out3, const uint2* irr, uint16_t* out16)
{
    const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
    if (i < n8) {
        uint32_t h[16];                       // 8 pixels x (rg, ba)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 v = ld_u4(irr2 + 4 * i + k);
            h[4 * k] = v.x; h[4 * k + 1] = v.y; h[4 * k + 2] = v.z; h[4 * k + 3] = v.w;
        }
        uint16_t o[24];
#pragma unroll
        for (int px = 0; px < 8; px++) {
            const uint32_t rg = h[2 * px], ba = h[2 * px + 1];
            const bool covered = (ba >> 16) != 0u;          // alpha is 1.0 or 0.0
            o[3 * px] = (uint16_t)((rg & 0x7fffu) | (covered ? 0u : 0x8000u));
            o[3 * px + 1] = (uint16_t)(rg >> 16);
            o[3 * px + 2] = (uint16_t)(ba & 0xffffu);
        }
#pragma unroll
        for (int k = 0; k < 3; k++)
            out3[3 * i + k] = make_uint4((uint32_t)o[8 * k] | ((uint32_t)o[8 * k + 1] << 16), (uint32_t)o[8 * k + 2] | ((uint32_t)o[8 * k + 3] << 16),
                                         (uint32_t)o[8 * k + 4] | ((uint32_t)o[8 * k + 5] << 16), (uint32_t)o[8 * k + 6] | ((uint32_t)o[8 * k + 7] << 16));
    } else if (i == n8) {                     // the last n % 8 pixels
        for (size_t p = 8 * n8; p < n; p++) {
            const uint2 v = irr[p];
            out16[3 * p] = (uint16_t)((v.x & 0x7fffu) | ((v.y >> 16) ? 0u : 0x8000u));
            out16[3 * p + 1] = (uint16_t)(v.x >> 16);
            out16[3 * p + 2] = (uint16_t)(v.y & 0xffffu);
        }
    }
}

// ------------------------------------------------------------------ display composite (outside the hot path)
__device__ __forceinline__ unsigned char srgb8(float x)
{
    x = fminf(fmaxf(x, 0.0f), 1.0f);
    const float e = x <= 0.0031308f ? x * 12.92f : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
    return (unsigned char)__float2int_rd(e * 255.0f + 0.5f);
}

__global__ void __launch_bounds__(kBlock) k_composite(int n, const uint2* __restrict__ irr, const uint2* __restrict__ albedo,
                                                      const uint2* __restrict__ direct, uchar4* __restrict__ composite,
                                                      uchar4* __restrict__ direct_srgb)
{
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float4 e = unpack_half4(irr[i]), a = unpack_half4(albedo[i]), d = unpack_half4(direct[i]);
    const float ip = 1.0f / RC_PI_F;
    // Bgra8UnormSrgb (src/window/app.rs:59-75); clear colour (0,0,0,1) where no geometry (src/renderer.rs:573-578)
    const float r = fmaf(a.x * ip, e.x, d.x), g = fmaf(a.y * ip, e.y, d.y), b = fmaf(a.z * ip, e.z, d.z);
    composite[i] = make_uchar4(srgb8(b), srgb8(g), srgb8(r), 255);
    direct_srgb[i] = make_uchar4(srgb8(d.z), srgb8(d.y), srgb8(d.x), 255);
}

// ------------------------------------------------------------------ debug / parity entry points
__global__ void __launch_bounds__(kBlock) k_trace_rays(DScene s, const float* __restrict__ rays, uint32_t n, float* __restrict__ hits)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float* r = rays + 8 * (size_t)i;
    const Hit h = trace(s, f3(r[0], r[1], r[2]), f3(r[4], r[5], r[6]), r[3], r[7]);
    float* o = hits + 4 * (size_t)i;
    o[0] = h.t; o[1] = h.u; o[2] = h.v; o[3] = __uint_as_float(h.prim);
}

__global__ void __launch_bounds__(kBlock) k_shade_points(DScene s, DLights L, const float* __restrict__ in, uint32_t n, float* __restrict__ out)
{
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float* r = in + 8 * (size_t)i;
    const uint32_t prim = __float_as_uint(r[0]);
    const float u = r[1], v = r[2];
    const float3 eye = f3(r[4], r[5], r[6]);
    const float* a = s.verts + 17 * (size_t)s.tris[3 * (size_t)prim];
    const float* b = s.verts + 17 * (size_t)s.tris[3 * (size_t)prim + 1];
    const float* c = s.verts + 17 * (size_t)s.tris[3 * (size_t)prim + 2];
    const float3 P = f3(lerp3(a[0], b[0], c[0], u, v), lerp3(a[1], b[1], c[1], u, v), lerp3(a[2], b[2], c[2], u, v));
    const Shade sh = shade_hit(s, L, prim, u, v, P, vnormalize(vsub(eye, P)));
    out[4 * (size_t)i] = sh.rad.x; out[4 * (size_t)i + 1] = sh.rad.y; out[4 * (size_t)i + 2] = sh.rad.z; out[4 * (size_t)i + 3] = 0.f;
}

inline unsigned blocks_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace

void launch_gbuffer(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, GBufferOut out, int DD0,
                    const float* dirs0, uint16_t* pixmask, cudaStream_t st)
{
    dim3 grid((tile.w + 31) / 32, (tile.h + 7) / 8);
    k_gbuffer<<<grid, kBlock, 0, st>>>(s, cam, L, tile, out, DD0, dirs0, DD0 <= 16 ? pixmask : nullptr);
}

size_t bin_tiles(TileRect tile) { return (size_t)((tile.w + kBinTile - 1) / kBinTile) * ((tile.h + kBinTile - 1) / kBinTile); }
size_t bin_list_entries(TileRect tile) { return bin_tiles(tile) * kBinCap; }

size_t bin_huge_bytes(uint32_t n_leaf_tris) { return (size_t)(n_leaf_tris ? n_leaf_tris : 1) * sizeof(HugeTri); }

void launch_gbuffer_binned(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, GBufferOut out, int DD0,
                           const float* dirs0, uint16_t* pixmask, uint32_t n_leaf_tris, unsigned int* bin_count, uint32_t* bin_lists,
                           unsigned int* huge_count, void* huge, cudaStream_t st)
{
    const int ntx = (tile.w + kBinTile - 1) / kBinTile, nty = (tile.h + kBinTile - 1) / kBinTile;
    cudaMemsetAsync(huge_count, 0, sizeof(unsigned int), st);
    if (n_leaf_tris) k_bin<<<(n_leaf_tris + kBinBlock - 1) / kBinBlock, kBinBlock, 0, st>>>(s, cam, tile, n_leaf_tris, ntx, bin_count, bin_lists, huge_count, (HugeTri*)huge);
    k_gbuffer_binned<<<dim3(ntx, nty), kBlock, 0, st>>>(s, cam, L, tile, out, DD0, dirs0, DD0 <= 16 ? pixmask : nullptr, ntx, bin_count, bin_lists,
                                                        huge_count, (const HugeTri*)huge);
}

void launch_direct(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, const float* depth, const uint32_t* prim,
                   const float2* bary, uint2* albedo, uint2* direct, cudaStream_t st)
{
    dim3 grid((tile.w + 31) / 32, (tile.h + 7) / 8);
    k_direct<<<grid, kBlock, 0, st>>>(s, cam, L, tile, depth, prim, bary, albedo, direct);
}

void launch_probes(const DScene& s, const DCamera& cam, const DLevelSet& ls, unsigned total, TileRect tile, float offset,
                   const float* depth, const uint32_t* prim, float4* origin, float4* normal, const uint16_t* pixmask,
                   uint32_t* need0, const uint32_t* occ, uint32_t frame, int ow, bool floating, cudaStream_t st)
{
    // floating probes: levels 0..2 a thread per probe, from level 3 up a warp per probe (k_probes).  The split is rounded down to
    // a multiple of 32 so that a warp never mixes the two modes (the few probes of level 2 behind it simply get a warp each too)
    unsigned thread_probes = (floating && ls.n > 3) ? ls.lv[3].probe_offset : total;
    if (thread_probes != total) thread_probes &= ~31u;
    const size_t threads = (size_t)thread_probes + (size_t)(total - thread_probes) * 32;
    if (floating)
        k_probes<true><<<blocks_for(threads), kBlock, 0, st>>>(s, cam, ls, total, thread_probes, tile, offset, depth, prim, origin, normal, pixmask,
                                                               need0, occ, frame, ow);
    else
        k_probes<false><<<blocks_for(threads), kBlock, 0, st>>>(s, cam, ls, total, thread_probes, tile, offset, depth, prim, origin, normal, pixmask,
                                                                need0, occ, frame, ow);
}

// lane distance of a texel's +dy neighbour inside the warp, or 0 when the 2x2 children of a lower direction do
// not share a warp under this mapping (then the caller runs launch_child_avg after the level)
int march_avg_ystep(int D, int map)
{
    if (D < 2 || (D & (D - 1))) return 0;
    if (map == MAP_DIR_TILE && !(D & 7)) return 8;
    if (map == MAP_LINEAR && D <= 16) return D;
    return 0;
}

void launch_link_entry(const DScene& s, const DLevelSet& ls, unsigned link_total, int entry_levels, const float4* origin,
                       const float4* normal, uint4* link_idx, float4* link_w, int4* entry, cudaStream_t st)
{
    EntryPlan plan{};
    plan.n = entry_levels;
    unsigned groups = 0;
    for (int l = 0; l < entry_levels; l++) {
        const DLevel& lv = ls.lv[l];
        plan.g[l] = l < 2 ? 2 : 1;
        plan.group_offset[l] = groups;
        groups += (unsigned)(((lv.sw + plan.g[l] - 1) / plan.g[l]) * ((lv.sh + plan.g[l] - 1) / plan.g[l]));
    }
    plan.group_offset[entry_levels] = groups;
    const unsigned lb = blocks_for(link_total), eb = blocks_for(groups);
    if (lb + eb) k_link_entry<<<lb + eb, kBlock, 0, st>>>(s, ls, plan, link_total, lb, origin, normal, link_idx, link_w, entry);
}

void launch_march(const DScene& s, const DLights& L, const DLevel& lv, const DLevel* up, bool top, float3 sky,
                  const float4* origin, const float* dirs, const float4* dirq, uint2* texels, const float4* up_avg,
                  const uint4* link_idx, const float4* link_w, const int4* entry, float4* avg_out, bool fused, int map, int occ, bool pdl,
                  bool compact, int max_blocks, const uint32_t* list, const unsigned int* count, int quad, bool up_const,
                  const unsigned int* count_triv, unsigned list_cap, cudaStream_t st)
{
    const int block = 128;
    if (map == MAP_DIR_TILE && (lv.D & 7)) map = MAP_LINEAR;   // the 8x4 direction tile needs D % 8 == 0
    int ystep = march_avg_ystep(lv.D, map);
    if (list) { map = MAP_LINEAR; ystep = quad ? 2 : 0; compact = false; }   // max_blocks: the caller's estimate of the list length
    if (!ystep || compact) avg_out = nullptr;
    size_t n = (size_t)lv.sw * lv.sh * lv.D * lv.D;
    if (map == MAP_PROBE_TILE) n = (size_t)((lv.sw + 7) / 8) * ((lv.sh + 3) / 4) * lv.D * lv.D * 32;
    const int UD = up_const ? -1 : (up ? up->D : 0);   // -1: up_avg holds the top probes' origins (far_field up_const)
    const int topi = top ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    size_t blocks = (n + block - 1) / block;
    if (max_blocks > 0 && !compact && blocks > (size_t)max_blocks) blocks = (size_t)max_blocks;   // resident grid, grid-stride loop
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)block);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && fused && !top) ? 1 : 0;   // only a kernel that waits on its predecessor may start early
#define RC_LAUNCH_MARCH(F, M, T)                                                                                              \
    (compact ? cudaLaunchKernelEx(&cfg, k_march_compact<F, M>, s, L, lv, UD, T, sky, map, origin, dirs, texels, up_avg, link_idx, link_w) \
             : cudaLaunchKernelEx(&cfg, k_march<F, M>, s, L, lv, UD, T, sky, map, n, origin, dirq, texels, up_avg, link_idx, link_w, entry, avg_out, ystep, list, count, quad, count_triv, list_cap))
    const bool f = fused && !top;
    const int t = f ? 0 : topi;
    if (occ >= 16) { if (f) RC_LAUNCH_MARCH(true, 16, t); else RC_LAUNCH_MARCH(false, 16, t); }
    else if (occ >= 12) { if (f) RC_LAUNCH_MARCH(true, 12, t); else RC_LAUNCH_MARCH(false, 12, t); }
    else if (occ >= 10) { if (f) RC_LAUNCH_MARCH(true, 10, t); else RC_LAUNCH_MARCH(false, 10, t); }
    else { if (f) RC_LAUNCH_MARCH(true, 8, t); else RC_LAUNCH_MARCH(false, 8, t); }
#undef RC_LAUNCH_MARCH
}

void launch_need_chain(const NeedChain& c, const float4* origin, const uint4* link_idx, const float4* link_w, uint32_t* need_all,
                       uint32_t* list_all, unsigned int* counts, cudaStream_t st)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kChainCtas);
    cfg.blockDim = dim3(kChainBlock);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kChainCtas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_need_chain, c, origin, link_idx, link_w, need_all, list_all, counts);
}

void launch_march_quad(const DScene& s, const DLights& L, const DLevel& lv, float3 sky, const float4* origin, const float4* dirq, uint2* texels,
                       const float4* up_avg, const uint4* link_idx, const float4* link_w, float4* avg_out, bool fused, int occ, bool pdl, int max_blocks,
                       const uint32_t* list, const unsigned int* count, bool up_const, cudaStream_t st)
{
    const int UD = up_const ? -1 : 0;
    cudaLaunchConfig_t cfg{};
    const size_t nq = (size_t)lv.sw * lv.sh * lv.D * lv.D / 4;
    size_t blocks = (nq + 63) / 64;
    if (max_blocks > 0 && blocks > (size_t)max_blocks) blocks = (size_t)max_blocks;
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(64);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && fused) ? 1 : 0;
    if (fused) {
        if (occ >= 16) cudaLaunchKernelEx(&cfg, k_march_quad<true, 16>, s, L, lv, UD, sky, origin, dirq, texels, up_avg, link_idx, link_w, avg_out, list, count);
        else if (occ >= 12) cudaLaunchKernelEx(&cfg, k_march_quad<true, 12>, s, L, lv, UD, sky, origin, dirq, texels, up_avg, link_idx, link_w, avg_out, list, count);
        else cudaLaunchKernelEx(&cfg, k_march_quad<true, 8>, s, L, lv, UD, sky, origin, dirq, texels, up_avg, link_idx, link_w, avg_out, list, count);
    } else {
        cudaLaunchKernelEx(&cfg, k_march_quad<false, 12>, s, L, lv, UD, sky, origin, dirq, texels, up_avg, link_idx, link_w, avg_out, list, count);
    }
}

void launch_march_pool(const DScene& s, const DLights& L, const DLevel& lv, float3 sky, const float4* origin, const float4* dirq, uint2* texels,
                       const float4* up_avg, const uint4* link_idx, const float4* link_w, const int4* entry, float4* avg_out, bool fused, int occ,
                       bool pdl, int max_blocks, const uint32_t* list, const unsigned int* count, const unsigned int* count_triv, unsigned list_cap,
                       int thresh, bool up_const, cudaStream_t st)
{
    const int UD = up_const ? -1 : 0;
    cudaLaunchConfig_t cfg{};
    const size_t n = (size_t)lv.sw * lv.sh * lv.D * lv.D;
    size_t blocks = (n + kPoolRays - 1) / kPoolRays;
    if (max_blocks > 0 && blocks > (size_t)max_blocks) blocks = (size_t)max_blocks;
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && fused) ? 1 : 0;
#define RC_LAUNCH_POOL(F, M) cudaLaunchKernelEx(&cfg, k_march_pool<F, M>, s, L, lv, UD, sky, origin, dirq, texels, up_avg, link_idx, link_w, entry, \
                                                avg_out, list, count, count_triv, list_cap, thresh)
    if (!fused) RC_LAUNCH_POOL(false, 10);
    else if (occ >= 12) RC_LAUNCH_POOL(true, 12);
    else if (occ >= 10) RC_LAUNCH_POOL(true, 10);
    else RC_LAUNCH_POOL(true, 8);
#undef RC_LAUNCH_POOL
}

void launch_need(const DLevel& lv, int Dr, int has_upper, int up_words, const float4* origin, const uint4* link_idx,
                 const float4* link_w, uint32_t* need, uint32_t* need_up, uint32_t* list, unsigned int* count, bool clear, bool pdl,
                 bool trigger, bool dir_major, int tile_order, bool append, int4 own, cudaStream_t st)
{
    const size_t total = (size_t)lv.sw * lv.sh * ((Dr * Dr + 31) / 32);
    if (!total) return;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks_for(total));
    cfg.blockDim = dim3(kBlock);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    // tile order needs whole (probe, row-pair) groups per warp: 256-thread blocks and 32 words per probe at Dr = 32 guarantee it
    cudaLaunchKernelEx(&cfg, k_need, lv, Dr, has_upper, up_words, origin, link_idx, link_w, need, need_up, list, count, clear ? 1 : 0,
                       trigger ? 1 : 0, dir_major ? 1 : 0, (tile_order && has_upper != 1 && (Dr == 32 || (tile_order > 1 && (Dr == 8 || Dr == 16)))) ? 1 : 0,
                       append ? 1 : 0, own);
}

int split_chunk() { return kSplitChunk; }

void launch_split(const SplitPlan& plan, cudaStream_t st)
{
    if (plan.n <= 0 || !plan.block_off[plan.n]) return;
    k_split<<<plan.block_off[plan.n], kSplitBlock, 0, st>>>(plan);
}

void launch_march_all(const DScene& s, const DLights& L, const DLevelSet& ls, const int* levels, int n, const int* map,
                      const size_t* dir_offset, int entry_levels, int top_level, float3 sky, const float4* origin,
                      const float* dirs, uint2* cascade, const int4* entry, int occ, cudaStream_t st)
{
    MarchPlan plan{};
    plan.n = n;
    unsigned blocks = 0;
    for (int k = 0; k < n; k++) {
        const DLevel& lv = ls.lv[levels[k]];
        int m = map[levels[k]];
        if (m == MAP_DIR_TILE && (lv.D & 7)) m = MAP_LINEAR;
        size_t cnt = (size_t)lv.sw * lv.sh * lv.D * lv.D;
        if (m == MAP_PROBE_TILE) cnt = (size_t)((lv.sw + 7) / 8) * ((lv.sh + 3) / 4) * lv.D * lv.D * 32;
        plan.level[k] = levels[k];
        plan.map[k] = m;
        plan.top[k] = levels[k] == top_level ? 1 : 0;
        plan.use_entry[k] = levels[k] < entry_levels ? 1 : 0;
        plan.dir_offset[k] = (unsigned)dir_offset[levels[k]];
        plan.block_offset[k] = blocks;
        blocks += (unsigned)((cnt + 127) / 128);
    }
    plan.block_offset[n] = blocks;
    if (!blocks) return;
    if (occ >= 12) k_march_all<12><<<blocks, 128, 0, st>>>(s, L, ls, plan, sky, origin, dirs, cascade, entry);
    else if (occ >= 10) k_march_all<10><<<blocks, 128, 0, st>>>(s, L, ls, plan, sky, origin, dirs, cascade, entry);
    else k_march_all<8><<<blocks, 128, 0, st>>>(s, L, ls, plan, sky, origin, dirs, cascade, entry);
}

void launch_march_persist(const DScene& s, const DLights& L, const DLevel& lv, const DLevel* up, bool top, float3 sky,
                          const float4* origin, const float* dirs, uint2* texels, const float4* up_avg,
                          const uint4* link_idx, const float4* link_w, bool fused, int map, int thresh, int grid_blocks,
                          unsigned int* counter, bool pdl, cudaStream_t st)
{
    if (map == MAP_DIR_TILE && (lv.D & 7)) map = MAP_LINEAR;
    unsigned long long n = (unsigned long long)lv.sw * lv.sh * lv.D * lv.D;
    if (map == MAP_PROBE_TILE) n = (unsigned long long)((lv.sw + 7) / 8) * ((lv.sh + 3) / 4) * lv.D * lv.D * 32ull;
    const int UD = up ? up->D : 0;
    const int topi = top ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid_blocks);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (fused && !top)
        cudaLaunchKernelEx(&cfg, k_march_persist<true>, s, L, lv, UD, 0, sky, map, n, thresh, counter, origin, dirs, texels, up_avg, link_idx, link_w);
    else
        cudaLaunchKernelEx(&cfg, k_march_persist<false>, s, L, lv, UD, topi, sky, map, n, thresh, counter, origin, dirs, texels, up_avg, link_idx, link_w);
}

int march_persist_blocks_per_sm()
{
    int a = 0, b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_march_persist<true>, 128, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_march_persist<false>, 128, 0);
    return a < b ? a : b;
}

void launch_fill_top(const DLevel& lv, float3 sky, const float4* origin, uint2* texels, float4* avg_out, cudaStream_t st)
{
    const size_t n2 = (size_t)lv.sw * lv.sh * (((size_t)lv.D * lv.D) >> 1);
    k_fill_top<<<blocks_for(n2), kBlock, 0, st>>>(lv, sky, origin, reinterpret_cast<uint4*>(texels), avg_out);
}

void launch_child_avg(const DLevel& lv, const uint2* texels, float4* avg_out, cudaStream_t st)
{
    const size_t n = (size_t)lv.sw * lv.sh * (((size_t)lv.D * lv.D) >> 2);
    if (n) k_child_avg<<<blocks_for(n), kBlock, 0, st>>>(lv, texels, avg_out);
}

void launch_merge(const DLevel& lv, const DLevel& up, float3 sky, const float4* origin, uint2* texels, const float4* up_avg,
                  const uint4* link_idx, const float4* link_w, cudaStream_t st)
{
    const size_t n = (size_t)lv.sw * lv.sh * lv.D * lv.D;
    k_merge<<<blocks_for(n), kBlock, 0, st>>>(lv, up.D, sky, origin, texels, up_avg, link_idx, link_w);
}

// The pairs of antipodal level-0 directions k_gather_mma<true> relies on (S3 table, D = 4)
bool gather_dirs_symmetric(const float* d)
{
    static const int pairs[8][2] = {{5, 15}, {6, 12}, {9, 3}, {10, 0}, {1, 14}, {2, 13}, {4, 11}, {7, 8}};
    for (auto& pr : pairs)
        for (int k = 0; k < 3; k++)
            if (!(d[3 * pr[0] + k] == -d[3 * pr[1] + k])) return false;   // +0 == -0: only the sign of a zero may differ
    return true;
}

static inline int floor_div_i(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

void launch_gather(const DCamera& cam, const DLevel& l0, TileRect tile, const float4* origin0, const uint2* texels0,
                   const float* dirs0, const float* depth, const uint32_t* normal, uint2* out, unsigned int* counts_in,
                   unsigned int* counts_out, const PeerOut& peer, int tiles_per_block, const GatherMma& mma, cudaStream_t st)
{
    dim3 grid((tile.w + 31) / 32, (tile.h + 7) / 8);
    const int DD = l0.D * l0.D;
    if (mma.enabled && DD == 16 && l0.P == 4 && mma.dirs0_host && mma.axis_nx && mma.axis_ny) {
        // cells of the global 4x4-pixel cell grid (origin at pixel (2, 2)) the tile touches
        const int cbx0 = floor_div_i(tile.x0 - 2, 4), cbx1 = floor_div_i(tile.x0 + tile.w - 1 - 2, 4);
        const int cby0 = floor_div_i(tile.y0 - 2, 4), cby1 = floor_div_i(tile.y0 + tile.h - 1 - 2, 4);
        const int ncx = cbx1 - cbx0 + 1, ncy = cby1 - cby0 + 1;
        const int nsteps = (ncy + kGmCellsY - 1) / kGmCellsY;
        const int spb = tiles_per_block > 0 ? tiles_per_block : 1;
        dim3 g2((ncx + kGmCellsX - 1) / kGmCellsX, (nsteps + spb - 1) / spb);
        Dirs16 dv;
        for (int k = 0; k < 48; k++) dv.v[k] = mma.dirs0_host[k];
        const size_t smem = sizeof(GatherMmaSmem);
        if (mma.symmetric)
            k_gather_mma<true><<<g2, 256, smem, st>>>(cam, l0, tile, dv, mma.axis_nx, mma.axis_ny, origin0, texels0, depth, normal, out, cbx0, cby0,
                                                      nsteps, spb, counts_in, counts_out, peer);
        else
            k_gather_mma<false><<<g2, 256, smem, st>>>(cam, l0, tile, dv, mma.axis_nx, mma.axis_ny, origin0, texels0, depth, normal, out, cbx0, cby0,
                                                       nsteps, spb, counts_in, counts_out, peer);
        return;
    }
    const int max_probes = ((32 + l0.P - 1) / l0.P + 2) * ((8 + l0.P - 1) / l0.P + 2);
    const size_t smem = (size_t)max_probes * ((DD / 2 + 1) * 16 + 16) + (size_t)DD * 3 * sizeof(float);
    // dynamic shared memory above the 48 KB default must be opted into per kernel AND per device (the attribute is
    // per-context state); a window that does not fit the SM at all is reported as a launch error by the runtime
    auto opt_in = [](const void* fn, size_t bytes) {
        if (bytes > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    };
    if (DD == 16 && tiles_per_block > 1) {   // software-pipelined column of tiles (two staging buffers)
        const size_t smem2 = 2 * (size_t)max_probes * ((DD / 2 + 1) * 16 + 16) + (size_t)DD * 3 * sizeof(float);
        if (smem2 <= 200 * 1024) {
            dim3 grid2((tile.w + 31) / 32, ((tile.h + 7) / 8 + tiles_per_block - 1) / tiles_per_block);
            opt_in((const void*)k_gather_pipe, smem2);
            k_gather_pipe<<<grid2, kBlock, smem2, st>>>(cam, l0, tile, origin0, texels0, dirs0, depth, normal, out, max_probes, tiles_per_block,
                                                       counts_in, counts_out, peer);
            return;
        }
    }
    if (DD == 16) {
        opt_in((const void*)k_gather<16>, smem);
        k_gather<16><<<grid, kBlock, smem, st>>>(cam, l0, tile, origin0, texels0, dirs0, depth, normal, out, max_probes, counts_in, counts_out, peer);
        return;
    }
    opt_in((const void*)k_gather<0>, smem);
    k_gather<0><<<grid, kBlock, smem, st>>>(cam, l0, tile, origin0, texels0, dirs0, depth, normal, out, max_probes, counts_in, counts_out, peer);
}

void launch_peer_begin(const PeerOut& peer, uint32_t* my_ctrl, cudaStream_t st) { k_peer_begin<<<1, 32, 0, st>>>(peer, my_ctrl); }
void launch_peer_publish(const PeerOut& peer, cudaStream_t st) { k_peer_publish<<<1, 32, 0, st>>>(peer); }
void launch_peer_wait(int world, uint32_t seq, uint32_t* my_ctrl, cudaStream_t st) { k_peer_wait<<<1, 32, 0, st>>>(world, seq, my_ctrl); }

void launch_copy_to_host(const void* src, void* dst_host_mapped, size_t bytes, int blocks, cudaStream_t st)
{
    k_copy_to_host<<<blocks, kBlock, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst_host_mapped), bytes / 16);
}

void launch_pack_rgb48(size_t n_pixels, const uint2* irradiance, void* out, cudaStream_t st)
{
    const size_t n8 = n_pixels / 8;
    k_pack_rgb48<<<blocks_for(n8 + 1), kBlock, 0, st>>>(n8, n_pixels, reinterpret_cast<const uint4*>(irradiance), reinterpret_cast<uint4*>(out),
                                                       irradiance, reinterpret_cast<uint16_t*>(out));
}

void launch_composite(TileRect tile, const uint2* irradiance, const uint2* albedo, const uint2* direct,
                      uchar4* composite, uchar4* direct_srgb, cudaStream_t st)
{
    const int n = tile.w * tile.h;
    k_composite<<<blocks_for((size_t)n), kBlock, 0, st>>>(n, irradiance, albedo, direct, composite, direct_srgb);
}

void launch_trace_rays(const DScene& s, const float* rays, uint32_t n, float* hits, cudaStream_t st)
{
    k_trace_rays<<<blocks_for(n), kBlock, 0, st>>>(s, rays, n, hits);
}

void launch_shade_points(const DScene& s, const DLights& L, const float* in, uint32_t n, float* out, cudaStream_t st)
{
    k_shade_points<<<blocks_for(n), kBlock, 0, st>>>(s, L, in, n, out);
}

}  // namespace rc
