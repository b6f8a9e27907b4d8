// rc_api.cu — the C ABI of include/rc_b200.h: context, device memory, frame scheduling.
// Mirrors DefaultRenderer::new / RenderStage::{update, resize, render}
// (src/renderer.rs:168-632) as a headless CUDA path.  There is no CPU fallback:
// without a CUDA device rc_create fails with RC_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "../../include/rc_b200.h"
#include "bvh.h"
#include "kernels.cuh"
#include "scene.h"

using namespace rc;

namespace {

thread_local std::string g_create_error;

struct HostModel {
    std::vector<float> stream;      // 17 floats / vertex
    std::vector<uint32_t> indices;  // reversed winding
    float material80[20];           // UniformMaterial + enable_bit + Ke
    int tex_c = -1, tex_n = -1;     // texture slots or -1
    std::string name;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    bool fresh = false;              // the last alloc() had to take new memory (contents undefined)
    // grow-only: a request that fits the current allocation reuses it (rc_set_tile re-tiles a context every few
    // frames and must not touch the allocator)
    cudaError_t alloc(size_t count)
    {
        fresh = false;
        if (count <= cap && p) { n = count; return cudaSuccess; }
        release();
        if (!count) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) { n = cap = count; fresh = true; }
        return e;
    }
    cudaError_t upload(const std::vector<T>& v)
    {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    // zero-filled when new memory was taken, or when `always` (a reused buffer keeps its — finite — contents otherwise)
    cudaError_t alloc_zero(size_t count, bool always = true)
    {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || !count) return e;
        if (!always && !fresh) return cudaSuccess;
        return cudaMemset(p, 0, count * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = cap = 0; }
};

}  // namespace

// Host-side scene: everything ObjScene::load + DefaultRenderer::new prepare on the CPU
// (src/primitives.rs:122-175, src/renderer.rs:370-497), flattened for upload.  Device-free.
struct rc_scene {
    std::string error;
    std::vector<HostModel> models;
    rc_scene_info info{};
    std::vector<float> verts;                 // [NV][17]
    std::vector<uint32_t> tris, tri_model;    // global ids
    std::vector<DMaterial> mats;
    std::vector<DTexture> tex;
    std::vector<uint8_t> tex_data;
    std::vector<float> v0, e1, e2;            // S5 set-up
    std::vector<uint8_t> skip;
    float diag = 0.f;
};

struct rc_ctx {
    rc_config cfg{};
    std::string scene_path, error;
    int device = 0;
    cudaStream_t stream = nullptr;

    rc_scene host;                            // kept for the parity getters

    // scene (device)
    DevBuf<float4> d_nodes, d_tri_geom, d_tri_eg;
    DevBuf<uint32_t> d_tris, d_tri_model;
    DevBuf<float> d_verts, d_srgb;
    DevBuf<DMaterial> d_mats;
    DevBuf<DTexture> d_tex;
    DevBuf<uint8_t> d_tex_data;
    DScene scene{};

    // frame state
    uint32_t W = 0, H = 0;
    TileRect tile{};
    uint32_t N = 0;
    float L0 = 0, t_far = 0, offset = 0;
    std::vector<DLevel> levels;
    std::vector<rc_level_info> level_info;
    std::vector<size_t> dir_offset;      // floats before level i in d_dirs
    std::vector<std::vector<float>> dirs_host;
    DevBuf<uint2> d_cascade;             // all levels, RGBA16F texels
    DevBuf<float4> d_origin, d_normal;   // all probes
    DevBuf<uint4> d_link_idx;
    DevBuf<float4> d_link_w;
    DevBuf<int4> d_entry;                // 2 x int4 per probe: BVH entry frontier (k_entry)
    DevBuf<float4> d_avg;                // levels >= 1: child averages per (probe, lower direction) for the merge (S8)
    std::vector<size_t> avg_offset;      // float4s before level i in d_avg (level 0 has none)
    float4* avg_of(uint32_t level) { return level >= 1 && level < N ? d_avg.p + avg_offset[level] : nullptr; }
    // direction culling (kernels.cu k_need): per level request masks, ray lists and list lengths
    DevBuf<uint32_t> d_need, d_list;
    DevBuf<uint16_t> d_pixmask;                     // per pixel: level-0 directions with cs_d > 0 (k_gbuffer)
    DevBuf<unsigned int> d_ray_count;
    // List lengths are only known on the device.  The frame's last kernel (k_gather) writes them to mapped pinned
    // host memory (posted writes, no copy, no sync); the next frames size their march grids from the latest values that have arrived
    // (x1.25 + slack).  The kernels loop over the list with a grid stride, so any grid size is correct — a
    // stale or missing estimate only costs empty blocks (too large) or a second loop trip (too small).
    unsigned int* h_ray_count = nullptr;            // pinned, 2 x RC_MAX_LEVELS entries: [i] = entries of level i's list (0xffffffff = nothing arrived yet),
                                                    // [RC_MAX_LEVELS + i] = its certain misses, bit 31 set if k_split classified the level in that frame
    // Split lists: k_split sorts certain misses (root-box test) to the back of a copy of the level's list, k_march skips their traversal.
    // 0 = off, 2 = every level, 1 = adaptive: the test costs ~40 instructions per requested ray and pays only where many rays leave
    // the scene at once, so a level is classified while the last measured share of certain misses is high (on at >= 25 %, off
    // below 15 %); every level >= 1 is measured in the first frame and in every 64th.  Any choice gives the same texels.
    int list_split = 1;
    bool split_on[RC_MAX_LEVELS] = {};
    int march_pool = 0, march_pool_thresh = 20;      // k_march_pool (block-local ray pool with refill) for the culled levels >= 1
    uint32_t split_mask = 0;                         // levels whose split list (d_list2) the frame being enqueued marches
    DevBuf<uint32_t> d_list2;
    unsigned int list_len(uint32_t i) const          // entries of level i's last list that arrived (0xffffffff: none yet)
    {
        const unsigned int a = *(volatile unsigned int*)(h_ray_count + i);
        return a;
    }
    bool split_level(uint32_t i)                     // classify level i in the frame being enqueued?
    {
        if (!list_split || march_quad) return false;
        if (list_split >= 2) return true;
        if (i == 0) return false;                    // level 0: short rays from the entry frontier, a certain miss costs the march next to nothing
        const unsigned int a = *(volatile unsigned int*)(h_ray_count + i), t = *(volatile unsigned int*)(h_ray_count + RC_MAX_LEVELS + i);
        if (a == 0xffffffffu) return true;           // nothing has arrived yet: measure
        if (t & 0x80000000u) {
            const double share = a ? (double)(t & 0x7fffffffu) / (double)a : 0.0;
            if (share >= 0.25) split_on[i] = true; else if (share < 0.15) split_on[i] = false;
            split_triv[i] = t & 0x7fffffffu;
        }
        return split_on[i] || split_measure_frame();
    }
    bool split_measure_frame() const { return (frame_id & 63u) == 1u; }
    // k_split is one more launch in the latency-bound chain before the march (~8 us however short the lists are) and a certain miss
    // saves the march ~70 ns per thousand entries... i.e. ~7 us per 100 k: below that the classification costs more than it saves
    // (teapot 1080p: only level 4 qualifies, 82 k certain misses, frame +0.8 % with the split)
    static constexpr unsigned kSplitMinMisses = 120000u;
    unsigned int split_triv[RC_MAX_LEVELS] = {};
    std::vector<size_t> need_offset, list_offset;   // words / entries before level i
    std::vector<int> need_res;                      // Dr_i: D_0 for levels 0 and 1, D_{i-1} above
    int cull = 1;                                   // rc_set_tuning("cull", 0) marches every texel
    int list_dir_major = 0;
    // levels >= 1 of the request chain in one cluster launch (k_need_chain, eight 1024-thread blocks with cluster barriers between
    // the levels).  Measured SLOWER: probes stage 0.128 -> 0.46 ms at 4K, 0.058 -> 0.11 at 1080p — level 1 alone is 130 k mask
    // words at 4K, sixteen passes of a cluster that occupies 8 of 148 SMs.  Off; the per-level launches stay.
    int need_fused = 0;
    int list_tiled = 1;                             // levels >= 2: ray lists in 4x2 quad-tile order (k_need tile_order)                         // bit i: level i's ray list is ordered direction-major inside each warp's share
    // rc_read_target_async: 0 = copy engine (cudaMemcpyAsync); n > 0 = n resident blocks of k_copy_to_host store the
    // target into the page-locked destination
    int copy_blocks = 0;
    int need_pdl = 0;                               // 1: the k_need chain (levels >= 1) uses programmatic dependent launch
    // k_gather_pipe: a block walks that many 32x8 pixel tiles with the next tile's probes prefetched; 1 = k_gather;
    // 0 = auto: 4 on large frames, 2 on small ones (measured: 4K 0.292 -> 0.252 ms with 4; 1080p 0.057 -> 0.050 ms with 2,
    // 0.054 with 4 — too few blocks per SM left)
    int gather_tiles = 0;
    int gather_tiles_eff() const
    {
        if (gather_tiles > 0) return gather_tiles;
        const size_t t = (size_t)((tile.w + 31) / 32) * ((tile.h + 7) / 8);
        if (gather_mma && levels[0].D == 4 && levels[0].P == 4) return t >= 16384 ? 8 : 4;   // k_gather_mma: steps per block (A/B r2d)
        return t >= 16384 ? 4 : 2;
    }
    // k_gather_mma (tensor-core gather, D0 = 4 and P0 = 4): 1 = default, 0 = the scalar kernels (A/B, exact S9 arithmetic)
    // 0: render without the peer-memory stores although peers are attached (every rank must switch in the same frame)
    int peer_stores = 1;
    // 0 (default): the finished tiles are gathered on rank 0 (final image gather); 1: every rank receives every tile
    // (all-gather: N times the NVLink traffic — at 8 ranks the stores cost more than the gather itself).  Same value on all ranks.
    int peer_broadcast = 0;
    // Halo exchange (RC_CFG_HALO_EXCHANGE on a tiled context): levels >= 1 are marched only for the probes this rank owns; the
    // caller moves request masks (before rc_render_lists) and child averages (after every rc_render_level) between the ranks
    bool exchange_active() const { return (cfg.flags & RC_CFG_HALO_EXCHANGE) && cfg.tile_w && cfg.tile_h && cull_possible(); }
    bool lists_pending = false;
    int gather_mma = 1;
    bool gather_sym = false;                        // the level-0 direction table is point-symmetric (gather_dirs_symmetric)
    DevBuf<float> d_axis;                           // S4: nx(x) for x < W, then ny(y) for y < H
    // primary visibility by triangle binning (k_bin / k_gbuffer_binned) instead of the per-pixel BVH traversal (k_gbuffer).
    // Bit-identical, and 40 % fewer instructions per pixel (859 vs 1410 warp-instructions per warp at 4K), but measured SLOWER
    // on every bundled scene (living_room 4K 0.50 vs 0.315 ms, test_room 0.15 vs 0.08, teapot 0.14 vs 0.13): k_bin is bound
    // by the latency of its returning atomics (0.14 ms, 7 % of the warp slots busy) and every tile block starts with a chain
    // of four dependent loads (count -> list -> triangles -> test) that four resident blocks per SM do not hide
    // (ncu r2l: a third of the stall samples in the block prologue).  Off by default; kept as an A/B path.
    int gbuffer_binned = 0;
    uint32_t n_leaf_tris = 0;
    DevBuf<uint32_t> d_occ;                         // 32x8-pixel occupancy stamps of the tile (GBufferOut::occ), frame_id = the stamp
    uint32_t frame_id = 0;
    DevBuf<unsigned int> d_bin_count, d_bin_huge_count;
    DevBuf<uint32_t> d_bin_lists;
    DevBuf<uint8_t> d_bin_huge;
    // rc_render records the frame's ~18 launches into a CUDA graph (stream capture) and submits it with ONE
    // cudaGraphLaunch; every frame is re-captured and the executable graph updated in place
    // (cudaGraphExecUpdate: camera, lights, grid sizes and the output slot are node parameters).
    int use_graph = 1;
    // tiled multi-GPU: every rank's gather stores its tile into all ranks' full-frame buffers over NVLink peer
    // mappings (rc_peer_export / rc_peer_attach); two slots per rank + a control block, one cudaMalloc shared by IPC
    struct Peer {
        int world = 0, rank = 0;
        void* local = nullptr;                  // this rank's shared allocation
        void* base[kMaxPeers] = {};             // every rank's allocation as mapped here (base[rank] == local)
        size_t slot_bytes = 0;
        uint32_t seq = 0;                       // frames delivered so far
        uint2* frame(int r, uint32_t q) const { return (uint2*)((char*)base[r] + (q & 1u) * slot_bytes); }
        uint32_t* ctrl(int r) const { return (uint32_t*)((char*)base[r] + 2 * slot_bytes); }
    } peer;
    cudaGraphExec_t graph_exec = nullptr;
    bool capturing = false;
    bool frame_culled = false;                      // the frame being recorded uses the ray lists
    bool frame_open = false;                        // rc_render_begin without its rc_render_end (which resets the list lengths)
    bool cull_possible() const
    {
        const uint32_t D0 = (uint32_t)levels[0].D;
        return cull && !(cfg.flags & RC_CFG_SEPARATE_MERGE) && !march_persist && !march_compact && !march_batch &&
               D0 * D0 <= 16 && (D0 & (D0 - 1)) == 0;
    }
    DevBuf<float> d_dirs;
    // per direction (w.xyz, 1/w.x), (1/w.y, 1/w.z, 0, 0): the slab-test reciprocals of S5 depend on the direction only,
    // so they are divided once here (IEEE, the same safe_inv expression) instead of three times per ray in k_march
    DevBuf<float4> d_dirq;
    DevBuf<float4> d_qinv;                         // per level >= 1 and quad: slab reciprocals of its four children (k_split)
    size_t qinv_offset[RC_MAX_LEVELS] = {};
    DevBuf<float> d_depth;
    DevBuf<uint32_t> d_prim, d_nrm;
    DevBuf<uint2> d_albedo, d_direct, d_irr, d_irr2;   // irradiance is double-buffered for rc_read_target_async
    int irr_slot = 0;                    // buffer written by the last rc_render
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_frame_done = nullptr, ev_copy_done[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};
    uint2* irr() { return irr_slot ? d_irr2.p : d_irr.p; }
    DevBuf<uchar4> d_composite, d_direct_srgb;
    DevBuf<uint16_t> d_irr_pack;         // RC_TARGET_IRRADIANCE_RGB48 staging (3 x float16 per pixel)
    DevBuf<float> d_dbg_in, d_dbg_out;
    DevBuf<unsigned int> d_counters;     // per-level ray-fetch counters of the persistent march

    DCamera cam{};
    DLights lights{};
    bool have_camera = false;
    bool composite_valid = false, direct_valid = false;
    DevBuf<float2> d_bary;               // barycentrics of the primary hit
    DCamera cam_rendered{};              // camera / lights of the last rendered frame (deferred direct pass)
    DLights lights_rendered{};

    cudaEvent_t ev[8]{};
    cudaEvent_t ev_level[RC_MAX_LEVELS + 1]{};   // ev_level[i] recorded after level i's kernels
    bool ev_recorded = false;
    bool frame_batched = false;          // the last frame went through render_levels_batched
    uint32_t launches = 0;
    cudaStream_t last_stream = nullptr;
    int march_map[RC_MAX_LEVELS];   // thread->texel mapping per level (kernels.cu MAP_*)
    int march_block = 128;   // blocks of 128 threads; march_occ = min resident blocks/SM the registers must allow
    int march_occ = 10;      // measured best of 8/10/12/16
    int march_persist = 0, march_thresh = 8, march_grid = 0, march_pdl = 1;   // persistent variant measured slower
    bool level_timing = false;
    int march_compact = 0;   // bit i: level i uses the block-compacting march kernel
    int fill_top = 1;        // fill a top level that cannot hit anything instead of marching it (exact)
    int march_waves = 0;     // > 0: march grid capped at SMs * occ * waves blocks (grid-stride loop); 0: one thread per ray
    int sm_count = 148;
    // levels 0..march_entry-1 start their traversal at the per-probe BVH entry frontier (k_link_entry); 0 = at the root.
    // Measured: with every texel marched the frontiers save 0.15 ms of a 4K living_room frame (L0 -24 %,
    // L1 -22 %, L2 -9 %); with direction culling half of those rays are gone and building the frontiers (35 us) costs
    // what they save, so the default is off.  -1 = the levels whose interval ends within a tenth of the scene
    // diagonal, on frames of at least 3 M texels per level.
    int march_entry = 0;
    // 1: march all levels in one launch, then merge top-down with k_merge; 0 (default): one fused march+merge kernel
    // per level, PDL-chained.  Measured: the single launch saves nothing over the PDL chain and the
    // separate merges cost more than the fused ones.
    int march_batch = 0;
    // culled levels >= 1: one 2x2 quad of texels per thread (k_march_quad) instead of one ray per thread; march_quad_occ = resident
    // 64-thread blocks per SM the register allocation must allow
    // Measured SLOWER (4K march 1.02-1.18 vs 0.685 ms, teapot 0.39-0.51 vs 0.30): 32 different quads per warp, four rounds of
    // traversal each waiting for its slowest lane — the per-ray overhead it saves is smaller than the divergence it adds.  Off.
    int march_quad = 0, march_quad_occ = 16;
    size_t texels_per_level() const { return levels.empty() ? 0 : (size_t)levels[0].sw * levels[0].sh * levels[0].D * levels[0].D; }
    bool batched() const
    {
        if (march_persist || level_timing) return false;
        return march_batch > 0;
    }
    bool top_fillable() const
    {
        // S7 shortcut, exact: a ray that starts on a surface (inside the scene's box grown by the probe offset) is
        // farther than the box diagonal from every triangle once t > diag + 2*offset -> the whole level misses
        const DLevel& L = levels[N - 1];
        return fill_top && L.t0 > host.diag * 1.001f + 2.0f * offset && (((size_t)L.texel_offset) & 1) == 0;
    }
    int entry_levels() const
    {
        if (march_entry >= 0) return march_entry < (int)N ? march_entry : (int)N;
        if (texels_per_level() < 3000000u) return 0;
        int n = 0;
        while (n < (int)N && levels[n].t1 <= 0.1f * host.diag) n++;
        return n;
    }

    bool fail(rc_status, const std::string& m) { error = m; return false; }
};

namespace {

#define CU_OK(ctx, call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            (ctx)->error = std::string(#call) + ": " + cudaGetErrorString(e__);           \
            return RC_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// rc_spec.h S3: equal-area octahedral directions, double -> float
void make_directions(int D, std::vector<float>& out)
{
    const double PI = 3.14159265358979323846;
    out.resize((size_t)D * D * 3);
    for (int dy = 0; dy < D; dy++)
        for (int dx = 0; dx < D; dx++) {
            double u = (2.0 * dx + 1.0) / D - 1.0, v = (2.0 * dy + 1.0) / D - 1.0;
            double d = 1.0 - (std::fabs(u) + std::fabs(v)), r = 1.0 - std::fabs(d);
            double phi = (r == 0.0) ? 0.0 : (PI / 4.0) * ((std::fabs(v) - std::fabs(u)) / r + 1.0);
            double f = r * std::sqrt(2.0 - r * r);
            float* o = &out[3 * ((size_t)dy * D + dx)];
            o[0] = (float)std::copysign(f * std::cos(phi), u);
            o[1] = (float)std::copysign(f * std::sin(phi), v);
            o[2] = (float)std::copysign(1.0 - r * r, d);
        }
}

bool invert4(const double m[16], double inv[16])   // column-major Gauss-Jordan with partial pivoting
{
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { a[r][c] = m[c * 4 + r]; a[r][c + 4] = (r == c) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int piv = c;
        for (int r = c + 1; r < 4; r++) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (a[piv][c] == 0.0) return false;
        if (piv != c) for (int k = 0; k < 8; k++) std::swap(a[c][k], a[piv][k]);
        double d = a[c][c];
        for (int k = 0; k < 8; k++) a[c][k] /= d;
        for (int r = 0; r < 4; r++)
            if (r != c) { double f = a[r][c]; for (int k = 0; k < 8; k++) a[r][k] -= f * a[c][k]; }
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) inv[c * 4 + r] = a[r][c + 4];
    return true;
}

// rc_spec.h S4
bool primary_basis(const rc_camera& c, DCamera& out)
{
    double m[16], inv[16];
    for (int i = 0; i < 16; i++) m[i] = c.view_proj[i];
    if (!invert4(m, inv)) return false;
    const double e[3] = {c.eye[0], c.eye[1], c.eye[2]};
    double A[4], B[4], C[4];
    for (int r = 0; r < 4; r++) { A[r] = inv[r]; B[r] = inv[4 + r]; C[r] = inv[8 + r] + inv[12 + r]; }
    double dx[3], dy[3], dc[3];
    for (int k = 0; k < 3; k++) { dx[k] = A[k] - e[k] * A[3]; dy[k] = B[k] - e[k] * B[3]; dc[k] = C[k] - e[k] * C[3]; }
    const double sc = (C[3] < 0 ? -1.0 : 1.0) / std::sqrt(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]);
    out.eye = make_float3(c.eye[0], c.eye[1], c.eye[2]);
    out.dx = make_float3((float)(dx[0] * sc), (float)(dx[1] * sc), (float)(dx[2] * sc));
    out.dy = make_float3((float)(dy[0] * sc), (float)(dy[1] * sc), (float)(dy[2] * sc));
    out.dc = make_float3((float)(dc[0] * sc), (float)(dc[1] * sc), (float)(dc[2] * sc));
    // S4b: rows 2 and 3 of the column-major view_proj, and the rows applied to (eye, 1) — float, fixed fma order
    const float* M = c.view_proj;
    const float ex = c.eye[0], ey = c.eye[1], ez = c.eye[2];
    out.clip_z = make_float4(M[2], M[6], M[10], fmaf(M[10], ez, fmaf(M[6], ey, fmaf(M[2], ex, M[14]))));
    out.clip_w = make_float4(M[3], M[7], M[11], fmaf(M[11], ez, fmaf(M[7], ey, fmaf(M[3], ex, M[15]))));
    out.row_x = make_float4(M[0], M[4], M[8], M[12]);
    out.row_y = make_float4(M[1], M[5], M[9], M[13]);
    out.row_w = make_float4(M[3], M[7], M[11], M[15]);
    return true;
}

inline int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Probe sub-grid every level must hold so the tile can be gathered and merged (S1 footprints).
void level_rects(uint32_t W, uint32_t H, uint32_t P0, uint32_t N, TileRect t, std::vector<DLevel>& lv)
{
    int xlo = 0, xhi = 0, ylo = 0, yhi = 0;
    for (uint32_t i = 0; i < N; i++) {
        DLevel& L = lv[i];
        L.P = (int)(P0 << i);
        L.gw = (int)((W + L.P - 1) / L.P);
        L.gh = (int)((H + L.P - 1) / L.P);
        if (i == 0) {
            xlo = clampi(floor_div(t.x0 - (int)P0 / 2, (int)P0), 0, L.gw - 1);
            xhi = clampi(floor_div(t.x0 + t.w - 1 - (int)P0 / 2, (int)P0) + 1, 0, L.gw - 1);
            ylo = clampi(floor_div(t.y0 - (int)P0 / 2, (int)P0), 0, L.gh - 1);
            yhi = clampi(floor_div(t.y0 + t.h - 1 - (int)P0 / 2, (int)P0) + 1, 0, L.gh - 1);
        } else {
            auto lo_of = [](int q) { return (q % 2 == 0) ? q / 2 - 1 : (q - 1) / 2; };
            xlo = clampi(lo_of(xlo), 0, L.gw - 1);
            xhi = clampi(lo_of(xhi) + 1, 0, L.gw - 1);
            ylo = clampi(lo_of(ylo), 0, L.gh - 1);
            yhi = clampi(lo_of(yhi) + 1, 0, L.gh - 1);
        }
        L.px0 = xlo; L.py0 = ylo; L.sw = xhi - xlo + 1; L.sh = yhi - ylo + 1;
    }
}

rc_status setup_frame(rc_ctx* c, uint32_t W, uint32_t H, bool retile = false)
{
    c->W = W; c->H = H;
    const rc_config& cfg = c->cfg;
    TileRect t{0, 0, (int)W, (int)H};
    if (cfg.tile_w && cfg.tile_h) {
        if (cfg.tile_x0 + cfg.tile_w > W || cfg.tile_y0 + cfg.tile_h > H) { c->error = "tile outside the frame"; return RC_ERR_INVALID_ARG; }
        t = TileRect{(int)cfg.tile_x0, (int)cfg.tile_y0, (int)cfg.tile_w, (int)cfg.tile_h};
    }
    c->tile = t;
    const uint32_t P0 = cfg.probe_spacing0 ? cfg.probe_spacing0 : RC_DEFAULT_P0;
    const uint32_t D0 = cfg.dir_res0 ? cfg.dir_res0 : RC_DEFAULT_D0;
    const uint32_t N = cfg.num_levels ? cfg.num_levels : RC_DEFAULT_LEVELS;
    if (N > RC_MAX_LEVELS || D0 < 2 || (D0 & 1) || P0 < 1 || (D0 << (N - 1)) > 4096) {
        c->error = "unsupported cascade parameters (need even D0 >= 2, N <= 10, D_top <= 4096)";
        return RC_ERR_INVALID_ARG;
    }
    {   // the gather stages the level-0 probes of a 32x8 pixel tile in shared memory (kernels.cu launch_gather)
        const size_t max_probes = (size_t)((32 + P0 - 1) / P0 + 2) * ((8 + P0 - 1) / P0 + 2), DD = (size_t)D0 * D0;
        if (max_probes * ((DD / 2 + 1) * 16 + 16) + DD * 12 > 227u * 1024u) {
            c->error = "unsupported cascade parameters: the gather's probe window (P0, D0) exceeds 227 KB of shared memory";
            return RC_ERR_INVALID_ARG;
        }
    }
    c->N = N;
    c->levels.assign(N, DLevel{});
    level_rects(W, H, P0, N, t, c->levels);
    c->level_info.assign(N, rc_level_info{});
    c->dir_offset.assign(N, 0);
    c->dirs_host.assign(N, {});
    size_t texels = 0, probes = 0, dirf = 0;
    std::vector<float> all_dirs;
    for (uint32_t i = 0; i < N; i++) {
        DLevel& L = c->levels[i];
        L.D = (int)(D0 << i);
        const double a = (std::pow(4.0, (double)i) - 1.0) / 3.0, b = (std::pow(4.0, (double)i + 1.0) - 1.0) / 3.0;
        L.t0 = (float)((double)c->L0 * a);                               // S2
        L.t1 = (i == N - 1) ? c->t_far : (float)((double)c->L0 * b);
        L.texel_offset = texels;
        L.probe_offset = (unsigned)probes;
        const size_t np = (size_t)L.sw * L.sh;
        texels += np * L.D * L.D;
        probes += np;
        make_directions(L.D, c->dirs_host[i]);
        c->dir_offset[i] = dirf;
        dirf += c->dirs_host[i].size();
        all_dirs.insert(all_dirs.end(), c->dirs_host[i].begin(), c->dirs_host[i].end());
        rc_level_info& I = c->level_info[i];
        I.spacing = L.P; I.dir_res = L.D; I.grid_w = L.gw; I.grid_h = L.gh;
        I.px0 = L.px0; I.py0 = L.py0; I.sub_w = L.sw; I.sub_h = L.sh;
        I.texel_offset = L.texel_offset; I.texel_count = np * L.D * L.D;
        I.t_begin = L.t0; I.t_end = L.t1;
    }
    const size_t npx = (size_t)t.w * t.h;
    // retile (rc_set_tile): buffers that are merely reused keep their contents — every value ever stored in them is finite,
    // which is all the zero-weight reads of culled texels need; the request masks are all-zero between frames (k_need)
    CU_OK(c, c->d_cascade.alloc_zero(texels, !retile));
    CU_OK(c, c->d_origin.alloc(probes));
    CU_OK(c, c->d_normal.alloc(probes));
    CU_OK(c, c->d_link_idx.alloc(probes));
    CU_OK(c, c->d_link_w.alloc(probes));
    CU_OK(c, c->d_entry.alloc(2 * probes));
    c->avg_offset.assign(N, 0);
    size_t avgs = 0;
    for (uint32_t i = 1; i < N; i++) {
        c->avg_offset[i] = avgs;
        avgs += (size_t)c->levels[i].sw * c->levels[i].sh * ((size_t)c->levels[i].D * c->levels[i].D / 4);
    }
    // zero-initialised: culled texels / averages are only ever multiplied by zero weights, so they must stay finite
    CU_OK(c, c->d_avg.alloc_zero(avgs ? avgs : 1, !retile));
    c->need_offset.assign(N + 1, 0);
    c->list_offset.assign(N + 1, 0);
    c->need_res.assign(N, 0);
    for (uint32_t i = 0; i < N; i++) {
        const int Dr = i == 0 ? c->levels[0].D : c->levels[i - 1].D;
        const size_t np = (size_t)c->levels[i].sw * c->levels[i].sh;
        c->need_res[i] = Dr;
        c->need_offset[i + 1] = c->need_offset[i] + np * (size_t)((Dr * Dr + 31) / 32);
        c->list_offset[i + 1] = c->list_offset[i] + np * (size_t)(Dr * Dr);
    }
    CU_OK(c, c->d_need.alloc_zero(c->need_offset[N], !retile));
    CU_OK(c, c->d_list.alloc(c->list_offset[N]));
    CU_OK(c, c->d_list2.alloc(c->list_offset[N]));
    CU_OK(c, c->d_ray_count.alloc_zero(3 * RC_MAX_LEVELS, !retile));
    if (!c->h_ray_count) CU_OK(c, cudaHostAlloc((void**)&c->h_ray_count, 2 * RC_MAX_LEVELS * sizeof(unsigned int), cudaHostAllocMapped));
    if (!retile) for (uint32_t i = 0; i < 2 * RC_MAX_LEVELS; i++) c->h_ray_count[i] = i < RC_MAX_LEVELS ? 0xffffffffu : 0u;   // (a re-tiled context keeps its grid-size estimates)
    if (!retile) CU_OK(c, c->d_dirs.upload(all_dirs));
    if (!retile) {   // S4's per-column / per-row terms, evaluated once with the very expressions primary_dir uses
        std::vector<float> axis((size_t)W + H);
        for (uint32_t x = 0; x < W; x++) axis[x] = (float)(2 * (int)x + 1) / (float)(int)W - 1.0f;
        for (uint32_t y = 0; y < H; y++) axis[(size_t)W + y] = 1.0f - (float)(2 * (int)y + 1) / (float)(int)H;
        CU_OK(c, c->d_axis.upload(axis));
        c->gather_sym = c->levels[0].D == 4 && gather_dirs_symmetric(c->dirs_host[0].data());
    }
    if (!retile) {
        std::vector<float4> q(2 * (all_dirs.size() / 3));
        auto safe_inv = [](float d) { return 1.0f / (std::fabs(d) > 1e-20f ? d : std::copysign(1e-20f, d)); };   // rc_device.cuh safe_inv
        for (size_t k = 0; k < all_dirs.size() / 3; k++) {
            const float x = all_dirs[3 * k], y = all_dirs[3 * k + 1], z = all_dirs[3 * k + 2];
            q[2 * k] = make_float4(x, y, z, safe_inv(x));
            q[2 * k + 1] = make_float4(safe_inv(y), safe_inv(z), 0.f, 0.f);
        }
        CU_OK(c, c->d_dirq.upload(q));
        // k_split's table: the same reciprocals, the four children of a quad (levels >= 1: direction (2x + i, 2y + j) of quad (x, y),
        // in the order (0,0) (1,0) (0,1) (1,1)) side by side — 3 x float4 per quad, consecutive quads consecutive
        std::vector<float4> qi;
        for (uint32_t i = 0; i < N; i++) {
            c->qinv_offset[i] = qi.size();
            const int D = c->levels[i].D;
            if (i == 0 || (D & 1)) continue;
            const float4* lq = q.data() + 2 * (c->dir_offset[i] / 3);
            for (int y = 0; y < D / 2; y++)
                for (int x = 0; x < D / 2; x++) {
                    float v[12];
                    for (int ch = 0; ch < 4; ch++) {
                        const size_t d = (size_t)(2 * y + (ch >> 1)) * D + 2 * x + (ch & 1);
                        v[3 * ch] = lq[2 * d].w; v[3 * ch + 1] = lq[2 * d + 1].x; v[3 * ch + 2] = lq[2 * d + 1].y;
                    }
                    for (int k = 0; k < 3; k++) qi.push_back(make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                }
        }
        if (qi.empty()) qi.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
        CU_OK(c, c->d_qinv.upload(qi));
    }
    CU_OK(c, c->d_occ.alloc_zero((size_t)((t.w + 31) / 32) * ((t.h + 7) / 8), !retile));    // stale stamps carry older frame ids
    CU_OK(c, c->d_bin_count.alloc_zero(bin_tiles(t)));          // always zeroed: a re-tiled context starts from empty candidate lists
    CU_OK(c, c->d_bin_lists.alloc(bin_list_entries(t)));
    CU_OK(c, c->d_bin_huge_count.alloc_zero(1));
    CU_OK(c, c->d_bin_huge.alloc(bin_huge_bytes(c->n_leaf_tris)));
    CU_OK(c, c->d_depth.alloc(npx));
    CU_OK(c, c->d_prim.alloc(npx));
    CU_OK(c, c->d_nrm.alloc(npx));
    CU_OK(c, c->d_bary.alloc(npx));
    CU_OK(c, c->d_pixmask.alloc(npx));
    CU_OK(c, c->d_albedo.alloc(npx));
    CU_OK(c, c->d_direct.alloc(npx));
    CU_OK(c, c->d_irr.alloc(npx));
    CU_OK(c, c->d_irr2.alloc(npx));
    c->copy_pending[0] = c->copy_pending[1] = false;
    CU_OK(c, c->d_composite.alloc(npx));
    CU_OK(c, c->d_direct_srgb.alloc(npx));
    c->cam.W = (int)W; c->cam.H = (int)H;
    c->ev_recorded = false;
    return RC_OK;
}

rc_status load_host_scene(const std::string& scene_path, uint32_t flags, rc_scene* c)
{
    std::vector<ObjScene> scenes;
    std::optional<Vec3> light;
    LoadError err;
    // light predicate of the reference: |mt| mt.name == "Light" (src/renderer.rs:176)
    if (!ObjScene::load(scene_path, [](const TobjMaterial& m) { return m.name == "Light"; }, scenes, light, err)) {
        c->error = "scene load failed: " + err.message;
        return RC_ERR_SCENE_LOAD;
    }
    if (scenes.empty()) { c->error = "scene has no models (src/renderer.rs:325-329 panics here)"; return RC_ERR_SCENE_LOAD; }
    const bool use_tex = !(flags & RC_CFG_NO_TEXTURES);

    std::vector<float>& verts = c->verts;
    std::vector<uint32_t>&tris = c->tris, &tri_model = c->tri_model;
    std::vector<DMaterial>& mats = c->mats;
    std::vector<DTexture>& tex = c->tex;
    std::vector<uint8_t>& tex_data = c->tex_data;
    uint32_t voff = 0;
    float bmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t m = 0; m < scenes.size(); m++) {
        const ObjScene& sc = scenes[m];
        HostModel hm;
        hm.name = sc.name();
        hm.stream = sc.vertex_stream();
        hm.indices = sc.indices();
        std::optional<Material> mat = sc.material(use_tex);
        UniformMaterial um = to_uniform(mat);
        DMaterial dm;
        memset(&dm, 0, sizeof(dm));
        memcpy(dm.ka, um.ambient, 16); memcpy(dm.kd, um.diffuse, 16); memcpy(dm.ks, um.specular, 16);
        dm.ns = um.shininess;
        dm.tex_c = dm.tex_n = -1;
        if (mat) {
            auto add_tex = [&](const std::shared_ptr<Image>& img) -> int {
                if (!img) return -1;
                DTexture t;
                while (tex_data.size() % 16) tex_data.push_back(0);
                t.offset = tex_data.size(); t.w = img->width; t.h = img->height;
                tex_data.insert(tex_data.end(), img->rgba.begin(), img->rgba.end());
                tex.push_back(t);
                return (int)tex.size() - 1;
            };
            dm.tex_c = add_tex(mat->color_texture);
            dm.tex_n = add_tex(mat->normal_texture);
            dm.ke[0] = mat->emission.x; dm.ke[1] = mat->emission.y; dm.ke[2] = mat->emission.z;
        }
        dm.ebit = (uint32_t)(dm.tex_c >= 0) | ((uint32_t)(dm.tex_n >= 0) << 1);   // src/renderer.rs:422-423
        fill_material_constants(dm);
        mats.push_back(dm);
        memcpy(hm.material80, &um, 64);
        memcpy(&hm.material80[16], &dm.ebit, 4);
        hm.material80[17] = dm.ke[0]; hm.material80[18] = dm.ke[1]; hm.material80[19] = dm.ke[2];
        hm.tex_c = dm.tex_c; hm.tex_n = dm.tex_n;
        const uint32_t nv = (uint32_t)(hm.stream.size() / 17);
        for (uint32_t i = 0; i < nv; i++)
            for (int k = 0; k < 3; k++) {
                bmin[k] = std::fmin(bmin[k], hm.stream[17 * (size_t)i + k]);
                bmax[k] = std::fmax(bmax[k], hm.stream[17 * (size_t)i + k]);
            }
        verts.insert(verts.end(), hm.stream.begin(), hm.stream.end());
        for (size_t t = 0; t + 2 < hm.indices.size(); t += 3) {
            tris.push_back(hm.indices[t] + voff); tris.push_back(hm.indices[t + 1] + voff); tris.push_back(hm.indices[t + 2] + voff);
            tri_model.push_back((uint32_t)m);
        }
        voff += nv;
        c->models.push_back(std::move(hm));
    }
    const uint32_t nt = (uint32_t)tri_model.size();

    // S5 triangle set-up: v0, e1 = v1 - v0, e2 = v2 - v0; skip zero-area line/point triangles
    c->v0.resize(3 * (size_t)nt); c->e1.resize(3 * (size_t)nt); c->e2.resize(3 * (size_t)nt);
    c->skip.assign(nt, 0);
    for (uint32_t t = 0; t < nt; t++) {
        const float* a = &verts[17 * (size_t)tris[3 * (size_t)t]];
        const float* b = &verts[17 * (size_t)tris[3 * (size_t)t + 1]];
        const float* cc = &verts[17 * (size_t)tris[3 * (size_t)t + 2]];
        for (int k = 0; k < 3; k++) { c->v0[3 * (size_t)t + k] = a[k]; c->e1[3 * (size_t)t + k] = b[k] - a[k]; c->e2[3 * (size_t)t + k] = cc[k] - a[k]; }
        c->skip[t] = !memcmp(a, b, 12) || !memcmp(a, cc, 12) || !memcmp(b, cc, 12);
    }
    const float dx = bmax[0] - bmin[0], dy = bmax[1] - bmin[1], dz = bmax[2] - bmin[2];
    c->diag = std::sqrt((dx * dx + dy * dy) + dz * dz);

    rc_scene_info& I = c->info;
    I.num_models = (uint32_t)c->models.size(); I.num_vertices = voff; I.num_triangles = nt;
    I.num_materials = (uint32_t)mats.size();
    I.num_textures = (uint32_t)tex.size();
    I.light_from_obj = light ? 1u : 0u;
    for (int k = 0; k < 3; k++) { I.bbox_min[k] = bmin[k]; I.bbox_max[k] = bmax[k]; }
    if (light) { I.obj_light[0] = light->x; I.obj_light[1] = light->y; I.obj_light[2] = light->z; }
    return RC_OK;
}

rc_status load_scene(rc_ctx* c)
{
    rc_status hs = load_host_scene(c->scene_path, c->cfg.flags, &c->host);
    if (hs != RC_OK) { c->error = c->host.error; return hs; }
    rc_scene& h = c->host;
    const uint32_t nt = h.info.num_triangles;
    const float diag = h.diag;
    Bvh bvh;
    int max_leaf = 3;      // A/B on the B200 (tools/ab_bvh.sh, frame ms for leaf 4 / 3 / 2): living_room 4K 1.206 / 1.182 / 1.185, cube 512 0.176 / 0.172 / 0.172,
                           // sonic 8K 4.70 / 4.69 / 4.75, teapot 1080p 0.529 / 0.538 / 0.526; SAH leaf termination (RC_BVH_NODE_COST) no better
    float node_cost = 0.f;
    if (const char* e = getenv("RC_BVH_LEAF")) max_leaf = atoi(e);
    if (const char* e = getenv("RC_BVH_NODE_COST")) node_cost = (float)atof(e);
    float split_budget = 0.f;
    if (const char* e = getenv("RC_BVH_SPLIT")) split_budget = (float)atof(e);
    build_bvh(h.v0.data(), h.e1.data(), h.e2.data(), h.skip.data(), nt, 1e-4f * diag, bvh, max_leaf, node_cost, split_budget);
    if (bvh.max_depth > 44) { c->error = "BVH deeper than the traversal stack"; return RC_ERR_SCENE_LOAD; }
    h.info.bvh_nodes = (uint32_t)bvh.nodes.size();

    std::vector<float4> nodes(bvh.nodes.size() * 4), geom(bvh.leaf_tris.size() * 3), eg((size_t)nt * 2);
    memcpy(nodes.data(), bvh.nodes.data(), bvh.nodes.size() * sizeof(BvhNode));
    const std::vector<float>&v0 = h.v0, &e1 = h.e1, &e2 = h.e2;
    for (size_t i = 0; i < bvh.leaf_tris.size(); i++) {
        const uint32_t t = bvh.leaf_tris[i];
        float idf;
        memcpy(&idf, &t, 4);
        geom[3 * i] = make_float4(v0[3 * (size_t)t], v0[3 * (size_t)t + 1], v0[3 * (size_t)t + 2], idf);
        geom[3 * i + 1] = make_float4(e1[3 * (size_t)t], e1[3 * (size_t)t + 1], e1[3 * (size_t)t + 2], 0.f);
        geom[3 * i + 2] = make_float4(e2[3 * (size_t)t], e2[3 * (size_t)t + 1], e2[3 * (size_t)t + 2], 0.f);
    }
    for (uint32_t t = 0; t < nt; t++) {
        eg[2 * (size_t)t] = make_float4(e1[3 * (size_t)t], e1[3 * (size_t)t + 1], e1[3 * (size_t)t + 2], 0.f);
        eg[2 * (size_t)t + 1] = make_float4(e2[3 * (size_t)t], e2[3 * (size_t)t + 1], e2[3 * (size_t)t + 2], 0.f);
    }
    std::vector<float> srgb(256);
    for (int i = 0; i < 256; i++) {   // S7: decode table in double
        const double x = i / 255.0;
        srgb[i] = (float)(x <= 0.04045 ? x / 12.92 : std::pow((x + 0.055) / 1.055, 2.4));
    }
    std::vector<DTexture> tex = h.tex;
    std::vector<uint8_t> tex_data = h.tex_data;
    if (tex.empty()) tex.push_back(DTexture{0, 1, 1});
    if (tex_data.empty()) tex_data.assign(16, 0);
    c->n_leaf_tris = (uint32_t)bvh.leaf_tris.size();
    if (geom.empty()) geom.assign(3, make_float4(0.f, 0.f, 0.f, 0.f));

    CU_OK(c, c->d_nodes.upload(nodes));
    CU_OK(c, c->d_tri_geom.upload(geom));
    CU_OK(c, c->d_tri_eg.upload(eg));
    CU_OK(c, c->d_tris.upload(h.tris));
    CU_OK(c, c->d_tri_model.upload(h.tri_model));
    CU_OK(c, c->d_verts.upload(h.verts));
    CU_OK(c, c->d_mats.upload(h.mats));
    CU_OK(c, c->d_tex.upload(tex));
    CU_OK(c, c->d_tex_data.upload(tex_data));
    CU_OK(c, c->d_srgb.upload(srgb));
    c->scene = DScene{c->d_nodes.p, c->d_tri_geom.p, c->d_tri_eg.p, c->d_tris.p, c->d_tri_model.p, c->d_verts.p,
                      c->d_mats.p, c->d_tex.p, c->d_tex_data.p, c->d_srgb.p};

    // S2 / S6 defaults
    c->L0 = c->cfg.interval0 > 0.f ? c->cfg.interval0 : diag / RC_INTERVAL0_DIVISOR;
    c->t_far = c->cfg.t_far > 0.f ? c->cfg.t_far : RC_TFAR_FACTOR * diag;
    c->offset = c->cfg.normal_offset > 0.f ? c->cfg.normal_offset : c->L0 / RC_OFFSET_DIVISOR;
    return RC_OK;
}

rc_status scene_model_stream(const rc_scene& s, std::string& error, uint32_t model, float* vertices, size_t vbytes,
                             uint32_t* indices, size_t ibytes, uint32_t* num_vertices, uint32_t* num_indices)
{
    if (model >= s.models.size()) return RC_ERR_INVALID_ARG;
    const HostModel& m = s.models[model];
    if (num_vertices) *num_vertices = (uint32_t)(m.stream.size() / 17);
    if (num_indices) *num_indices = (uint32_t)m.indices.size();
    if (vertices) {
        if (vbytes < m.stream.size() * 4) { error = "vertex buffer too small"; return RC_ERR_BUFFER_SIZE; }
        memcpy(vertices, m.stream.data(), m.stream.size() * 4);
    }
    if (indices) {
        if (ibytes < m.indices.size() * 4) { error = "index buffer too small"; return RC_ERR_BUFFER_SIZE; }
        memcpy(indices, m.indices.data(), m.indices.size() * 4);
    }
    return RC_OK;
}

void peer_release(rc_ctx* c)
{
    rc_ctx::Peer& p = c->peer;
    for (int r = 0; r < p.world; r++)
        if (r != p.rank && p.base[r]) cudaIpcCloseMemHandle(p.base[r]);
    if (p.local) cudaFree(p.local);
    p = rc_ctx::Peer();
}


void destroy_ctx(rc_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_level) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    c->d_nodes.release(); c->d_tri_geom.release(); c->d_tri_eg.release(); c->d_tris.release(); c->d_tri_model.release();
    c->d_verts.release(); c->d_srgb.release(); c->d_mats.release(); c->d_tex.release(); c->d_tex_data.release();
    c->d_cascade.release(); c->d_origin.release(); c->d_normal.release(); c->d_link_idx.release(); c->d_link_w.release(); c->d_entry.release(); c->d_avg.release(); c->d_need.release(); c->d_list.release(); c->d_list2.release(); c->d_pixmask.release();
    if (c->h_ray_count) cudaFreeHost(c->h_ray_count);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    peer_release(c); c->d_ray_count.release();
    c->d_dirs.release(); c->d_dirq.release(); c->d_qinv.release(); c->d_axis.release(); c->d_occ.release(); c->d_bin_count.release(); c->d_bin_lists.release(); c->d_bin_huge_count.release(); c->d_bin_huge.release(); c->d_depth.release(); c->d_prim.release(); c->d_nrm.release(); c->d_albedo.release();
    c->d_bary.release(); c->d_direct.release(); c->d_irr.release(); c->d_irr2.release();
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_frame_done) cudaEventDestroy(c->ev_frame_done);
    for (auto& e : c->ev_copy_done) if (e) cudaEventDestroy(e);
    c->d_composite.release(); c->d_direct_srgb.release(); c->d_irr_pack.release();
    c->d_dbg_in.release(); c->d_dbg_out.release(); c->d_counters.release();
    delete c;
}

enum { EV_START = 0, EV_GBUF, EV_PROBES, EV_LEVELS, EV_GATHER, EV_MARCH_ALL /* batched path: after k_march_all, before the merges */ };

}  // namespace

extern "C" {

uint32_t rc_abi_version(void) { return RC_ABI_VERSION; }

const char* rc_last_error(const rc_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

rc_status rc_create(const rc_config* cfg, rc_ctx** out)
{
    if (!out) return RC_ERR_INVALID_ARG;
    *out = nullptr;
    if (!cfg || cfg->struct_size != sizeof(rc_config) || !cfg->scene_path || !cfg->width || !cfg->height) {
        g_create_error = "rc_create: bad config (struct_size / scene_path / size)";
        return RC_ERR_INVALID_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        g_create_error = "rc_create: no usable CUDA device (this library has no CPU fallback)";
        return RC_ERR_NO_DEVICE;
    }
    rc_ctx* c = new rc_ctx();
    c->cfg = *cfg;
    c->device = cfg->device;
    std::string path = cfg->scene_path;
    if (!path.empty() && path[0] != '/' && cfg->resource_root && cfg->resource_root[0])
        path = std::string(cfg->resource_root) + "/" + path;   // RESOURCE_PATH.join(obj_path), src/primitives.rs:106
    c->scene_path = path;
    c->cfg.scene_path = nullptr; c->cfg.resource_root = nullptr;
    rc_status st = RC_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) st = RC_ERR_CUDA;
    if (st == RC_OK && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) st = RC_ERR_CUDA;
    if (st == RC_OK) for (auto& e : c->ev) if (cudaEventCreate(&e) != cudaSuccess) { st = RC_ERR_CUDA; break; }
    if (st == RC_OK) {   // lowest priority: an SM-driven read-back must not take SMs from the frame that is being rendered
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, lo) != cudaSuccess) st = RC_ERR_CUDA;
    }
    if (st == RC_OK && cudaEventCreateWithFlags(&c->ev_frame_done, cudaEventDisableTiming) != cudaSuccess) st = RC_ERR_CUDA;
    if (st == RC_OK) for (auto& e : c->ev_copy_done) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { st = RC_ERR_CUDA; break; }
    if (st == RC_OK) for (auto& e : c->ev_level) if (cudaEventCreate(&e) != cudaSuccess) { st = RC_ERR_CUDA; break; }
    if (st != RC_OK) c->error = "CUDA context / stream creation failed";
    try {                       // never let an exception (std::bad_alloc from a hostile asset) cross the C ABI
        if (st == RC_OK) st = load_scene(c);
        if (st == RC_OK) st = setup_frame(c, cfg->width, cfg->height);
    } catch (const std::exception& e) {
        c->error = std::string("rc_create: ") + e.what();
        st = RC_ERR_SCENE_LOAD;
    }
    if (st != RC_OK) { g_create_error = c->error; destroy_ctx(c); return st; }
    {   // tuning knobs (A/B runs): RC_MARCH_MAP = one char per level, L linear / D direction tile / P probe tile
        const char* mm = getenv("RC_MARCH_MAP");
        for (uint32_t i = 0; i < RC_MAX_LEVELS; i++) {
            int def = 1;   // direction tiles
            char ch = (mm && strlen(mm) > i) ? mm[i] : 0;
            c->march_map[i] = ch == 'L' ? 0 : ch == 'D' ? 1 : ch == 'P' ? 2 : def;
        }
        const char* mb = getenv("RC_MARCH_BLOCK");
        if (mb && (atoi(mb) == 64 || atoi(mb) == 128 || atoi(mb) == 256 || atoi(mb) == 512)) c->march_block = atoi(mb);
        if (const char* e = getenv("RC_MARCH_PERSIST")) c->march_persist = atoi(e);
        if (const char* e = getenv("RC_MARCH_OCC")) c->march_occ = atoi(e);
        if (const char* e = getenv("RC_MARCH_COMPACT")) c->march_compact = (int)strtol(e, nullptr, 0);
        if (const char* e = getenv("RC_MARCH_WAVES")) c->march_waves = atoi(e);
        if (const char* e = getenv("RC_MARCH_THRESH")) c->march_thresh = atoi(e) < 1 ? 1 : (atoi(e) > 32 ? 32 : atoi(e));
        if (const char* e = getenv("RC_MARCH_PDL")) c->march_pdl = atoi(e);
        if (const char* e = getenv("RC_MARCH_ENTRY")) c->march_entry = atoi(e) < -1 ? -1 : atoi(e);
        if (const char* e = getenv("RC_MARCH_BATCH")) c->march_batch = atoi(e) > 0 ? 1 : 0;
        if (const char* e = getenv("RC_MARCH_QUAD")) c->march_quad = atoi(e) != 0;
        if (const char* e = getenv("RC_MARCH_QUAD_OCC")) c->march_quad_occ = atoi(e);
        if (const char* e = getenv("RC_CULL")) c->cull = atoi(e) != 0;
        if (const char* e = getenv("RC_NEED_PDL")) c->need_pdl = atoi(e) != 0;
        if (const char* e = getenv("RC_COPY_BLOCKS")) c->copy_blocks = atoi(e) < 0 ? 0 : (atoi(e) > 1024 ? 1024 : atoi(e));
        if (const char* e = getenv("RC_NEED_FUSED")) c->need_fused = atoi(e) != 0;
        if (const char* e = getenv("RC_MARCH_POOL")) c->march_pool = atoi(e) != 0;
        if (const char* e = getenv("RC_MARCH_POOL_THRESH")) c->march_pool_thresh = atoi(e) < 1 ? 1 : (atoi(e) > 32 ? 32 : atoi(e));
        if (const char* e = getenv("RC_LIST_SPLIT")) c->list_split = atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e));
        if (const char* e = getenv("RC_LIST_TILED")) c->list_tiled = atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e));
        if (const char* e = getenv("RC_LIST_DIRMAJOR")) c->list_dir_major = (int)strtol(e, nullptr, 0);
        if (const char* e = getenv("RC_GATHER_MMA")) c->gather_mma = atoi(e) != 0;
        if (const char* e = getenv("RC_GBUFFER_BINNED")) c->gbuffer_binned = atoi(e) != 0;
        if (const char* e = getenv("RC_GATHER_TILES")) c->gather_tiles = atoi(e) < 0 ? 0 : (atoi(e) > 64 ? 64 : atoi(e));
        if (const char* e = getenv("RC_GRAPH")) c->use_graph = atoi(e) != 0;
        cudaDeviceProp prop;
        int bps = march_persist_blocks_per_sm();
        if (cudaGetDeviceProperties(&prop, c->device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
        if (c->sm_count > 0 && bps > 0) c->march_grid = c->sm_count * bps;
        else c->march_persist = 0;
        if (const char* e = getenv("RC_MARCH_GRID")) if (atoi(e) > 0) c->march_grid = atoi(e);
        if (c->d_counters.alloc(RC_MAX_LEVELS) != cudaSuccess) c->march_persist = 0;
    }
    c->lights.n = 1;    // AppState::light_position default [0,0,0] (src/app.rs)
    c->lights.flags = RC_UPD_ENABLE_NORMAL_MAP;
    *out = c;
    return RC_OK;
}

void rc_destroy(rc_ctx* ctx) { destroy_ctx(ctx); }

rc_status rc_update(rc_ctx* c, const rc_camera* cam, const rc_light* lights, uint32_t n_lights, uint32_t flags)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if (!cam || (n_lights && !lights) || n_lights > RC_MAX_LIGHTS) { c->error = "rc_update: bad arguments"; return RC_ERR_INVALID_ARG; }
    if (!primary_basis(*cam, c->cam)) { c->error = "rc_update: singular view_proj matrix"; return RC_ERR_INVALID_ARG; }
    c->cam.W = (int)c->W; c->cam.H = (int)c->H;
    c->cam.clip = (c->cfg.flags & RC_CFG_RASTER_CLIP) ? 1 : 0;
    {   // screen-space rectangle of the scene's bounding box (grown by 1e-3 of the diagonal and by two pixels), in double
        const rc_scene_info& I = c->host.info;
        double x0 = 1e30, y0 = 1e30, x1 = -1e30, y1 = -1e30;
        bool whole = false;
        const double pad = 1e-3 * (double)c->host.diag;
        for (int k = 0; k < 8 && !whole; k++) {
            const double p[3] = {(k & 1 ? I.bbox_max[0] + pad : I.bbox_min[0] - pad), (k & 2 ? I.bbox_max[1] + pad : I.bbox_min[1] - pad),
                                 (k & 4 ? I.bbox_max[2] + pad : I.bbox_min[2] - pad)};
            const float* M = cam->view_proj;
            const double cx = M[0] * p[0] + M[4] * p[1] + M[8] * p[2] + M[12], cy = M[1] * p[0] + M[5] * p[1] + M[9] * p[2] + M[13];
            const double cw = M[3] * p[0] + M[7] * p[1] + M[11] * p[2] + M[15];
            if (!(cw > 1e-6 * (std::fabs(cx) + std::fabs(cy) + 1.0))) { whole = true; break; }
            const double sx = (cx / cw * 0.5 + 0.5) * c->W - 0.5, sy = (1.0 - (cy / cw * 0.5 + 0.5)) * c->H - 0.5;
            x0 = std::min(x0, sx); x1 = std::max(x1, sx); y0 = std::min(y0, sy); y1 = std::max(y1, sy);
        }
        if (whole || !(x0 == x0) || !(y0 == y0) || !(x1 == x1) || !(y1 == y1)) { c->cam.sb_x0 = 0; c->cam.sb_y0 = 0; c->cam.sb_x1 = (int)c->W - 1; c->cam.sb_y1 = (int)c->H - 1; }
        else {
            auto cl = [](double v, int hi) { return (int)std::min(std::max(v, -1.0), (double)hi); };
            c->cam.sb_x0 = cl(std::floor(x0) - 2, (int)c->W); c->cam.sb_y0 = cl(std::floor(y0) - 2, (int)c->H);
            c->cam.sb_x1 = cl(std::ceil(x1) + 2, (int)c->W); c->cam.sb_y1 = cl(std::ceil(y1) + 2, (int)c->H);
        }
    }
    c->lights.n = (int)n_lights;
    for (uint32_t i = 0; i < n_lights; i++) memcpy(c->lights.pos[i], lights[i].position, 16);
    c->lights.flags = flags;
    c->have_camera = true;
    return RC_OK;
}

rc_status rc_resize(rc_ctx* c, uint32_t width, uint32_t height)
{
    if (!c || !width || !height) return RC_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->cfg.width = width; c->cfg.height = height;
    c->cfg.tile_w = c->cfg.tile_h = 0;   // a resize resets the tile to the full frame
    c->have_camera = false;              // aspect changed: the caller must rc_update (Projection::resize)
    // the peer exchange buffers were sized for the old frame (rc_peer_export) and every rank's mapping of them is
    // stale: detach, so that the gather cannot store a larger frame into them; the caller re-exports / re-attaches
    peer_release(c);
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    return setup_frame(c, width, height);
}

rc_status rc_set_tile(rc_ctx* c, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if ((w == 0) != (h == 0) || (w && (x0 + w > c->W || y0 + h > c->H))) { c->error = "rc_set_tile: tile outside the frame"; return RC_ERR_INVALID_ARG; }
    cudaSetDevice(c->device);
    // the buffers are re-laid-out in place: every frame in flight must have finished with them
    if (c->last_stream) CU_OK(c, cudaStreamSynchronize(c->last_stream));
    CU_OK(c, cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CU_OK(c, cudaStreamSynchronize(c->copy_stream));
    c->cfg.tile_x0 = x0; c->cfg.tile_y0 = y0; c->cfg.tile_w = w; c->cfg.tile_h = h;
    c->composite_valid = c->direct_valid = false;
    return setup_frame(c, c->W, c->H, true);      // camera, lights, peers and tuning are untouched
}

// The k_need launches of one frame.  append = true: masks are consumed into the ray lists (and pushed up) level by level —
// the single-pass form.  append = false: masks only (first pass of the halo exchange).
static int4 owned_rect(const rc_ctx* c, uint32_t i);

// pass = 0: the single-pass form (masks consumed into the ray lists and pushed up, level by level);
// pass = 1: masks only (first pass of the halo exchange);  pass = 2: lists of the owned probes from the completed masks
static rc_status enqueue_need_chain(rc_ctx* c, cudaStream_t st, int pass)
{
    const uint32_t n_lists = c->top_fillable() ? c->N - 1 : c->N;
    const bool append = pass != 1, push = pass != 2;
    auto has_upper_of = [&](uint32_t i) { return (push && i + 1 < n_lists) ? (i == 0 ? 1 : 2) : 0; };
    auto own_of = [&](uint32_t i) { return (pass == 2 && i >= 1) ? owned_rect(c, i) : make_int4(0, 0, -1, -1); };
    const uint32_t fused_from = (c->need_fused && n_lists >= 3) ? 1u : n_lists;      // levels >= fused_from go into ONE cluster launch
    for (uint32_t i = 0; i < n_lists && i < fused_from; i++) {
        const DLevel& L = c->levels[i];
        const int has_upper = has_upper_of(i);
        const int up_res = i + 1 < c->N ? c->need_res[i + 1] : 0;
        launch_need(L, c->need_res[i], has_upper, (up_res * up_res + 31) / 32, c->d_origin.p + L.probe_offset,
                    c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset, c->d_need.p + c->need_offset[i],
                    has_upper ? c->d_need.p + c->need_offset[i + 1] : nullptr, c->d_list.p + c->list_offset[i],
                    c->d_ray_count.p + i, append && i >= 1, c->need_pdl && i >= 1 && pass == 0, c->need_pdl && i + 1 < n_lists && pass == 0,
                    ((c->list_dir_major >> i) & 1) != 0, i >= 1 ? c->list_tiled : 0, append, own_of(i), st);
        c->launches++;
    }
    if (fused_from < n_lists) {
        NeedChain ch{};
        ch.first = (int)fused_from; ch.last = (int)n_lists - 1; ch.append = append ? 1 : 0;
        for (uint32_t i = fused_from; i < n_lists; i++) {
            const int up_res = i + 1 < c->N ? c->need_res[i + 1] : 0;
            ch.lv[i] = c->levels[i];
            ch.Dr[i] = c->need_res[i];
            ch.has_upper[i] = has_upper_of(i);
            ch.up_words[i] = (up_res * up_res + 31) / 32;
            ch.clear[i] = append ? 1 : 0;
            ch.dir_major[i] = (c->list_dir_major >> i) & 1;
            ch.tile_order[i] = (c->list_tiled && ch.has_upper[i] != 1 && (ch.Dr[i] == 32 || (c->list_tiled > 1 && (ch.Dr[i] == 8 || ch.Dr[i] == 16)))) ? 1 : 0;
            ch.need_off[i] = c->need_offset[i];
            ch.need_up_off[i] = i + 1 < c->N ? c->need_offset[i + 1] : c->need_offset[i];
            ch.list_off[i] = c->list_offset[i];
            ch.own[i] = own_of(i);
        }
        launch_need_chain(ch, c->d_origin.p, c->d_link_idx.p, c->d_link_w.p, c->d_need.p, c->d_list.p, c->d_ray_count.p, st);
        c->launches++;
    }
    if (append) {
        // split lists (k_split): one launch over every level chosen for this frame, sized like the march grids from the last
        // list lengths that arrived (grid-stride inside a level's blocks; the whole capacity while nothing has arrived)
        SplitPlan plan{};
        c->split_mask = 0;
        bool chosen[RC_MAX_LEVELS] = {};
        bool measuring = c->list_split >= 2 || c->split_measure_frame();
        unsigned long long expected_misses = 0;
        for (uint32_t i = 0; i < n_lists; i++) {
            chosen[i] = c->split_level(i);
            if (chosen[i] && c->h_ray_count[i] == 0xffffffffu) measuring = true;      // nothing measured yet
            if (chosen[i]) expected_misses += c->split_triv[i];
        }
        const bool worth_it = measuring || expected_misses >= rc_ctx::kSplitMinMisses;
        for (uint32_t i = 0; i < n_lists; i++) {
            if (!chosen[i] || !worth_it) continue;
            const size_t cap = c->list_offset[i + 1] - c->list_offset[i];
            const unsigned int prev = c->list_len(i);
            const double entries = prev == 0xffffffffu ? (double)cap : std::min((double)prev * 1.25 + 4096.0, (double)cap);
            const int k = plan.n++;
            SplitJob& jb = plan.job[k];
            jb.lv = c->levels[i];
            jb.level = (int)i;
            jb.cap = (unsigned)cap;
            jb.root = c->scene.nodes;
            jb.origin = c->d_origin.p + c->levels[i].probe_offset;
            jb.dirq = c->d_dirq.p + 2 * (c->dir_offset[i] / 3);
            jb.qinv = c->d_qinv.p + c->qinv_offset[i];
            jb.list_a = c->d_list.p + c->list_offset[i];
            jb.list_b = c->d_list2.p + c->list_offset[i];
            jb.counts = c->d_ray_count.p;
            plan.block_off[k + 1] = plan.block_off[k] + (unsigned)std::min(entries / (double)split_chunk() + 1.0, 1.0e6);
            c->split_mask |= 1u << i;
        }
        if (plan.n) {
            launch_split(plan, st);
            c->launches++;
        }
    }
    return RC_OK;
}

// probes of level i this context OWNS: those whose anchor pixel (S1) lies inside the tile; inclusive sub-grid coordinates,
// z < x when none
static int4 owned_rect(const rc_ctx* c, uint32_t i)
{
    const DLevel& L = c->levels[i];
    auto range = [](int t0, int tn, int P, int full, int sub0, int subn, int& lo, int& hi) {
        // anchor(p) = min(p*P + P/2, full - 1) in [t0, t0 + tn)
        lo = 1; hi = 0;
        for (int p = sub0; p < sub0 + subn; p++) {
            const int a = std::min(p * P + P / 2, full - 1);
            if (a >= t0 && a < t0 + tn) { if (hi < lo) lo = p - sub0; hi = p - sub0; }
        }
    };
    int x0, x1, y0, y1;
    range(c->tile.x0, c->tile.w, L.P, (int)c->W, L.px0, L.sw, x0, x1);
    range(c->tile.y0, c->tile.h, L.P, (int)c->H, L.py0, L.sh, y0, y1);
    if (x1 < x0 || y1 < y0) return make_int4(1, 1, 0, 0);
    return make_int4(x0, y0, x1, y1);
}

// stage-timing events: inside a stream capture they must become EXTERNAL event-record nodes, otherwise the host
// cannot synchronise on / time them (cudaErrorInvalidValue)
static cudaError_t record_event(rc_ctx* c, cudaEvent_t ev, cudaStream_t st)
{
    return cudaEventRecordWithFlags(ev, st, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

rc_status rc_render_begin(rc_ctx* c, void* stream)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if (!c->have_camera) { c->error = "rc_render before rc_update"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    c->last_stream = st;
    c->launches = 0;
    if (c->march_persist) CU_OK(c, cudaMemsetAsync(c->d_counters.p, 0, RC_MAX_LEVELS * sizeof(unsigned int), st));
    c->composite_valid = false;
    c->direct_valid = false;
    c->frame_batched = false;
    c->cam_rendered = c->cam;
    c->lights_rendered = c->lights;
    CU_OK(c, record_event(c, c->ev[EV_START], st));
    c->frame_culled = c->cull_possible();
    if (c->frame_open && c->frame_culled)           // the last frame never reached its gather: its list lengths are still set
        CU_OK(c, cudaMemsetAsync(c->d_ray_count.p, 0, 3 * RC_MAX_LEVELS * sizeof(unsigned int), st));
    c->frame_open = true;
    c->frame_id++;
    if (c->frame_id == 0) c->frame_id = 1;
    GBufferOut gb{c->d_depth.p, c->d_prim.p, c->d_nrm.p, c->d_bary.p, (c->cfg.flags & RC_CFG_FLOATING_PROBES) ? c->d_occ.p : nullptr, c->frame_id,
                  (c->tile.w + 31) / 32};
    if (c->gbuffer_binned) {
        launch_gbuffer_binned(c->scene, c->cam, c->lights, c->tile, gb, c->levels[0].D * c->levels[0].D, c->d_dirs.p,
                              c->frame_culled ? c->d_pixmask.p : nullptr, c->n_leaf_tris, c->d_bin_count.p, c->d_bin_lists.p,
                              c->d_bin_huge_count.p, c->d_bin_huge.p, st);
        c->launches += c->n_leaf_tris ? 2 : 1;
    } else {
        launch_gbuffer(c->scene, c->cam, c->lights, c->tile, gb, c->levels[0].D * c->levels[0].D, c->d_dirs.p,
                       c->frame_culled ? c->d_pixmask.p : nullptr, st);
        c->launches++;
    }
    CU_OK(c, record_event(c, c->ev[EV_GBUF], st));
    DLevelSet ls;
    ls.n = (int)c->N;
    for (uint32_t i = 0; i < c->N; i++) ls.lv[i] = c->levels[i];
    const DLevel& top = c->levels[c->N - 1];
    const unsigned n_probes = top.probe_offset + (unsigned)(top.sw * top.sh);
    launch_probes(c->scene, c->cam, ls, n_probes, c->tile, c->offset, c->d_depth.p, c->d_prim.p, c->d_origin.p, c->d_normal.p,
                  c->frame_culled ? c->d_pixmask.p : nullptr, c->d_need.p, c->d_occ.p, c->frame_id, (c->tile.w + 31) / 32,
                  (c->cfg.flags & RC_CFG_FLOATING_PROBES) != 0, st);
    c->launches++;
    {
        const int ne = c->march_persist ? 0 : c->entry_levels();
        if (c->N > 1 || ne > 0) {
            launch_link_entry(c->scene, ls, top.probe_offset, ne, c->d_origin.p, c->d_normal.p, c->d_link_idx.p, c->d_link_w.p, c->d_entry.p, st);
            c->launches++;
        }
    }
    if (c->frame_culled) {
        // request masks bottom-up + one ray list per level; a top level that is filled needs neither.
        // Halo exchange (RC_CFG_HALO_EXCHANGE): this pass only propagates the masks — the lists are built by rc_render_lists
        // after the caller has sent the requests of the probes this rank does not own to their owners
        rc_status s = enqueue_need_chain(c, st, c->exchange_active() ? 1 : 0);
        if (s != RC_OK) return s;
        c->lists_pending = c->exchange_active();
    }
    CU_OK(c, record_event(c, c->ev[EV_PROBES], st));
    CU_OK(c, cudaGetLastError());
    return RC_OK;
}

rc_status rc_render_lists(rc_ctx* c, void* stream)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if (!c->frame_open || !c->lists_pending) { c->error = "rc_render_lists: only after rc_render_begin of a halo-exchange context"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    // second pass over the (now complete) masks: list the requests of the probes this rank owns, clear everything.
    // Level 0 is never exchanged: every level-0 probe of the sub-grid (the tile's probes and their one-probe ring) is marched here.
    rc_status sl = enqueue_need_chain(c, st, 2);
    if (sl != RC_OK) return sl;
    c->lists_pending = false;
    CU_OK(c, cudaGetLastError());
    return RC_OK;
}

rc_status rc_exchange_level_info(rc_ctx* c, uint32_t level, rc_exchange_info* out)
{
    if (!c || !out || level >= c->N) return RC_ERR_INVALID_ARG;
    const DLevel& L = c->levels[level];
    memset(out, 0, sizeof(*out));
    const int Dr = c->need_res[level];
    out->need_ptr = c->d_need.p + c->need_offset[level];
    out->need_words_per_probe = (uint32_t)((Dr * Dr + 31) / 32);
    out->avg_ptr = c->avg_of(level);
    out->avg_float4_per_probe = level >= 1 ? (uint32_t)((L.D * L.D) / 4) : 0u;
    out->px0 = L.px0; out->py0 = L.py0; out->sub_w = (uint32_t)L.sw; out->sub_h = (uint32_t)L.sh;
    const int4 o = owned_rect(c, level);
    out->own_x0 = o.x; out->own_y0 = o.y; out->own_x1 = o.z; out->own_y1 = o.w;
    out->exchanged = (c->exchange_active() && level >= 1 && !(level + 1 == c->N && c->top_fillable())) ? 1u : 0u;
    return RC_OK;
}

rc_status rc_render_level(rc_ctx* c, uint32_t level, void* stream)
{
    if (!c || level >= c->N) return RC_ERR_INVALID_ARG;
    if (c->lists_pending) { c->error = "rc_render_level: halo-exchange context — call rc_render_lists after exchanging the request masks"; return RC_ERR_STATE; }
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const bool fused = !(c->cfg.flags & RC_CFG_SEPARATE_MERGE);
    const bool top = (level == c->N - 1);
    const DLevel& L = c->levels[level];
    const DLevel* U = top ? nullptr : &c->levels[level + 1];
    const float3 sky = make_float3(c->cfg.sky[0], c->cfg.sky[1], c->cfg.sky[2]);
    uint2* tex = c->d_cascade.p + L.texel_offset;
    const float4* up = top ? nullptr : c->avg_of(level + 1);   // the merge reads level i+1 through its child averages
    float4* my_avg = c->avg_of(level);                          // ... and level i-1 will read this level's
    // culled frames never materialise a top level that cannot hit anything: level N-2 evaluates its far field
    // from the top probes' validity alone (far_field's up_const)
    const bool const_top = c->frame_culled && c->top_fillable();
    if (top && const_top) {
        if (c->level_timing) CU_OK(c, cudaEventRecord(c->ev_level[level], st));
        return RC_OK;
    }
    const bool up_const = const_top && level + 2 == c->N;
    if (up_const) up = c->d_origin.p + U->probe_offset;
    if (top && c->top_fillable()) {
        launch_fill_top(L, sky, c->d_origin.p + L.probe_offset, tex, my_avg, st);
        c->launches++;
        if (c->level_timing) CU_OK(c, cudaEventRecord(c->ev_level[level], st));
        CU_OK(c, cudaGetLastError());
        return RC_OK;
    }
    const bool compact = ((c->march_compact >> level) & 1) != 0;
    int eff_map = c->march_map[level];
    if (eff_map == 1 && (L.D & 7)) eff_map = 0;   // as launch_march: the 8x4 direction tile needs D % 8 == 0
    // the march kernel leaves the child averages itself when it finalises the level and the 2x2 children share a warp
    const bool culled = c->frame_culled && !c->march_persist && !compact && fused;
    int list_blocks = 0;   // 0 = the upper bound (every texel)
    if (culled && c->h_ray_count) {
        const unsigned int prev = c->list_len(level);   // entries of the last list that arrived
        if (prev != 0xffffffffu) {
            const double threads = (double)prev * (level >= 1 ? 4.0 : 1.0) * 1.25;
            list_blocks = (int)std::min(threads / 128.0 + 64.0, 2.0e9);
        }
    }
    const bool avg_in_kernel = my_avg && (fused || top) && !c->march_persist && !compact && (culled || march_avg_ystep(L.D, eff_map) != 0);
    const bool quad_kernel = culled && level >= 1 && !top && c->march_quad;
    const bool split = culled && ((c->split_mask >> level) & 1u) != 0;      // this frame's k_split partitioned the level's list
    const bool pool_kernel = culled && level >= 1 && !top && c->march_pool && !quad_kernel;
    if (pool_kernel) {
        int pblocks = 0;
        if (c->h_ray_count) {
            const unsigned int prev = c->list_len(level);
            if (prev != 0xffffffffu) pblocks = (int)std::min((double)prev * 4.0 * 1.25 / 1024.0 + 16.0, 2.0e9);
        }
        launch_march_pool(c->scene, c->lights, L, sky, c->d_origin.p + L.probe_offset, c->d_dirq.p + 2 * (c->dir_offset[level] / 3), tex, up,
                          c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset,
                          (int)level < c->entry_levels() ? c->d_entry.p + 2 * (size_t)L.probe_offset : nullptr, avg_in_kernel ? my_avg : nullptr,
                          fused, c->march_occ, c->march_pdl != 0, pblocks, (split ? c->d_list2.p : c->d_list.p) + c->list_offset[level],
                          c->d_ray_count.p + (split ? 2 * RC_MAX_LEVELS : 0) + level, split ? c->d_ray_count.p + RC_MAX_LEVELS + level : nullptr,
                          (unsigned)(c->list_offset[level + 1] - c->list_offset[level]), c->march_pool_thresh, up_const, st);
    } else if (quad_kernel) {
        int qblocks = 0;
        if (c->h_ray_count) {
            const unsigned int prev = c->list_len(level);
            if (prev != 0xffffffffu) qblocks = (int)std::min((double)prev * 1.25 / 64.0 + 64.0, 2.0e9);
        }
        launch_march_quad(c->scene, c->lights, L, sky, c->d_origin.p + L.probe_offset, c->d_dirq.p + 2 * (c->dir_offset[level] / 3), tex, up,
                          c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset, my_avg, fused, c->march_quad_occ, c->march_pdl != 0,
                          qblocks, c->d_list.p + c->list_offset[level], c->d_ray_count.p + level, up_const, st);
    } else if (c->march_persist)
        launch_march_persist(c->scene, c->lights, L, U, top, sky, c->d_origin.p + L.probe_offset, c->d_dirs.p + c->dir_offset[level], tex, up,
                             c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset, fused, c->march_map[level], c->march_thresh,
                             c->march_grid, c->d_counters.p + level, c->march_pdl && fused && !top, st);
    else
        launch_march(c->scene, c->lights, L, U, top, sky, c->d_origin.p + L.probe_offset, c->d_dirs.p + c->dir_offset[level],
                     c->d_dirq.p + 2 * (c->dir_offset[level] / 3), tex, up,
                     c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset,
                     (int)level < c->entry_levels() ? c->d_entry.p + 2 * (size_t)L.probe_offset : nullptr,
                     avg_in_kernel ? my_avg : nullptr, fused, c->march_map[level], c->march_occ, c->march_pdl != 0, compact,
                     culled ? list_blocks : (c->march_waves > 0 ? c->sm_count * c->march_occ * c->march_waves : 0),
                     culled ? (split ? c->d_list2.p : c->d_list.p) + c->list_offset[level] : nullptr,
                     c->d_ray_count.p + (split ? 2 * RC_MAX_LEVELS : 0) + level, level >= 1 ? 1 : 0, up_const,
                     split ? c->d_ray_count.p + RC_MAX_LEVELS + level : nullptr, (unsigned)(c->list_offset[level + 1] - c->list_offset[level]), st);
    c->launches++;
    if (!fused && !top) {
        launch_merge(L, *U, sky, c->d_origin.p + L.probe_offset, tex, up, c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset, st);
        c->launches++;
    }
    if (my_avg && !avg_in_kernel) {
        launch_child_avg(L, tex, my_avg, st);
        c->launches++;
    }
    // per-level events sit between the level kernels and would defeat their PDL overlap: opt-in only
    if (c->level_timing) CU_OK(c, cudaEventRecord(c->ev_level[level], st));
    CU_OK(c, cudaGetLastError());
    return RC_OK;
}

rc_status rc_set_tuning(rc_ctx* c, const char* key, int value)
{
    if (!c || !key) return RC_ERR_INVALID_ARG;
    const std::string k = key;
    if (k == "level_timing") c->level_timing = value != 0;
    else if (k == "march_persist") c->march_persist = (value != 0 && c->march_grid > 0 && c->d_counters.p) ? 1 : 0;
    else if (k == "march_thresh") c->march_thresh = value < 1 ? 1 : (value > 32 ? 32 : value);
    else if (k == "march_pdl") c->march_pdl = value != 0;
    else if (k == "march_occ" && value >= 8 && value <= 16) c->march_occ = value;
    else if (k == "march_compact" && value >= 0) c->march_compact = value;
    else if (k == "march_waves" && value >= 0) c->march_waves = value;
    else if (k == "fill_top") c->fill_top = value != 0;
    else if (k == "march_entry" && value >= -1) c->march_entry = value;
    else if (k == "march_batch" && value >= 0 && value <= 1) c->march_batch = value;
    else if (k == "march_quad" && value >= 0 && value <= 1) c->march_quad = value;
    else if (k == "march_quad_occ" && value >= 8 && value <= 16) c->march_quad_occ = value;
    else if (k == "cull" && value >= 0 && value <= 1) c->cull = value;
    else if (k == "gather_tiles" && value >= 0 && value <= 64) c->gather_tiles = value;
    else if (k == "gather_mma" && value >= 0 && value <= 1) c->gather_mma = value;
    else if (k == "gbuffer_binned" && value >= 0 && value <= 1) c->gbuffer_binned = value;
    else if (k == "peer_stores" && value >= 0 && value <= 1) c->peer_stores = value;
    else if (k == "peer_broadcast" && value >= 0 && value <= 1) c->peer_broadcast = value;
    else if (k == "need_pdl" && value >= 0 && value <= 1) c->need_pdl = value;
    else if (k == "copy_blocks" && value >= 0 && value <= 1024) c->copy_blocks = value;
    else if (k == "list_dir_major" && value >= 0) c->list_dir_major = value;
    else if (k == "list_tiled" && value >= 0 && value <= 2) c->list_tiled = value;
    else if (k == "list_split" && value >= 0 && value <= 2) c->list_split = value;
    else if (k == "march_pool" && value >= 0 && value <= 1) c->march_pool = value;
    else if (k == "march_pool_thresh" && value >= 1 && value <= 32) c->march_pool_thresh = value;
    else if (k == "need_fused" && value >= 0 && value <= 1) c->need_fused = value;
    else if (k == "graph" && value >= 0 && value <= 1) c->use_graph = value;
    else if (k == "march_block" && (value == 64 || value == 128 || value == 256 || value == 512)) c->march_block = value;
    else if (k == "march_grid" && value > 0) c->march_grid = value;
    else if (k.rfind("march_map", 0) == 0 && k.size() == 10 && k[9] >= '0' && k[9] <= '9' && value >= 0 && value <= 2)
        c->march_map[k[9] - '0'] = value;
    else { c->error = "rc_set_tuning: unknown key or bad value: " + k; return RC_ERR_INVALID_ARG; }
    return RC_OK;
}

// the irradiance is double-buffered for rc_read_target_async: flip, and wait for a read-back still using the new slot
static rc_status next_output_slot(rc_ctx* c, cudaStream_t st)
{
    c->irr_slot ^= 1;
    if (c->copy_pending[c->irr_slot]) {   // an asynchronous read-back of this buffer may still be in flight
        CU_OK(c, cudaStreamWaitEvent(st, c->ev_copy_done[c->irr_slot], 0));
        c->copy_pending[c->irr_slot] = false;
    }
    return RC_OK;
}

rc_status rc_render_end(rc_ctx* c, void* stream)
{
    if (!c) return RC_ERR_INVALID_ARG;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CU_OK(c, record_event(c, c->ev[EV_LEVELS], st));
    if (!c->capturing) {                  // (a captured frame did this before the capture began)
        rc_status s = next_output_slot(c, st);
        if (s != RC_OK) return s;
    }
    const DLevel& L0 = c->levels[0];
    PeerOut po{};
    if (c->peer.world && c->peer_stores) {
        if ((size_t)c->W * c->H * sizeof(uint2) > c->peer.slot_bytes) { c->error = "peer frame buffers are smaller than the frame (re-export after a resize)"; return RC_ERR_STATE; }
        c->peer.seq++;
        po.world = c->peer.world; po.rank = c->peer.rank; po.W = (int)c->W; po.seq = c->peer.seq;
        po.dst_mask = c->peer_broadcast ? ((1u << c->peer.world) - 1u) : 1u;
        for (int r = 0; r < c->peer.world; r++) { po.frame[r] = c->peer.frame(r, c->peer.seq); po.ctrl[r] = c->peer.ctrl(r); }
        launch_peer_begin(po, c->peer.ctrl(c->peer.rank), st);
        c->launches++;
    }
    launch_gather(c->cam, L0, c->tile, c->d_origin.p + L0.probe_offset, c->d_cascade.p + L0.texel_offset, c->d_dirs.p,
                  c->d_depth.p, c->d_nrm.p, c->irr(), c->d_ray_count.p, c->frame_culled ? c->h_ray_count : nullptr, po, c->gather_tiles_eff(),
                  GatherMma{c->gather_mma, c->gather_sym ? 1 : 0, c->dirs_host[0].data(), c->d_axis.p, c->d_axis.p + c->W}, st);
    c->launches++;
    if (po.world) {
        launch_peer_publish(po, st);
        c->launches++;
    }
    CU_OK(c, record_event(c, c->ev[EV_GATHER], st));
    CU_OK(c, cudaGetLastError());
    c->frame_open = false;
    c->ev_recorded = true;
    return RC_OK;
}

// small frames: every level's march in one launch, then the merges top-down (rc_ctx::march_batch)
static rc_status render_levels_batched(rc_ctx* c, void* stream)
{
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const float3 sky = make_float3(c->cfg.sky[0], c->cfg.sky[1], c->cfg.sky[2]);
    DLevelSet ls;
    ls.n = (int)c->N;
    for (uint32_t i = 0; i < c->N; i++) ls.lv[i] = c->levels[i];
    int order[RC_MAX_LEVELS], n = 0;
    int first = (int)c->N - 1;
    if (c->top_fillable()) {
        const DLevel& T = c->levels[c->N - 1];
        launch_fill_top(T, sky, c->d_origin.p + T.probe_offset, c->d_cascade.p + T.texel_offset, c->avg_of(c->N - 1), st);
        c->launches++;
        first--;
    }
    for (int i = first; i >= 0; i--) order[n++] = i;
    if (n) {
        launch_march_all(c->scene, c->lights, ls, order, n, c->march_map, c->dir_offset.data(), c->entry_levels(), (int)c->N - 1, sky,
                         c->d_origin.p, c->d_dirs.p, c->d_cascade.p, c->d_entry.p, c->march_occ, st);
        c->launches++;
    }
    CU_OK(c, record_event(c, c->ev[EV_MARCH_ALL], st));
    if (first == (int)c->N - 1 && c->N > 1) {   // a marched top level is final as it is
        launch_child_avg(c->levels[c->N - 1], c->d_cascade.p + c->levels[c->N - 1].texel_offset, c->avg_of(c->N - 1), st);
        c->launches++;
    }
    for (int i = (int)c->N - 2; i >= 0; i--) {
        const DLevel &L = c->levels[i], &U = c->levels[i + 1];
        launch_merge(L, U, sky, c->d_origin.p + L.probe_offset, c->d_cascade.p + L.texel_offset, c->avg_of(i + 1),
                     c->d_link_idx.p + L.probe_offset, c->d_link_w.p + L.probe_offset, st);
        c->launches++;
        if (i >= 1) {
            launch_child_avg(L, c->d_cascade.p + L.texel_offset, c->avg_of(i), st);
            c->launches++;
        }
    }
    CU_OK(c, cudaGetLastError());
    return RC_OK;
}

static rc_status render_enqueue(rc_ctx* c, void* stream);

rc_status rc_render(rc_ctx* c, void* stream)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if (c->exchange_active()) {
        c->error = "rc_render: a halo-exchange context renders through rc_render_begin / rc_render_lists / rc_render_level / rc_render_end "
                   "with the caller's exchanges in between (include/rc_b200.h)";
        return RC_ERR_STATE;
    }
    if (!c->use_graph || c->level_timing || !c->have_camera) return render_enqueue(c, stream);
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    rc_status s = next_output_slot(c, st);     // waits on an event from outside the capture: must precede it
    if (s != RC_OK) return s;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        c->use_graph = 0;
        c->irr_slot ^= 1;                      // render_enqueue flips again
        return render_enqueue(c, stream);
    }
    c->capturing = true;
    s = render_enqueue(c, stream);
    c->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (s != RC_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (s != RC_OK) return s;
        c->use_graph = 0;                      // this driver / stream cannot capture the frame: plain launches from now on
        c->irr_slot ^= 1;
        return render_enqueue(c, stream);
    }
    if (c->graph_exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(c->graph_exec, graph, &info) != cudaSuccess) {   // topology changed (tuning, resize, ...)
            cudaGetLastError();
            cudaGraphExecDestroy(c->graph_exec);
            c->graph_exec = nullptr;
        }
    }
    if (!c->graph_exec && cudaGraphInstantiate(&c->graph_exec, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphDestroy(graph);
        c->graph_exec = nullptr;
        c->use_graph = 0;
        c->irr_slot ^= 1;
        return render_enqueue(c, stream);
    }
    cudaGraphDestroy(graph);
    CU_OK(c, cudaGraphLaunch(c->graph_exec, st));
    return RC_OK;
}

static rc_status render_enqueue(rc_ctx* c, void* stream)
{
    rc_status s = rc_render_begin(c, stream);
    if (s != RC_OK) return s;
    c->frame_batched = c->batched();
    if (c->frame_batched) {
        s = render_levels_batched(c, stream);
        if (s != RC_OK) return s;
        return rc_render_end(c, stream);
    }
    for (int i = (int)c->N - 1; i >= 0; i--) {
        s = rc_render_level(c, (uint32_t)i, stream);
        if (s != RC_OK) return s;
    }
    return rc_render_end(c, stream);
}

rc_status rc_synchronize(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    if (c->last_stream) CU_OK(c, cudaStreamSynchronize(c->last_stream));
    CU_OK(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

rc_status rc_target_bytes(rc_ctx* c, rc_target which, size_t* bytes)
{
    if (!c || !bytes) return RC_ERR_INVALID_ARG;
    const size_t npx = (size_t)c->tile.w * c->tile.h;
    switch ((int)which) {
    case RC_TARGET_IRRADIANCE: case RC_TARGET_DIRECT: case RC_TARGET_ALBEDO: *bytes = npx * 8; return RC_OK;
    case RC_TARGET_IRRADIANCE_RGB48: *bytes = npx * 6; return RC_OK;
    case RC_TARGET_DEPTH: case RC_TARGET_NORMAL: case RC_TARGET_PRIM: case RC_TARGET_COMPOSITE: case RC_TARGET_DIRECT_SRGB8:
        *bytes = npx * 4; return RC_OK;
    default: break;
    }
    if ((int)which >= RC_TARGET_CASCADE0 && (uint32_t)((int)which - RC_TARGET_CASCADE0) < c->N) {
        *bytes = (size_t)c->level_info[(int)which - RC_TARGET_CASCADE0].texel_count * 8;
        return RC_OK;
    }
    c->error = "unknown target";
    return RC_ERR_INVALID_ARG;
}

rc_status rc_read_target(rc_ctx* c, rc_target which, void* host_dst, size_t bytes)
{
    if (!c || !host_dst) return RC_ERR_INVALID_ARG;
    size_t need = 0;
    rc_status s = rc_target_bytes(c, which, &need);
    if (s != RC_OK) return s;
    if (bytes < need) { c->error = "rc_read_target: buffer too small"; return RC_ERR_BUFFER_SIZE; }
    cudaSetDevice(c->device);
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    const void* src = nullptr;
    const int w_ = (int)which;
    if ((w_ == RC_TARGET_DIRECT || w_ == RC_TARGET_ALBEDO || w_ == RC_TARGET_COMPOSITE || w_ == RC_TARGET_DIRECT_SRGB8) && !c->direct_valid) {
        if (!c->ev_recorded) { c->error = "rc_read_target before rc_render"; return RC_ERR_STATE; }
        // deferred fs_main (src/shader.wgsl:76-100) from the stored visibility of the last rendered frame
        launch_direct(c->scene, c->cam_rendered, c->lights_rendered, c->tile, c->d_depth.p, c->d_prim.p, c->d_bary.p,
                      c->d_albedo.p, c->d_direct.p, st);
        CU_OK(c, cudaGetLastError());
        c->direct_valid = true;
    }
    switch ((int)which) {
    case RC_TARGET_IRRADIANCE: src = c->irr(); break;
    case RC_TARGET_IRRADIANCE_RGB48:
        CU_OK(c, c->d_irr_pack.alloc(((size_t)c->tile.w * c->tile.h * 6 + 15) / 16 * 16 / 2));
        launch_pack_rgb48((size_t)c->tile.w * c->tile.h, c->irr(), c->d_irr_pack.p, st);
        src = c->d_irr_pack.p;
        break;
    case RC_TARGET_DIRECT: src = c->d_direct.p; break;
    case RC_TARGET_ALBEDO: src = c->d_albedo.p; break;
    case RC_TARGET_DEPTH: src = c->d_depth.p; break;
    case RC_TARGET_NORMAL: src = c->d_nrm.p; break;
    case RC_TARGET_PRIM: src = c->d_prim.p; break;
    case RC_TARGET_COMPOSITE: case RC_TARGET_DIRECT_SRGB8:
        if (!c->composite_valid) {
            launch_composite(c->tile, c->irr(), c->d_albedo.p, c->d_direct.p, c->d_composite.p, c->d_direct_srgb.p, st);
            c->composite_valid = true;
        }
        src = which == RC_TARGET_COMPOSITE ? (const void*)c->d_composite.p : (const void*)c->d_direct_srgb.p;
        break;
    default: src = c->d_cascade.p + c->levels[(int)which - RC_TARGET_CASCADE0].texel_offset; break;
    }
    CU_OK(c, cudaMemcpyAsync(host_dst, src, need, cudaMemcpyDeviceToHost, st));
    CU_OK(c, cudaStreamSynchronize(st));
    return RC_OK;
}

rc_status rc_read_target_async(rc_ctx* c, rc_target which, void* host_dst, size_t bytes, uint32_t* ticket)
{
    if (!c || !host_dst || !ticket) return RC_ERR_INVALID_ARG;
    if (which != RC_TARGET_IRRADIANCE && which != RC_TARGET_IRRADIANCE_RGB48) {
        c->error = "rc_read_target_async: only the irradiance targets are double-buffered";
        return RC_ERR_INVALID_ARG;
    }
    const bool rgb48 = which == RC_TARGET_IRRADIANCE_RGB48;
    const size_t need = (size_t)c->tile.w * c->tile.h * (rgb48 ? 6 : 8);
    if (bytes < need) { c->error = "rc_read_target_async: buffer too small"; return RC_ERR_BUFFER_SIZE; }
    cudaSetDevice(c->device);
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    const int slot = c->irr_slot;
    CU_OK(c, cudaEventRecord(c->ev_frame_done, st));
    CU_OK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_frame_done, 0));
    const void* src = c->irr();
    if (rgb48) {   // packed on the copy stream (which is ordered: pack i, copy i, pack i+1, ...), so one staging buffer serves
        CU_OK(c, c->d_irr_pack.alloc(((size_t)c->tile.w * c->tile.h * 6 + 15) / 16 * 16 / 2));
        launch_pack_rgb48((size_t)c->tile.w * c->tile.h, c->irr(), c->d_irr_pack.p, c->copy_stream);
        src = c->d_irr_pack.p;
    }
    bool by_sm = false;
    if (c->copy_blocks > 0 && need % 16 == 0) {   // SM-driven copy: only into page-locked memory the device can address
        cudaPointerAttributes pa{};
        void* dev_view = nullptr;
        if (cudaPointerGetAttributes(&pa, host_dst) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
            cudaHostGetDevicePointer(&dev_view, host_dst, 0) == cudaSuccess && dev_view) {
            launch_copy_to_host(src, dev_view, need, c->copy_blocks, c->copy_stream);
            by_sm = true;
        } else {
            cudaGetLastError();
        }
    }
    if (!by_sm) CU_OK(c, cudaMemcpyAsync(host_dst, src, need, cudaMemcpyDeviceToHost, c->copy_stream));
    CU_OK(c, cudaEventRecord(c->ev_copy_done[slot], c->copy_stream));
    c->copy_pending[slot] = true;
    *ticket = (uint32_t)slot;
    return RC_OK;
}

rc_status rc_read_wait(rc_ctx* c, uint32_t ticket)
{
    if (!c || ticket > 1) return RC_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    CU_OK(c, cudaEventSynchronize(c->ev_copy_done[ticket]));
    return RC_OK;
}

rc_status rc_stage_times(rc_ctx* c, float* ms, uint32_t n)
{
    if (!c || !ms) return RC_ERR_INVALID_ARG;
    if (!c->ev_recorded) { c->error = "rc_stage_times before rc_render"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    // (events recorded by graph nodes only change state when the graph runs: wait for the stream, not the event)
    CU_OK(c, cudaStreamSynchronize(c->last_stream ? c->last_stream : c->stream));
    CU_OK(c, cudaEventSynchronize(c->ev[EV_GATHER]));
    float t[RC_STAGE_COUNT] = {0};
    cudaEventElapsedTime(&t[RC_STAGE_GBUFFER], c->ev[EV_START], c->ev[EV_GBUF]);
    cudaEventElapsedTime(&t[RC_STAGE_PROBES], c->ev[EV_GBUF], c->ev[EV_PROBES]);
    if (c->frame_batched) {
        cudaEventElapsedTime(&t[RC_STAGE_MARCH], c->ev[EV_PROBES], c->ev[EV_MARCH_ALL]);   // (top fill +) k_march_all
        cudaEventElapsedTime(&t[RC_STAGE_MERGE], c->ev[EV_MARCH_ALL], c->ev[EV_LEVELS]);   // k_merge x (N-1)
    } else {
        cudaEventElapsedTime(&t[RC_STAGE_MARCH], c->ev[EV_PROBES], c->ev[EV_LEVELS]);      // march fused with merge
        t[RC_STAGE_MERGE] = 0.f;
    }
    cudaEventElapsedTime(&t[RC_STAGE_GATHER], c->ev[EV_LEVELS], c->ev[EV_GATHER]);
    cudaEventElapsedTime(&t[RC_STAGE_FRAME], c->ev[EV_START], c->ev[EV_GATHER]);
    for (uint32_t i = 0; i < n && i < RC_STAGE_COUNT; i++) ms[i] = t[i];
    return RC_OK;
}

rc_status rc_level_times(rc_ctx* c, float* ms, uint32_t n)
{
    if (!c || !ms) return RC_ERR_INVALID_ARG;
    if (!c->ev_recorded || !c->level_timing) { c->error = "rc_level_times needs rc_set_tuning(level_timing, 1) and a rendered frame"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    CU_OK(c, cudaEventSynchronize(c->ev[EV_GATHER]));
    for (uint32_t i = 0; i < n && i < c->N; i++) {
        // levels run top-down: level i starts when level i+1 (or the probe stage) ended
        cudaEvent_t begin = (i == c->N - 1) ? c->ev[EV_PROBES] : c->ev_level[i + 1];
        ms[i] = 0.f;
        cudaEventElapsedTime(&ms[i], begin, c->ev_level[i]);
    }
    return RC_OK;
}

rc_status rc_launch_count(rc_ctx* c, uint32_t* launches)
{
    if (!c || !launches) return RC_ERR_INVALID_ARG;
    *launches = c->launches;
    return RC_OK;
}

// ---- peer-memory frame exchange (tiled multi-GPU; kernels.cu k_gather / k_peer_*)
rc_status rc_peer_export(rc_ctx* c, void* handle, size_t handle_bytes)
{
    if (!c || !handle) return RC_ERR_INVALID_ARG;
    if (handle_bytes < sizeof(cudaIpcMemHandle_t)) { c->error = "rc_peer_export: handle buffer smaller than 64 bytes"; return RC_ERR_BUFFER_SIZE; }
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    peer_release(c);
    rc_ctx::Peer& p = c->peer;
    p.slot_bytes = (((size_t)c->W * c->H * sizeof(uint2)) + 255) & ~(size_t)255;
    const size_t bytes = 2 * p.slot_bytes + kPeerCtrlWords * sizeof(uint32_t);
    CU_OK(c, cudaMalloc(&p.local, bytes));
    CU_OK(c, cudaMemset(p.local, 0, bytes));
    cudaIpcMemHandle_t h;
    CU_OK(c, cudaIpcGetMemHandle(&h, p.local));
    memcpy(handle, &h, sizeof(h));
    return RC_OK;
}

rc_status rc_peer_attach(rc_ctx* c, const void* handles, uint32_t world, uint32_t rank)
{
    if (!c || !handles || world < 1 || world > (uint32_t)kMaxPeers || rank >= world) return RC_ERR_INVALID_ARG;
    if (!c->peer.local) { c->error = "rc_peer_attach before rc_peer_export"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    rc_ctx::Peer& p = c->peer;
    for (uint32_t r = 0; r < world; r++) {
        if (r == rank) { p.base[r] = p.local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + r * sizeof(h), sizeof(h));
        CU_OK(c, cudaIpcOpenMemHandle(&p.base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    p.rank = (int)rank;
    p.world = (int)world;     // from the next rc_render on, k_gather also stores into every rank's frame buffer
    p.seq = 0;
    return RC_OK;
}

rc_status rc_peer_wait(rc_ctx* c, void* stream)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if (!c->peer.world || !c->peer.seq) { c->error = "rc_peer_wait: no frame has been rendered with peers attached"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (!c->peer_broadcast && c->peer.rank != 0) return RC_OK;     // only rank 0 receives frames: nothing to wait for here
    launch_peer_wait(c->peer.world, c->peer.seq, c->peer.ctrl(c->peer.rank), st);
    CU_OK(c, cudaGetLastError());
    return RC_OK;
}

rc_status rc_peer_frame(rc_ctx* c, void** device_ptr, size_t* bytes, uint32_t* timeouts)
{
    if (!c || !device_ptr) return RC_ERR_INVALID_ARG;
    if (!c->peer.world) { c->error = "rc_peer_frame: peers not attached"; return RC_ERR_STATE; }
    *device_ptr = c->peer.frame(c->peer.rank, c->peer.seq);
    if (bytes) *bytes = (size_t)c->W * c->H * sizeof(uint2);
    if (timeouts) {
        cudaSetDevice(c->device);
        CU_OK(c, cudaMemcpy(timeouts, c->peer.ctrl(c->peer.rank) + kPeerError, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    return RC_OK;
}

rc_status rc_rays_marched(rc_ctx* c, uint32_t* rays, uint32_t n)
{
    if (!c || !rays) return RC_ERR_INVALID_ARG;
    if (!c->ev_recorded) { c->error = "rc_rays_marched before rc_render"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    CU_OK(c, cudaStreamSynchronize(c->last_stream ? c->last_stream : c->stream));
    for (uint32_t i = 0; i < n; i++) {
        uint32_t v = 0xffffffffu;
        if (i < c->N && c->frame_culled) {
            if (i + 1 == c->N && c->top_fillable()) v = 0;
            else v = c->list_len(i) * (i >= 1 ? 4u : 1u);   // lists above level 0 hold 2x2 quads
        }
        rays[i] = v;
    }
    return RC_OK;
}

rc_status rc_get_ray_list(rc_ctx* c, uint32_t level, uint32_t* entries, size_t bytes, uint32_t* count)
{
    if (!c || !count || level >= c->N) return RC_ERR_INVALID_ARG;
    if (!c->ev_recorded || !c->frame_culled) { c->error = "rc_get_ray_list needs a rendered frame with direction culling on"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    CU_OK(c, cudaStreamSynchronize(c->last_stream ? c->last_stream : c->stream));
    uint32_t n = (level + 1 == c->N && c->top_fillable()) ? 0u : c->h_ray_count[level];
    if (n == 0xffffffffu) n = 0u;
    *count = n;
    const size_t cap = c->list_offset[level + 1] - c->list_offset[level];
    size_t copy = n < bytes / 4 ? n : bytes / 4;
    if (copy > cap) copy = cap;
    if (entries && copy) CU_OK(c, cudaMemcpy(entries, c->d_list.p + c->list_offset[level], copy * 4, cudaMemcpyDeviceToHost));
    return RC_OK;
}

rc_status rc_get_split_list(rc_ctx* c, uint32_t level, uint32_t* entries, size_t bytes, uint32_t* n_enter, uint32_t* n_miss)
{
    if (!c || !n_enter || !n_miss || level >= c->N) return RC_ERR_INVALID_ARG;
    if (!c->ev_recorded || !c->frame_culled) { c->error = "rc_get_split_list needs a rendered frame with direction culling on"; return RC_ERR_STATE; }
    cudaSetDevice(c->device);
    CU_OK(c, cudaStreamSynchronize(c->last_stream ? c->last_stream : c->stream));
    const unsigned int len = c->h_ray_count[level], triv = c->h_ray_count[RC_MAX_LEVELS + level];
    if (!((c->split_mask >> level) & 1u) || len == 0xffffffffu || !(triv & 0x80000000u)) {
        c->error = "rc_get_split_list: the last frame did not classify this level (rc_set_tuning list_split)";
        return RC_ERR_STATE;
    }
    const uint32_t nm = triv & 0x7fffffffu, ne = len - nm;
    *n_enter = ne; *n_miss = nm;
    const size_t cap = c->list_offset[level + 1] - c->list_offset[level], room = bytes / 4;
    const size_t copy_a = std::min<size_t>(ne, room), copy_b = std::min<size_t>(nm, room - copy_a);
    const uint32_t* lb = c->d_list2.p + c->list_offset[level];
    if (entries && copy_a) CU_OK(c, cudaMemcpy(entries, lb, copy_a * 4, cudaMemcpyDeviceToHost));
    if (entries && copy_b) CU_OK(c, cudaMemcpy(entries + copy_a, lb + (cap - copy_b), copy_b * 4, cudaMemcpyDeviceToHost));   // (appended from the back)
    return RC_OK;
}

rc_status rc_get_levels(rc_ctx* c, rc_level_info* out, uint32_t max_levels, uint32_t* num_levels)
{
    if (!c) return RC_ERR_INVALID_ARG;
    if (num_levels) *num_levels = c->N;
    if (out) for (uint32_t i = 0; i < c->N && i < max_levels; i++) out[i] = c->level_info[i];
    return RC_OK;
}

rc_status rc_get_scene_info(rc_ctx* c, rc_scene_info* out)
{
    if (!c || !out) return RC_ERR_INVALID_ARG;
    *out = c->host.info;
    return RC_OK;
}

rc_status rc_get_tile(rc_ctx* c, uint32_t xywh[4])
{
    if (!c || !xywh) return RC_ERR_INVALID_ARG;
    xywh[0] = (uint32_t)c->tile.x0; xywh[1] = (uint32_t)c->tile.y0; xywh[2] = (uint32_t)c->tile.w; xywh[3] = (uint32_t)c->tile.h;
    return RC_OK;
}

rc_status rc_get_intervals(rc_ctx* c, float out3[3])
{
    if (!c || !out3) return RC_ERR_INVALID_ARG;
    out3[0] = c->L0; out3[1] = c->t_far; out3[2] = c->offset;
    return RC_OK;
}

rc_status rc_get_directions(rc_ctx* c, uint32_t level, float* out, size_t bytes)
{
    if (!c || !out || level >= c->N) return RC_ERR_INVALID_ARG;
    const auto& d = c->dirs_host[level];
    if (bytes < d.size() * sizeof(float)) { c->error = "rc_get_directions: buffer too small"; return RC_ERR_BUFFER_SIZE; }
    memcpy(out, d.data(), d.size() * sizeof(float));
    return RC_OK;
}

rc_status rc_get_model_stream(rc_ctx* c, uint32_t model, float* vertices, size_t vbytes, uint32_t* indices, size_t ibytes,
                              uint32_t* num_vertices, uint32_t* num_indices)
{
    if (!c) return RC_ERR_INVALID_ARG;
    return scene_model_stream(c->host, c->error, model, vertices, vbytes, indices, ibytes, num_vertices, num_indices);
}

rc_status rc_get_model_material(rc_ctx* c, uint32_t model, void* out80, size_t bytes)
{
    if (!c || !out80 || model >= c->host.models.size()) return RC_ERR_INVALID_ARG;
    if (bytes < 80) return RC_ERR_BUFFER_SIZE;
    memcpy(out80, c->host.models[model].material80, 80);
    return RC_OK;
}

/* ---- device-free scene ingest (≙ ObjScene::load on the CPU) ---- */
rc_status rc_scene_load(const char* obj_path, uint32_t flags, rc_scene** out)
{
    if (!obj_path || !out) return RC_ERR_INVALID_ARG;
    *out = nullptr;
    rc_scene* s = new rc_scene();
    rc_status st;
    try {
        st = load_host_scene(obj_path, flags, s);
    } catch (const std::exception& e) {
        s->error = std::string("rc_scene_load: ") + e.what();
        st = RC_ERR_SCENE_LOAD;
    }
    if (st != RC_OK) { g_create_error = s->error; delete s; return st; }
    *out = s;
    return RC_OK;
}

void rc_scene_free(rc_scene* s) { delete s; }

rc_status rc_scene_get_info(const rc_scene* s, rc_scene_info* out)
{
    if (!s || !out) return RC_ERR_INVALID_ARG;
    *out = s->info;
    return RC_OK;
}

rc_status rc_scene_model_stream(rc_scene* s, uint32_t model, float* vertices, size_t vbytes, uint32_t* indices, size_t ibytes,
                                uint32_t* num_vertices, uint32_t* num_indices)
{
    if (!s) return RC_ERR_INVALID_ARG;
    return scene_model_stream(*s, s->error, model, vertices, vbytes, indices, ibytes, num_vertices, num_indices);
}

rc_status rc_scene_model_material(const rc_scene* s, uint32_t model, void* out80, size_t bytes)
{
    if (!s || !out80 || model >= s->models.size()) return RC_ERR_INVALID_ARG;
    if (bytes < 80) return RC_ERR_BUFFER_SIZE;
    memcpy(out80, s->models[model].material80, 80);
    return RC_OK;
}

rc_status rc_decode_image_file(const char* path, uint8_t* rgba, size_t bytes, uint32_t* width, uint32_t* height)
{
    if (!path) return RC_ERR_INVALID_ARG;
    std::string why;
    std::shared_ptr<Image> img;
    try {
        img = load_image_rgba8(path, &why);
    } catch (const std::exception& e) {
        why = e.what();
    }
    if (!img) { g_create_error = why; return RC_ERR_SCENE_LOAD; }
    if (width) *width = img->width;
    if (height) *height = img->height;
    if (!rgba) return RC_OK;
    if (bytes < img->rgba.size()) return RC_ERR_BUFFER_SIZE;
    memcpy(rgba, img->rgba.data(), img->rgba.size());
    return RC_OK;
}

rc_status rc_scene_model_name(const rc_scene* s, uint32_t model, char* out, size_t bytes)
{
    if (!s || !out || !bytes || model >= s->models.size()) return RC_ERR_INVALID_ARG;
    snprintf(out, bytes, "%s", s->models[model].name.c_str());
    return RC_OK;
}

rc_status rc_scene_model_texture(const rc_scene* s, uint32_t model, uint32_t which, uint8_t* rgba, size_t bytes,
                                 uint32_t* width, uint32_t* height)
{
    if (!s || model >= s->models.size() || which > 1) return RC_ERR_INVALID_ARG;
    const int slot = which ? s->models[model].tex_n : s->models[model].tex_c;
    if (width) *width = slot < 0 ? 0u : s->tex[slot].w;
    if (height) *height = slot < 0 ? 0u : s->tex[slot].h;
    if (slot < 0 || !rgba) return RC_OK;
    const size_t need = (size_t)s->tex[slot].w * s->tex[slot].h * 4;
    if (bytes < need) return RC_ERR_BUFFER_SIZE;
    memcpy(rgba, s->tex_data.data() + s->tex[slot].offset, need);
    return RC_OK;
}

rc_status rc_trace_rays(rc_ctx* c, const float* rays, uint32_t n, float* hits)
{
    if (!c || !rays || !hits) return RC_ERR_INVALID_ARG;
    if (!n) return RC_OK;
    cudaSetDevice(c->device);
    CU_OK(c, c->d_dbg_in.alloc((size_t)n * 8));
    CU_OK(c, c->d_dbg_out.alloc((size_t)n * 4));
    CU_OK(c, cudaMemcpyAsync(c->d_dbg_in.p, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    launch_trace_rays(c->scene, c->d_dbg_in.p, n, c->d_dbg_out.p, c->stream);
    CU_OK(c, cudaGetLastError());
    CU_OK(c, cudaMemcpyAsync(hits, c->d_dbg_out.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

rc_status rc_shade_points(rc_ctx* c, const float* in, uint32_t n, float* out)
{
    if (!c || !in || !out) return RC_ERR_INVALID_ARG;
    if (!n) return RC_OK;
    cudaSetDevice(c->device);
    CU_OK(c, c->d_dbg_in.alloc((size_t)n * 8));
    CU_OK(c, c->d_dbg_out.alloc((size_t)n * 4));
    CU_OK(c, cudaMemcpyAsync(c->d_dbg_in.p, in, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    launch_shade_points(c->scene, c->lights, c->d_dbg_in.p, n, c->d_dbg_out.p, c->stream);
    CU_OK(c, cudaGetLastError());
    CU_OK(c, cudaMemcpyAsync(out, c->d_dbg_out.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

rc_status rc_cascade_device_ptr(rc_ctx* c, uint32_t level, void** dev_ptr, size_t* bytes)
{
    if (!c || level >= c->N || !dev_ptr) return RC_ERR_INVALID_ARG;
    *dev_ptr = c->d_cascade.p + c->levels[level].texel_offset;
    if (bytes) *bytes = (size_t)c->level_info[level].texel_count * 8;
    return RC_OK;
}

rc_status rc_irradiance_device_ptr(rc_ctx* c, void** dev_ptr, size_t* bytes)
{
    if (!c || !dev_ptr) return RC_ERR_INVALID_ARG;
    *dev_ptr = c->irr();
    if (bytes) *bytes = (size_t)c->tile.w * c->tile.h * 8;
    return RC_OK;
}

}  // extern "C"
