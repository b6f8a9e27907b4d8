// kernels.cuh — launch wrappers of the sm_100a kernels (kernels.cu).
#pragma once
#include "rc_device.cuh"

namespace rc {

struct GBufferOut {
    float* depth; uint32_t* prim; uint32_t* normal; float2* bary;
    // occupancy of the tile in 32x8-pixel cells: occ[cell] == frame  <=>  some pixel of the cell saw geometry in this frame
    // (a stamp instead of a flag: nothing has to be cleared between frames).  k_probes skips the S6 candidate search of cells
    // that are empty as a whole — most of the sky.
    uint32_t* occ; uint32_t frame; int ow;
};

struct TileRect { int x0, y0, w, h; };

// k_entry work decomposition: level l is covered by groups of g[l] x g[l] probes, one thread each
// k_march_all work decomposition: launch slot k marches level level[k] with blocks [block_offset[k], block_offset[k+1])
struct MarchPlan {
    unsigned block_offset[RC_MAX_LEVELS + 1];
    unsigned dir_offset[RC_MAX_LEVELS];
    int level[RC_MAX_LEVELS], map[RC_MAX_LEVELS], top[RC_MAX_LEVELS], use_entry[RC_MAX_LEVELS];
    int n;
};
// Peer-memory exchange of the finished tiles (tiled multi-GPU): every rank's k_gather stores its tile into all
// ranks' full-frame buffers over NVLink.  world == 0: disabled.
constexpr int kMaxPeers = 8;
enum { kPeerArrived = 0, kPeerReleased = kMaxPeers, kPeerError = 2 * kMaxPeers, kPeerCtrlWords = 2 * kMaxPeers + 8 };
struct PeerOut {
    uint2* frame[kMaxPeers];       // rank d's full-frame buffer (this frame's slot), W x H RGBA16F
    uint32_t* ctrl[kMaxPeers];     // rank d's control block
    int world, rank, W;
    uint32_t seq;                  // frame number, from 1
    uint32_t dst_mask;             // ranks that receive the tiles (bit d): rank 0 alone = final image gather, all = all-gather
};
struct EntryPlan { unsigned group_offset[RC_MAX_LEVELS + 1]; int g[RC_MAX_LEVELS]; int n; };

// pixmask != null (direction culling, DD0 = D0^2 <= 16): also stores per pixel the mask of level-0 directions
// the gather will weight with cs_d > 0
void launch_gbuffer(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, GBufferOut out, int DD0,
                    const float* dirs0, uint16_t* pixmask, cudaStream_t st);
// Primary visibility by triangle binning (k_bin + k_gbuffer_binned): same outputs as launch_gbuffer, bit for bit.
// bin_count: bin_tiles(tile) counters, all zero when the frame starts (the raster kernel zeroes what it consumes);
// bin_lists: bin_list_entries(tile) triangle slots; n_leaf_tris: entries of DScene::tri_geom / 3
size_t bin_tiles(TileRect tile);
size_t bin_list_entries(TileRect tile);
size_t bin_huge_bytes(uint32_t n_leaf_tris);     // the frame's list of triangles too large to bin (tested by every tile)
void launch_gbuffer_binned(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, GBufferOut out, int DD0,
                           const float* dirs0, uint16_t* pixmask, uint32_t n_leaf_tris, unsigned int* bin_count, uint32_t* bin_lists,
                           unsigned int* huge_count, void* huge, cudaStream_t st);
// direction culling, level lv (bottom-up): appends the requests of `need` (bits at resolution Dr, probes of lv) to
// `list` / `count` and pushes them to the upper level's masks `need_up` (has_upper: 0 none, 1 same resolution
// [level 0 -> 1], 2 expanded 2x; up_words = mask words per upper probe) through k_link's tables; clear: zero the
// consumed words (levels >= 1, whose masks are accumulated with atomicOr and must be empty for the next frame)
void launch_need(const DLevel& lv, int Dr, int has_upper, int up_words, const float4* origin, const uint4* link_idx,
                 const float4* link_w, uint32_t* need, uint32_t* need_up, uint32_t* list, unsigned int* count, bool clear, bool pdl,
                 bool trigger, bool dir_major, int tile_order, bool append, int4 own, cudaStream_t st);
// append = false: only propagate the masks (halo exchange, first pass); append = true with own.z >= 0: list only the requests of
// the probes inside own = (x0, y0, x1, y1), inclusive sub-grid coordinates (the probes this rank owns)
// pdl: programmatic dependent launch on the previous level's k_need (the origins / link tables must be older);
// trigger: the next launch in the stream is a pdl k_need, so this one may release it early;
// dir_major: order each warp's list entries by request (direction) first, probe second;
// tile_order: a probe's quads are appended 4x2-tile by 4x2-tile instead of row by row (1: request resolution 32 only —
// measured: level 4 of the 4K frame 0.175 -> 0.161 ms, resolutions 8 / 16 unchanged or slower; 2: resolutions 8, 16, 32)
// deferred fs_main: albedo / direct colour from the stored visibility (on demand)
void launch_direct(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, const float* depth, const uint32_t* prim,
                   const float2* bary, uint2* albedo, uint2* direct, cudaStream_t st);
// all levels' probes in one launch; anchors inside the tile reuse the G-buffer hit; with pixmask (direction
// culling) every level-0 probe also gets need0[probe] = OR of the masks of the pixels it serves
void launch_probes(const DScene& s, const DCamera& cam, const DLevelSet& ls, unsigned total, TileRect tile, float offset,
                   const float* depth, const uint32_t* prim, float4* origin, float4* normal, const uint16_t* pixmask,
                   uint32_t* need0, const uint32_t* occ, uint32_t frame, int ow, bool floating, cudaStream_t st);
// One launch, two independent jobs that only read the probe origins:
//  link:  per lower probe (levels 0..N-2, `link_total` probes) the 4 upper probe slots (sub-grid linear) and
//         normalised weights (w.x < 0: no valid upper probe)
//  entry: per probe of levels 0..entry_levels-1 the entry frontier of the BVH for the ball B(origin, t1) —
//         2 x int4 links per probe, valid first, padded with 0x80000000 (rc_device.cuh trace(..., entry))
void launch_link_entry(const DScene& s, const DLevelSet& ls, unsigned link_total, int entry_levels, const float4* origin,
                       const float4* normal, uint4* link_idx, float4* link_w, int4* entry, cudaStream_t st);
// march level lv; fused != 0 also merges with the (already merged) upper level, read through its child averages
// `up_avg` (float4 per upper probe and lower direction: 0.25*(((c0+c1)+c2)+c3), S8); entry: this level's
// frontiers or null; avg_out: where to leave this level's own child averages when the kernel finalises the
// level — only honoured when march_avg_ystep(D, map) != 0, otherwise call launch_child_avg afterwards.
// list / count (direction culling): march only the requested texels (quad = 0, level 0) or 2x2 quads (quad = 1)
// listed by launch_need; avg_out is then always honoured.  up_const: the upper level is an unmaterialised top
// level that cannot hit anything — up_avg then holds the top probes' origins (far_field)
// dirs: this level's direction table (the compacting variant reads it); dirq: the same directions with their slab
// reciprocals, 2 x float4 per direction: (w, 1/w.x), (1/w.y, 1/w.z, 0, 0) — what k_march reads
void launch_march(const DScene& s, const DLights& L, const DLevel& lv, const DLevel* up, bool top, float3 sky,
                  const float4* origin, const float* dirs, const float4* dirq, uint2* texels, const float4* up_avg,
                  const uint4* link_idx, const float4* link_w, const int4* entry, float4* avg_out, bool fused, int map, int occ,
                  bool pdl, bool compact, int max_blocks, const uint32_t* list, const unsigned int* count, int quad, bool up_const,
                  const unsigned int* count_triv, unsigned list_cap, cudaStream_t st);
// culled levels >= 1, block-local ray pool with refill of finished lanes (k_march_pool); max_blocks: pools of 1024 rays
void launch_march_pool(const DScene& s, const DLights& L, const DLevel& lv, float3 sky, const float4* origin, const float4* dirq, uint2* texels,
                       const float4* up_avg, const uint4* link_idx, const float4* link_w, const int4* entry, float4* avg_out, bool fused, int occ,
                       bool pdl, int max_blocks, const uint32_t* list, const unsigned int* count, const unsigned int* count_triv, unsigned list_cap,
                       int thresh, bool up_const, cudaStream_t st);
// culled levels >= 1, one 2x2 quad of texels per thread (k_march_quad): list = the level's quad list, count = its length;
// occ = resident 64-thread blocks per SM the register allocation must allow (8 -> 128 regs, 12 -> 80, 16 -> 64)
void launch_march_quad(const DScene& s, const DLights& L, const DLevel& lv, float3 sky, const float4* origin, const float4* dirq, uint2* texels,
                       const float4* up_avg, const uint4* link_idx, const float4* link_w, float4* avg_out, bool fused, int occ, bool pdl, int max_blocks,
                       const uint32_t* list, const unsigned int* count, bool up_const, cudaStream_t st);
// levels first..last of the request chain in one launch (a thread-block cluster with a cluster barrier between the levels);
// per level the same parameters as launch_need, offsets in words / entries from need_all / list_all; probe arrays are whole-frame
// split ray lists (k_split): a level's list (list_a, length counts[level]) is copied into list_b — entries whose rays all miss the
// BVH root's two child boxes from the back of the level's `cap` entries (counted in counts[RC_MAX_LEVELS + level], bit 31 set as
// "classified"), the others from the front (counts[2 * RC_MAX_LEVELS + level]).
struct SplitJob {
    DLevel lv;
    int level;
    unsigned cap;
    const float4* root;       // node 0 of the BVH
    const float4* origin;     // the level's probe origins
    const float4* dirq;       // the level's direction + slab-reciprocal table (2 x float4 per direction): level 0
    const float4* qinv;       // levels >= 1: 3 x float4 per quad = the reciprocals of its children (0,0) (1,0) (0,1) (1,1)
    const uint32_t* list_a;
    uint32_t* list_b;
    unsigned int* counts;
};
// blocks [block_off[k], block_off[k+1]) of one launch work on job[k]
struct SplitPlan {
    int n;
    unsigned block_off[RC_MAX_LEVELS + 1];
    SplitJob job[RC_MAX_LEVELS];
};
int split_chunk();
void launch_split(const SplitPlan& plan, cudaStream_t st);
struct NeedChain {
    DLevel lv[RC_MAX_LEVELS];
    int Dr[RC_MAX_LEVELS], has_upper[RC_MAX_LEVELS], up_words[RC_MAX_LEVELS], clear[RC_MAX_LEVELS], dir_major[RC_MAX_LEVELS], tile_order[RC_MAX_LEVELS];
    unsigned long long need_off[RC_MAX_LEVELS], need_up_off[RC_MAX_LEVELS], list_off[RC_MAX_LEVELS];
    int4 own[RC_MAX_LEVELS];
    int first, last, append;
};
void launch_need_chain(const NeedChain& c, const float4* origin, const uint4* link_idx, const float4* link_w, uint32_t* need_all,
                       uint32_t* list_all, unsigned int* counts, cudaStream_t st);
int march_avg_ystep(int D, int map);
// child averages of a finalised level from its texels (paths whose march kernel does not write them itself)
void launch_child_avg(const DLevel& lv, const uint2* texels, float4* avg_out, cudaStream_t st);
// every listed level's rays in one launch, unmerged (the caller runs launch_merge top-down afterwards);
// levels[k] in launch order, map[level], dir_offset[level] in floats into dirs, origin / cascade / entry: whole-frame bases
void launch_march_all(const DScene& s, const DLights& L, const DLevelSet& ls, const int* levels, int n, const int* map,
                      const size_t* dir_offset, int entry_levels, int top_level, float3 sky, const float4* origin,
                      const float* dirs, uint2* cascade, const int4* entry, int occ, cudaStream_t st);
// persistent variant: resident grid, dynamic ray fetch with lane replacement, PDL-chained across levels
void launch_march_persist(const DScene& s, const DLights& L, const DLevel& lv, const DLevel* up, bool top, float3 sky,
                          const float4* origin, const float* dirs, uint2* texels, const float4* up_avg,
                          const uint4* link_idx, const float4* link_w, bool fused, int map, int thresh, int grid_blocks,
                          unsigned int* counter, bool pdl, cudaStream_t st);
int march_persist_blocks_per_sm();
// top level whose interval lies entirely outside the scene bounds: fill with (sky, 1) instead of marching
void launch_fill_top(const DLevel& lv, float3 sky, const float4* origin, uint2* texels, float4* avg_out, cudaStream_t st);
void launch_merge(const DLevel& lv, const DLevel& up, float3 sky, const float4* origin, uint2* texels, const float4* up_avg,
                  const uint4* link_idx, const float4* link_w, cudaStream_t st);
// Tensor-core gather (k_gather_mma; D0 = 4, P0 = 4 only): dirs0_host = the level-0 direction table on the host (it travels in
// the kernel parameter block), axis_nx / axis_ny = device tables of S4's nx(x), ny(y) for the full frame, symmetric =
// gather_dirs_symmetric(dirs0_host).  enabled = 0 (or other cascade parameters) selects the scalar kernels.
struct GatherMma { int enabled; int symmetric; const float* dirs0_host; const float* axis_nx; const float* axis_ny; };
bool gather_dirs_symmetric(const float* dirs0_host);
// tiles_per_block > 1 (and D0 = 4): software-pipelined variant, a block walks a column of that many 32x8 pixel tiles
void launch_gather(const DCamera& cam, const DLevel& l0, TileRect tile, const float4* origin0, const uint2* texels0,
                   const float* dirs0, const float* depth, const uint32_t* normal, uint2* out, unsigned int* counts_in,
                   unsigned int* counts_out, const PeerOut& peer, int tiles_per_block, const GatherMma& mma, cudaStream_t st);
void launch_peer_begin(const PeerOut& peer, uint32_t* my_ctrl, cudaStream_t st);
void launch_peer_publish(const PeerOut& peer, cudaStream_t st);
void launch_peer_wait(int world, uint32_t seq, uint32_t* my_ctrl, cudaStream_t st);
// SM-driven device -> pinned-host copy (bytes % 16 == 0, dst = device-visible address of page-locked host memory)
void launch_copy_to_host(const void* src, void* dst_host_mapped, size_t bytes, int blocks, cudaStream_t st);
// RGBA16F irradiance -> RGB48 (3 x float16 per pixel, sign bit of r = "no geometry"): the 6-byte read-back format
void launch_pack_rgb48(size_t n_pixels, const uint2* irradiance, void* out, cudaStream_t st);
void launch_composite(TileRect tile, const uint2* irradiance, const uint2* albedo, const uint2* direct,
                      uchar4* composite, uchar4* direct_srgb, cudaStream_t st);
void launch_trace_rays(const DScene& s, const float* rays, uint32_t n, float* hits, cudaStream_t st);
void launch_shade_points(const DScene& s, const DLights& L, const float* in, uint32_t n, float* out, cudaStream_t st);

}  // namespace rc
