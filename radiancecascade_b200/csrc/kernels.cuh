// kernels.cuh — launch wrappers of the sm_100a kernels (kernels.cu).
#pragma once
#include "rc_device.cuh"

namespace rc {

struct GBufferOut {
    float* depth; uint32_t* prim; uint32_t* normal; float2* bary;
};

struct TileRect { int x0, y0, w, h; };

void launch_gbuffer(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, GBufferOut out, cudaStream_t st);
// deferred fs_main: albedo / direct colour from the stored visibility (on demand)
void launch_direct(const DScene& s, const DCamera& cam, const DLights& L, TileRect tile, const float* depth, const uint32_t* prim,
                   const float2* bary, uint2* albedo, uint2* direct, cudaStream_t st);
// all levels' probes in one launch; anchors inside the tile reuse the G-buffer hit
void launch_probes(const DScene& s, const DCamera& cam, const DLevelSet& ls, unsigned total, TileRect tile, float offset,
                   const float* depth, const uint32_t* prim, float4* origin, float4* normal, cudaStream_t st);
// per lower probe (levels 0..N-2, `total` probes): the 4 upper probe slots (sub-grid linear) and
// normalised weights (w.x < 0: no valid upper probe)
void launch_link(const DLevelSet& ls, unsigned total, const float4* origin, const float4* normal, uint4* link_idx,
                 float4* link_w, cudaStream_t st);
// march level lv; fused != 0 also merges with the (already merged) upper level
void launch_march(const DScene& s, const DLights& L, const DLevel& lv, const DLevel* up, bool top, float3 sky,
                  const float4* origin, const float* dirs, uint2* texels, const uint2* up_texels,
                  const uint4* link_idx, const float4* link_w, bool fused, int map, int occ, bool pdl, bool compact,
                  int max_blocks, cudaStream_t st);
// persistent variant: resident grid, dynamic ray fetch with lane replacement, PDL-chained across levels
void launch_march_persist(const DScene& s, const DLights& L, const DLevel& lv, const DLevel* up, bool top, float3 sky,
                          const float4* origin, const float* dirs, uint2* texels, const uint2* up_texels,
                          const uint4* link_idx, const float4* link_w, bool fused, int map, int thresh, int grid_blocks,
                          unsigned int* counter, bool pdl, cudaStream_t st);
int march_persist_blocks_per_sm();
// top level whose interval lies entirely outside the scene bounds: fill with (sky, 1) instead of marching
void launch_fill_top(const DLevel& lv, float3 sky, const float4* origin, uint2* texels, cudaStream_t st);
void launch_merge(const DLevel& lv, const DLevel& up, float3 sky, const float4* origin, uint2* texels, const uint2* up_texels,
                  const uint4* link_idx, const float4* link_w, cudaStream_t st);
void launch_gather(const DCamera& cam, const DLevel& l0, TileRect tile, const float4* origin0, const uint2* texels0,
                   const float* dirs0, const float* depth, const uint32_t* normal, uint2* out, cudaStream_t st);
void launch_composite(TileRect tile, const uint2* irradiance, const uint2* albedo, const uint2* direct,
                      uchar4* composite, uchar4* direct_srgb, cudaStream_t st);
void launch_trace_rays(const DScene& s, const float* rays, uint32_t n, float* hits, cudaStream_t st);
void launch_shade_points(const DScene& s, const DLights& L, const float* in, uint32_t n, float* out, cudaStream_t st);

}  // namespace rc
