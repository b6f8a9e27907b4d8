/*
 * rc_b200.h — C ABI of librc_b200.so: the B200-native radiance-cascade GI path.
 *
 * This is the drop-in boundary.  Each entry point names the
 * reference interface it stands in for (paths relative to the reference
 * repository jw910731/RadianceCascade).  The reference has NO cascade / ray
 * march / merge / gather code; the GI stages behind this ABI
 * implement the builder-owned specification in include/rc_spec.h, while the
 * scene ingest, camera, material and direct-lighting inputs follow the
 * reference's source.
 *
 * Conventions: plain C, no torch / CUDA types in signatures (the stream is an
 * opaque void* holding a cudaStream_t, NULL = the context's own stream).
 * All entry points return rc_status and never abort across the ABI; a
 * human-readable message for the last failure of a context is available from
 * rc_last_error().  One context = one GPU = one owning host thread; per frame
 * the call order is rc_update -> rc_render -> rc_read_target, mirroring
 * App::handle_redraw (src/window/app.rs:221-267).
 */
#ifndef RC_B200_H
#define RC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_ABI_VERSION 1u

typedef int32_t rc_status;
enum {
    RC_OK = 0,
    RC_ERR_INVALID_ARG = 1,
    RC_ERR_SCENE_LOAD = 2,   /* ≙ the .unwrap() panic at src/renderer.rs:176 */
    RC_ERR_CUDA = 3,
    RC_ERR_NO_DEVICE = 4,    /* no CUDA device: there is NO CPU fallback */
    RC_ERR_BUFFER_SIZE = 5,
    RC_ERR_STATE = 6
};

/* rc_config.flags */
#define RC_CFG_SEPARATE_MERGE 0x1u /* run march and merge as separate kernels (debug / A-B) instead of the fused path */
#define RC_CFG_NO_TEXTURES    0x2u /* ignore map_Kd / map_Bump (as if the files were missing, src/primitives.rs:390-404) */
/* Tiled context with EXCHANGED instead of recomputed halos: cascade levels >= 1 are marched only for the probes
 * whose anchor pixel lies in this context's tile ("owned"); the caller moves two things between the ranks with any transport
 * (NCCL send/recv in radiancecascade_b200/distributed.py): the request masks of the probes a rank needs but does not own, to
 * their owners, between rc_render_begin and rc_render_lists; and each finished level's child averages of owned probes to the
 * ranks whose sub-grid holds them, after every rc_render_level.  Level 0 and an unmaterialised top level are never exchanged.
 * rc_render on such a context returns RC_ERR_STATE. */
#define RC_CFG_HALO_EXCHANGE  0x4u
/* The reference's clip volume and depth test for primary visibility (rc_spec.h S4b): pixels and probe anchors see the closest
 * fragment between the near and far planes of the projection inside view_proj (src/camera.rs:77-79; src/app.rs:26 near 0.1,
 * far 100), exactly what the render pass keeps (Depth32Float, Less, clear 1.0: src/renderer.rs:354-360, 585-592).
 * Without the flag primary rays see [0, inf): the GI benchmark cameras use far = 4 x the scene diagonal either way. */
#define RC_CFG_RASTER_CLIP    0x8u
/* rc_spec.h S6, optional: a probe whose anchor pixel sees no geometry floats to the first anchor of the finer levels' probes inside
 * its cell that does, so that no valid probe is left without a valid upper probe (the far field is then never dropped next to a
 * silhouette).  Costs 4-8 % of a frame. */
#define RC_CFG_FLOATING_PROBES 0x10u

/* rc_update flags.  bit0 ≙ AppState::enable_normal_map (src/app.rs:18, src/renderer.rs:620-631) */
#define RC_UPD_ENABLE_NORMAL_MAP 0x1u

typedef struct rc_ctx rc_ctx;

/* ≙ arguments of DefaultRenderer::new(device, config, queue, state, path)
 * (src/renderer.rs:168-174): surface size, scene path; plus the cascade
 * parameters of include/rc_spec.h and the CUDA device ordinal.  Zero means
 * "default" for every cascade field. */
typedef struct rc_config {
    uint32_t struct_size;    /* = sizeof(rc_config) */
    uint32_t width, height;  /* full frame in pixels (≙ SurfaceConfiguration) */
    int32_t  device;         /* CUDA ordinal */
    const char* scene_path;  /* .obj file (≙ `path`, joined onto resource_root when relative) */
    const char* resource_root; /* ≙ RESOURCE_PATH (src/primitives.rs:12); may be NULL */
    uint32_t probe_spacing0; /* P0, pixels, default 4 */
    uint32_t dir_res0;       /* D0, directions per axis at level 0, default 4 */
    uint32_t num_levels;     /* N, default 6 */
    float    interval0;      /* L0 world units; <= 0: bbox_diag/256 */
    float    t_far;          /* end of the top interval; <= 0: 4*bbox_diag */
    float    normal_offset;  /* probe lift along the geometric normal; <= 0: 1e-3*L0... see rc_spec.h */
    float    sky[3];         /* radiance of a top-level miss, default 0 */
    uint32_t flags;          /* RC_CFG_* */
    /* Screen-space tile rendered by this context (multi-GPU).
     * tile_w == 0 or tile_h == 0 means the full frame. */
    uint32_t tile_x0, tile_y0, tile_w, tile_h;
} rc_config;

/* Same 80-byte layout as UniformCamera (src/camera.rs:9-15): column-major
 * proj*view matrix, then (eye, 1). */
typedef struct rc_camera {
    float view_proj[16];
    float eye[4];
} rc_camera;

/* Same 16-byte layout as UniformLight (src/primitives.rs:14-18). */
typedef struct rc_light {
    float position[4];
} rc_light;

typedef enum rc_target {
    RC_TARGET_IRRADIANCE = 0, /* float16 RGBA  [tile_h][tile_w][4]  linear HDR irradiance E, a = 1 where geometry */
    RC_TARGET_DIRECT     = 1, /* float16 RGBA  [tile_h][tile_w][4]  fs_main output (linear), a = 1 where geometry */
    RC_TARGET_DEPTH      = 2, /* float32       [tile_h][tile_w]     primary-ray distance, < 0 where no geometry */
    RC_TARGET_NORMAL     = 3, /* uint32        [tile_h][tile_w]     shading normal, 2 x snorm16 octahedral (standard mapping) */
    RC_TARGET_ALBEDO     = 4, /* float16 RGBA  [tile_h][tile_w][4]  linear albedo (`color` of fs_main) */
    RC_TARGET_PRIM       = 5, /* uint32        [tile_h][tile_w]     global triangle id, 0xffffffff where no geometry */
    RC_TARGET_COMPOSITE  = 6, /* uint8 BGRA    [tile_h][tile_w][4]  sRGB-encoded albedo*E/pi + direct (Bgra8UnormSrgb, src/window/app.rs:59-75) */
    RC_TARGET_DIRECT_SRGB8 = 7, /* uint8 BGRA  [tile_h][tile_w][4]  what the reference presents: sRGB_encode(fs_main) */
    RC_TARGET_IRRADIANCE_RGB48 = 8, /* float16 RGB [tile_h][tile_w][3]  the irradiance without its alpha channel, bit-exact: the sign bit of r is
                                     * set where the pixel has no geometry (E >= 0 everywhere else).  6 instead of 8 bytes per pixel over
                                     * PCIe; also accepted by rc_read_target_async */
    RC_TARGET_CASCADE0   = 16 /* + level i: float16 RGBA, merged cascade level i, layout in rc_spec.h */
} rc_target;

/* Integer layout of one cascade level, as used by the kernels (must be
 * bit-exact with the oracle's tables). */
typedef struct rc_level_info {
    uint32_t spacing;      /* P_i */
    uint32_t dir_res;      /* D_i */
    uint32_t grid_w, grid_h;     /* probes of the full frame at this level */
    int32_t  px0, py0;     /* first probe column / row held by this context */
    uint32_t sub_w, sub_h; /* probes held by this context (== grid for a full frame) */
    uint64_t texel_offset; /* first texel of the level in the cascade buffer */
    uint64_t texel_count;  /* sub_w*sub_h*dir_res^2 */
    float    t_begin, t_end;
} rc_level_info;

typedef struct rc_scene_info {
    uint32_t num_models, num_vertices, num_triangles, num_materials, num_textures;
    uint32_t bvh_nodes, light_from_obj; /* light_from_obj ≙ AppState::given_light_position */
    float    bbox_min[3], bbox_max[3];
    float    obj_light[3];
} rc_scene_info;

/* Stage indices for rc_stage_times. */
enum { RC_STAGE_GBUFFER = 0, RC_STAGE_PROBES = 1, RC_STAGE_MARCH = 2, RC_STAGE_MERGE = 3,
       RC_STAGE_GATHER = 4, RC_STAGE_FRAME = 5, RC_STAGE_COUNT = 6 };

/* ≙ DefaultRenderer::new (src/renderer.rs:168-555): loads the scene
 * (ObjScene::load, src/primitives.rs:122-175), builds vertex streams, materials,
 * textures, the BVH, and sizes every device buffer.  Fails with
 * RC_ERR_NO_DEVICE when no CUDA device is present. */
rc_status rc_create(const rc_config* cfg, rc_ctx** out);

/* ≙ the per-frame host->device writes: queue.write_buffer(camera_buffer,
 * UniformCamera) and queue.write_buffer(light_buffer, UniformLight)
 * (src/window/app.rs:112-130) followed by RenderStage::update
 * (src/renderer.rs:620-631).  The reference supports one light; n_lights > 1
 * sums the diffuse and specular terms over lights. */
rc_status rc_update(rc_ctx* ctx, const rc_camera* cam, const rc_light* lights,
                    uint32_t n_lights, uint32_t flags);

/* ≙ RenderStage::resize (src/renderer.rs:615-618) + Projection::resize. */
rc_status rc_resize(rc_ctx* ctx, uint32_t width, uint32_t height);

/* Multi-GPU (no reference counterpart): move this context's screen-space tile inside the unchanged frame, e.g. to re-balance
 * the tiles of a tiled frame from measured frame times.  Cheap: device buffers are grow-only and reused, camera / lights /
 * peer mappings / tuning stay; waits for frames in flight.  w == h == 0 selects the full frame. */
rc_status rc_set_tile(rc_ctx* ctx, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h);

/* ≙ RenderStage::render (src/renderer.rs:559-613): enqueues the frame on
 * `stream` (a cudaStream_t, NULL = context stream); does not synchronise and
 * does not allocate.  Stages: G-buffer, probe placement, per-level march
 * (+merge), gather. */
rc_status rc_render(rc_ctx* ctx, void* stream);

/* Waits for the last rc_render and copies a target into host memory
 * (≙ reading back the colour attachment the caller owns). */
rc_status rc_read_target(rc_ctx* ctx, rc_target which, void* host_dst, size_t bytes);
/* Pipelined read-back of RC_TARGET_IRRADIANCE (double-buffered on the device): enqueues the copy of the
 * frame just rendered on a separate copy stream and returns a ticket; the next rc_update / rc_render may be
 * issued immediately and overlaps the copy.  rc_read_wait(ticket) blocks until host_dst is complete.
 * host_dst should be page-locked, and at most two copies may be outstanding (one per buffer).  To keep the GPU's
 * queue full, submit frame i (rc_update, rc_render) BEFORE waiting for the read-back of frame i-2 and only then call
 * rc_read_target_async for frame i (bench.py e2e loop); the device side is ordered by the library. */
rc_status rc_read_target_async(rc_ctx* ctx, rc_target which, void* host_dst, size_t bytes, uint32_t* ticket);
rc_status rc_read_wait(rc_ctx* ctx, uint32_t ticket);
/* Size in bytes rc_read_target needs for `which`. */
rc_status rc_target_bytes(rc_ctx* ctx, rc_target which, size_t* bytes);

/* CUDA-event time of each stage of the last rendered frame, in ms. */
rc_status rc_stage_times(rc_ctx* ctx, float* ms, uint32_t n);
/* CUDA-event time of each cascade level's march(+merge) kernels in the last frame, ms[level].
 * Needs rc_set_tuning(ctx, "level_timing", 1): the events sit between the level kernels and
 * switch off their programmatic-dependent-launch overlap, so they are off by default. */
rc_status rc_level_times(rc_ctx* ctx, float* ms, uint32_t n);
/* Runtime tuning knobs (A/B measurement; defaults are the measured best): "level_timing" 0/1,
 * "march_persist" 0/1 (persistent ray-replacement march vs one ray per thread), "march_thresh"
 * 1..32 (refill when fewer lanes are busy), "march_pdl" 0/1, "march_block" 64..512, "march_grid",
 * "march_map<level>" 0 linear / 1 direction tile / 2 probe tile, "march_entry" -1 auto / n levels that start their
 * traversal at per-probe BVH entry frontiers, "march_batch" 0/1 (all levels in one launch + separate merges),
 * "cull" 0/1 (direction culling), "graph" 0/1 (submit the frame as one CUDA graph), "gather_tiles" 0 auto / 1 the
 * one-tile-per-block gather / n tiles per block with prefetch, "list_dir_major" bit i: level i's ray list is ordered
 * direction-major, "need_pdl" 0/1 (programmatic dependent launch along the k_need chain), "copy_blocks" 0 = the copy
 * engine / n = rc_read_target_async stores the frame into page-locked memory from n resident blocks,
 * "list_split" 0 off / 1 adaptive (default) / 2 every level: k_split sorts the list entries whose rays all miss the BVH
 * root's two child boxes to the back of the level's list and k_march skips their traversal (bit-identical texels),
 * "march_pool" 0/1 + "march_pool_thresh" 1..32 (block-local ray pool with refill of finished lanes; measured slower),
 * "gather_mma", "list_tiled", "need_fused", "march_quad", "gbuffer_binned", "peer_stores", "peer_broadcast": DESIGN.md §4-5. */
rc_status rc_set_tuning(rc_ctx* ctx, const char* key, int value);
/* Number of kernels rc_render launches per frame. */
rc_status rc_launch_count(rc_ctx* ctx, uint32_t* launches);
/* Direction culling (default on; rc_set_tuning "cull" 0 marches every texel): rays[level] = texels of the level
 * that the last frame actually marched — the texels some pixel's irradiance depends on with a non-zero weight;
 * every other texel of the cascade is left unspecified.  UINT32_MAX where the level was not culled (every texel
 * marched), 0 for a top level that is constant and never materialised.  Waits for the frame to finish. */
rc_status rc_rays_marched(rc_ctx* ctx, uint32_t* rays, uint32_t n);
/* Debug / parity entry: the ray list of `level` as the last culled frame built it (an index table: must match the
 * culling rule of include/rc_spec.h bit for bit, in any order).  Entry = probe * R^2 + r with probe the sub-grid
 * linear probe index; level 0: R = D_0, r = texel (dy * D_0 + dx); level i >= 1: R = D_{i-1}, r = the 2x2 quad of
 * level-i texels below direction r of level i-1.  *count = entries of the list; at most bytes / 4 are copied. */
rc_status rc_get_ray_list(rc_ctx* ctx, uint32_t level, uint32_t* entries, size_t bytes, uint32_t* count);
/* Debug / parity entry for the split ray lists (rc_set_tuning "list_split"): the list of `level` as k_split partitioned it
 * in the last frame — first the *n_enter entries whose rays are traversed, then the *n_miss entries classified as certain
 * misses (every ray of the entry fails the slab test of the BVH root's two child boxes over the level's interval, so
 * none of them can hit a triangle).  Same entry format as rc_get_ray_list; together the two parts are that list.
 * RC_ERR_STATE when the last frame did not classify the level.  At most bytes / 4 entries are copied. */
rc_status rc_get_split_list(rc_ctx* ctx, uint32_t level, uint32_t* entries, size_t bytes, uint32_t* n_enter, uint32_t* n_miss);

/* ---- Tiled multi-GPU: final-image exchange through NVLink peer memory (one context per GPU, one process each,
 * all on one node).  Replaces the all-gather of the finished tiles: k_gather stores every pixel of this rank's tile
 * straight into EVERY rank's full-frame buffer as it is produced, then publishes a per-rank "delivered frame n"
 * flag; rc_peer_wait enqueues the acquire of all ranks' flags.  Two buffer slots alternate, and a rank announces
 * that it is done with frame n when it starts frame n+1 in stream order — so consume the assembled frame on the
 * stream you render on, before the next rc_render.
 * By default the tiles are gathered on rank 0 only (final image gather); rc_set_tuning("peer_broadcast", 1) on every rank
 * makes every rank receive the whole frame (all-gather, N times the traffic).
 *   rc_peer_export  allocates this rank's buffers and returns their 64-byte CUDA IPC handle
 *   rc_peer_attach  handles = world x 64 bytes, rank-ordered (exchange them with any host-side all-gather)
 *   rc_peer_wait    enqueue: wait until every rank's tile of the last rendered frame has arrived here
 *   rc_peer_frame   device pointer of the assembled W x H RGBA16F frame; timeouts (optional) = waits that gave up
 *                   after ~2 s instead of hanging the GPU (0 in a healthy run) */
rc_status rc_peer_export(rc_ctx* ctx, void* handle, size_t handle_bytes);
rc_status rc_peer_attach(rc_ctx* ctx, const void* handles, uint32_t world, uint32_t rank);
rc_status rc_peer_wait(rc_ctx* ctx, void* stream);
rc_status rc_peer_frame(rc_ctx* ctx, void** device_ptr, size_t* bytes, uint32_t* timeouts);

rc_status rc_get_levels(rc_ctx* ctx, rc_level_info* out, uint32_t max_levels, uint32_t* num_levels);
rc_status rc_get_scene_info(rc_ctx* ctx, rc_scene_info* out);
/* The tile this context renders: x0, y0, w, h in full-frame pixels. */
rc_status rc_get_tile(rc_ctx* ctx, uint32_t xywh[4]);
/* The cascade intervals actually in use: L0, t_far, probe normal offset. */
rc_status rc_get_intervals(rc_ctx* ctx, float out3[3]);
/* Direction table of a level: float[dir_res*dir_res][3], row-major (dy, dx). */
rc_status rc_get_directions(rc_ctx* ctx, uint32_t level, float* out, size_t bytes);

/* ≙ the 17-float interleaved vertex buffer (src/renderer.rs:371-410) and the
 * winding-reversed index buffer (src/primitives.rs:369-376) of model `m`.
 * Pass NULL buffers to query counts. */
rc_status rc_get_model_stream(rc_ctx* ctx, uint32_t model, float* vertices, size_t vbytes,
                              uint32_t* indices, size_t ibytes,
                              uint32_t* num_vertices, uint32_t* num_indices);
/* ≙ UniformMaterial (64 bytes, src/primitives.rs:37-73) followed by enable_bit
 * (u32) and emission Ke (3 floats, GI only): 20 x 4 bytes. */
rc_status rc_get_model_material(rc_ctx* ctx, uint32_t model, void* out80, size_t bytes);

/* Debug / parity entry: closest-hit query of arbitrary rays against the
 * context's BVH.  rays: float[n][8] = origin xyz, tmin, dir xyz, tmax.
 * hits: float[n][4] = t (<0 miss), u, v, prim id as float bits. */
rc_status rc_trace_rays(rc_ctx* ctx, const float* rays, uint32_t n, float* hits);
/* Debug / parity entry: fs_main (src/shader.wgsl:76-100) evaluated on the GPU at
 * hit points.  in: float[n][8] = prim id bits, u, v, pad, view-origin xyz, pad. out: float[n][4]. */
rc_status rc_shade_points(rc_ctx* ctx, const float* in, uint32_t n, float* out);

/* Halo exchange: what the caller needs to move level `level`'s request masks and child averages.  Both buffers are row-major over
 * the context's probe sub-grid [sub_h][sub_w]: need = need_words_per_probe uint32 per probe (requests this rank's pixels make of
 * the probe), avg = avg_float4_per_probe float4 per probe (the averages the level below merges from, written for the owned
 * probes by rc_render_level).  own_* = the owned probes (inclusive sub-grid coordinates; own_x1 < own_x0: none).
 * exchanged = 0 for level 0 and for a top level that is never materialised. */
typedef struct rc_exchange_info {
    void*    need_ptr;
    void*    avg_ptr;
    uint32_t need_words_per_probe, avg_float4_per_probe;
    int32_t  px0, py0;
    uint32_t sub_w, sub_h;
    int32_t  own_x0, own_y0, own_x1, own_y1;
    uint32_t exchanged, pad;
} rc_exchange_info;
rc_status rc_exchange_level_info(rc_ctx* ctx, uint32_t level, rc_exchange_info* out);
/* Halo exchange, between rc_render_begin and the first rc_render_level: builds the ray lists from the request masks (which the
 * caller has completed with the other ranks' requests) for the probes this context owns. */
rc_status rc_render_lists(rc_ctx* ctx, void* stream);

/* Device pointer + size of the merged level (debug / custom transports). */
rc_status rc_cascade_device_ptr(rc_ctx* ctx, uint32_t level, void** dev_ptr, size_t* bytes);
rc_status rc_irradiance_device_ptr(rc_ctx* ctx, void** dev_ptr, size_t* bytes);
/* Split rc_render for halo exchange: levels are processed top-down; after
 * rc_render_level(i) the caller exchanges level i's border ring, then calls
 * rc_render_level(i-1); rc_render_begin does G-buffer + probes, rc_render_end the gather. */
rc_status rc_render_begin(rc_ctx* ctx, void* stream);
rc_status rc_render_level(rc_ctx* ctx, uint32_t level, void* stream);
rc_status rc_render_end(rc_ctx* ctx, void* stream);

/* Device-free scene ingest: the CPU part of ObjScene::load (src/primitives.rs:122-175)
 * and of DefaultRenderer::new's mesh preparation (src/renderer.rs:370-497) — tobj-style
 * loading, winding reversal, TBN, 17-float vertex stream, UniformMaterial, enable_bit,
 * texture decode.  Needs no GPU; rc_create runs exactly this and then uploads. */
typedef struct rc_scene rc_scene;
rc_status rc_scene_load(const char* obj_path, uint32_t flags /* RC_CFG_NO_TEXTURES */, rc_scene** out);
void rc_scene_free(rc_scene* scene);
rc_status rc_scene_get_info(const rc_scene* scene, rc_scene_info* out); /* bvh_nodes = 0: no BVH is built here */
rc_status rc_scene_model_stream(rc_scene* scene, uint32_t model, float* vertices, size_t vbytes,
                                uint32_t* indices, size_t ibytes, uint32_t* num_vertices, uint32_t* num_indices);
rc_status rc_scene_model_material(const rc_scene* scene, uint32_t model, void* out80, size_t bytes);
rc_status rc_scene_model_name(const rc_scene* scene, uint32_t model, char* out, size_t bytes);
/* which: 0 colour map (map_Kd), 1 normal map (map_Bump); width = height = 0 when the model has none */
rc_status rc_scene_model_texture(const rc_scene* scene, uint32_t model, uint32_t which, uint8_t* rgba, size_t bytes,
                                 uint32_t* width, uint32_t* height);

/* ≙ image::ImageReader::open(path).decode().to_rgba8() as used by Scene::material and
 * Texture::from_image (src/primitives.rs:391-404, src/texture.rs:85-93): built-in PNG and JPEG
 * decoders.  Pass rgba = NULL to query the size. */
rc_status rc_decode_image_file(const char* path, uint8_t* rgba, size_t bytes, uint32_t* width, uint32_t* height);

rc_status rc_synchronize(rc_ctx* ctx);
void rc_destroy(rc_ctx* ctx);
const char* rc_last_error(const rc_ctx* ctx); /* ctx may be NULL: last rc_create failure */
uint32_t rc_abi_version(void);

/* Host-side façade keeping the reference's camera math callable from C
 * (glam conventions). */
/* ≙ Camera::calc_matrix (src/camera.rs:43-52). out16 column-major. */
void rc_camera_view_matrix(const float position[3], float yaw, float pitch, float out16[16]);
/* ≙ Projection::calc_matrix (src/camera.rs:77-79); fovy in radians. */
void rc_projection_matrix(float fovy, float aspect, float znear, float zfar, float out16[16]);
/* ≙ UniformCamera::from_camera_project (src/camera.rs:17-22). */
void rc_uniform_camera(const float position[3], float yaw, float pitch,
                       float fovy, float aspect, float znear, float zfar, rc_camera* out);
/* look-at helper for synthetic orbit paths: same look_to_rh arithmetic with dir = normalize(target - position). */
void rc_uniform_camera_look_at(const float position[3], const float target[3],
                               float fovy, float aspect, float znear, float zfar, rc_camera* out);

#ifdef __cplusplus
}
#endif
#endif /* RC_B200_H */
