/* tauray_b200.h -- C ABI of the B200-native path-tracing hot path.
 *
 * This library replaces three pieces of vga-group/tauray behind its renderer/stage surface:
 *   rt_pipeline            (src/rt_pipeline.cc:13-192)           -> CUDA kernel launch table
 *   path_tracer_stage      (src/path_tracer_stage.cc:30-147)     -> tr_pt_create / tr_pt_render
 *   acceleration_structure (src/acceleration_structure.cc:8-443) -> on-device LBVH built by tr_pt_set_scene
 * plus the multi-GPU gather of stitch_stage/device_transfer (src/stitch_stage.cc:128-196,
 * src/device_transfer.cc:140-347) -> strip buffers exchanged over NVLink + tr_pt_unshuffle.
 *
 * Conventions: every entry point returns 0 or a negative tr_status, never
 * throws; the caller owns all input memory (it is copied during the call); the library owns
 * device memory; one host thread per context; all GPU work is ordered on the context's stream
 * (the stream plays the role of the reference's `dependencies`). There is NO CPU fallback:
 * tr_pt_create fails with TR_ERR_NO_DEVICE when no sm_100 device is usable.
 *
 * All table structs below are the reference's own GPU wire formats (GLSL `scalar` layout),
 * so Tauray's scene_stage can hand over its staging data unchanged. Matrices are
 * column-major (glm), as in the reference.
 */
#ifndef TAURAY_B200_H
#define TAURAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TR_B200_ABI_VERSION 2

typedef enum tr_status {
    TR_OK = 0,
    TR_ERR_INVALID_ARGUMENT = -1,
    TR_ERR_NO_DEVICE = -2,       /* no CUDA device / not sm_100: there is no fallback */
    TR_ERR_CUDA = -3,            /* a CUDA runtime call failed, see tr_pt_last_error */
    TR_ERR_UNSUPPORTED = -4,     /* option combination outside the hot-path scope */
    TR_ERR_NO_SCENE = -5,
    TR_ERR_OUT_OF_MEMORY = -6
} tr_status;

/* ---- enums: values equal the reference's ------------------------------------------- */
/* rt_stage::sampler_type, src/rt_stage.hh:21-27 */
enum { TR_SAMPLER_UNIFORM_RANDOM = 0, TR_SAMPLER_SOBOL_OWEN = 1,
       TR_SAMPLER_SOBOL_Z_ORDER_2D = 2, TR_SAMPLER_SOBOL_Z_ORDER_3D = 3 };
/* film_filter, src/rt_common.hh:8-13 */
enum { TR_FILM_POINT = 0, TR_FILM_BOX = 1, TR_FILM_BLACKMAN_HARRIS = 2 };
/* multiple_importance_sampling_mode, src/rt_common.hh:16-21 */
enum { TR_MIS_DISABLED = 0, TR_MIS_BALANCE_HEURISTIC = 1, TR_MIS_POWER_HEURISTIC = 2 };
/* bounce_sampling_mode, src/rt_common.hh:24-29 */
enum { TR_BOUNCE_HEMISPHERE = 0, TR_BOUNCE_COSINE_HEMISPHERE = 1, TR_BOUNCE_MATERIAL = 2 };
/* tri_light_sampling_mode, src/rt_common.hh:32-37 */
enum { TR_TRI_LIGHT_AREA = 0, TR_TRI_LIGHT_SOLID_ANGLE = 1, TR_TRI_LIGHT_HYBRID = 2 };
/* distribution_strategy, src/distribution_strategy.hh:8-18 */
enum { TR_DISTRIBUTION_DUPLICATE = 0, TR_DISTRIBUTION_SCANLINE = 1,
       TR_DISTRIBUTION_SHUFFLED_STRIPS = 2 };
/* camera::projection_type (perspective / orthographic share one layout, src/camera.cc:397-407) */
enum { TR_PROJECTION_PERSPECTIVE = 0, TR_PROJECTION_ORTHOGRAPHIC = 1 };
/* feature_stage::feature, src/feature_stage.cc:35-66 (the AOV tracer) */
enum { TR_FEATURE_ALBEDO = 0, TR_FEATURE_WORLD_NORMAL = 1, TR_FEATURE_VIEW_NORMAL = 2,
       TR_FEATURE_WORLD_POS = 3, TR_FEATURE_VIEW_POS = 4, TR_FEATURE_DISTANCE = 5,
       TR_FEATURE_INSTANCE_ID = 9 };
/* kernel organisation of the hot loop (ours) */
enum { TR_KERNEL_WAVEFRONT = 0, TR_KERNEL_MEGAKERNEL = 1 };
/* Which rgen program a render runs. TR_ESTIMATOR_DIRECT == direct_stage (src/direct_stage.cc:29-126 +
 * shader/direct.rgen:50-132): one camera ray per pass, samples_per_pass next-event-estimation samples at
 * the first hit, no MIS (direct_stage sets no MIS define, so nee_mis_pdf returns the light pdf), sphere
 * lights hidden from the camera ray (mask 0xFF^0x02); max_ray_depth only bounds the Sobol dimension
 * groups; mis_mode, hide_lights, russian roulette, clamping and regularisation do not apply. */
enum { TR_ESTIMATOR_PATH_TRACER = 0, TR_ESTIMATOR_DIRECT = 1 };
/* texture formats accepted in tr_texture */
enum { TR_TEXTURE_RGBA8_UNORM = 0, TR_TEXTURE_RGBA32F = 1, TR_TEXTURE_RGBA16F = 2 };
/* Layers of gbuffer_target the path tracer writes (TR_GBUFFER_ENTRIES, src/gbuffer.hh:18-51;
 * write_all_outputs, shader/path_tracer.glsl:535-576). Texel formats are the reference's. */
enum {
    TR_GBUFFER_COLOR = 0,               /* rgba32f (src/rt_renderer.cc:189-190); always present */
    TR_GBUFFER_DIFFUSE = 1,             /* rgba16f; demodulated diffuse light, a = 1/second-hit distance; running mean */
    TR_GBUFFER_REFLECTION = 2,          /* rgba16f; demodulated reflected light; running mean */
    TR_GBUFFER_ALBEDO = 3,              /* rgba16f; written by the first sample only, like all layers below */
    TR_GBUFFER_MATERIAL = 4,            /* rgba8 unorm: metallic, roughness, ior/4, transmittance (gbuffer.glsl:256-260) */
    TR_GBUFFER_NORMAL = 5,              /* rg16 snorm: octahedral mapped normal (math.glsl:480-485) */
    TR_GBUFFER_POS = 6,                 /* rgba32f world position */
    TR_GBUFFER_SCREEN_MOTION = 7,       /* rgba32f: previous-frame uv + depth (camera.glsl:61-67) */
    TR_GBUFFER_INSTANCE_ID = 8,         /* r32 sint */
    TR_GBUFFER_LAYER_COUNT = 9
};
/* environment_map::projection, src/environment_map.hh:12-15 */
enum { TR_ENVMAP_NONE = -1, TR_ENVMAP_LAT_LONG = 0 };

/* ---- scene tables ------------------------------------------------------------------ */
/* mesh::vertex, src/mesh.hh:19-25 == shader/scene.glsl:14-20 */
typedef struct tr_vertex {
    float pos[3];
    float normal[3];
    float uv[2];
    float tangent[4];
} tr_vertex; /* 48 B */

/* material_buffer, src/scene_stage.cc:17-30 == shader/material.glsl:9-22 */
typedef struct tr_material {
    float albedo_factor[4];
    float metallic_roughness_factor[4]; /* x = metallic, y = roughness */
    float emission_factor[4];
    float transmittance;
    float ior;
    float normal_factor;
    uint32_t flags;                     /* bit0 double sided, bit1 transient */
    int32_t albedo_tex_id;              /* -1 = none */
    int32_t metallic_roughness_tex_id;
    int32_t normal_tex_id;
    int32_t emission_tex_id;
} tr_material; /* 80 B */

/* instance_buffer, src/scene_stage.cc:32-44 == shader/scene.glsl:43-53 */
typedef struct tr_instance {
    int32_t light_base_id;              /* -1: not emissive; else first tri-light index */
    int32_t sh_grid_index;
    uint32_t pad;
    float shadow_terminator_mul;
    float model[16];
    float model_normal[16];             /* inverse transpose of model */
    float model_prev[16];
    tr_material mat;
} tr_instance; /* 288 B */

/* directional_light_entry, src/scene_stage.cc:46-62 == shader/light.glsl:7-13 */
typedef struct tr_directional_light {
    float color[3];
    int32_t shadow_map_index;
    float dir[3];
    float dir_cutoff;                   /* cos(radians(angle)) */
} tr_directional_light; /* 32 B */

/* point_light_entry (also spotlights), src/scene_stage.cc:64-98 == shader/light.glsl:15-27 */
typedef struct tr_point_light {
    float color[3];
    float dir[3];
    float pos[3];
    float radius;
    float dir_cutoff;
    float dir_falloff;                  /* 0: not a spotlight */
    float cutoff_radius;
    float spot_radius;
    int32_t shadow_map_index;
    int32_t padding;
} tr_point_light; /* 64 B */

/* tri_light_entry, src/scene_stage.cc:103-111 == shader/light.glsl:29-37; built on the device
 * (extract_tri_lights.comp:17-54), exposed only through tr_pt_read_tri_lights for tests. */
typedef struct tr_tri_light {
    float pos[3][3];
    uint32_t emission_factor;           /* R9G9B9E5 */
    uint32_t instance_id;
    uint32_t primitive_id;
    uint32_t uv[3];                     /* half2 */
    int32_t emission_tex_id;
} tr_tri_light; /* 64 B */

/* matrix_camera_data_buffer, src/camera.cc:397-407 == shader/camera.glsl:13-23 */
typedef struct tr_camera_data {
    float view[16];
    float view_inverse[16];
    float view_proj[16];
    float proj_inverse[16];
    float origin[4];
    float dof_params[4];                /* focus distance, aperture radius, angle, sides */
    float projection_info[4];
    float pan[4];
} tr_camera_data; /* 320 B */

/* camera_pair, shader/scene.glsl:176-180 */
typedef struct tr_camera_pair {
    tr_camera_data current;
    tr_camera_data previous;
} tr_camera_pair; /* 640 B */

/* alias_table_entry, src/environment_map.hh:37-43 == shader/alias_table.glsl:7-13 */
typedef struct tr_alias_table_entry {
    uint32_t alias_id;
    uint32_t probability;
    float pdf;
    float alias_pdf;
} tr_alias_table_entry; /* 16 B */

/* One entry per instance: the mesh it references (mesh::get_vertices/get_indices,
 * src/mesh.hh:60-63) and the opaque flag the reference gives the BLAS geometry
 * (= !material::potentially_transparent(), src/material.cc:7-11; src/scene_stage.cc:1410). */
typedef struct tr_mesh_ref {
    const tr_vertex* vertices;
    const uint32_t* indices;
    uint32_t vertex_count;
    uint32_t index_count;
    uint32_t opaque;
    uint32_t reserved;
} tr_mesh_ref;

/* A texture of the sampler table (src/sampler_table.cc); bilinear, repeat, LOD 0. */
typedef struct tr_texture {
    const void* data;
    uint32_t width;
    uint32_t height;
    uint32_t format;                    /* TR_TEXTURE_* */
    uint32_t reserved;
} tr_texture;

/* What scene_stage flattens for the tracer (src/scene_stage.cc:1066-1116,1277-1354). */
typedef struct tr_pt_scene_desc {
    uint32_t instance_count;
    uint32_t point_light_count;
    uint32_t directional_light_count;
    uint32_t texture_count;
    const tr_instance* instances;
    const tr_mesh_ref* meshes;          /* instance_count entries */
    const tr_point_light* point_lights;
    const tr_directional_light* directional_lights;
    const tr_texture* textures;
    float environment_factor[4];        /* scene_metadata.environment_factor */
    int32_t environment_proj;           /* TR_ENVMAP_NONE (-1) or TR_ENVMAP_LAT_LONG */
    uint32_t reserved;
    /* Texels of the environment map when environment_proj >= 0 (RGBA16F like the reference's .hdr
     * textures, src/texture.cc:498-510, or RGBA32F); sampled bilinearly with repeat in x and
     * clamp-to-edge in y (envmap_sampler, src/scene_stage.cc:259-265). The importance-sampling
     * alias table (environment_map::generate_alias_table, src/environment_map.cc:39-140 +
     * shader/alias_table_importance.comp) is built by tr_pt_set_scene. NULL otherwise. */
    const tr_texture* environment_map;
} tr_pt_scene_desc;

/* ---- options: field names and meaning follow path_tracer_stage::options
 * (src/path_tracer_stage.hh:13-30) <- rt_camera_stage::options (src/rt_camera_stage.hh:17-25)
 * <- rt_stage::options (src/rt_stage.hh:29-40). tr_pt_default_options() fills the CLI
 * defaults of src/options.hh (which create_renderer always copies over the struct defaults,
 * src/tauray.cc:392-403,472-488). ----------------------------------------------------- */
typedef struct tr_pt_options {
    /* rt_stage::options */
    int32_t max_ray_depth;              /* 8 */
    float min_ray_dist;                 /* 1e-4 */
    int32_t rng_seed;                   /* 0 */
    int32_t local_sampler;              /* TR_SAMPLER_UNIFORM_RANDOM */
    /* rt_camera_stage::options */
    int32_t samples_per_pixel;          /* 1 */
    int32_t samples_per_pass;           /* 1 */
    int32_t projection;                 /* TR_PROJECTION_PERSPECTIVE */
    int32_t transparent_background;     /* 0 */
    uint32_t active_viewport_count;     /* 1 */
    /* path_tracer_stage::options */
    int32_t use_shadow_terminator_fix;  /* 0 */
    int32_t use_white_albedo_on_first_bounce; /* 0 */
    int32_t hide_lights;                /* 0 */
    int32_t film;                       /* TR_FILM_POINT */
    int32_t mis_mode;                   /* TR_MIS_POWER_HEURISTIC */
    float film_radius;                  /* 0.5 */
    float russian_roulette_delta;       /* 0 = off */
    float indirect_clamping;            /* 0 = off */
    float regularization_gamma;         /* 0 = off */
    int32_t depth_of_field;             /* 0 */
    float sampling_weight_point_lights; /* 1 */
    float sampling_weight_directional_lights; /* 1 */
    float sampling_weight_envmap;       /* 1 */
    float sampling_weight_emissive_triangles; /* 1 */
    int32_t bounce_mode;                /* TR_BOUNCE_MATERIAL */
    int32_t tri_light_mode;             /* TR_TRI_LIGHT_SOLID_ANGLE (CLI default) */
    /* ours */
    int32_t kernel_mode;                /* TR_KERNEL_WAVEFRONT */
    int32_t estimator;                  /* TR_ESTIMATOR_PATH_TRACER; TR_ESTIMATOR_DIRECT == direct_stage */
    int32_t reserved[4];
} tr_pt_options;

/* distribution_params, src/distribution_strategy.hh:20-27 */
typedef struct tr_pt_distribution {
    uint32_t size[2];
    int32_t strategy;
    uint32_t index;
    uint32_t count;
    int32_t primary;
} tr_pt_distribution;

/* Counters of the last tr_pt_render (the reference's timer "path tracing", src/rt_stage.cc:60).
 * Instrumentation flags (tr_pt_set_instrumentation): TR_INSTRUMENT_VISITS counts BVH node visits
 * and triangle tests (slower kernels); TR_INSTRUMENT_TIMING brackets every kernel launch with CUDA
 * events on the render stream and sums the durations per kernel class. */
#define TR_INSTRUMENT_VISITS 1
#define TR_INSTRUMENT_TIMING 2
typedef struct tr_pt_stats {
    uint64_t closest_rays;              /* traversal queries issued (closest hit) */
    uint64_t shadow_rays;               /* traversal queries issued (shadow / any hit) */
    uint64_t paths;                     /* camera paths started */
    uint64_t closest_node_visits;       /* TR_INSTRUMENT_VISITS */
    uint64_t closest_triangle_tests;
    uint64_t shadow_node_visits;
    uint64_t shadow_triangle_tests;
    uint64_t kernel_launches;           /* kernels launched by the last tr_pt_render */
    uint64_t culled_shadow_rays;        /* NEE shadow rays NOT traced because the sample's BSDF weight is exactly 0
                                           (the result cannot depend on visibility); shadow_rays + culled = the
                                           reference's shadow-ray count */
    float render_ms;                    /* CUDA-event time of the last tr_pt_render */
    float closest_ms;                   /* TR_INSTRUMENT_TIMING: summed over launches of the class */
    float shadow_ms;
    float shade_ms;
    float other_ms;
    uint32_t closest_launches;
    uint32_t shadow_launches;
    uint32_t shade_launches;
    float bvh_build_ms;                 /* last tr_pt_set_scene */
    uint32_t bvh_node_count;
    uint32_t bvh_depth;
    uint32_t triangle_count;
    uint32_t tri_light_count;
    uint32_t reserved;
} tr_pt_stats;

typedef struct tr_pt_ctx tr_pt_ctx;     /* one per GPU == one path_tracer_stage */

/* ---- entry points ---------------------------------------------------------------------- */

/* CLI defaults (src/options.hh:43-49,170-214,292-294,358-366,438-488,513-532). */
void tr_pt_default_options(tr_pt_options* opt);

/* == path_tracer_stage ctor (src/path_tracer_stage.cc:30-116): selects the kernel permutation
 * (the reference's shader #defines) from `opt`. */
int tr_pt_create(tr_pt_ctx** out, int cuda_device, const tr_pt_options* opt);
void tr_pt_destroy(tr_pt_ctx* ctx);

/* == scene_stage upload + ensure_blas + TLAS build (src/scene_stage.cc:821-872,1419-1468,
 * 1620-1723): copies the tables, pre-transforms vertices (pre_transform.comp), extracts
 * tri-lights (extract_tri_lights.comp) and builds the LBVH on the device. */
int tr_pt_set_scene(tr_pt_ctx* ctx, const tr_pt_scene_desc* scene);

/* == the per-frame part of scene_stage::update for animated scenes: instance-buffer refresh
 * (src/scene_stage.cc:1066-1116), pre_transform + tri-light re-extraction and the acceleration
 * structure update (scene_stage::record_as_build / record_tri_light_extraction,
 * src/scene_stage.cc:1572-1675; top_level_acceleration_structure::rebuild,
 * src/acceleration_structure.cc:376-443). `instances` replaces the scene's instance records
 * (transforms, materials; same count, same meshes, same light_base_id layout). Everything runs on
 * the device from the vertex data tr_pt_set_scene uploaded: world-space vertices, tri lights and a
 * full LBVH rebuild (a rebuild, not a refit: it costs a few ms and keeps the tree quality). The
 * result is identical to tr_pt_set_scene with the new records. Accumulation is not reset, as in the
 * reference (the application calls reset_accumulation). */
int tr_pt_update_instances(tr_pt_ctx* ctx, const tr_instance* instances, uint32_t instance_count);

/* How tr_pt_update_instances brings the acceleration structure up to date:
 * TR_UPDATE_REBUILD (default) -- a new radix tree every frame (above);
 * TR_UPDATE_REFIT -- the tree tr_pt_set_scene built (the quality build) keeps its topology, slot assignment and
 *   triangle order; the leaf triangles are re-fetched and the quantised child boxes recomputed bottom-up. This is what
 *   the reference asks of the driver for animated geometry (bottom-level `eUpdate` builds of fast-build+update BLASes,
 *   src/acceleration_structure.cc:187-262; top level :376-443): cheaper than a build and as good as the original tree
 *   while the motion is moderate. The image never depends on the tree, so both modes render identical pictures. */
#define TR_UPDATE_REBUILD 0
#define TR_UPDATE_REFIT 1
int tr_pt_set_update_mode(tr_pt_ctx* ctx, int mode);

/* == camera UBO upload (src/scene_stage.cc:1145-1174); `pairs` = view_count tr_camera_pair. */
int tr_pt_set_cameras(tr_pt_ctx* ctx, const tr_camera_pair* pairs, uint32_t view_count);
/* The same in stream order, for frames in flight: the table is copied on `cuda_stream` -- the stream the following
 * tr_pt_render calls use -- so it lands after the frames already enqueued there and before the next one, and the host
 * does not wait for the frame in flight. `pairs` must be pinned host memory and stay untouched until the copy has run.
 * A different `view_count` than the current one re-allocates the table and blocks like tr_pt_set_cameras. */
int tr_pt_set_cameras_async(tr_pt_ctx* ctx, const tr_camera_pair* pairs, uint32_t view_count, void* cuda_stream);

/* == rt_camera_stage::reset_distribution_params (src/rt_camera_stage.cc:141-145) */
int tr_pt_set_distribution(tr_pt_ctx* ctx, const tr_pt_distribution* dist);

/* == rt_camera_stage::reset_accumulated_samples + rt_stage::reset_sample_counter
 * (src/rt_renderer.cc:67-81) */
int tr_pt_reset_accumulation(tr_pt_ctx* ctx, int reset_sample_counter);

/* == stage::run (src/stage.cc:35-76) for path_tracer_stage: all passes of one frame
 * (samples_per_pixel / samples_per_pass launches of the hot loop), asynchronous on `cuda_stream`
 * (NULL = the context's own stream). */
int tr_pt_render(tr_pt_ctx* ctx, void* cuda_stream);

/* == frames in flight on the device (the reference keeps MAX_FRAMES_IN_FLIGHT = 2 command buffers going,
 * src/context.hh; rt_renderer::render is called for frame k + 1 while frame k executes). tr_pt_render == begin + end.
 * tr_pt_render_begin enqueues, on a stream of the context's own, everything of the next frame that only touches the
 * frame's own state: camera rays, traversal and shading of every bounce (ordered behind what `cuda_stream` holds at the
 * time of the call, e.g. tr_pt_set_cameras_async). tr_pt_render_end enqueues on `cuda_stream`, behind that work, the
 * accumulation into the colour target and the stores of the stitch -- the only part that touches shared state -- for the
 * OLDEST begun frame. Call order `begin(0); { begin(k + 1); end(k); }` keeps two frames' traversal launches queued: the
 * SMs one frame's persistent grids release while they drain go to the other frame. At most two
 * frames may be begun. Results (tr_pt_read_*, tr_pt_get_stats, tr_pt_get_full_image_info) refer to the frame ended
 * last. Frames that need more than one wavefront batch, aux G-buffer layers, samples_per_pass > 1, the megakernel or
 * the instrumentation are rendered whole by tr_pt_render_end (one frame begun at a time). Scene changes
 * (tr_pt_set_scene, tr_pt_update_instances) need every begun frame ended. */
int tr_pt_render_begin(tr_pt_ctx* ctx, void* cuda_stream);
int tr_pt_render_end(tr_pt_ctx* ctx, void* cuda_stream);

/* == feature_stage (src/feature_stage.cc:21-110, shader/rt_feature.rgen): one AOV pass. */
int tr_pt_render_feature(tr_pt_ctx* ctx, int feature, const float default_value[4],
                         void* cuda_stream);

/* Blocks until the context's stream is idle (the reference waits on timeline semaphores). */
int tr_pt_sync(tr_pt_ctx* ctx);

/* Device pointer + geometry of the colour target this context renders into:
 * primary: size.x * size.y * views float4 (the full image, gbuffer `color`, rgba32f);
 * non-primary: the linear strip image of get_distribution_target_size(): count * views float4
 * (src/distribution_strategy.cc:5-19). */
int tr_pt_get_color_buffer(tr_pt_ctx* ctx, void** device_ptr, size_t* bytes);

/* == headless staging copy (src/headless.cc:125-171): D2H of the colour target. */
int tr_pt_read_color(tr_pt_ctx* ctx, float* rgba32f_host, size_t bytes);

/* The same without blocking the host, for frames in flight (the reference keeps MAX_FRAMES_IN_FLIGHT = 2 staging
 * copies going, src/headless.cc:125-171 + src/context.hh): the copy is enqueued on `copy_stream` (a cudaStream_t of the
 * caller, not the render stream), ordered after the frame rendered last; `rgba32f_host` (pinned memory) is valid once
 * `copy_stream` has reached that point. A later tr_pt_render that writes the SAME colour target waits for the copy on
 * the device; one that renders into another target (tr_pt_set_color_buffer with a second buffer) runs beside it. */
int tr_pt_read_color_async(tr_pt_ctx* ctx, float* rgba32f_host, size_t bytes, void* copy_stream);

/* Accumulation checkpoint / resume. The reference has none for accumulation state (animation renders
 * resume per frame with --skip-frames, src/options.hh:116-119); SURVEY.md section 8(f) lists the
 * dump/reload of the fp32 accumulation buffer + counters as the natural addition for long
 * accumulations (BASELINE configs[3]: 4096 spp). The state of a context is its colour target
 * (tr_pt_read_color), rt_stage::frame_counter (-> sampling_data.sample_counter, src/rt_stage.cc:76-85)
 * and rt_camera_stage::accumulated_samples (src/rt_camera_stage.cc:86,100).
 * tr_pt_restore_accumulation writes the colour target (same size as tr_pt_read_color gives for the
 * current cameras + distribution) and sets both counters: the next tr_pt_render continues the
 * running mean exactly where the checkpointed context stopped (bit-identical to an uninterrupted
 * run). The state also records what the counters only make sense with -- samples_per_pixel
 * (sample_counter = frame_counter * samples_per_pixel), the target size and the enabled aux G-buffer
 * layers -- and the restore fails with TR_ERR_INVALID_ARGUMENT when the context differs. With aux layers
 * enabled they are part of the checkpoint (first-sample layers are never written again once
 * accumulated_samples != 0): dump them with tr_pt_read_gbuffer_layer, and after
 * tr_pt_restore_accumulation write every enabled layer back with tr_pt_write_gbuffer_layer;
 * tr_pt_render fails with TR_ERR_INVALID_ARGUMENT while one is still missing. */
typedef struct tr_pt_accumulation_state {
    uint32_t frame_counter;
    uint32_t accumulated_samples;
    uint32_t samples_per_pixel;
    uint32_t gbuffer_mask;
    uint64_t color_pixels;              /* float4 texels of the colour target (all views) */
} tr_pt_accumulation_state;
int tr_pt_get_accumulation_state(tr_pt_ctx* ctx, tr_pt_accumulation_state* out);
int tr_pt_restore_accumulation(tr_pt_ctx* ctx, const float* rgba32f_host, size_t bytes,
                               const tr_pt_accumulation_state* state);

/* == stitch_shuffled_strips.comp:21-63 / stitch_scanline.comp:21-50 for ONE source device (the
 * strategy is the context's distribution strategy). Shuffled strips: scatters `strip`
 * (count*views float4, device memory, e.g. one slice of an all-gather result) of the device that
 * rendered slots [index, index+count) into `full_image` (size.x*size.y*views float4). Scanline:
 * `strip` holds the rows of device `index` of `count` devices (size.x * ceil((size.y-index)/count)
 * float4 per view); row ly lands on image row ly*count+index.
 * blend_ratio < 1 mixes with the image's previous content, mix(old, new, blend_ratio)
 * (stitch_shuffled_strips.comp:50-57; stitch_stage::set_blend_ratio, src/stitch_stage.cc:112). */
int tr_pt_unshuffle(tr_pt_ctx* ctx, const void* strip_device, uint32_t index, uint32_t count,
                    void* full_image_device, float blend_ratio, void* cuda_stream);

/* Fused render->stitch over NVLink peer memory: the last accumulation pass of tr_pt_render
 * also stores every finished pixel into each registered peer image (device pointers mapped
 * into this process, e.g. with cudaIpcOpenMemHandle) at its un-shuffled position, so no
 * separate gather/stitch pass runs. Pass n = 0 to disable. */
int tr_pt_set_peer_images(tr_pt_ctx* ctx, void* const* full_image_device_ptrs, int n);
/* The fused stitch's blend ratio for the NEXT frames (stitch_stage::set_blend_ratio): the reference sets
 * 1/(accumulated_frames+1) for the frame after a workload change under accumulation and 1 again after
 * it (src/rt_renderer.cc:124-128,178-182). Default 1 (plain store). */
int tr_pt_set_stitch_blend(tr_pt_ctx* ctx, float blend_ratio);

/* Cross-process exchange of full images for tr_pt_set_peer_images (one process per GPU):
 * tr_pt_alloc_full_image allocates (or returns) a library-owned size.x*size.y*views float4 image,
 * tr_pt_ipc_get_handle exports it as a 64-byte cudaIpcMemHandle_t, tr_pt_ipc_open_handle maps a
 * peer's handle into this process (cudaIpcOpenMemHandle), tr_pt_ipc_close_handle unmaps it. */
int tr_pt_alloc_full_image(tr_pt_ctx* ctx, void** device_ptr, size_t* bytes);
int tr_pt_read_full_image(tr_pt_ctx* ctx, float* rgba32f_host, size_t bytes);
/* The full image is an arena of two halves, each [colour | every aux layer enabled with tr_pt_set_gbuffer_layers
 * when tr_pt_alloc_full_image ran]. The pointer tr_pt_alloc_full_image returns is the arena base = the colour of
 * half 0, and its `bytes` the size of one colour image, as before. The reference stitches every transferred
 * gbuffer_target entry (src/rt_renderer.cc:370-397, src/stitch_stage.cc:156-171): with peers set, tr_pt_render also
 * stores this context's texels of the enabled aux layers into every peer arena (all arenas share one layout, so the
 * layers, image size and view count must be the same on every rank).
 * tr_pt_set_stitch_double_buffer(1): frames alternate between the halves (the parity of the number of tr_pt_render calls
 * since tr_pt_set_peer_images, so all ranks must render the same frames): frame k+1 of a fast rank stores into the
 * half nobody reads while a slower rank still copies frame k out, which a single image cannot guarantee with one
 * barrier per frame (the reference uses per-frame-in-flight targets + timeline semaphores, src/stitch_stage.cc:128-196).
 * Blending (tr_pt_set_stitch_blend) reads the previous content from the other half. tr_pt_read_full_image /
 * _range / tr_pt_read_full_layer read the half the last tr_pt_render wrote. Default: off (half 0 only).
 * tr_pt_set_render_in_full_image(1): the context's colour target and aux layers ARE its full image (the reference's
 * display device renders in place into the stitched image and keeps accumulating across workload changes,
 * src/rt_renderer.cc:165-175); needs a primary distribution. With double buffering the running means continue from
 * the other half and layers written only by the first sample are carried over. */
typedef struct tr_pt_full_image_info {
    void* base;
    uint64_t arena_bytes;
    uint64_t half_bytes;
    uint64_t layer_offset[TR_GBUFFER_LAYER_COUNT];   /* byte offset inside a half; ~0 = layer absent */
    uint64_t pixels;                                 /* size.x * size.y * views */
    uint32_t generation;                             /* increments whenever the arena is reallocated */
    uint32_t current_half;                           /* the half the last tr_pt_render wrote */
} tr_pt_full_image_info;
int tr_pt_get_full_image_info(tr_pt_ctx* ctx, tr_pt_full_image_info* out);
int tr_pt_set_stitch_double_buffer(tr_pt_ctx* ctx, int enable);
int tr_pt_set_render_in_full_image(tr_pt_ctx* ctx, int enable);
/* D2H of part of the stitched colour image (bytes [byte_offset, byte_offset + bytes) of the current half): lets every
 * rank of a multi-GPU job copy its own share of the image into one shared pinned host buffer over its own PCIe link
 * instead of all ranks (or one) copying everything (headless staging copy, src/headless.cc:125-171). */
int tr_pt_read_full_image_range(tr_pt_ctx* ctx, void* host_dst, size_t byte_offset, size_t bytes);
/* The same without blocking the host (frames in flight): enqueued on `copy_stream`, which the caller has already ordered
 * after the frame's barrier (the fused stitch is complete only once EVERY rank has finished the frame). The copy reads
 * the half the last tr_pt_render wrote; with tr_pt_set_stitch_double_buffer the next frame writes the other half, and
 * the caller lets its copy finish before it passes the NEXT frame's barrier, so no rank can store into a half that is
 * still being read (tauray_b200/renderer.py: read_back_async). `host_dst` must be pinned. */
int tr_pt_read_full_image_range_async(tr_pt_ctx* ctx, void* host_dst, size_t byte_offset, size_t bytes, void* copy_stream);
int tr_pt_read_full_layer(tr_pt_ctx* ctx, int layer, void* host_dst, size_t bytes);
/* Single-process multi-GPU (the reference's own topology, src/context.cc:415): let this context's device
 * store into memory of `peer_cuda_device` (cudaDeviceEnablePeerAccess); then pass the peer context's
 * tr_pt_alloc_full_image pointer to tr_pt_set_peer_images directly, no IPC handle needed. */
int tr_pt_enable_peer_access(tr_pt_ctx* ctx, int peer_cuda_device);
int tr_pt_ipc_get_handle(tr_pt_ctx* ctx, unsigned char handle[64]);
int tr_pt_ipc_open_handle(tr_pt_ctx* ctx, const unsigned char handle[64], void** device_ptr);
int tr_pt_ipc_close_handle(tr_pt_ctx* ctx, void* device_ptr);
/* Render into caller-owned device memory (e.g. a torch tensor registered with NCCL) instead of the
 * library's own colour target; NULL restores the internal target. `bytes` must equal the size
 * tr_pt_get_color_buffer reports for the current cameras/distribution. */
int tr_pt_set_color_buffer(tr_pt_ctx* ctx, void* device_ptr, size_t bytes);

/* Auxiliary G-buffer layers (gbuffer_target entries other than colour): `layer_mask` has bit
 * (1 << TR_GBUFFER_x) set for every layer the target has; the library allocates them with the
 * colour target's layout (views x layer pixels; full image on the primary device, strip otherwise).
 * They are written by tr_pt_render exactly as write_all_outputs does (shader/path_tracer.glsl:
 * 547-574): diffuse/reflection are running means in rgba16f, the others are written when no sample
 * has been accumulated yet. Default: none. With peer images set they are stitched across GPUs like the colour
 * (see tr_pt_get_full_image_info). */
int tr_pt_set_gbuffer_layers(tr_pt_ctx* ctx, uint32_t layer_mask);
int tr_pt_get_gbuffer_layer(tr_pt_ctx* ctx, int layer, void** device_ptr, size_t* bytes);
int tr_pt_read_gbuffer_layer(tr_pt_ctx* ctx, int layer, void* host_dst, size_t bytes);
/* The inverse of tr_pt_read_gbuffer_layer, for resuming from a checkpoint (see tr_pt_restore_accumulation). */
int tr_pt_write_gbuffer_layer(tr_pt_ctx* ctx, int layer, const void* host_src, size_t bytes);

int tr_pt_get_stats(tr_pt_ctx* ctx, tr_pt_stats* stats);
int tr_pt_set_instrumentation(tr_pt_ctx* ctx, int flags);

/* Test hooks (device results copied to the host). */
int tr_pt_read_tri_lights(tr_pt_ctx* ctx, tr_tri_light* out, uint32_t capacity);
/* The environment map's alias table (width * height entries, row-major); returns the number of
 * entries written, or a negative error. */
int tr_pt_read_alias_table(tr_pt_ctx* ctx, tr_alias_table_entry* out, uint32_t capacity);
/* Traces `n` rays (origin.xyz,tmin / dir.xyz,tmax as 8 floats each) through the LBVH with the
 * closest-hit kernel; out = n * {int instance_id, int primitive_id, float u, float v, float t}. */
int tr_pt_trace_rays(tr_pt_ctx* ctx, const float* rays, uint32_t n, void* hits_out);
/* Fills `out` (n*4 uint32) with the sampler draw `draw_index` (-1: film sample) for n
 * pixel coordinates (n*4 uint32: x, y, view, sample). Exercises the device sampler only. */
int tr_pt_sample_test(tr_pt_ctx* ctx, const uint32_t* coords, uint32_t n, int draw_index,
                      uint32_t* out);

const char* tr_pt_last_error(tr_pt_ctx* ctx);   /* ctx may be NULL: last create error */
int tr_pt_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
