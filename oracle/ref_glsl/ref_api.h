/* ref_api.h -- TEST INFRASTRUCTURE ONLY. C interface of oracle/_ref/libref_glsl_<variant>.so: the reference's own
 * shader code (read from /root/reference at build time, never copied here) running on the CPU. See build_ref.py. */
#ifndef TAURAY_REF_API_H
#define TAURAY_REF_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* closest raw triangle hit of ray = {origin.xyz, tmin, dir.xyz, tmax} with tmin < t < tmax, no alpha test:
 * hit = {int32 instance, int32 primitive, float u, float v, float t}; instance < 0 = none. Must be thread-safe. */
typedef int (*ref_trace_fn)(void* user, const float ray[8], void* hit);

typedef struct ref_texture { const void* data; uint32_t width, height, format, reserved; } ref_texture;

typedef struct ref_scene {
    uint32_t instance_count, point_light_count, directional_light_count, tri_light_count, texture_count, view_count;
    const void* instances;                 /* tr_instance[instance_count] (288 B wire format) */
    const void* const* world_vertices;     /* per instance: pre-transformed tr_vertex[] (pre_transform.comp) */
    const void* const* indices;            /* per instance: uint32[] */
    const uint32_t* opaque;                /* per instance: VK_GEOMETRY_OPAQUE_BIT */
    const void* point_lights;              /* tr_point_light[] */
    const void* directional_lights;        /* tr_directional_light[] */
    const void* tri_lights;                /* tr_tri_light[] (extract_tri_lights.comp) */
    const ref_texture* textures;
    ref_texture environment_map;
    const void* alias_table;               /* tr_alias_table_entry[] */
    const void* cameras;                   /* tr_camera_pair[view_count] */
    float environment_factor[4];
    int32_t environment_proj;
    uint32_t reserved;
    ref_trace_fn trace;
    void* trace_user;
} ref_scene;

typedef struct ref_frame {
    uint32_t width, height, views;
    uint32_t previous_samples;             /* push constant: pass * samples_per_pass */
    uint32_t samples_accumulated;          /* distribution UBO */
    uint32_t sample_counter, rng_seed;     /* sampling UBO */
    int32_t antialiasing;                  /* film != point */
    float min_ray_dist, indirect_clamping, film_radius, russian_roulette_delta, regularization_gamma;
    int32_t threads;
} ref_frame;

typedef struct ref_counters { uint64_t closest_rays, shadow_rays; } ref_counters;

/* One pass of shader/path_tracer.rgen over width x height x views launch ids (DISTRIBUTION_STRATEGY 0); the targets are
 * rgba32f images [view][y][x] that the shader accumulates into exactly as on the device (gbuffer.glsl). */
int ref_pt_render(const ref_scene* scene, const ref_frame* frame, float* color, float* diffuse, float* reflection, ref_counters* counters);
/* shader/pre_transform.comp for one instance: object-space -> world-space vertices */
int ref_pre_transform(const void* instance_record, const void* vertices_in, void* vertices_out, uint32_t vertex_count);
/* shader/extract_tri_lights.comp for one instance (pre-transformed vertices) */
int ref_extract_tri_lights(const void* instances, uint32_t instance_id, const void* world_vertices, const void* indices,
                           uint32_t triangle_count, void* tri_lights_out);
/* Function-level entry: applies function `fn` of the shader include files to n packed argument records
 * (32-bit words, layouts in ref_unit_fn.cc / oracle.cc:orc_function -- the same table on both sides). */
int ref_function(int fn, const uint32_t* in, uint32_t* out, uint32_t n);
int ref_function_words(int fn, int* words_in, int* words_out);
const char* ref_variant_name(void);

#ifdef __cplusplus
}
#endif
#endif
