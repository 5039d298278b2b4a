/* oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU restatement of Tauray's path tracer (the parity oracle). It has the
 * same shape as include/tauray_b200.h so that a test can drive both with the same inputs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product never does (see oracle/README.md).
 */
#ifndef TAURAY_ORACLE_H
#define TAURAY_ORACLE_H
#include "../include/tauray_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

int orc_create(orc_ctx** out, const tr_pt_options* opt, int threads);
void orc_destroy(orc_ctx* ctx);
int orc_set_scene(orc_ctx* ctx, const tr_pt_scene_desc* scene);
int orc_set_cameras(orc_ctx* ctx, const tr_camera_pair* pairs, uint32_t view_count);
int orc_set_distribution(orc_ctx* ctx, const tr_pt_distribution* dist);
int orc_reset_accumulation(orc_ctx* ctx, int reset_sample_counter);
/* One frame: samples_per_pixel / samples_per_pass passes. max_slots > 0 renders only the first
 * max_slots launch ids of every view (a bounded sample for the CPU baseline). */
int orc_render(orc_ctx* ctx, uint64_t max_slots);
int orc_render_feature(orc_ctx* ctx, int feature, const float default_value[4]);
int orc_read_color(orc_ctx* ctx, float* rgba32f, size_t bytes);
int orc_get_stats(orc_ctx* ctx, tr_pt_stats* stats);
int orc_write_color(orc_ctx* ctx, const float* in, size_t bytes);
int orc_read_tri_lights(orc_ctx* ctx, tr_tri_light* out, uint32_t capacity);
int orc_set_gbuffer_layers(orc_ctx* ctx, uint32_t layer_mask);
int orc_read_gbuffer_layer(orc_ctx* ctx, int layer, void* out, size_t bytes);
int orc_read_alias_table(orc_ctx* ctx, tr_alias_table_entry* out, uint32_t capacity);
int orc_trace_rays(orc_ctx* ctx, const float* rays, uint32_t n, void* hits_out);
int orc_sample_test(orc_ctx* ctx, const uint32_t* coords, uint32_t n, int draw_index, uint32_t* out);
/* scatter one strip into a full image on the host (stitch_shuffled_strips.comp:21-63) */
int orc_unshuffle(orc_ctx* ctx, const float* strip, uint32_t index, uint32_t count, float* full_image, float blend_ratio);
/* elementary functions of "TR math", for unit tests: op 0 sin,1 cos,2 atan2(a,b),3 pow(a,b),4 log2,5 exp2,
 * 6 asin, 7 acos */
int orc_math(int op, const float* a, const float* b, float* out, uint32_t n);
/* Single-ray closest hit (thread-safe, no alpha test): user = the orc_ctx. The ray caster of the reference-shader
 * harness (oracle/ref_glsl/ref_api.h: ref_trace_fn). */
int orc_trace_one(void* user, const float* ray8, void* hit_out);
/* The restatement's version of function `fn` of the table in oracle/ref_glsl/ref_unit_fn.cc over n packed records. */
int orc_function(orc_ctx* ctx, int fn, const uint32_t* in, uint32_t* out, uint32_t n);
const char* orc_last_error(orc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
