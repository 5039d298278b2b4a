// ref_unit_pt.cc -- TEST INFRASTRUCTURE ONLY. The reference's path tracer pipeline (shader/path_tracer.rgen + its hit,
// any-hit, miss and intersection shaders, read from /root/reference and rewritten lexically by build_ref.py) compiled
// as C++ and driven like vkCmdTraceRaysKHR would: one raygen invocation per launch id. Ours in this file: the launch
// loop, the resource bindings and traceRayEXT, which uses a caller-provided ray caster (the oracle's BVH) as the
// "hardware" for triangle candidates and runs the reference's own any-hit / closest-hit / miss / intersection shaders.
#include "glsl_shim.hpp"
#include "ref_api.h"
#include <atomic>
#include <thread>
#include <vector>

namespace glsl { namespace pt {
#include "ref_prelude.hpp"

// ---- the reference's shader text --------------------------------------------------------------------
#include "path_tracer.rgen.h"
#include "rt_common.rahit.h"
#include "rt_common.rchit.h"
#include "rt_common.rmiss.h"
#include "rt_common_shadow.rahit.h"
#include "rt_common_shadow.rchit.h"
#include "rt_common_shadow.rmiss.h"
#include "rt_common_point_light.rint.h"
#include "rt_common_point_light.rchit.h"

// ---- ours: the "device" around it -----------------------------------------------------------------------
static ref_trace_fn g_trace = nullptr;
static void* g_trace_user = nullptr;
static const uint32_t* g_opaque = nullptr;
static std::atomic<uint64_t> g_closest_rays{0}, g_shadow_rays{0};

struct raw_hit { int32_t inst, prim; float u, v, t; };

static void set_hit_builtins(const raw_hit& h)
{
    attribs = vec2(h.u, h.v);
    gl_InstanceCustomIndexEXT = h.inst;
    gl_GeometryIndexEXT = 0;
    gl_PrimitiveID = h.prim;
    gl_HitTEXT = h.t;
}

void traceRayEXT(const accelerationStructureEXT&, uint ray_flags, uint cull_mask, uint sbt_offset, uint, uint miss_index,
                 vec3 origin, float tmin, vec3 direction, float tmax, int)
{
    gl_WorldRayOriginEXT = origin;
    gl_WorldRayDirectionEXT = direction;
    const bool shadow = sbt_offset == 1;
    (shadow ? g_shadow_rays : g_closest_rays)++;
    float cur_tmin = tmin, best_t = tmax;
    bool have_triangle = false;
    raw_hit tri{};
    // Triangle candidates in increasing t: opaque geometry commits at once, the rest goes through the any-hit shader
    // (the order in which a device runs any-hit shaders is unspecified; increasing t is one valid order).
    for(;;)
    {
        const float ray[8] = {origin.x, origin.y, origin.z, cur_tmin, direction.x, direction.y, direction.z, best_t};
        raw_hit h;
        g_trace(g_trace_user, ray, &h);
        if(h.inst < 0) break;
        if(!g_opaque[h.inst])
        {
            set_hit_builtins(h);
            rt_anyhit_action = 0;
            if(shadow) rt_common_shadow_rahit_main();
            else rt_common_rahit_main();
            if(rt_anyhit_action == 1) { cur_tmin = h.t; continue; }     // ignoreIntersectionEXT: look behind it
        }
        tri = h;
        best_t = h.t;
        have_triangle = true;
        break;
    }
    // Procedural geometry: the sphere lights (TLAS mask 0x02, src/scene_stage.cc:1356-1400), intersection shader
    int light = -1;
    if(cull_mask & 0x02u)
    {
        for(uint i = 0; i < scene_metadata.point_light_count; ++i)
        {
            if(point_lights.lights[i].radius == 0.0f) continue;          // degenerate AABB: never entered
            gl_PrimitiveID = (int)i;
            rt_reported_t = -1.0f;
            rt_common_point_light_rint_main();
            if(rt_reported_t > 0.0f && rt_reported_t >= tmin && rt_reported_t <= best_t)
            {
                if(have_triangle && rt_reported_t == best_t) continue;
                best_t = rt_reported_t;
                light = (int)i;
            }
        }
    }
    (void)ray_flags;
    if(light >= 0)
    {
        gl_PrimitiveID = light;
        gl_HitTEXT = best_t;
        if(shadow) rt_common_shadow_rchit_main(); else rt_common_point_light_rchit_main();
    }
    else if(have_triangle)
    {
        set_hit_builtins(tri);
        if(shadow) rt_common_shadow_rchit_main(); else rt_common_rchit_main();
    }
    else if(miss_index == 0) rt_common_rmiss_main();
    else rt_common_shadow_rmiss_main();
}

static std::vector<vertex_buffer> g_vertex_buffers;
static std::vector<index_buffer> g_index_buffers;
static std::vector<sampler2D> g_textures;

static sampler2D make_sampler(const ref_texture& t, int clamp_y)
{
    sampler2D s;
    s.data = t.data; s.width = (int)t.width; s.height = (int)t.height; s.format = (int)t.format; s.clamp_y = clamp_y;
    return s;
}

}} // namespace glsl::pt

using namespace glsl;
using namespace glsl::pt;

static_assert(sizeof(pt::instance) == 288 && sizeof(pt::vertex) == 48 && sizeof(pt::material) == 80, "wire formats");
static_assert(sizeof(pt::directional_light) == 32 && sizeof(pt::point_light) == 64 && sizeof(pt::tri_light) == 64, "wire formats");
static_assert(sizeof(pt::camera_pair) == 640 && sizeof(pt::alias_table_entry) == 16, "wire formats");

extern "C" int ref_pt_render(const ref_scene* s, const ref_frame* f, float* color, float* diffuse, float* reflection, ref_counters* counters)
{
    if(!s || !f || !color || !s->trace) return -1;
    // ---- bindings (set 1: scene, src/scene_stage.cc:1277-1354)
    instances.o = (pt::instance*)s->instances;
    directional_lights.lights = (pt::directional_light*)s->directional_lights;
    point_lights.lights = (pt::point_light*)s->point_lights;
    tri_lights.lights = (pt::tri_light*)s->tri_lights;
    g_vertex_buffers.resize(s->instance_count);
    g_index_buffers.resize(s->instance_count);
    for(uint32_t i = 0; i < s->instance_count; ++i)
    {
        g_vertex_buffers[i].v = (pt::vertex*)s->world_vertices[i];
        g_index_buffers[i].i = (uint*)s->indices[i];
    }
    vertices = g_vertex_buffers.data();
    indices = g_index_buffers.data();
    g_textures.resize(s->texture_count);
    for(uint32_t i = 0; i < s->texture_count; ++i) g_textures[i] = make_sampler(s->textures[i], 0);
    textures = g_textures.data();
    if(s->environment_proj >= 0) environment_map_tex = make_sampler(s->environment_map, 1);
    environment_map_alias_table.entries = (pt::alias_table_entry*)s->alias_table;
    scene_metadata.instance_count = s->instance_count;
    scene_metadata.point_light_count = s->point_light_count;
    scene_metadata.directional_light_count = s->directional_light_count;
    scene_metadata.tri_light_count = s->tri_light_count;
    scene_metadata.environment_factor = vec4(s->environment_factor[0], s->environment_factor[1], s->environment_factor[2], s->environment_factor[3]);
    scene_metadata.environment_proj = s->environment_proj;
    camera.pairs = (pt::camera_pair*)s->cameras;
    g_trace = s->trace; g_trace_user = s->trace_user; g_opaque = s->opaque;
    // ---- set 0: distribution + sampling UBOs, targets, push constants
    distribution.size = uvec2(f->width, f->height);
    distribution.index = 0; distribution.count = 1; distribution.primary = 1;
    distribution.samples_accumulated = f->samples_accumulated;
    sampling_data.sample_counter = f->sample_counter;
    sampling_data.rng_seed = f->rng_seed;
    color_target.data = color; color_target.width = (int)f->width; color_target.height = (int)f->height; color_target.layers = (int)f->views;
    diffuse_target = color_target; diffuse_target.data = diffuse;
    reflection_target = color_target; reflection_target.data = reflection;
    control.samples = SAMPLES_PER_PASS;
    control.previous_samples = f->previous_samples;
    control.min_ray_dist = f->min_ray_dist;
    control.indirect_clamping = f->indirect_clamping;
    control.film_radius = f->film_radius;
    control.russian_roulette_delta = f->russian_roulette_delta;
    control.antialiasing = f->antialiasing;
    control.regularization_gamma = f->regularization_gamma;
    g_closest_rays = 0; g_shadow_rays = 0;

    // ---- vkCmdTraceRaysKHR(width, height, views): one raygen invocation per launch id
    std::atomic<uint32_t> next_row{0};
    const uint32_t rows = f->height * f->views;
    auto worker = [&]() {
        for(;;)
        {
            const uint32_t r = next_row.fetch_add(1);
            if(r >= rows) break;
            const uint32_t view = r / f->height, y = r % f->height;
            for(uint32_t x = 0; x < f->width; ++x)
            {
                gl_LaunchSizeEXT = uvec3(f->width, f->height, f->views);
                gl_LaunchIDEXT = uvec3(x, y, view);
                path_tracer_rgen_main();
            }
        }
    };
    const int threads = f->threads > 0 ? f->threads : (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for(int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for(auto& t: pool) t.join();
    if(counters) { counters->closest_rays = g_closest_rays; counters->shadow_rays = g_shadow_rays; }
    return 0;
}

extern "C" const char* ref_variant_name(void) { return REF_VARIANT_NAME; }
