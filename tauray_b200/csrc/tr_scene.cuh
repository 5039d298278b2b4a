// tr_scene.cuh -- device-side scene, BVH and launch-parameter structs shared by the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/tauray_b200.h"

namespace trd {

// ---- BVH layout in HBM -----------------------------------------------------
// Node = 4 x float4 = 64 B, one 128-bit load each (Aila/Laine layout):
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   n3 = (ref0, ref1, -, -) as int bits
// ref >= 0: inner node index; ref < 0: leaf, ~ref = (first_triangle << 3) | (count - 1), count <= 8.
// An absent child has an inverted box (lo = +inf, hi = -inf) and is never entered.
// Triangle = 3 x float4 = 48 B in leaf (Morton) order:
//   t0 = (v0.xyz, instance_id bits), t1 = (v1.xyz, primitive_id bits), t2 = (v2.xyz, flags bits)
// flags bit0 = opaque geometry (VK_GEOMETRY_OPAQUE_BIT, src/scene_stage.cc:1410).
#define TR_BVH_MAX_LEAF 4
#define TR_BVH_STACK 96
#define TR_TRI_FLAG_OPAQUE 1u

struct DeviceBvh
{
    const float4* nodes;
    const float4* tris;
    uint32_t node_count;
    uint32_t tri_count;
    uint32_t top_nodes;     // leading nodes (BFS order) that fit the shared-memory top-of-tree cache
};

// Per-instance shading record: the material (src/scene_stage.cc:17-30) plus where the instance's
// pre-transformed vertices / indices start in the concatenated buffers.
struct DeviceInstance
{
    tr_material mat;          // 80 B
    int32_t light_base_id;
    uint32_t vertex_base;
    uint32_t index_base;
    uint32_t opaque;
    float shadow_terminator_mul;
    uint32_t pad[3];
};                            // 112 B

struct DeviceTexture
{
    const void* data;
    uint32_t width, height, format, pad;
};

struct DeviceScene
{
    DeviceBvh bvh;
    const DeviceInstance* instances;
    const tr_vertex* vertices;      // world space (pre_transform.comp)
    const uint32_t* indices;
    const tr_point_light* point_lights;
    const tr_directional_light* directional_lights;
    const tr_tri_light* tri_lights;
    const DeviceTexture* textures;
    uint32_t instance_count;
    uint32_t point_light_count;
    uint32_t directional_light_count;
    uint32_t tri_light_count;
    float environment_factor[4];
    int32_t environment_proj;
    uint32_t any_non_opaque;
};

// distribution UBO (src/rt_camera_stage.cc:17-24) + push constants (src/path_tracer_stage.cc:10-20)
// + sampling UBO (src/rt_stage.cc:16-20), flattened.
struct FrameParams
{
    uint32_t size_x, size_y;
    uint32_t index;
    uint32_t count;             // scanline: device count; shuffled strips: slots of this device
    uint32_t strips_b;          // calculate_shuffled_strips_b(size)
    uint32_t region_size;
    int32_t strategy;
    int32_t primary;
    uint32_t samples_accumulated;
    uint32_t ray_count_x, ray_count_y;
    uint32_t views;
    uint32_t layer_size;        // pixels per view in the colour target
    uint32_t sample_counter;
    uint32_t rng_seed;
    uint32_t previous_samples;  // pass * samples_per_pass
    uint32_t max_sobol_bounces;
};

struct Counters
{
    unsigned long long closest_rays, shadow_rays, paths;
    unsigned long long closest_nodes, closest_tris, shadow_nodes, shadow_tris;
};

} // namespace trd
