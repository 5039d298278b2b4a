// tr_scene.cuh -- device-side scene, BVH and launch-parameter structs shared by the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/tauray_b200.h"

namespace trd {

// ---- BVH layout in HBM -----------------------------------------------------
// 8-wide compressed node = 5 x float4 = 80 B (after Ylitie, Karras, Laine 2017), five 128-bit loads:
//   n0 = (p.x, p.y, p.z, [ex | ey<<8 | ez<<16])                   p = box origin, e* = exponent byte of
//                                                                  the per-axis scale 2^e
//   n1 = (child_base, triangle_base, imask_oct[0..3], [leaf8 | two8<<8])
//                                                                  imask_oct[o] = slots holding inner
//                                                                  children, as bit (slot ^ o): the mask
//                                                                  in the order a ray of octant o visits
//                                                                  them (imask_oct[0] = plain slot mask;
//                                                                  octant o >= 4 = bit reversal of o ^ 7);
//                                                                  leaf8 = slots holding a leaf, two8 =
//                                                                  leaf slots with two triangles
//   n2 = (qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7])     quantised child boxes, one byte per
//   n3 = (qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7])     slot and bound: lo = p + q * 2^e,
//   n4 = (qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7])     rounded outwards
// The j-th inner child (in slot order) is node child_base + j. Children sit in slots chosen per octant
// so that visiting (slot ^ octinv) from high to low is roughly front to back.
// Triangle = 3 x float4 = 48 B; the triangles of one node are contiguous from triangle_base in slot
// order, one or two per leaf slot:
//   t0 = (v0.xyz, instance_id bits), t1 = (v1.xyz, primitive_id bits), t2 = (v2.xyz, flags bits)
// flags bit0 = opaque geometry (VK_GEOMETRY_OPAQUE_BIT, src/scene_stage.cc:1410).
#define TR_BVH_NODE_VEC4 5
#define TR_BVH_LEAF_TRIS 2
#define TR_BVH_STACK 48
// the wavefront kernel postpones a triangle group (one extra stack entry) only while fewer entries than this are stacked
#define TR_BVH_POSTPONE_CAP 16
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
    uint32_t triangle_base;   // first record of the instance in DeviceScene::shade_tris
    uint32_t pad[2];
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
    // Per-triangle shading records, 6 float4 = 96 B: for each of the three vertices (pos.xyz, normal.x) (normal.yz, uv.xy)
    // -- the first 32 bytes of the world-space tr_vertex, gathered through the index buffer once per scene (update), so
    // that shading a hit is ONE contiguous 96-byte fetch instead of indices -> three scattered vertices.
    const float4* shade_tris;
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
    uint32_t has_textures;                         // texture_count > 0: otherwise every material texture id is -1
    DeviceTexture envmap;                          // lat-long texels (environment_proj >= 0)
    const tr_alias_table_entry* alias_table;       // envmap.width * envmap.height entries
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
    unsigned long long culled_shadow_rays;
};

} // namespace trd
