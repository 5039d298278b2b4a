// tr_build.h -- host entry points of the on-device scene preprocessing / LBVH build (tr_bvh.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>
#include "../../include/tauray_b200.h"

struct tr_build_instance
{
    uint32_t vertex_base;     // into the concatenated world-space vertex buffer
    uint32_t index_base;
    uint32_t triangle_base;   // running triangle count before this instance
    uint32_t triangle_count;
    uint32_t opaque;
};

struct tr_build_input
{
    const tr_vertex* vertices;            // device, world space
    const uint32_t* indices;              // device
    const tr_build_instance* instances;   // host
    uint32_t instance_count;
    uint32_t triangle_count;
    uint32_t quality;                     // 1: binned-SAH hierarchy (static scenes), 0: radix tree (per-frame rebuilds)
};

struct tr_build_output
{
    float4* nodes;            // device, TR_BVH_NODE_VEC4 float4 per node, breadth-first order; caller frees
    float4* tris;             // device, 3 float4 per triangle in leaf order; caller frees
    uint32_t node_count;
    uint32_t depth;           // number of breadth-first levels
    float eps;                // box inflation
    std::vector<uint32_t> level_offsets;   // depth + 1 entries: nodes [level_offsets[l], level_offsets[l + 1]) form level l
};

cudaError_t tr_build_pre_transform(cudaStream_t stream, const tr_vertex* in, tr_vertex* out, uint32_t count,
                                   const float* model_dev, const float* model_normal_dev);
// 96-byte shading record per triangle (pos, normal, uv of its three world-space vertices), see DeviceScene::shade_tris
cudaError_t tr_build_shade_tris(cudaStream_t stream, const tr_vertex* verts, const uint32_t* indices, uint32_t triangle_count, float4* out);
cudaError_t tr_build_extract_tri_lights(cudaStream_t stream, const tr_vertex* verts, const uint32_t* indices,
                                        uint32_t triangle_count, uint32_t instance_id, int32_t light_base_id,
                                        const float emission[3], int32_t emission_tex_id, tr_tri_light* out);
// importance = luminance * texel solid angle of a lat-long environment map (RGBA16F / RGBA32F texels on the device)
cudaError_t tr_build_envmap_importance(cudaStream_t stream, const void* texels, uint32_t format, uint32_t width, uint32_t height,
                                       float* importance_dev);
cudaError_t tr_build_bvh(cudaStream_t stream, const tr_build_input& in, tr_build_output* out);
// Refit of an existing wide tree to moved vertices (the BLAS / TLAS `update` of src/acceleration_structure.cc:187-262,
// :376-443): the topology, the slot assignment and the triangle order stay, the leaf triangles are re-fetched and the
// quantised child boxes recomputed level by level from the bottom. `instances_dev` = in.instances on the device,
// `scratch` = 6 floats per node + 1 word.
cudaError_t tr_build_refit_bvh(cudaStream_t stream, const tr_build_input& in, const tr_build_instance* instances_dev,
                               float4* nodes, float4* tris, uint32_t node_count, const std::vector<uint32_t>& level_offsets,
                               float* scratch);
