// tr_render.h -- host-side launch interface of the render kernels (tr_render.cu).
#pragma once
#include <cuda_runtime.h>
#include "tr_scene.cuh"

// Wavefront path state, structure-of-arrays in HBM; one entry per launch slot of the pass
// (slot = (view, y, x) of the launch grid). 16-byte records so every access is one
// coalesced 128-bit load/store per thread.
struct tr_wavefront
{
    uint32_t capacity;          // slots
    float4* pos_pdf;            // ray origin / hit position, bsdf_pdf
    float4* dir_seed;           // ray direction, payload.random_seed (bits)
    float4* atten_reg;          // attenuation, regularization
    float4* lobes;              // primary_lobes
    float4* diffuse;            // diffuse accumulator (rgb)
    float4* reflection;         // reflection accumulator (rgb)
    float4* fh_emission;        // first-hit emission (rgb)
    float4* fh_albedo;          // first-hit albedo rgb + metallic... (albedo.a in fh_emission.w)
    uint4* rs_seed;             // PCG stream state (uniform-random sampler only)
    int4* hits;                 // instance, primitive, u, v (bits)
    float4* nee;                // 4 float4 per slot: the pending NeeJob
    float4* sum_color;          // per-pass sample sum when samples_per_pass > 1
    uint32_t* queue[2];         // active path ids, ping-pong
    uint32_t* shadow_queue;
    uint32_t* counts;           // [0..2*D): active counts per bounce (ping-pong slot b), shadow counts
};

namespace trd { struct RenderParams; }

cudaError_t tr_render_init_constants();
// One pass (= samples_per_pass samples of every slot) of path_tracer.rgen.
cudaError_t tr_render_pass_wavefront(cudaStream_t stream, const trd::RenderParams& P, const tr_wavefront& wf,
                                     int sm_count, uint64_t* launches);
cudaError_t tr_render_pass_megakernel(cudaStream_t stream, const trd::RenderParams& P, int sm_count, uint64_t* launches);
cudaError_t tr_render_feature(cudaStream_t stream, const trd::RenderParams& P, int feature, const float default_value[4]);
cudaError_t tr_render_trace_rays(cudaStream_t stream, const trd::RenderParams& P, const float* rays_dev, uint32_t n, void* hits_dev);
cudaError_t tr_render_sample_test(cudaStream_t stream, const trd::RenderParams& P, const uint32_t* coords_dev, uint32_t n,
                                  int draw_index, uint32_t* out_dev);
cudaError_t tr_render_unshuffle(cudaStream_t stream, const trd::FrameParams& f, const float4* strip, uint32_t index,
                                uint32_t count, float4* full_image);
