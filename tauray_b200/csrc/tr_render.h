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

// Per-launch CUDA-event timing (TR_INSTRUMENT_TIMING): one event before every kernel launch on the
// render stream; a launch's duration is the gap to the next event.
enum { TR_KCLASS_CLOSEST = 0, TR_KCLASS_SHADOW = 1, TR_KCLASS_SHADE = 2, TR_KCLASS_OTHER = 3, TR_KCLASS_END = 4 };
struct tr_launch_timer
{
    static constexpr int MAX_MARKS = 8192;
    cudaEvent_t events[MAX_MARKS];
    unsigned char cls[MAX_MARKS];
    int created = 0;
    int used = 0;
    bool overflow = false;
    void reset() { used = 0; overflow = false; }
    void mark(cudaStream_t stream, int kclass)
    {
        if(used >= MAX_MARKS) { overflow = true; return; }
        if(used >= created)
        {
            if(cudaEventCreate(&events[created]) != cudaSuccess) { overflow = true; return; }
            created++;
        }
        cls[used] = (unsigned char)kclass;
        cudaEventRecord(events[used], stream);
        used++;
    }
    void destroy() { for(int i = 0; i < created; ++i) cudaEventDestroy(events[i]); created = used = 0; }
};

// Second stream + events that let the shadow traversal of bounce b run beside the closest-hit traversal of bounce
// b + 1 (both depend only on the shade pass of bounce b): the tail of one persistent grid, where SMs drain, is filled
// by the blocks of the other. Optional (null = everything on the render stream).
struct tr_overlap
{
    cudaStream_t aux = nullptr;
    cudaEvent_t shaded = nullptr;       // recorded on the render stream after k_wf_shade(b)
    cudaEvent_t shadowed = nullptr;     // recorded on aux after k_wf_trace<shadow>(b)
};

cudaError_t tr_render_init_constants();
const uint4* tr_render_sobol_chunks();
// One pass (= samples_per_pass samples of every slot) of path_tracer.rgen.
// `phase`: TR_PHASE_ALL, or -- for frames in flight (tr_pt_render_begin / tr_pt_render_end) -- TR_PHASE_TRACE (everything up to
// the last shade pass: only the wavefront state is written) and TR_PHASE_FINISH (k_wf_finish: the accumulation into the colour
// target and the peer stores) as two calls on different streams. Single-sample passes only.
enum { TR_PHASE_TRACE = 1, TR_PHASE_FINISH = 2, TR_PHASE_ALL = 3 };
cudaError_t tr_render_pass_wavefront(cudaStream_t stream, const trd::RenderParams& P, const tr_wavefront& wf,
                                     int sm_count, uint64_t* launches, tr_launch_timer* timer, uint32_t batch,
                                     const tr_overlap* overlap = nullptr, int phase = TR_PHASE_ALL);
cudaError_t tr_render_pass_megakernel(cudaStream_t stream, const trd::RenderParams& P, int sm_count, uint64_t* launches,
                                      tr_launch_timer* timer);
cudaError_t tr_render_pass_direct(cudaStream_t stream, const trd::RenderParams& P, int sm_count, uint64_t* launches,
                                      tr_launch_timer* timer);
// Copies this rank's texels of the enabled aux G-buffer layers into every peer's full-image arena (end of a frame).
cudaError_t tr_render_stitch_aux_layers(cudaStream_t stream, const trd::RenderParams& P, const void* const src[],
                                        const int src_full[], int sm_count);
cudaError_t tr_render_feature(cudaStream_t stream, const trd::RenderParams& P, int feature, const float default_value[4]);
cudaError_t tr_render_trace_rays(cudaStream_t stream, const trd::RenderParams& P, const float* rays_dev, uint32_t n, void* hits_dev);
cudaError_t tr_render_sample_test(cudaStream_t stream, const trd::RenderParams& P, const uint32_t* coords_dev, uint32_t n,
                                  int draw_index, uint32_t* out_dev);
cudaError_t tr_render_unshuffle(cudaStream_t stream, const trd::FrameParams& f, const float4* strip, uint32_t index,
                                uint32_t count, float4* full_image, float blend_ratio);
cudaError_t tr_render_unscanline(cudaStream_t stream, const trd::FrameParams& f, const float4* strip, uint32_t index,
                                uint32_t count, float4* full_image, float blend_ratio);
