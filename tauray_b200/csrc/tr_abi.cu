// tr_abi.cu -- the C ABI of include/tauray_b200.h: context, scene upload, frame loop.
//
// Host-side mirror of what the reference does around the hot loop:
//   tr_pt_create            path_tracer_stage ctor                 src/path_tracer_stage.cc:30-116
//   tr_pt_set_scene         scene_stage upload + ensure_blas       src/scene_stage.cc:821-872,1066-1116,1498-1541
//   tr_pt_render            rt_stage::update + record_command_buffers + rt_camera_stage::update
//                           src/rt_stage.cc:76-133, src/rt_camera_stage.cc:74-101, path_tracer_stage.cc:118-147
// There is no CPU fallback anywhere in this file: without a usable sm_100 device every entry point
// fails with TR_ERR_NO_DEVICE / TR_ERR_CUDA.
#include "tr_path.cuh"
#include "tr_render.h"
#include "tr_build.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace trd;

struct tr_pt_ctx
{
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t last_stream = nullptr;   // stream of the last render (may be the caller's)
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // tr_pt_read_color_async: copies still reading a colour target (a render into the same target waits for them)
    struct PendingCopy { const void* src = nullptr; cudaEvent_t done = nullptr; };
    PendingCopy pending_copies[4];
    int pending_next = 0;
    tr_overlap overlap{};                 // second stream for the shadow traversal (tr_render.h)
    tr_pt_options opt{};
    std::string error;

    // scene
    bool has_scene = false;
    DeviceScene scene{};
    std::vector<void*> scene_allocs;
    // cameras / distribution / targets
    tr_camera_pair* cameras = nullptr;
    uint32_t views = 0;
    tr_pt_distribution dist{};
    bool has_dist = false;
    float4* color = nullptr;        // the target in use (own_color or an external buffer)
    float4* own_color = nullptr;
    float stitch_blend = 1.0f;
    uint32_t restore_pending_layers = 0;                    // aux layers a restored checkpoint still has to write back
    uint32_t gbuffer_mask = 0;                              // aux layers of gbuffer_target (tr_pt_set_gbuffer_layers)
    void* gbuffer[TR_GBUFFER_LAYER_COUNT] = {};
    size_t gbuffer_pixels = 0;
    float* prev_from_world = nullptr;                       // scene allocation: 16 floats per instance
    // what tr_pt_update_instances needs to redo the transform-dependent part of the scene on the device
    struct Geometry
    {
        tr_vertex* v_raw = nullptr;          // object-space vertices as uploaded (scene allocation)
        tr_vertex* v_world = nullptr;
        uint32_t* indices = nullptr;
        float* matrices = nullptr;           // 32 floats per instance: model, model_normal
        DeviceInstance* d_instances = nullptr;
        tr_tri_light* d_tri_lights = nullptr;
        float4* shade_tris = nullptr;        // scene allocation: 6 float4 per triangle
        float4* bvh_nodes = nullptr;         // owned here, replaced by every rebuild
        float4* bvh_tris = nullptr;
        std::vector<uint32_t> level_offsets; // breadth-first levels of the node array (refit runs them bottom-up)
        tr_build_instance* d_binst = nullptr;   // refit: binst on the device, allocated on first use
        float* refit_scratch = nullptr;      // refit: 6 floats per node + 1 word
        uint32_t refit_scratch_nodes = 0;
        std::vector<tr_build_instance> binst;
        std::vector<DeviceInstance> dinst;
        std::vector<uint32_t> vertex_counts;
        uint64_t triangle_total = 0;
        uint64_t tri_light_count = 0;
    } geom;
    float4* external_color = nullptr;
    size_t external_bytes = 0;
    size_t color_pixels = 0;
    // library-owned stitched image (tr_pt_alloc_full_image): an arena of two halves (ping-pong per frame when
    // stitch_double_buffer is on), each [colour | enabled aux layers]
    float4* full_image = nullptr;
    size_t full_image_pixels = 0;
    size_t full_half_bytes = 0, full_arena_bytes = 0;
    size_t full_layer_off[TR_GBUFFER_LAYER_COUNT] = {};
    uint32_t full_mask = 0;
    uint32_t full_generation = 0;
    int stitch_double_buffer = 0;
    uint32_t stitch_frame = 0;          // renders since the peers were set: its parity picks the half to write
    uint32_t full_cur_half = 0;         // the half the last render wrote
    int render_in_full = 0;             // the colour target (and aux layers) ARE the full image (display device)
    bool wrote_first_sample_layers = false;
    tr_launch_timer* timer = nullptr;
    tr_wavefront wf{};
    std::vector<void*> wf_allocs;
    Counters* counters = nullptr;
    // frame state (rt_stage::frame_counter, rt_camera_stage::accumulated_samples)
    uint32_t frame_counter = 0;
    int accumulated_samples = 0;
    int instrumentation = 0;
    uint32_t pass_batch = 1;
    tr_pt_stats stats{};
    bool rendered = false;
    int update_mode = TR_UPDATE_REBUILD;     // tr_pt_set_update_mode
    // frames in flight (tr_pt_render_begin / tr_pt_render_end): up to two begun frames, each with its own wavefront state,
    // streams, counters and camera table; `deferred` = the frame is not eligible and is rendered whole by tr_pt_render_end
    struct Flight
    {
        tr_wavefront wf{};
        std::vector<void*> allocs;
        tr_overlap overlap{};
        cudaStream_t stream = nullptr;
        cudaEvent_t traced = nullptr, finished = nullptr, ev_begin = nullptr, ev_end = nullptr;
        Counters* counters = nullptr;
        tr_camera_pair* cameras = nullptr;
        uint32_t camera_views = 0;
        RenderParams* P = nullptr;
        uint32_t batch = 1;
        uint32_t read_half = 0;
        uint64_t launches = 0;
        bool used = false, deferred = false;
        uint32_t deferred_frame_counter = 0;
        int deferred_accumulated = 0;
    };
    Flight flights[2];
    int flight_head = 0, flight_count = 0;
    cudaEvent_t flight_call = nullptr;
    Counters* stats_counters = nullptr;      // whose counters / events tr_pt_get_stats reads (the last finished frame's)
    cudaEvent_t stats_begin = nullptr, stats_end = nullptr;
    float4* peers[8] = {};
    int peer_count = 0;
    int peer_self = -1;
};

static std::string g_create_error;

static int fail(tr_pt_ctx* ctx, int code, const std::string& msg)
{
    if(ctx) ctx->error = msg; else g_create_error = msg;
    return code;
}
static int fail_cuda(tr_pt_ctx* ctx, cudaError_t e, const char* what)
{
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return fail(ctx, e == cudaErrorMemoryAllocation ? TR_ERR_OUT_OF_MEMORY : TR_ERR_CUDA, msg);
}
#define TRY_CUDA(ctx, x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) return fail_cuda(ctx, e_, #x); } while(0)

namespace {
struct DeviceGuard
{
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

template<typename T> cudaError_t dev_alloc(std::vector<void*>& list, T** p, size_t count)
{
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
    if(e == cudaSuccess) list.push_back(*p);
    return e;
}

void free_list(std::vector<void*>& list)
{
    for(void* p: list) cudaFree(p);
    list.clear();
}

uint32_t host_pcg(uint32_t seed) // src/math.cc:280-285
{
    seed = seed * 747796405u + 2891336453u;
    seed = ((seed >> ((seed >> 28) + 4)) ^ seed) * 277803737u;
    return (seed >> 22) ^ seed;
}

unsigned calculate_shuffled_strips_b(uint32_t w, uint32_t h) // src/distribution_strategy.cc:62-69
{
    unsigned n = w * h;
    unsigned b = 31;
    while((n >> b) < 128 && b > 0) b--;
    return b;
}

int check_options(const tr_pt_options* o, std::string& why)
{
    if(o->max_ray_depth < 1 || o->max_ray_depth > 64) { why = "max_ray_depth must be in [1, 64]"; return TR_ERR_INVALID_ARGUMENT; }
    if(o->samples_per_pass < 1 || o->samples_per_pixel < o->samples_per_pass) { why = "bad samples_per_pixel/samples_per_pass"; return TR_ERR_INVALID_ARGUMENT; }
    if(o->projection != TR_PROJECTION_PERSPECTIVE && o->projection != TR_PROJECTION_ORTHOGRAPHIC)
    { why = "only matrix cameras are supported"; return TR_ERR_UNSUPPORTED; }
    if(o->local_sampler < 0 || o->local_sampler > TR_SAMPLER_SOBOL_Z_ORDER_3D) { why = "unknown sampler"; return TR_ERR_INVALID_ARGUMENT; }
    if(o->film < 0 || o->film > 2 || o->mis_mode < 0 || o->mis_mode > 2 || o->bounce_mode < 0 || o->bounce_mode > 2 ||
       o->tri_light_mode < 0 || o->tri_light_mode > 2 || o->kernel_mode < 0 || o->kernel_mode > 1)
    { why = "enum option out of range"; return TR_ERR_INVALID_ARGUMENT; }
    if(o->estimator != TR_ESTIMATOR_PATH_TRACER && o->estimator != TR_ESTIMATOR_DIRECT) { why = "unknown estimator"; return TR_ERR_INVALID_ARGUMENT; }
    return TR_OK;
}

void fill_frame(const tr_pt_ctx* c, FrameParams& f)
{
    const tr_pt_distribution& d = c->dist;
    f.size_x = d.size[0]; f.size_y = d.size[1];
    f.index = d.index; f.count = d.count;
    f.strategy = d.strategy; f.primary = d.primary;
    f.strips_b = calculate_shuffled_strips_b(d.size[0], d.size[1]);
    f.region_size = ((d.size[0] * d.size[1]) + (1u << f.strips_b) - 1) >> f.strips_b;
    switch(d.strategy)
    {
    case TR_DISTRIBUTION_SCANLINE:
        f.ray_count_x = d.size[0];
        f.ray_count_y = (d.size[1] - d.index + d.count - 1) / d.count;
        break;
    case TR_DISTRIBUTION_SHUFFLED_STRIPS:
        f.ray_count_x = d.count; f.ray_count_y = 1;
        break;
    default:
        f.ray_count_x = d.size[0]; f.ray_count_y = d.size[1];
        break;
    }
    f.views = c->views;
    f.layer_size = (d.primary || d.strategy == TR_DISTRIBUTION_DUPLICATE) ? d.size[0] * d.size[1] : f.ray_count_x * f.ray_count_y;
    f.samples_accumulated = 0; f.sample_counter = 0; f.rng_seed = 0; f.previous_samples = 0;
    f.max_sobol_bounces = (uint32_t)(c->opt.max_ray_depth > 8 ? 8 : c->opt.max_ray_depth);
}

void fill_params(const tr_pt_ctx* c, RenderParams& P)
{
    memset(&P, 0, sizeof(P));
    P.scene = c->scene;
    P.opt = c->opt;
    fill_frame(c, P.frame);
    P.cameras = c->cameras;
    P.color = c->color;
    P.counters = c->counters;
    P.sobol_chunks = tr_render_sobol_chunks();
    P.stats = (c->instrumentation & TR_INSTRUMENT_VISITS) ? 1 : 0;
    for(int i = 0; i < c->peer_count; ++i) P.peers[i] = c->peers[i];
    P.peer_count = c->peer_count;
    P.peer_self = c->peer_self;
    P.last_pass = 0;
    P.stitch_blend = c->stitch_blend;
    P.color_prev = c->color;
    for(int l = 1; l < TR_GBUFFER_LAYER_COUNT; ++l) P.gbuffer[l] = ((c->gbuffer_mask >> l) & 1u) ? c->gbuffer[l] : nullptr;
    for(int l = 1; l < TR_GBUFFER_LAYER_COUNT; ++l) P.gbuffer_prev[l] = P.gbuffer[l];
    for(int l = 0; l < TR_GBUFFER_LAYER_COUNT; ++l) P.full_layer[l] = c->full_layer_off[l];
    P.full_cur = P.full_prev = 0;
    P.gbuffer_mask = c->gbuffer_mask;
    P.prev_from_world = c->prev_from_world;
}

// Shared state (cameras, scene, wavefront buffers, counters) may still be in use by a render that runs on a
// caller-provided stream: wait for it before mutating any of it or rendering on another stream.
cudaError_t wait_last_stream(tr_pt_ctx* c, cudaStream_t next)
{
    if(c->last_stream && c->last_stream != next) return cudaStreamSynchronize(c->last_stream);
    return cudaSuccess;
}

// Blocking setters also wait for the frames in flight (tr_pt_render_begin): those run on their own streams.
cudaError_t quiesce(tr_pt_ctx* c, cudaStream_t next)
{
    for(auto& f : c->flights)
        if(f.stream && f.used)
        {
            cudaError_t e = cudaStreamSynchronize(f.stream);
            if(e == cudaSuccess && f.overlap.aux) e = cudaStreamSynchronize(f.overlap.aux);
            if(e != cudaSuccess) return e;
        }
    return wait_last_stream(c, next);
}

const size_t k_layer_texel_bytes[TR_GBUFFER_LAYER_COUNT] = {16, 8, 8, 8, 4, 4, 16, 16, 4};

void free_gbuffer(tr_pt_ctx* c)
{
    for(int l = 1; l < TR_GBUFFER_LAYER_COUNT; ++l)
    {
        if(c->gbuffer[l]) cudaFree(c->gbuffer[l]);
        c->gbuffer[l] = nullptr;
    }
    c->gbuffer_pixels = 0;
}

// (re)allocates the colour target and the wavefront buffers for the current cameras/distribution
static int alloc_wavefront(tr_pt_ctx* c, tr_wavefront& w, std::vector<void*>& allocs, uint32_t slots)
{
    free_list(allocs);
    memset(&w, 0, sizeof(w));
    const size_t n = std::max<uint32_t>(slots, 1);
    TRY_CUDA(c, dev_alloc(allocs, &w.pos_pdf, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.dir_seed, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.atten_reg, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.lobes, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.diffuse, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.reflection, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.fh_emission, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.fh_albedo, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.rs_seed, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.hits, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.nee, 4 * n));
    TRY_CUDA(c, dev_alloc(allocs, &w.sum_color, c->opt.samples_per_pass > 1 ? 3 * n : 1));   // colour, diffuse, reflection sums
    TRY_CUDA(c, dev_alloc(allocs, &w.queue[0], n));
    TRY_CUDA(c, dev_alloc(allocs, &w.queue[1], n));
    TRY_CUDA(c, dev_alloc(allocs, &w.shadow_queue, n));
    TRY_CUDA(c, dev_alloc(allocs, &w.counts, 4 * 64 + 2));
    w.capacity = slots;
    return TR_OK;
}

int prepare_targets(tr_pt_ctx* c)
{
    if(!c->has_scene) return fail(c, TR_ERR_NO_SCENE, "tr_pt_set_scene has not been called");
    if(!c->has_dist || c->views == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "cameras / distribution not set");
    FrameParams f;
    fill_frame(c, f);
    const size_t pixels = (size_t)f.layer_size * f.views;
    if(c->render_in_full && c->full_image)
    {
        c->color = reinterpret_cast<float4*>(reinterpret_cast<char*>(c->full_image) + c->full_cur_half * c->full_half_bytes);
        c->color_pixels = pixels;
    }
    else if(c->external_color)
    {
        if(c->external_bytes != pixels * sizeof(float4))
            return fail(c, TR_ERR_INVALID_ARGUMENT, "external colour buffer has " + std::to_string(c->external_bytes) +
                        " bytes, the target needs " + std::to_string(pixels * sizeof(float4)));
        c->color = c->external_color;
        c->color_pixels = pixels;
    }
    else if(pixels != c->color_pixels || !c->own_color || c->color != c->own_color)
    {
        if(c->own_color) cudaFree(c->own_color);
    free_gbuffer(c);
        c->own_color = nullptr;
        TRY_CUDA(c, cudaMalloc((void**)&c->own_color, std::max<size_t>(pixels, 1) * sizeof(float4)));
        TRY_CUDA(c, cudaMemsetAsync(c->own_color, 0, std::max<size_t>(pixels, 1) * sizeof(float4), c->stream));
        c->color = c->own_color;
        c->color_pixels = pixels;
    }
    // aux G-buffer layers share the colour target's layout
    if(c->gbuffer_pixels != pixels) free_gbuffer(c);
    for(int l = 1; l < TR_GBUFFER_LAYER_COUNT; ++l)
    {
        const bool want = (c->gbuffer_mask >> l) & 1u;
        if(want && !c->gbuffer[l])
        {
            const size_t bytes = std::max<size_t>(pixels, 1) * k_layer_texel_bytes[l];
            TRY_CUDA(c, cudaMalloc(&c->gbuffer[l], bytes));
            TRY_CUDA(c, cudaMemsetAsync(c->gbuffer[l], 0, bytes, c->stream));
        }
        else if(!want && c->gbuffer[l]) { cudaFree(c->gbuffer[l]); c->gbuffer[l] = nullptr; }
    }
    c->gbuffer_pixels = pixels;
    const uint64_t slots64 = (uint64_t)f.ray_count_x * f.ray_count_y * f.views;
    if(slots64 >= (1ull << 31)) return fail(c, TR_ERR_UNSUPPORTED, "more than 2^31 launch slots per pass");
    // passes of one sample each are traced `pass_batch` at a time (more rays in flight per launch)
    const uint32_t pass_count = (uint32_t)(c->opt.samples_per_pixel / c->opt.samples_per_pass);
    uint32_t batch = 1;
    if(c->opt.samples_per_pass == 1 && slots64 > 0)
    {
        const uint64_t budget = 16ull << 20;   // <= 16 Mi slots (~4 GB of state) in flight
        batch = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(pass_count, 8), budget / slots64));
    }
    c->pass_batch = batch;
    const uint32_t slots = (uint32_t)(slots64 * batch);
    if(slots64 * batch >= (1ull << 31)) return fail(c, TR_ERR_UNSUPPORTED, "more than 2^31 launch slots in flight");
    if(c->opt.kernel_mode == TR_KERNEL_WAVEFRONT && c->opt.estimator == TR_ESTIMATOR_PATH_TRACER && (c->wf.capacity < slots || !c->wf.counts))
    {
        const int rc = alloc_wavefront(c, c->wf, c->wf_allocs, slots);
        if(rc) return rc;
    }
    return TR_OK;
}

// Layout of the full-image arena for the current image size, view count and aux layers.
void full_image_layout(const tr_pt_ctx* c, size_t pixels, uint32_t mask, size_t off[TR_GBUFFER_LAYER_COUNT], size_t* half_bytes)
{
    size_t at = 0;
    for(int l = 0; l < TR_GBUFFER_LAYER_COUNT; ++l)
    {
        off[l] = 0;
        if(l != TR_GBUFFER_COLOR && !((mask >> l) & 1u)) continue;
        off[l] = at;
        at += (pixels * k_layer_texel_bytes[l] + 255) & ~(size_t)255;
    }
    *half_bytes = at;
    (void)c;
}

// Picks the half of the full-image arena this frame writes and, for a context that renders in place into its full
// image (the reference's display device), points the targets at it.
int select_full_image_half(tr_pt_ctx* c, RenderParams& P)
{
    if(c->peer_count == 0 && !c->render_in_full) return TR_OK;
    const uint32_t cur = c->stitch_double_buffer ? (c->stitch_frame & 1u) : 0u;
    const uint32_t prev = c->stitch_double_buffer ? (cur ^ 1u) : 0u;
    c->stitch_frame++;
    c->full_cur_half = cur;
    P.full_cur = cur * c->full_half_bytes;
    P.full_prev = prev * c->full_half_bytes;
    if(c->peer_count > 0 && c->gbuffer_mask && (!c->full_image || c->full_mask != c->gbuffer_mask))
    {
        // the peers' arenas are laid out like ours: without our own arena the layer offsets are unknown
        if(c->full_image) return fail(c, TR_ERR_INVALID_ARGUMENT, "aux G-buffer layers changed: call tr_pt_alloc_full_image again (and re-exchange it) before rendering");
        size_t half = 0;
        full_image_layout(c, (size_t)c->dist.size[0] * c->dist.size[1] * c->views, c->gbuffer_mask, c->full_layer_off, &half);
        for(int l = 0; l < TR_GBUFFER_LAYER_COUNT; ++l) P.full_layer[l] = c->full_layer_off[l];
    }
    if(c->render_in_full)
    {
        if(!c->full_image) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_set_render_in_full_image needs tr_pt_alloc_full_image");
        if(!c->dist.primary) return fail(c, TR_ERR_INVALID_ARGUMENT, "rendering in place into the full image needs a primary distribution");
        if(c->full_image_pixels != c->color_pixels || c->full_mask != c->gbuffer_mask)
            return fail(c, TR_ERR_INVALID_ARGUMENT, "the full image does not match the cameras / distribution / layers: call tr_pt_alloc_full_image again");
        char* base = reinterpret_cast<char*>(c->full_image);
        P.color = reinterpret_cast<float4*>(base + P.full_cur);
        P.color_prev = reinterpret_cast<const float4*>(base + P.full_prev);
        for(int l = 1; l < TR_GBUFFER_LAYER_COUNT; ++l)
        {
            if(!((c->gbuffer_mask >> l) & 1u)) continue;
            P.gbuffer[l] = base + P.full_cur + c->full_layer_off[l];
            P.gbuffer_prev[l] = base + P.full_prev + c->full_layer_off[l];
        }
        c->color = P.color;
    }
    return TR_OK;
}

void free_scene(tr_pt_ctx* c)
{
    free_list(c->scene_allocs);
    if(c->geom.bvh_nodes) cudaFree(c->geom.bvh_nodes);
    if(c->geom.bvh_tris) cudaFree(c->geom.bvh_tris);
    if(c->geom.d_binst) cudaFree(c->geom.d_binst);
    if(c->geom.refit_scratch) cudaFree(c->geom.refit_scratch);
    c->geom.d_binst = nullptr; c->geom.refit_scratch = nullptr; c->geom.refit_scratch_nodes = 0; c->geom.level_offsets.clear();
    c->geom = tr_pt_ctx::Geometry();
    c->prev_from_world = nullptr;
    memset(&c->scene, 0, sizeof(c->scene));
    c->has_scene = false;
}
} // namespace

extern "C" {

int tr_pt_abi_version(void) { return TR_B200_ABI_VERSION; }

void tr_pt_default_options(tr_pt_options* o)
{
    if(!o) return;
    memset(o, 0, sizeof(*o));
    o->max_ray_depth = 8;
    o->min_ray_dist = 1e-4f;
    o->rng_seed = 0;
    o->local_sampler = TR_SAMPLER_UNIFORM_RANDOM;
    o->samples_per_pixel = 1;
    o->samples_per_pass = 1;
    o->projection = TR_PROJECTION_PERSPECTIVE;
    o->active_viewport_count = 1;
    o->film = TR_FILM_POINT;
    o->mis_mode = TR_MIS_POWER_HEURISTIC;
    o->film_radius = 0.5f;
    o->sampling_weight_point_lights = 1.0f;
    o->sampling_weight_directional_lights = 1.0f;
    o->sampling_weight_envmap = 1.0f;
    o->sampling_weight_emissive_triangles = 1.0f;
    o->bounce_mode = TR_BOUNCE_MATERIAL;
    o->tri_light_mode = TR_TRI_LIGHT_SOLID_ANGLE;
    o->kernel_mode = TR_KERNEL_WAVEFRONT;
}

int tr_pt_create(tr_pt_ctx** out, int cuda_device, const tr_pt_options* opt)
{
    if(!out || !opt) return fail(nullptr, TR_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    std::string why;
    int rc = check_options(opt, why);
    if(rc) return fail(nullptr, rc, why);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if(e != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        return fail(nullptr, TR_ERR_NO_DEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) + "); tauray_b200 has no CPU fallback");
    }
    if(cuda_device < 0 || cuda_device >= count) return fail(nullptr, TR_ERR_INVALID_ARGUMENT, "cuda_device out of range");
    cudaDeviceProp prop;
    if(cudaGetDeviceProperties(&prop, cuda_device) != cudaSuccess) return fail(nullptr, TR_ERR_NO_DEVICE, "cannot query device");
    if(prop.major != 10)
        return fail(nullptr, TR_ERR_NO_DEVICE, "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    ", this library carries sm_100a code only");
    DeviceGuard guard(cuda_device);
    tr_pt_ctx* c = new tr_pt_ctx();
    c->device = cuda_device;
    c->sm_count = prop.multiProcessorCount;
    c->opt = *opt;
    // direct_stage defines no MIS_* macro: nee_mis_pdf falls through to the light pdf (path_tracer.glsl:78-89)
    if(c->opt.estimator == TR_ESTIMATOR_DIRECT) c->opt.mis_mode = TR_MIS_DISABLED;
    auto bail = [&](cudaError_t err, const char* what) { int r = fail_cuda(nullptr, err, what); tr_pt_destroy(c); return r; };
    if((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if((e = cudaEventCreate(&c->ev_begin)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaEventCreate(&c->ev_end)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaStreamCreateWithFlags(&c->overlap.aux, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if((e = cudaEventCreateWithFlags(&c->overlap.shaded, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaEventCreateWithFlags(&c->overlap.shadowed, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaMalloc((void**)&c->counters, sizeof(Counters))) != cudaSuccess) return bail(e, "cudaMalloc(counters)");
    if((e = cudaMemset(c->counters, 0, sizeof(Counters))) != cudaSuccess) return bail(e, "cudaMemset(counters)");
    if((e = tr_render_init_constants()) != cudaSuccess) return bail(e, "upload of the Sobol table");
    *out = c;
    return TR_OK;
}

void tr_pt_destroy(tr_pt_ctx* c)
{
    if(!c) return;
    DeviceGuard guard(c->device);
    if(c->stream) cudaStreamSynchronize(c->stream);
    for(auto& f : c->flights)
    {
        if(f.stream) cudaStreamSynchronize(f.stream);
        if(f.overlap.aux) { cudaStreamSynchronize(f.overlap.aux); cudaStreamDestroy(f.overlap.aux); }
        if(f.overlap.shaded) cudaEventDestroy(f.overlap.shaded);
        if(f.overlap.shadowed) cudaEventDestroy(f.overlap.shadowed);
        for(cudaEvent_t ev : {f.traced, f.finished, f.ev_begin, f.ev_end}) if(ev) cudaEventDestroy(ev);
        if(f.stream) cudaStreamDestroy(f.stream);
        if(f.counters) cudaFree(f.counters);
        if(f.cameras) cudaFree(f.cameras);
        free_list(f.allocs);
        delete f.P;
    }
    if(c->flight_call) cudaEventDestroy(c->flight_call);
    free_scene(c);
    free_list(c->wf_allocs);
    if(c->cameras) cudaFree(c->cameras);
    if(c->own_color) cudaFree(c->own_color);
    if(c->full_image) cudaFree(c->full_image);
    if(c->timer) { c->timer->destroy(); delete c->timer; }
    if(c->counters) cudaFree(c->counters);
    for(auto& pc : c->pending_copies) if(pc.done) cudaEventDestroy(pc.done);
    if(c->ev_begin) cudaEventDestroy(c->ev_begin);
    if(c->ev_end) cudaEventDestroy(c->ev_end);
    if(c->overlap.shaded) cudaEventDestroy(c->overlap.shaded);
    if(c->overlap.shadowed) cudaEventDestroy(c->overlap.shadowed);
    if(c->overlap.aux) cudaStreamDestroy(c->overlap.aux);
    if(c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// M = model_prev * inverse(model) (column-major wire matrices): what `prev_pos = model_prev * inverse(model) *
// model_pos` (rt.glsl:75) applies to a pre-transformed vertex. GLSL does not pin inverse(); defined here as
// inverse = adjugate / determinant and the product, all in double, rounded once to fp32.
static void previous_from_world(const float model[16], const float model_prev[16], float out[16])
{
    auto at = [&](int row, int col) { return (double)model[col * 4 + row]; };
    auto det3 = [&](const int r[3], const int c[3]) {
        return at(r[0], c[0]) * (at(r[1], c[1]) * at(r[2], c[2]) - at(r[1], c[2]) * at(r[2], c[1]))
             - at(r[0], c[1]) * (at(r[1], c[0]) * at(r[2], c[2]) - at(r[1], c[2]) * at(r[2], c[0]))
             + at(r[0], c[2]) * (at(r[1], c[0]) * at(r[2], c[1]) - at(r[1], c[1]) * at(r[2], c[0]));
    };
    double cofactor[4][4];
    for(int row = 0; row < 4; ++row)
        for(int col = 0; col < 4; ++col)
        {
            int r[3], cidx[3];
            for(int i = 0, k = 0; i < 4; ++i) if(i != row) r[k++] = i;
            for(int i = 0, k = 0; i < 4; ++i) if(i != col) cidx[k++] = i;
            const double m = det3(r, cidx);
            cofactor[row][col] = ((row + col) & 1) ? -m : m;
        }
    const double det = at(0, 0) * cofactor[0][0] + at(0, 1) * cofactor[0][1] + at(0, 2) * cofactor[0][2] + at(0, 3) * cofactor[0][3];
    for(int col = 0; col < 4; ++col)
        for(int row = 0; row < 4; ++row)
        {
            double acc = 0.0;
            for(int k = 0; k < 4; ++k) acc += (double)model_prev[k * 4 + row] * (cofactor[col][k] / det);   // inverse[k][col]
            out[col * 4 + row] = (float)acc;
        }
}

// The host half of environment_map::generate_alias_table (src/environment_map.cc:63-120): normalise the
// importances to mean 1, one sweep pairing under-full ("light") with over-full ("heavy") texels
// (arXiv:1903.00227), then the per-texel solid-angle pdfs. `imp` is normalised in place.
static void build_alias_table(std::vector<float>& imp, uint32_t width, uint32_t height, std::vector<tr_alias_table_entry>& table)
{
    const uint32_t n = width * height;
    double total = 0.0;
    for(float v: imp) total += v;
    const float inv_mean = (float)(1.0 / (total / (double)n));
    for(float& v: imp) v *= inv_mean;

    table.resize(n);
    for(uint32_t k = 0; k < n; ++k) table[k] = tr_alias_table_entry{k, 0xFFFFFFFFu, 1.0f, 1.0f};
    // 2^32 * v as the x86-64 build of the reference converts it: through 64 bits, truncated to 32
    auto fixed32 = [](float v) { return (uint32_t)(long long)std::ldexp(v, 32); };
    auto next_light = [&](uint32_t k) { while(k < n && imp[k] > 1.0f) ++k; return k; };
    auto next_heavy = [&](uint32_t k) { while(k < n && imp[k] <= 1.0f) ++k; return k; };
    uint32_t light = next_light(0), heavy = next_heavy(0);
    float carry = heavy < n ? imp[heavy] : 0.0f;   // what is left of the current heavy texel
    while(heavy < n)
    {
        if(carry > 1.0f)
        {
            // top a light texel up from the heavy one
            if(light >= n) break;   // the reference tests `>` and would write one entry past the end
            table[light].probability = fixed32(imp[light]);
            table[light].alias_id = heavy;
            carry = (carry + imp[light]) - 1.0f;
            light = next_light(light + 1);
        }
        else
        {
            // the heavy texel has become light itself: it is topped up by the next heavy one
            const uint32_t done = heavy;
            table[done].probability = fixed32(carry);
            heavy = next_heavy(heavy + 1);
            if(heavy < n)
            {
                table[done].alias_id = heavy;
                carry = (carry + imp[heavy]) - 1.0f;
            }
        }
    }
    std::vector<float> row_sin(height);
    for(uint32_t y = 0; y < height; ++y) row_sin[y] = (float)std::sin(((float)y + 0.5f) / (float)height * M_PI);
    for(uint32_t k = 0; k < n; ++k)
    {
        const uint32_t a = table[k].alias_id;
        table[k].pdf = (float)(imp[k] / (2.0f * M_PI * M_PI * row_sin[k / width]));
        table[k].alias_pdf = (float)(imp[a] / (2.0f * M_PI * M_PI * row_sin[a / width]));
    }
}

// The transform-dependent part of the scene, all on the device: pre_transform.comp, instance table,
// model_prev * inverse(model), extract_tri_lights.comp and the BVH (the reference's per-frame
// scene_stage::record_as_build / record_tri_light_extraction + instance-buffer refresh,
// src/scene_stage.cc:1572-1675). Used by tr_pt_set_scene and tr_pt_update_instances.
static int transform_and_build(tr_pt_ctx* c, const tr_instance* instances, uint32_t instance_count, bool quality_build, bool refit = false)
{
    tr_pt_ctx::Geometry& g = c->geom;
    cudaStream_t st = c->stream;
    std::vector<float> hm(32 * (size_t)instance_count);
    std::vector<float> prev((size_t)instance_count * 16);
    for(uint32_t i = 0; i < instance_count; ++i)
    {
        memcpy(&hm[32 * (size_t)i], instances[i].model, 64);
        memcpy(&hm[32 * (size_t)i + 16], instances[i].model_normal, 64);
        // model_prev * inverse(model) per instance, for the screen-motion layer (rt.glsl:73-79)
        previous_from_world(instances[i].model, instances[i].model_prev, &prev[(size_t)i * 16]);
        DeviceInstance& d = g.dinst[i];
        d.mat = instances[i].mat;
        d.light_base_id = instances[i].light_base_id;
        d.shadow_terminator_mul = instances[i].shadow_terminator_mul;
        d.triangle_base = g.binst[i].triangle_base;
    }
    if(instance_count)
    {
        TRY_CUDA(c, cudaMemcpyAsync(g.matrices, hm.data(), hm.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        TRY_CUDA(c, cudaMemcpyAsync(g.d_instances, g.dinst.data(), sizeof(DeviceInstance) * g.dinst.size(), cudaMemcpyHostToDevice, st));
        TRY_CUDA(c, cudaMemcpyAsync(c->prev_from_world, prev.data(), sizeof(float) * prev.size(), cudaMemcpyHostToDevice, st));
    }
    for(uint32_t i = 0; i < instance_count; ++i)
        TRY_CUDA(c, tr_build_pre_transform(st, g.v_raw + g.binst[i].vertex_base, g.v_world + g.binst[i].vertex_base, g.vertex_counts[i],
                                           g.matrices + 32 * (size_t)i, g.matrices + 32 * (size_t)i + 16));
    // ---- per-triangle shading records
    if(g.shade_tris)
        for(uint32_t i = 0; i < instance_count; ++i)
            TRY_CUDA(c, tr_build_shade_tris(st, g.v_world + g.binst[i].vertex_base, g.indices + g.binst[i].index_base, g.binst[i].triangle_count,
                                            g.shade_tris + 6 * (size_t)g.binst[i].triangle_base));
    // ---- extract_tri_lights.comp
    if(g.tri_light_count) TRY_CUDA(c, cudaMemsetAsync(g.d_tri_lights, 0, sizeof(tr_tri_light) * g.tri_light_count, st));
    for(uint32_t i = 0; i < instance_count; ++i)
    {
        if(g.dinst[i].light_base_id < 0) continue;
        TRY_CUDA(c, tr_build_extract_tri_lights(st, g.v_world + g.binst[i].vertex_base, g.indices + g.binst[i].index_base, g.binst[i].triangle_count,
                                                i, g.dinst[i].light_base_id, g.dinst[i].mat.emission_factor, g.dinst[i].mat.emission_tex_id, g.d_tri_lights));
    }
    // ---- LBVH
    TRY_CUDA(c, cudaEventRecord(c->ev_begin, st));
    tr_build_input bin;
    bin.vertices = g.v_world; bin.indices = g.indices; bin.instances = g.binst.data();
    bin.instance_count = instance_count; bin.triangle_count = (uint32_t)g.triangle_total;
    bin.quality = quality_build ? 1u : 0u;
    if(refit && g.bvh_nodes && g.bvh_tris && !g.level_offsets.empty())
    {
        // the tree keeps its topology: re-fetch the leaf triangles, recompute the child boxes bottom-up (tr_bvh.cu)
        const uint32_t node_count = c->scene.bvh.node_count;
        if(!g.d_binst && instance_count)
        {
            TRY_CUDA(c, cudaMalloc((void**)&g.d_binst, sizeof(tr_build_instance) * instance_count));
            TRY_CUDA(c, cudaMemcpyAsync(g.d_binst, g.binst.data(), sizeof(tr_build_instance) * instance_count, cudaMemcpyHostToDevice, st));
        }
        if(g.refit_scratch_nodes < node_count)
        {
            if(g.refit_scratch) cudaFree(g.refit_scratch);
            g.refit_scratch = nullptr; g.refit_scratch_nodes = 0;
            TRY_CUDA(c, cudaMalloc((void**)&g.refit_scratch, (6 * (size_t)node_count + 1) * sizeof(float)));
            g.refit_scratch_nodes = node_count;
        }
        TRY_CUDA(c, tr_build_refit_bvh(st, bin, g.d_binst, g.bvh_nodes, g.bvh_tris, node_count, g.level_offsets, g.refit_scratch));
        TRY_CUDA(c, cudaEventRecord(c->ev_end, st));
        TRY_CUDA(c, cudaStreamSynchronize(st));
        float refit_ms = 0.0f;
        cudaEventElapsedTime(&refit_ms, c->ev_begin, c->ev_end);
        c->stats.bvh_build_ms = refit_ms;
        return TR_OK;
    }
    tr_build_output bout;
    TRY_CUDA(c, tr_build_bvh(st, bin, &bout));
    g.level_offsets = bout.level_offsets;
    // the new tree is owned from here on, whatever happens next
    if(g.bvh_nodes) cudaFree(g.bvh_nodes);
    if(g.bvh_tris) cudaFree(g.bvh_tris);
    g.bvh_nodes = bout.nodes; g.bvh_tris = bout.tris;
    c->scene.bvh.nodes = nullptr; c->scene.bvh.tris = nullptr;
    TRY_CUDA(c, cudaEventRecord(c->ev_end, st));
    TRY_CUDA(c, cudaStreamSynchronize(st));   // also covers the host staging vectors above
    float build_ms = 0.0f;
    cudaEventElapsedTime(&build_ms, c->ev_begin, c->ev_end);
    // The wavefront kernel's stack holds at most one sibling group per level plus the postponed triangle groups,
    // and it postpones only while fewer than TR_BVH_POSTPONE_CAP entries are stacked (tr_render.cu).
    if(bout.depth + TR_BVH_POSTPONE_CAP > TR_BVH_STACK)
    {
        free_scene(c);
        return fail(c, TR_ERR_UNSUPPORTED, "BVH depth " + std::to_string(bout.depth) + " exceeds the traversal stack");
    }
    DeviceScene& ds = c->scene;
    ds.bvh.nodes = bout.nodes; ds.bvh.tris = bout.tris; ds.bvh.node_count = bout.node_count; ds.bvh.tri_count = (uint32_t)g.triangle_total;
    ds.bvh.top_nodes = 0;
    c->stats.bvh_build_ms = build_ms;
    c->stats.bvh_node_count = bout.node_count;
    c->stats.bvh_depth = bout.depth;
    return TR_OK;
}

int tr_pt_set_scene(tr_pt_ctx* c, const tr_pt_scene_desc* s)
{
    if(!c || !s) return fail(c, TR_ERR_INVALID_ARGUMENT, "null argument");
    if(s->environment_proj > TR_ENVMAP_LAT_LONG) return fail(c, TR_ERR_UNSUPPORTED, "unknown environment map projection");
    if(s->environment_proj >= 0)
    {
        const tr_texture* e = s->environment_map;
        if(!e || !e->data || !e->width || !e->height) return fail(c, TR_ERR_INVALID_ARGUMENT, "environment_proj >= 0 needs environment_map texels");
        if(e->format != TR_TEXTURE_RGBA16F && e->format != TR_TEXTURE_RGBA32F)
            return fail(c, TR_ERR_INVALID_ARGUMENT, "environment map must be RGBA16F or RGBA32F");
        if((uint64_t)e->width * e->height >= (1ull << 31)) return fail(c, TR_ERR_UNSUPPORTED, "environment map too large");
    }
    if(s->instance_count && (!s->instances || !s->meshes)) return fail(c, TR_ERR_INVALID_ARGUMENT, "null instance tables");
    if(c->flight_count) return fail(c, TR_ERR_INVALID_ARGUMENT, "frames in flight: call tr_pt_render_end first");
    DeviceGuard guard(c->device);
    TRY_CUDA(c, quiesce(c, c->stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    free_scene(c);
    std::vector<void*>& A = c->scene_allocs;
    cudaStream_t st = c->stream;

    // ---- layout of the concatenated buffers
    tr_pt_ctx::Geometry& g = c->geom;
    g.binst.assign(s->instance_count, tr_build_instance());
    g.dinst.assign(s->instance_count, DeviceInstance());
    g.vertex_counts.assign(s->instance_count, 0u);
    std::vector<tr_build_instance>& binst = g.binst;
    std::vector<DeviceInstance>& dinst = g.dinst;
    uint64_t vtotal = 0, itotal = 0, ttotal = 0;
    uint64_t tri_light_count = 0;
    bool any_non_opaque = false;
    for(uint32_t i = 0; i < s->instance_count; ++i)
    {
        const tr_mesh_ref& m = s->meshes[i];
        if(m.index_count % 3) return fail(c, TR_ERR_INVALID_ARGUMENT, "index_count is not a multiple of 3");
        if((m.vertex_count && !m.vertices) || (m.index_count && !m.indices)) return fail(c, TR_ERR_INVALID_ARGUMENT, "null mesh data");
        binst[i].vertex_base = (uint32_t)vtotal;
        binst[i].index_base = (uint32_t)itotal;
        binst[i].triangle_base = (uint32_t)ttotal;
        binst[i].triangle_count = m.index_count / 3;
        binst[i].opaque = m.opaque ? 1u : 0u;
        g.vertex_counts[i] = m.vertex_count;
        if(!m.opaque) any_non_opaque = true;
        DeviceInstance& d = dinst[i];
        memset(&d, 0, sizeof(d));
        d.vertex_base = (uint32_t)vtotal;
        d.index_base = (uint32_t)itotal;
        d.opaque = binst[i].opaque;
        const tr_material& mat = s->instances[i].mat;
        const int32_t tex_ids[4] = {mat.albedo_tex_id, mat.metallic_roughness_tex_id, mat.normal_tex_id, mat.emission_tex_id};
        for(int32_t t: tex_ids)
            if(t >= (int32_t)s->texture_count) return fail(c, TR_ERR_INVALID_ARGUMENT, "texture id out of range");
        if(s->instances[i].light_base_id >= 0)
            tri_light_count = std::max<uint64_t>(tri_light_count, (uint64_t)s->instances[i].light_base_id + binst[i].triangle_count);
        vtotal += m.vertex_count; itotal += m.index_count; ttotal += binst[i].triangle_count;
    }
    if(vtotal >= (1ull << 32) || itotal >= (1ull << 32) || ttotal >= (1ull << 28))
        return fail(c, TR_ERR_UNSUPPORTED, "scene too large (>= 2^28 triangles)");
    g.triangle_total = ttotal;
    g.tri_light_count = tri_light_count;

    // ---- upload (object space; the world-space copy is made by transform_and_build)
    TRY_CUDA(c, dev_alloc(A, &g.v_raw, vtotal));
    TRY_CUDA(c, dev_alloc(A, &g.matrices, 32 * (size_t)s->instance_count));
    TRY_CUDA(c, dev_alloc(A, &g.v_world, vtotal));
    TRY_CUDA(c, dev_alloc(A, &g.indices, itotal));
    TRY_CUDA(c, dev_alloc(A, &g.d_instances, s->instance_count));
    TRY_CUDA(c, dev_alloc(A, &c->prev_from_world, (size_t)s->instance_count * 16));
    TRY_CUDA(c, dev_alloc(A, &g.d_tri_lights, tri_light_count));
    {
        static const bool use_shade_tris = !(getenv("TR_TUNE_SHADE_TRIS") && atoi(getenv("TR_TUNE_SHADE_TRIS")) == 0);
        if(use_shade_tris) TRY_CUDA(c, dev_alloc(A, &g.shade_tris, 6 * (size_t)ttotal));
    }
    for(uint32_t i = 0; i < s->instance_count; ++i)
    {
        const tr_mesh_ref& m = s->meshes[i];
        if(m.vertex_count) TRY_CUDA(c, cudaMemcpyAsync(g.v_raw + binst[i].vertex_base, m.vertices, sizeof(tr_vertex) * (size_t)m.vertex_count, cudaMemcpyHostToDevice, st));
        if(m.index_count) TRY_CUDA(c, cudaMemcpyAsync(g.indices + binst[i].index_base, m.indices, sizeof(uint32_t) * (size_t)m.index_count, cudaMemcpyHostToDevice, st));
    }

    tr_point_light* d_pl = nullptr;
    tr_directional_light* d_dl = nullptr;
    TRY_CUDA(c, dev_alloc(A, &d_pl, s->point_light_count));
    TRY_CUDA(c, dev_alloc(A, &d_dl, s->directional_light_count));
    if(s->point_light_count) TRY_CUDA(c, cudaMemcpyAsync(d_pl, s->point_lights, sizeof(tr_point_light) * s->point_light_count, cudaMemcpyHostToDevice, st));
    if(s->directional_light_count) TRY_CUDA(c, cudaMemcpyAsync(d_dl, s->directional_lights, sizeof(tr_directional_light) * s->directional_light_count, cudaMemcpyHostToDevice, st));

    std::vector<DeviceTexture> htex(s->texture_count);
    for(uint32_t i = 0; i < s->texture_count; ++i)
    {
        const tr_texture& t = s->textures[i];
        if(!t.data || t.width == 0 || t.height == 0 || t.format > TR_TEXTURE_RGBA16F) return fail(c, TR_ERR_INVALID_ARGUMENT, "bad texture");
        const size_t bytes = (size_t)t.width * t.height * (t.format == TR_TEXTURE_RGBA32F ? 16 : t.format == TR_TEXTURE_RGBA16F ? 8 : 4);
        uint8_t* d = nullptr;
        TRY_CUDA(c, dev_alloc(A, &d, bytes));
        TRY_CUDA(c, cudaMemcpyAsync(d, t.data, bytes, cudaMemcpyHostToDevice, st));
        htex[i].data = d; htex[i].width = t.width; htex[i].height = t.height; htex[i].format = t.format; htex[i].pad = 0;
    }
    DeviceTexture* d_tex = nullptr;
    TRY_CUDA(c, dev_alloc(A, &d_tex, s->texture_count));
    if(s->texture_count) TRY_CUDA(c, cudaMemcpyAsync(d_tex, htex.data(), sizeof(DeviceTexture) * htex.size(), cudaMemcpyHostToDevice, st));

    // ---- environment map + alias table (environment_map::generate_alias_table, src/environment_map.cc:39-140)
    DeviceTexture env_tex{};
    tr_alias_table_entry* d_alias = nullptr;
    if(s->environment_proj >= 0)
    {
        const tr_texture& e = *s->environment_map;
        const size_t texels = (size_t)e.width * e.height;
        const size_t bytes = texels * (e.format == TR_TEXTURE_RGBA32F ? 16 : 8);
        uint8_t* d = nullptr;
        float* d_importance = nullptr;
        TRY_CUDA(c, dev_alloc(A, &d, bytes));
        TRY_CUDA(c, dev_alloc(A, &d_alias, texels));
        TRY_CUDA(c, cudaMalloc(&d_importance, sizeof(float) * texels));
        std::vector<float> importance(texels);
        cudaError_t e1 = cudaMemcpyAsync(d, e.data, bytes, cudaMemcpyHostToDevice, st);
        if(e1 == cudaSuccess) e1 = tr_build_envmap_importance(st, d, e.format, e.width, e.height, d_importance);
        if(e1 == cudaSuccess) e1 = cudaMemcpyAsync(importance.data(), d_importance, sizeof(float) * texels, cudaMemcpyDeviceToHost, st);
        if(e1 == cudaSuccess) e1 = cudaStreamSynchronize(st);
        cudaFree(d_importance);
        if(e1 != cudaSuccess) return fail_cuda(c, e1, "environment map importances");
        std::vector<tr_alias_table_entry> table;
        build_alias_table(importance, e.width, e.height, table);
        TRY_CUDA(c, cudaMemcpyAsync(d_alias, table.data(), sizeof(tr_alias_table_entry) * texels, cudaMemcpyHostToDevice, st));
        TRY_CUDA(c, cudaStreamSynchronize(st));   // `table` is pageable host memory
        env_tex.data = d; env_tex.width = e.width; env_tex.height = e.height; env_tex.format = e.format; env_tex.pad = 0;
    }

    // ---- transforms, tri lights, BVH
    {
        // static upload: the quality (binned SAH) build, like the reference's fast-trace static BLAS
        const int rc = transform_and_build(c, s->instances, s->instance_count, true);
        if(rc != TR_OK) return rc;
    }
    DeviceScene& ds = c->scene;
    ds.instances = g.d_instances; ds.vertices = g.v_world; ds.indices = g.indices; ds.shade_tris = g.shade_tris;
    ds.point_lights = d_pl; ds.directional_lights = d_dl; ds.tri_lights = g.d_tri_lights; ds.textures = d_tex;
    ds.instance_count = s->instance_count; ds.point_light_count = s->point_light_count;
    ds.directional_light_count = s->directional_light_count; ds.tri_light_count = (uint32_t)tri_light_count;
    memcpy(ds.environment_factor, s->environment_factor, 16);
    ds.environment_proj = s->environment_proj;
    ds.any_non_opaque = any_non_opaque ? 1u : 0u;
    ds.has_textures = s->texture_count ? 1u : 0u;
    ds.envmap = env_tex;
    ds.alias_table = d_alias;
    c->has_scene = true;
    c->stats.triangle_count = (uint32_t)ttotal;
    c->stats.tri_light_count = (uint32_t)tri_light_count;
    return TR_OK;
}

int tr_pt_update_instances(tr_pt_ctx* c, const tr_instance* instances, uint32_t instance_count)
{
    if(!c || (instance_count && !instances)) return fail(c, TR_ERR_INVALID_ARGUMENT, "null argument");
    if(!c->has_scene) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_update_instances needs a scene");
    if(instance_count != c->scene.instance_count) return fail(c, TR_ERR_INVALID_ARGUMENT, "instance count differs from the scene's");
    const tr_pt_ctx::Geometry& g = c->geom;
    uint64_t tri_light_count = 0;
    for(uint32_t i = 0; i < instance_count; ++i)
        if(instances[i].light_base_id >= 0)
            tri_light_count = std::max<uint64_t>(tri_light_count, (uint64_t)instances[i].light_base_id + g.binst[i].triangle_count);
    if(tri_light_count != g.tri_light_count) return fail(c, TR_ERR_INVALID_ARGUMENT, "the tri-light layout (light_base_id) may not change; use tr_pt_set_scene");
    if(c->flight_count) return fail(c, TR_ERR_INVALID_ARGUMENT, "frames in flight: call tr_pt_render_end first");
    DeviceGuard guard(c->device);
    TRY_CUDA(c, quiesce(c, c->stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    // per-frame update: the fast radix-tree rebuild, or a refit of the tree the scene has (tr_pt_set_update_mode)
    const int rc = transform_and_build(c, instances, instance_count, false, c->update_mode == TR_UPDATE_REFIT);
    if(rc != TR_OK) c->has_scene = false;     // a half-updated scene must not be rendered: the next render fails loudly
    return rc;
}

int tr_pt_set_update_mode(tr_pt_ctx* c, int mode)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(mode != TR_UPDATE_REBUILD && mode != TR_UPDATE_REFIT) return fail(c, TR_ERR_INVALID_ARGUMENT, "unknown update mode");
    c->update_mode = mode;
    return TR_OK;
}

int tr_pt_set_cameras(tr_pt_ctx* c, const tr_camera_pair* pairs, uint32_t view_count)
{
    if(!c || !pairs || view_count == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(c->device);
    TRY_CUDA(c, quiesce(c, c->stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    if(view_count != c->views || !c->cameras)
    {
        if(c->cameras) cudaFree(c->cameras);
        c->cameras = nullptr;
        TRY_CUDA(c, cudaMalloc((void**)&c->cameras, sizeof(tr_camera_pair) * view_count));
        c->views = view_count;
    }
    TRY_CUDA(c, cudaMemcpyAsync(c->cameras, pairs, sizeof(tr_camera_pair) * view_count, cudaMemcpyHostToDevice, c->stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    return TR_OK;
}

int tr_pt_set_cameras_async(tr_pt_ctx* c, const tr_camera_pair* pairs, uint32_t view_count, void* cuda_stream)
{
    if(!c || !pairs || view_count == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "null argument");
    if(!cuda_stream || view_count != c->views || !c->cameras) return tr_pt_set_cameras(c, pairs, view_count);
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // frames rendered on another stream read the table too: they finish first (normally there are none)
    TRY_CUDA(c, wait_last_stream(c, st));
    TRY_CUDA(c, cudaMemcpyAsync(c->cameras, pairs, sizeof(tr_camera_pair) * view_count, cudaMemcpyHostToDevice, st));
    return TR_OK;
}

int tr_pt_set_distribution(tr_pt_ctx* c, const tr_pt_distribution* d)
{
    if(!c || !d) return fail(c, TR_ERR_INVALID_ARGUMENT, "null argument");
    if(d->size[0] == 0 || d->size[1] == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "empty image");
    if(d->strategy < 0 || d->strategy > 2) return fail(c, TR_ERR_INVALID_ARGUMENT, "bad distribution strategy");
    if(d->strategy == TR_DISTRIBUTION_SCANLINE && (d->count == 0 || d->index >= d->count))
        return fail(c, TR_ERR_INVALID_ARGUMENT, "bad scanline index/count");
    if((uint64_t)d->size[0] * d->size[1] >= (1ull << 31)) return fail(c, TR_ERR_UNSUPPORTED, "image too large");
    c->dist = *d;
    if(d->strategy == TR_DISTRIBUTION_DUPLICATE) c->dist.primary = 1;
    c->has_dist = true;
    return TR_OK;
}

int tr_pt_reset_accumulation(tr_pt_ctx* c, int reset_sample_counter)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    c->accumulated_samples = 0;
    c->restore_pending_layers = 0;
    if(reset_sample_counter) c->frame_counter = 0;
    return TR_OK;
}

static int render_whole(tr_pt_ctx* c, void* cuda_stream);

int tr_pt_render(tr_pt_ctx* c, void* cuda_stream)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(c->flight_count) return fail(c, TR_ERR_INVALID_ARGUMENT, "frames in flight: call tr_pt_render_end first");
    return render_whole(c, cuda_stream);
}

static int render_whole(tr_pt_ctx* c, void* cuda_stream)
{
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);
    if(rc) return rc;
    if(c->restore_pending_layers & c->gbuffer_mask)
        return fail(c, TR_ERR_INVALID_ARGUMENT, "restored checkpoint: write the aux G-buffer layers back (tr_pt_write_gbuffer_layer) before rendering");
    // A caller-provided stream is ordered after our own setup work.
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    if(st != c->stream) TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    TRY_CUDA(c, wait_last_stream(c, st));
    // asynchronous read-backs (tr_pt_read_color_async) of the target this frame writes finish first, on the device
    for(auto& pc : c->pending_copies)
        if(pc.src && pc.src == c->color) { TRY_CUDA(c, cudaStreamWaitEvent(st, pc.done, 0)); pc.src = nullptr; }
    RenderParams P;
    fill_params(c, P);
    rc = select_full_image_half(c, P);
    if(rc) return rc;
    // src/rt_stage.cc:76-85 (sampling UBO), src/rt_camera_stage.cc:74-101 (distribution UBO)
    P.frame.sample_counter = c->frame_counter * (uint32_t)c->opt.samples_per_pixel;
    P.frame.rng_seed = c->opt.rng_seed != 0 ? host_pcg((uint32_t)c->opt.rng_seed) : 0;
    P.frame.samples_accumulated = (uint32_t)c->accumulated_samples;
    c->frame_counter++;
    c->accumulated_samples += c->opt.samples_per_pixel;

    TRY_CUDA(c, cudaMemsetAsync(c->counters, 0, sizeof(Counters), st));
    TRY_CUDA(c, cudaEventRecord(c->ev_begin, st));
    uint64_t launches = 0;
    tr_launch_timer* timer = nullptr;
    if(c->instrumentation & TR_INSTRUMENT_TIMING)
    {
        if(!c->timer) c->timer = new tr_launch_timer();
        c->timer->reset();
        timer = c->timer;
    }
    else if(c->timer) c->timer->reset();
    const uint32_t pass_count = (uint32_t)(c->opt.samples_per_pixel / c->opt.samples_per_pass); // path_tracer_stage.cc:37
    for(uint32_t pass = 0; pass < pass_count;)
    {
        const bool direct = c->opt.estimator == TR_ESTIMATOR_DIRECT;      // one kernel organisation: per-thread loop
        const bool wavefront = !direct && c->opt.kernel_mode == TR_KERNEL_WAVEFRONT;
        const uint32_t batch = wavefront ? std::min(c->pass_batch, pass_count - pass) : 1u;
        P.frame.previous_samples = pass * (uint32_t)c->opt.samples_per_pass; // path_tracer_stage.cc:141
        P.last_pass = pass + batch == pass_count;
        cudaError_t e = wavefront ?
            tr_render_pass_wavefront(st, P, c->wf, c->sm_count, &launches, timer, batch, &c->overlap) :
            direct ? tr_render_pass_direct(st, P, c->sm_count, &launches, timer) :
            tr_render_pass_megakernel(st, P, c->sm_count, &launches, timer);
        if(e != cudaSuccess) return fail_cuda(c, e, "render pass launch");
        pass += batch;
    }
    // aux G-buffer layers across devices: this rank's texels -> every peer's arena (src/rt_renderer.cc:370-397)
    if(c->peer_count > 0 && c->gbuffer_mask)
    {
        const bool first_sample_frame = P.frame.samples_accumulated == 0;
        const void* src[TR_GBUFFER_LAYER_COUNT] = {};
        int src_full[TR_GBUFFER_LAYER_COUNT] = {};
        for(int l = 1; l < TR_GBUFFER_LAYER_COUNT; ++l)
        {
            if(!((c->gbuffer_mask >> l) & 1u)) continue;
            const bool running_mean = l == TR_GBUFFER_DIFFUSE || l == TR_GBUFFER_REFLECTION;
            // in-place targets: what this frame did not write is carried over from the half the previous frame wrote
            src[l] = (c->render_in_full && !running_mean && !first_sample_frame) ? P.gbuffer_prev[l] : P.gbuffer[l];
            src_full[l] = 0;
        }
        P.last_pass = 1;
        TRY_CUDA(c, tr_render_stitch_aux_layers(st, P, src, src_full, c->sm_count));
        launches++;
    }
    TRY_CUDA(c, cudaEventRecord(c->ev_end, st));
    c->last_stream = st;
    c->stats.kernel_launches = launches;
    c->stats_counters = c->counters; c->stats_begin = c->ev_begin; c->stats_end = c->ev_end;
    c->rendered = true;
    return TR_OK;
}

// ---- frames in flight --------------------------------------------------------------------------------------------
// tr_pt_render_begin enqueues everything of a frame that only touches the frame's own wavefront state -- camera rays, the
// traversal and shade passes of every bounce -- on a stream of its own; tr_pt_render_end enqueues, on the caller's stream
// and after that work, the one kernel that touches shared state: the accumulation into the colour target and the peer
// stores of the stitch. With begin(k + 1) issued before end(k) the traversal launches of two frames are in the queue
// at any time, and the SMs one frame's persistent grids release while they drain go to the other's.
static bool flight_eligible(const tr_pt_ctx* c)
{
    const uint32_t pass_count = (uint32_t)(c->opt.samples_per_pixel / c->opt.samples_per_pass);
    return c->opt.kernel_mode == TR_KERNEL_WAVEFRONT && c->opt.estimator == TR_ESTIMATOR_PATH_TRACER && c->opt.samples_per_pass == 1 &&
           pass_count <= c->pass_batch && c->gbuffer_mask == 0 && c->instrumentation == 0;
}

int tr_pt_render_begin(tr_pt_ctx* c, void* cuda_stream)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(c->flight_count >= 2) return fail(c, TR_ERR_INVALID_ARGUMENT, "two frames are in flight already: call tr_pt_render_end");
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);
    if(rc) return rc;
    if(c->restore_pending_layers & c->gbuffer_mask)
        return fail(c, TR_ERR_INVALID_ARGUMENT, "restored checkpoint: write the aux G-buffer layers back (tr_pt_write_gbuffer_layer) before rendering");
    tr_pt_ctx::Flight& F = c->flights[(c->flight_head + c->flight_count) % 2];
    // a frame the split does not fit is only noted here and rendered whole, frame by frame, by tr_pt_render_end; while such
    // a frame is pending the following ones are deferred too (the order of the frames is the order of the calls)
    F.deferred = !flight_eligible(c) || (c->flight_count && c->flights[c->flight_head].deferred);
    if(F.deferred)
    {
        // the frame's place in the accumulation is taken NOW (a tr_pt_reset_accumulation for the next frame may follow before
        // this one is ended), like the split path does
        F.deferred_frame_counter = c->frame_counter;
        F.deferred_accumulated = c->accumulated_samples;
        c->frame_counter++;
        c->accumulated_samples += c->opt.samples_per_pixel;
        c->flight_count++;
        return TR_OK;
    }
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    if(!F.stream)
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);              // lo = the LOWEST priority: finish + barrier kernels of the caller go first
        TRY_CUDA(c, cudaStreamCreateWithPriority(&F.stream, cudaStreamNonBlocking, lo));
        TRY_CUDA(c, cudaStreamCreateWithPriority(&F.overlap.aux, cudaStreamNonBlocking, lo));
        TRY_CUDA(c, cudaEventCreateWithFlags(&F.overlap.shaded, cudaEventDisableTiming));
        TRY_CUDA(c, cudaEventCreateWithFlags(&F.overlap.shadowed, cudaEventDisableTiming));
        TRY_CUDA(c, cudaEventCreateWithFlags(&F.traced, cudaEventDisableTiming));
        TRY_CUDA(c, cudaEventCreateWithFlags(&F.finished, cudaEventDisableTiming));
        TRY_CUDA(c, cudaEventCreate(&F.ev_begin));
        TRY_CUDA(c, cudaEventCreate(&F.ev_end));
        TRY_CUDA(c, cudaMalloc((void**)&F.counters, sizeof(Counters)));
        F.P = new RenderParams();
    }
    if(!c->flight_call) TRY_CUDA(c, cudaEventCreateWithFlags(&c->flight_call, cudaEventDisableTiming));
    FrameParams fp;
    fill_frame(c, fp);
    const uint32_t pass_count = (uint32_t)(c->opt.samples_per_pixel / c->opt.samples_per_pass);
    const uint64_t slots = (uint64_t)fp.ray_count_x * fp.ray_count_y * fp.views * pass_count;
    if(F.wf.capacity < slots || !F.wf.counts)
    {
        if(F.used) { TRY_CUDA(c, cudaStreamSynchronize(F.stream)); TRY_CUDA(c, cudaStreamSynchronize(F.overlap.aux)); }
        rc = alloc_wavefront(c, F.wf, F.allocs, (uint32_t)slots);
        if(rc) return rc;
    }
    if(F.camera_views != c->views || !F.cameras)
    {
        if(F.used) TRY_CUDA(c, cudaStreamSynchronize(F.stream));
        if(F.cameras) cudaFree(F.cameras);
        F.cameras = nullptr;
        TRY_CUDA(c, cudaMalloc((void**)&F.cameras, sizeof(tr_camera_pair) * c->views));
        F.camera_views = c->views;
    }
    RenderParams& P = *F.P;
    fill_params(c, P);
    rc = select_full_image_half(c, P);
    if(rc) return rc;
    F.read_half = c->full_cur_half;
    P.frame.sample_counter = c->frame_counter * (uint32_t)c->opt.samples_per_pixel;
    P.frame.rng_seed = c->opt.rng_seed != 0 ? host_pcg((uint32_t)c->opt.rng_seed) : 0;
    P.frame.samples_accumulated = (uint32_t)c->accumulated_samples;
    P.frame.previous_samples = 0;
    P.last_pass = 1;
    P.counters = F.counters;
    P.cameras = F.cameras;
    c->frame_counter++;
    c->accumulated_samples += c->opt.samples_per_pixel;
    F.batch = pass_count;
    F.launches = 0;
    // On the caller's stream, behind whatever it has enqueued (a camera upload): the frame's own copy of the camera table
    // -- the next frame's upload must not change the table under this frame. By then the frame that used this slot two
    // frames ago has been finished on the same stream (tr_pt_render_end), so its kernels no longer read the copy.
    TRY_CUDA(c, wait_last_stream(c, st));
    TRY_CUDA(c, cudaMemcpyAsync(F.cameras, c->cameras, sizeof(tr_camera_pair) * c->views, cudaMemcpyDeviceToDevice, st));
    TRY_CUDA(c, cudaEventRecord(c->flight_call, st));
    TRY_CUDA(c, cudaStreamWaitEvent(F.stream, c->flight_call, 0));
    if(F.used) TRY_CUDA(c, cudaStreamWaitEvent(F.stream, F.finished, 0));     // the slot's wavefront state is free again
    TRY_CUDA(c, cudaMemsetAsync(F.counters, 0, sizeof(Counters), F.stream));
    TRY_CUDA(c, cudaEventRecord(F.ev_begin, F.stream));
    cudaError_t e = tr_render_pass_wavefront(F.stream, P, F.wf, c->sm_count, &F.launches, nullptr, F.batch, &F.overlap, TR_PHASE_TRACE);
    if(e != cudaSuccess) return fail_cuda(c, e, "render pass launch");
    TRY_CUDA(c, cudaEventRecord(F.traced, F.stream));
    F.used = true;
    c->flight_count++;
    return TR_OK;
}

int tr_pt_render_end(tr_pt_ctx* c, void* cuda_stream)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(c->flight_count == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_render_end without tr_pt_render_begin");
    DeviceGuard guard(c->device);
    tr_pt_ctx::Flight& F = c->flights[c->flight_head];
    if(F.deferred)
    {
        c->flight_head = (c->flight_head + 1) % 2;
        c->flight_count--;
        const uint32_t now_counter = c->frame_counter;
        const int now_accumulated = c->accumulated_samples;
        c->frame_counter = F.deferred_frame_counter;
        c->accumulated_samples = F.deferred_accumulated;
        const int rc = render_whole(c, cuda_stream);
        c->frame_counter = now_counter;
        c->accumulated_samples = now_accumulated;
        return rc;
    }
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    RenderParams& P = *F.P;
    // asynchronous read-backs of the target this frame writes finish first, on the device
    for(auto& pc : c->pending_copies)
        if(pc.src && pc.src == P.color) { TRY_CUDA(c, cudaStreamWaitEvent(st, pc.done, 0)); pc.src = nullptr; }
    TRY_CUDA(c, cudaStreamWaitEvent(st, F.traced, 0));
    cudaError_t e = tr_render_pass_wavefront(st, P, F.wf, c->sm_count, &F.launches, nullptr, F.batch, nullptr, TR_PHASE_FINISH);
    if(e != cudaSuccess) return fail_cuda(c, e, "render pass launch");
    TRY_CUDA(c, cudaEventRecord(F.finished, st));
    TRY_CUDA(c, cudaEventRecord(F.ev_end, st));
    TRY_CUDA(c, cudaEventRecord(c->ev_end, st));               // tr_pt_read_color_async orders its copy behind this
    c->color = P.color;
    c->full_cur_half = F.read_half;                            // what the read-backs of this frame address
    c->last_stream = st;
    c->stats.kernel_launches = F.launches;
    c->stats_counters = F.counters; c->stats_begin = F.ev_begin; c->stats_end = F.ev_end;
    c->rendered = true;
    c->flight_head = (c->flight_head + 1) % 2;
    c->flight_count--;
    return TR_OK;
}

int tr_pt_render_feature(tr_pt_ctx* c, int feature, const float default_value[4], void* cuda_stream)
{
    if(!c || !default_value) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);
    if(rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    if(st != c->stream) TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    TRY_CUDA(c, wait_last_stream(c, st));
    RenderParams P;
    fill_params(c, P);
    TRY_CUDA(c, tr_render_feature(st, P, feature, default_value));
    c->last_stream = st;
    return TR_OK;
}

int tr_pt_sync(tr_pt_ctx* c)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if(c->last_stream && c->last_stream != c->stream) TRY_CUDA(c, cudaStreamSynchronize(c->last_stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    return TR_OK;
}

int tr_pt_get_color_buffer(tr_pt_ctx* c, void** device_ptr, size_t* bytes)
{
    if(!c || !device_ptr || !bytes) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);
    if(rc) return rc;
    *device_ptr = c->color;
    *bytes = c->color_pixels * sizeof(float4);
    return TR_OK;
}

int tr_pt_read_color(tr_pt_ctx* c, float* out, size_t bytes)
{
    if(!c || !out) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if(!c->color) return fail(c, TR_ERR_INVALID_ARGUMENT, "nothing rendered yet");
    if(bytes != c->color_pixels * sizeof(float4))
        return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch: target has " + std::to_string(c->color_pixels * sizeof(float4)) + " bytes");
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    TRY_CUDA(c, cudaMemcpyAsync(out, c->color, bytes, cudaMemcpyDeviceToHost, st));
    TRY_CUDA(c, cudaStreamSynchronize(st));
    return TR_OK;
}

int tr_pt_read_color_async(tr_pt_ctx* c, float* out, size_t bytes, void* copy_stream)
{
    if(!c || !out || !copy_stream) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if(!c->color || !c->rendered) return fail(c, TR_ERR_INVALID_ARGUMENT, "nothing rendered yet");
    if(bytes != c->color_pixels * sizeof(float4))
        return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch: target has " + std::to_string(c->color_pixels * sizeof(float4)) + " bytes");
    cudaStream_t cs = (cudaStream_t)copy_stream;
    tr_pt_ctx::PendingCopy& pc = c->pending_copies[c->pending_next];
    c->pending_next = (c->pending_next + 1) % 4;
    if(!pc.done) TRY_CUDA(c, cudaEventCreateWithFlags(&pc.done, cudaEventDisableTiming));
    else if(pc.src) TRY_CUDA(c, cudaEventSynchronize(pc.done));      // a slot is reused only after its copy has finished
    TRY_CUDA(c, cudaStreamWaitEvent(cs, c->ev_end, 0));              // after the frame rendered last
    TRY_CUDA(c, cudaMemcpyAsync(out, c->color, bytes, cudaMemcpyDeviceToHost, cs));
    TRY_CUDA(c, cudaEventRecord(pc.done, cs));
    pc.src = c->color;
    return TR_OK;
}

int tr_pt_get_accumulation_state(tr_pt_ctx* c, tr_pt_accumulation_state* out)
{
    if(!c || !out) return TR_ERR_INVALID_ARGUMENT;
    out->frame_counter = c->frame_counter;
    out->accumulated_samples = (uint32_t)c->accumulated_samples;
    out->samples_per_pixel = (uint32_t)c->opt.samples_per_pixel;
    out->gbuffer_mask = c->gbuffer_mask;
    out->color_pixels = c->color_pixels;
    return TR_OK;
}

int tr_pt_restore_accumulation(tr_pt_ctx* c, const float* rgba32f_host, size_t bytes, const tr_pt_accumulation_state* state)
{
    if(!c || !rgba32f_host || !state) return fail(c, TR_ERR_INVALID_ARGUMENT, "null argument");
    if(state->accumulated_samples > 0x7FFFFFFFu) return fail(c, TR_ERR_INVALID_ARGUMENT, "accumulated_samples out of range");
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);      // needs scene, cameras and distribution: they define the target
    if(rc) return rc;
    if(bytes != c->color_pixels * sizeof(float4) || state->color_pixels != c->color_pixels)
        return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch: target has " + std::to_string(c->color_pixels * sizeof(float4)) + " bytes");
    if(state->samples_per_pixel != (uint32_t)c->opt.samples_per_pixel)
        return fail(c, TR_ERR_INVALID_ARGUMENT, "the checkpoint was written with samples_per_pixel = " + std::to_string(state->samples_per_pixel));
    if(state->gbuffer_mask != c->gbuffer_mask)
        return fail(c, TR_ERR_INVALID_ARGUMENT, "the checkpoint's G-buffer layers differ from the context's");
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    TRY_CUDA(c, cudaMemcpyAsync(c->color, rgba32f_host, bytes, cudaMemcpyHostToDevice, st));
    TRY_CUDA(c, cudaStreamSynchronize(st));
    c->frame_counter = state->frame_counter;
    c->accumulated_samples = (int)state->accumulated_samples;
    // first-sample layers are never written again once samples are accumulated: they have to come from the checkpoint
    c->restore_pending_layers = state->accumulated_samples ? c->gbuffer_mask : 0u;
    return TR_OK;
}

int tr_pt_set_gbuffer_layers(tr_pt_ctx* c, uint32_t layer_mask)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(layer_mask >> TR_GBUFFER_LAYER_COUNT) return fail(c, TR_ERR_INVALID_ARGUMENT, "unknown G-buffer layer");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    TRY_CUDA(c, cudaStreamSynchronize(st));
    c->gbuffer_mask = layer_mask & ~(1u << TR_GBUFFER_COLOR);
    if(c->has_scene && c->has_dist && c->views) return prepare_targets(c);
    return TR_OK;
}

int tr_pt_get_gbuffer_layer(tr_pt_ctx* c, int layer, void** device_ptr, size_t* bytes)
{
    if(!c || layer < 0 || layer >= TR_GBUFFER_LAYER_COUNT) return TR_ERR_INVALID_ARGUMENT;
    if(layer == TR_GBUFFER_COLOR) return tr_pt_get_color_buffer(c, device_ptr, bytes);
    if(!((c->gbuffer_mask >> layer) & 1u)) return fail(c, TR_ERR_INVALID_ARGUMENT, "G-buffer layer not enabled");
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);
    if(rc) return rc;
    if(device_ptr) *device_ptr = c->gbuffer[layer];
    if(c->render_in_full && c->full_image && ((c->full_mask >> layer) & 1u) && device_ptr)   // in-place targets live in the arena
        *device_ptr = reinterpret_cast<char*>(c->full_image) + c->full_cur_half * c->full_half_bytes + c->full_layer_off[layer];
    if(bytes) *bytes = c->gbuffer_pixels * k_layer_texel_bytes[layer];
    return TR_OK;
}

int tr_pt_read_gbuffer_layer(tr_pt_ctx* c, int layer, void* out, size_t bytes)
{
    if(!c || !out || layer < 0 || layer >= TR_GBUFFER_LAYER_COUNT) return TR_ERR_INVALID_ARGUMENT;
    if(layer == TR_GBUFFER_COLOR) return tr_pt_read_color(c, (float*)out, bytes);
    if(!((c->gbuffer_mask >> layer) & 1u) || !c->gbuffer[layer]) return fail(c, TR_ERR_INVALID_ARGUMENT, "G-buffer layer not enabled or nothing rendered yet");
    if(c->render_in_full && c->full_image && ((c->full_mask >> layer) & 1u)) return tr_pt_read_full_layer(c, layer, out, bytes);
    DeviceGuard guard(c->device);
    const size_t have = c->gbuffer_pixels * k_layer_texel_bytes[layer];
    if(bytes != have) return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch: layer has " + std::to_string(have) + " bytes");
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    TRY_CUDA(c, cudaMemcpyAsync(out, c->gbuffer[layer], bytes, cudaMemcpyDeviceToHost, st));
    TRY_CUDA(c, cudaStreamSynchronize(st));
    return TR_OK;
}

int tr_pt_write_gbuffer_layer(tr_pt_ctx* c, int layer, const void* src, size_t bytes)
{
    if(!c || !src || layer <= TR_GBUFFER_COLOR || layer >= TR_GBUFFER_LAYER_COUNT) return fail(c, TR_ERR_INVALID_ARGUMENT, "bad layer");
    if(!((c->gbuffer_mask >> layer) & 1u)) return fail(c, TR_ERR_INVALID_ARGUMENT, "G-buffer layer not enabled");
    DeviceGuard guard(c->device);
    int rc = prepare_targets(c);
    if(rc) return rc;
    const size_t have = c->gbuffer_pixels * k_layer_texel_bytes[layer];
    if(bytes != have) return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch: layer has " + std::to_string(have) + " bytes");
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    TRY_CUDA(c, cudaMemcpyAsync(c->gbuffer[layer], src, bytes, cudaMemcpyHostToDevice, st));
    TRY_CUDA(c, cudaStreamSynchronize(st));
    c->restore_pending_layers &= ~(1u << layer);
    return TR_OK;
}

// D2H of `bytes` bytes at `offset` of layer `layer` in the half of the full-image arena the last render wrote
static int read_full(tr_pt_ctx* c, int layer, void* out, size_t offset, size_t bytes, cudaStream_t async_stream = nullptr)
{
    if(!c || !out || layer < 0 || layer >= TR_GBUFFER_LAYER_COUNT) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if(!c->full_image) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_alloc_full_image has not been called");
    if(layer != TR_GBUFFER_COLOR && !((c->full_mask >> layer) & 1u)) return fail(c, TR_ERR_INVALID_ARGUMENT, "the full image has no such layer");
    const size_t have = c->full_image_pixels * k_layer_texel_bytes[layer];
    if(offset > have || bytes > have - offset) return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch: the layer has " + std::to_string(have) + " bytes");
    if(bytes == 0) return TR_OK;
    const char* src = reinterpret_cast<const char*>(c->full_image) + c->full_cur_half * c->full_half_bytes + c->full_layer_off[layer] + offset;
    if(async_stream)
    {
        TRY_CUDA(c, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, async_stream));
        return TR_OK;
    }
    cudaStream_t st = c->last_stream ? c->last_stream : c->stream;
    TRY_CUDA(c, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, st));
    TRY_CUDA(c, cudaStreamSynchronize(st));
    return TR_OK;
}

int tr_pt_read_full_image_range_async(tr_pt_ctx* c, void* out, size_t byte_offset, size_t bytes, void* copy_stream)
{
    if(!copy_stream) return TR_ERR_INVALID_ARGUMENT;
    return read_full(c, TR_GBUFFER_COLOR, out, byte_offset, bytes, (cudaStream_t)copy_stream);
}

int tr_pt_read_full_image(tr_pt_ctx* c, float* out, size_t bytes)
{
    if(c && c->full_image && bytes != c->full_image_pixels * sizeof(float4)) return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch");
    return read_full(c, TR_GBUFFER_COLOR, out, 0, bytes);
}

int tr_pt_read_full_image_range(tr_pt_ctx* c, void* out, size_t byte_offset, size_t bytes)
{
    return read_full(c, TR_GBUFFER_COLOR, out, byte_offset, bytes);
}

int tr_pt_read_full_layer(tr_pt_ctx* c, int layer, void* out, size_t bytes)
{
    if(c && c->full_image && layer >= 0 && layer < TR_GBUFFER_LAYER_COUNT && bytes != c->full_image_pixels * k_layer_texel_bytes[layer])
        return fail(c, TR_ERR_INVALID_ARGUMENT, "size mismatch");
    return read_full(c, layer, out, 0, bytes);
}

int tr_pt_unshuffle(tr_pt_ctx* c, const void* strip_device, uint32_t index, uint32_t count, void* full_image_device, float blend_ratio,
                    void* cuda_stream)
{
    if(!c || !strip_device || !full_image_device) return TR_ERR_INVALID_ARGUMENT;
    if(!c->has_dist || c->views == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "cameras / distribution not set");
    if(c->dist.strategy == TR_DISTRIBUTION_DUPLICATE) return fail(c, TR_ERR_UNSUPPORTED, "unshuffle: nothing to stitch under the duplicate distribution");
    if(c->dist.strategy == TR_DISTRIBUTION_SCANLINE && (count == 0 || index >= count)) return fail(c, TR_ERR_INVALID_ARGUMENT, "scanline: index must be < count (device index / device count)");
    DeviceGuard guard(c->device);
    FrameParams f;
    fill_frame(c, f);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    if(c->dist.strategy == TR_DISTRIBUTION_SCANLINE)
        TRY_CUDA(c, tr_render_unscanline(st, f, (const float4*)strip_device, index, count, (float4*)full_image_device, blend_ratio));
    else
        TRY_CUDA(c, tr_render_unshuffle(st, f, (const float4*)strip_device, index, count, (float4*)full_image_device, blend_ratio));
    return TR_OK;
}

int tr_pt_set_peer_images(tr_pt_ctx* c, void* const* ptrs, int n)
{
    if(!c || n < 0 || n > 8 || (n > 0 && !ptrs)) return fail(c, TR_ERR_INVALID_ARGUMENT, "0..8 peer images");
    c->peer_self = -1;
    for(int i = 0; i < n; ++i)
    {
        c->peers[i] = (float4*)ptrs[i];
        if(c->full_image && ptrs[i] == (void*)c->full_image) c->peer_self = i;
    }
    c->peer_count = n;
    c->stitch_frame = 0;
    return TR_OK;
}

int tr_pt_set_stitch_blend(tr_pt_ctx* c, float blend_ratio)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(!(blend_ratio >= 0.0f && blend_ratio <= 1.0f)) return fail(c, TR_ERR_INVALID_ARGUMENT, "blend ratio must be in [0, 1]");
    c->stitch_blend = blend_ratio;
    return TR_OK;
}

int tr_pt_get_stats(tr_pt_ctx* c, tr_pt_stats* stats)
{
    if(!c || !stats) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if(c->rendered)
    {
        cudaEvent_t sb = c->stats_begin ? c->stats_begin : c->ev_begin, se = c->stats_end ? c->stats_end : c->ev_end;
        TRY_CUDA(c, cudaEventSynchronize(se));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, sb, se);
        c->stats.render_ms = ms;
        Counters h;
        TRY_CUDA(c, cudaMemcpy(&h, c->stats_counters ? c->stats_counters : c->counters, sizeof(h), cudaMemcpyDeviceToHost));
        c->stats.closest_rays = h.closest_rays;
        c->stats.shadow_rays = h.shadow_rays;
        c->stats.paths = h.paths;
        c->stats.closest_node_visits = h.closest_nodes;
        c->stats.closest_triangle_tests = h.closest_tris;
        c->stats.shadow_node_visits = h.shadow_nodes;
        c->stats.shadow_triangle_tests = h.shadow_tris;
        c->stats.culled_shadow_rays = h.culled_shadow_rays;
        c->stats.closest_ms = c->stats.shadow_ms = c->stats.shade_ms = c->stats.other_ms = 0.0f;
        c->stats.closest_launches = c->stats.shadow_launches = c->stats.shade_launches = 0;
        if(c->timer && c->timer->used > 1)
        {
            tr_launch_timer& t = *c->timer;
            for(int i = 0; i + 1 < t.used; ++i)
            {
                if(t.cls[i] == TR_KCLASS_END) continue;
                float d = 0.0f;
                if(cudaEventElapsedTime(&d, t.events[i], t.events[i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
                switch(t.cls[i])
                {
                case TR_KCLASS_CLOSEST: c->stats.closest_ms += d; c->stats.closest_launches++; break;
                case TR_KCLASS_SHADOW: c->stats.shadow_ms += d; c->stats.shadow_launches++; break;
                case TR_KCLASS_SHADE: c->stats.shade_ms += d; c->stats.shade_launches++; break;
                default: c->stats.other_ms += d; break;
                }
            }
        }
    }
    *stats = c->stats;
    return TR_OK;
}

int tr_pt_set_instrumentation(tr_pt_ctx* c, int flags)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    c->instrumentation = flags & (TR_INSTRUMENT_VISITS | TR_INSTRUMENT_TIMING);
    return TR_OK;
}

int tr_pt_set_color_buffer(tr_pt_ctx* c, void* device_ptr, size_t bytes)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    c->external_color = (float4*)device_ptr;
    c->external_bytes = device_ptr ? bytes : 0;
    if(!device_ptr) { c->color = nullptr; c->color_pixels = 0; }
    return TR_OK;
}

int tr_pt_alloc_full_image(tr_pt_ctx* c, void** device_ptr, size_t* bytes)
{
    if(!c || !device_ptr) return TR_ERR_INVALID_ARGUMENT;
    if(!c->has_dist || c->views == 0) return fail(c, TR_ERR_INVALID_ARGUMENT, "cameras / distribution not set");
    DeviceGuard guard(c->device);
    const size_t pixels = (size_t)c->dist.size[0] * c->dist.size[1] * c->views;
    if(!c->full_image || c->full_image_pixels != pixels || c->full_mask != c->gbuffer_mask)
    {
        TRY_CUDA(c, wait_last_stream(c, c->stream));
        TRY_CUDA(c, cudaStreamSynchronize(c->stream));
        if(c->full_image) cudaFree(c->full_image);
        c->full_image = nullptr;
        c->full_image_pixels = 0;
        size_t half = 0;
        full_image_layout(c, pixels, c->gbuffer_mask, c->full_layer_off, &half);
        TRY_CUDA(c, cudaMalloc((void**)&c->full_image, 2 * half));
        TRY_CUDA(c, cudaMemset(c->full_image, 0, 2 * half));
        c->full_image_pixels = pixels;
        c->full_half_bytes = half;
        c->full_arena_bytes = 2 * half;
        c->full_mask = c->gbuffer_mask;
        c->full_generation++;
        c->full_cur_half = 0;
        c->stitch_frame = 0;
    }
    *device_ptr = c->full_image;
    if(bytes) *bytes = pixels * sizeof(float4);
    return TR_OK;
}

int tr_pt_get_full_image_info(tr_pt_ctx* c, tr_pt_full_image_info* out)
{
    if(!c || !out) return TR_ERR_INVALID_ARGUMENT;
    if(!c->full_image) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_alloc_full_image has not been called");
    memset(out, 0, sizeof(*out));
    out->base = c->full_image;
    out->arena_bytes = c->full_arena_bytes;
    out->half_bytes = c->full_half_bytes;
    for(int l = 0; l < TR_GBUFFER_LAYER_COUNT; ++l)
        out->layer_offset[l] = (l == TR_GBUFFER_COLOR || ((c->full_mask >> l) & 1u)) ? c->full_layer_off[l] : ~0ull;
    out->pixels = c->full_image_pixels;
    out->generation = c->full_generation;
    out->current_half = c->full_cur_half;
    return TR_OK;
}

int tr_pt_set_stitch_double_buffer(tr_pt_ctx* c, int enable)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    c->stitch_double_buffer = enable ? 1 : 0;
    c->stitch_frame = 0;
    return TR_OK;
}

int tr_pt_set_render_in_full_image(tr_pt_ctx* c, int enable)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(enable && !c->full_image) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_alloc_full_image has not been called");
    c->render_in_full = enable ? 1 : 0;
    if(!enable && c->color != c->own_color && c->color != c->external_color) { c->color = nullptr; c->color_pixels = 0; }
    return TR_OK;
}

int tr_pt_enable_peer_access(tr_pt_ctx* c, int peer)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(peer == c->device) return TR_OK;
    DeviceGuard guard(c->device);
    int can = 0;
    TRY_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, peer));
    if(!can) return fail(c, TR_ERR_UNSUPPORTED, "no peer access between the devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if(e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return TR_OK; }
    if(e != cudaSuccess) return fail_cuda(c, e, "cudaDeviceEnablePeerAccess");
    return TR_OK;
}

int tr_pt_ipc_get_handle(tr_pt_ctx* c, unsigned char handle[64])
{
    if(!c || !handle) return TR_ERR_INVALID_ARGUMENT;
    if(!c->full_image) return fail(c, TR_ERR_INVALID_ARGUMENT, "tr_pt_alloc_full_image has not been called");
    DeviceGuard guard(c->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    TRY_CUDA(c, cudaIpcGetMemHandle(&h, c->full_image));
    memcpy(handle, &h, 64);
    return TR_OK;
}

int tr_pt_ipc_open_handle(tr_pt_ctx* c, const unsigned char handle[64], void** device_ptr)
{
    if(!c || !handle || !device_ptr) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    TRY_CUDA(c, cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TR_OK;
}

int tr_pt_ipc_close_handle(tr_pt_ctx* c, void* device_ptr)
{
    if(!c || !device_ptr) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    TRY_CUDA(c, cudaIpcCloseMemHandle(device_ptr));
    return TR_OK;
}

int tr_pt_read_tri_lights(tr_pt_ctx* c, tr_tri_light* out, uint32_t capacity)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(!c->has_scene) return fail(c, TR_ERR_NO_SCENE, "no scene");
    DeviceGuard guard(c->device);
    const uint32_t n = std::min(capacity, c->scene.tri_light_count);
    if(out && n) TRY_CUDA(c, cudaMemcpy(out, c->scene.tri_lights, sizeof(tr_tri_light) * n, cudaMemcpyDeviceToHost));
    return (int)c->scene.tri_light_count;
}

int tr_pt_read_alias_table(tr_pt_ctx* c, tr_alias_table_entry* out, uint32_t capacity)
{
    if(!c) return TR_ERR_INVALID_ARGUMENT;
    if(!c->has_scene) return fail(c, TR_ERR_NO_SCENE, "no scene");
    DeviceGuard guard(c->device);
    const uint32_t total = c->scene.environment_proj >= 0 ? c->scene.envmap.width * c->scene.envmap.height : 0u;
    const uint32_t n = std::min(capacity, total);
    if(out && n) TRY_CUDA(c, cudaMemcpy(out, c->scene.alias_table, sizeof(tr_alias_table_entry) * n, cudaMemcpyDeviceToHost));
    return (int)total;
}

int tr_pt_trace_rays(tr_pt_ctx* c, const float* rays, uint32_t n, void* hits_out)
{
    if(!c || (n && (!rays || !hits_out))) return TR_ERR_INVALID_ARGUMENT;
    if(!c->has_scene) return fail(c, TR_ERR_NO_SCENE, "no scene");
    DeviceGuard guard(c->device);
    if(n == 0) return TR_OK;
    float* d_rays = nullptr;
    void* d_hits = nullptr;
    std::vector<void*> tmp;
    struct TmpFree { std::vector<void*>& l; ~TmpFree() { free_list(l); } } tmp_free{tmp};
    TRY_CUDA(c, dev_alloc(tmp, &d_rays, 8 * (size_t)n));
    TRY_CUDA(c, dev_alloc(tmp, (uint8_t**)&d_hits, 20 * (size_t)n));
    TRY_CUDA(c, cudaMemcpyAsync(d_rays, rays, 32 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    RenderParams P;
    memset(&P, 0, sizeof(P));
    P.scene = c->scene;
    P.opt = c->opt;
    TRY_CUDA(c, tr_render_trace_rays(c->stream, P, d_rays, n, d_hits));
    TRY_CUDA(c, cudaMemcpyAsync(hits_out, d_hits, 20 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    return TR_OK;
}

int tr_pt_sample_test(tr_pt_ctx* c, const uint32_t* coords, uint32_t n, int draw_index, uint32_t* out)
{
    if(!c || (n && (!coords || !out))) return TR_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(c->device);
    if(n == 0) return TR_OK;
    uint32_t *d_in = nullptr, *d_out = nullptr;
    std::vector<void*> tmp;
    struct TmpFree { std::vector<void*>& l; ~TmpFree() { free_list(l); } } tmp_free{tmp};
    TRY_CUDA(c, dev_alloc(tmp, &d_in, 4 * (size_t)n));
    TRY_CUDA(c, dev_alloc(tmp, &d_out, 4 * (size_t)n));
    TRY_CUDA(c, cudaMemcpyAsync(d_in, coords, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    RenderParams P;
    memset(&P, 0, sizeof(P));
    P.opt = c->opt;
    P.frame.sample_counter = c->frame_counter * (uint32_t)c->opt.samples_per_pixel;
    P.frame.rng_seed = c->opt.rng_seed != 0 ? host_pcg((uint32_t)c->opt.rng_seed) : 0;
    P.frame.max_sobol_bounces = (uint32_t)(c->opt.max_ray_depth > 8 ? 8 : c->opt.max_ray_depth);
    P.sobol_chunks = tr_render_sobol_chunks();
    TRY_CUDA(c, tr_render_sample_test(c->stream, P, d_in, n, draw_index, d_out));
    TRY_CUDA(c, cudaMemcpyAsync(out, d_out, 16 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    TRY_CUDA(c, cudaStreamSynchronize(c->stream));
    return TR_OK;
}

const char* tr_pt_last_error(tr_pt_ctx* c) { return c ? c->error.c_str() : g_create_error.c_str(); }

} // extern "C"
