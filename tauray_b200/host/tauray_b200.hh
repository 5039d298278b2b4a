// tauray_b200.hh -- C++ host mirror of the reference's stage/renderer interface for the path tracer,
// sitting directly above the C ABI (include/tauray_b200.h). Same names, argument meaning and error
// behaviour (std::runtime_error, as src/main.cc:36-42 expects) as:
//   tr::distribution_strategy / distribution_params / get_device_distribution_params ...
//                                   src/distribution_strategy.{hh,cc}
//   tr::path_tracer_stage           src/path_tracer_stage.{hh,cc}  (+ rt_camera_stage / rt_stage options)
//   tr::rt_renderer<path_tracer_stage>  src/rt_renderer.{hh,cc}    (render, reset_accumulation,
//                                                                   set_device_workloads)
//   tr::load_balancer               src/load_balancer.{hh,cc}
// A Tauray maintainer swaps `#include "path_tracer_stage.hh"` for this header in the renderer that
// should run on B200s. Header-only; link with -ltauray_b200.
#pragma once
#include "../../include/tauray_b200.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace tr_b200
{

struct uvec2 { unsigned x = 0, y = 0; };

// ---- src/distribution_strategy.hh -------------------------------------------------------------
enum distribution_strategy
{
    DISTRIBUTION_DUPLICATE = 0,
    DISTRIBUTION_SCANLINE = 1,
    DISTRIBUTION_SHUFFLED_STRIPS = 2
};

struct distribution_params
{
    uvec2 size;
    distribution_strategy strategy = DISTRIBUTION_SCANLINE;
    unsigned index = 0;
    unsigned count = 1;
    bool primary = true;
};

inline uvec2 get_distribution_render_size(const distribution_params& params)
{
    switch(params.strategy)
    {
    case DISTRIBUTION_DUPLICATE: return params.size;
    case DISTRIBUTION_SCANLINE:
        return uvec2{params.size.x, (params.size.y - params.index + params.count - 1) / params.count};
    case DISTRIBUTION_SHUFFLED_STRIPS: return uvec2{params.count, 1};
    }
    return uvec2{};
}
inline uvec2 get_distribution_target_size(const distribution_params& params)
{
    if(params.primary) return params.size;
    if(params.strategy == DISTRIBUTION_SHUFFLED_STRIPS)
        return uvec2{params.size.x, (params.count + params.size.x - 1) / params.size.x};
    return get_distribution_render_size(params);
}
inline uvec2 get_ray_count(const distribution_params& params) { return get_distribution_render_size(params); }

inline unsigned calculate_shuffled_strips_b(uvec2 size)
{
    unsigned n = size.x * size.y;
    unsigned b = 31;
    while((n >> b) < 128 && b > 0) b--;
    return b;
}
inline size_t get_region_size(size_t image_size, unsigned b)
{
    size_t n_regions = size_t(1) << b;
    return (image_size + n_regions - 1) / n_regions;
}
inline unsigned calculate_shuffled_strips_pixels_per_device(uvec2 size, float max_ratio)
{
    unsigned b = calculate_shuffled_strips_b(size);
    return (unsigned)std::ceil(max_ratio * get_region_size(size_t(size.x) * size.y, b) * (1u << b));
}
inline distribution_params get_device_distribution_params(
    uvec2 full_image_size, distribution_strategy strategy, double workload_offset, double workload_size,
    unsigned device_index, unsigned device_count, bool primary)
{
    distribution_params d;
    d.strategy = strategy;
    d.size = full_image_size;
    d.primary = primary;
    if(strategy == DISTRIBUTION_SHUFFLED_STRIPS)
    {
        unsigned before = calculate_shuffled_strips_pixels_per_device(full_image_size, (float)workload_offset);
        unsigned after = calculate_shuffled_strips_pixels_per_device(full_image_size, (float)(workload_offset + workload_size));
        d.index = before;
        d.count = after - before;
    }
    else
    {
        d.index = device_index;
        d.count = device_count;
    }
    return d;
}

// ---- src/rt_common.hh ----------------------------------------------------------------------------
enum class film_filter { POINT = 0, BOX, BLACKMAN_HARRIS };
enum class multiple_importance_sampling_mode { MIS_DISABLED, MIS_BALANCE_HEURISTIC, MIS_POWER_HEURISTIC };
enum class bounce_sampling_mode { HEMISPHERE, COSINE_HEMISPHERE, MATERIAL };
enum class tri_light_sampling_mode { AREA, SOLID_ANGLE, HYBRID };
struct light_sampling_weights
{
    float point_lights = 1.0f;
    float directional_lights = 1.0f;
    float envmap = 1.0f;
    float emissive_triangles = 1.0f;
};

inline void check(int rc, tr_pt_ctx* ctx, const char* what)
{
    if(rc < 0) throw std::runtime_error(std::string(what) + ": " + tr_pt_last_error(ctx));
}

// ---- src/path_tracer_stage.hh (+ rt_camera_stage.hh, rt_stage.hh) --------------------------------
class path_tracer_stage
{
public:
    enum class sampler_type { UNIFORM_RANDOM = 0, SOBOL_OWEN, SOBOL_Z_ORDER_2D, SOBOL_Z_ORDER_3D };
    enum projection_type { PERSPECTIVE = 0, ORTHOGRAPHIC = 1 };

    struct options
    {
        // rt_stage::options (struct defaults of src/rt_stage.hh:29-40)
        int max_ray_depth = 8;
        float min_ray_dist = 0.001f;
        int rng_seed = 0;
        sampler_type local_sampler = sampler_type::UNIFORM_RANDOM;
        size_t max_passes_per_command_buffer = 0;   // accepted, meaningless without command buffers
        // rt_camera_stage::options (src/rt_camera_stage.hh:17-25)
        distribution_params distribution;
        size_t active_viewport_count = 1;
        int samples_per_pixel = 1;
        int samples_per_pass = 1;
        projection_type projection = PERSPECTIVE;
        bool transparent_background = false;
        // path_tracer_stage::options (src/path_tracer_stage.hh:13-30)
        bool use_shadow_terminator_fix = false;
        bool use_white_albedo_on_first_bounce = false;
        bool hide_lights = false;
        film_filter film = film_filter::BLACKMAN_HARRIS;
        multiple_importance_sampling_mode mis_mode = multiple_importance_sampling_mode::MIS_POWER_HEURISTIC;
        float film_radius = 1.0f;
        float russian_roulette_delta = 0;
        float indirect_clamping = 0;
        float regularization_gamma = 0.0f;
        bool depth_of_field = false;
        light_sampling_weights sampling_weights;
        bounce_sampling_mode bounce_mode = bounce_sampling_mode::MATERIAL;
        tri_light_sampling_mode tri_light_mode = tri_light_sampling_mode::HYBRID;
        // not a reference option (the reference has one stage class per rgen program): direct_stage sets it
        int estimator = TR_ESTIMATOR_PATH_TRACER;
    };

    // `scene` plays scene_stage's role: the flattened tables (src/scene_stage.cc:1066-1116,1277-1354);
    // `cameras` = the camera UBO (src/camera.cc:431-462).
    path_tracer_stage(int cuda_device, const tr_pt_scene_desc& scene, const tr_camera_pair* cameras,
                      uint32_t view_count, const options& opt)
    : opt(opt)
    {
        tr_pt_options o = to_abi(opt);
        check(tr_pt_create(&ctx, cuda_device, &o), nullptr, "tr_pt_create");
        try
        {
            check(tr_pt_set_scene(ctx, &scene), ctx, "tr_pt_set_scene");
            check(tr_pt_set_cameras(ctx, cameras, view_count), ctx, "tr_pt_set_cameras");
            reset_distribution_params(opt.distribution);
        }
        catch(...) { tr_pt_destroy(ctx); throw; }
    }
    path_tracer_stage(const path_tracer_stage&) = delete;
    path_tracer_stage& operator=(const path_tracer_stage&) = delete;
    ~path_tracer_stage() { tr_pt_destroy(ctx); }

    // stage::run (src/stage.cc:35-76): `stream` carries the ordering the reference expresses as dependencies.
    void run(void* cuda_stream = nullptr) { check(tr_pt_render(ctx, cuda_stream), ctx, "tr_pt_render"); }
    // two frames in flight (MAX_FRAMES_IN_FLIGHT = 2): run == run_begin + run_end; call run_begin for frame k + 1 before
    // run_end for frame k and the traversal work of both frames is queued on the device
    void run_begin(void* cuda_stream = nullptr) { check(tr_pt_render_begin(ctx, cuda_stream), ctx, "tr_pt_render_begin"); }
    void run_end(void* cuda_stream = nullptr) { check(tr_pt_render_end(ctx, cuda_stream), ctx, "tr_pt_render_end"); }
    void sync() { check(tr_pt_sync(ctx), ctx, "tr_pt_sync"); }
    void reset_accumulated_samples() { check(tr_pt_reset_accumulation(ctx, 0), ctx, "tr_pt_reset_accumulation"); }
    void reset_sample_counter() { check(tr_pt_reset_accumulation(ctx, 1), ctx, "tr_pt_reset_accumulation"); }
    void reset_distribution_params(distribution_params distribution)
    {
        opt.distribution = distribution;
        tr_pt_distribution d;
        d.size[0] = distribution.size.x; d.size[1] = distribution.size.y;
        d.strategy = (int)distribution.strategy;
        d.index = distribution.index; d.count = distribution.count;
        d.primary = distribution.primary ? 1 : 0;
        check(tr_pt_set_distribution(ctx, &d), ctx, "tr_pt_set_distribution");
    }
    void set_cameras(const tr_camera_pair* cameras, uint32_t view_count)
    {
        check(tr_pt_set_cameras(ctx, cameras, view_count), ctx, "tr_pt_set_cameras");
    }
    // the per-frame part of scene_stage::update for animated scenes (src/scene_stage.cc:1572-1675):
    // new instance records -> world-space vertices, tri lights and the acceleration structure, on the device
    void update_instances(const tr_instance* instances, uint32_t instance_count)
    {
        check(tr_pt_update_instances(ctx, instances, instance_count), ctx, "tr_pt_update_instances");
    }
    // TR_UPDATE_REBUILD (default) or TR_UPDATE_REFIT: how update_instances brings the acceleration structure up to date
    void set_update_mode(int mode) { check(tr_pt_set_update_mode(ctx, mode), ctx, "tr_pt_set_update_mode"); }
    std::vector<float> read_color()
    {
        void* p; size_t bytes;
        check(tr_pt_get_color_buffer(ctx, &p, &bytes), ctx, "tr_pt_get_color_buffer");
        std::vector<float> out(bytes / sizeof(float));
        check(tr_pt_read_color(ctx, out.data(), bytes), ctx, "tr_pt_read_color");
        return out;
    }
    // frames in flight: the staging copy of the frame rendered last, on the caller's copy stream (pinned host memory)
    void read_color_async(float* pinned_host, size_t bytes, void* copy_stream)
    {
        check(tr_pt_read_color_async(ctx, pinned_host, bytes, copy_stream), ctx, "tr_pt_read_color_async");
    }
    // accumulation checkpoint / resume (not in the reference)
    tr_pt_accumulation_state accumulation_state()
    {
        tr_pt_accumulation_state st;
        check(tr_pt_get_accumulation_state(ctx, &st), ctx, "tr_pt_get_accumulation_state");
        return st;
    }
    void restore_accumulation(const std::vector<float>& color, const tr_pt_accumulation_state& st)
    {
        check(tr_pt_restore_accumulation(ctx, color.data(), color.size() * sizeof(float), &st), ctx, "tr_pt_restore_accumulation");
    }
    // aux G-buffer layers are part of a checkpoint: dump with read_gbuffer_layer, write back after restore_accumulation
    std::vector<unsigned char> read_gbuffer_layer(int layer)
    {
        void* p; size_t bytes;
        check(tr_pt_get_gbuffer_layer(ctx, layer, &p, &bytes), ctx, "tr_pt_get_gbuffer_layer");
        std::vector<unsigned char> out(bytes);
        check(tr_pt_read_gbuffer_layer(ctx, layer, out.data(), bytes), ctx, "tr_pt_read_gbuffer_layer");
        return out;
    }
    void write_gbuffer_layer(int layer, const std::vector<unsigned char>& texels)
    {
        check(tr_pt_write_gbuffer_layer(ctx, layer, texels.data(), texels.size()), ctx, "tr_pt_write_gbuffer_layer");
    }
    tr_pt_stats get_stats() { tr_pt_stats s; check(tr_pt_get_stats(ctx, &s), ctx, "tr_pt_get_stats"); return s; }
    tr_pt_ctx* handle() { return ctx; }
    const options& get_options() const { return opt; }

    static tr_pt_options to_abi(const options& opt)
    {
        tr_pt_options o;
        tr_pt_default_options(&o);
        o.max_ray_depth = opt.max_ray_depth;
        o.min_ray_dist = opt.min_ray_dist;
        o.rng_seed = opt.rng_seed;
        o.local_sampler = (int)opt.local_sampler;
        o.samples_per_pixel = opt.samples_per_pixel;
        o.samples_per_pass = opt.samples_per_pass;
        o.projection = (int)opt.projection;
        o.transparent_background = opt.transparent_background;
        o.active_viewport_count = (uint32_t)opt.active_viewport_count;
        o.use_shadow_terminator_fix = opt.use_shadow_terminator_fix;
        o.use_white_albedo_on_first_bounce = opt.use_white_albedo_on_first_bounce;
        o.hide_lights = opt.hide_lights;
        o.film = (int)opt.film;
        o.mis_mode = (int)opt.mis_mode;
        o.film_radius = opt.film_radius;
        o.russian_roulette_delta = opt.russian_roulette_delta;
        o.indirect_clamping = opt.indirect_clamping;
        o.regularization_gamma = opt.regularization_gamma;
        o.depth_of_field = opt.depth_of_field;
        o.sampling_weight_point_lights = opt.sampling_weights.point_lights;
        o.sampling_weight_directional_lights = opt.sampling_weights.directional_lights;
        o.sampling_weight_envmap = opt.sampling_weights.envmap;
        o.sampling_weight_emissive_triangles = opt.sampling_weights.emissive_triangles;
        o.bounce_mode = (int)opt.bounce_mode;
        o.tri_light_mode = (int)opt.tri_light_mode;
        o.estimator = opt.estimator;
        return o;
    }

private:
    tr_pt_ctx* ctx = nullptr;
    options opt;
};

// ---- src/direct_stage.hh: the direct-light renderer (shader/direct.rgen) -----------------------------
// Same stage surface as path_tracer_stage; the options are direct_stage::options (src/direct_stage.hh:13-22):
// film, film_radius, sampling_weights, bounce_mode, tri_light_mode on top of rt_camera_stage::options.
class direct_stage: public path_tracer_stage
{
public:
    struct options
    {
        int max_ray_depth = 8;
        float min_ray_dist = 0.001f;
        int rng_seed = 0;
        sampler_type local_sampler = sampler_type::UNIFORM_RANDOM;
        distribution_params distribution;
        size_t active_viewport_count = 1;
        int samples_per_pixel = 1;
        int samples_per_pass = 1;
        projection_type projection = PERSPECTIVE;
        bool transparent_background = false;
        film_filter film = film_filter::BLACKMAN_HARRIS;
        float film_radius = 1.0f;
        light_sampling_weights sampling_weights;
        bounce_sampling_mode bounce_mode = bounce_sampling_mode::MATERIAL;
        tri_light_sampling_mode tri_light_mode = tri_light_sampling_mode::HYBRID;
    };
    direct_stage(int cuda_device, const tr_pt_scene_desc& scene, const tr_camera_pair* cameras, uint32_t view_count,
                 const options& opt)
    : path_tracer_stage(cuda_device, scene, cameras, view_count, widen(opt)) {}

    static path_tracer_stage::options widen(const options& d)
    {
        path_tracer_stage::options o;
        o.max_ray_depth = d.max_ray_depth; o.min_ray_dist = d.min_ray_dist; o.rng_seed = d.rng_seed;
        o.local_sampler = d.local_sampler; o.distribution = d.distribution;
        o.active_viewport_count = d.active_viewport_count;
        o.samples_per_pixel = d.samples_per_pixel; o.samples_per_pass = d.samples_per_pass;
        o.projection = d.projection; o.transparent_background = d.transparent_background;
        o.film = d.film; o.film_radius = d.film_radius; o.sampling_weights = d.sampling_weights;
        o.bounce_mode = d.bounce_mode; o.tri_light_mode = d.tri_light_mode;
        o.estimator = TR_ESTIMATOR_DIRECT;
        return o;
    }
};

// ---- src/load_balancer.{hh,cc} -------------------------------------------------------------------
class load_balancer
{
public:
    explicit load_balancer(size_t device_count, const std::vector<double>& initial_weights = {})
    : workloads(initial_weights)
    {
        workloads.resize(device_count);
        double sum = 0, add = 0;
        for(double w: workloads) sum += w;
        if(sum == 0) { add = 1.0; sum = (double)workloads.size(); }
        for(double& w: workloads) w = (std::max(w, 0.0) + add) / sum;
    }
    // `times` = the per-device "path tracing" timer (here tr_pt_stats::render_ms)
    const std::vector<double>& update(const std::vector<double>& times)
    {
        double sum_speed = 0;
        for(size_t i = 0; i < workloads.size(); ++i) sum_speed += std::max(workloads[i] / times[i], 0.0);
        if(sum_speed > 0 && std::isfinite(sum_speed))
            for(size_t i = 0; i < workloads.size(); ++i)
                workloads[i] = workloads[i] * 0.9 + (workloads[i] / times[i] / sum_speed) * 0.1;
        return workloads;
    }
    std::vector<double> workloads;
};

// ---- src/rt_renderer.{hh,cc} for Pipeline = path_tracer_stage, one process driving N GPUs --------
class rt_renderer
{
public:
    struct options: path_tracer_stage::options
    {
        distribution_strategy strategy = DISTRIBUTION_SHUFFLED_STRIPS;
        bool accumulate = false;    // rt_renderer::options::accumulate
    };

    rt_renderer(const std::vector<int>& cuda_devices, uvec2 size, const options& opt)
    : devices(cuda_devices), size(size), opt(opt), workloads(cuda_devices.size(), 1.0 / cuda_devices.size()) {}

    // scene_stage::update on every device (the scene is replicated per device like Tauray's per_device<> buffers)
    void update_instances(const tr_instance* instances, uint32_t instance_count)
    {
        for(auto& s: stages) s->update_instances(instances, instance_count);
    }
    // acceleration-structure update of animated scenes: TR_UPDATE_REFIT (the reference's eUpdate builds) or TR_UPDATE_REBUILD
    void set_update_mode(int mode)
    {
        for(auto& s: stages) s->set_update_mode(mode);
    }

    // renderer::set_scene (src/rt_renderer.cc:60-65): builds one stage per device
    void set_scene(const tr_pt_scene_desc& scene, const tr_camera_pair* cameras, uint32_t view_count)
    {
        stages.clear();
        views = view_count;
        for(size_t i = 0; i < devices.size(); ++i)
        {
            path_tracer_stage::options o = opt;
            o.distribution = device_params(i);
            stages.emplace_back(new path_tracer_stage(devices[i], scene, cameras, view_count, o));
        }
        connect_display_image();
        accumulated_frames = 0;
    }
    // The other gbuffer_target entries (src/gbuffer.hh:18-51): rendered by every stage and gathered into the display
    // device's image like the colour (src/rt_renderer.cc:370-397). `layer_mask`: bit (1 << TR_GBUFFER_x) per layer.
    void set_gbuffer_layers(uint32_t layer_mask)
    {
        for(auto& s: stages) check(tr_pt_set_gbuffer_layers(s->handle(), layer_mask), s->handle(), "tr_pt_set_gbuffer_layers");
        connect_display_image();
    }
    // one stitched aux layer on the display device, in the layer's texel format
    std::vector<unsigned char> read_layer(int layer)
    {
        if(stages.size() == 1) return stages[0]->read_gbuffer_layer(layer);
        tr_pt_full_image_info info;
        check(tr_pt_get_full_image_info(stages[0]->handle(), &info), stages[0]->handle(), "tr_pt_get_full_image_info");
        static const size_t texel_bytes[TR_GBUFFER_LAYER_COUNT] = {16, 8, 8, 8, 4, 4, 16, 16, 4};
        std::vector<unsigned char> out((size_t)info.pixels * texel_bytes[layer]);
        check(tr_pt_read_full_layer(stages[0]->handle(), layer, out.data(), out.size()), stages[0]->handle(), "tr_pt_read_full_layer");
        return out;
    }
    // renderer::reset_accumulation (src/rt_renderer.cc:67-81)
    void reset_accumulation(bool reset_sample_counter = true)
    {
        for(auto& s: stages)
        {
            s->reset_accumulated_samples();
            if(reset_sample_counter) s->reset_sample_counter();
        }
        accumulated_frames = 0;
        set_blend_ratio(1.0f);
    }
    // renderer::render (src/rt_renderer.cc:84-133)
    void render()
    {
        // every stage touches only the pixels of its own strip, in the display image too: no ordering between stages
        for(auto& s: stages) s->run();
        for(auto& s: stages) s->sync();
        if(blend_ratio < 1.0f) set_blend_ratio(1.0f);   // "Reset temporary blending from stitching", :126-127
        accumulated_frames++;
    }
    // renderer::set_device_workloads (src/rt_renderer.cc:136-183)
    void set_device_workloads(const std::vector<double>& ratios)
    {
        if(opt.strategy != DISTRIBUTION_SHUFFLED_STRIPS) return;
        double cumulative = 0;
        workloads.resize(ratios.size());
        for(size_t i = 0; i < ratios.size(); ++i)
        {
            workloads[i] = std::min(std::max(ratios[i], 0.0), 1.0 - cumulative);
            cumulative += workloads[i];
        }
        for(size_t i = 0; i < stages.size(); ++i)
        {
            stages[i]->reset_distribution_params(device_params(i));
            // only the primary device renders in place and can keep accumulating (:165-175)
            if(i != 0 && stages.size() > 1) stages[i]->reset_accumulated_samples();
        }
        if(opt.accumulate && stages.size() > 1) set_blend_ratio(1.0f / (float)(accumulated_frames + 1));
    }
    // the stitched image on the display device (headless staging copy, src/headless.cc:125-171)
    std::vector<float> read_image()
    {
        if(stages.size() == 1) return stages[0]->read_color();
        std::vector<float> out((size_t)size.x * size.y * views * 4);
        check(tr_pt_read_full_image(stages[0]->handle(), out.data(), out.size() * sizeof(float)), stages[0]->handle(), "tr_pt_read_full_image");
        return out;
    }
    std::vector<double> device_times_ms()
    {
        std::vector<double> t;
        for(auto& s: stages) t.push_back(s->get_stats().render_ms);
        return t;
    }
    path_tracer_stage& stage(size_t i) { return *stages[i]; }
    size_t device_count() const { return devices.size(); }

private:
    void connect_display_image()
    {
        if(devices.size() <= 1) return;
        // display device = device 0 (src/context.cc): it renders in place into the display image
        // (primary distribution); every other stage stores its finished pixels (colour and aux layers) straight
        // into that image over NVLink (replaces device_transfer + stitch_stage)
        void* full = nullptr; size_t bytes = 0;
        check(tr_pt_alloc_full_image(stages[0]->handle(), &full, &bytes), stages[0]->handle(), "tr_pt_alloc_full_image");
        check(tr_pt_set_render_in_full_image(stages[0]->handle(), 1), stages[0]->handle(), "tr_pt_set_render_in_full_image");
        for(size_t i = 1; i < devices.size(); ++i)
        {
            check(tr_pt_enable_peer_access(stages[i]->handle(), devices[0]), stages[i]->handle(), "tr_pt_enable_peer_access");
            void* peers[1] = {full};
            check(tr_pt_set_peer_images(stages[i]->handle(), peers, 1), stages[i]->handle(), "tr_pt_set_peer_images");
        }
    }
    distribution_params device_params(size_t i) const
    {
        if(devices.size() == 1)
        {
            distribution_params d;
            d.size = size; d.strategy = opt.strategy; d.primary = true; d.index = 0;
            d.count = opt.strategy == DISTRIBUTION_SHUFFLED_STRIPS ? calculate_shuffled_strips_pixels_per_device(size, 1.0f) : 1;
            return d;
        }
        double offset = 0;
        for(size_t k = 0; k < i; ++k) offset += workloads[k];
        return get_device_distribution_params(size, opt.strategy, offset, workloads[i], (unsigned)i, (unsigned)devices.size(), i == 0);
    }
    // stitch_stage::set_blend_ratio for the fused stitch of the non-primary stages
    void set_blend_ratio(float r)
    {
        blend_ratio = r;
        for(size_t i = 1; i < stages.size(); ++i)
            check(tr_pt_set_stitch_blend(stages[i]->handle(), r), stages[i]->handle(), "tr_pt_set_stitch_blend");
    }

    std::vector<int> devices;
    uvec2 size;
    options opt;
    std::vector<double> workloads;
    std::vector<std::unique_ptr<path_tracer_stage>> stages;
    uint32_t views = 0;
    uint32_t accumulated_frames = 0;
    float blend_ratio = 1.0f;
};

} // namespace tr_b200
