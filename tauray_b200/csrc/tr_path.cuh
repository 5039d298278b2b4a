// tr_path.cuh -- the per-path logic of Tauray's path tracer, split at the traversal calls so the
// same device functions serve the wavefront kernels (state in HBM between stages) and the
// single-launch megakernel (state in registers):
//
//   path_begin      = path_tracer.rgen:90-100 (sampler init + get_world_camera_ray)
//   closest_hit     = traceRayEXT (closest) incl. any-hit alpha test and sphere lights
//   shade_bounce    = evaluate_ray loop body, path_tracer.glsl:405-497, up to the NEE shadow ray
//   shadow_visibility = shadow_ray, path_tracer.glsl:35-52
//   nee_finish      = path_tracer.glsl:336-340 + :453-469 (after the shadow ray returned)
//   path_color      = path_tracer.rgen:112
//   accumulate      = write_all_outputs + accumulate_gbuffer_color (gbuffer.glsl:18-28)
#pragma once
#include "tr_shading.cuh"
#include "tr_trace.cuh"

namespace trd {

struct RenderParams
{
    DeviceScene scene;
    tr_pt_options opt;
    FrameParams frame;
    const tr_camera_pair* cameras;
    float4* color;              // colour target of this device
    const float4* color_prev;   // where the first pass of a frame reads the accumulated value (== color unless the target
                                // is the double-buffered full image: then the other half of the ping-pong pair)
    Counters* counters;
    const uint4* sobol_chunks;  // g_sobol_chunks (global memory)
    int stats;                  // count node visits / triangle tests
    // fused stitch: full images of the peer GPUs (NVLink-mapped), written by the last pass
    // Every peer pointer is the base of that rank's full-image arena: two halves (ping-pong per frame) of
    // [colour | enabled aux layers]; full_cur / full_prev = byte offset of the half this frame writes / the previous
    // frame wrote, full_layer[l] = byte offset of layer l inside a half (colour = 0).
    float4* peers[8];
    int peer_count;
    int peer_self;              // index of this rank's own arena in peers[], -1 if it is not among them
    int last_pass;
    float stitch_blend;         // stitch_stage blend_ratio; 1 = plain store
    size_t full_cur, full_prev;
    size_t full_layer[TR_GBUFFER_LAYER_COUNT];
    // aux G-buffer layers (NULL when absent), texel formats of src/gbuffer.hh:18-51
    void* gbuffer[TR_GBUFFER_LAYER_COUNT];
    const void* gbuffer_prev[TR_GBUFFER_LAYER_COUNT];   // see color_prev
    uint32_t gbuffer_mask;
    const float* prev_from_world;   // 16 floats per instance: model_prev * inverse(model), column-major
    // traversal tuning (defaults in tr_render.cu; TR_TUNE_* environment overrides for experiments)
    int tune_fetch_threshold;
    int tune_postpone_div;
    int tune_tma;               // bit 0: Sobol tables of the shade kernel via TMA, bit 1: top-of-tree nodes in shared memory
    int tune_coop_groups;       // k_wf_trace_coop: posted triangle groups per warp that start a triangle phase
    int tune_coop_wait;         // ... or this many lanes out of nodes and waiting for their triangles
    int tune_shade_prefetch;    // k_wf_shade: prefetch the next path's records (0 off, 1 L2, 2 L1)
    int tune_split_stage;       // k_wf_trace<..., SPLIT>: run the split stage
    int tune_split_steps;       // ... and split a ray early once it has taken this many node steps (<= 255)
    unsigned long long tune_probe_bin_ns;
    int probe_slot;             // TR_DEBUG_DRAIN: index of this launch in the frame's timeline
    unsigned long long* probe;  // TR_DEBUG_DRAIN: per launch {first warp start, first / last warp that found the queue empty, last exit} (ns)
};

// Path state carried from bounce to bounce (evaluate_ray's locals).
struct PathState
{
    vec3 pos;
    vec3 view;
    vec3 attenuation;
    float bsdf_pdf;
    float regularization;
    bsdf_lobes primary_lobes;
    vec3 diffuse;
    vec3 reflection;
    float aux_alpha;            // diffuse.a = reflection.a (1 / distance to the second hit)
    uint32_t random_seed;       // payload.random_seed
    // first hit (first_hit_material.{emission, albedo, metallic})
    vec3 fh_emission;
    vec4 fh_albedo;
    float fh_metallic;
};

// A next-event-estimation sample waiting for its shadow ray.
struct NeeJob
{
    vec3 dir;
    float tmax;
    vec3 contrib;
    float mis_pdf;              // nee_mis_pdf(light_pdf, bsdf_pdf)
    vec3 attenuation;
    vec3 modulate;              // modulate_bsdf(mat, lobes), bounce != 0 only
    float w_diffuse;            // demodulation weights of the lobes in effect for this sample
    float w_reflection;
    int needs_shadow_ray;
    int bounce;
};

// ---------------------------------------------------------------------------------------------
// pixel mapping: shader/rt.glsl:170-222
// ---------------------------------------------------------------------------------------------
TRD uint32_t permute_region_id(const FrameParams& f, uint32_t i)
{
    uint32_t region_id = i / f.region_size;
    uint32_t k = __brev(region_id) >> (32 - f.strips_b);
    return k * f.region_size + i % f.region_size;
}

// slot = launch id flattened as (view, y, x). Returns false when the launch id renders nothing.
TRD bool map_slot(const FrameParams& f, uint32_t slot, uint32_t& px, uint32_t& py, uint32_t& view, uint32_t& write_index)
{
    const uint32_t per_view = f.ray_count_x * f.ray_count_y;
    view = slot / per_view;
    const uint32_t r = slot - view * per_view;
    const uint32_t lx = r % f.ray_count_x, ly = r / f.ray_count_x;
    if(f.strategy == TR_DISTRIBUTION_SHUFFLED_STRIPS)
    {
        const uint32_t j = permute_region_id(f, f.index + lx);
        if(j >= f.size_x * f.size_y) return false;
        px = j % f.size_x;
        py = j / f.size_x;
        write_index = f.primary ? j : lx;
    }
    else if(f.strategy == TR_DISTRIBUTION_SCANLINE)
    {
        px = lx;
        py = ly * f.count + f.index;
        write_index = f.primary ? py * f.size_x + px : ly * f.size_x + lx;
    }
    else
    {
        px = lx; py = ly;
        write_index = ly * f.size_x + lx;
    }
    write_index += view * f.layer_size;
    return true;
}

// ---------------------------------------------------------------------------------------------
// textures: bilinear, repeat, LOD 0
// ---------------------------------------------------------------------------------------------
TRD vec4 texel(const DeviceTexture& t, int x, int y)
{
    size_t i = (size_t)y * t.width + x;
    if(t.format == TR_TEXTURE_RGBA32F)
    {
        float4 f = __ldg(reinterpret_cast<const float4*>(t.data) + i);
        return V4(f.x, f.y, f.z, f.w);
    }
    if(t.format == TR_TEXTURE_RGBA16F)
    {
        uint2 h = __ldg(reinterpret_cast<const uint2*>(t.data) + i);
        return V4(half_to_float_(h.x & 0xFFFFu), half_to_float_(h.x >> 16), half_to_float_(h.y & 0xFFFFu), half_to_float_(h.y >> 16));
    }
    uchar4 b = __ldg(reinterpret_cast<const uchar4*>(t.data) + i);
    return V4((float)b.x / 255.0f, (float)b.y / 255.0f, (float)b.z / 255.0f, (float)b.w / 255.0f);
}
TRD vec4 sample_texture(const DeviceScene& s, int id, vec2 uv)
{
    const DeviceTexture t = s.textures[id];
    float x = uv.x * (float)t.width - 0.5f;
    float y = uv.y * (float)t.height - 0.5f;
    if(!(fabsf(x) < 1e9f)) x = 0.0f;
    if(!(fabsf(y) < 1e9f)) y = 0.0f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float fx = x - fx0, fy = y - fy0;
    int w = (int)t.width, h = (int)t.height;
    int x0 = (int)fx0 % w; if(x0 < 0) x0 += w;
    int y0 = (int)fy0 % h; if(y0 < 0) y0 += h;
    int x1 = x0 + 1 == w ? 0 : x0 + 1;
    int y1 = y0 + 1 == h ? 0 : y0 + 1;
    vec4 top = mix_(texel(t, x0, y0), texel(t, x1, y0), fx);
    vec4 bot = mix_(texel(t, x0, y1), texel(t, x1, y1), fx);
    return mix_(top, bot, fy);
}

// envmap_sampler (src/scene_stage.cc:259-265): linear, repeat in x, clamp-to-edge in y, LOD 0
TRD vec4 sample_envmap_texture(const DeviceScene& s, vec2 uv)
{
    const DeviceTexture t = s.envmap;
    float x = uv.x * (float)t.width - 0.5f;
    float y = uv.y * (float)t.height - 0.5f;
    if(!(fabsf(x) < 1e9f)) x = 0.0f;
    if(!(fabsf(y) < 1e9f)) y = 0.0f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float fx = x - fx0, fy = y - fy0;
    int w = (int)t.width, h = (int)t.height;
    int x0 = (int)fx0 % w; if(x0 < 0) x0 += w;
    int x1 = x0 + 1 == w ? 0 : x0 + 1;
    int y0 = min(max((int)fy0, 0), h - 1);
    int y1 = min(max((int)fy0 + 1, 0), h - 1);
    vec4 top = mix_(texel(t, x0, y0), texel(t, x1, y0), fx);
    vec4 bot = mix_(texel(t, x0, y1), texel(t, x1, y1), fx);
    return mix_(top, bot, fy);
}

// ---------------------------------------------------------------------------------------------
// environment map importance sampling: shader/alias_table.glsl:23-37, shader/rt.glsl:251-300
// ---------------------------------------------------------------------------------------------
// ids outside the table read as zero (Vulkan robust buffer access): the rounding in
// latlong_direction_to_pixel_id can produce row == height.
TRD tr_alias_table_entry alias_entry(const DeviceScene& s, int i)
{
    tr_alias_table_entry e;
    e.alias_id = 0u; e.probability = 0u; e.pdf = 0.0f; e.alias_pdf = 0.0f;
    if(i >= 0 && (uint32_t)i < s.envmap.width * s.envmap.height)
    {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(s.alias_table) + i);
        e.alias_id = q.x; e.probability = q.y; e.pdf = __uint_as_float(q.z); e.alias_pdf = __uint_as_float(q.w);
    }
    return e;
}

TRD int latlong_direction_to_pixel_id(vec3 dir, int sx, int sy)
{
    const float u = atan2_(dir.z, dir.x) * 0.5f / TR_PI + 0.5f;
    const float v = asin_(-dir.y) / TR_PI + 0.5f;
    const int px = (int)(u * (float)sx + 0.5f);
    const int py = (int)(v * (float)sy + 0.5f);
    return px + py * sx;
}

TRD vec3 uv_to_latlong_direction(vec2 uv)
{
    const float ux = (uv.x - 0.5f) * TR_PI, uy = (uv.y - 0.5f) * TR_PI;
    vec3 dir = V3(cos_(2.0f * ux), -sin_(uy), sin_(2.0f * ux));
    const float r = sqrtf(1.0f - dir.y * dir.y);
    dir.x *= r; dir.z *= r;
    return dir;
}

TRD vec3 sample_environment_map(const DeviceScene& s, uvec4 rand, vec3& dir, float& length, float& pdf)
{
    vec3 color = V3(s.environment_factor[0], s.environment_factor[1], s.environment_factor[2]);
    const uint32_t sx = s.envmap.width, sy = s.envmap.height;
    const uint32_t pixel_count = sx * sy;
    const uint32_t ipx = min(rand.x / (0xFFFFFFFFu / sx), sx - 1u);
    const uint32_t ipy = min(rand.y / (0xFFFFFFFFu / sy), sy - 1u);
    int i = (int)(ipx + ipy * sx);
    const tr_alias_table_entry at = alias_entry(s, i);
    pdf = at.pdf;
    if(rand.z > at.probability)
    {
        i = (int)at.alias_id;
        pdf = at.alias_pdf;
    }
    const uint32_t px = (uint32_t)i % sx, py = (uint32_t)i / sx;
    const float offx = (float)(rand.x * pixel_count) * TR_INV_UINT32_MAX;
    const float offy = (float)(rand.y * pixel_count) * TR_INV_UINT32_MAX;
    const vec2 uv = V2(((float)(int)px + offx) / (float)sx, ((float)(int)py + offy) / (float)sy);
    dir = uv_to_latlong_direction(uv);
    color = color * xyz(sample_envmap_texture(s, uv));
    length = inf_();   // RAY_MAX_DIST = float(1e39), shader/rt.glsl:25
    return color;
}

TRD float sample_environment_map_pdf(const DeviceScene& s, vec3 dir)
{
    return alias_entry(s, latlong_direction_to_pixel_id(dir, (int)s.envmap.width, (int)s.envmap.height)).pdf;
}

// ---------------------------------------------------------------------------------------------
// vertex fetch: shader/rt.glsl:28-117 (PRE_TRANSFORMED_VERTICES)
// ---------------------------------------------------------------------------------------------
struct TriVerts
{
    float4 a0, a1, a2;  // vertex 0: pos.xyz normal.x | normal.yz uv.xy | tangent
    float4 b0, b1, b2;
    float4 c0, c1, c2;
};

TRD void fetch_triangle(const DeviceScene& s, const DeviceInstance& inst, int primitive_id, bool want_tangent, TriVerts& tv)
{
    if(s.shade_tris && !want_tangent)
    {
        const float4* r = s.shade_tris + 6 * ((size_t)inst.triangle_base + (size_t)primitive_id);
        tv.a0 = __ldg(r); tv.a1 = __ldg(r + 1);
        tv.b0 = __ldg(r + 2); tv.b1 = __ldg(r + 3);
        tv.c0 = __ldg(r + 4); tv.c1 = __ldg(r + 5);
        return;
    }
    const uint32_t* idx = s.indices + inst.index_base + 3 * (size_t)primitive_id;
    const uint32_t i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
    const float4* v0 = reinterpret_cast<const float4*>(s.vertices + inst.vertex_base + i0);
    const float4* v1 = reinterpret_cast<const float4*>(s.vertices + inst.vertex_base + i1);
    const float4* v2 = reinterpret_cast<const float4*>(s.vertices + inst.vertex_base + i2);
    tv.a0 = __ldg(v0); tv.a1 = __ldg(v0 + 1);
    tv.b0 = __ldg(v1); tv.b1 = __ldg(v1 + 1);
    tv.c0 = __ldg(v2); tv.c1 = __ldg(v2 + 1);
    if(want_tangent) { tv.a2 = __ldg(v0 + 2); tv.b2 = __ldg(v1 + 2); tv.c2 = __ldg(v2 + 2); }
}

TRD vec2 interpolated_uv(const DeviceScene& s, const DeviceInstance& inst, int primitive_id, float bu, float bv)
{
    const uint32_t* idx = s.indices + inst.index_base + 3 * (size_t)primitive_id;
    const uint32_t i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(s.vertices + inst.vertex_base + i0) + 1);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(s.vertices + inst.vertex_base + i1) + 1);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(s.vertices + inst.vertex_base + i2) + 1);
    vec3 b = V3(1.0f - bu - bv, bu, bv);
    return V2(a1.z * b.x + b1.z * b.y + c1.z * b.z, a1.w * b.x + b1.w * b.y + c1.w * b.z);
}

// alpha of rt_common.rahit / rt_common_shadow.rahit (shader/rt.glsl:136-144)
TRD float hit_alpha(const DeviceScene& s, int instance_id, int primitive_id, float u, float v)
{
    const DeviceInstance& inst = s.instances[instance_id];
    float alpha = inst.mat.albedo_factor[3];
    if(inst.mat.albedo_tex_id >= 0)
        alpha *= sample_texture(s, inst.mat.albedo_tex_id, interpolated_uv(s, inst, primitive_id, u, v)).w;
    return alpha;
}

struct SurfaceVertex
{
    vec3 pos;
    vec3 hard_normal;
    vec3 smooth_normal;
    vec3 mapped_normal;
    vec2 uv;
    vec3 tangent, bitangent;
    bool back_facing;
};

// shader/rt.glsl:28-101; `origin` is the ray origin (for the tri-light pdf of emissive instances)
TRD SurfaceVertex get_interpolated_vertex(const RenderParams& P, const DeviceInstance& inst, vec3 view, float bu, float bv,
                                          int primitive_id, vec3 origin, float& pdf)
{
    TriVerts tv;
    const bool want_tangent = inst.mat.normal_tex_id >= 0;
    fetch_triangle(P.scene, inst, primitive_id, want_tangent, tv);
    const vec3 b = V3(1.0f - bu - bv, bu, bv);
    const vec3 p0 = V3(tv.a0.x, tv.a0.y, tv.a0.z), p1 = V3(tv.b0.x, tv.b0.y, tv.b0.z), p2 = V3(tv.c0.x, tv.c0.y, tv.c0.z);
    const vec3 n0 = V3(tv.a0.w, tv.a1.x, tv.a1.y), n1 = V3(tv.b0.w, tv.b1.x, tv.b1.y), n2 = V3(tv.c0.w, tv.c1.x, tv.c1.y);
    SurfaceVertex o;
    o.pos = p0 * b.x + p1 * b.y + p2 * b.z;
    pdf = 0.0f;
    if(P.opt.sampling_weight_emissive_triangles > 0 && inst.light_base_id >= 0)
        pdf = sample_triangle_light_pdf(P.opt.tri_light_mode, o.pos - origin, p0 - origin, p1 - origin, p2 - origin);
    o.smooth_normal = normalize(n0 * b.x + n1 * b.y + n2 * b.z);
    if(want_tangent)
    {
        vec3 t = V3(tv.a2.x, tv.a2.y, tv.a2.z) * b.x + V3(tv.b2.x, tv.b2.y, tv.b2.z) * b.y + V3(tv.c2.x, tv.c2.y, tv.c2.z) * b.z;
        float tw = tv.a2.w * b.x + tv.b2.w * b.y + tv.c2.w * b.z;
        o.tangent = normalize(t);
        o.bitangent = normalize(cross(o.smooth_normal, o.tangent) * tw);
    }
    else o.tangent = o.bitangent = V3(0);
    o.uv = V2(tv.a1.z * b.x + tv.b1.z * b.y + tv.c1.z * b.z, tv.a1.w * b.x + tv.b1.w * b.y + tv.c1.w * b.z);
    o.hard_normal = normalize(cross(p1 - p0, p2 - p0));
    o.back_facing = dot(o.hard_normal, view) > 0;
    if(o.back_facing)
    {
        o.smooth_normal = -o.smooth_normal;
        o.hard_normal = -o.hard_normal;
    }
    o.mapped_normal = o.smooth_normal;
    return o;
}

// shader/scene.glsl:88-152
TRD sampled_material sample_material(const RenderParams& P, const DeviceInstance& inst, SurfaceVertex& v)
{
    const tr_material& mat = inst.mat;
    sampled_material res;
    res.albedo = V4(mat.albedo_factor[0], mat.albedo_factor[1], mat.albedo_factor[2], mat.albedo_factor[3]);
    if(P.scene.has_textures && mat.albedo_tex_id >= 0)
    {
        vec4 tex_col = sample_texture(P.scene, mat.albedo_tex_id, v.uv);
        vec3 lin = inverse_srgb_correction(xyz(tex_col));
        res.albedo = res.albedo * V4(lin, tex_col.w);
    }
    vec2 mr = V2(mat.metallic_roughness_factor[0], mat.metallic_roughness_factor[1]);
    if(P.scene.has_textures && mat.metallic_roughness_tex_id >= 0)
    {
        vec4 t = sample_texture(P.scene, mat.metallic_roughness_tex_id, v.uv);
        mr = mr * V2(t.z, t.y);
    }
    res.metallic = mr.x;
    res.roughness = mr.y * mr.y;
    if(P.scene.has_textures && mat.normal_tex_id >= 0)
    {
        mat3 tbn;
        tbn.c0 = v.tangent; tbn.c1 = v.bitangent; tbn.c2 = v.smooth_normal;
        vec4 t = sample_texture(P.scene, mat.normal_tex_id, v.uv);
        vec3 ts_normal = normalize(V3(t.x * 2.0f - 1.0f, t.y * 2.0f - 1.0f, t.z * 2.0f - 1.0f));
        v.mapped_normal = normalize(mul(tbn, ts_normal * V3(mat.normal_factor, mat.normal_factor, 1.0f)));
        if(any_nan(v.mapped_normal)) v.mapped_normal = v.smooth_normal;
    }
    res.emission = V3(mat.emission_factor[0], mat.emission_factor[1], mat.emission_factor[2]);
    if(P.scene.has_textures && mat.emission_tex_id >= 0)
        res.emission = res.emission * xyz(sample_texture(P.scene, mat.emission_tex_id, v.uv));
    res.transmittance = mat.transmittance;
    if(v.back_facing && res.transmittance > 0.0001f)
    {
        res.ior_in = mat.ior;
        res.ior_out = 1.0f;
    }
    else
    {
        res.ior_in = 1.0f;
        res.ior_out = mat.ior;
    }
    float f0 = (res.ior_out - res.ior_in) / (res.ior_out + res.ior_in);
    f0 *= f0;
    res.f0 = f0;
    return res;
}

// ---------------------------------------------------------------------------------------------
// traceRayEXT (closest): rt_common.rchit/.rmiss/.rahit + rt_common_point_light.rint/.rchit.
// Result encoding = hit_payload (shader/rt_common_payload.glsl:4-21): instance >= 0 triangle;
// else primitive >= 0 sphere light with u = t; else miss. Any-hit invocations are emulated in
// increasing-t order: each non-opaque candidate draws pcg(random_seed).
// ---------------------------------------------------------------------------------------------
template<bool STATS>
TRD void closest_hit(const RenderParams& P, vec3 pos, float tmin, vec3 dir, bool include_lights,
                     uint32_t& random_seed, Hit& out, TraceStats& st)
{
    float cur_tmin = tmin;
    for(;;)
    {
        Hit h;
        h.instance_id = -1; h.primitive_id = -1; h.u = 0; h.v = 0; h.t = 0; h.flags = 0;
        const bool tri_hit = traverse<false, STATS>(P.scene.bvh, pos, dir, cur_tmin, inf_(), h, st,
                                                    [](int, int, float, float, float) {});
        int light_id = -1;
        float light_t = 0.0f;
        if(include_lights)
        {
            for(uint32_t i = 0; i < P.scene.point_light_count; ++i)
            {
                const tr_point_light& pl = P.scene.point_lights[i];
                if(pl.radius == 0.0f) continue;
                float t = sphere_intersection(ld3(pl.pos), pl.radius, pos, dir);
                if(!(t > 0)) continue;
                if(!(t >= cur_tmin)) continue;
                if(light_id < 0 || t < light_t) { light_id = (int)i; light_t = t; }
            }
        }
        if(light_id >= 0 && (!tri_hit || light_t < h.t))
        {
            out.instance_id = -1; out.primitive_id = light_id; out.u = light_t; out.v = 0; out.t = light_t; out.flags = 0;
            return;
        }
        if(!tri_hit)
        {
            out.instance_id = -1; out.primitive_id = -1; out.u = 0; out.v = 0; out.t = 0; out.flags = 0;
            return;
        }
        if(!(h.flags & TR_TRI_FLAG_OPAQUE))
        {
            float alpha_cutoff = (float)pcg(random_seed) * TR_INV_UINT32_MAX;
            if(hit_alpha(P.scene, h.instance_id, h.primitive_id, h.u, h.v) <= alpha_cutoff)
            {
                cur_tmin = h.t;
                continue;
            }
        }
        out = h;
        return;
    }
}

// shader/path_tracer.glsl:35-52 + rt_common_shadow.{rchit,rahit,rmiss}; light spheres excluded
// (mask 0xFD). Non-opaque factors are multiplied in (t, instance, primitive) order.
template<bool STATS>
TRD float shadow_visibility(const RenderParams& P, vec3 pos, float min_dist, vec3 dir, float max_dist, TraceStats& st)
{
    constexpr int MAXF = 6;
    float ft[MAXF], ff[MAXF];
    int fi[MAXF], fp[MAXF];
    int nf = 0;
    float overflow = 1.0f;
    Hit h;
    const DeviceScene& s = P.scene;
    const bool blocked = traverse<true, STATS>(s.bvh, pos, dir, min_dist, max_dist, h, st,
        [&](int inst, int prim, float u, float v, float t) {
            float f = 1.0f - hit_alpha(s, inst, prim, u, v);
            if(nf < MAXF) { ft[nf] = t; ff[nf] = f; fi[nf] = inst; fp[nf] = prim; nf++; }
            else overflow *= f;
        });
    if(blocked) return 0.0f;
    float visibility = 1.0f;
    // selection in (t, instance, primitive) order
    for(int k = 0; k < nf; ++k)
    {
        int m = -1;
        for(int j = 0; j < nf; ++j)
        {
            if(fi[j] == -2) continue;
            if(m < 0 || ft[j] < ft[m] || (ft[j] == ft[m] && (fi[j] < fi[m] || (fi[j] == fi[m] && fp[j] < fp[m])))) m = j;
        }
        visibility *= ff[m];
        fi[m] = -2;
        if(visibility == 0.0f) return 0.0f;
    }
    return visibility * overflow;
}

// ---------------------------------------------------------------------------------------------
// NEE helpers: shader/rt.glsl:302-335, shader/path_tracer.glsl:54-89, 203-300, 356-365
// ---------------------------------------------------------------------------------------------
struct NeeProbs { float point, triangle, directional, envmap; };

TRD NeeProbs get_nee_sampling_probabilities(const RenderParams& P)
{
    NeeProbs p;
    p.point = (P.opt.sampling_weight_point_lights > 0 && P.scene.point_light_count > 0) ? P.opt.sampling_weight_point_lights : 0.0f;
    p.triangle = (P.opt.sampling_weight_emissive_triangles > 0 && P.scene.tri_light_count > 0) ? P.opt.sampling_weight_emissive_triangles : 0.0f;
    p.directional = (P.opt.sampling_weight_directional_lights > 0 && P.scene.directional_light_count > 0) ? P.opt.sampling_weight_directional_lights : 0.0f;
    p.envmap = (P.opt.sampling_weight_envmap > 0 && P.scene.environment_proj >= 0) ? P.opt.sampling_weight_envmap : 0.0f;
    float sum = p.point + p.triangle + p.directional + p.envmap;
    float inv_sum = sum <= 0.0f ? 0.0f : (1.0f / sum + 1e-5f);
    p.point *= inv_sum;
    p.triangle *= inv_sum;
    p.directional *= inv_sum;
    p.envmap *= inv_sum;
    return p;
}

struct IntersectionPdf { float point_light_pdf, directional_light_pdf, tri_light_pdf, envmap_pdf; };

TRD uint32_t max1(uint32_t c) { return c > 1u ? c : 1u; }

TRD float bsdf_mis_pdf(const RenderParams& P, const IntersectionPdf& nee_pdf, float bsdf_pdf)
{
    if(bsdf_pdf == 0.0f) return 1.0f;
    NeeProbs pr = get_nee_sampling_probabilities(P);
    float avg_nee_pdf =
        nee_pdf.directional_light_pdf * pr.directional / (float)max1(P.scene.directional_light_count) +
        nee_pdf.tri_light_pdf * pr.triangle / (float)max1(P.scene.tri_light_count) +
        nee_pdf.envmap_pdf * pr.envmap +
        nee_pdf.point_light_pdf * pr.point / (float)max1(P.scene.point_light_count);
    if(P.opt.mis_mode == TR_MIS_POWER_HEURISTIC)
        return (avg_nee_pdf * avg_nee_pdf + bsdf_pdf * bsdf_pdf) / bsdf_pdf;
    else if(P.opt.mis_mode == TR_MIS_BALANCE_HEURISTIC)
        return avg_nee_pdf + bsdf_pdf;
    return avg_nee_pdf > 0 ? inf_() : bsdf_pdf;
}

TRD float nee_mis_pdf(const RenderParams& P, float nee_pdf, float bsdf_pdf)
{
    if(nee_pdf <= 0.0f) return -nee_pdf;
    if(P.opt.mis_mode == TR_MIS_POWER_HEURISTIC)
        return (nee_pdf * nee_pdf + bsdf_pdf * bsdf_pdf) / nee_pdf;
    else if(P.opt.mis_mode == TR_MIS_BALANCE_HEURISTIC)
        return nee_pdf + bsdf_pdf;
    return nee_pdf;
}

TRD float clamp_contribution_mul(const RenderParams& P, vec3 contrib)
{
    if(P.opt.indirect_clamping > 0.0f)
    {
        float m = rgb_to_luminance(contrib);
        if(m > P.opt.indirect_clamping) return P.opt.indirect_clamping / m;
    }
    return 1;
}

TRD void correct_lobes_for_normal_map(vec3 sample_dir, vec3 geometric_normal, bsdf_lobes& lobes)
{
    if(dot(geometric_normal, sample_dir) < 0)
    {
        lobes.diffuse = 0;
        lobes.dielectric_reflection = 0;
        lobes.metallic_reflection = 0;
    }
    else lobes.transmission = 0;
}

TRD bool nee_enabled(const RenderParams& P)
{
    return (P.opt.sampling_weight_point_lights > 0 && P.scene.point_light_count > 0) ||
        (P.opt.sampling_weight_directional_lights > 0 && P.scene.directional_light_count > 0) ||
        (P.opt.sampling_weight_emissive_triangles > 0 && P.scene.tri_light_count > 0) ||
        (P.opt.sampling_weight_envmap > 0 && P.scene.environment_proj >= 0);
}

// shader/path_tracer.glsl:203-289
TRD vec3 sample_explicit_light(const RenderParams& P, uvec4 rand_uint, vec3 pos, vec3& out_dir, float& out_length, float& pdf)
{
    const NeeProbs pr = get_nee_sampling_probabilities(P);
    vec4 u = to_unit(rand_uint);
    if(P.opt.sampling_weight_point_lights > 0 && (u.w -= pr.point) < 0)
    {
        const int light_count = (int)P.scene.point_light_count;
        int light_index = clampi_((int)(u.z * (float)light_count), 0, light_count - 1);
        float weight = (float)(light_count > 1 ? light_count : 1);
        const tr_point_light pl = P.scene.point_lights[light_index];
        vec3 color;
        sample_point_light(pl, V2(u.x, u.y), pos, out_dir, out_length, color, pdf);
        pdf *= pr.point / weight;
        return color;
    }
    else if(P.opt.sampling_weight_emissive_triangles > 0 && (u.w -= pr.triangle) < 0)
    {
        const int light_count = (int)P.scene.tri_light_count;
        int light_index = clampi_((int)(u.z * (float)light_count), 0, light_count - 1);
        const float4* tlp = reinterpret_cast<const float4*>(P.scene.tri_lights + light_index);
        const float4 l0 = __ldg(tlp), l1 = __ldg(tlp + 1), l2 = __ldg(tlp + 2), l3 = __ldg(tlp + 3);
        vec3 A = V3(l0.x, l0.y, l0.z) - pos;
        vec3 B = V3(l0.w, l1.x, l1.y) - pos;
        vec3 C = V3(l1.z, l1.w, l2.x) - pos;
        vec3 color = r9g9b9e5_to_rgb(__float_as_uint(l2.y));
        float tri_pdf = 0.0f;
        out_dir = sample_triangle_light(P.opt.tri_light_mode, V2(u.x, u.y), A, B, C, tri_pdf);
        out_length = ray_plane_intersection_dist(out_dir, A, B, C);
        if(isinf_(tri_pdf) || tri_pdf <= 0 || out_length <= P.opt.min_ray_dist || any_nan(out_dir))
        {
            pdf = 1.0f;
            out_dir = V3(0);
            return V3(0);
        }
        const int emission_tex_id = __float_as_int(l3.w);
        if(emission_tex_id >= 0)
        {
            vec3 bary = get_barycentric_coords(out_dir * out_length, A, B, C);
            const uint32_t h0 = __float_as_uint(l3.x), h1 = __float_as_uint(l3.y), h2 = __float_as_uint(l3.z);
            vec2 uv0 = V2(half_to_float_(h0 & 0xFFFFu), half_to_float_(h0 >> 16));
            vec2 uv1 = V2(half_to_float_(h1 & 0xFFFFu), half_to_float_(h1 >> 16));
            vec2 uv2 = V2(half_to_float_(h2 & 0xFFFFu), half_to_float_(h2 >> 16));
            vec2 uv = bary.x * uv0 + bary.y * uv1 + bary.z * uv2;
            color = color * xyz(sample_texture(P.scene, emission_tex_id, uv));
        }
        out_length -= P.opt.min_ray_dist;
        pdf = pr.triangle * tri_pdf / (float)light_count;
        return color;
    }
    else if(P.opt.sampling_weight_envmap > 0 && (u.w -= pr.envmap) < 0)
    {
        vec3 color = sample_environment_map(P.scene, rand_uint, out_dir, out_length, pdf);
        pdf *= pr.envmap;
        return color;
    }
    else if(P.opt.sampling_weight_directional_lights > 0 && (u.w -= pr.directional) < 0)
    {
        const int light_count = (int)P.scene.directional_light_count;
        int light_index = clampi_((int)(u.z * (float)light_count), 0, light_count - 1);
        const tr_directional_light dl = P.scene.directional_lights[light_index];
        out_length = inf_();
        vec3 color;
        sample_directional_light(dl, V2(u.x, u.y), out_dir, color, pdf);
        pdf *= pr.directional / (float)light_count;
        return color;
    }
    pdf = 1.0f;
    out_dir = V3(0);
    out_length = 0.0f;
    return V3(0);
}

TRD uvec4 generate_ray_sample_uint(const RenderParams& P, const SobolTables& tables, const uvec4& ss_seed, uvec4& rs_seed, uint32_t bounce_index)
{
    if(P.opt.local_sampler == TR_SAMPLER_SOBOL_OWEN)
        return get_shuffled_scrambled_sobol_pt_uint(tables, ss_seed, bounce_index, P.frame.max_sobol_bounces);
    if(P.opt.local_sampler >= TR_SAMPLER_SOBOL_Z_ORDER_2D)   // get_sobol_z_sample_uint, sobol_z_sampler.glsl:19-22
        return generate_sobol_sample(tables, ss_seed.x, bounce_index, P.frame.max_sobol_bounces);
    return pcg4d(rs_seed);
}

// ---------------------------------------------------------------------------------------------
// camera: shader/camera.glsl:25-59, shader/rt.glsl:234-247, shader/path_tracer.glsl:504-533
// ---------------------------------------------------------------------------------------------
TRD void get_camera_ray(const RenderParams& P, const tr_camera_data& cam, vec2 pixel_coord, vec2 screen_size,
                        vec2 dof_u, vec3& origin, vec3& dir)
{
    vec2 uv = pixel_coord / screen_size;
    uv = V2(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f);
    if(P.opt.projection == TR_PROJECTION_PERSPECTIVE)
    {
        if(P.opt.depth_of_field)
        {
            vec2 aperture_offset = cam.dof_params[3] == 0 ?
                sample_concentric_disk(dof_u) :
                sample_regular_polygon(dof_u, cam.dof_params[2], (uint32_t)cam.dof_params[3]);
            vec3 view_origin = V3(aperture_offset.x * cam.dof_params[1], aperture_offset.y * cam.dof_params[1], 0);
            vec3 view_dir = xyz(mul4(cam.proj_inverse, V4(uv.x, uv.y, 1, 1))) * cam.dof_params[0];
            view_dir = normalize(view_dir - view_origin);
            origin = xyz(mul4(cam.view_inverse, V4(view_origin, 1.0f)));
            dir = normalize(xyz(mul4(cam.view_inverse, V4(view_dir, 0))));
        }
        else
        {
            origin = V3(cam.origin[0], cam.origin[1], cam.origin[2]);
            vec4 t = mul4(cam.proj_inverse, V4(uv.x, uv.y, 1, 1));
            dir = normalize(xyz(mul4(cam.view_inverse, V4(t.x, t.y, t.z, 0))));
        }
    }
    else
    {
        origin = xyz(mul4(cam.view_inverse, mul4(cam.proj_inverse, V4(uv.x, uv.y, 0, 1))));
        dir = normalize(xyz(mul4(cam.view_inverse, V4(0, 0, -1, 0))));
    }
}

TRD void get_screen_camera_ray(const RenderParams& P, const tr_camera_data& cam, uint32_t px, uint32_t py,
                               vec2 pixel_offset, vec2 dof_u, vec3& origin, vec3& dir)
{
    vec2 p = V2((float)(int)px + (pixel_offset.x * 0.5f + 0.5f), (float)(int)py + (pixel_offset.y * 0.5f + 0.5f));
    p.y = (float)P.frame.size_y - p.y;
    get_camera_ray(P, cam, p, V2((float)P.frame.size_x, (float)P.frame.size_y), dof_u, origin, dir);
}

// path_tracer.rgen:90-100 + evaluate_ray prologue (path_tracer.glsl:376-384)
TRD void path_begin(const RenderParams& P, uint32_t px, uint32_t py, uint32_t view, uint32_t sample_in_pass,
                    PathState& s, uvec4& ss_seed, uvec4& rs_seed)
{
    local_sampler ls = init_local_sampler(px, py, view, P.frame.previous_samples + sample_in_pass,
                                          P.frame.sample_counter, P.frame.rng_seed, P.opt.local_sampler);
    vec2 cam_offset = V2(0, 0);
    if(P.opt.film != TR_FILM_POINT)
    {
        vec4 r = to_unit(pcg4d(ls.rs_seed));
        if(P.opt.film == TR_FILM_BOX) cam_offset = V2(r.x * 2.0f - 1.0f, r.y * 2.0f - 1.0f);
        else cam_offset = sample_blackman_harris_concentric_disk(V2(r.x, r.y)) * 2.0f;
        cam_offset = cam_offset * (2.0f * P.opt.film_radius);
    }
    vec2 dof_u = V2(0.5f, 0.5f);
    if(P.opt.depth_of_field)
    {
        vec4 r = to_unit(pcg4d(ls.rs_seed));
        dof_u = V2(r.x, r.y);
    }
    get_screen_camera_ray(P, P.cameras[view].current, px, py, cam_offset, dof_u, s.pos, s.view);
    s.attenuation = V3(1);
    s.diffuse = V3(0);
    s.reflection = V3(0);
    s.aux_alpha = 0.0f;
    s.regularization = 1.0f;
    s.bsdf_pdf = 0.0f;
    s.primary_lobes = LOBES(0, 0, 0, 1);
    s.random_seed = pcg4d(ls.rs_seed).x;
    s.fh_emission = V3(0);
    s.fh_albedo = V4(0, 0, 0, 0);
    s.fh_metallic = 1.0f;
    ss_seed = ls.ss_seed;
    rs_seed = ls.rs_seed;
}

TRD void add_demodulated(PathState& s, vec3 light_color, float w_diffuse, float w_reflection)
{
    s.diffuse += light_color * w_diffuse;
    s.reflection += light_color * w_reflection;
}

// get_intersection_info (path_tracer.glsl:91-201) for a hit record `h` of the ray (origin, view).
struct HitInfo
{
    IntersectionPdf nee_pdf;
    sampled_material mat;
    vec3 light;
    vec3 pos, hard_normal, mapped_normal;
    int instance_id;
    bool surface;
};

TRD void get_intersection_info(const RenderParams& P, const Hit& h, vec3 origin, vec3 view, HitInfo& out)
{
    IntersectionPdf nee_pdf;
    nee_pdf.point_light_pdf = 0; nee_pdf.directional_light_pdf = 0; nee_pdf.tri_light_pdf = 0; nee_pdf.envmap_pdf = 0;
    sampled_material mat;
    mat.albedo = V4(0, 0, 0, 0); mat.metallic = 1; mat.roughness = 0; mat.emission = V3(0);
    mat.transmittance = 0; mat.ior_in = 0; mat.ior_out = 0; mat.f0 = 0;
    vec3 light = V3(0);
    vec3 v_pos, v_hard = V3(0), v_mapped;
    bool surface;
    if(h.instance_id >= 0)
    {
        const DeviceInstance& inst = P.scene.instances[h.instance_id];
        float pdf;
        SurfaceVertex vd = get_interpolated_vertex(P, inst, view, h.u, h.v, h.primitive_id, origin, pdf);
        mat = sample_material(P, inst, vd);
        mat.albedo.w = 1.0f;
        if(P.opt.sampling_weight_emissive_triangles > 0)
        {
            nee_pdf.tri_light_pdf = pdf == 0.0f ? 0.0f : pdf;
            light = mat.emission;
            mat.emission = V3(0);
        }
        v_pos = vd.pos; v_hard = vd.hard_normal; v_mapped = vd.mapped_normal;
        surface = true;
    }
    else if(h.primitive_id >= 0)
    {
        const tr_point_light pl = P.scene.point_lights[h.primitive_id];
        vec3 color = get_spotlight_intensity(pl, view) * ld3(pl.color) / (pl.radius * pl.radius * TR_PI);
        if(P.opt.sampling_weight_point_lights > 0)
        {
            light = color;
            nee_pdf.point_light_pdf = sample_point_light_pdf(pl, origin);
        }
        else mat.emission = color;
        v_pos = origin + h.u * view;
        v_mapped = normalize(v_pos - ld3(pl.pos));
        mat.albedo = V4(0, 0, 0, 1);
        surface = false;
    }
    else
    {
        vec3 env = V3(P.scene.environment_factor[0], P.scene.environment_factor[1], P.scene.environment_factor[2]);
        if(P.scene.environment_proj >= 0)
        {
            vec2 uv;
            uv.y = asin_(-view.y) / TR_PI + 0.5f;
            uv.x = atan2_(view.z, view.x) / (2.0f * TR_PI) + 0.5f;
            env = env * xyz(sample_envmap_texture(P.scene, uv));
        }
        for(uint32_t i = 0; i < P.scene.directional_light_count; ++i)
        {
            const tr_directional_light dl = P.scene.directional_lights[i];
            if(dl.dir_cutoff >= 1.0f) continue;
            float visible = step_(dl.dir_cutoff, dot(view, -ld3(dl.dir)));
            vec3 col = visible * ld3(dl.color) / (2.0f * TR_PI * (1.0f - dl.dir_cutoff));
            if(P.opt.sampling_weight_directional_lights > 0)
            {
                light += col;
                nee_pdf.directional_light_pdf += visible * sample_directional_light_pdf(dl);
            }
            else mat.emission += col;
        }
        v_pos = origin;
        v_mapped = -view;
        if(P.opt.sampling_weight_envmap > 0)
        {
            light += env;
            nee_pdf.envmap_pdf = P.scene.environment_proj >= 0 ? sample_environment_map_pdf(P.scene, view) : 0.0f;
        }
        else mat.emission += env;
        surface = false;
    }
    out.nee_pdf = nee_pdf; out.mat = mat; out.light = light;
    out.pos = v_pos; out.hard_normal = v_hard; out.mapped_normal = v_mapped;
    out.instance_id = surface ? h.instance_id : -1;
    out.surface = surface;
}

// The body of evaluate_ray's bounce loop (path_tracer.glsl:405-497) for a hit record `h`.
// Returns true when the path continues (s.pos / s.view hold the next ray); `job.bounce >= 0`
// when an NEE sample was drawn (finish it with nee_finish, after tracing the shadow ray if
// job.needs_shadow_ray).
TRD bool shade_bounce(const RenderParams& P, const SobolTables& tables, uint32_t bounce, const Hit& h, PathState& s,
                      const uvec4& ss_seed, uvec4& rs_seed, NeeJob& job)
{
    job.bounce = -1;
    const uint32_t max_bounces = (uint32_t)P.opt.max_ray_depth;
    const vec3 origin = s.pos;
    const vec3 view = s.view;

    HitInfo hi;
    get_intersection_info(P, h, origin, view, hi);
    const IntersectionPdf nee_pdf = hi.nee_pdf;
    sampled_material mat = hi.mat;
    vec3 light = hi.light;
    const vec3 v_pos = hi.pos, v_hard = hi.hard_normal, v_mapped = hi.mapped_normal;
    const bool surface = hi.surface;
    const bool terminal = !surface || bounce == max_bounces - 1;

    // ---- MIS-weighted emission, path_tracer.glsl:411-435
    float mis_pdf = bsdf_mis_pdf(P, nee_pdf, s.bsdf_pdf);
    float mis_weight = 1.0f;
    if(s.bsdf_pdf != 0)
    {
        s.attenuation /= s.bsdf_pdf;
        mis_weight = s.bsdf_pdf / mis_pdf;
    }
    light = s.attenuation * mis_weight * (mat.emission + light);
    if(bounce != 0) light *= clamp_contribution_mul(P, light);
    add_demodulated(s, light, s.primary_lobes.diffuse + s.primary_lobes.transmission,
                    s.primary_lobes.dielectric_reflection + s.primary_lobes.metallic_reflection);
    if(bounce == 0)
    {
        s.fh_emission = light;
        s.fh_albedo = mat.albedo;
        s.fh_metallic = mat.metallic;
    }

    // ---- path space regularisation, :437-444
    if(P.opt.regularization_gamma != 0.0f)
    {
        if(s.bsdf_pdf != 0.0f)
            s.regularization *= max_(1 - P.opt.regularization_gamma / pow_(s.bsdf_pdf, 0.25f), 0.0f);
        mat.roughness = 1.0f - ((1.0f - mat.roughness) * s.regularization);
    }
    if(terminal) return false;
    if(bounce == 1) s.aux_alpha = 1.0f / length(v_pos - origin);   // diffuse.a = reflection.a, path_tracer.glsl:470-471

    const mat3 tbn = create_tangent_space(v_mapped);
    const vec3 shading_view = view_to_tangent_space(view, tbn);

    // ---- next event estimation up to the shadow ray, :302-339 and :449-456
    // The NEE draw is an argument of next_event_estimation(): it is evaluated (and advances the PCG
    // stream) even when no light class is enabled.
    const uvec4 nee_rand = generate_ray_sample_uint(P, tables, ss_seed, rs_seed, bounce * 2);
    if(nee_enabled(P))
    {
        vec3 out_dir;
        float out_length = 0.0f;
        float light_pdf;
        vec3 contrib = sample_explicit_light(P, nee_rand, v_pos, out_dir, out_length, light_pdf);
        vec3 shading_light = mul(out_dir, tbn);
        bsdf_lobes lobes = LOBES(0, 0, 0, 0);
        float bsdf_pdf = material_bsdf_pdf(P.opt.bounce_mode, shading_light, shading_view, mat, lobes);
        correct_lobes_for_normal_map(out_dir, v_hard, lobes);
        job.bounce = (int)bounce;
        job.dir = out_dir;
        job.tmax = out_length;
        job.contrib = contrib;
        job.needs_shadow_ray = (contrib.x > 0.0001f || contrib.y > 0.0001f || contrib.z > 0.0001f) ? 1 : 0;
        job.mis_pdf = nee_mis_pdf(P, light_pdf, bsdf_pdf);
        job.attenuation = s.attenuation;
        if(bounce != 0)
        {
            job.modulate = modulate_bsdf(xyz(mat.albedo), lobes);
            job.w_diffuse = s.primary_lobes.diffuse + s.primary_lobes.transmission;
            job.w_reflection = s.primary_lobes.dielectric_reflection + s.primary_lobes.metallic_reflection;
        }
        else
        {
            job.modulate = V3(1);
            job.w_diffuse = lobes.diffuse + lobes.transmission;
            job.w_reflection = lobes.dielectric_reflection + lobes.metallic_reflection;
        }
    }
    else
    {
        // next_event_estimation returns vec3(0): radiance = attenuation * 0 is still accumulated
        // (it can only matter through NaN/Inf attenuation, kept for fidelity).
        job.bounce = (int)bounce;
        job.dir = V3(0); job.tmax = 0; job.contrib = V3(0); job.needs_shadow_ray = 0; job.mis_pdf = 1.0f;
        job.attenuation = s.attenuation;
        bsdf_lobes lobes = LOBES(0, 0, 0, 0);
        job.modulate = bounce != 0 ? modulate_bsdf(xyz(mat.albedo), lobes) : V3(1);
        if(bounce != 0)
        {
            job.w_diffuse = s.primary_lobes.diffuse + s.primary_lobes.transmission;
            job.w_reflection = s.primary_lobes.dielectric_reflection + s.primary_lobes.metallic_reflection;
        }
        else { job.w_diffuse = 0; job.w_reflection = 0; }
    }

    // ---- sample the BSDF for the next ray, :476-497
    bsdf_lobes lobes = LOBES(0, 0, 0, 0);
    vec4 ray_sample = to_unit(generate_ray_sample_uint(P, tables, ss_seed, rs_seed, bounce * 2 + 1));
    vec3 out_dir;
    float pdf;
    material_bsdf_sample(P.opt.bounce_mode, ray_sample, shading_view, mat, out_dir, lobes, pdf);
    s.bsdf_pdf = pdf;
    s.view = mul(tbn, out_dir);
    correct_lobes_for_normal_map(v_hard, s.view, lobes);
    if(bounce != 0) s.attenuation *= modulate_bsdf(xyz(mat.albedo), lobes);
    else s.primary_lobes = lobes;
    s.pos = v_pos;
    if(P.opt.russian_roulette_delta > 0)
    {
        float qi = min_(1.0f, 1.0f / P.opt.russian_roulette_delta);
        if(ray_sample.w > qi) return false;
    }
    if(max_(s.attenuation.x, max_(s.attenuation.y, s.attenuation.z)) <= 0.0f) return false;
    return true;
}

// path_tracer.glsl:336-340 + :453-469, given the shadow ray's visibility (ignored unless
// job.needs_shadow_ray).
TRD void nee_finish(const RenderParams& P, const NeeJob& job, float visibility, PathState& s)
{
    vec3 contrib = job.contrib;
    if(job.needs_shadow_ray) contrib *= visibility;
    contrib /= job.mis_pdf;
    vec3 radiance = job.attenuation * contrib;
    if(job.bounce != 0)
    {
        radiance *= job.modulate;
        radiance *= clamp_contribution_mul(P, radiance);
    }
    add_demodulated(s, radiance, job.w_diffuse, job.w_reflection);
}

// A pending NEE sample whose shadow ray cannot change the image: every lobe weight that multiplies the
// radiance is exactly 0 and the radiance itself is finite for visibility 1 (then it is finite -- namely
// 0 -- for visibility 0 as well), so the reference would add +-0 whatever the ray returns. Such rays are
// not traced (counted in culled_shadow_rays); the result is bit-identical.
TRD bool nee_shadow_ray_is_moot(const NeeJob& job)
{
    if(!job.needs_shadow_ray) return false;
    const bool zero_weight = job.bounce != 0 ?
        ((job.modulate.x == 0.0f && job.modulate.y == 0.0f && job.modulate.z == 0.0f) ||
         (job.w_diffuse == 0.0f && job.w_reflection == 0.0f)) :
        (job.w_diffuse == 0.0f && job.w_reflection == 0.0f);
    if(!zero_weight) return false;
    const vec3 r = job.attenuation * (job.contrib / job.mis_pdf);
    const float big = 1e30f;   // finite with room for the remaining (zero) multiplications
    return fabsf(r.x) < big && fabsf(r.y) < big && fabsf(r.z) < big;
}

// path_tracer.rgen:108-117
TRD vec3 path_color(const RenderParams& P, const PathState& s)
{
    vec3 albedo = P.opt.use_white_albedo_on_first_bounce ? V3(1) : xyz(s.fh_albedo);
    return s.fh_emission + modulate_color(albedo, s.fh_metallic, s.diffuse, s.reflection);
}

// ---------------------------------------------------------------------------------------------
// aux G-buffer stores of write_all_outputs (path_tracer.glsl:547-574); texel formats src/gbuffer.hh:18-51
// ---------------------------------------------------------------------------------------------
TRD uint2 pack_rgba16f(vec4 v)
{
    return make_uint2(float_to_half_(v.x) | (float_to_half_(v.y) << 16), float_to_half_(v.z) | (float_to_half_(v.w) << 16));
}
TRD vec4 unpack_rgba16f(uint2 q)
{
    return V4(half_to_float_(q.x & 0xFFFFu), half_to_float_(q.x >> 16), half_to_float_(q.y & 0xFFFFu), half_to_float_(q.y >> 16));
}
// accumulate_gbuffer_diffuse / _reflection (gbuffer.glsl:68-78,118-128) on an rgba16f layer
TRD void accumulate_rgba16f(uint2* layer, const uint2* layer_prev, uint32_t index, vec4 v, uint32_t samples, uint32_t previous_samples)
{
    if(previous_samples != 0)
    {
        const vec4 prev = unpack_rgba16f(layer_prev[index]);
        const uint32_t tot = samples + previous_samples;
        v = mix_(v, prev, (float)previous_samples / (float)tot);
    }
    layer[index] = pack_rgba16f(v);
}
// image-store conversions: clamp, scale, round to nearest even; NaN -> 0
TRD uint32_t float_to_unorm8_(float v)
{
    if(!(v > 0.0f)) return 0u;
    if(v >= 1.0f) return 255u;
    return (uint32_t)__float2int_rn(v * 255.0f);
}
TRD uint32_t float_to_snorm16_(float v)
{
    if(isnan_(v)) return 0u;
    const float c = v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v);
    return (uint32_t)__float2int_rn(c * 32767.0f) & 0xFFFFu;
}
TRD vec2 octahedral_pack(vec3 n) // math.glsl:480-485
{
    n = n / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    if(n.z >= 0.0f) return V2(n.x, n.y);
    return V2((1.0f - fabsf(n.y)) * (step_(0.0f, n.x) * 2.0f - 1.0f), (1.0f - fabsf(n.x)) * (step_(0.0f, n.y) * 2.0f - 1.0f));
}
// The layers written by the first sample only (prev_samples == 0), from the first hit of the path.
TRD void write_first_sample_layers(const RenderParams& P, uint32_t texel, uint32_t view, const HitInfo& hi)
{
    if(P.gbuffer[TR_GBUFFER_ALBEDO]) reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_ALBEDO])[texel] = pack_rgba16f(hi.mat.albedo);
    if(P.gbuffer[TR_GBUFFER_MATERIAL])
    {
        const float ior = hi.mat.ior_out / hi.mat.ior_in;   // pack_gbuffer_material, gbuffer.glsl:256-260
        reinterpret_cast<uint32_t*>(P.gbuffer[TR_GBUFFER_MATERIAL])[texel] =
            float_to_unorm8_(hi.mat.metallic) | (float_to_unorm8_(hi.mat.roughness) << 8) |
            (float_to_unorm8_(ior * 0.25f) << 16) | (float_to_unorm8_(hi.mat.transmittance) << 24);
    }
    if(P.gbuffer[TR_GBUFFER_NORMAL])
    {
        const vec2 o = octahedral_pack(hi.mapped_normal);
        reinterpret_cast<uint32_t*>(P.gbuffer[TR_GBUFFER_NORMAL])[texel] = float_to_snorm16_(o.x) | (float_to_snorm16_(o.y) << 16);
    }
    if(P.gbuffer[TR_GBUFFER_POS])
        reinterpret_cast<float4*>(P.gbuffer[TR_GBUFFER_POS])[texel] = make_float4(hi.pos.x, hi.pos.y, hi.pos.z, 0.0f);
    if(P.gbuffer[TR_GBUFFER_SCREEN_MOTION])
    {
        // prev_pos (rt.glsl:73-79, PRE_TRANSFORMED_VERTICES) + get_camera_projection (camera.glsl:61-67)
        vec3 prev_pos = hi.pos;
        if(hi.instance_id >= 0) prev_pos = xyz(mul4(P.prev_from_world + 16 * (size_t)hi.instance_id, V4(hi.pos, 1.0f)));
        const vec4 pp = mul4(P.cameras[view].previous.view_proj, V4(prev_pos, 1.0f));
        reinterpret_cast<float4*>(P.gbuffer[TR_GBUFFER_SCREEN_MOTION])[texel] =
            make_float4(pp.x / pp.w * 0.5f + 0.5f, pp.y / pp.w * 0.5f + 0.5f, pp.w, 0.0f);
    }
    if(P.gbuffer[TR_GBUFFER_INSTANCE_ID]) reinterpret_cast<int32_t*>(P.gbuffer[TR_GBUFFER_INSTANCE_ID])[texel] = hi.instance_id;
}

// write_all_outputs (path_tracer.glsl:535-576) + accumulate_gbuffer_color (gbuffer.glsl:18-28)
TRD float4 accumulate(const RenderParams& P, vec3 color, float alpha, uint32_t write_index, uint32_t pass_offset = 0)
{
    const uint32_t samples = (uint32_t)P.opt.samples_per_pass;
    const uint32_t prev_samples = P.frame.samples_accumulated + P.frame.previous_samples + pass_offset * samples;
    vec4 out = V4(color, alpha);
    if(prev_samples != 0)
    {
        // the first pass of a frame continues from what the previous frame left
        const float4 p = (P.frame.previous_samples + pass_offset == 0 ? P.color_prev : P.color)[write_index];
        const uint32_t total = samples + prev_samples;
        out = mix_(out, V4(p.x, p.y, p.z, p.w), (float)prev_samples / (float)total);
    }
    const float4 o = make_float4(out.x, out.y, out.z, out.w);
    P.color[write_index] = o;
    return o;
}

template<typename T> TRD T* peer_texels(const RenderParams& P, int k, size_t half, int layer)
{
    return reinterpret_cast<T*>(reinterpret_cast<char*>(P.peers[k]) + half + P.full_layer[layer]);
}

// Fused stitch (replaces device_transfer + stitch_shuffled_strips.comp:21-63): the finished pixel
// goes straight to its un-shuffled position in every peer's full image over NVLink.
TRD void store_to_peers(const RenderParams& P, float4 o, uint32_t px, uint32_t py, uint32_t view)
{
    if(P.peer_count == 0 || !P.last_pass) return;
    const size_t dst = (size_t)view * P.frame.size_x * P.frame.size_y + (size_t)py * P.frame.size_x + px;
    if(P.stitch_blend < 1.0f)
    {
        // stitch_shuffled_strips.comp:50-57: every image holds the same previous content, so all stay identical.
        // The previous content is read from this rank's own image when it has one (no remote loads).
        for(int k = 0; k < P.peer_count; ++k)
        {
            const int from = P.peer_self >= 0 ? P.peer_self : k;
            const float4 q = peer_texels<float4>(P, from, P.full_prev, TR_GBUFFER_COLOR)[dst];
            const vec4 m = mix_(V4(q.x, q.y, q.z, q.w), V4(o.x, o.y, o.z, o.w), P.stitch_blend);
            peer_texels<float4>(P, k, P.full_cur, TR_GBUFFER_COLOR)[dst] = make_float4(m.x, m.y, m.z, m.w);
        }
        return;
    }
    for(int k = 0; k < P.peer_count; ++k)
        if(k != P.peer_self || peer_texels<float4>(P, k, P.full_cur, TR_GBUFFER_COLOR) != P.color)   // rendered in place already
            peer_texels<float4>(P, k, P.full_cur, TR_GBUFFER_COLOR)[dst] = o;
}

} // namespace trd
