// oracle_shading.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Scalar restatement of the material / BSDF / light code of Tauray's path tracer:
//   shader/material.glsl, shader/ggx.glsl, shader/light.glsl, shader/scene.glsl:88-152,
//   shader/sampling.glsl, shader/random_sampler.glsl, shader/sobol_owen_sampler.glsl,
//   shader/math.glsl:125-179.
// Expression order follows the GLSL text; see oracle_math.h for the arithmetic definition.
#pragma once
#include "oracle_math.h"
#include "../include/tauray_b200.h"
#include "../include/tr_sobol_table.inc"

namespace orc {

// ---------------------------------------------------------------------------------------------
// samplers
// ---------------------------------------------------------------------------------------------
struct local_sampler
{
    uvec4 rs_seed; // random_sampler.seed
    uvec4 ss_seed; // sobol_owen_sampler.seed
};

// shader/math.glsl:168-179
static inline uint32_t owen_scramble_1(uint32_t x, uint32_t seed)
{
    x = bitreverse_(x);
    x ^= x * 0x3D20ADEAu;
    x += seed;
    x *= (seed >> 16u) | 1u;
    x ^= x * 0x05526C56u;
    x ^= x * 0x53A22864u;
    return bitreverse_(x);
}
static inline uvec4 owen_scramble_2d(uvec4 x, uvec4 seed)
{
    return {owen_scramble_1(x.x, seed.x), owen_scramble_1(x.y, seed.y),
            owen_scramble_1(x.z, seed.z), owen_scramble_1(x.w, seed.w)};
}

// shader/math.glsl:125-153; MAX_SOBOL_BOUNCES = min(MAX_BOUNCES, 8)
// (shader/sobol_lookup_table.glsl:4-14). The loop really does stop before the MSB.
static inline uvec4 generate_sobol_sample(uint32_t index, uint32_t bounce, uint32_t max_sobol_bounces)
{
    uvec4 x = {0, 0, 0, 0};
    if(bounce >= max_sobol_bounces)
    {
        x = {index, bounce, bounce * index, 0};
        return pcg4d(x);
    }
    for(int bit = findLSB_(index); bit < findMSB_(index); bit++)
    {
        uint32_t mask = (index >> bit) & 1u;
        if(mask != 0)
        {
            const unsigned int* row = TR_SOBOL_TABLE[bounce * 32 + bit];
            x.x ^= row[0]; x.y ^= row[1]; x.z ^= row[2]; x.w ^= row[3];
        }
    }
    return x;
}

// shader/sobol_owen_sampler.glsl:11-21 (the sampler is passed by value; pcg4d mutates the copy)
static inline uvec4 get_shuffled_scrambled_sobol_pt_uint(uvec4 seed, uint32_t bounce, uint32_t max_sobol_bounces)
{
    uint32_t index = seed.w;
    seed.w = bounce;
    uvec4 h = pcg4d(seed);
    uint32_t shuffled_index = owen_scramble_1(index, h.x);
    seed.w = index;
    uvec4 yzwx = {seed.y, seed.z, seed.w, seed.x};
    return owen_scramble_2d(generate_sobol_sample(shuffled_index, bounce, max_sobol_bounces), yzwx);
}

// ---- Z-order Sobol sampler (shader/sobol_z_sampler.glsl) ------------------------------------------
// shader/math.glsl:85-93 (only .x of the result is used here)
static inline uint32_t pcg2d_x(uint32_t sx, uint32_t sy)
{
    sx = sx * 1664525u + 1013904223u; sy = sy * 1664525u + 1013904223u;
    uint32_t ox = sx, oy = sy;
    sx = ox + oy * 1664525u; sy = oy + ox * 1664525u;
    sx = (sx >> 16) ^ sx; sy = (sy >> 16) ^ sy;
    ox = sx; oy = sy;
    sx = ox + oy * 1664525u; sy = oy + ox * 1664525u;
    sx = (sx >> 16) ^ sx;
    return sx;
}
// shader/math.glsl:155-166
static inline uint32_t get_permutation_n(int n, uint32_t permutation, uint32_t dimension)
{
    uint32_t res = 0;
    for(int i = n - 1; i >= 0; --i)
    {
        uint32_t q = permutation % (uint32_t)(n - i);
        permutation /= (uint32_t)(n - i);
        if((uint32_t)i == dimension) res = q;
        if(dimension > (uint32_t)i) res += (res >= q) ? 1u : 0u;
    }
    return res;
}
// shader/math.glsl:181-203
static inline uint32_t owen_scramble_4d(uint32_t x, uint32_t seed)
{
    uint32_t result = 0;
    for(uint32_t i = 0; i < 32; i += 2)
        result |= get_permutation_n(4, pcg2d_x(seed, x & ((~3u) << i)) % 24u, (x >> i) & 3u) << i;
    return result;
}
static inline uint32_t owen_scramble_8d(uint32_t x, uint32_t seed)
{
    uint32_t result = 0;
    for(uint32_t i = 0; i < 32; i += 3)
        result |= get_permutation_n(8, pcg2d_x(seed, x & ((~7u) << i)) % 40320u, (x >> i) & 7u) << i;
    return result;
}
// shader/math.glsl:260-278
static inline uint32_t morton_spread_2d(uint32_t x)
{
    x &= 0x0000ffffu;
    x = (x ^ (x << 8)) & 0x00ff00ffu;
    x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u;
    x = (x ^ (x << 1)) & 0x55555555u;
    return x;
}
static inline uint32_t morton_2d(uint32_t x, uint32_t y) { return morton_spread_2d(x) + 2u * morton_spread_2d(y); }
static inline uint32_t morton_spread_3d(uint32_t x)
{
    x &= 0x000003ffu;
    x = (x ^ (x << 16)) & 0xff0000ffu;
    x = (x ^ (x << 8)) & 0x0300f00fu;
    x = (x ^ (x << 4)) & 0x030c30c3u;
    x = (x ^ (x << 2)) & 0x09249249u;
    return x;
}
static inline uint32_t morton_3d(uint32_t x, uint32_t y, uint32_t z)
{
    return morton_spread_3d(x) + 2u * morton_spread_3d(y) + 4u * morton_spread_3d(z);
}
// shader/sobol_z_sampler.glsl:24-40 (coord.z -- the view -- is not part of the key)
static inline uint32_t init_sobol_z_index(uvec4 coord, int curve_dims)
{
    if(curve_dims == 3) return owen_scramble_8d(morton_3d(coord.x, coord.y, coord.w), coord.w >> 10u);
    return owen_scramble_4d(morton_2d(coord.x, coord.y), coord.w);
}

// shader/sampling.glsl:32-45 + random_sampler.glsl:11-19 + sobol_owen_sampler.glsl:23-28;
// z_curve_dims: 0 = sobol-owen / uniform-random, 2 / 3 = sobol-z2 / sobol-z3 (ss_seed.x = sobol_index)
static inline local_sampler init_local_sampler(uvec4 coord, uint32_t sample_counter, uint32_t rng_seed, int z_curve_dims = 0)
{
    local_sampler ls;
    coord.w += sample_counter;
    coord.z += rng_seed;
    uvec4 s = coord;
    ls.ss_seed = pcg4d(s);
    if(z_curve_dims) ls.ss_seed = uvec4{init_sobol_z_index(coord, z_curve_dims), 0u, 0u, 0u};
    ls.rs_seed = coord;
    ls.rs_seed.y ^= pcg(ls.rs_seed.x);
    ls.rs_seed.z ^= pcg(ls.rs_seed.y);
    ls.rs_seed.w ^= pcg(ls.rs_seed.z);
    return ls;
}

static inline vec4 to_unit(uvec4 u)
{
    return V4((float)u.x * TR_INV_UINT32_MAX, (float)u.y * TR_INV_UINT32_MAX,
              (float)u.z * TR_INV_UINT32_MAX, (float)u.w * TR_INV_UINT32_MAX);
}

// ---------------------------------------------------------------------------------------------
// materials (shader/material.glsl)
// ---------------------------------------------------------------------------------------------
struct sampled_material
{
    vec4 albedo;
    float metallic;
    float roughness;
    vec3 emission;
    float transmittance;
    float ior_in;
    float ior_out;
    float f0;
    uint32_t flags;
    float shadow_terminator_mul;
};

struct bsdf_lobes
{
    float transmission;
    float diffuse;
    float dielectric_reflection;
    float metallic_reflection;
};

// material.glsl:52-55
static inline vec3 modulate_bsdf(const sampled_material& mat, bsdf_lobes bsdf)
{
    vec3 a = xyz(mat.albedo);
    float s = bsdf.metallic_reflection + bsdf.transmission + bsdf.diffuse;
    return V3(a.x * s + bsdf.dielectric_reflection, a.y * s + bsdf.dielectric_reflection,
              a.z * s + bsdf.dielectric_reflection);
}
// material.glsl:57-65
static inline vec3 modulate_color(const sampled_material& mat, vec3 diffuse, vec3 reflected)
{
    float approx_fresnel = 0.02f;
    vec3 a = xyz(mat.albedo);
    diffuse = diffuse * a * (1 - mat.metallic);
    reflected = reflected * mix_(V3(approx_fresnel), a, mat.metallic) / mix_(approx_fresnel, 1, mat.metallic);
    return diffuse + reflected;
}
// material.glsl:67-73
static inline void add_demodulated_color(bsdf_lobes primary, vec3 light_color, vec3& diffuse, vec3& reflected)
{
    diffuse += light_color * (primary.diffuse + primary.transmission);
    reflected += light_color * (primary.dielectric_reflection + primary.metallic_reflection);
}

// ---------------------------------------------------------------------------------------------
// GGX (shader/ggx.glsl)
// ---------------------------------------------------------------------------------------------
// ggx.glsl:30-33
static inline float ggx_fresnel_schlick(float cos_d, float f0)
{
    return f0 + (1.0f - f0) * pow5_(max_(1.0f - cos_d, 0.0f));
}
// ggx.glsl:36-49
static inline float ggx_fresnel(float cos_d, const sampled_material& mat)
{
    if(mat.ior_in > mat.ior_out)
    {
        float inv_eta = mat.ior_in / mat.ior_out;
        float sin_theta2 = inv_eta * inv_eta * (1.0f - cos_d * cos_d);
        if(sin_theta2 >= 1.0f) return 1.0f;
        cos_d = sqrtf(1.0f - sin_theta2);
    }
    else if(mat.ior_in == mat.ior_out) return 0.0f;
    return ggx_fresnel_schlick(cos_d, mat.f0);
}
// ggx.glsl:54-67
static inline float fresnel_importance(float cos_d, const sampled_material& mat)
{
    if(mat.ior_in > mat.ior_out)
    {
        float inv_eta = mat.ior_in / mat.ior_out;
        float sin_theta2 = inv_eta * inv_eta * (1.0f - cos_d * cos_d);
        if(sin_theta2 >= 1.0f) return 1.0f;
        cos_d = sqrtf(1.0f - sin_theta2);
    }
    else if(mat.ior_in == mat.ior_out) return 0.0f;
    return mat.f0 + (max_(1.0f - mat.roughness, mat.f0) - mat.f0) * pow5_(1.0f - cos_d);
}
// ggx.glsl:76-79
static inline float ggx_fresnel_refl(float cos_d, const sampled_material& mat)
{
    return ggx_fresnel_schlick(cos_d, mat.f0);
}
// ggx.glsl:82-87
static inline float ggx_masking(float v_dot_n, float v_dot_h, float a)
{
    float a2 = a * a;
    return step_(0.0f, v_dot_n * v_dot_h) * 2.0f /
        (1.0f + sqrtf(1.0f + a2 / (v_dot_n * v_dot_n) - a2));
}
// ggx.glsl:90-97
static inline float ggx_masking_shadowing(float v_dot_n, float v_dot_h, float l_dot_n, float l_dot_h, float a)
{
    float a2 = a * a;
    return step_(0.0f, v_dot_n * v_dot_h) * step_(0.0f, l_dot_n * l_dot_h) * 4.0f /
        ((1.0f + sqrtf(1.0f + a2 / max_(v_dot_n * v_dot_n, 1e-18f) - a2)) *
         (1.0f + sqrtf(1.0f + a2 / max_(l_dot_n * l_dot_n, 1e-18f) - a2)));
}
// ggx.glsl:101-109
static inline float ggx_masking_shadowing_predivided(float v_dot_n, float v_dot_h, float l_dot_n, float l_dot_h, float a)
{
    float a2 = a * a;
    float denom1 = fabsf(l_dot_n) * sqrtf(a2 + (1.0f - a2) * v_dot_n * v_dot_n);
    float denom2 = fabsf(v_dot_n) * sqrtf(a2 + (1.0f - a2) * l_dot_n * l_dot_n);
    return step_(0.0f, v_dot_n * v_dot_h) * step_(0.0f, l_dot_n * l_dot_h) * 0.5f / (denom1 + denom2);
}
// ggx.glsl:114-119
static inline float ggx_distribution(float h_dot_n, float a)
{
    float a2 = a * a;
    float denom = h_dot_n * h_dot_n * (a2 - 1.0f) + 1.0f;
    return a2 / (TR_PI * denom * denom);
}
// ggx.glsl:123-147
static inline void ggx_brdf_inner(vec3 out_dir, vec3 view_dir, vec3 h, float fresnel, float distribution,
                                  float cos_d, const sampled_material& mat, bsdf_lobes& bsdf)
{
    float cos_l = out_dir.z;
    float cos_v = view_dir.z;
    float geometry = ggx_masking_shadowing_predivided(cos_v, cos_d, cos_l, dot(out_dir, h), mat.roughness);
    float kd = (1.0f - fresnel) * (1.0f - mat.metallic) * (1.0f - mat.transmittance);
    cos_l = max_(cos_l, 0.0f);
    bsdf.diffuse += kd * cos_l / TR_PI;
    bsdf.dielectric_reflection += fresnel * geometry * distribution * cos_l * (1.0f - mat.metallic);
    bsdf.metallic_reflection += geometry * distribution * cos_l * mat.metallic;
}
// ggx.glsl:215-236
static inline vec3 ggx_vndf_sample(vec3 view, float roughness, float u1, float u2)
{
    vec3 v = normalize(V3(roughness * view.x, roughness * view.y, view.z));
    vec3 t1 = v.z < 0.9999f ? normalize(cross(v, V3(0, 0, 1))) : V3(1, 0, 0);
    vec3 t2 = cross(t1, v);
    float inv_a = 1.0f + v.z;
    float a = 1.0f / inv_a;
    float r = sqrtf(u1);
    float phi = u2 < a ? u2 * inv_a * TR_PI : TR_PI + (u2 - a) / (1.0f - a) * TR_PI;
    float sp, cp;
    sincos_(phi, &sp, &cp);
    float p1 = r * cp;
    float p2 = r * sp * (u2 < a ? 1.0f : v.z);
    float p3 = sqrtf(max_(0.0f, 1.0f - p1 * p1 - p2 * p2));
    vec3 n = p1 * t1 + p2 * t2 + p3 * v;
    return normalize(V3(roughness * n.x, roughness * n.y, max_(0.0f, n.z)));
}

#define TR_LOBE_ALL 0
#define TR_LOBE_DIFFUSE 1
#define TR_LOBE_TRANSMISSION 2
#define TR_LOBE_REFLECTION 3

// ggx.glsl:240-375 with eval_all_lobes = true (ggx_bsdf_sample, :377-388)
static inline void ggx_bsdf_sample(vec4 uniform_random, vec3 view_dir, const sampled_material& mat,
                                   vec3& out_dir, bsdf_lobes& bsdf, float& pdf)
{
    const bool zero_roughness = mat.roughness < 0.001f;
    vec3 h = zero_roughness ? V3(0, 0, 1) : ggx_vndf_sample(view_dir, mat.roughness, uniform_random.x, uniform_random.y);
    float cos_d = dot(view_dir, h);
    float fresnel = ggx_fresnel(cos_d, mat);
    float cos_v = view_dir.z;
    float max_albedo = max_(mat.albedo.x, max_(mat.albedo.y, mat.albedo.z));
    float specular_cutoff = mix_(1, fresnel_importance(view_dir.z, mat), (1 - mat.metallic) * max_albedo);
    float diffuse_cutoff = 1.0f - mat.transmittance;
    float specular_probability = specular_cutoff;
    float diffuse_probability = (1.0f - specular_cutoff) * diffuse_cutoff;
    float transmissive_probability = (1.0f - specular_cutoff) * (1.0f - diffuse_cutoff);
    float u = uniform_random.z;

    if(u <= specular_cutoff)
    { // Reflective
        out_dir = reflect_(-view_dir, h);
        float cos_l = out_dir.z;
        float cos_h = h.z;
        float G1 = ggx_masking(cos_v, cos_d, mat.roughness);
        float D = zero_roughness ? 4 * cos_l * cos_v : ggx_distribution(cos_h, mat.roughness);
        pdf = G1 * D / (4 * fabsf(cos_v)) * specular_probability;
        float diffuse_pdf = zero_roughness ? 0 : pdf_cosine_hemisphere(out_dir) * diffuse_probability;
        pdf += diffuse_pdf;
        ggx_brdf_inner(out_dir, view_dir, h, fresnel, D, cos_d, mat, bsdf);
        if(zero_roughness)
        {
            bsdf.diffuse = 0;
            bsdf.dielectric_reflection /= pdf;
            bsdf.metallic_reflection /= pdf;
            pdf = 0;
        }
    }
    else
    {
        u = clamp_((u - specular_cutoff) / (1 - specular_cutoff), 0.0f, 0.99999f);
        if(u <= diffuse_cutoff)
        { // Diffuse
            u = clamp_(u / diffuse_cutoff, 0.0f, 0.99999f);
            out_dir = sample_cosine_hemisphere(V2(u, uniform_random.w));
            h = normalize(view_dir + out_dir);
            float cos_h = h.z;
            cos_d = dot(view_dir, h);
            fresnel = ggx_fresnel_refl(cos_d, mat);
            float G1 = ggx_masking(cos_v, cos_d, mat.roughness);
            float D = zero_roughness ? 0 : ggx_distribution(cos_h, mat.roughness);
            pdf = pdf_cosine_hemisphere(out_dir) * diffuse_probability;
            float specular_pdf = G1 * D / (4 * fabsf(cos_v)) * specular_probability;
            pdf += specular_pdf;
            ggx_brdf_inner(out_dir, view_dir, h, fresnel, D, cos_d, mat, bsdf);
            if(zero_roughness)
            {
                bsdf.dielectric_reflection = 0;
                bsdf.metallic_reflection = 0;
            }
        }
        else
        { // Transmissive
            out_dir = normalize(refract_(-view_dir, h, mat.ior_in / mat.ior_out));
            if(any_nan(out_dir))
            {
                out_dir = V3(0);
                pdf = 0;
                return;
            }
            float cos_l = out_dir.z;
            float cos_h = h.z;
            float cos_o = dot(out_dir, h);
            float G2 = ggx_masking_shadowing(cos_v, cos_d, cos_l, cos_o, mat.roughness);
            float G1 = ggx_masking(cos_v, cos_d, mat.roughness);
            float D = zero_roughness ? 4 * cos_l * cos_v : ggx_distribution(cos_h, mat.roughness);
            float denom = mat.ior_in / mat.ior_out * cos_d + cos_o;
            bsdf.transmission += fabsf(cos_d * cos_o) * mat.transmittance * (1.0f - mat.metallic) * (1.0f - fresnel) * G2 * D / (denom * denom * fabsf(cos_v));
            pdf = (fabsf(cos_d * cos_o) * G1 * D) / (denom * denom * fabsf(cos_v)) * transmissive_probability;
            if(zero_roughness)
            {
                bsdf.transmission /= pdf;
                pdf = 0;
            }
        }
    }
}

// ggx.glsl:403-491 with lobe = MATERIAL_LOBE_ALL (ggx_bsdf_pdf, :493-500)
static inline float ggx_bsdf_pdf(vec3 out_dir, vec3 view_dir, const sampled_material& mat, bsdf_lobes& bsdf)
{
    float cos_l = out_dir.z;
    float cos_v = view_dir.z;
    vec3 h;
    if(cos_l > 0) h = normalize(view_dir + out_dir);
    else h = (mat.ior_in > mat.ior_out ? 1.0f : -1.0f) * normalize(mat.ior_out * out_dir + mat.ior_in * view_dir);
    float cos_h = h.z;
    float cos_d = dot(view_dir, h);
    float cos_o = dot(out_dir, h);
    float fresnel = ggx_fresnel(cos_d, mat);
    float geometry = ggx_masking_shadowing_predivided(cos_v, cos_d, cos_l, cos_o, mat.roughness);
    const bool zero_roughness = mat.roughness < 0.001f;
    float distribution = zero_roughness ? 0 : ggx_distribution(cos_h, mat.roughness);
    float max_albedo = max_(mat.albedo.x, max_(mat.albedo.y, mat.albedo.z));
    float specular_cutoff = mix_(1, fresnel_importance(view_dir.z, mat), (1 - mat.metallic) * max_albedo);
    float diffuse_cutoff = 1.0f - mat.transmittance;
    float specular_probability = specular_cutoff;
    float diffuse_probability = (1.0f - specular_cutoff) * diffuse_cutoff;
    float transmissive_probability = (1.0f - specular_cutoff) * (1.0f - diffuse_cutoff);
    float G1 = ggx_masking(cos_v, cos_d, mat.roughness);
    float pdf = 0.0f;
    if(cos_l > 0)
    {
        float kd = (1.0f - fresnel) * (1.0f - mat.metallic) * (1.0f - mat.transmittance);
        {
            float diffuse_pdf = pdf_cosine_hemisphere(out_dir) * diffuse_probability;
            if(!isnan_(diffuse_pdf) && !isinf_(diffuse_pdf) && diffuse_pdf > 0.0f)
            {
                bsdf.diffuse += kd * cos_l / TR_PI;
                pdf += diffuse_pdf;
            }
        }
        {
            float specular_pdf = G1 * distribution / (4 * fabsf(cos_v)) * specular_probability;
            if(!isnan_(specular_pdf) && !isinf_(specular_pdf) && specular_pdf > 0.0f)
            {
                bsdf.dielectric_reflection += fresnel * geometry * distribution * cos_l * (1.0f - mat.metallic);
                bsdf.metallic_reflection += geometry * distribution * cos_l * mat.metallic;
                pdf += specular_pdf;
            }
        }
    }
    else
    {
        float denom = mat.ior_in / mat.ior_out * cos_d + cos_o;
        geometry *= 4.0f;
        float transmit_pdf = (fabsf(cos_d * cos_o) * G1 * distribution) / (fabsf(cos_v) * denom * denom * TR_PI) * transmissive_probability;
        if(!isnan_(transmit_pdf) && !isinf_(transmit_pdf) && transmit_pdf > 0.0f)
        {
            bsdf.transmission += -cos_l * fabsf(cos_d * cos_o) * mat.transmittance * (1.0f - mat.metallic) * (1.0f - fresnel) * geometry * distribution / (denom * denom);
            pdf += transmit_pdf;
        }
    }
    return pdf;
}

// ggx.glsl:502-531
static inline void material_bsdf_sample(int bounce_mode, vec4 uniform_random, vec3 view_dir,
                                        const sampled_material& mat, vec3& out_dir, bsdf_lobes& bsdf, float& pdf)
{
    if(bounce_mode == TR_BOUNCE_HEMISPHERE)
    {
        if(mat.transmittance > 0.0f)
        {
            out_dir = sample_sphere(V2(uniform_random.x, uniform_random.y));
            pdf = 0.25f / TR_PI;
        }
        else
        {
            out_dir = sample_hemisphere(V2(uniform_random.x, uniform_random.y));
            pdf = 0.5f / TR_PI;
        }
        ggx_bsdf_pdf(out_dir, view_dir, mat, bsdf);
    }
    else if(bounce_mode == TR_BOUNCE_COSINE_HEMISPHERE)
    {
        float split = mat.transmittance * 0.5f;
        out_dir = (uniform_random.z < split ? -1.0f : 1.0f) * sample_cosine_hemisphere(V2(uniform_random.x, uniform_random.y));
        pdf = fabsf(out_dir.z / TR_PI) * (uniform_random.z < split ? split : 1.0f - split);
        ggx_bsdf_pdf(out_dir, view_dir, mat, bsdf);
    }
    else ggx_bsdf_sample(uniform_random, view_dir, mat, out_dir, bsdf, pdf);
}
// ggx.glsl:533-552
static inline float material_bsdf_pdf(int bounce_mode, vec3 out_dir, vec3 view_dir,
                                      const sampled_material& mat, bsdf_lobes& bsdf)
{
    if(bounce_mode == TR_BOUNCE_HEMISPHERE)
    {
        ggx_bsdf_pdf(out_dir, view_dir, mat, bsdf);
        if(mat.transmittance == 0 && out_dir.z <= 0) return 0.0f;
        return mat.transmittance > 0.0f ? 0.25f / TR_PI : 0.5f / TR_PI;
    }
    else if(bounce_mode == TR_BOUNCE_COSINE_HEMISPHERE)
    {
        ggx_bsdf_pdf(out_dir, view_dir, mat, bsdf);
        if(mat.transmittance == 0 && out_dir.z <= 0) return 0.0f;
        float split = mat.transmittance * 0.5f;
        return fabsf(out_dir.z / TR_PI) * (out_dir.z < 0 ? split : 1.0f - split);
    }
    return ggx_bsdf_pdf(out_dir, view_dir, mat, bsdf);
}

// ---------------------------------------------------------------------------------------------
// lights (shader/light.glsl)
// ---------------------------------------------------------------------------------------------
static inline vec3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }

// light.glsl:45-57
static inline float get_spotlight_intensity(const tr_point_light& l, vec3 dir)
{
    if(l.dir_falloff > 0)
    {
        float cutoff = dot(dir, -ld3(l.dir));
        cutoff = cutoff > l.dir_cutoff ?
            1.0f - pow_(max_(1.0f - cutoff, 0.0f) / (1.0f - l.dir_cutoff), l.dir_falloff) : 0.0f;
        return cutoff;
    }
    else return 1.0f;
}
// light.glsl:74-104
static inline void sample_point_light(const tr_point_light& pl, vec2 u, vec3 pos, vec3& out_dir,
                                      float& out_length, vec3& color, float& pdf)
{
    vec3 dir = pos - ld3(pl.pos);
    float dist2 = dot(dir, dir);
    float k = 1.0f - pl.radius * pl.radius / dist2;
    float dir_cutoff = k > 0 ? sqrtf(k) : -1.0f;
    out_dir = sample_cone(u, -normalize(dir), dir_cutoff);
    float b = dot(dir, out_dir);
    out_length = -b - sqrtf(max_(b * b - dist2 + pl.radius * pl.radius, 0.0f));
    color = get_spotlight_intensity(pl, normalize(-dir)) * ld3(pl.color);
    if(pl.radius == 0.0f) pdf = -dist2;
    else
    {
        color /= pl.radius * pl.radius * TR_PI;
        pdf = 1 / (2.0f * TR_PI * (1.0f - dir_cutoff));
    }
}
// light.glsl:106-115
static inline float sample_point_light_pdf(const tr_point_light& pl, vec3 pos)
{
    vec3 dir = pos - ld3(pl.pos);
    float dist2 = dot(dir, dir);
    float k = 1.0f - pl.radius * pl.radius / dist2;
    float dir_cutoff = k > 0 ? sqrtf(k) : -1.0f;
    if(pl.radius == 0.0f) return 0;
    else return 1 / (2.0f * TR_PI * (1.0f - dir_cutoff));
}
// light.glsl:117-127
static inline void sample_directional_light(const tr_directional_light& dl, vec2 u, vec3& out_dir, vec3& color, float& pdf)
{
    out_dir = sample_cone(u, -ld3(dl.dir), dl.dir_cutoff);
    pdf = dl.dir_cutoff >= 1.0f ? -1.0f : 1.0f / (2.0f * TR_PI * (1.0f - dl.dir_cutoff));
    color = pdf > 0 ? ld3(dl.color) * pdf : ld3(dl.color);
}
// light.glsl:129-132
static inline float sample_directional_light_pdf(const tr_directional_light& dl)
{
    return dl.dir_cutoff >= 1.0f ? 0.0f : 1.0f / (2.0f * TR_PI * (1.0f - dl.dir_cutoff));
}
// light.glsl:134-179
static inline float sample_triangle_light_pdf(int mode, vec3 P, vec3 A, vec3 B, vec3 C)
{
    if(mode == TR_TRI_LIGHT_AREA) return triangle_area_pdf(P, A, B, C);
    if(mode == TR_TRI_LIGHT_SOLID_ANGLE)
        return 1.0f / spherical_triangle_solid_angle(normalize(A), normalize(B), normalize(C));
    float solid_angle = spherical_triangle_solid_angle(normalize(A), normalize(B), normalize(C));
    return solid_angle > 1e-6f ? 1.0f / solid_angle : triangle_area_pdf(P, A, B, C);
}
static inline vec3 sample_triangle_light(int mode, vec2 u, vec3 A, vec3 B, vec3 C, float& pdf)
{
    if(mode == TR_TRI_LIGHT_AREA)
    {
        vec3 P = sample_triangle_area(u, A, B, C);
        pdf = triangle_area_pdf(P, A, B, C);
        return normalize(P);
    }
    if(mode == TR_TRI_LIGHT_SOLID_ANGLE) return sample_spherical_triangle(u, A, B, C, pdf);
    float solid_angle = spherical_triangle_solid_angle(normalize(A), normalize(B), normalize(C));
    if(solid_angle > 1e-6f) return sample_spherical_triangle(u, A, B, C, pdf);
    vec3 P = sample_triangle_area(u, A, B, C);
    pdf = triangle_area_pdf(P, A, B, C);
    return normalize(P);
}

// shader/rt_common.glsl:36-47
static inline float sphere_intersection(vec3 sphere_pos, float sphere_radius, vec3 ray_origin, vec3 ray_direction)
{
    vec3 oc = ray_origin - sphere_pos;
    float a = dot(ray_direction, ray_direction);
    float b = 2.0f * dot(oc, ray_direction);
    float c = dot(oc, oc) - sphere_radius * sphere_radius;
    float discriminant = b * b - 4.0f * a * c;
    if(discriminant < 0) return -1.0f;
    else return (-b - sqrtf(discriminant)) / (2.0f * a);
}

} // namespace orc
