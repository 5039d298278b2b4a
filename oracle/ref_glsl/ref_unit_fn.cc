// ref_unit_fn.cc -- TEST INFRASTRUCTURE ONLY. Function-level entry into the reference's shader include files
// (shader/math.glsl, sampling.glsl, color.glsl, material.glsl, ggx.glsl, light.glsl, camera.glsl, alias_table.glsl, read
// from /root/reference and rewritten lexically by build_ref.py). ref_function(fn, in, out, n) applies function `fn`
// to n packed records of 32-bit words; oracle.cc:orc_function implements the SAME table with the oracle's restatement,
// and tests/test_ref_glsl.py compares the two (integers bit for bit, floats within 1e-5 relative).
#include "glsl_shim.hpp"
#include "ref_api.h"

namespace glsl { namespace fn {
#include "ref_prelude.hpp"
void traceRayEXT(const accelerationStructureEXT&, uint, uint, uint, uint, uint, vec3, float, vec3, float, int) {}
#define SAMPLING_DATA_BINDING 1
#include "sampling.glsl.h"
#include "ggx.glsl.h"
#include "light.glsl.h"
#include "alias_table.glsl.h"
}}

using namespace glsl;
using namespace glsl::fn;

namespace {
struct rd
{
    const uint32_t* p;
    uint32_t u() { return *p++; }
    int i() { return (int)*p++; }
    float f() { return uintBitsToFloat(*p++); }
    vec2 v2() { float a = f(), b = f(); return vec2(a, b); }
    vec3 v3() { float a = f(), b = f(), c = f(); return vec3(a, b, c); }
    vec4 v4() { float a = f(), b = f(), c = f(), e = f(); return vec4(a, b, c, e); }
    uvec4 u4() { uint32_t a = u(), b = u(), c = u(), e = u(); return uvec4(a, b, c, e); }
    sampled_material mat()
    {
        sampled_material m;
        m.albedo = v4(); m.metallic = f(); m.roughness = f(); m.emission = vec3(0); m.transmittance = f();
        m.ior_in = f(); m.ior_out = f(); m.f0 = f(); m.flags = 0; m.shadow_terminator_mul = 1.0f;
        return m;
    }
    point_light pl() { point_light l; memcpy(&l, p, 64); p += 16; return l; }
    directional_light dl() { directional_light l; memcpy(&l, p, 32); p += 8; return l; }
};
struct wr
{
    uint32_t* p;
    void u(uint32_t v) { *p++ = v; }
    void f(float v) { *p++ = floatBitsToUint(v); }
    void v2(vec2 v) { f(v.x); f(v.y); }
    void v3(vec3 v) { f(v.x); f(v.y); f(v.z); }
    void u4(uvec4 v) { u(v.x); u(v.y); u(v.z); u(v.w); }
    void lobes(const bsdf_lobes& l) { f(l.transmission); f(l.diffuse); f(l.dielectric_reflection); f(l.metallic_reflection); }
};
// words in / words out per record, indexed by function id (keep in step with oracle.cc:orc_function)
const int k_words[][2] = {
    {1, 2}, {4, 8}, {2, 4}, {8, 4}, {7, 8}, {2, 1}, {2, 1}, {2, 1}, {3, 1}, {3, 9},
    {2, 2}, {2, 3}, {6, 3}, {4, 2}, {2, 2}, {11, 4}, {9, 1}, {11, 3}, {12, 1}, {12, 3},
    {17, 8}, {16, 5}, {21, 8}, {19, 1}, {10, 7}, {11, 4}, {12, 1}, {1, 3}, {3, 1}, {14, 3},
    {16, 3}, {86, 6}, {3, 2}, {3, 3}, {3, 1}, {5, 1}, {2, 3}, {6, 3}, {19, 1}, {12, 1},
    {12, 3}, {2, 3}, {2, 3}, {1, 1},
};
const int k_count = (int)(sizeof(k_words) / sizeof(k_words[0]));
}

extern "C" int ref_function_words(int id, int* words_in, int* words_out)
{
    if(id < 0 || id >= k_count) return -1;
    *words_in = k_words[id][0]; *words_out = k_words[id][1];
    return 0;
}

extern "C" int ref_function(int id, const uint32_t* in, uint32_t* out, uint32_t n)
{
    if(id < 0 || id >= k_count) return -1;
    for(uint32_t k = 0; k < n; ++k)
    {
        rd r{in + (size_t)k * k_words[id][0]};
        wr w{out + (size_t)k * k_words[id][1]};
        switch(id)
        {
        case 0: { uint s = r.u(); uint v = pcg(s); w.u(v); w.u(s); break; }
        case 1: { uvec4 s = r.u4(); uvec4 v = pcg4d(s); w.u4(v); w.u4(s); break; }
        case 2: { uint index = r.u(), bounce = r.u(); w.u4(generate_sobol_sample(index, bounce)); break; }
        case 3: { uvec4 x = r.u4(), s = r.u4(); w.u4(owen_scramble_2d(x, s)); break; }
        case 4:
        {   // init_local_sampler(coord) with the sampling UBO, one bounce draw, then the next film/alpha draw
            uvec4 coord = r.u4();
            sampling_data.sample_counter = r.u(); sampling_data.rng_seed = r.u();
            uint bounce = r.u();
            local_sampler ls = init_local_sampler(coord);
            w.u4(generate_ray_sample_uint(ls, bounce));
            w.u4(generate_uniform_random_uint(ls.rs));
            break;
        }
        case 5: { uint x = r.u(), s = r.u(); w.u(owen_scramble_4d(x, s)); break; }
        case 6: { uint x = r.u(), s = r.u(); w.u(owen_scramble_8d(x, s)); break; }
        case 7: { uint x = r.u(), y = r.u(); w.u(morton_2d(uvec2(x, y))); break; }
        case 8: { uint x = r.u(), y = r.u(), z = r.u(); w.u(morton_3d(uvec3(x, y, z))); break; }
        case 9: { mat3 m = create_tangent_space(r.v3()); w.v3(m[0]); w.v3(m[1]); w.v3(m[2]); break; }
        case 10: w.v2(sample_concentric_disk(r.v2())); break;
        case 11: w.v3(sample_cosine_hemisphere(r.v2())); break;
        case 12: { vec2 u = r.v2(); vec3 d = r.v3(); float c = r.f(); w.v3(sample_cone(u, d, c)); break; }
        case 13: { vec2 u = r.v2(); float a = r.f(); uint s = r.u(); w.v2(sample_regular_polygon(u, a, s)); break; }
        case 14: w.v2(sample_blackman_harris_concentric_disk(r.v2())); break;
        case 15: { vec2 xi = r.v2(); vec3 A = r.v3(), B = r.v3(), C = r.v3(); float pdf = 0; w.v3(sample_spherical_triangle(xi, A, B, C, pdf)); w.f(pdf); break; }
        case 16: { vec3 A = r.v3(), B = r.v3(), C = r.v3(); w.f(spherical_triangle_solid_angle(A, B, C)); break; }
        case 17: { vec2 u = r.v2(); vec3 A = r.v3(), B = r.v3(), C = r.v3(); w.v3(sample_triangle_area(u, A, B, C)); break; }
        case 18: { vec3 P = r.v3(), A = r.v3(), B = r.v3(), C = r.v3(); w.f(triangle_area_pdf(P, A, B, C)); break; }
        case 19: { vec3 v = r.v3(); mat3 m; m[0] = r.v3(); m[1] = r.v3(); m[2] = r.v3(); w.v3(view_to_tangent_space(v, m)); break; }
        case 20:
        {
            vec4 u = r.v4(); vec3 view = r.v3(); sampled_material m = r.mat();
            vec3 out_dir = vec3(0); bsdf_lobes lobes = bsdf_lobes(0, 0, 0, 0); float pdf = 0;
            material_bsdf_sample(u, view, m, out_dir, lobes, pdf);
            w.v3(out_dir); w.lobes(lobes); w.f(pdf);
            break;
        }
        case 21:
        {
            vec3 out_dir = r.v3(), view = r.v3(); sampled_material m = r.mat();
            bsdf_lobes lobes = bsdf_lobes(0, 0, 0, 0);
            float pdf = material_bsdf_pdf(out_dir, view, m, lobes);
            w.lobes(lobes); w.f(pdf);
            break;
        }
        case 22:
        {
            point_light pl = r.pl(); vec2 u = r.v2(); vec3 pos = r.v3();
            vec3 out_dir, color; float len = 0, pdf = 0;
            sample_point_light(pl, u, pos, out_dir, len, color, pdf);
            w.v3(out_dir); w.f(len); w.v3(color); w.f(pdf);
            break;
        }
        case 23: { point_light pl = r.pl(); vec3 pos = r.v3(); w.f(sample_point_light_pdf(pl, pos)); break; }
        case 24:
        {
            directional_light dl = r.dl(); vec2 u = r.v2();
            vec3 out_dir, color; float pdf = 0;
            sample_directional_light(dl, u, out_dir, color, pdf);
            w.v3(out_dir); w.v3(color); w.f(pdf);
            break;
        }
        case 25: { vec2 u = r.v2(); vec3 A = r.v3(), B = r.v3(), C = r.v3(); float pdf = 0; w.v3(sample_triangle_light(u, A, B, C, pdf)); w.f(pdf); break; }
        case 26: { vec3 P = r.v3(), A = r.v3(), B = r.v3(), C = r.v3(); w.f(sample_triangle_light_pdf(P, A, B, C)); break; }
        case 27: w.v3(r9g9b9e5_to_rgb(r.u())); break;
        case 28: w.u(rgb_to_r9g9b9e5(r.v3())); break;
        case 29: { sampled_material m = r.mat(); bsdf_lobes l; l.transmission = r.f(); l.diffuse = r.f(); l.dielectric_reflection = r.f(); l.metallic_reflection = r.f(); w.v3(modulate_bsdf(m, l)); break; }
        case 30: { sampled_material m = r.mat(); vec3 d = r.v3(), refl = r.v3(); w.v3(modulate_color(m, d, refl)); break; }
        case 31:
        {
            camera_data cam; memcpy(&cam, r.p, 320); r.p += 80;
            vec2 pixel = r.v2(), size = r.v2(), dof_u = r.v2();
            vec3 o, d;
            get_camera_ray(cam, pixel, size, dof_u, o, d);
            w.v3(o); w.v3(d);
            break;
        }
        case 32: w.v2(octahedral_pack(r.v3())); break;
        case 33: w.v3(inverse_srgb_correction(r.v3())); break;
        case 34: w.f(rgb_to_luminance(r.v3())); break;
        case 35: { vec3 d = r.v3(); int sx = r.i(), sy = r.i(); w.u((uint32_t)latlong_direction_to_pixel_id(d, ivec2(sx, sy))); break; }
        case 36: w.v3(uv_to_latlong_direction(r.v2())); break;
        case 37: { vec3 v = r.v3(); float rough = r.f(), u1 = r.f(), u2 = r.f(); w.v3(ggx_vndf_sample(v, rough, u1, u2)); break; }
        case 38: { point_light pl = r.pl(); vec3 d = r.v3(); w.f(get_spotlight_intensity(pl, d)); break; }
        case 39: { vec3 d = r.v3(), A = r.v3(), B = r.v3(), C = r.v3(); w.f(ray_plane_intersection_dist(d, A, B, C)); break; }
        case 40: { vec3 p = r.v3(), A = r.v3(), B = r.v3(), C = r.v3(); w.v3(get_barycentric_coords(p, A, B, C)); break; }
        case 41: w.v3(sample_sphere(r.v2())); break;
        case 42: w.v3(sample_hemisphere(r.v2())); break;
        case 43: w.f(sample_blackman_harris(r.f())); break;
        }
    }
    return 0;
}
