// ref_prelude.hpp -- TEST INFRASTRUCTURE ONLY. What a Vulkan ray-tracing device provides around the reference's
// shader code, restated for the CPU: opaque resource types, the texture unit, image load/store, the gl_* built-in
// variables and the ray-tracing built-ins. Included inside `namespace glsl { namespace <unit> {` BEFORE the rewritten
// shader text (oracle/ref_glsl/build_ref.py); ours, not the reference's.
#pragma once

struct sampler2D
{
    const void* data = nullptr;
    int width = 0, height = 0;
    int format = 0;            // TR_TEXTURE_*: 0 rgba8 unorm, 1 rgba32f, 2 rgba16f
    int clamp_y = 0;           // envmap_sampler: repeat in x, clamp-to-edge in y (src/scene_stage.cc:259-265)
};
struct image2DArray { float* data = nullptr; int width = 0, height = 0, layers = 0; };
struct iimage2DArray { int* data = nullptr; int width = 0, height = 0, layers = 0; };
struct accelerationStructureEXT { int unused = 0; };

static thread_local uvec3 gl_LaunchIDEXT, gl_LaunchSizeEXT, gl_GlobalInvocationID;
static thread_local vec3 gl_WorldRayOriginEXT, gl_WorldRayDirectionEXT;
static thread_local int gl_InstanceCustomIndexEXT = 0, gl_GeometryIndexEXT = 0, gl_PrimitiveID = 0;
static thread_local float gl_HitTEXT = 0.0f;
static thread_local vec2 attribs;                       // hitAttributeEXT
static thread_local int rt_anyhit_action = 0;           // 0 accept, 1 ignoreIntersectionEXT, 2 terminateRayEXT
static thread_local float rt_reported_t = -1.0f;        // reportIntersectionEXT
static const uint gl_RayFlagsNoneEXT = 0u, gl_RayFlagsOpaqueEXT = 1u, gl_RayFlagsTerminateOnFirstHitEXT = 4u;
#define nonuniformEXT(x) (x)
#define ignoreIntersectionEXT do { rt_anyhit_action = 1; return; } while(0)
#define terminateRayEXT do { rt_anyhit_action = 2; return; } while(0)

inline float half_bits_to_float(uint h)
{
    const uint sign = (h >> 15) & 1u, ex = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
    float v;
    if(ex == 0) v = std::ldexp((float)man, -24);
    else if(ex == 31) v = man ? NAN : INFINITY;
    else v = std::ldexp((float)(man | 0x400u), (int)ex - 25);
    return sign ? -v : v;
}
inline uint float_to_half_bits(float f)      // round to nearest even
{
    uint u = floatBitsToUint(f);
    const uint sign = (u >> 16) & 0x8000u;
    u &= 0x7FFFFFFFu;
    if(u >= 0x7F800000u) return sign | (u > 0x7F800000u ? 0x7E00u : 0x7C00u);
    if(u >= 0x477FF000u) return sign | 0x7C00u;
    if(u < 0x33000001u) return sign;
    int e = (int)(u >> 23) - 127;
    uint m = (u & 0x7FFFFFu) | 0x800000u;
    int shift = e < -14 ? 13 + (-14 - e) : 13;
    uint half_m = m >> shift;
    const uint rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if(rem > halfway || (rem == halfway && (half_m & 1u))) half_m++;
    if(e < -14) return sign | half_m;
    return sign | (((uint)(e + 15) << 10) + (half_m - 0x400u));
}
inline vec2 unpackHalf2x16(uint v) { return vec2(half_bits_to_float(v & 0xFFFFu), half_bits_to_float(v >> 16)); }
inline uint packHalf2x16(vec2 v) { return float_to_half_bits(v.x) | (float_to_half_bits(v.y) << 16); }

// ---- texture unit: VK_FILTER_LINEAR, LOD 0, repeat (envmap: clamp-to-edge in y), unnormalised texel centres at +0.5
inline vec4 fetch_texel(const sampler2D& t, int x, int y)
{
    const size_t i = ((size_t)y * t.width + x) * 4;
    if(t.format == 1) { const float* f = (const float*)t.data; return vec4(f[i], f[i + 1], f[i + 2], f[i + 3]); }
    if(t.format == 2)
    {
        const uint16_t* h = (const uint16_t*)t.data;
        return vec4(half_bits_to_float(h[i]), half_bits_to_float(h[i + 1]), half_bits_to_float(h[i + 2]), half_bits_to_float(h[i + 3]));
    }
    const uint8_t* b = (const uint8_t*)t.data;
    return vec4(b[i] / 255.0f, b[i + 1] / 255.0f, b[i + 2] / 255.0f, b[i + 3] / 255.0f);
}
inline vec4 texture(const sampler2D& t, vec2 uv)
{
    float x = uv.x * (float)t.width - 0.5f, y = uv.y * (float)t.height - 0.5f;
    if(!(std::fabs(x) < 1e9f)) x = 0.0f;
    if(!(std::fabs(y) < 1e9f)) y = 0.0f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const float fx = x - fx0, fy = y - fy0;
    const int w = t.width, h = t.height;
    int x0 = (int)fx0 % w; if(x0 < 0) x0 += w;
    const int x1 = x0 + 1 == w ? 0 : x0 + 1;
    int y0, y1;
    if(t.clamp_y)
    {
        y0 = (int)fy0; y1 = y0 + 1;
        y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0);
        y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
    }
    else
    {
        y0 = (int)fy0 % h; if(y0 < 0) y0 += h;
        y1 = y0 + 1 == h ? 0 : y0 + 1;
    }
    const vec4 top = mix(fetch_texel(t, x0, y0), fetch_texel(t, x1, y0), fx);
    const vec4 bot = mix(fetch_texel(t, x0, y1), fetch_texel(t, x1, y1), fx);
    return mix(top, bot, fy);
}
inline ivec2 textureSize(const sampler2D& t, int) { return ivec2(t.width, t.height); }
inline vec4 texelFetch(const sampler2D& t, ivec2 p, int) { return fetch_texel(t, p.x, p.y); }

inline vec4 imageLoad(const image2DArray& img, ivec3 p)
{
    const float* f = img.data + (((size_t)p.z * img.height + p.y) * img.width + p.x) * 4;
    return vec4(f[0], f[1], f[2], f[3]);
}
inline void imageStore(image2DArray& img, ivec3 p, vec4 v)
{
    if(!img.data) return;
    float* f = img.data + (((size_t)p.z * img.height + p.y) * img.width + p.x) * 4;
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
inline ivec4 imageLoad(const iimage2DArray& img, ivec3 p) { return ivec4(img.data[((size_t)p.z * img.height + p.y) * img.width + p.x], 0, 0, 0); }
inline void imageStore(iimage2DArray& img, ivec3 p, ivec4 v) { if(img.data) img.data[((size_t)p.z * img.height + p.y) * img.width + p.x] = v.x; }

// ray-tracing built-ins, defined after the shader text (they call the shaders' own entry points)
void traceRayEXT(const accelerationStructureEXT& tlas, uint ray_flags, uint cull_mask, uint sbt_offset, uint sbt_stride, uint miss_index,
                 vec3 origin, float tmin, vec3 direction, float tmax, int payload_location);
inline bool reportIntersectionEXT(float t, uint) { rt_reported_t = t; return true; }
