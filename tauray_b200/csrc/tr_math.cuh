// tr_math.cuh -- "TR math" on the device: the arithmetic definition of the GLSL built-ins and of
// shader/math.glsl that the hot path uses.
//
// GLSL leaves the precision of sin/cos/atan/pow/normalize and FMA contraction to the driver, so
// this build pins one definition: every operation is an IEEE-754 binary32 +,-,*,/,sqrt (round to
// nearest even) or an explicit __fmaf_rn, in exactly the order written. The translation unit is
// compiled with -fmad=false (no implicit contraction), default -prec-div/-prec-sqrt and no
// -use_fast_math. The CPU oracle states the same definition independently, which makes
// GPU-vs-oracle parity bit-exact.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define TRD __device__ __forceinline__

namespace trd {

TRD float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
TRD float min_(float a, float b) { return b < a ? b : a; }
TRD float max_(float a, float b) { return a < b ? b : a; }
TRD float clamp_(float x, float lo, float hi) { return min_(max_(x, lo), hi); }
TRD int clampi_(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
TRD float step_(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
TRD float mix_(float x, float y, float a) { return fma_(y, a, x * (1.0f - a)); }
TRD float fract_(float x) { return x - floorf(x); }
TRD bool isnan_(float x) { return x != x; }
TRD bool isinf_(float x) { return (__float_as_uint(x) & 0x7FFFFFFFu) == 0x7F800000u; }
TRD float exp2i_(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }
TRD float inf_() { return __uint_as_float(0x7F800000u); }

#define TR_PI 3.14159265359f
#define TR_M_1_SQRT3 0.57735026918962576451f
#define TR_INV_UINT32_MAX 2.3283064365386963e-10f

TRD void sincos_(float x, float* s, float* c)
{
    float q = floorf(fma_(x, 0.636619772f, 0.5f));
    float r = fma_(q, -1.57079601e+00f, x);
    r = fma_(q, -3.13916473e-07f, r);
    r = fma_(q, -5.39030253e-15f, r);
    int qi = (int)q;
    float z = r * r;
    float sp = fma_(z, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fma_(sp, z, -1.6666654611e-1f);
    float sr = fma_(sp * z, r, r);
    float cp = fma_(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fma_(cp, z, 4.166664568298827e-2f);
    float cr = fma_(cp * z, z, fma_(z, -0.5f, 1.0f));
    int k = qi & 3;
    float ss = (k & 1) ? cr : sr;
    float cc = (k & 1) ? sr : cr;
    *s = (k & 2) ? -ss : ss;
    *c = (k == 1 || k == 2) ? -cc : cc;
}

TRD float atan2_(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float mx = max_(ax, ay), mn = min_(ax, ay);
    float a = mx == 0.0f ? 0.0f : mn / mx;
    float base = 0.0f, t = a;
    if(a > 0.4142135679721832f)
    {
        t = (a - 1.0f) / (a + 1.0f);
        base = 0.7853981852531433f;
    }
    float z = t * t;
    float p = fma_(z, 8.05374449538e-2f, -1.38776856032e-1f);
    p = fma_(p, z, 1.99777106478e-1f);
    p = fma_(p, z, -3.33329491539e-1f);
    float r = base + fma_(p * z, t, t);
    if(ay > ax) r = 1.5707963705062866f - r;
    if(x < 0.0f) r = 3.1415927410125732f - r;
    return y < 0.0f ? -r : r;
}
TRD float sin_(float x) { float s, c; sincos_(x, &s, &c); return s; }
TRD float cos_(float x) { float s, c; sincos_(x, &s, &c); return c; }
TRD float asin_(float x) { return atan2_(x, sqrtf((1.0f - x) * (1.0f + x))); }
TRD float acos_(float x) { return atan2_(sqrtf((1.0f - x) * (1.0f + x)), x); }

TRD float log2_(float x)
{
    uint32_t u = __float_as_uint(x);
    int e = (int)(u >> 23) - 127;
    float m = __uint_as_float((u & 0x007FFFFFu) | 0x3F800000u);
    if(m > 1.41421356f) { m *= 0.5f; e += 1; }
    float f = m - 1.0f;
    float z = f * f;
    float p = fma_(f, 7.0376836292e-2f, -1.1514610310e-1f);
    p = fma_(p, f, 1.1676998740e-1f);
    p = fma_(p, f, -1.2420140846e-1f);
    p = fma_(p, f, 1.4249322787e-1f);
    p = fma_(p, f, -1.6668057665e-1f);
    p = fma_(p, f, 2.0000714765e-1f);
    p = fma_(p, f, -2.4999993993e-1f);
    p = fma_(p, f, 3.3333331174e-1f);
    float ln = fma_(p * f, z, fma_(z, -0.5f, f));
    return fma_(ln, 1.4426950408889634f, (float)e);
}
TRD float exp2_(float x)
{
    if(x > 127.99f) return inf_();
    if(x < -126.0f) return 0.0f;
    float n = floorf(x + 0.5f);
    float f = x - n;
    float p = fma_(f, 1.535336188319500e-4f, 1.339887440266574e-3f);
    p = fma_(p, f, 9.618437357674640e-3f);
    p = fma_(p, f, 5.550332471162809e-2f);
    p = fma_(p, f, 2.402264791363012e-1f);
    p = fma_(p, f, 6.931472028550421e-1f);
    float r = fma_(p, f, 1.0f);
    return r * exp2i_((int)n);
}
TRD float pow_(float x, float y)
{
    if(x == 0.0f) return y > 0.0f ? 0.0f : inf_();
    if(x < 0.0f || isnan_(x)) return __uint_as_float(0x7FC00000u);
    return exp2_(y * log2_(x));
}
TRD float pow5_(float x) { float x2 = x * x; return x2 * x2 * x; }

TRD uint32_t float_to_half_(float f)
{
    uint32_t x = __float_as_uint(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x007FFFFFu;
    int exp = (int)((x >> 23) & 0xFF);
    if(exp == 255) return sign | 0x7C00u | (mant ? 0x200u : 0u);
    int e = exp - 127 + 15;
    if(e >= 31) return sign | 0x7C00u;
    if(e <= 0)
    {
        if(e < -10) return sign;
        mant |= 0x00800000u;
        int shift = 14 - e;
        uint32_t h = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1);
        if(rem > half || (rem == half && (h & 1u))) h++;
        return sign | h;
    }
    uint32_t h = ((uint32_t)e << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1FFFu;
    if(rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return sign | h;
}
TRD float half_to_float_(uint32_t h)
{
    uint32_t sign = (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t mant = h & 0x3FFu;
    if(exp == 0)
    {
        if(mant == 0) return __uint_as_float(sign);
        float m = (float)mant * exp2i_(-24);
        return sign ? -m : m;
    }
    if(exp == 31) return __uint_as_float(sign | 0x7F800000u | (mant << 13));
    return __uint_as_float(sign | ((exp + 112u) << 23) | (mant << 13));
}

TRD int findLSB_(uint32_t x) { return x == 0 ? -1 : __ffs((int)x) - 1; }
TRD int findMSB_(uint32_t x) { return x == 0 ? -1 : 31 - __clz((int)x); }

// ---------------------------------------------------------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct uvec4 { uint32_t x, y, z, w; };

TRD vec2 V2(float x, float y) { vec2 v; v.x = x; v.y = y; return v; }
TRD vec3 V3(float x, float y, float z) { vec3 v; v.x = x; v.y = y; v.z = z; return v; }
TRD vec3 V3(float s) { return V3(s, s, s); }
TRD vec4 V4(float x, float y, float z, float w) { vec4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
TRD vec4 V4(vec3 a, float w) { return V4(a.x, a.y, a.z, w); }
TRD vec3 xyz(vec4 v) { return V3(v.x, v.y, v.z); }
TRD vec3 ld3(const float* p) { return V3(p[0], p[1], p[2]); }
TRD vec3 f3(float4 v) { return V3(v.x, v.y, v.z); }

TRD vec2 operator+(vec2 a, vec2 b) { return V2(a.x + b.x, a.y + b.y); }
TRD vec2 operator-(vec2 a, vec2 b) { return V2(a.x - b.x, a.y - b.y); }
TRD vec2 operator*(vec2 a, float s) { return V2(a.x * s, a.y * s); }
TRD vec2 operator*(float s, vec2 a) { return V2(a.x * s, a.y * s); }
TRD vec2 operator*(vec2 a, vec2 b) { return V2(a.x * b.x, a.y * b.y); }
TRD vec2 operator/(vec2 a, vec2 b) { return V2(a.x / b.x, a.y / b.y); }

TRD vec3 operator+(vec3 a, vec3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
TRD vec3 operator-(vec3 a, vec3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
TRD vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
TRD vec3 operator*(vec3 a, vec3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
TRD vec3 operator*(vec3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
TRD vec3 operator*(float s, vec3 a) { return V3(a.x * s, a.y * s, a.z * s); }
TRD vec3 operator/(vec3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
TRD vec3 operator/(vec3 a, vec3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
TRD vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
TRD vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
TRD vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
TRD vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
TRD vec4 operator/(vec4 a, float s) { return V4(a.x / s, a.y / s, a.z / s, a.w / s); }

TRD vec4 operator+(vec4 a, vec4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
TRD vec4 operator*(vec4 a, float s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); }
TRD vec4 operator*(vec4 a, vec4 b) { return V4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

TRD float dot(vec2 a, vec2 b) { return fma_(a.y, b.y, a.x * b.x); }
TRD float dot(vec3 a, vec3 b) { return fma_(a.z, b.z, fma_(a.y, b.y, a.x * b.x)); }
TRD vec3 cross(vec3 a, vec3 b)
{
    return V3(fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x)));
}
TRD float length(vec3 a) { return sqrtf(dot(a, a)); }
TRD vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
TRD float inversesqrt_(float x) { return 1.0f / sqrtf(x); }
TRD vec3 mix_(vec3 x, vec3 y, float a) { return V3(mix_(x.x, y.x, a), mix_(x.y, y.y, a), mix_(x.z, y.z, a)); }
TRD vec4 mix_(vec4 x, vec4 y, float a) { return V4(mix_(x.x, y.x, a), mix_(x.y, y.y, a), mix_(x.z, y.z, a), mix_(x.w, y.w, a)); }
TRD bool any_nan(vec3 v) { return isnan_(v.x) || isnan_(v.y) || isnan_(v.z); }
TRD vec3 reflect_(vec3 I, vec3 N)
{
    float t = 2.0f * dot(N, I);
    return V3(fma_(-t, N.x, I.x), fma_(-t, N.y, I.y), fma_(-t, N.z, I.z));
}
TRD vec3 refract_(vec3 I, vec3 N, float eta)
{
    float d = dot(N, I);
    float k = fma_(-(eta * eta), fma_(-d, d, 1.0f), 1.0f);
    if(k < 0.0f) return V3(0.0f);
    float t = fma_(eta, d, sqrtf(k));
    return V3(fma_(-t, N.x, eta * I.x), fma_(-t, N.y, eta * I.y), fma_(-t, N.z, eta * I.z));
}

struct mat3 { vec3 c0, c1, c2; };

TRD vec3 mul(const mat3& m, vec3 v)
{
    return V3(
        fma_(m.c2.x, v.z, fma_(m.c1.x, v.y, m.c0.x * v.x)),
        fma_(m.c2.y, v.z, fma_(m.c1.y, v.y, m.c0.y * v.x)),
        fma_(m.c2.z, v.z, fma_(m.c1.z, v.y, m.c0.z * v.x)));
}
TRD vec3 mul(vec3 v, const mat3& m) { return V3(dot(v, m.c0), dot(v, m.c1), dot(v, m.c2)); }
// column-major 4x4 (16 floats) times vec4
TRD vec4 mul4(const float* m, vec4 v)
{
    vec4 r;
    r.x = fma_(m[12], v.w, fma_(m[8], v.z, fma_(m[4], v.y, m[0] * v.x)));
    r.y = fma_(m[13], v.w, fma_(m[9], v.z, fma_(m[5], v.y, m[1] * v.x)));
    r.z = fma_(m[14], v.w, fma_(m[10], v.z, fma_(m[6], v.y, m[2] * v.x)));
    r.w = fma_(m[15], v.w, fma_(m[11], v.z, fma_(m[7], v.y, m[3] * v.x)));
    return r;
}
TRD mat3 upper3(const float* m)
{
    mat3 r;
    r.c0 = V3(m[0], m[1], m[2]); r.c1 = V3(m[4], m[5], m[6]); r.c2 = V3(m[8], m[9], m[10]);
    return r;
}
TRD float determinant(const mat3& m) { return dot(m.c0, cross(m.c1, m.c2)); }

// shader/math.glsl:12-31
TRD vec3 create_tangent(vec3 normal)
{
    vec3 major;
    if(fabsf(normal.x) < TR_M_1_SQRT3) major = V3(1, 0, 0);
    else if(fabsf(normal.y) < TR_M_1_SQRT3) major = V3(0, 1, 0);
    else major = V3(0, 0, 1);
    return normalize(cross(normal, major));
}
TRD mat3 create_tangent_space(vec3 normal)
{
    mat3 m;
    m.c0 = create_tangent(normal);
    m.c1 = cross(normal, m.c0);
    m.c2 = normal;
    return m;
}
// shader/math.glsl:472-478
TRD vec3 view_to_tangent_space(vec3 view, const mat3& tbn)
{
    vec3 tview = mul(-view, tbn);
    if(tview.z < 1e-5f) tview = V3(tview.x, tview.y, max_(tview.z, 1e-5f));
    return normalize(tview);
}

// shader/math.glsl:75-81
TRD uint32_t pcg(uint32_t& seed)
{
    seed = seed * 747796405u + 2891336453u;
    seed = ((seed >> ((seed >> 28) + 4)) ^ seed) * 277803737u;
    seed = (seed >> 22) ^ seed;
    return seed;
}
// shader/math.glsl:105-112
TRD uvec4 pcg4d(uvec4& s)
{
    s.x = s.x * 1664525u + 1013904223u;
    s.y = s.y * 1664525u + 1013904223u;
    s.z = s.z * 1664525u + 1013904223u;
    s.w = s.w * 1664525u + 1013904223u;
    uvec4 t = s;
    s.x += t.y * t.w; s.y += t.z * t.x; s.z += t.x * t.y; s.w += t.y * t.z;
    s.x ^= s.x >> 16; s.y ^= s.y >> 16; s.z ^= s.z >> 16; s.w ^= s.w >> 16;
    t = s;
    s.x += t.y * t.w; s.y += t.z * t.x; s.z += t.x * t.y; s.w += t.y * t.z;
    return s;
}

// shader/math.glsl:205-217
TRD vec2 sample_concentric_disk(vec2 u)
{
    vec2 uo = V2(2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f);
    vec2 abs_uo = V2(fabsf(uo.x), fabsf(uo.y));
    if(abs_uo.x < 0.0001f && abs_uo.y < 0.0001f) return V2(0, 0);
    vec2 rt = (abs_uo.x > abs_uo.y) ?
        V2(uo.x, TR_PI / 4 * (uo.y / uo.x)) :
        V2(uo.y, TR_PI / 2 - TR_PI / 4 * (uo.x / uo.y));
    float s, c;
    sincos_(rt.y, &s, &c);
    return V2(rt.x * c, rt.x * s);
}
// shader/math.glsl:219-227
TRD float sample_blackman_harris(float u)
{
    bool flip = u > 0.5f;
    u = flip ? 1 - u : u;
    float vx = -0.33518669f * pow_(u, 0.5f);
    float vy = -0.51620529f * pow_(u, 0.3333333333f);
    float vz = 1.87406934f * pow_(u, 0.25f);
    float vw = -0.66315464f * pow_(u, 0.2f);
    float s = 0.29627329f * u + vx + vy + vz + vw;
    return flip ? 1 - s : s;
}
// shader/math.glsl:229-241
TRD vec2 sample_blackman_harris_concentric_disk(vec2 u)
{
    vec2 uo = V2(2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f);
    vec2 abs_uo = V2(fabsf(uo.x), fabsf(uo.y));
    if(abs_uo.x < 0.0001f && abs_uo.y < 0.0001f) return V2(0, 0);
    vec2 rt = (abs_uo.x > abs_uo.y) ?
        V2(u.x, TR_PI / 4 * (uo.y / uo.x)) :
        V2(u.y, TR_PI / 2 - TR_PI / 4 * (uo.x / uo.y));
    float s, c;
    sincos_(rt.y, &s, &c);
    float r = 2.0f * sample_blackman_harris(rt.x) - 1.0f;
    return V2(r * c, r * s);
}
// shader/math.glsl:281-292
TRD vec2 sample_regular_polygon(vec2 u, float angle, uint32_t sides)
{
    float fs = (float)sides;
    float side = floorf(u.x * fs);
    u.x = fract_(u.x * fs);
    float side_radians = (2.0f * TR_PI) / fs;
    float a1 = side_radians * side + angle;
    float a2 = side_radians * (side + 1) + angle;
    float s1, c1, s2, c2;
    sincos_(a1, &s1, &c1);
    sincos_(a2, &s2, &c2);
    vec2 b = V2(s1, c1);
    vec2 c = V2(s2, c2);
    if(u.x + u.y > 1) u = V2(1 - u.x, 1 - u.y);
    return b * u.x + c * u.y;
}
// shader/math.glsl:294-303
TRD vec3 sample_cosine_hemisphere(vec2 u)
{
    vec2 d = sample_concentric_disk(u);
    return V3(d.x, d.y, sqrtf(max_(0, 1 - dot(d, d))));
}
TRD float pdf_cosine_hemisphere(vec3 dir) { return max_(dir.z, 0.0f) * (1.0f / TR_PI); }
// shader/math.glsl:305-327
TRD vec3 sample_sphere(vec2 u)
{
    float cos_theta = 2 * u.x - 1;
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float phi = u.y * 2 * TR_PI;
    float s, c;
    sincos_(phi, &s, &c);
    return V3(c * sin_theta, s * sin_theta, cos_theta);
}
TRD vec3 sample_hemisphere(vec2 u)
{
    float cos_theta = u.x;
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float phi = u.y * 2 * TR_PI;
    float s, c;
    sincos_(phi, &s, &c);
    return V3(c * sin_theta, s * sin_theta, cos_theta);
}
// shader/math.glsl:342-358
TRD vec3 sample_cone(vec2 u, vec3 dir, float cos_theta_min)
{
    float cos_theta = mix_(1.0f, cos_theta_min, u.x);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float phi = u.y * 2 * TR_PI;
    float s, c;
    sincos_(phi, &s, &c);
    vec3 o = mul(create_tangent_space(dir), V3(c * sin_theta, s * sin_theta, cos_theta));
    return dot(o, dir) <= cos_theta_min ? dir : o;
}
// shader/math.glsl:360-371
TRD vec3 sample_triangle_area(vec2 u, vec3 A, vec3 B, vec3 C)
{
    float alpha = u.x;
    float beta = u.y;
    if(alpha + beta > 1)
    {
        alpha = 1 - alpha;
        beta = 1 - beta;
    }
    float gamma = 1 - beta - alpha;
    return alpha * A + beta * B + gamma * C;
}
// shader/math.glsl:373-382
TRD float determinant_accurate(vec3 nA, vec3 nB, vec3 nC)
{
    float div = inversesqrt_(2.0f * fabsf(nB.x) + 2.0f);
    float e = nB.x > 0 ? div : -div;
    vec3 h = nB * div + V3(e, 0, 0);
    vec3 a = nA - 2.0f * h * dot(h, nA);
    vec3 c = nC - 2.0f * h * dot(h, nC);
    return fabsf(a.y * c.z - c.y * a.z);
}
// shader/math.glsl:385-416
TRD vec3 sample_spherical_triangle(vec2 xi, vec3 A, vec3 B, vec3 C, float& pdf)
{
    vec3 nA = normalize(A);
    vec3 nB = normalize(B);
    vec3 nC = normalize(C);

    float dAB = dot(nA, nB);
    float dBC = dot(nB, nC);
    float dAC = dot(nA, nC);

    float div = inversesqrt_(2.0f * fabsf(nB.x) + 2.0f);
    float e = nB.x > 0 ? div : -div;
    vec3 h = nB * div + V3(e, 0, 0);

    vec3 a = nA - 2.0f * h * (dAB * div + e * nA.x);
    vec3 c = nC - 2.0f * h * (dBC * div + e * nC.x);
    float G0 = fabsf(a.y * c.z - c.y * a.z);
    float G1 = dAC + dBC;
    float G2 = 1.0f + dAB;

    float solid_angle = 2.0f * atan2_(G0, G1 + G2);
    pdf = 1.0f / solid_angle;
    float chosen_split = xi.x * solid_angle * 0.5f;
    float ss, cs;
    sincos_(chosen_split, &ss, &cs);
    vec3 r = (G0 * cs - G1 * ss) * nA + G2 * ss * nC;

    vec3 Ch = 2.0f * dot(nA, r) * r / dot(r, r) - nA;
    float d = dot(Ch, nB);
    float z = 1 - xi.y + d * xi.y;
    float st = sqrtf((1.0f - z * z) / (1.0f - d * d));
    return (z - st * d) * nB + st * Ch;
}
// shader/math.glsl:419-426
TRD float spherical_triangle_solid_angle(vec3 nA, vec3 nB, vec3 nC)
{
    return 2.0f * atan2_(
        determinant_accurate(nA, nB, nC),
        1.0f + (dot(nA, nB) + (dot(nB, nC) + dot(nA, nC))));
}
// shader/math.glsl:438-444
TRD float triangle_area_pdf(vec3 p, vec3 a, vec3 b, vec3 c)
{
    vec3 normal = cross(a - b, a - c);
    float p_dist2 = dot(p, p);
    return 2.0f * p_dist2 * sqrtf(p_dist2) / fabsf(dot(normal, p));
}
// shader/math.glsl:447-453
TRD float ray_plane_intersection_dist(vec3 dir, vec3 A, vec3 B, vec3 C)
{
    vec3 n = normalize(cross(A - B, A - C));
    float w = dot(A, n);
    return fabsf(w / dot(n, dir));
}
// shader/math.glsl:455-470
TRD vec3 get_barycentric_coords(vec3 p, vec3 A, vec3 B, vec3 C)
{
    vec3 ba = B - A, ca = C - A, pa = p - A;
    float bb = dot(ba, ba);
    float bc = dot(ba, ca);
    float cc = dot(ca, ca);
    float pb = dot(pa, ba);
    float pc = dot(pa, ca);
    float denom = 1.0f / (bb * cc - bc * bc);
    vec3 bary;
    bary.y = (cc * pb - bc * pc) * denom;
    bary.z = (bb * pc - bc * pb) * denom;
    bary.x = 1.0f - bary.y - bary.z;
    return bary;
}

// shader/color.glsl:7-12
TRD vec3 inverse_srgb_correction(vec3 col)
{
    float c[3] = {col.x, col.y, col.z};
    float o[3];
#pragma unroll
    for(int i = 0; i < 3; ++i)
    {
        float low = c[i] * 0.07739938f;
        float high = pow_(fma_(c[i], 0.94786729f, 0.05213270f), 2.4f);
        o[i] = 0.04045f < c[i] ? high : low;
    }
    return V3(o[0], o[1], o[2]);
}
// shader/color.glsl:14-17
TRD float rgb_to_luminance(vec3 col) { return dot(col, V3(0.2126f, 0.7152f, 0.0722f)); }
// shader/color.glsl:19-28
TRD uint32_t rgb_to_r9g9b9e5(vec3 color)
{
    int ex0, ex1, ex2;
    frexpf(color.x, &ex0);
    frexpf(color.y, &ex1);
    frexpf(color.z, &ex2);
    int e = clampi_(ex0 > ex1 ? (ex0 > ex2 ? ex0 : ex2) : (ex1 > ex2 ? ex1 : ex2), -16, 15);
    float scale = exp2i_(-e);
    int r = clampi_((int)floorf(color.x * scale * 512.0f), 0, 511);
    int g = clampi_((int)floorf(color.y * scale * 512.0f), 0, 511);
    int b = clampi_((int)floorf(color.z * scale * 512.0f), 0, 511);
    return (uint32_t)r | ((uint32_t)g << 9) | ((uint32_t)b << 18) | ((uint32_t)(e + 16) << 27);
}
// shader/color.glsl:30-34
TRD vec3 r9g9b9e5_to_rgb(uint32_t rgbe)
{
    int r = (int)(rgbe & 0x1FFu), g = (int)((rgbe >> 9) & 0x1FFu), b = (int)((rgbe >> 18) & 0x1FFu);
    int a = (int)((rgbe >> 27) & 0x1FFu);
    float s = exp2i_(a - 16);
    return V3((float)r * (1.0f / 512.0f) * s, (float)g * (1.0f / 512.0f) * s, (float)b * (1.0f / 512.0f) * s);
}

} // namespace trd
