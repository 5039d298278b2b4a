// glsl_shim.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// A GLSL-as-C++ shim: just enough of the GLSL 4.60 type system and built-in functions for the reference's own
// shader include files (shader/math.glsl, sampling.glsl, ggx.glsl, light.glsl, ... of /root/reference) to compile
// as C++ after the purely lexical rewrite of oracle/ref_glsl/build_ref.py (parameter qualifiers -> references,
// float literals get an `f`, array constructors -> brace lists). The shader text itself is never copied into this
// repository: build_ref.py reads it from /root/reference at build time and writes only into oracle/_ref/.
//
// Semantics follow the GLSL specification (chapter 8) with IEEE binary32 arithmetic and the host libm for the
// transcendental functions; where the specification leaves precision open (normalize, pow, sin, ...) this is one
// valid implementation, which is why the function-level parity tests compare floating-point results within 1e-5
// relative and integer results bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace glsl {

using uint = uint32_t;

template<typename T, int N> struct vec_t;

// ---- swizzles ------------------------------------------------------------------------------------
// A swizzle proxy aliases the storage of the vector it is a union member of (M components) and names K of them.
template<typename T, int M, int... I> struct swz
{
    T d[M];
    static constexpr int K = sizeof...(I);
    using V = vec_t<T, K>;
    // GLSL's implicit conversions: int -> uint, int / uint -> float (4.1.10); everything else needs a constructor
    template<typename U> static constexpr bool implicit_to = (std::is_same<T, int>::value && std::is_same<U, uint32_t>::value) ||
        (std::is_integral<T>::value && !std::is_same<T, bool>::value && std::is_same<U, float>::value);
    static constexpr bool is_prefix() { const int idx[K] = {I...}; for(int k = 0; k < K; ++k) if(idx[k] != k) return false; return true; }
    operator V() const { return get(); }
    // a swizzle naming the leading components in order is an l-value with V's own layout: usable as an inout argument
    template<typename Dummy = void, typename = std::enable_if_t<is_prefix(), Dummy>> operator V&() { return *reinterpret_cast<V*>(d); }
    template<typename U, typename = std::enable_if_t<!std::is_same<U, T>::value>> explicit(!implicit_to<U>) operator vec_t<U, K>() const { return vec_t<U, K>(get()); }
    V get() const { V v; const int idx[K] = {I...}; for(int k = 0; k < K; ++k) v.d[k] = d[idx[k]]; return v; }
    swz& set(const V& v) { const int idx[K] = {I...}; for(int k = 0; k < K; ++k) d[idx[k]] = v.d[k]; return *this; }
    swz& operator=(const V& v) { return set(v); }
    swz& operator=(const swz& o) { return set(o.get()); }
    template<int M2, int... J> swz& operator=(const swz<T, M2, J...>& o) { return set(o.get()); }
    swz& operator+=(const V& v) { return set(get() + v); }
    swz& operator-=(const V& v) { return set(get() - v); }
    swz& operator*=(const V& v) { return set(get() * v); }
    swz& operator/=(const V& v) { return set(get() / v); }
    swz& operator+=(T s) { return set(get() + s); }
    swz& operator-=(T s) { return set(get() - s); }
    swz& operator*=(T s) { return set(get() * s); }
    swz& operator/=(T s) { return set(get() / s); }
};

// GLSL constructors take any mix of scalars and vectors and flatten them, converting every component.
template<typename T, int N, typename A, typename = std::enable_if_t<std::is_arithmetic<A>::value>>
inline void glsl_append(T (&d)[N], int& k, A a) { if(k < N) d[k] = T(a); ++k; }
template<typename T, int N, typename U, int M>
inline void glsl_append(T (&d)[N], int& k, const vec_t<U, M>& v) { for(int i = 0; i < M; ++i) { if(k < N) d[k] = T(v.d[i]); ++k; } }
template<typename T, int N, typename U, int M, int... I>
inline void glsl_append(T (&d)[N], int& k, const swz<U, M, I...>& s) { glsl_append(d, k, s.get()); }
#define GLSL_CTORS(N) \
    vec_t() { for(int i = 0; i < N; ++i) d[i] = T(0); } \
    explicit vec_t(T s) { for(int i = 0; i < N; ++i) d[i] = s; } \
    vec_t(const vec_t& o) { for(int i = 0; i < N; ++i) d[i] = o.d[i]; } \
    template<typename U, int M, typename = std::enable_if_t<(M >= N) && !(M == N && std::is_same<U, T>::value)>> \
    explicit(!(M == N && ((std::is_same<U, int>::value && std::is_same<T, uint32_t>::value) || \
                          (std::is_integral<U>::value && !std::is_same<U, bool>::value && std::is_same<T, float>::value)))) \
    vec_t(const vec_t<U, M>& o) { for(int i = 0; i < N; ++i) d[i] = T(o.d[i]); } \
    template<typename A0, typename A1, typename... Rest> vec_t(const A0& a0, const A1& a1, const Rest&... rest) \
    { int k = 0; glsl_append(d, k, a0); glsl_append(d, k, a1); (glsl_append(d, k, rest), ...); } \
    vec_t& operator=(const vec_t& o) { for(int i = 0; i < N; ++i) d[i] = o.d[i]; return *this; } \
    T& operator[](int i) { return d[i]; } \
    const T& operator[](int i) const { return d[i]; }

#define GLSL_SWZ2(M, a, b, A, B) swz<T, M, A, B> a##b;
#define GLSL_SWZ3(M, a, b, c, A, B, C) swz<T, M, A, B, C> a##b##c;
#define GLSL_SWZ4(M, a, b, c, e, A, B, C, E) swz<T, M, A, B, C, E> a##b##c##e;

template<typename T> struct vec_t<T, 2>
{
    union {
        T d[2];
        struct { T x, y; };
        struct { T r, g; };
        GLSL_SWZ2(2, r, g, 0, 1) GLSL_SWZ2(2, g, r, 1, 0)
        GLSL_SWZ2(2, x, x, 0, 0) GLSL_SWZ2(2, x, y, 0, 1) GLSL_SWZ2(2, y, x, 1, 0) GLSL_SWZ2(2, y, y, 1, 1)
    };
    GLSL_CTORS(2)
};

template<typename T> struct vec_t<T, 3>
{
    union {
        T d[3];
        struct { T x, y, z; };
        struct { T r, g, b; };
        GLSL_SWZ2(3, x, y, 0, 1) GLSL_SWZ2(3, y, x, 1, 0) GLSL_SWZ2(3, x, z, 0, 2) GLSL_SWZ2(3, z, x, 2, 0)
        GLSL_SWZ2(3, y, z, 1, 2) GLSL_SWZ2(3, z, y, 2, 1)
        GLSL_SWZ3(3, x, y, z, 0, 1, 2) GLSL_SWZ3(3, y, z, x, 1, 2, 0) GLSL_SWZ3(3, z, x, y, 2, 0, 1)
        GLSL_SWZ3(3, x, z, y, 0, 2, 1) GLSL_SWZ3(3, y, x, z, 1, 0, 2) GLSL_SWZ3(3, z, y, x, 2, 1, 0)
        GLSL_SWZ3(3, r, g, b, 0, 1, 2) GLSL_SWZ2(3, r, g, 0, 1) GLSL_SWZ2(3, b, g, 2, 1)
    };
    GLSL_CTORS(3)
};

template<typename T> struct vec_t<T, 4>
{
    union {
        T d[4];
        struct { T x, y, z, w; };
        struct { T r, g, b, a; };
        GLSL_SWZ2(4, x, y, 0, 1) GLSL_SWZ2(4, y, x, 1, 0) GLSL_SWZ2(4, x, z, 0, 2) GLSL_SWZ2(4, z, w, 2, 3) GLSL_SWZ2(4, y, z, 1, 2)
        GLSL_SWZ3(4, x, y, z, 0, 1, 2) GLSL_SWZ3(4, x, y, w, 0, 1, 3) GLSL_SWZ3(4, y, z, w, 1, 2, 3) GLSL_SWZ3(4, r, g, b, 0, 1, 2)
        GLSL_SWZ3(4, y, z, x, 1, 2, 0) GLSL_SWZ3(4, z, x, y, 2, 0, 1)
        GLSL_SWZ4(4, x, y, z, w, 0, 1, 2, 3) GLSL_SWZ4(4, y, z, x, y, 1, 2, 0, 1) GLSL_SWZ4(4, w, x, y, z, 3, 0, 1, 2)
        GLSL_SWZ4(4, y, z, w, x, 1, 2, 3, 0) GLSL_SWZ4(4, z, w, x, y, 2, 3, 0, 1)
        GLSL_SWZ4(4, r, g, b, a, 0, 1, 2, 3) GLSL_SWZ2(4, r, g, 0, 1) GLSL_SWZ2(4, b, g, 2, 1)
    };
    GLSL_CTORS(4)
};

using vec2 = vec_t<float, 2>; using vec3 = vec_t<float, 3>; using vec4 = vec_t<float, 4>;
using ivec2 = vec_t<int, 2>; using ivec3 = vec_t<int, 3>; using ivec4 = vec_t<int, 4>;
using uvec2 = vec_t<uint, 2>; using uvec3 = vec_t<uint, 3>; using uvec4 = vec_t<uint, 4>;
using bvec2 = vec_t<bool, 2>; using bvec3 = vec_t<bool, 3>; using bvec4 = vec_t<bool, 4>;

// ---- operators: NON-template overloads per concrete type, so that swizzle proxies convert implicitly ----------
#define GLSL_FOR(V, N, expr) V r_; for(int i = 0; i < N; ++i) r_.d[i] = (expr); return r_;
#define GLSL_BINOP(V, S, N, op) \
    inline V operator op(const V& a, const V& b) { GLSL_FOR(V, N, a.d[i] op b.d[i]) } \
    inline V operator op(const V& a, S b) { GLSL_FOR(V, N, a.d[i] op b) } \
    inline V operator op(S a, const V& b) { GLSL_FOR(V, N, a op b.d[i]) } \
    inline V& operator op##=(V& a, const V& b) { a = a op b; return a; } \
    inline V& operator op##=(V& a, S b) { a = a op b; return a; }
#define GLSL_ARITH(V, S, N) GLSL_BINOP(V, S, N, +) GLSL_BINOP(V, S, N, -) GLSL_BINOP(V, S, N, *) GLSL_BINOP(V, S, N, /) \
    inline V operator-(const V& a) { GLSL_FOR(V, N, -a.d[i]) }
#define GLSL_INTOPS(V, S, N) GLSL_BINOP(V, S, N, %) GLSL_BINOP(V, S, N, &) GLSL_BINOP(V, S, N, |) GLSL_BINOP(V, S, N, ^) \
    GLSL_BINOP(V, S, N, <<) GLSL_BINOP(V, S, N, >>) inline V operator~(const V& a) { GLSL_FOR(V, N, ~a.d[i]) }
GLSL_ARITH(vec2, float, 2) GLSL_ARITH(vec3, float, 3) GLSL_ARITH(vec4, float, 4)
GLSL_ARITH(ivec2, int, 2) GLSL_ARITH(ivec3, int, 3) GLSL_ARITH(ivec4, int, 4)
GLSL_ARITH(uvec2, uint, 2) GLSL_ARITH(uvec3, uint, 3) GLSL_ARITH(uvec4, uint, 4)
GLSL_INTOPS(ivec2, int, 2) GLSL_INTOPS(ivec3, int, 3) GLSL_INTOPS(ivec4, int, 4)
GLSL_INTOPS(uvec2, uint, 2) GLSL_INTOPS(uvec3, uint, 3) GLSL_INTOPS(uvec4, uint, 4)

// swizzle (op) swizzle of the same component type: the natural vector type (keeps overload resolution unambiguous)
#define GLSL_SWZ_OP(op) \
    template<typename T, int M1, int... I, int M2, int... J> \
    inline auto operator op(const swz<T, M1, I...>& a, const swz<T, M2, J...>& b) { return a.get() op b.get(); } \
    template<typename T, int M1, int... I> \
    inline auto operator op(const swz<T, M1, I...>& a, const vec_t<T, (int)sizeof...(I)>& b) { return a.get() op b; } \
    template<typename T, int M1, int... I> \
    inline auto operator op(const vec_t<T, (int)sizeof...(I)>& a, const swz<T, M1, I...>& b) { return a op b.get(); }
GLSL_SWZ_OP(+) GLSL_SWZ_OP(-) GLSL_SWZ_OP(*) GLSL_SWZ_OP(/) GLSL_SWZ_OP(%) GLSL_SWZ_OP(&) GLSL_SWZ_OP(|) GLSL_SWZ_OP(^) GLSL_SWZ_OP(<<) GLSL_SWZ_OP(>>)

#define GLSL_MIXED(FV, IV, N, op) \
    inline FV operator op(const FV& a, const IV& b) { GLSL_FOR(FV, N, a.d[i] op float(b.d[i])) } \
    inline FV operator op(const IV& a, const FV& b) { GLSL_FOR(FV, N, float(a.d[i]) op b.d[i]) }
#define GLSL_MIXED_ALL(FV, IV, N) GLSL_MIXED(FV, IV, N, +) GLSL_MIXED(FV, IV, N, -) GLSL_MIXED(FV, IV, N, *) GLSL_MIXED(FV, IV, N, /)
GLSL_MIXED_ALL(vec2, ivec2, 2) GLSL_MIXED_ALL(vec3, ivec3, 3) GLSL_MIXED_ALL(vec4, ivec4, 4)
GLSL_MIXED_ALL(vec2, uvec2, 2) GLSL_MIXED_ALL(vec3, uvec3, 3) GLSL_MIXED_ALL(vec4, uvec4, 4)

// ---- scalar built-ins ----------------------------------------------------------------------------
template<typename A, typename B> using common_t = std::common_type_t<A, B>;
#define GLSL_ARITH_T(...) typename = std::enable_if_t<(std::is_arithmetic<__VA_ARGS__>::value)>
template<typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline common_t<A, B> min(A x, B y) { using C = common_t<A, B>; return C(y) < C(x) ? C(y) : C(x); }
template<typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline common_t<A, B> max(A x, B y) { using C = common_t<A, B>; return C(x) < C(y) ? C(y) : C(x); }
template<typename A, typename B, typename C3, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C3>::value>>
inline common_t<A, common_t<B, C3>> clamp(A x, B lo, C3 hi) { using C = common_t<A, common_t<B, C3>>; return glsl::min(glsl::max(C(x), C(lo)), C(hi)); }
template<typename A, typename B, typename C3, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_floating_point<C3>::value>>
inline float mix(A x, B y, C3 a) { return float(x) * (1.0f - float(a)) + float(y) * float(a); }
template<typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline float step(A edge, B x) { return float(x) < float(edge) ? 0.0f : 1.0f; }
template<typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline float pow(A x, B y) { return std::pow(float(x), float(y)); }
template<typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline float atan(A y, B x) { return std::atan2(float(y), float(x)); }
inline float atan(float x) { return std::atan(x); }
inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline int sign(int x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float round(float x) { return std::nearbyint(x); }     // GLSL: implementation's choice for .5; roundEven is allowed
inline float fract(float x) { return x - std::floor(x); }
template<typename A, GLSL_ARITH_T(A)> inline float sqrt(A x) { return std::sqrt(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float inversesqrt(A x) { return 1.0f / std::sqrt(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float sin(A x) { return std::sin(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float cos(A x) { return std::cos(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float tan(A x) { return std::tan(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float asin(A x) { return std::asin(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float acos(A x) { return std::acos(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float exp2(A x) { return std::exp2(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float log2(A x) { return std::log2(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float exp(A x) { return std::exp(float(x)); }
template<typename A, GLSL_ARITH_T(A)> inline float log(A x) { return std::log(float(x)); }
inline float fma(float a, float b, float c) { return std::fma(a, b, c); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return std::isinf(x); }
inline int bitCount(uint x) { return __builtin_popcount(x); }
inline int bitCount(int x) { return __builtin_popcount((uint)x); }
inline int findLSB(uint x) { return x == 0 ? -1 : __builtin_ctz(x); }
inline int findMSB(uint x) { return x == 0 ? -1 : 31 - __builtin_clz(x); }
inline int findLSB(int x) { return findLSB((uint)x); }
inline uint bitfieldReverse(uint x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}
inline float frexp(float x, int& e) { return std::frexp(x, &e); }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }

// ---- component-wise vector built-ins ------------------------------------------------------------
#define GLSL_MAP1(R, V, N, fn) inline R fn(const V& a) { GLSL_FOR(R, N, fn(a.d[i])) }
#define GLSL_MAP2(V, S, N, fn) \
    inline V fn(const V& a, const V& b) { GLSL_FOR(V, N, fn(a.d[i], b.d[i])) } \
    inline V fn(const V& a, S b) { GLSL_FOR(V, N, fn(a.d[i], b)) }
#define GLSL_FLOATVEC(V, B, IV, N) \
    GLSL_MAP1(V, V, N, abs) GLSL_MAP1(V, V, N, sign) GLSL_MAP1(V, V, N, floor) GLSL_MAP1(V, V, N, ceil) GLSL_MAP1(V, V, N, round) GLSL_MAP1(V, V, N, fract) \
    GLSL_MAP1(V, V, N, sqrt) GLSL_MAP1(V, V, N, inversesqrt) GLSL_MAP1(V, V, N, sin) GLSL_MAP1(V, V, N, cos) GLSL_MAP1(V, V, N, exp2) GLSL_MAP1(V, V, N, log2) \
    GLSL_MAP1(B, V, N, isnan) GLSL_MAP1(B, V, N, isinf) \
    GLSL_MAP2(V, float, N, min) GLSL_MAP2(V, float, N, max) GLSL_MAP2(V, float, N, pow) \
    inline V step(const V& e, const V& x) { GLSL_FOR(V, N, step(e.d[i], x.d[i])) } \
    inline V step(float e, const V& x) { GLSL_FOR(V, N, step(e, x.d[i])) } \
    inline V clamp(const V& x, const V& lo, const V& hi) { GLSL_FOR(V, N, clamp(x.d[i], lo.d[i], hi.d[i])) } \
    inline V clamp(const V& x, float lo, float hi) { GLSL_FOR(V, N, clamp(x.d[i], lo, hi)) } \
    inline V mix(const V& x, const V& y, const V& a) { GLSL_FOR(V, N, mix(x.d[i], y.d[i], a.d[i])) } \
    inline V mix(const V& x, const V& y, float a) { GLSL_FOR(V, N, mix(x.d[i], y.d[i], a)) } \
    inline V mix(const V& x, const V& y, const B& a) { GLSL_FOR(V, N, a.d[i] ? y.d[i] : x.d[i]) } \
    inline V fma(const V& a, const V& b, const V& c) { GLSL_FOR(V, N, fma(a.d[i], b.d[i], c.d[i])) } \
    inline V frexp(const V& x, IV& e) { GLSL_FOR(V, N, frexp(x.d[i], e.d[i])) } \
    inline float dot(const V& a, const V& b) { float s = a.d[0] * b.d[0]; for(int i = 1; i < N; ++i) s += a.d[i] * b.d[i]; return s; } \
    inline float length(const V& a) { return std::sqrt(dot(a, a)); } \
    inline float distance(const V& a, const V& b) { return length(a - b); } \
    inline V normalize(const V& a) { return a / length(a); } \
    inline V reflect(const V& I, const V& Nn) { return I - 2.0f * dot(Nn, I) * Nn; } \
    inline V refract(const V& I, const V& Nn, float eta) \
    { const float d_ = dot(Nn, I); const float k = 1.0f - eta * eta * (1.0f - d_ * d_); if(k < 0.0f) return V(0.0f); return eta * I - (eta * d_ + std::sqrt(k)) * Nn; } \
    inline B lessThan(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] < b.d[i]) } \
    inline B lessThanEqual(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] <= b.d[i]) } \
    inline B greaterThan(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] > b.d[i]) } \
    inline B greaterThanEqual(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] >= b.d[i]) } \
    inline B equal(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] == b.d[i]) }
GLSL_FLOATVEC(vec2, bvec2, ivec2, 2) GLSL_FLOATVEC(vec3, bvec3, ivec3, 3) GLSL_FLOATVEC(vec4, bvec4, ivec4, 4)

#define GLSL_INTVEC(V, S, B, IV, N) \
    GLSL_MAP2(V, S, N, min) GLSL_MAP2(V, S, N, max) \
    inline V clamp(const V& x, const V& lo, const V& hi) { GLSL_FOR(V, N, clamp(x.d[i], lo.d[i], hi.d[i])) } \
    inline V clamp(const V& x, S lo, S hi) { GLSL_FOR(V, N, clamp(x.d[i], lo, hi)) } \
    inline IV bitCount(const V& a) { GLSL_FOR(IV, N, bitCount(a.d[i])) } \
    inline B equal(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] == b.d[i]) } \
    inline B lessThan(const V& a, const V& b) { GLSL_FOR(B, N, a.d[i] < b.d[i]) }
GLSL_INTVEC(ivec2, int, bvec2, ivec2, 2) GLSL_INTVEC(ivec3, int, bvec3, ivec3, 3) GLSL_INTVEC(ivec4, int, bvec4, ivec4, 4)
GLSL_INTVEC(uvec2, uint, bvec2, ivec2, 2) GLSL_INTVEC(uvec3, uint, bvec3, ivec3, 3) GLSL_INTVEC(uvec4, uint, bvec4, ivec4, 4)
GLSL_MAP1(ivec2, ivec2, 2, abs) GLSL_MAP1(ivec3, ivec3, 3, abs) GLSL_MAP1(ivec4, ivec4, 4, abs)
GLSL_MAP1(ivec2, ivec2, 2, sign) GLSL_MAP1(ivec3, ivec3, 3, sign) GLSL_MAP1(ivec4, ivec4, 4, sign)
GLSL_MAP1(uvec2, uvec2, 2, bitfieldReverse) GLSL_MAP1(uvec3, uvec3, 3, bitfieldReverse) GLSL_MAP1(uvec4, uvec4, 4, bitfieldReverse)

#define GLSL_EQ(V, N) \
    inline bool operator==(const V& a, const V& b) { for(int i = 0; i < N; ++i) if(!(a.d[i] == b.d[i])) return false; return true; } \
    inline bool operator!=(const V& a, const V& b) { return !(a == b); }
GLSL_EQ(vec2, 2) GLSL_EQ(vec3, 3) GLSL_EQ(vec4, 4) GLSL_EQ(ivec2, 2) GLSL_EQ(ivec3, 3) GLSL_EQ(ivec4, 4) GLSL_EQ(uvec2, 2) GLSL_EQ(uvec3, 3) GLSL_EQ(uvec4, 4)
inline bool all(const bvec2& b) { return b.d[0] && b.d[1]; }
inline bool all(const bvec3& b) { return b.d[0] && b.d[1] && b.d[2]; }
inline bool all(const bvec4& b) { return b.d[0] && b.d[1] && b.d[2] && b.d[3]; }
inline bool any(const bvec2& b) { return b.d[0] || b.d[1]; }
inline bool any(const bvec3& b) { return b.d[0] || b.d[1] || b.d[2]; }
inline bool any(const bvec4& b) { return b.d[0] || b.d[1] || b.d[2] || b.d[3]; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

// ---- matrices (column-major, m[c] = column c) ---------------------------------------------------------
struct mat4;
struct mat3
{
    vec3 c[3];
    mat3() {}
    mat3(const vec3& a, const vec3& b, const vec3& e) { c[0] = a; c[1] = b; c[2] = e; }
    explicit mat3(float s) { c[0] = vec3(s, 0, 0); c[1] = vec3(0, s, 0); c[2] = vec3(0, 0, s); }
    explicit inline mat3(const mat4& m);
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4
{
    vec4 c[4];
    mat4() {}
    mat4(const vec4& a, const vec4& b, const vec4& e, const vec4& f) { c[0] = a; c[1] = b; c[2] = e; c[3] = f; }
    explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m) { for(int i = 0; i < 3; ++i) c[i] = vec3(m.c[i].x, m.c[i].y, m.c[i].z); }
inline float determinant(const mat3& m) { return dot(m.c[0], cross(m.c[1], m.c[2])); }
inline vec3 operator*(const mat3& m, const vec3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
inline vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
inline vec4 operator*(const mat4& m, const vec4& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z + m.c[3] * v.w; }
inline vec4 operator*(const vec4& v, const mat4& m) { return vec4(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2]), dot(v, m.c[3])); }
inline mat4 operator*(const mat4& a, const mat4& b) { return mat4(a * b.c[0], a * b.c[1], a * b.c[2], a * b.c[3]); }
inline mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b.c[0], a * b.c[1], a * b.c[2]); }
inline mat3 transpose(const mat3& m) { return mat3(vec3(m.c[0].x, m.c[1].x, m.c[2].x), vec3(m.c[0].y, m.c[1].y, m.c[2].y), vec3(m.c[0].z, m.c[1].z, m.c[2].z)); }

} // namespace glsl
