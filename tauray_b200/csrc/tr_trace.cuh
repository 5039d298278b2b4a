// tr_trace.cuh -- software replacement of traceRayEXT: BVH traversal + ray/triangle predicate.
//
// What the result depends on is only the predicate, never the tree:
//   * watertight Woop/Benthin/Wald (JCGT 2013) ray-triangle test in fp32 with the double fallback,
//     48-byte triangle records fetched as three 128-bit loads;
//   * a hit needs tmin < t < tmax; the smallest t wins, equal t is broken by the lower
//     (instance_id, primitive_id) -- so any conservative traversal order gives the same answer;
//   * barycentrics (u, v) = weights of vertex 1 and vertex 2 (VK hit attributes, shader/rt.glsl:43).
// Box slabs are conservative: node boxes are inflated at build time (tr_bvh.cu) so the fused
// slab arithmetic can never cull a triangle the predicate accepts.
#pragma once
#include "tr_math.cuh"
#include "tr_scene.cuh"

namespace trd {

struct RayPre
{
    int kx, ky, kz;
    float Sx, Sy, Sz;
    vec3 org;
};

TRD float sel(vec3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

TRD RayPre precompute_ray(vec3 org, vec3 dir)
{
    RayPre r;
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    r.kx = r.kz + 1; if(r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if(r.ky == 3) r.ky = 0;
    if(sel(dir, r.kz) < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    float dz = sel(dir, r.kz);
    r.Sx = sel(dir, r.kx) / dz;
    r.Sy = sel(dir, r.ky) / dz;
    r.Sz = 1.0f / dz;
    r.org = org;
    return r;
}

TRD bool intersect_triangle(const RayPre& r, vec3 v0, vec3 v1, vec3 v2, float tmin, float tmax,
                            float& t_out, float& u_out, float& v_out)
{
    vec3 A = v0 - r.org, B = v1 - r.org, C = v2 - r.org;
    float Akz = sel(A, r.kz), Bkz = sel(B, r.kz), Ckz = sel(C, r.kz);
    float Ax = fma_(-r.Sx, Akz, sel(A, r.kx));
    float Ay = fma_(-r.Sy, Akz, sel(A, r.ky));
    float Bx = fma_(-r.Sx, Bkz, sel(B, r.kx));
    float By = fma_(-r.Sy, Bkz, sel(B, r.ky));
    float Cx = fma_(-r.Sx, Ckz, sel(C, r.kx));
    float Cy = fma_(-r.Sy, Ckz, sel(C, r.ky));

    // separately rounded products: exactly antisymmetric edge functions => watertight edges
    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
    if(U == 0.0f || V == 0.0f || W == 0.0f)
    {
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if(det == 0.0f) return false;
    float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    float T = fma_(W, Cz, fma_(V, Bz, U * Az));
    float rcp = 1.0f / det;
    float t = T * rcp;
    if(!(t > tmin && t < tmax)) return false;
    t_out = t;
    u_out = V * rcp;
    v_out = W * rcp;
    return true;
}

struct Hit
{
    int instance_id;   // -1: none
    int primitive_id;
    float u, v, t;
    uint32_t flags;
};

struct TraceStats
{
    uint32_t nodes, tris;
};

// Direction components this close to zero are replaced for the SLAB test only (the triangle
// predicate uses the true direction): keeps 0 * inf out of the interval arithmetic.
TRD float safe_inv(float d)
{
    const float eps = 1e-20f;
    float a = fabsf(d) < eps ? (d < 0.0f ? -eps : eps) : d;
    return 1.0f / a;
}

#define TR_SLAB(n0x, n0y, n0z, n0w, n2lo, n2hi, tmin_, tmax_, cmin, cmax)              \
    {                                                                                      \
        float lox = fma_(n0x, idx, -oodx), hix = fma_(n0y, idx, -oodx);                    \
        float loy = fma_(n0z, idy, -oody), hiy = fma_(n0w, idy, -oody);                    \
        float loz = fma_(n2lo, idz, -oodz), hiz = fma_(n2hi, idz, -oodz);                  \
        cmin = fmaxf(fmaxf(fminf(lox, hix), fminf(loy, hiy)), fmaxf(fminf(loz, hiz), tmin_)); \
        cmax = fminf(fminf(fmaxf(lox, hix), fmaxf(loy, hiy)), fminf(fmaxf(loz, hiz), tmax_)); \
    }

// Closest hit in (tmin, tmax). `ANY`: return at the first OPAQUE hit (shadow rays); non-opaque
// hits are reported through `on_transparent` and traversal continues.
template<bool ANY, bool STATS, typename F>
TRD bool traverse(const DeviceBvh& bvh, vec3 org, vec3 dir, float tmin, float tmax, Hit& best,
                  TraceStats& st, F&& on_transparent)
{
    const RayPre rp = precompute_ray(org, dir);
    const float idx = safe_inv(dir.x), idy = safe_inv(dir.y), idz = safe_inv(dir.z);
    const float oodx = org.x * idx, oody = org.y * idy, oodz = org.z * idz;
    float best_t = tmax;
    bool found = false;
    int stack[TR_BVH_STACK];
    int sp = 0;
    int node = 0;
    const float4* __restrict__ nodes = bvh.nodes;
    const float4* __restrict__ tris = bvh.tris;

    for(;;)
    {
        if(node >= 0)
        {
            if(STATS) st.nodes++;
            const float4 n0 = __ldg(nodes + 4 * node + 0);
            const float4 n1 = __ldg(nodes + 4 * node + 1);
            const float4 n2 = __ldg(nodes + 4 * node + 2);
            const float4 n3 = __ldg(nodes + 4 * node + 3);
            float c0min, c0max, c1min, c1max;
            TR_SLAB(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, tmin, best_t, c0min, c0max);
            TR_SLAB(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, tmin, best_t, c1min, c1max);
            const bool h0 = c0min <= c0max;
            const bool h1 = c1min <= c1max;
            int r0 = __float_as_int(n3.x), r1 = __float_as_int(n3.y);
            if(h0 && h1)
            {
                if(c1min < c0min) { int t = r0; r0 = r1; r1 = t; }
                stack[sp++] = r1;
                node = r0;
            }
            else if(h0) node = r0;
            else if(h1) node = r1;
            else
            {
                if(sp == 0) break;
                node = stack[--sp];
            }
        }
        else
        {
            const uint32_t leaf = (uint32_t)~node;
            const uint32_t first = leaf >> 3;
            const uint32_t count = (leaf & 7u) + 1u;
            for(uint32_t i = 0; i < count; ++i)
            {
                if(STATS) st.tris++;
                const float4 t0 = __ldg(tris + 3 * (first + i) + 0);
                const float4 t1 = __ldg(tris + 3 * (first + i) + 1);
                const float4 t2 = __ldg(tris + 3 * (first + i) + 2);
                float t, u, v;
                if(!intersect_triangle(rp, f3(t0), f3(t1), f3(t2), tmin, tmax, t, u, v)) continue;
                const int inst = __float_as_int(t0.w), prim = __float_as_int(t1.w);
                const uint32_t flags = __float_as_uint(t2.w);
                if(ANY)
                {
                    if(flags & TR_TRI_FLAG_OPAQUE)
                    {
                        best.instance_id = inst; best.primitive_id = prim; best.u = u; best.v = v; best.t = t; best.flags = flags;
                        return true;
                    }
                    on_transparent(inst, prim, u, v, t);
                    continue;
                }
                const bool better = !found || t < best.t ||
                    (t == best.t && (inst < best.instance_id || (inst == best.instance_id && prim < best.primitive_id)));
                if(better)
                {
                    found = true;
                    best.instance_id = inst; best.primitive_id = prim; best.u = u; best.v = v; best.t = t; best.flags = flags;
                    best_t = t;
                }
            }
            if(sp == 0) break;
            node = stack[--sp];
        }
    }
    return found;
}

} // namespace trd
