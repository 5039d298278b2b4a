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

// Per-ray constants of the watertight ray-triangle test (Woop, Benthin, Wald 2013): kz = the dominant
// direction axis, (kx, ky) = the two axes after it in cyclic order, S = the shear that maps the ray onto +z.
// The paper (and the oracle) additionally swap kx and ky when dir[kz] < 0 to keep the winding; that swap
// negates U, V, W, det and T exactly (every operation is sign-symmetric), so t = T / det and the barycentrics
// V / det, W / det come out bit-identical without it. Leaving it out makes the vertex permutation a pure
// rotation by kz: two selects per component on two per-ray predicates instead of three-way selects on three
// independent indices.
struct RayPre
{
    int kz;
    float Sx, Sy, Sz;
    vec3 org;
};

TRD float sel(vec3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

TRD RayPre precompute_ray(vec3 org, vec3 dir)
{
    RayPre r;
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    int kx = r.kz + 1; if(kx == 3) kx = 0;
    int ky = kx + 1; if(ky == 3) ky = 0;
    float dz = sel(dir, r.kz);
    r.Sx = sel(dir, kx) / dz;
    r.Sy = sel(dir, ky) / dz;
    r.Sz = 1.0f / dz;
    r.org = org;
    return r;
}

// (v[kx], v[ky], v[kz]) for kx = kz + 1, ky = kz + 2 (mod 3)
TRD vec3 rotate_to_kz(vec3 v, bool kz0, bool kz1)
{
    vec3 r;
    r.x = kz0 ? v.y : (kz1 ? v.z : v.x);
    r.y = kz0 ? v.z : (kz1 ? v.x : v.y);
    r.z = kz0 ? v.x : (kz1 ? v.y : v.z);
    return r;
}

TRD bool intersect_triangle(const RayPre& r, vec3 v0, vec3 v1, vec3 v2, float tmin, float tmax,
                            float& t_out, float& u_out, float& v_out)
{
    const bool kz0 = r.kz == 0, kz1 = r.kz == 1;
    const vec3 A = rotate_to_kz(v0 - r.org, kz0, kz1), B = rotate_to_kz(v1 - r.org, kz0, kz1), C = rotate_to_kz(v2 - r.org, kz0, kz1);
    float Ax = fma_(-r.Sx, A.z, A.x);
    float Ay = fma_(-r.Sy, A.z, A.y);
    float Bx = fma_(-r.Sx, B.z, B.x);
    float By = fma_(-r.Sy, B.z, B.y);
    float Cx = fma_(-r.Sx, C.z, C.x);
    float Cy = fma_(-r.Sy, C.z, C.y);

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
    float Az = r.Sz * A.z, Bz = r.Sz * B.z, Cz = r.Sz * C.z;
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

// Per-ray constants of the wide-node slab test.
struct RaySlab
{
    float idx, idy, idz;     // 1 / dir (clamped away from 0)
    float ox, oy, oz;        // ray origin
    uint32_t octinv;         // bit k set when the direction component is >= 0 (x = 4, y = 2, z = 1)
    uint32_t sel0;           // PRMT selector of the child processed first in a half, see qbyte_to_float
};

TRD RaySlab make_ray_slab(vec3 org, vec3 dir)
{
    RaySlab r;
    r.idx = safe_inv(dir.x); r.idy = safe_inv(dir.y); r.idz = safe_inv(dir.z);
    r.ox = org.x; r.oy = org.y; r.oz = org.z;
    r.octinv = (dir.x >= 0.0f ? 4u : 0u) | (dir.y >= 0.0f ? 2u : 0u) | (dir.z >= 0.0f ? 1u : 0u);
    r.sel0 = 0x7044u | ((r.octinv & 3u) << 8);
    return r;
}

// Quantised bounds are stored as 0x80 | q with q in [0, 127]: one PRMT drops the byte into bits
// 16..23 of 0x3F000000, giving the float 1 + q/128 exactly -- no integer->float conversion (I2F runs
// at a fraction of the FMA rate). The slab then is t = (1 + q/128) * (128 a) + (b - 128 a).
// The selector is a register: nibble 2 names the byte, and the byte taken for the K-th processed child
// of a half is K ^ (octinv & 3), so that children are tested in the ray's octant order and the hit bit
// of the K-th test is the constant 1 << K.
TRD float qbyte_to_float(uint32_t w, uint32_t sel)
{
    return __uint_as_float(__byte_perm(w, 0x3F000000u, sel));
}

// Two independent IEEE fp32 FMAs in one issue slot (Blackwell FFMA2): (d0, d1) = (a0, a1) * b + c.
// Each half rounds exactly like fmaf, so this changes the instruction count only.
TRD void fma2_bcast(float& d0, float& d1, float a0, float a1, float b, float c)
{
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %4};\n\t"
        "mov.b64 rc, {%5, %5};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}

// One 80-byte compressed wide node (tr_scene.cuh): five 128-bit loads, eight quantised child boxes.
// Returns the node group G = (child_base, inner-hit bits [31:24] | imask [7:0]) and the triangle
// group Gt = (triangle_base, leaf-hit bits [7:0] | leaf slots [15:8] | two-triangle slots [23:16]).
// Hit bits are in octant order, position (slot ^ octinv): the build assigns near-to-far slots per
// octant, so taking the highest set bit first visits children roughly front to back without sorting.
// The node stores its inner-child mask already permuted into that order for octants 0..3; octant o >= 4
// is the bit reversal of octant o ^ 7. Empty slots carry an inverted box and can never be hit.
TRD void intersect_node8_regs(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const RaySlab& r,
                              float tmin, float tmax, uint2& G, uint2& Gt);

TRD void intersect_node8(const float4* __restrict__ nodes, uint32_t node_index, const RaySlab& r,
                         float tmin, float tmax, uint2& G, uint2& Gt)
{
    const float4* n = nodes + TR_BVH_NODE_VEC4 * (size_t)node_index;
    intersect_node8_regs(__ldg(n + 0), __ldg(n + 1), __ldg(n + 2), __ldg(n + 3), __ldg(n + 4), r, tmin, tmax, G, Gt);
}

// The same from a generic pointer: the node may live in shared memory (top of the tree staged by TMA) or in global memory.
TRD void intersect_node8_at(const float4* n, const RaySlab& r, float tmin, float tmax, uint2& G, uint2& Gt)
{
    intersect_node8_regs(n[0], n[1], n[2], n[3], n[4], r, tmin, tmax, G, Gt);
}

// NEAREST: also reports, in `near_bit`, the hit bit (24..31 in G.y) of the INNER child with the smallest entry distance
// (0 = none): the closest-hit kernel descends into it first and leaves the rest of the group in octant order. The key
// packs the entry distance (a non-negative float orders like its bits) with the test index in the low 3 bits.
#ifndef TR_SIGNBITS
#define TR_SIGNBITS 0
#endif
template<bool NEAREST>
TRD void intersect_node8_core(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const RaySlab& r,
                              float tmin, float tmax, uint2& G, uint2& Gt, uint32_t& near_bit)
{
    const uint32_t e = __float_as_uint(n0.w);
    // 128 * 2^e: the exponent byte already includes the factor (tr_bvh.cu writes e + 7)
    const float sx = __uint_as_float((e & 0xFFu) << 23);
    const float sy = __uint_as_float(((e >> 8) & 0xFFu) << 23);
    const float sz = __uint_as_float(((e >> 16) & 0xFFu) << 23);
    const float ax = sx * r.idx, ay = sy * r.idy, az = sz * r.idz;
    const float bx = (n0.x - r.ox) * r.idx - ax, by = (n0.y - r.oy) * r.idy - ay, bz = (n0.z - r.oz) * r.idz - az;
    const uint32_t sel0 = r.sel0, sel1 = sel0 ^ 0x100u, sel2 = sel0 ^ 0x200u, sel3 = sel0 ^ 0x300u;
    uint32_t acc[2] = {0u, 0u};
    uint32_t near_key[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
    for(int half = 0; half < 2; ++half)
    {
        const uint32_t qlox = __float_as_uint(half ? n2.y : n2.x), qloy = __float_as_uint(half ? n2.w : n2.z);
        const uint32_t qloz = __float_as_uint(half ? n3.y : n3.x), qhix = __float_as_uint(half ? n3.w : n3.z);
        const uint32_t qhiy = __float_as_uint(half ? n4.y : n4.x), qhiz = __float_as_uint(half ? n4.w : n4.z);
        const uint32_t nx = r.idx < 0.0f ? qhix : qlox, fx = r.idx < 0.0f ? qlox : qhix;
        const uint32_t ny = r.idy < 0.0f ? qhiy : qloy, fy = r.idy < 0.0f ? qloy : qhiy;
        const uint32_t nz = r.idz < 0.0f ? qhiz : qloz, fz = r.idz < 0.0f ? qloz : qhiz;
#define TR_CHILD(K, SEL)                                                                                   \
        {                                                                                                  \
            float tnx, tfx, tny, tfy, tnz, tfz;                                                            \
            fma2_bcast(tnx, tfx, qbyte_to_float(nx, SEL), qbyte_to_float(fx, SEL), ax, bx);      \
            fma2_bcast(tny, tfy, qbyte_to_float(ny, SEL), qbyte_to_float(fy, SEL), ay, by);      \
            fma2_bcast(tnz, tfz, qbyte_to_float(nz, SEL), qbyte_to_float(fz, SEL), az, bz);      \
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                   \
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));                                   \
            if(TR_SIGNBITS && !NEAREST) acc[half] = __funnelshift_l(__float_as_uint(__fsub_rn(cmax, cmin)), acc[half], 1); \
            else if(cmin <= cmax) acc[half] |= 1u << K;                                                    \
            if(NEAREST) near_key[half] = min(near_key[half], cmin <= cmax ? ((__float_as_uint(cmin) & ~3u) | (uint32_t)K) : 0xFFFFFFFFu); \
        }
        // TR_SIGNBITS: the hit bit is the (inverted) sign of cmax - cmin, shifted into the mask by a funnel shift -- one
        // FADD (FMA pipe) + one SHF instead of FSETP + SEL/IADD on the ALU pipe, which is the pipe that bounds the kernel.
        // A node that holds an acceptable hit has cmin < t <= cmax, so the difference is > 0 and cannot round below +0.
        if(TR_SIGNBITS && !NEAREST) { TR_CHILD(3, sel3) TR_CHILD(2, sel2) TR_CHILD(1, sel1) TR_CHILD(0, sel0) acc[half] = ~acc[half] & 0xFu; }
        else { TR_CHILD(0, sel0) TR_CHILD(1, sel1) TR_CHILD(2, sel2) TR_CHILD(3, sel3) }
#undef TR_CHILD
    }
    const uint32_t hi_octant = r.octinv & 4u;
    const uint32_t hits = (acc[0] << hi_octant) | (acc[1] << (4u - hi_octant));
    const uint32_t perm_word = __float_as_uint(n1.z);
    uint32_t imask_oct = (perm_word >> (8u * ((hi_octant ? ~r.octinv : r.octinv) & 3u))) & 0xFFu;
    if(hi_octant) imask_oct = __brev(imask_oct) >> 24;
    G.x = __float_as_uint(n1.x);
    G.y = ((hits & imask_oct) << 24) | (perm_word & 0xFFu);
    if(NEAREST)
    {
        // nearest hit of each half -> its position in `hits`; leaf slots are not descended into, so only an inner child counts
        const uint32_t p0 = (near_key[0] & 3u) + hi_octant, p1 = (near_key[1] & 3u) + (4u - hi_octant);
        const bool v0 = near_key[0] != 0xFFFFFFFFu && ((imask_oct >> p0) & 1u), v1 = near_key[1] != 0xFFFFFFFFu && ((imask_oct >> p1) & 1u);
        const bool first0 = v0 && (!v1 || (near_key[0] & ~3u) <= (near_key[1] & ~3u));
        near_bit = first0 ? 24u + p0 : (v1 ? 24u + p1 : 0u);
    }
    Gt.x = __float_as_uint(n1.y);
    Gt.y = (hits & ~imask_oct) | (__float_as_uint(n1.w) << 8);
}

TRD void intersect_node8_regs(const float4 n0, const float4 n1, const float4 n2, const float4 n3, const float4 n4, const RaySlab& r,
                              float tmin, float tmax, uint2& G, uint2& Gt)
{
    uint32_t unused;
    intersect_node8_core<false>(n0, n1, n2, n3, n4, r, tmin, tmax, G, Gt, unused);
}

TRD void intersect_node8_nearest(const float4* __restrict__ nodes, uint32_t node_index, const RaySlab& r,
                                 float tmin, float tmax, uint2& G, uint2& Gt, uint32_t& near_bit)
{
    const float4* n = nodes + TR_BVH_NODE_VEC4 * (size_t)node_index;
    intersect_node8_core<true>(__ldg(n + 0), __ldg(n + 1), __ldg(n + 2), __ldg(n + 3), __ldg(n + 4), r, tmin, tmax, G, Gt, near_bit);
}

// Takes the next triangle out of a triangle group: the leaf slot with the highest octant-order hit
// bit, first triangle first. Triangles of a node are stored in slot order, one or two per leaf
// slot, so the index of a slot's first triangle is the number of leaf slots below it plus the number of
// two-triangle slots below it. Bit 31 of Gt.x marks "second triangle of the top slot pending".
TRD uint32_t pop_triangle(uint2& Gt, uint32_t octinv)
{
    const uint32_t p = 31u - (uint32_t)__clz((int)(Gt.y & 0xFFu));
    const uint32_t s = p ^ octinv;
    const uint32_t below = ~(0xFFFFFFFFu << s);
    const uint32_t second = Gt.x >> 31;
    const uint32_t index = (Gt.x & 0x7FFFFFFFu) + __popc((Gt.y >> 8) & below) + __popc((Gt.y >> 16) & below) + second;
    if(((Gt.y >> (16u + s)) & 1u) > second) Gt.x |= 0x80000000u;
    else { Gt.x &= 0x7FFFFFFFu; Gt.y &= ~(1u << p); }
    return index;
}

// Index of the inner child behind hit bit `bit` (24..31) of group G.
TRD uint32_t child_of_bit(const uint2& G, uint32_t bit, uint32_t octinv)
{
    const uint32_t slot = (bit - 24u) ^ octinv;
    return G.x + __popc(G.y & 0xFFu & ((1u << slot) - 1u));
}

// Straightforward per-thread traversal (megakernel, feature tracer, test hook): closest hit in
// (tmin, tmax). `ANY`: return at the first OPAQUE hit (shadow rays); non-opaque hits are reported
// through `on_transparent` and traversal continues. The persistent wavefront kernel (tr_render.cu)
// runs the same node/triangle code with dynamic ray fetch.
template<bool ANY, bool STATS, typename F>
TRD bool traverse(const DeviceBvh& bvh, vec3 org, vec3 dir, float tmin, float tmax, Hit& best,
                  TraceStats& st, F&& on_transparent)
{
    const RayPre rp = precompute_ray(org, dir);
    const RaySlab rs = make_ray_slab(org, dir);
    float best_t = tmax;
    bool found = false;
    uint2 stack[TR_BVH_STACK];
    int sp = 0;
    const float4* __restrict__ tris = bvh.tris;
    uint2 G, Gt;
    if(STATS) st.nodes++;
    intersect_node8(bvh.nodes, 0, rs, tmin, best_t, G, Gt);
    for(;;)
    {
        while(Gt.y & 0xFFu)
        {
            const uint32_t tri_index = pop_triangle(Gt, rs.octinv);
            if(STATS) st.tris++;
            const float4* tp = tris + 3 * (size_t)tri_index;
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
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
        if((G.y & 0xFF000000u) == 0u)
        {
            if(sp == 0) break;
            G = stack[--sp];
        }
        const uint32_t bit = 31u - (uint32_t)__clz((int)G.y);
        G.y &= ~(1u << bit);
        const uint32_t child = child_of_bit(G, bit, rs.octinv);
        if(G.y & 0xFF000000u) stack[sp++] = G;
        if(STATS) st.nodes++;
        intersect_node8(bvh.nodes, child, rs, tmin, best_t, G, Gt);
    }
    return found;
}

} // namespace trd
