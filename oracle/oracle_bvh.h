// oracle_bvh.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Stand-in for the driver's acceleration structure behind traceRayEXT
// (src/acceleration_structure.cc:66-269,296-443; call sites shader/path_tracer.glsl:35-52,387-403).
// A top-down binned-SAH BVH over world-space triangles -- deliberately a DIFFERENT tree from the
// product's LBVH -- traversed with a conservative slab test. What is pinned is the
// ray/triangle predicate, not the tree:
//   * watertight Woop/Benthin/Wald (JCGT 2013) intersection in fp32 with the double fallback;
//   * a hit needs tmin < t < tmax; of several hits the smallest t wins and equal t is broken by
//     the lower (instance_id, primitive_id), so the answer does not depend on traversal order;
//   * barycentrics = (V/det, W/det) = weights of vertex 1 and 2, as VK hit attributes
//     (shader/rt.glsl:43).
#pragma once
#include "oracle_math.h"
#include <algorithm>
#include <vector>

namespace orc {

struct tri_record
{
    vec3 v0, v1, v2;
    int32_t instance_id;
    int32_t primitive_id;
    uint32_t opaque;
};

struct ray_pre
{
    int kx, ky, kz;
    float Sx, Sy, Sz;
    vec3 org;
};

static inline float comp(vec3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

static inline ray_pre precompute_ray(vec3 org, vec3 dir)
{
    ray_pre r;
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    r.kx = r.kz + 1; if(r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if(r.ky == 3) r.ky = 0;
    if(comp(dir, r.kz) < 0.0f) std::swap(r.kx, r.ky);
    float dz = comp(dir, r.kz);
    r.Sx = comp(dir, r.kx) / dz;
    r.Sy = comp(dir, r.ky) / dz;
    r.Sz = 1.0f / dz;
    r.org = org;
    return r;
}

// Returns true and fills t,u,v when the ray meets the triangle with tmin < t < tmax.
static inline bool intersect_triangle(
    const ray_pre& r, vec3 v0, vec3 v1, vec3 v2, float tmin, float tmax,
    float& t_out, float& u_out, float& v_out
){
    vec3 A = v0 - r.org, B = v1 - r.org, C = v2 - r.org;
    float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    float Ax = fma_(-r.Sx, Akz, comp(A, r.kx));
    float Ay = fma_(-r.Sy, Akz, comp(A, r.ky));
    float Bx = fma_(-r.Sx, Bkz, comp(B, r.kx));
    float By = fma_(-r.Sy, Bkz, comp(B, r.ky));
    float Cx = fma_(-r.Sx, Ckz, comp(C, r.kx));
    float Cy = fma_(-r.Sy, Ckz, comp(C, r.ky));

    // Edge functions: separately rounded products keep them exactly antisymmetric, which is
    // what makes shared edges watertight.
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if(U == 0.0f || V == 0.0f || W == 0.0f)
    {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
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

struct hit_record
{
    int32_t instance_id = -1;
    int32_t primitive_id = -1;
    float u = 0, v = 0;
    float t = 0;
};

struct bvh_node
{
    float lo[3], hi[3];
    uint32_t left;      // inner: index of left child (right = left+1); leaf: first triangle
    uint32_t count;     // 0 = inner
};

class bvh
{
public:
    std::vector<tri_record> tris; // reordered by build
    std::vector<bvh_node> nodes;
    mutable uint64_t dummy = 0;

    void build(std::vector<tri_record>&& input)
    {
        tris = std::move(input);
        nodes.clear();
        size_t n = tris.size();
        if(n == 0) return;
        float maxabs = 0.0f;
        cent.resize(n); blo.resize(n); bhi.resize(n);
        for(size_t i = 0; i < n; ++i)
        {
            const tri_record& t = tris[i];
            vec3 lo = V3(std::min({t.v0.x, t.v1.x, t.v2.x}), std::min({t.v0.y, t.v1.y, t.v2.y}), std::min({t.v0.z, t.v1.z, t.v2.z}));
            vec3 hi = V3(std::max({t.v0.x, t.v1.x, t.v2.x}), std::max({t.v0.y, t.v1.y, t.v2.y}), std::max({t.v0.z, t.v1.z, t.v2.z}));
            blo[i] = lo; bhi[i] = hi;
            cent[i] = V3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
            maxabs = std::max({maxabs, fabsf(lo.x), fabsf(lo.y), fabsf(lo.z), fabsf(hi.x), fabsf(hi.y), fabsf(hi.z)});
        }
        eps = maxabs * (1.0f / 65536.0f);
        order.resize(n);
        for(size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
        nodes.reserve(2 * n);
        nodes.push_back({});
        build_node(0, 0, (uint32_t)n);
        std::vector<tri_record> sorted(n);
        for(size_t i = 0; i < n; ++i) sorted[i] = tris[order[i]];
        tris.swap(sorted);
        cent.clear(); blo.clear(); bhi.clear(); order.clear();
        cent.shrink_to_fit(); blo.shrink_to_fit(); bhi.shrink_to_fit(); order.shrink_to_fit();
    }

    // Closest hit with the order-independent tie-break. `skip_*`: a triangle to ignore
    // (used by the in-t-order any-hit emulation). Counters are optional.
    bool closest(vec3 org, vec3 dir, float tmin, float tmax, hit_record& out,
                 uint64_t* node_visits = nullptr, uint64_t* tri_tests = nullptr) const
    {
        if(nodes.empty()) return false;
        ray_pre rp = precompute_ray(org, dir);
        vec3 idir = V3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        bool found = false;
        float best_t = tmax;
        uint32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while(sp)
        {
            const bvh_node& nd = nodes[stack[--sp]];
            if(node_visits) ++*node_visits;
            if(!slab(nd, org, dir, idir, tmin, best_t)) continue;
            if(nd.count == 0)
            {
                stack[sp++] = nd.left;
                stack[sp++] = nd.left + 1;
                continue;
            }
            for(uint32_t i = nd.left; i < nd.left + nd.count; ++i)
            {
                const tri_record& tr = tris[i];
                float t, u, v;
                if(tri_tests) ++*tri_tests;
                // `<= best` must stay reachable for the tie-break: test against tmax, compare after.
                if(!intersect_triangle(rp, tr.v0, tr.v1, tr.v2, tmin, tmax, t, u, v)) continue;
                bool better = !found || t < out.t ||
                    (t == out.t && (tr.instance_id < out.instance_id ||
                        (tr.instance_id == out.instance_id && tr.primitive_id < out.primitive_id)));
                if(better)
                {
                    found = true;
                    out.t = t; out.u = u; out.v = v;
                    out.instance_id = tr.instance_id;
                    out.primitive_id = tr.primitive_id;
                    best_t = t;
                }
            }
        }
        return found;
    }

    // Visits every triangle hit with tmin < t < tmax, in no particular order; the callback
    // returns false to stop. Used for shadow rays.
    template<typename F>
    void any(vec3 org, vec3 dir, float tmin, float tmax, F&& on_hit,
             uint64_t* node_visits = nullptr, uint64_t* tri_tests = nullptr) const
    {
        if(nodes.empty()) return;
        ray_pre rp = precompute_ray(org, dir);
        vec3 idir = V3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        uint32_t stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while(sp)
        {
            const bvh_node& nd = nodes[stack[--sp]];
            if(node_visits) ++*node_visits;
            if(!slab(nd, org, dir, idir, tmin, tmax)) continue;
            if(nd.count == 0)
            {
                stack[sp++] = nd.left;
                stack[sp++] = nd.left + 1;
                continue;
            }
            for(uint32_t i = nd.left; i < nd.left + nd.count; ++i)
            {
                const tri_record& tr = tris[i];
                float t, u, v;
                if(tri_tests) ++*tri_tests;
                if(!intersect_triangle(rp, tr.v0, tr.v1, tr.v2, tmin, tmax, t, u, v)) continue;
                if(!on_hit(tr, t, u, v)) return;
            }
        }
    }

private:
    std::vector<vec3> cent, blo, bhi;
    std::vector<uint32_t> order;
    float eps = 0.0f;

    // Conservative: boxes are inflated by eps at build time, and the interval test has slack.
    static bool slab(const bvh_node& nd, vec3 o, vec3 d, vec3 idir, float tmin, float tmax)
    {
        float t0 = tmin, t1 = tmax;
        const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z}, ii[3] = {idir.x, idir.y, idir.z};
        for(int a = 0; a < 3; ++a)
        {
            if(dd[a] == 0.0f)
            {
                if(oo[a] < nd.lo[a] || oo[a] > nd.hi[a]) return false;
                continue;
            }
            float ta = (nd.lo[a] - oo[a]) * ii[a];
            float tb = (nd.hi[a] - oo[a]) * ii[a];
            if(ta > tb) std::swap(ta, tb);
            ta -= fabsf(ta) * 4e-7f;
            tb += fabsf(tb) * 4e-7f;
            if(ta > t0) t0 = ta;
            if(tb < t1) t1 = tb;
        }
        return t0 <= t1;
    }

    void build_node(uint32_t node_index, uint32_t first, uint32_t last)
    {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for(uint32_t i = first; i < last; ++i)
        {
            uint32_t p = order[i];
            const float l[3] = {blo[p].x, blo[p].y, blo[p].z}, h[3] = {bhi[p].x, bhi[p].y, bhi[p].z};
            const float c[3] = {cent[p].x, cent[p].y, cent[p].z};
            for(int a = 0; a < 3; ++a)
            {
                lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]);
                clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]);
            }
        }
        for(int a = 0; a < 3; ++a)
        {
            nodes[node_index].lo[a] = lo[a] - eps;
            nodes[node_index].hi[a] = hi[a] + eps;
        }
        uint32_t n = last - first;
        if(n <= 4)
        {
            nodes[node_index].left = first;
            nodes[node_index].count = n;
            return;
        }
        int axis = 0;
        float ext = chi[0] - clo[0];
        for(int a = 1; a < 3; ++a) if(chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
        uint32_t mid = (first + last) / 2;
        if(ext > 0.0f)
        {
            constexpr int BINS = 16;
            struct bin { float lo[3], hi[3]; uint32_t n; };
            bin bins[BINS];
            for(auto& b: bins) { for(int a = 0; a < 3; ++a) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; } b.n = 0; }
            float scale = BINS / ext;
            auto bin_of = [&](uint32_t p) {
                float c = axis == 0 ? cent[p].x : (axis == 1 ? cent[p].y : cent[p].z);
                int b = (int)((c - clo[axis]) * scale);
                return b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
            };
            for(uint32_t i = first; i < last; ++i)
            {
                uint32_t p = order[i];
                bin& b = bins[bin_of(p)];
                const float l[3] = {blo[p].x, blo[p].y, blo[p].z}, h[3] = {bhi[p].x, bhi[p].y, bhi[p].z};
                for(int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], l[a]); b.hi[a] = std::max(b.hi[a], h[a]); }
                b.n++;
            }
            auto area = [](const float* l, const float* h) {
                float dx = h[0] - l[0], dy = h[1] - l[1], dz = h[2] - l[2];
                return dx * dy + dy * dz + dz * dx;
            };
            float right_area[BINS]; uint32_t right_n[BINS];
            {
                float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY}; uint32_t c = 0;
                for(int b = BINS - 1; b > 0; --b)
                {
                    for(int a = 0; a < 3; ++a) { l[a] = std::min(l[a], bins[b].lo[a]); h[a] = std::max(h[a], bins[b].hi[a]); }
                    c += bins[b].n;
                    right_area[b] = c ? area(l, h) : 0.0f; right_n[b] = c;
                }
            }
            float best = INFINITY; int best_split = -1;
            {
                float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY}; uint32_t c = 0;
                for(int b = 0; b < BINS - 1; ++b)
                {
                    for(int a = 0; a < 3; ++a) { l[a] = std::min(l[a], bins[b].lo[a]); h[a] = std::max(h[a], bins[b].hi[a]); }
                    c += bins[b].n;
                    if(c == 0 || right_n[b + 1] == 0) continue;
                    float cost = area(l, h) * c + right_area[b + 1] * right_n[b + 1];
                    if(cost < best) { best = cost; best_split = b; }
                }
            }
            if(best_split >= 0)
            {
                auto it = std::partition(order.begin() + first, order.begin() + last,
                    [&](uint32_t p) { return bin_of(p) <= best_split; });
                mid = (uint32_t)(it - order.begin());
            }
            if(best_split < 0 || mid == first || mid == last)
            {
                mid = (first + last) / 2;
                std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + last,
                    [&](uint32_t a, uint32_t b) {
                        float ca = axis == 0 ? cent[a].x : (axis == 1 ? cent[a].y : cent[a].z);
                        float cb = axis == 0 ? cent[b].x : (axis == 1 ? cent[b].y : cent[b].z);
                        return ca < cb;
                    });
            }
        }
        uint32_t left = (uint32_t)nodes.size();
        nodes.push_back({});
        nodes.push_back({});
        nodes[node_index].left = left;
        nodes[node_index].count = 0;
        build_node(left, first, mid);
        build_node(left + 1, mid, last);
    }
};

} // namespace orc
