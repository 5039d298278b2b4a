// tr_bvh.cu -- on-device acceleration-structure build: the replacement of
// bottom_/top_level_acceleration_structure (src/acceleration_structure.cc:66-269,296-443) and of
// scene_stage::ensure_blas (src/scene_stage.cc:821-872) for the default merged-static strategy.
//
// Pipeline (all kernels, no host round trips except two scalar read-backs):
//   1. pre_transform     world-space vertices per instance          (shader/pre_transform.comp:26-42)
//   2. extract_tri_lights world-space emissive triangles            (shader/extract_tri_lights.comp:17-54)
//   3. gather            48-byte triangle records + bounds (min/max reduction)
//   4. morton            63-bit Morton code of the triangle-box centre
//   5. sort              LSD radix sort, 8 x 8-bit passes; ranking by warp match/ballot + scans
//                        (the reference's radix_sort.cc wraps a Vulkan sort it never calls)
//   6. hierarchy         Karras 2012 binary radix tree over (code, index) keys
//   7. refit             bottom-up AABBs with one atomic flag per inner node
//   8. emit              breadth-first collapse into an 8-wide compressed BVH (80-byte nodes, <= 2
//                        triangles per leaf slot), boxes inflated by eps = 2^-16 * max|coordinate|
#include "tr_scene.cuh"
#include "tr_math.cuh"
#include "tr_build.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cstdlib>

namespace trd {

#define CUDA_OK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) return e_; } while(0)

// ---------------------------------------------------------------------------------------------
// 1. pre_transform.comp
// ---------------------------------------------------------------------------------------------
__global__ void k_pre_transform(const tr_vertex* __restrict__ in, tr_vertex* __restrict__ out, uint32_t count,
                                const float* __restrict__ model, const float* __restrict__ model_normal)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= count) return;
    tr_vertex v = in[i];
    float m[16], mnm[16];
#pragma unroll
    for(int k = 0; k < 16; ++k) { m[k] = model[k]; mnm[k] = model_normal[k]; }
    mat3 mn = upper3(mnm);
    vec3 p = xyz(mul4(m, V4(v.pos[0], v.pos[1], v.pos[2], 1)));
    vec3 n = normalize(mul(mn, ld3(v.normal)));
    vec3 t = normalize(mul(mn, V3(v.tangent[0], v.tangent[1], v.tangent[2])));
    float tw = v.tangent[3];
    if(determinant(mn) < 0) { n = -n; t = -t; tw = -tw; }
    v.pos[0] = p.x; v.pos[1] = p.y; v.pos[2] = p.z;
    v.normal[0] = n.x; v.normal[1] = n.y; v.normal[2] = n.z;
    v.tangent[0] = t.x; v.tangent[1] = t.y; v.tangent[2] = t.z; v.tangent[3] = tw;
    out[i] = v;
}

// ---------------------------------------------------------------------------------------------
// 2. extract_tri_lights.comp (PRE_TRANSFORMED_VERTICES variant)
// ---------------------------------------------------------------------------------------------
__global__ void k_extract_tri_lights(const tr_vertex* __restrict__ verts, const uint32_t* __restrict__ indices,
                                     uint32_t triangle_count, uint32_t instance_id, int32_t light_base_id,
                                     float ex, float ey, float ez, int32_t emission_tex_id, tr_tri_light* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= triangle_count) return;
    tr_tri_light l;
    l.emission_tex_id = emission_tex_id;
    l.emission_factor = rgb_to_r9g9b9e5(V3(ex, ey, ez));
    l.instance_id = instance_id;
    l.primitive_id = i;
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
        const tr_vertex& v = verts[indices[3 * i + k]];
        l.pos[k][0] = v.pos[0]; l.pos[k][1] = v.pos[1]; l.pos[k][2] = v.pos[2];
        l.uv[k] = float_to_half_(v.uv[0]) | (float_to_half_(v.uv[1]) << 16);
    }
    out[light_base_id + i] = l;
}

// ---------------------------------------------------------------------------------------------
// 2b. per-triangle shading records (DeviceScene::shade_tris)
// ---------------------------------------------------------------------------------------------
__global__ void k_shade_tris(const tr_vertex* __restrict__ verts, const uint32_t* __restrict__ indices, uint32_t triangle_count,
                             float4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= triangle_count) return;
    for(int k = 0; k < 3; ++k)
    {
        const float4* v = reinterpret_cast<const float4*>(verts + indices[3 * (size_t)i + k]);
        out[6 * (size_t)i + 2 * k] = v[0];
        out[6 * (size_t)i + 2 * k + 1] = v[1];
    }
}

// ---------------------------------------------------------------------------------------------
// 3. gather triangles + scene bounds
// ---------------------------------------------------------------------------------------------
// Order-preserving float <-> uint mapping for atomicMin/Max.
__device__ __forceinline__ uint32_t f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

__global__ void k_gather(const tr_vertex* __restrict__ verts, const uint32_t* __restrict__ indices, uint32_t triangle_count,
                         uint32_t instance_id, uint32_t opaque, uint32_t out_base, float4* __restrict__ tris,
                         uint32_t* __restrict__ bounds /* 6 ordered uints: lo xyz, hi xyz */)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if(i < triangle_count)
    {
        const tr_vertex& a = verts[indices[3 * i + 0]];
        const tr_vertex& b = verts[indices[3 * i + 1]];
        const tr_vertex& c = verts[indices[3 * i + 2]];
        float4* o = tris + 3 * (size_t)(out_base + i);
        o[0] = make_float4(a.pos[0], a.pos[1], a.pos[2], __int_as_float((int)instance_id));
        o[1] = make_float4(b.pos[0], b.pos[1], b.pos[2], __int_as_float((int)i));
        o[2] = make_float4(c.pos[0], c.pos[1], c.pos[2], __uint_as_float(opaque ? TR_TRI_FLAG_OPAQUE : 0u));
#pragma unroll
        for(int k = 0; k < 3; ++k)
        {
            lo[k] = fminf(fminf(a.pos[k], b.pos[k]), c.pos[k]);
            hi[k] = fmaxf(fmaxf(a.pos[k], b.pos[k]), c.pos[k]);
        }
    }
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
        for(int off = 16; off > 0; off >>= 1)
        {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], off));
        }
    }
    if((threadIdx.x & 31) == 0 && lo[0] <= hi[0])
    {
#pragma unroll
        for(int k = 0; k < 3; ++k)
        {
            atomicMin(bounds + k, f2ord(lo[k]));
            atomicMax(bounds + 3 + k, f2ord(hi[k]));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 4. Morton codes (63 bit, 21 per axis)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long expand21(uint32_t v)
{
    unsigned long long x = v & 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ tris, uint32_t n, const uint32_t* __restrict__ bounds,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float4 a = tris[3 * (size_t)i], b = tris[3 * (size_t)i + 1], c = tris[3 * (size_t)i + 2];
    float lo[3] = {ord2f(bounds[0]), ord2f(bounds[1]), ord2f(bounds[2])};
    float hi[3] = {ord2f(bounds[3]), ord2f(bounds[4]), ord2f(bounds[5])};
    float p[3] = {
        0.5f * (fminf(fminf(a.x, b.x), c.x) + fmaxf(fmaxf(a.x, b.x), c.x)),
        0.5f * (fminf(fminf(a.y, b.y), c.y) + fmaxf(fmaxf(a.y, b.y), c.y)),
        0.5f * (fminf(fminf(a.z, b.z), c.z) + fmaxf(fmaxf(a.z, b.z), c.z))};
    uint32_t q[3];
#pragma unroll
    for(int k = 0; k < 3; ++k)
    {
        float ext = hi[k] - lo[k];
        float t = ext > 0.0f ? (p[k] - lo[k]) / ext : 0.0f;
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        q[k] = min((uint32_t)(t * 2097152.0f), 2097151u);
    }
    keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    vals[i] = i;
}

// ---------------------------------------------------------------------------------------------
// 5. radix sort: 8-bit digits; per-tile histogram -> global scan -> stable scatter.
// Ranking inside a tile uses __match_any_sync (peers with my digit in the warp) + popc for the
// in-warp rank, a per-warp digit table in shared memory and a scan over warps.
// ---------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__global__ void k_sort_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist, uint32_t tiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE;
    for(int it = 0; it < SORT_ITEMS; ++it)
    {
        uint32_t i = base + it * SORT_THREADS + threadIdx.x;
        if(i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan over hist[digit][tile] (digit-major), single block
__global__ void k_sort_scan(uint32_t* __restrict__ hist, uint32_t total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    for(uint32_t base = 0; base < total; base += blockDim.x)
    {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < total ? hist[i] : 0;
        uint32_t x = v;
        for(int off = 1; off < 32; off <<= 1)
        {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
            if((threadIdx.x & 31) >= off) x += y;
        }
        if((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if(threadIdx.x < 32)
        {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? warp_sums[threadIdx.x] : 0;
            for(int off = 1; off < 32; off <<= 1)
            {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, off);
                if(threadIdx.x >= off) w += y;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        uint32_t warp_off = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
        uint32_t incl = x + warp_off + carry;
        if(i < total) hist[i] = incl - v;
        __syncthreads();
        if(threadIdx.x == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
}

__global__ void k_sort_scatter(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                               unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                               uint32_t n, int shift, const uint32_t* __restrict__ hist, uint32_t tiles)
{
    __shared__ uint32_t digit_base[256];           // global offset of this tile's next key of each digit
    __shared__ uint32_t warp_count[SORT_THREADS / 32][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    digit_base[threadIdx.x] = hist[threadIdx.x * tiles + blockIdx.x];
    const uint32_t base = blockIdx.x * SORT_TILE;
    for(int it = 0; it < SORT_ITEMS; ++it)
    {
        for(int w = 0; w < SORT_THREADS / 32; ++w) warp_count[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + it * SORT_THREADS + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = valid ? keys_in[i] : ~0ull;
        uint32_t digit = valid ? ((uint32_t)(key >> shift) & 255u) : 256u;
        // peers: lanes of this warp holding the same digit
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
        uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if(valid && rank_in_warp == 0) warp_count[warp][digit] = __popc(peers);
        __syncthreads();
        // exclusive scan over warps for each digit (thread d handles digit d)
        {
            uint32_t run = 0;
            for(int w = 0; w < SORT_THREADS / 32; ++w)
            {
                uint32_t c = warp_count[w][threadIdx.x];
                warp_count[w][threadIdx.x] = run;
                run += c;
            }
            __syncthreads();
            if(valid)
            {
                uint32_t dst = digit_base[digit] + warp_count[warp][digit] + rank_in_warp;
                keys_out[dst] = key;
                vals_out[dst] = vals_in[i];
            }
            __syncthreads();
            digit_base[threadIdx.x] += run;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// 6. Karras hierarchy. Inner node i in [0, n-2]; children encoded as index (inner) or ~leaf.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j)
{
    if(j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if(a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ children,
                            int2* __restrict__ ranges, int* __restrict__ inner_parent, int* __restrict__ leaf_parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while(delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for(int t = lmax >> 1; t >= 1; t >>= 1)
        if(delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do
    {
        t = (t + 1) >> 1;
        if(delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while(t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? ~gamma : gamma;
    int right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    ranges[i] = make_int2(first, last);
    if(left >= 0) inner_parent[left] = i; else leaf_parent[~left] = i;
    if(right >= 0) inner_parent[right] = i; else leaf_parent[~right] = i;
    if(i == 0) inner_parent[0] = -1;
}

// ---------------------------------------------------------------------------------------------
// 7. bottom-up refit
// ---------------------------------------------------------------------------------------------
__global__ void k_refit(const float4* __restrict__ tris, const uint32_t* __restrict__ order, int n,
                        const int2* __restrict__ children, const int* __restrict__ inner_parent,
                        const int* __restrict__ leaf_parent, float* __restrict__ leaf_boxes /* n*6 */,
                        float* __restrict__ inner_boxes /* (n-1)*6 */, int* __restrict__ flags, int climb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint32_t t = order[i];
    const float4 a = tris[3 * (size_t)t], b = tris[3 * (size_t)t + 1], c = tris[3 * (size_t)t + 2];
    float box[6] = {
        fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z),
        fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z)};
#pragma unroll
    for(int k = 0; k < 6; ++k) leaf_boxes[6 * (size_t)i + k] = box[k];
    if(n == 1 || !climb) return;
    int node = leaf_parent[i];
    while(node >= 0)
    {
        __threadfence();
        if(atomicAdd(&flags[node], 1) == 0) return; // first arrival: the sibling will finish
        __threadfence();
        const int2 ch = children[node];
        const volatile float* l = ch.x >= 0 ? inner_boxes + 6 * (size_t)ch.x : leaf_boxes + 6 * (size_t)(~ch.x);
        const volatile float* r = ch.y >= 0 ? inner_boxes + 6 * (size_t)ch.y : leaf_boxes + 6 * (size_t)(~ch.y);
#pragma unroll
        for(int k = 0; k < 3; ++k)
        {
            inner_boxes[6 * (size_t)node + k] = fminf(l[k], r[k]);
            inner_boxes[6 * (size_t)node + 3 + k] = fmaxf(l[3 + k], r[3 + k]);
        }
        node = inner_parent[node];
    }
}

// ---------------------------------------------------------------------------------------------
// 6b. PLOC hierarchy (Meister & Bittner 2018, "Parallel Locally-Ordered Clustering"): optional.
// Clusters start as the Morton-sorted triangles; every round each cluster finds, within +-PLOC_RADIUS
// positions, the neighbour whose merged box has the smallest surface area; mutual nearest
// neighbours merge into a new binary node; survivors are compacted in order. Selected with
// TR_TUNE_BUILDER=ploc; on the benchmark scene it traces shadow rays ~9% faster and closest-hit rays ~8%
// slower than the radix tree of section 6 (measured, profiles/), so the radix tree stays the default.
// ---------------------------------------------------------------------------------------------
constexpr int PLOC_RADIUS = 16;
constexpr int PLOC_BLOCK = 256;

__global__ void k_ploc_init(int n, const float* __restrict__ leaf_boxes, float4* __restrict__ lo, float4* __restrict__ hi,
                            int* __restrict__ cl_node)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float* b = leaf_boxes + 6 * (size_t)i;
    lo[i] = make_float4(b[0], b[1], b[2], 0.0f);
    hi[i] = make_float4(b[3], b[4], b[5], 0.0f);
    cl_node[i] = ~i;
}

__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_nn(int n, const float4* __restrict__ lo, const float4* __restrict__ hi, int* __restrict__ nn)
{
    __shared__ float4 s_lo[PLOC_BLOCK + 2 * PLOC_RADIUS], s_hi[PLOC_BLOCK + 2 * PLOC_RADIUS];
    const int base = blockIdx.x * PLOC_BLOCK - PLOC_RADIUS;
    for(int k = threadIdx.x; k < PLOC_BLOCK + 2 * PLOC_RADIUS; k += PLOC_BLOCK)
    {
        const int g = base + k;
        if(g >= 0 && g < n) { s_lo[k] = lo[g]; s_hi[k] = hi[g]; }
    }
    __syncthreads();
    const int i = blockIdx.x * PLOC_BLOCK + threadIdx.x;
    if(i >= n) return;
    const float4 al = s_lo[threadIdx.x + PLOC_RADIUS], ah = s_hi[threadIdx.x + PLOC_RADIUS];
    float best = FLT_MAX;
    int best_j = -1;
    for(int d = -PLOC_RADIUS; d <= PLOC_RADIUS; ++d)
    {
        const int j = i + d;
        if(d == 0 || j < 0 || j >= n) continue;
        const float4 bl = s_lo[threadIdx.x + PLOC_RADIUS + d], bh = s_hi[threadIdx.x + PLOC_RADIUS + d];
        const float dx = fmaxf(ah.x, bh.x) - fminf(al.x, bl.x);
        const float dy = fmaxf(ah.y, bh.y) - fminf(al.y, bl.y);
        const float dz = fmaxf(ah.z, bh.z) - fminf(al.z, bl.z);
        const float area = dx * dy + dy * dz + dz * dx;
        if(area < best) { best = area; best_j = j; }
    }
    nn[i] = best_j;
}

__global__ void k_ploc_merge(int n, const int* __restrict__ nn, int* __restrict__ cl_node, float4* __restrict__ lo, float4* __restrict__ hi,
                             int2* __restrict__ children, float* __restrict__ inner_boxes, int* __restrict__ counts,
                             int* __restrict__ node_counter, int* __restrict__ valid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const int j = nn[i];
    if(j < 0 || nn[j] != i) { valid[i] = 1; return; }
    if(i > j) { valid[i] = 0; return; }
    // i < j: merge j into i
    const int id = atomicAdd(node_counter, 1);
    const int a = cl_node[i], b = cl_node[j];
    children[id] = make_int2(a, b);
    counts[id] = (a < 0 ? 1 : counts[a]) + (b < 0 ? 1 : counts[b]);
    const float4 al = lo[i], ah = hi[i], bl = lo[j], bh = hi[j];
    const float4 l = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.0f);
    const float4 h = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.0f);
    float* o = inner_boxes + 6 * (size_t)id;
    o[0] = l.x; o[1] = l.y; o[2] = l.z; o[3] = h.x; o[4] = h.y; o[5] = h.z;
    lo[i] = l; hi[i] = h;
    cl_node[i] = id;
    valid[i] = 1;
}

__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_count(int n, const int* __restrict__ valid, uint32_t* __restrict__ block_sums)
{
    const int i = blockIdx.x * PLOC_BLOCK + threadIdx.x;
    const int v = i < n ? valid[i] : 0;
    const int c = __syncthreads_count(v);
    if(threadIdx.x == 0) block_sums[blockIdx.x] = (uint32_t)c;
    if(threadIdx.x == 0 && blockIdx.x == 0) block_sums[gridDim.x] = 0u;   // becomes the total after the exclusive scan
}

__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_compact(int n, const int* __restrict__ valid, const uint32_t* __restrict__ block_offsets,
                                                              const int* __restrict__ node_in, const float4* __restrict__ lo_in, const float4* __restrict__ hi_in,
                                                              int* __restrict__ node_out, float4* __restrict__ lo_out, float4* __restrict__ hi_out)
{
    __shared__ uint32_t warp_base[PLOC_BLOCK / 32];
    const int i = blockIdx.x * PLOC_BLOCK + threadIdx.x;
    const int v = i < n ? valid[i] : 0;
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if(lane == 0) warp_base[warp] = __popc(ballot);
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t run = 0;
        for(int w = 0; w < PLOC_BLOCK / 32; ++w) { uint32_t c = warp_base[w]; warp_base[w] = run; run += c; }
    }
    __syncthreads();
    if(v)
    {
        const uint32_t dst = block_offsets[blockIdx.x] + warp_base[warp] + __popc(ballot & ((1u << lane) - 1u));
        node_out[dst] = node_in[i];
        lo_out[dst] = lo_in[i];
        hi_out[dst] = hi_in[i];
    }
}


// ---------------------------------------------------------------------------------------------
// 6c. Top-down binned-SAH hierarchy (Wald 2007, "On fast construction of SAH-based bounding volume hierarchies"),
// built breadth first on the device: the quality build for static geometry -- what the reference asks of the driver
// with ePreferFastTrace for its static BLAS (src/acceleration_structure.cc:150-262). One thread block per node and
// level: centroid bounds, 16 bins per axis (bounds + counts through shared-memory atomics), the 45 split candidates
// evaluated in parallel, then a stable partition of the node's primitive range into the other ping-pong buffer.
// A split that leaves one side empty (all centroids in one bin) falls back to the median of the range. The result
// feeds the same 8-wide collapse as the radix tree: children / counts / inner_boxes by node id, leaves = positions in
// the final primitive order.
// ---------------------------------------------------------------------------------------------
constexpr int SAH_BINS = 16;
constexpr int SAH_BLOCK = 128;

struct SahTask
{
    int begin, end;      // primitive range in the current buffer
    int id;              // inner node id
    float lo[3], hi[3];  // bounds of the node's triangles
};

__global__ void k_sah_boxes(const float4* __restrict__ tris, uint32_t n, float* __restrict__ boxes, uint32_t* __restrict__ perm)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float4 a = tris[3 * (size_t)i], b = tris[3 * (size_t)i + 1], c = tris[3 * (size_t)i + 2];
    float* o = boxes + 6 * (size_t)i;
    o[0] = fminf(fminf(a.x, b.x), c.x); o[1] = fminf(fminf(a.y, b.y), c.y); o[2] = fminf(fminf(a.z, b.z), c.z);
    o[3] = fmaxf(fmaxf(a.x, b.x), c.x); o[4] = fmaxf(fmaxf(a.y, b.y), c.y); o[5] = fmaxf(fmaxf(a.z, b.z), c.z);
    perm[i] = i;
}

__device__ __forceinline__ float half_area(const float* lo, const float* hi)
{
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(SAH_BLOCK) k_sah_level(const SahTask* __restrict__ tasks, int task_count, const float* __restrict__ boxes,
                                                         const uint32_t* __restrict__ perm_in, uint32_t* __restrict__ perm_out,
                                                         uint32_t* __restrict__ perm_final, int2* __restrict__ children, int* __restrict__ counts,
                                                         float* __restrict__ inner_boxes, SahTask* __restrict__ next_tasks,
                                                         int* __restrict__ next_count, int* __restrict__ node_counter, int force_median,
                                                         int axis_mode, int level, float axis_penalty)
{
    __shared__ uint32_t s_cb[6];                               // centroid bounds, ordered ints
    __shared__ uint32_t s_bin[3][SAH_BINS][6];                 // bin bounds, ordered ints
    __shared__ int s_cnt[3][SAH_BINS];
    __shared__ float s_cost[3 * (SAH_BINS - 1)];
    __shared__ int s_best, s_nl, s_left_off, s_right_off, s_ids[2];
    __shared__ uint32_t s_lbox[6], s_rbox[6];                  // child bounds, ordered ints
    __shared__ int s_warp_l[SAH_BLOCK / 32], s_warp_r[SAH_BLOCK / 32];
    const int tid = threadIdx.x;
    const SahTask task = tasks[blockIdx.x];
    const int begin = task.begin, end = task.end, count = end - begin;

    if(tid < 6) s_cb[tid] = tid < 3 ? 0xFFFFFFFFu : 0u;
    for(int k = tid; k < 3 * SAH_BINS * 6; k += SAH_BLOCK) (&s_bin[0][0][0])[k] = ((k % 6) < 3) ? 0xFFFFFFFFu : 0u;
    for(int k = tid; k < 3 * SAH_BINS; k += SAH_BLOCK) (&s_cnt[0][0])[k] = 0;
    __syncthreads();
    // ---- centroid bounds
    {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for(int i = begin + tid; i < end; i += SAH_BLOCK)
        {
            const float* b = boxes + 6 * (size_t)perm_in[i];
            for(int a = 0; a < 3; ++a)
            {
                const float c = 0.5f * (b[a] + b[3 + a]);
                lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c);
            }
        }
        for(int a = 0; a < 3; ++a)
        {
            for(int off = 16; off > 0; off >>= 1)
            {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], off));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], off));
            }
            if((tid & 31) == 0 && lo[a] <= hi[a]) { atomicMin(&s_cb[a], f2ord(lo[a])); atomicMax(&s_cb[3 + a], f2ord(hi[a])); }
        }
    }
    __syncthreads();
    float cmin[3], scale[3];
    for(int a = 0; a < 3; ++a)
    {
        cmin[a] = ord2f(s_cb[a]);
        const float ext = ord2f(s_cb[3 + a]) - cmin[a];
        scale[a] = ext > 0.0f ? (float)SAH_BINS * (1.0f - 1e-6f) / ext : 0.0f;
    }
    // ---- binning
    for(int i = begin + tid; i < end; i += SAH_BLOCK)
    {
        const float* b = boxes + 6 * (size_t)perm_in[i];
        for(int a = 0; a < 3; ++a)
        {
            const float c = 0.5f * (b[a] + b[3 + a]);
            int k = (int)((c - cmin[a]) * scale[a]);
            k = max(0, min(SAH_BINS - 1, k));
            atomicAdd(&s_cnt[a][k], 1);
            for(int j = 0; j < 3; ++j) { atomicMin(&s_bin[a][k][j], f2ord(b[j])); atomicMax(&s_bin[a][k][3 + j], f2ord(b[3 + j])); }
        }
    }
    __syncthreads();
    // ---- the 45 split candidates: split s of axis a puts bins [0, s) left
    if(tid < 3 * (SAH_BINS - 1))
    {
        const int a = tid / (SAH_BINS - 1), sp = tid % (SAH_BINS - 1) + 1;
        float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        float rlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        int nl = 0, nr = 0;
        for(int k = 0; k < SAH_BINS; ++k)
        {
            const int c = s_cnt[a][k];
            if(c == 0) continue;
            float* lo = k < sp ? llo : rlo;
            float* hi = k < sp ? lhi : rhi;
            for(int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], ord2f(s_bin[a][k][j])); hi[j] = fmaxf(hi[j], ord2f(s_bin[a][k][3 + j])); }
            if(k < sp) nl += c; else nr += c;
        }
        float cost = (nl == 0 || nr == 0 || force_median) ? FLT_MAX : half_area(llo, lhi) * (float)nl + half_area(rlo, rhi) * (float)nr;
        // axis_mode 1: only the axis (level mod 3) competes (an octree-like arrangement that the octant-ordered traversal
        // of the wide nodes likes); 2: only the axis of the largest centroid extent. Other axes remain the fallback.
        if(axis_mode == 1 && a != level % 3) cost = cost == FLT_MAX ? FLT_MAX : cost * axis_penalty;
        if(axis_mode == 2)
        {
            const float e0 = ord2f(s_cb[3]) - ord2f(s_cb[0]), e1 = ord2f(s_cb[4]) - ord2f(s_cb[1]), e2 = ord2f(s_cb[5]) - ord2f(s_cb[2]);
            const int longest = e0 >= e1 ? (e0 >= e2 ? 0 : 2) : (e1 >= e2 ? 1 : 2);
            if(a != longest) cost = cost == FLT_MAX ? FLT_MAX : cost * axis_penalty;
        }
        s_cost[tid] = cost;
    }
    __syncthreads();
    if(tid == 0)
    {
        int best = -1;
        float bc = FLT_MAX;
        for(int k = 0; k < 3 * (SAH_BINS - 1); ++k)
            if(s_cost[k] < bc) { bc = s_cost[k]; best = k; }
        s_best = best;
        int nl = count / 2;                       // median fallback
        if(best >= 0)
        {
            const int a = best / (SAH_BINS - 1), sp = best % (SAH_BINS - 1) + 1;
            nl = 0;
            for(int k = 0; k < sp; ++k) nl += s_cnt[a][k];
        }
        s_nl = nl;
        s_left_off = 0; s_right_off = 0;
        // inner ids for the children that are not single triangles
        const int inner = (nl > 1 ? 1 : 0) + (count - nl > 1 ? 1 : 0);
        const int base = inner ? atomicAdd(node_counter, inner) : 0;
        s_ids[0] = nl > 1 ? base : -1;
        s_ids[1] = count - nl > 1 ? base + (nl > 1 ? 1 : 0) : -1;
    }
    if(tid < 6) { s_lbox[tid] = tid < 3 ? 0xFFFFFFFFu : 0u; s_rbox[tid] = tid < 3 ? 0xFFFFFFFFu : 0u; }
    __syncthreads();
    const int best = s_best, nl = s_nl;
    const int axis = best >= 0 ? best / (SAH_BINS - 1) : 0, split = best >= 0 ? best % (SAH_BINS - 1) + 1 : 0;
    // ---- stable partition into the other buffer, child bounds on the way
    float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float rlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, rhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(int chunk = begin; chunk < end; chunk += SAH_BLOCK)
    {
        const int i = chunk + tid;
        const bool valid = i < end;
        uint32_t prim = 0;
        bool left = false;
        if(valid)
        {
            prim = perm_in[i];
            const float* b = boxes + 6 * (size_t)prim;
            if(best >= 0)
            {
                const float c = 0.5f * (b[axis] + b[3 + axis]);
                int k = (int)((c - cmin[axis]) * scale[axis]);
                k = max(0, min(SAH_BINS - 1, k));
                left = k < split;
            }
            else left = (i - begin) < nl;
            float* lo = left ? llo : rlo;
            float* hi = left ? lhi : rhi;
            for(int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], b[j]); hi[j] = fmaxf(hi[j], b[3 + j]); }
        }
        const uint32_t lm = __ballot_sync(0xFFFFFFFFu, valid && left), rm = __ballot_sync(0xFFFFFFFFu, valid && !left);
        const int w = tid >> 5, lane = tid & 31;
        if(lane == 0) { s_warp_l[w] = __popc(lm); s_warp_r[w] = __popc(rm); }
        __syncthreads();
        int lbase = s_left_off, rbase = s_right_off;
        for(int k = 0; k < w; ++k) { lbase += s_warp_l[k]; rbase += s_warp_r[k]; }
        if(valid)
        {
            const int dst = left ? begin + lbase + __popc(lm & ((1u << lane) - 1u)) : begin + nl + rbase + __popc(rm & ((1u << lane) - 1u));
            perm_out[dst] = prim;
            if((left && nl == 1) || (!left && count - nl == 1)) perm_final[dst] = prim;
        }
        __syncthreads();
        if(tid == 0)
        {
            int tl = 0, tr = 0;
            for(int k = 0; k < SAH_BLOCK / 32; ++k) { tl += s_warp_l[k]; tr += s_warp_r[k]; }
            s_left_off += tl; s_right_off += tr;
        }
        __syncthreads();
    }
    for(int j = 0; j < 3; ++j)
    {
        for(int off = 16; off > 0; off >>= 1)
        {
            llo[j] = fminf(llo[j], __shfl_xor_sync(0xFFFFFFFFu, llo[j], off)); lhi[j] = fmaxf(lhi[j], __shfl_xor_sync(0xFFFFFFFFu, lhi[j], off));
            rlo[j] = fminf(rlo[j], __shfl_xor_sync(0xFFFFFFFFu, rlo[j], off)); rhi[j] = fmaxf(rhi[j], __shfl_xor_sync(0xFFFFFFFFu, rhi[j], off));
        }
        if((tid & 31) == 0)
        {
            if(llo[j] <= lhi[j]) { atomicMin(&s_lbox[j], f2ord(llo[j])); atomicMax(&s_lbox[3 + j], f2ord(lhi[j])); }
            if(rlo[j] <= rhi[j]) { atomicMin(&s_rbox[j], f2ord(rlo[j])); atomicMax(&s_rbox[3 + j], f2ord(rhi[j])); }
        }
    }
    __syncthreads();
    if(tid == 0)
    {
        const int nr = count - nl;
        children[task.id] = make_int2(nl > 1 ? s_ids[0] : ~begin, nr > 1 ? s_ids[1] : ~(begin + nl));
        counts[task.id] = count;
        for(int j = 0; j < 6; ++j) inner_boxes[6 * (size_t)task.id + j] = j < 3 ? task.lo[j] : task.hi[j - 3];
        const int pushes = (nl > 1 ? 1 : 0) + (nr > 1 ? 1 : 0);
        if(pushes)
        {
            int at = atomicAdd(next_count, pushes);
            if(nl > 1)
            {
                SahTask t;
                t.begin = begin; t.end = begin + nl; t.id = s_ids[0];
                for(int j = 0; j < 3; ++j) { t.lo[j] = ord2f(s_lbox[j]); t.hi[j] = ord2f(s_lbox[3 + j]); }
                next_tasks[at++] = t;
            }
            if(nr > 1)
            {
                SahTask t;
                t.begin = begin + nl; t.end = end; t.id = s_ids[1];
                for(int j = 0; j < 3; ++j) { t.lo[j] = ord2f(s_rbox[j]); t.hi[j] = ord2f(s_rbox[3 + j]); }
                next_tasks[at++] = t;
            }
        }
    }
}

// leaf boxes + triangle records in the final primitive order
__global__ void k_sah_finalize(const float4* __restrict__ tris, const float* __restrict__ boxes, const uint32_t* __restrict__ perm, uint32_t n,
                               float4* __restrict__ tris_out, float* __restrict__ leaf_boxes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint32_t t = perm[i];
    for(int w = 0; w < 3; ++w) tris_out[3 * (size_t)i + w] = tris[3 * (size_t)t + w];
    for(int k = 0; k < 6; ++k) leaf_boxes[6 * (size_t)i + k] = boxes[6 * (size_t)t + k];
}

__global__ void k_counts_from_ranges(int n_inner, const int2* __restrict__ ranges, int* __restrict__ counts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n_inner) counts[i] = ranges[i].y - ranges[i].x + 1;
}

// ---------------------------------------------------------------------------------------------
// 8. breadth-first emission of the 8-wide compressed BVH
// ---------------------------------------------------------------------------------------------
// One thread per wide node. The node's subtree root is a Karras inner node; its child set starts
// as the two binary children and the child with the largest surface area is opened (replaced by
// its two children) until there are 8 or nothing can be opened. A binary subtree with
// <= TR_BVH_LEAF_TRIS triangles becomes a leaf child. Children are assigned to slots per octant
// (greedy on dot(centroid offset, slot sign)), boxes are quantised to 8 bits outwards.
__device__ __forceinline__ int range_size(const int* __restrict__ counts, int c) { return c < 0 ? 1 : counts[c]; }

// Sorted-triangle positions of the (<= 3) leaves below binary node c.
__device__ __forceinline__ int collect_leaves(const int2* __restrict__ children, int c, int* out)
{
    if(c < 0) { out[0] = ~c; return 1; }
    int stack[4], sp = 0, n = 0;
    stack[sp++] = c;
    while(sp)
    {
        const int2 ch = children[stack[--sp]];
        if(ch.y < 0) { if(n < 3) out[n] = ~ch.y; n++; } else stack[sp++] = ch.y;
        if(ch.x < 0) { if(n < 3) out[n] = ~ch.x; n++; } else stack[sp++] = ch.x;
    }
    return n < 3 ? n : 3;
}

// 80-byte node, tr_scene.cuh. `count_of_slot[s]` = triangles of the leaf in slot s (1..2), 0 for inner /
// empty slots.
__device__ __forceinline__ void write_wide_node(float4* __restrict__ out, const float* nlo, const float* nhi, int n,
                                                const float (*clo)[3], const float (*chi)[3], const int* slot_child,
                                                const uint32_t* count_of_slot, uint32_t imask, uint32_t child_base, uint32_t tri_base)
{
    // per-axis quantum 2^e with 127 * 2^e >= extent (one step of margin for the outward rounding);
    // 7-bit quanta, stored as 0x80 | q so the traversal can build 1 + q/128 with one PRMT. The
    // exponent byte holds e + 7 (+127 bias): the traversal's scale is 128 * 2^e.
    uint32_t ebyte[3];
    float scale[3];
    for(int a = 0; a < 3; ++a)
    {
        float ext = fmaxf(nhi[a] - nlo[a], 1e-30f);
        int e;
        frexpf(ext / 126.0f, &e);          // ext/126 = m * 2^e, m in [0.5, 1)  =>  2^e >= ext/126
        e = max(-126, min(120, e));
        ebyte[a] = (uint32_t)(e + 7 + 127);
        scale[a] = __uint_as_float((uint32_t)(e + 127) << 23);
    }
    // empty slots: inverted box (lo = 127, hi = 0) -> never hit
    uint32_t q[6][2] = {{0xFFFFFFFFu, 0xFFFFFFFFu}, {0xFFFFFFFFu, 0xFFFFFFFFu}, {0xFFFFFFFFu, 0xFFFFFFFFu},
                        {0x80808080u, 0x80808080u}, {0x80808080u, 0x80808080u}, {0x80808080u, 0x80808080u}};
    uint32_t leaf8 = 0, two8 = 0;
    for(int s = 0; s < 8; ++s)
    {
        const int c = slot_child[s];
        if(c < 0) continue;
        if(count_of_slot[s] >= 1u) leaf8 |= 1u << s;
        if(count_of_slot[s] >= 2u) two8 |= 1u << s;
        for(int a = 0; a < 3; ++a)
        {
            float fl = floorf((clo[c][a] - nlo[a]) / scale[a]);
            float fh = ceilf((chi[c][a] - nlo[a]) / scale[a]);
            int ql = max(0, min(127, (int)fl));
            int qh = max(0, min(127, (int)fh));
            while(ql > 0 && nlo[a] + (float)ql * scale[a] > clo[c][a]) ql--;
            while(qh < 127 && nlo[a] + (float)qh * scale[a] < chi[c][a]) qh++;
            const int sh = 8 * (s & 3);
            q[a][s >> 2] = (q[a][s >> 2] & ~(0xFFu << sh)) | ((0x80u | (uint32_t)ql) << sh);
            q[3 + a][s >> 2] = (q[3 + a][s >> 2] & ~(0xFFu << sh)) | ((0x80u | (uint32_t)qh) << sh);
        }
    }
    // the inner-child mask in the visiting order of ray octants 0..3: bit (slot ^ o); octant o >= 4 is the
    // bit reversal of octant o ^ 7 (tr_trace.cuh)
    uint32_t perm = 0;
    for(uint32_t o = 0; o < 4; ++o)
    {
        uint32_t m = 0;
        for(uint32_t s = 0; s < 8; ++s)
            if(imask & (1u << s)) m |= 1u << (s ^ o);
        perm |= m << (8 * o);
    }
    const uint32_t e4 = ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16);
    out[0] = make_float4(nlo[0], nlo[1], nlo[2], __uint_as_float(e4));
    out[1] = make_float4(__uint_as_float(child_base), __uint_as_float(tri_base), __uint_as_float(perm), __uint_as_float(leaf8 | (two8 << 8)));
    out[2] = make_float4(__uint_as_float(q[0][0]), __uint_as_float(q[0][1]), __uint_as_float(q[1][0]), __uint_as_float(q[1][1]));
    out[3] = make_float4(__uint_as_float(q[2][0]), __uint_as_float(q[2][1]), __uint_as_float(q[3][0]), __uint_as_float(q[3][1]));
    out[4] = make_float4(__uint_as_float(q[4][0]), __uint_as_float(q[4][1]), __uint_as_float(q[5][0]), __uint_as_float(q[5][1]));
    (void)n;
}

// In-order (left to right) leaf positions below binary node c, at most `cap`.
__device__ __forceinline__ int collect_leaves_ordered(const int2* __restrict__ children, int c, int* out, int cap)
{
    if(c < 0) { out[0] = ~c; return 1; }
    int stack[20], sp = 0, n = 0;
    stack[sp++] = c;
    while(sp)
    {
        const int v = stack[--sp];
        if(v < 0) { if(n < cap) out[n] = ~v; n++; continue; }
        const int2 ch = children[v];
        if(sp < 19) { stack[sp++] = ch.y; stack[sp++] = ch.x; }
    }
    return n < cap ? n : cap;
}

constexpr int WIDE_BOTTOM_TRIS = 16;   // a subtree of at most 8 x 2 triangles becomes ONE node of leaf slots

__global__ void k_wide_level(int leaf_tris, const int* __restrict__ frontier, int frontier_size, int out_base,
                             const int2* __restrict__ children, const int* __restrict__ ranges,
                             const float* __restrict__ leaf_boxes, const float* __restrict__ inner_boxes, float eps,
                             const float4* __restrict__ tris_sorted, float4* __restrict__ nodes, float4* __restrict__ tris_out,
                             int* __restrict__ next_frontier, int* __restrict__ next_count, int next_out_base,
                             uint32_t* __restrict__ tri_counter, int bottom_tris, int keep_closed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= frontier_size) return;
    const int root = frontier[i];
    // the node's children: either binary subtrees (inner >= 0: becomes a wide node, else a leaf slot) ...
    int n = 0;
    int child_inner[8];          // binary node id of an inner child, -1 for a leaf slot
    int child_cnt[8];            // triangles of a leaf slot (1..2)
    int child_leaf[8][2];        // their positions in tris_sorted
    float clo[8][3], chi[8][3];
    const int root_count = range_size(ranges, root);
    if(leaf_tris >= 2 && root_count <= bottom_tris && root_count <= WIDE_BOTTOM_TRIS)
    {
        // ... or, for a small subtree, ONE bottom node: its triangles in order, cut into <= 8 slots of 1-2 consecutive
        // triangles so that the sum of (slot box area x triangles in the slot) is minimal (dynamic programme over
        // (triangles consumed, slots used)). Denser bottom nodes = fewer node steps per ray: the slab test of a node
        // costs the same for 3 or 8 occupied slots.
        int pos[WIDE_BOTTOM_TRIS];
        const int m = collect_leaves_ordered(children, root, pos, WIDE_BOTTOM_TRIS);
        float a1[WIDE_BOTTOM_TRIS], a2[WIDE_BOTTOM_TRIS];   // cost of a slot holding triangle j alone / j-1 and j
        for(int j = 0; j < m; ++j)
        {
            const float* b = leaf_boxes + 6 * (size_t)pos[j];
            a1[j] = (b[3] - b[0]) * (b[4] - b[1]) + (b[4] - b[1]) * (b[5] - b[2]) + (b[5] - b[2]) * (b[3] - b[0]);
            a2[j] = FLT_MAX;
            if(j > 0)
            {
                const float* p = leaf_boxes + 6 * (size_t)pos[j - 1];
                const float dx = fmaxf(b[3], p[3]) - fminf(b[0], p[0]), dy = fmaxf(b[4], p[4]) - fminf(b[1], p[1]), dz = fmaxf(b[5], p[5]) - fminf(b[2], p[2]);
                a2[j] = 2.0f * (dx * dy + dy * dz + dz * dx);
            }
        }
        // f[j][k]: best cost of the first j triangles in exactly k slots; choice[j][k] = size of the last slot
        float f[WIDE_BOTTOM_TRIS + 1][9];
        unsigned char choice[WIDE_BOTTOM_TRIS + 1][9];
        for(int j = 0; j <= m; ++j) for(int k = 0; k <= 8; ++k) { f[j][k] = FLT_MAX; choice[j][k] = 0; }
        f[0][0] = 0.0f;
        for(int j = 1; j <= m; ++j)
            for(int k = 1; k <= 8; ++k)
            {
                float best = FLT_MAX; unsigned char ch = 0;
                if(f[j - 1][k - 1] < FLT_MAX) { best = f[j - 1][k - 1] + a1[j - 1]; ch = 1; }
                if(j >= 2 && f[j - 2][k - 1] < FLT_MAX && f[j - 2][k - 1] + a2[j - 1] < best) { best = f[j - 2][k - 1] + a2[j - 1]; ch = 2; }
                f[j][k] = best; choice[j][k] = ch;
            }
        int bk = -1;
        for(int k = 1; k <= 8; ++k) if(f[m][k] < FLT_MAX && (bk < 0 || f[m][k] < f[m][bk])) bk = k;
        // walk back
        n = bk;
        int j = m;
        for(int k = bk; k >= 1; --k)
        {
            const int sz = choice[j][k];
            child_inner[k - 1] = -1;
            child_cnt[k - 1] = sz;
            child_leaf[k - 1][0] = pos[j - sz];
            child_leaf[k - 1][1] = sz == 2 ? pos[j - 1] : pos[j - sz];
            j -= sz;
        }
        for(int k = 0; k < n; ++k)
            for(int a = 0; a < 3; ++a)
            {
                const float* b0 = leaf_boxes + 6 * (size_t)child_leaf[k][0];
                const float* b1 = leaf_boxes + 6 * (size_t)child_leaf[k][1];
                clo[k][a] = fminf(b0[a], b1[a]) - eps;
                chi[k][a] = fmaxf(b0[3 + a], b1[3 + a]) + eps;
            }
    }
    else
    {
        int cand[8];
        n = 2;
        cand[0] = children[root].x;
        cand[1] = children[root].y;
        while(n < 8)
        {
            int best = -1;
            float best_area = -1.0f;
            for(int k = 0; k < n; ++k)
            {
                const int c = cand[k];
                if(c < 0 || range_size(ranges, c) <= keep_closed) continue;
                const float* b = inner_boxes + 6 * (size_t)c;
                const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
                const float area = dx * dy + dy * dz + dz * dx;
                if(area > best_area) { best_area = area; best = k; }
            }
            if(best < 0) break;
            const int c = cand[best];
            cand[best] = children[c].x;
            cand[n++] = children[c].y;
        }
        for(int k = 0; k < n; ++k)
        {
            const int c = cand[k];
            const float* b = c < 0 ? leaf_boxes + 6 * (size_t)(~c) : inner_boxes + 6 * (size_t)c;
            for(int a = 0; a < 3; ++a) { clo[k][a] = b[a] - eps; chi[k][a] = b[3 + a] + eps; }
            const int cnt = range_size(ranges, c);
            if(c < 0 || cnt <= leaf_tris)
            {
                int leaves[3];
                const int got = collect_leaves(children, c, leaves);
                child_inner[k] = -1;
                child_cnt[k] = got < cnt ? got : cnt;
                child_leaf[k][0] = leaves[0];
                child_leaf[k][1] = got > 1 ? leaves[1] : leaves[0];
            }
            else { child_inner[k] = c; child_cnt[k] = 0; child_leaf[k][0] = child_leaf[k][1] = 0; }
        }
    }
    float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(int k = 0; k < n; ++k)
        for(int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], clo[k][a]); nhi[a] = fmaxf(nhi[a], chi[k][a]); }
    // greedy slot assignment: maximise dot(child centre - node centre, sign(slot))
    int slot_child[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    bool assigned[8] = {false, false, false, false, false, false, false, false};
    for(int it = 0; it < n; ++it)
    {
        float best = -FLT_MAX;
        int bc = -1, bs = -1;
        for(int k = 0; k < n; ++k)
        {
            if(assigned[k]) continue;
            const float cx = 0.5f * (clo[k][0] + chi[k][0]) - 0.5f * (nlo[0] + nhi[0]);
            const float cy = 0.5f * (clo[k][1] + chi[k][1]) - 0.5f * (nlo[1] + nhi[1]);
            const float cz = 0.5f * (clo[k][2] + chi[k][2]) - 0.5f * (nlo[2] + nhi[2]);
            for(int s = 0; s < 8; ++s)
            {
                if(slot_child[s] >= 0) continue;
                const float cost = ((s & 4) ? cx : -cx) + ((s & 2) ? cy : -cy) + ((s & 1) ? cz : -cz);
                if(cost > best) { best = cost; bc = k; bs = s; }
            }
        }
        assigned[bc] = true;
        slot_child[bs] = bc;
    }
    // leaf children: triangles become contiguous from tri_base; inner children: contiguous nodes
    uint32_t tri_total = 0, inner_total = 0;
    for(int k = 0; k < n; ++k)
    {
        if(child_inner[k] < 0) tri_total += (uint32_t)child_cnt[k];
        else inner_total++;
    }
    const uint32_t tri_base = tri_total ? atomicAdd(tri_counter, tri_total) : 0u;
    const int inner_base = inner_total ? atomicAdd(next_count, (int)inner_total) : 0;
    uint32_t count_of_slot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t imask = 0, tri_off = 0, inner_off = 0;
    for(int s = 0; s < 8; ++s)
    {
        const int k = slot_child[s];
        if(k < 0) continue;
        if(child_inner[k] < 0)
        {
            for(int t = 0; t < child_cnt[k]; ++t)
                for(int w = 0; w < 3; ++w)
                    tris_out[3 * (size_t)(tri_base + tri_off + t) + w] = tris_sorted[3 * (size_t)child_leaf[k][t] + w];
            count_of_slot[s] = (uint32_t)child_cnt[k];
            tri_off += (uint32_t)child_cnt[k];
        }
        else
        {
            next_frontier[inner_base + (int)inner_off] = child_inner[k];
            inner_off++;
            imask |= 1u << s;
        }
    }
    write_wide_node(nodes + TR_BVH_NODE_VEC4 * (size_t)(out_base + i), nlo, nhi, n, clo, chi, slot_child, count_of_slot, imask,
                    (uint32_t)(next_out_base + inner_base), tri_base);
}

// Whole scene fits one leaf child (n <= TR_BVH_LEAF_TRIS), or is empty (n == 0).
__global__ void k_wide_single(int n, const float* __restrict__ leaf_boxes, const float* __restrict__ inner_boxes, float eps,
                              const float4* __restrict__ tris_sorted, float4* __restrict__ nodes, float4* __restrict__ tris_out)
{
    float clo[8][3], chi[8][3], nlo[3] = {0, 0, 0}, nhi[3] = {1, 1, 1};
    int slot_child[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    uint32_t count_of_slot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if(n > 0)
    {
        const float* box = n == 1 ? leaf_boxes : inner_boxes;
        for(int a = 0; a < 3; ++a)
        {
            clo[0][a] = nlo[a] = box[a] - eps;
            chi[0][a] = nhi[a] = box[3 + a] + eps;
        }
        slot_child[0] = 0;
        count_of_slot[0] = (uint32_t)n;
        for(int t = 0; t < 3 * n; ++t) tris_out[t] = tris_sorted[t];
    }
    write_wide_node(nodes, nlo, nhi, n > 0 ? 1 : 0, clo, chi, slot_child, count_of_slot, 0u, 0u, 0u);
}

__global__ void k_reorder_tris(const float4* __restrict__ tris, const uint32_t* __restrict__ order, uint32_t n, float4* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint32_t t = order[i];
    out[3 * (size_t)i + 0] = tris[3 * (size_t)t + 0];
    out[3 * (size_t)i + 1] = tris[3 * (size_t)t + 1];
    out[3 * (size_t)i + 2] = tris[3 * (size_t)t + 2];
}

// ---------------------------------------------------------------------------------------------
// 9. refit of the wide tree to moved vertices (src/acceleration_structure.cc:187-262 `eUpdate`)
// ---------------------------------------------------------------------------------------------
// The leaf records keep their order and their (instance, primitive) words; only the positions are re-fetched.
__global__ void k_refit_leaf_tris(const tr_vertex* __restrict__ verts, const uint32_t* __restrict__ indices,
                                  const tr_build_instance* __restrict__ inst, uint32_t n, float4* __restrict__ tris,
                                  uint32_t* __restrict__ maxabs_bits)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if(i < n)
    {
        float4* o = tris + 3 * (size_t)i;
        const int instance = __float_as_int(o[0].w), prim = __float_as_int(o[1].w);
        const tr_build_instance I = inst[instance];
        const uint32_t* idx = indices + I.index_base + 3 * (size_t)prim;
        for(int k = 0; k < 3; ++k)
        {
            const tr_vertex& v = verts[I.vertex_base + idx[k]];
            o[k] = make_float4(v.pos[0], v.pos[1], v.pos[2], o[k].w);
            m = fmaxf(m, fmaxf(fabsf(v.pos[0]), fmaxf(fabsf(v.pos[1]), fabsf(v.pos[2]))));
        }
    }
    for(int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(maxabs_bits, __float_as_uint(m));   // non-negative floats order like their bits
}

// One thread per wide node of one level (the levels below are done): exact bounds of every occupied slot -- an inner
// child's from `node_bounds`, a leaf slot's from its one or two triangles --, then the same inflation, quantisation and
// masks as at build time (write_wide_node) with the slot assignment, child base and triangle base the node already has.
__global__ void k_refit_level(float4* __restrict__ nodes, const float4* __restrict__ tris, float* __restrict__ node_bounds,
                              uint32_t begin, uint32_t end, const uint32_t* __restrict__ maxabs_bits)
{
    const uint32_t node = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if(node >= end) return;
    const float eps = __uint_as_float(*maxabs_bits) * (1.0f / 65536.0f);
    float4* out = nodes + TR_BVH_NODE_VEC4 * (size_t)node;
    const float4 n1 = out[1];
    const uint32_t child_base = __float_as_uint(n1.x), tri_base = __float_as_uint(n1.y);
    const uint32_t imask = __float_as_uint(n1.z) & 0xFFu;            // octant 0 = slot order
    const uint32_t leaf8 = __float_as_uint(n1.w) & 0xFFu, two8 = (__float_as_uint(n1.w) >> 8) & 0xFFu;
    float clo[8][3], chi[8][3];
    int slot_child[8];
    uint32_t count_of_slot[8];
    float elo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ehi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int n = 0;
    for(int s = 0; s < 8; ++s)
    {
        slot_child[s] = -1;
        count_of_slot[s] = 0u;
        const uint32_t below = (1u << s) - 1u;
        float lo[3], hi[3];
        if((imask >> s) & 1u)
        {
            const float* b = node_bounds + 6 * (size_t)(child_base + __popc(imask & below));
            for(int a = 0; a < 3; ++a) { lo[a] = b[a]; hi[a] = b[3 + a]; }
        }
        else if((leaf8 >> s) & 1u)
        {
            const uint32_t first = tri_base + __popc(leaf8 & below) + __popc(two8 & below);
            const uint32_t cnt = 1u + ((two8 >> s) & 1u);
            for(int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
            for(uint32_t t = 0; t < 3u * cnt; ++t)
            {
                const float4 v = tris[3 * (size_t)first + t];
                lo[0] = fminf(lo[0], v.x); lo[1] = fminf(lo[1], v.y); lo[2] = fminf(lo[2], v.z);
                hi[0] = fmaxf(hi[0], v.x); hi[1] = fmaxf(hi[1], v.y); hi[2] = fmaxf(hi[2], v.z);
            }
            count_of_slot[s] = cnt;
        }
        else continue;
        slot_child[s] = s;
        n++;
        for(int a = 0; a < 3; ++a)
        {
            clo[s][a] = lo[a] - eps; chi[s][a] = hi[a] + eps;
            elo[a] = fminf(elo[a], lo[a]); ehi[a] = fmaxf(ehi[a], hi[a]);
        }
    }
    float* nb = node_bounds + 6 * (size_t)node;
    if(n == 0)
    {
        for(int a = 0; a < 3; ++a) { nb[a] = 0.0f; nb[3 + a] = 0.0f; }
        return;                                                    // the empty scene's only node stays as it is
    }
    float nlo[3], nhi[3];
    for(int a = 0; a < 3; ++a) { nlo[a] = elo[a] - eps; nhi[a] = ehi[a] + eps; nb[a] = elo[a]; nb[3 + a] = ehi[a]; }
    write_wide_node(out, nlo, nhi, n, clo, chi, slot_child, count_of_slot, imask, child_base, tri_base);
}

__global__ void k_init_bounds(uint32_t* bounds)
{
    if(threadIdx.x < 3) bounds[threadIdx.x] = 0xFFFFFFFFu;
    else if(threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

} // namespace trd

using namespace trd;

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Build temporaries: one arena (a single cudaMalloc / cudaFree -- both synchronise the device and dominated
// the build time when every array had its own) with a bump pointer; anything that does not fit falls back to
// its own allocation.
struct ScopedFree
{
    std::vector<void*> ptrs;
    char* arena = nullptr;
    size_t arena_size = 0, arena_used = 0;
    ~ScopedFree() { for(void* p: ptrs) cudaFree(p); if(arena) cudaFree(arena); }
    cudaError_t reserve(size_t bytes)
    {
        cudaError_t e = cudaMalloc((void**)&arena, bytes);
        if(e != cudaSuccess) { arena = nullptr; cudaGetLastError(); return cudaSuccess; }   // fall back to per-array allocations
        arena_size = bytes;
        return cudaSuccess;
    }
    template<typename T> cudaError_t alloc(T** p, size_t count)
    {
        const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~(size_t)255;
        if(arena && arena_used + bytes <= arena_size)
        {
            *p = reinterpret_cast<T*>(arena + arena_used);
            arena_used += bytes;
            return cudaSuccess;
        }
        cudaError_t e = cudaMalloc((void**)p, bytes);
        if(e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

// optional phase timing of the build (TR_BUILD_TIMING=1 prints to stderr)
struct PhaseTimer
{
    cudaStream_t stream;
    bool on;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    explicit PhaseTimer(cudaStream_t s): stream(s), on(getenv("TR_BUILD_TIMING") != nullptr) {}
    void mark(const char* name)
    {
        if(!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, stream);
        marks.push_back({name, e});
    }
    ~PhaseTimer()
    {
        if(!on) return;
        cudaStreamSynchronize(stream);
        for(size_t i = 1; i < marks.size(); ++i)
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, "[tr_build] %-12s %.3f ms\n", marks[i - 1].first, ms);
        }
        for(auto& m: marks) cudaEventDestroy(m.second);
    }
};

cudaError_t tr_build_pre_transform(cudaStream_t stream, const tr_vertex* in, tr_vertex* out, uint32_t count,
                                   const float* model_dev, const float* model_normal_dev)
{
    if(count == 0) return cudaSuccess;
    k_pre_transform<<<div_up(count, 256), 256, 0, stream>>>(in, out, count, model_dev, model_normal_dev);
    return cudaGetLastError();
}

cudaError_t tr_build_shade_tris(cudaStream_t stream, const tr_vertex* verts, const uint32_t* indices, uint32_t triangle_count, float4* out)
{
    if(triangle_count == 0) return cudaSuccess;
    k_shade_tris<<<div_up(triangle_count, 256), 256, 0, stream>>>(verts, indices, triangle_count, out);
    return cudaGetLastError();
}

cudaError_t tr_build_extract_tri_lights(cudaStream_t stream, const tr_vertex* verts, const uint32_t* indices,
                                        uint32_t triangle_count, uint32_t instance_id, int32_t light_base_id,
                                        const float emission[3], int32_t emission_tex_id, tr_tri_light* out)
{
    if(triangle_count == 0) return cudaSuccess;
    k_extract_tri_lights<<<div_up(triangle_count, 256), 256, 0, stream>>>(
        verts, indices, triangle_count, instance_id, light_base_id, emission[0], emission[1], emission[2], emission_tex_id, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// environment map importances: shader/alias_table_importance.comp:19-31 with
// latlong_texel_solid_angle (shader/alias_table.glsl:15-20)
// ---------------------------------------------------------------------------------------------
__global__ void k_envmap_importance(const void* __restrict__ texels, uint32_t format, int sx, int sy, float* __restrict__ importance)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if(x >= sx || y >= sy) return;
    const size_t i = (size_t)y * sx + x;
    vec3 c;
    if(format == TR_TEXTURE_RGBA32F)
    {
        const float4 f = reinterpret_cast<const float4*>(texels)[i];
        c = V3(f.x, f.y, f.z);
    }
    else
    {
        const uint2 h = reinterpret_cast<const uint2*>(texels)[i];
        c = V3(half_to_float_(h.x & 0xFFFFu), half_to_float_(h.x >> 16), half_to_float_(h.y & 0xFFFFu));
    }
    const float y0 = (float)y / (float)sy, y1 = (float)(y + 1) / (float)sy;
    const float solid_angle = 2.0f * TR_PI * (cos_(TR_PI * y0) - cos_(TR_PI * y1)) / (float)sx;
    importance[i] = dot(c, V3(0.2126f, 0.7152f, 0.0722f)) * solid_angle;
}

cudaError_t tr_build_envmap_importance(cudaStream_t stream, const void* texels, uint32_t format, uint32_t width, uint32_t height,
                                       float* importance_dev)
{
    dim3 block(16, 16), grid(div_up(width, 16), div_up(height, 16));
    k_envmap_importance<<<grid, block, 0, stream>>>(texels, format, (int)width, (int)height, importance_dev);
    return cudaGetLastError();
}

cudaError_t tr_build_bvh(cudaStream_t stream, const tr_build_input& in, tr_build_output* out)
{
    const uint32_t n = in.triangle_count;
    out->nodes = nullptr; out->tris = nullptr; out->node_count = 0; out->depth = 0; out->level_offsets.clear();
    ScopedFree tmp;
    PhaseTimer timer(stream);
    {
        // 2 x 48 n triangle records, 80 n node upper bound, ~115 n sort/hierarchy arrays (+ ~100 n for PLOC)
        const bool big = getenv("TR_TUNE_BUILDER") && std::strcmp(getenv("TR_TUNE_BUILDER"), "lbvh") != 0;
        CUDA_OK(tmp.reserve((size_t)n * ((big || in.quality) ? 420 : 300) + (4u << 20)));
    }
    float4* tris_unsorted = nullptr;
    uint32_t* bounds = nullptr;
    CUDA_OK(tmp.alloc(&tris_unsorted, 3 * (size_t)n));
    CUDA_OK(tmp.alloc(&bounds, 8));
    timer.mark("gather");
    k_init_bounds<<<1, 32, 0, stream>>>(bounds);
    for(uint32_t i = 0; i < in.instance_count; ++i)
    {
        const tr_build_instance& inst = in.instances[i];
        if(inst.triangle_count == 0) continue;
        k_gather<<<div_up(inst.triangle_count, 256), 256, 0, stream>>>(
            in.vertices + inst.vertex_base, in.indices + inst.index_base, inst.triangle_count, i, inst.opaque,
            inst.triangle_base, tris_unsorted, bounds);
    }
    CUDA_OK(cudaGetLastError());

    // Morton-ordered records and the nodes at their upper bound (n - 1 inner nodes, +1 for the degenerate root)
    // are temporaries; the results are copied into exact-size allocations at the end.
    float4* tris_sorted = nullptr;
    CUDA_OK(tmp.alloc(&tris_sorted, 3 * (size_t)n));
    float4* nodes = nullptr;
    const size_t max_nodes = std::max<size_t>(n, 2);
    CUDA_OK(tmp.alloc(&nodes, max_nodes * TR_BVH_NODE_VEC4));
    auto fail = [&](cudaError_t err) { return err; };
#define CUDA_OK2(x) do { cudaError_t e2_ = (x); if(e2_ != cudaSuccess) return fail(e2_); } while(0)

    uint32_t hb[6];
    CUDA_OK2(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    CUDA_OK2(cudaStreamSynchronize(stream));
    float maxabs = 0.0f;
    for(int k = 0; k < 6; ++k)
    {
        uint32_t u = hb[k];
        uint32_t bits = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
        float f;
        memcpy(&f, &bits, 4);
        if(n > 0) maxabs = std::max(maxabs, std::fabs(f));
    }
    const float eps = maxabs * (1.0f / 65536.0f);
    out->eps = eps;

    if(n == 0)
    {
        // empty scene: one node without children
        float4 *nodes_final = nullptr, *tris_empty = nullptr;
        CUDA_OK2(cudaMalloc((void**)&nodes_final, TR_BVH_NODE_VEC4 * sizeof(float4)));
        cudaError_t e0 = cudaMalloc((void**)&tris_empty, 3 * sizeof(float4));
        if(e0 != cudaSuccess) { cudaFree(nodes_final); return e0; }
        k_wide_single<<<1, 1, 0, stream>>>(0, nullptr, nullptr, eps, nullptr, nodes_final, tris_empty);
        e0 = cudaGetLastError();
        if(e0 == cudaSuccess) e0 = cudaStreamSynchronize(stream);
        if(e0 != cudaSuccess) { cudaFree(nodes_final); cudaFree(tris_empty); return e0; }
        out->nodes = nodes_final; out->tris = tris_empty; out->node_count = 1; out->depth = 1;
        out->level_offsets = {0u, 1u};
        return cudaSuccess;
    }

    // builder: the quality build (binned SAH) for static scenes, the radix tree for per-frame rebuilds;
    // TR_TUNE_BUILDER=lbvh|sah|ploc overrides
    const char* builder_env = getenv("TR_TUNE_BUILDER");
    const bool use_sah = n >= 2 && (builder_env ? std::strcmp(builder_env, "sah") == 0 : in.quality != 0);
    int2* children = nullptr;
    int* counts = nullptr;
    float *leaf_boxes = nullptr, *inner_boxes = nullptr;
    int root_node = 0;
    if(use_sah)
    {
        float* boxes = nullptr;
        uint32_t *perm_a = nullptr, *perm_b = nullptr, *perm_final = nullptr;
        SahTask *tasks_a = nullptr, *tasks_b = nullptr;
        int *next_count = nullptr, *node_counter = nullptr;
        CUDA_OK2(tmp.alloc(&boxes, 6 * (size_t)n));
        CUDA_OK2(tmp.alloc(&perm_a, n)); CUDA_OK2(tmp.alloc(&perm_b, n)); CUDA_OK2(tmp.alloc(&perm_final, n));
        CUDA_OK2(tmp.alloc(&tasks_a, n / 2 + 2)); CUDA_OK2(tmp.alloc(&tasks_b, n / 2 + 2));
        CUDA_OK2(tmp.alloc(&next_count, 1)); CUDA_OK2(tmp.alloc(&node_counter, 1));
        CUDA_OK2(tmp.alloc(&children, n)); CUDA_OK2(tmp.alloc(&counts, n));
        CUDA_OK2(tmp.alloc(&leaf_boxes, 6 * (size_t)n)); CUDA_OK2(tmp.alloc(&inner_boxes, 6 * (size_t)n));
        timer.mark("sah");
        k_sah_boxes<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, n, boxes, perm_a);
        SahTask root;
        root.begin = 0; root.end = (int)n; root.id = 0;
        for(int k = 0; k < 3; ++k)
        {
            const uint32_t ul = hb[k], uh = hb[3 + k];
            const uint32_t bl = (ul & 0x80000000u) ? (ul & 0x7FFFFFFFu) : ~ul, bh = (uh & 0x80000000u) ? (uh & 0x7FFFFFFFu) : ~uh;
            memcpy(&root.lo[k], &bl, 4); memcpy(&root.hi[k], &bh, 4);
        }
        const int one = 1;
        CUDA_OK2(cudaMemcpyAsync(tasks_a, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
        CUDA_OK2(cudaMemcpyAsync(node_counter, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
        // Split axis: the axis (level mod 3) is preferred -- another axis wins only if its SAH cost is 1.5x lower. That keeps
        // the three binary levels a wide node swallows close to a 2x2x2 arrangement, which is what the octant-ordered
        // traversal of the 8-wide nodes visits front to back. Measured on the atrium (closest / shadow ms per frame):
        // radix tree 11.29 / 5.45, free axis choice 11.78 / 4.86, strict cycling 11.12 / 5.24, preference 1.5x 10.83 / 5.11.
        static const int sah_axis_mode = getenv("TR_TUNE_SAH_AXES") ? atoi(getenv("TR_TUNE_SAH_AXES")) : 1;
        static const float sah_axis_penalty = getenv("TR_TUNE_SAH_PENALTY") ? (float)atof(getenv("TR_TUNE_SAH_PENALTY")) : 1.5f;
        int task_count = 1, level = 0;
        while(task_count > 0)
        {
            CUDA_OK2(cudaMemsetAsync(next_count, 0, sizeof(int), stream));
            // beyond 48 levels only median splits: bounds the depth of pathological inputs (<= 48 + log2(n) binary levels)
            k_sah_level<<<task_count, SAH_BLOCK, 0, stream>>>(tasks_a, task_count, boxes, perm_a, perm_b, perm_final, children, counts, inner_boxes,
                                                             tasks_b, next_count, node_counter, level >= 48 ? 1 : 0, sah_axis_mode, level, sah_axis_penalty);
            CUDA_OK2(cudaMemcpyAsync(&task_count, next_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_OK2(cudaStreamSynchronize(stream));
            std::swap(tasks_a, tasks_b);
            std::swap(perm_a, perm_b);
            if(++level > 4096) return fail(cudaErrorUnknown);
        }
        k_sah_finalize<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, boxes, perm_final, n, tris_sorted, leaf_boxes);
        CUDA_OK2(cudaGetLastError());
        root_node = 0;
    }
    unsigned long long *keys_a = nullptr, *keys_b = nullptr;
    uint32_t *vals_a = nullptr, *vals_b = nullptr, *hist = nullptr;
    const uint32_t tiles = div_up(n, SORT_TILE);
    if(!use_sah)
    {
    CUDA_OK2(tmp.alloc(&keys_a, n)); CUDA_OK2(tmp.alloc(&keys_b, n));
    CUDA_OK2(tmp.alloc(&vals_a, n)); CUDA_OK2(tmp.alloc(&vals_b, n));
    CUDA_OK2(tmp.alloc(&hist, 256 * (size_t)tiles));
    timer.mark("morton+sort");
    k_morton<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, n, bounds, keys_a, vals_a);
    for(int pass = 0; pass < 8; ++pass)
    {
        const int shift = pass * 8;
        k_sort_hist<<<tiles, SORT_THREADS, 0, stream>>>(keys_a, n, shift, hist, tiles);
        k_sort_scan<<<1, 1024, 0, stream>>>(hist, 256 * tiles);
        k_sort_scatter<<<tiles, SORT_THREADS, 0, stream>>>(keys_a, vals_a, keys_b, vals_b, n, shift, hist, tiles);
        std::swap(keys_a, keys_b);
        std::swap(vals_a, vals_b);
    }
    CUDA_OK2(cudaGetLastError());
    k_reorder_tris<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, vals_a, n, tris_sorted);

    int2* ranges = nullptr;
    int *inner_parent = nullptr, *leaf_parent = nullptr, *flags = nullptr;
    CUDA_OK2(tmp.alloc(&children, n)); CUDA_OK2(tmp.alloc(&ranges, n));
    CUDA_OK2(tmp.alloc(&inner_parent, n)); CUDA_OK2(tmp.alloc(&leaf_parent, n)); CUDA_OK2(tmp.alloc(&flags, n));
    CUDA_OK2(tmp.alloc(&leaf_boxes, 6 * (size_t)n)); CUDA_OK2(tmp.alloc(&inner_boxes, 6 * (size_t)n));
    CUDA_OK2(cudaMemsetAsync(flags, 0, sizeof(int) * n, stream));
    timer.mark("hierarchy");
    static const bool use_karras = !(getenv("TR_TUNE_BUILDER") && std::strcmp(getenv("TR_TUNE_BUILDER"), "ploc") == 0);
    CUDA_OK2(tmp.alloc(&counts, n));
    if(use_karras || n < 2)
    {
        if(n > 1) k_hierarchy<<<div_up(n - 1, 256), 256, 0, stream>>>(keys_a, (int)n, children, ranges, inner_parent, leaf_parent);
        k_refit<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, vals_a, (int)n, children, inner_parent, leaf_parent, leaf_boxes, inner_boxes, flags, 1);
        if(n > 1) k_counts_from_ranges<<<div_up(n - 1, 256), 256, 0, stream>>>((int)n - 1, ranges, counts);
        CUDA_OK2(cudaGetLastError());
    }
    else
    {
        k_refit<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, vals_a, (int)n, children, inner_parent, leaf_parent, leaf_boxes, inner_boxes, flags, 0);
        float4 *lo_a = nullptr, *hi_a = nullptr, *lo_b = nullptr, *hi_b = nullptr;
        int *node_a = nullptr, *node_b = nullptr, *nn = nullptr, *valid = nullptr, *node_counter = nullptr;
        uint32_t* block_sums = nullptr;
        CUDA_OK2(tmp.alloc(&lo_a, n)); CUDA_OK2(tmp.alloc(&hi_a, n)); CUDA_OK2(tmp.alloc(&lo_b, n)); CUDA_OK2(tmp.alloc(&hi_b, n));
        CUDA_OK2(tmp.alloc(&node_a, n)); CUDA_OK2(tmp.alloc(&node_b, n)); CUDA_OK2(tmp.alloc(&nn, n)); CUDA_OK2(tmp.alloc(&valid, n));
        CUDA_OK2(tmp.alloc(&node_counter, 1)); CUDA_OK2(tmp.alloc(&block_sums, div_up(n, PLOC_BLOCK) + 1));
        CUDA_OK2(cudaMemsetAsync(node_counter, 0, sizeof(int), stream));
        k_ploc_init<<<div_up(n, 256), 256, 0, stream>>>((int)n, leaf_boxes, lo_a, hi_a, node_a);
        uint32_t m = n;
        int rounds = 0;
        while(m > 1)
        {
            const uint32_t blocks = div_up(m, PLOC_BLOCK);
            k_ploc_nn<<<blocks, PLOC_BLOCK, 0, stream>>>((int)m, lo_a, hi_a, nn);
            k_ploc_merge<<<blocks, PLOC_BLOCK, 0, stream>>>((int)m, nn, node_a, lo_a, hi_a, children, inner_boxes, counts, node_counter, valid);
            k_ploc_count<<<blocks, PLOC_BLOCK, 0, stream>>>((int)m, valid, block_sums);
            k_sort_scan<<<1, 1024, 0, stream>>>(block_sums, blocks + 1);
            k_ploc_compact<<<blocks, PLOC_BLOCK, 0, stream>>>((int)m, valid, block_sums, node_a, lo_a, hi_a, node_b, lo_b, hi_b);
            uint32_t next_m = 0;
            CUDA_OK2(cudaMemcpyAsync(&next_m, block_sums + blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CUDA_OK2(cudaStreamSynchronize(stream));
            if(next_m == 0 || next_m >= m) return fail(cudaErrorUnknown);   // cannot happen: the closest pair always merges
            m = next_m;
            std::swap(node_a, node_b); std::swap(lo_a, lo_b); std::swap(hi_a, hi_b);
            if(++rounds > 4096) return fail(cudaErrorUnknown);
        }
        CUDA_OK2(cudaMemcpyAsync(&root_node, node_a, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_OK2(cudaStreamSynchronize(stream));
    }
    }   // !use_sah

    timer.mark("wide");
    static const int leaf_tris_env = getenv("TR_TUNE_LEAF") ? atoi(getenv("TR_TUNE_LEAF")) : TR_BVH_LEAF_TRIS;
    const int leaf_tris = std::max(1, std::min(TR_BVH_LEAF_TRIS, leaf_tris_env));
    // subtrees of at most this many triangles become one bottom node of (up to 8) leaf slots; 0 = off (TR_TUNE_BOTTOM)
    // measured on the atrium (ms/frame, triangle tests per closest ray): off 20.54 / 5.50, 8: 20.31 / 5.07, 12: 20.27 / 4.98,
    // 16: 20.30 / 5.07 -- node steps stay at 17.5, the gain is the tighter leaf boxes of the optimal cut
    static const int bottom_tris = getenv("TR_TUNE_BOTTOM") ? atoi(getenv("TR_TUNE_BOTTOM")) : 12;
    // the greedy opening leaves subtrees of at most this many triangles closed (they become ONE bottom node at the next level
    // instead of being cut into several tiny ones); TR_TUNE_KEEP_CLOSED, 0 = open everything above the leaf size. Measured on
    // the atrium (TR_DEBUG_BVH_STATS: 201 k of the 261 k nodes are bottom nodes, 52 k of them with only 2-3 leaf slots): keeping
    // them closed costs more than it saves -- 0: 20.10 ms / 17.49 node steps per closest ray, 4: 20.30 / 17.98, 8: 20.34 / 18.11,
    // 12: 20.43 / 18.04 -- the parent's node step culls eight small tight boxes better than one visit of a fuller node
    static const int keep_closed_env = getenv("TR_TUNE_KEEP_CLOSED") ? atoi(getenv("TR_TUNE_KEEP_CLOSED")) : 0;
    const int keep_closed = std::max(leaf_tris, std::min(keep_closed_env, bottom_tris));
    uint32_t node_count = 0, depth = 0;
    // tris_sorted holds the Morton order; the wide emission re-packs triangles per node into tris_final
    float4* tris_final = nullptr;
    CUDA_OK2(cudaMalloc((void**)&tris_final, std::max<size_t>(3 * (size_t)n, 1) * sizeof(float4)));
    auto fail2 = [&](cudaError_t err) { cudaFree(tris_final); return fail(err); };
#define CUDA_OK3(x) do { cudaError_t e3_ = (x); if(e3_ != cudaSuccess) return fail2(e3_); } while(0)
    if(n <= (uint32_t)leaf_tris)
    {
        k_wide_single<<<1, 1, 0, stream>>>((int)n, leaf_boxes, inner_boxes, eps, tris_sorted, nodes, tris_final);
        node_count = 1; depth = 1;
        out->level_offsets = {0u, 1u};
    }
    else
    {
        int *frontier_a = nullptr, *frontier_b = nullptr, *next_count = nullptr;
        uint32_t* tri_counter = nullptr;
        CUDA_OK3(tmp.alloc(&frontier_a, n)); CUDA_OK3(tmp.alloc(&frontier_b, n)); CUDA_OK3(tmp.alloc(&next_count, 1));
        CUDA_OK3(tmp.alloc(&tri_counter, 1));
        CUDA_OK3(cudaMemcpyAsync(frontier_a, &root_node, sizeof(int), cudaMemcpyHostToDevice, stream));
        CUDA_OK3(cudaMemsetAsync(tri_counter, 0, sizeof(uint32_t), stream));
        int frontier_size = 1;
        int out_base = 0;
        out->level_offsets.assign(1, 0u);
        while(frontier_size > 0)
        {
            CUDA_OK3(cudaMemsetAsync(next_count, 0, sizeof(int), stream));
            const int next_out_base = out_base + frontier_size;
            k_wide_level<<<div_up((uint32_t)frontier_size, 128), 128, 0, stream>>>(
                leaf_tris, frontier_a, frontier_size, out_base, children, counts, leaf_boxes, inner_boxes, eps, tris_sorted, nodes,
                tris_final, frontier_b, next_count, next_out_base, tri_counter, bottom_tris, keep_closed);
            int next_size = 0;
            CUDA_OK3(cudaMemcpyAsync(&next_size, next_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_OK3(cudaStreamSynchronize(stream));
            out_base = next_out_base;
            out->level_offsets.push_back((uint32_t)out_base);
            frontier_size = next_size;
            std::swap(frontier_a, frontier_b);
            depth++;
        }
        node_count = (uint32_t)out_base;
    }
    timer.mark("finalize");
    // exact-size node array for the caller (the upper-bound array lives in the arena)
    float4* nodes_final = nullptr;
    CUDA_OK3(cudaMalloc((void**)&nodes_final, (size_t)node_count * TR_BVH_NODE_VEC4 * sizeof(float4)));
    cudaError_t ef = cudaMemcpyAsync(nodes_final, nodes, (size_t)node_count * TR_BVH_NODE_VEC4 * sizeof(float4), cudaMemcpyDeviceToDevice, stream);
    if(ef == cudaSuccess) ef = cudaStreamSynchronize(stream);
    if(ef != cudaSuccess) { cudaFree(nodes_final); return fail2(ef); }
    timer.mark("end");
    if(getenv("TR_DEBUG_BVH_STATS"))
    {
        // development aid: how full are the wide nodes, level by level
        std::vector<float4> h((size_t)node_count * TR_BVH_NODE_VEC4);
        cudaMemcpy(h.data(), nodes_final, h.size() * sizeof(float4), cudaMemcpyDeviceToHost);
        const std::vector<uint32_t>& lo = out->level_offsets;
        for(size_t l = 0; l + 1 < lo.size(); ++l)
        {
            unsigned long long inner = 0, leaf = 0, two = 0, hist[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bottom = 0;
            for(uint32_t k = lo[l]; k < lo[l + 1]; ++k)
            {
                uint32_t wz, ww;
                memcpy(&wz, &h[(size_t)k * TR_BVH_NODE_VEC4 + 1].z, 4); memcpy(&ww, &h[(size_t)k * TR_BVH_NODE_VEC4 + 1].w, 4);
                const int ni = __builtin_popcount(wz & 0xFFu), nl = __builtin_popcount(ww & 0xFFu), nt = __builtin_popcount((ww >> 8) & 0xFFu);
                inner += ni; leaf += nl; two += nt; hist[ni + nl]++;
                if(ni == 0) bottom++;
            }
            const double cnt = (double)(lo[l + 1] - lo[l]);
            fprintf(stderr, "bvh level %2zu: %8u nodes (%7llu without inner children)  inner/node %.2f  leaf slots/node %.2f  tris/node %.2f  children histogram", l,
                    lo[l + 1] - lo[l], bottom, inner / cnt, leaf / cnt, (leaf + two) / cnt);
            for(int c = 1; c <= 8; ++c) fprintf(stderr, " %d:%llu", c, hist[c]);
            fprintf(stderr, "\n");
        }
    }
    nodes = nodes_final;
    tris_sorted = tris_final;
    out->nodes = nodes;
    out->tris = tris_sorted;
    out->node_count = node_count;
    out->depth = depth;
    return cudaSuccess;
}

cudaError_t tr_build_refit_bvh(cudaStream_t stream, const tr_build_input& in, const tr_build_instance* instances_dev,
                               float4* nodes, float4* tris, uint32_t node_count, const std::vector<uint32_t>& level_offsets,
                               float* scratch)
{
    if(level_offsets.size() < 2 || level_offsets.back() != node_count) return cudaErrorInvalidValue;
    uint32_t* maxabs_bits = reinterpret_cast<uint32_t*>(scratch + 6 * (size_t)node_count);
    cudaError_t e = cudaMemsetAsync(maxabs_bits, 0, sizeof(uint32_t), stream);
    if(e != cudaSuccess) return e;
    const uint32_t n = in.triangle_count;
    if(n) k_refit_leaf_tris<<<div_up(n, 256), 256, 0, stream>>>(in.vertices, in.indices, instances_dev, n, tris, maxabs_bits);
    for(size_t l = level_offsets.size() - 1; l-- > 0;)
    {
        const uint32_t begin = level_offsets[l], end = level_offsets[l + 1];
        if(end > begin) k_refit_level<<<div_up(end - begin, 128), 128, 0, stream>>>(nodes, tris, scratch, begin, end, maxabs_bits);
    }
    return cudaGetLastError();
}
