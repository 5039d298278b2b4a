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
//   8. emit              breadth-first re-layout, subtrees of <= TR_BVH_MAX_LEAF triangles collapsed
//                        into leaves, boxes inflated by eps = 2^-16 * max|coordinate|
#include "tr_scene.cuh"
#include "tr_math.cuh"
#include "tr_build.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

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
                        float* __restrict__ inner_boxes /* (n-1)*6 */, int* __restrict__ flags)
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
    if(n == 1) return;
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
// 8. breadth-first emission with leaf collapse
// ---------------------------------------------------------------------------------------------
// frontier entry: Karras inner node id; its output slot is (level_base + position in frontier).
__device__ __forceinline__ bool is_leaf_range(int2 range) { return range.y - range.x + 1 <= TR_BVH_MAX_LEAF; }

__global__ void k_emit_level(const int* __restrict__ frontier, int frontier_size, int out_base,
                             const int2* __restrict__ children, const int2* __restrict__ ranges,
                             const float* __restrict__ leaf_boxes, const float* __restrict__ inner_boxes,
                             float eps, float4* __restrict__ nodes, int* __restrict__ next_frontier,
                             int* __restrict__ next_count, int next_out_base)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= frontier_size) return;
    const int node = frontier[i];
    const int2 ch = children[node];
    int ref[2];
    float lo[2][3], hi[2][3];
#pragma unroll
    for(int c = 0; c < 2; ++c)
    {
        const int child = c == 0 ? ch.x : ch.y;
        const float* box;
        if(child < 0)
        {
            const int leaf = ~child;
            ref[c] = ~((leaf << 3) | 0);
            box = leaf_boxes + 6 * (size_t)leaf;
        }
        else
        {
            const int2 rg = ranges[child];
            box = inner_boxes + 6 * (size_t)child;
            if(is_leaf_range(rg)) ref[c] = ~((rg.x << 3) | (rg.y - rg.x));
            else
            {
                int slot = atomicAdd(next_count, 1);
                next_frontier[slot] = child;
                ref[c] = next_out_base + slot;
            }
        }
#pragma unroll
        for(int k = 0; k < 3; ++k) { lo[c][k] = box[k] - eps; hi[c][k] = box[3 + k] + eps; }
    }
    float4* o = nodes + 4 * (size_t)(out_base + i);
    o[0] = make_float4(lo[0][0], hi[0][0], lo[0][1], hi[0][1]);
    o[1] = make_float4(lo[1][0], hi[1][0], lo[1][1], hi[1][1]);
    o[2] = make_float4(lo[0][2], hi[0][2], lo[1][2], hi[1][2]);
    o[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), 0.0f, 0.0f);
}

__global__ void k_emit_single_leaf(int n, const float* __restrict__ leaf_boxes, const float* __restrict__ inner_boxes,
                                   float eps, float4* __restrict__ nodes)
{
    // whole scene fits one leaf: root = { leaf(0, n), absent }
    const float* box = n == 1 ? leaf_boxes : inner_boxes;
    const float inf = __uint_as_float(0x7F800000u);
    nodes[0] = make_float4(box[0] - eps, box[3] + eps, box[1] - eps, box[4] + eps);
    nodes[1] = make_float4(inf, -inf, inf, -inf);
    nodes[2] = make_float4(box[2] - eps, box[5] + eps, inf, -inf);
    nodes[3] = make_float4(__int_as_float(~((0 << 3) | (n - 1))), __int_as_float(~0), 0.0f, 0.0f);
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

__global__ void k_init_bounds(uint32_t* bounds)
{
    if(threadIdx.x < 3) bounds[threadIdx.x] = 0xFFFFFFFFu;
    else if(threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

} // namespace trd

using namespace trd;

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

struct ScopedFree
{
    std::vector<void*> ptrs;
    ~ScopedFree() { for(void* p: ptrs) cudaFree(p); }
    template<typename T> cudaError_t alloc(T** p, size_t count)
    {
        cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
        if(e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

cudaError_t tr_build_pre_transform(cudaStream_t stream, const tr_vertex* in, tr_vertex* out, uint32_t count,
                                   const float* model_dev, const float* model_normal_dev)
{
    if(count == 0) return cudaSuccess;
    k_pre_transform<<<div_up(count, 256), 256, 0, stream>>>(in, out, count, model_dev, model_normal_dev);
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

cudaError_t tr_build_bvh(cudaStream_t stream, const tr_build_input& in, tr_build_output* out)
{
    const uint32_t n = in.triangle_count;
    out->nodes = nullptr; out->tris = nullptr; out->node_count = 0; out->depth = 0;
    ScopedFree tmp;
    float4* tris_unsorted = nullptr;
    uint32_t* bounds = nullptr;
    CUDA_OK(tmp.alloc(&tris_unsorted, 3 * (size_t)n));
    CUDA_OK(tmp.alloc(&bounds, 8));
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

    float4* tris_sorted = nullptr;
    CUDA_OK(cudaMalloc((void**)&tris_sorted, std::max<size_t>(3 * (size_t)n, 1) * sizeof(float4)));
    // node count upper bound: n - 1 inner nodes (+1 for the degenerate root)
    float4* nodes = nullptr;
    const size_t max_nodes = std::max<size_t>(n, 2);
    cudaError_t e = cudaMalloc((void**)&nodes, max_nodes * 4 * sizeof(float4));
    if(e != cudaSuccess) { cudaFree(tris_sorted); return e; }
    auto fail = [&](cudaError_t err) { cudaFree(tris_sorted); cudaFree(nodes); return err; };
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
        // empty scene: a root whose children are both absent
        const float inf = INFINITY;
        float4 h[4] = {make_float4(inf, -inf, inf, -inf), make_float4(inf, -inf, inf, -inf), make_float4(inf, -inf, inf, -inf),
                       make_float4(0, 0, 0, 0)};
        int r = ~0;
        memcpy(&h[3].x, &r, 4); memcpy(&h[3].y, &r, 4);
        CUDA_OK2(cudaMemcpyAsync(nodes, h, sizeof(h), cudaMemcpyHostToDevice, stream));
        CUDA_OK2(cudaStreamSynchronize(stream));
        out->nodes = nodes; out->tris = tris_sorted; out->node_count = 1; out->depth = 1;
        return cudaSuccess;
    }

    unsigned long long *keys_a = nullptr, *keys_b = nullptr;
    uint32_t *vals_a = nullptr, *vals_b = nullptr, *hist = nullptr;
    const uint32_t tiles = div_up(n, SORT_TILE);
    CUDA_OK2(tmp.alloc(&keys_a, n)); CUDA_OK2(tmp.alloc(&keys_b, n));
    CUDA_OK2(tmp.alloc(&vals_a, n)); CUDA_OK2(tmp.alloc(&vals_b, n));
    CUDA_OK2(tmp.alloc(&hist, 256 * (size_t)tiles));
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

    int2 *children = nullptr, *ranges = nullptr;
    int *inner_parent = nullptr, *leaf_parent = nullptr, *flags = nullptr;
    float *leaf_boxes = nullptr, *inner_boxes = nullptr;
    CUDA_OK2(tmp.alloc(&children, n)); CUDA_OK2(tmp.alloc(&ranges, n));
    CUDA_OK2(tmp.alloc(&inner_parent, n)); CUDA_OK2(tmp.alloc(&leaf_parent, n)); CUDA_OK2(tmp.alloc(&flags, n));
    CUDA_OK2(tmp.alloc(&leaf_boxes, 6 * (size_t)n)); CUDA_OK2(tmp.alloc(&inner_boxes, 6 * (size_t)n));
    CUDA_OK2(cudaMemsetAsync(flags, 0, sizeof(int) * n, stream));
    if(n > 1) k_hierarchy<<<div_up(n - 1, 256), 256, 0, stream>>>(keys_a, (int)n, children, ranges, inner_parent, leaf_parent);
    k_refit<<<div_up(n, 256), 256, 0, stream>>>(tris_unsorted, vals_a, (int)n, children, inner_parent, leaf_parent, leaf_boxes, inner_boxes, flags);
    CUDA_OK2(cudaGetLastError());

    uint32_t node_count = 0, depth = 0;
    if(n <= TR_BVH_MAX_LEAF)
    {
        k_emit_single_leaf<<<1, 1, 0, stream>>>((int)n, leaf_boxes, inner_boxes, eps, nodes);
        node_count = 1; depth = 1;
    }
    else
    {
        int *frontier_a = nullptr, *frontier_b = nullptr, *next_count = nullptr;
        CUDA_OK2(tmp.alloc(&frontier_a, n)); CUDA_OK2(tmp.alloc(&frontier_b, n)); CUDA_OK2(tmp.alloc(&next_count, 1));
        int zero = 0;
        CUDA_OK2(cudaMemcpyAsync(frontier_a, &zero, sizeof(int), cudaMemcpyHostToDevice, stream));
        int frontier_size = 1;
        int out_base = 0;
        while(frontier_size > 0)
        {
            CUDA_OK2(cudaMemsetAsync(next_count, 0, sizeof(int), stream));
            const int next_out_base = out_base + frontier_size;
            k_emit_level<<<div_up((uint32_t)frontier_size, 256), 256, 0, stream>>>(
                frontier_a, frontier_size, out_base, children, ranges, leaf_boxes, inner_boxes, eps, nodes,
                frontier_b, next_count, next_out_base);
            int next_size = 0;
            CUDA_OK2(cudaMemcpyAsync(&next_size, next_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_OK2(cudaStreamSynchronize(stream));
            out_base = next_out_base;
            frontier_size = next_size;
            std::swap(frontier_a, frontier_b);
            depth++;
        }
        node_count = (uint32_t)out_base;
    }
    CUDA_OK2(cudaStreamSynchronize(stream));
    out->nodes = nodes;
    out->tris = tris_sorted;
    out->node_count = node_count;
    out->depth = depth;
    return cudaSuccess;
}
