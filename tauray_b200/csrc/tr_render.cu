// tr_render.cu -- the render kernels: the replacement of rt_pipeline::trace_rays
// (src/rt_pipeline.cc:13-19) + shader/path_tracer.rgen for one pass.
//
// Wavefront organisation (kernel_mode = TR_KERNEL_WAVEFRONT), per pass and bounce b:
//   k_wf_generate       camera rays for every launch slot                    -> queue[0]
//   k_wf_trace_closest  closest-hit traversal of queue[b&1]                  -> hits
//   k_wf_shade          hit shading, emission MIS, NEE sample, BSDF sample   -> queue[(b+1)&1], shadow_queue
//   k_wf_trace_shadow   any-hit traversal of shadow_queue, NEE accumulate
//   k_wf_finish         path_tracer.rgen:112 + running-mean accumulate (+ NVLink peer stores)
// Survivors are compacted into the next queue with a warp ballot + prefix popcount and ONE
// atomicAdd per warp, so later bounces run on dense warps. Queue sizes never visit the host:
// kernels are launched as persistent grids (SM count x resident blocks) that read the counts
// from HBM and stride over the queue.
// kernel_mode = TR_KERNEL_MEGAKERNEL runs the same device functions in one launch (state in
// registers); it exists as the A/B baseline for the wavefront numbers.
#include "tr_path.cuh"
#include "tr_render.h"
#include "tr_tma.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace trd {

constexpr int WF_BLOCK = 128;
#ifndef TRACE_BLOCK
#define TRACE_BLOCK 128          // threads per block of the persistent traversal kernel (64: measured)
#endif
#ifndef TRACE_MIN_BLOCKS_CLOSEST
#define TRACE_MIN_BLOCKS_CLOSEST 9
#endif
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 4
#endif
#ifndef TRACE_MIN_BLOCKS_SHADOW
#define TRACE_MIN_BLOCKS_SHADOW 9
#endif

// ---- SoA <-> registers ---------------------------------------------------------------------------
TRD void store_ray(const tr_wavefront& wf, uint32_t id, const PathState& s)
{
    wf.pos_pdf[id] = make_float4(s.pos.x, s.pos.y, s.pos.z, s.bsdf_pdf);
    wf.dir_seed[id] = make_float4(s.view.x, s.view.y, s.view.z, __uint_as_float(s.random_seed));
}
TRD void store_state(const tr_wavefront& wf, uint32_t id, const PathState& s)
{
    store_ray(wf, id, s);
    wf.atten_reg[id] = make_float4(s.attenuation.x, s.attenuation.y, s.attenuation.z, s.regularization);
    wf.lobes[id] = make_float4(s.primary_lobes.transmission, s.primary_lobes.diffuse,
                               s.primary_lobes.dielectric_reflection, s.primary_lobes.metallic_reflection);
    wf.diffuse[id] = make_float4(s.diffuse.x, s.diffuse.y, s.diffuse.z, s.aux_alpha);
    wf.reflection[id] = make_float4(s.reflection.x, s.reflection.y, s.reflection.z, 0.0f);
}
TRD void store_first_hit(const tr_wavefront& wf, uint32_t id, const PathState& s)
{
    wf.fh_emission[id] = make_float4(s.fh_emission.x, s.fh_emission.y, s.fh_emission.z, s.fh_albedo.w);
    wf.fh_albedo[id] = make_float4(s.fh_albedo.x, s.fh_albedo.y, s.fh_albedo.z, s.fh_metallic);
}
TRD void load_state(const tr_wavefront& wf, uint32_t id, PathState& s)
{
    const float4 a = wf.pos_pdf[id], b = wf.dir_seed[id], c = wf.atten_reg[id], d = wf.lobes[id];
    const float4 e = wf.diffuse[id], f = wf.reflection[id];
    s.pos = V3(a.x, a.y, a.z); s.bsdf_pdf = a.w;
    s.view = V3(b.x, b.y, b.z); s.random_seed = __float_as_uint(b.w);
    s.attenuation = V3(c.x, c.y, c.z); s.regularization = c.w;
    s.primary_lobes = LOBES(d.x, d.y, d.z, d.w);
    s.diffuse = V3(e.x, e.y, e.z);
    s.reflection = V3(f.x, f.y, f.z);
    s.aux_alpha = e.w;
}
TRD void load_first_hit(const tr_wavefront& wf, uint32_t id, PathState& s)
{
    const float4 a = wf.fh_emission[id], b = wf.fh_albedo[id];
    s.fh_emission = V3(a.x, a.y, a.z);
    s.fh_albedo = V4(b.x, b.y, b.z, a.w);
    s.fh_metallic = b.w;
}
TRD void store_job(const tr_wavefront& wf, uint32_t id, const NeeJob& j)
{
    float4* o = wf.nee + 4 * (size_t)id;
    o[0] = make_float4(j.dir.x, j.dir.y, j.dir.z, j.tmax);
    o[1] = make_float4(j.contrib.x, j.contrib.y, j.contrib.z, j.mis_pdf);
    o[2] = make_float4(j.attenuation.x, j.attenuation.y, j.attenuation.z, j.w_diffuse);
    o[3] = make_float4(j.modulate.x, j.modulate.y, j.modulate.z, j.w_reflection);
}
TRD void load_job(const tr_wavefront& wf, uint32_t id, int bounce, NeeJob& j)
{
    const float4* o = wf.nee + 4 * (size_t)id;
    const float4 a = o[0], b = o[1], c = o[2], d = o[3];
    j.dir = V3(a.x, a.y, a.z); j.tmax = a.w;
    j.contrib = V3(b.x, b.y, b.z); j.mis_pdf = b.w;
    j.attenuation = V3(c.x, c.y, c.z); j.w_diffuse = c.w;
    j.modulate = V3(d.x, d.y, d.z); j.w_reflection = d.w;
    j.needs_shadow_ray = 1;
    j.bounce = bounce;
}

// Warp-aggregated append: one atomicAdd per warp. Must be reached by all 32 lanes.
TRD void warp_push(uint32_t* __restrict__ queue, uint32_t* __restrict__ count, bool pred, uint32_t value)
{
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, pred);
    if(mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if(lane == leader) base = atomicAdd(count, __popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if(pred) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// kind: 0 = closest-hit traversal counters, 1 = shadow traversal counters
TRD void flush_counters(Counters* c, unsigned long long closest, unsigned long long shadow, unsigned long long paths,
                        unsigned long long nodes, unsigned long long tris, int kind)
{
    // warp reduce then one atomic per warp and counter
    for(int off = 16; off > 0; off >>= 1)
    {
        closest += __shfl_xor_sync(0xFFFFFFFFu, closest, off);
        shadow += __shfl_xor_sync(0xFFFFFFFFu, shadow, off);
        paths += __shfl_xor_sync(0xFFFFFFFFu, paths, off);
        nodes += __shfl_xor_sync(0xFFFFFFFFu, nodes, off);
        tris += __shfl_xor_sync(0xFFFFFFFFu, tris, off);
    }
    if((threadIdx.x & 31) == 0)
    {
        if(closest) atomicAdd(&c->closest_rays, closest);
        if(shadow) atomicAdd(&c->shadow_rays, shadow);
        if(paths) atomicAdd(&c->paths, paths);
        if(nodes) atomicAdd(kind ? &c->shadow_nodes : &c->closest_nodes, nodes);
        if(tris) atomicAdd(kind ? &c->shadow_tris : &c->closest_tris, tris);
    }
}

// TR_DEBUG_DRAIN (development aid): every wavefront kernel notes when its first thread started and its last one left
TRD unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int PROBE_TIMELINE = 136 * 2 * 16;     // offset of the per-launch {start, end} pairs behind the per-bounce drain records
constexpr int PROBE_LAUNCHES = 128;
TRD void probe_begin(const RenderParams& P) { if(P.probe && (threadIdx.x & 31) == 0) atomicMin(P.probe + PROBE_TIMELINE + 2 * P.probe_slot, globaltimer_ns()); }
TRD void probe_end(const RenderParams& P) { if(P.probe && (threadIdx.x & 31) == 0) atomicMax(P.probe + PROBE_TIMELINE + 2 * P.probe_slot + 1, globaltimer_ns()); }

TRD void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
TRD void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

// ---- wavefront kernels ---------------------------------------------------------------------------
// `batch` passes of one sample each are traced together (slot = pass * base_slots + base slot): the samples
// of different passes are independent, only their accumulation order matters, and k_wf_finish keeps it.
__global__ void __launch_bounds__(WF_BLOCK) k_wf_generate(const RenderParams P, const tr_wavefront wf, uint32_t sample_in_pass,
                                                           uint32_t base_slots, uint32_t total_slots,
                                                           uint32_t* __restrict__ queue, uint32_t* __restrict__ count)
{
    unsigned long long paths = 0;
    probe_begin(P);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounded = (total_slots + 31u) & ~31u;
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < rounded; slot += stride)
    {
        bool valid = false;
        if(slot < total_slots)
        {
            uint32_t px, py, view, widx;
            const uint32_t pass = slot / base_slots;
            valid = map_slot(P.frame, slot - pass * base_slots, px, py, view, widx);
            if(valid)
            {
                PathState s;
                uvec4 ss, rs;
                path_begin(P, px, py, view, sample_in_pass + pass * (uint32_t)P.opt.samples_per_pass, s, ss, rs);
                // only the ray: the rest of a fresh path's state is constant and k_wf_shade sets it itself at bounce 0
                store_ray(wf, slot, s);
                // one 16-byte sampler record per slot: the PCG stream state (uniform-random sampler) or the
                // constant Sobol-Owen seed (sobol-owen sampler) / Sobol index (sobol-z samplers)
                if(P.opt.local_sampler != TR_SAMPLER_UNIFORM_RANDOM) wf.rs_seed[slot] = make_uint4(ss.x, ss.y, ss.z, ss.w);
                else wf.rs_seed[slot] = make_uint4(rs.x, rs.y, rs.z, rs.w);
                paths++;
            }
        }
        warp_push(queue, count, valid, slot);
    }
    flush_counters(P.counters, 0, 0, paths, 0, 0, 0);
    probe_end(P);
}

// ---- persistent-threads traversal kernel ---------------------------------------------------------
// One kernel for both ray kinds (ANY = shadow rays):
//   * the grid is SM count x resident blocks; every warp fetches rays from the queue with ONE
//     atomicAdd per refill (ballot of the idle lanes) and refills as soon as fewer than
//     FETCH_THRESHOLD lanes are still traversing, so warps stay dense while rays of very different
//     length retire (measured: the naive one-ray-per-thread loop ran at 5-11 of 32 lanes);
//   * 8-wide compressed BVH (tr_scene.cuh): one node visit = five 128-bit loads and eight quantised
//     slab tests; the traversal stack holds 64-bit node groups (child base + hit bits) in a
//     shared-memory short stack ([entry][thread]), entries beyond SHORT_STACK spill to local memory;
//   * node phase and triangle phase are separate loops; a lane with triangle hits waits for the warp's
//     node phase only while other lanes still look for their first triangle group.
#ifndef TR_SHORT_STACK
#define TR_SHORT_STACK 8
#endif
constexpr int SHORT_STACK = TR_SHORT_STACK;
constexpr int FETCH_THRESHOLD = 24;
constexpr int MAX_FACTORS = 6;

// Per-ray state that is touched only when a ray is fetched, hits a triangle or retires lives in shared memory
// ([word][thread], conflict-free) instead of registers: the kernel runs at 64 registers / 8 blocks per SM, and what
// the compiler spilled at that budget was reloaded from local memory on every node step.
enum { COLD_ID = 0, COLD_SX, COLD_SY, COLD_SZ, COLD_KZ, COLD_INST, COLD_PRIM, COLD_U, COLD_V, COLD_FLAGS, COLD_SPLIT, COLD_WORDS,
       COLD_NF = COLD_INST, COLD_OVERFLOW = COLD_PRIM };

// TOPSM: the first TOP_NODES nodes of the breadth-first node array (root, its children, their children) are staged in
// shared memory with one TMA bulk copy per block and read from there (TR_TUNE_TMA bit 1; measured).
constexpr int TOP_NODES = 73;
//
// SPLIT (tail splitting): once the queue is empty a launch ends with a few lanes finishing rays of several hundred node
// steps while the rest of the GPU idles -- a constant ~0.17 ms per launch that is 3-6 % of a full-frame launch and a third
// of a 1/8-frame one. Idle lanes of the warp then take over subtrees from the busy ones: a lane
// with stacked node groups gives its BOTTOM stack entry (the largest pending subtree) to an idle lane together with the
// ray (shuffles + the shared-memory column); the helper traverses it like any ray, helpers donate in turn, and a
// finished helper hands its best hit back to the ray's own lane, which retires the ray when no helper is left. Closest
// hit: the (t, instance, primitive) minimum is order independent, so the result is the sequential one bit for bit; every
// round the lanes of a ray share the best hit so far (the culling bound). Shadow rays (opaque scenes): any helper's
// blocker blocks the ray and ends the others. COLD_SPLIT = the ray's own lane | helpers outstanding << 8.
//
// FUSED: ONE launch traces the closest-hit rays of bounce `bounce` and the shadow rays of bounce `bounce - 1` -- both
// depend only on the shade pass before them. Two persistent grids cannot share the machine (the first one fills it, the
// second one only gets the SMs the first one drains from), so on two streams the pair still paid most of both drain
// phases; here the work list is [closest queue | shadow queue], every lane fetches from it, and the lanes that the long
// closest-hit rays leave idle are filled with shadow rays: one drain phase per bounce. The ray kind is per lane (bit 2 of
// the COLD_KZ word); node steps are the same code for both kinds. Opaque scenes only (no any-hit factor lists); ANY is
// ignored.
template<bool ANY, bool STATS, bool TOPSM = false, bool NEAR = false, bool SPLIT = false, bool FUSED = false>
__global__ void __launch_bounds__(TRACE_BLOCK, ANY ? TRACE_MIN_BLOCKS_SHADOW : TRACE_MIN_BLOCKS_CLOSEST) k_wf_trace(const RenderParams P, const tr_wavefront wf, uint32_t bounce,
                                                        const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
                                                        uint32_t* __restrict__ work_counter,
                                                        const uint32_t* __restrict__ queue2 = nullptr, const uint32_t* __restrict__ count2 = nullptr)
{
    __shared__ uint2 s_stack[SHORT_STACK * TRACE_BLOCK];
    __shared__ uint32_t s_cold[COLD_WORDS * TRACE_BLOCK];
    __shared__ __align__(128) float4 s_top[TOPSM ? TOP_NODES * TR_BVH_NODE_VEC4 : 1];
    __shared__ uint64_t s_top_mbar;
    uint2 lstack[TR_BVH_STACK - SHORT_STACK];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t n_first = *count;                              // FUSED: closest-hit rays first (the long ones), then the shadow rays
    const uint32_t n = FUSED ? n_first + *count2 : n_first;
    const float4* __restrict__ nodes = P.scene.bvh.nodes;
    // TR_DEBUG_DRAIN (development aid): when does the first warp start, when do the first and the last warp find the
    // queue empty, when does the last warp leave
    unsigned long long* const probe = SPLIT && P.probe ? P.probe + 136 * (2 * bounce + (ANY ? 1 : 0)) : nullptr;
    bool probe_seen = false;
    if(SPLIT && probe && lane == 0) atomicMin(probe + 0, globaltimer_ns());
    if(SPLIT) probe_begin(P);
    uint32_t top_count = 0;
    if(TOPSM)
    {
        top_count = P.scene.bvh.node_count < (uint32_t)TOP_NODES ? P.scene.bvh.node_count : (uint32_t)TOP_NODES;
        if(tid == 0) mbar_init(&s_top_mbar, 1);
        __syncthreads();
        if(tid == 0) tma_load_bulk(s_top, nodes, top_count * TR_BVH_NODE_VEC4 * (uint32_t)sizeof(float4), &s_top_mbar);
        mbar_wait(&s_top_mbar, 0);
    }
    const float4* __restrict__ tris = P.scene.bvh.tris;
    const bool include_lights = (FUSED || !ANY) && !(P.opt.hide_lights && bounce == 0);
    const float ray_tmin = ANY ? P.opt.min_ray_dist : (bounce == 0 ? 0.0f : P.opt.min_ray_dist);
    const bool transparency = !FUSED && P.scene.any_non_opaque != 0;
    // the kind of the lane's ray: a template constant, or (FUSED) bit 2 of the COLD_KZ word
#define KIND_ANY() (FUSED ? (COLD(COLD_KZ) >> 2) != 0u : ANY)

    TraceStats st; st.nodes = 0; st.tris = 0;
    // rays fetched by this thread: closest-hit rays in bits 0..11, shadow rays (FUSED) in bits 12..23; SPLIT: bits 24..31 =
    // node steps of the current ray (wraps at 256)
    uint32_t rays = 0;
    int warp_helpers = 0;                 // SPLIT: helper lanes at work in this warp (warp-uniform)

    // per-lane ray + traversal state held in registers
    uint2 G = make_uint2(0u, 0u), Gt = make_uint2(0u, 0u);
    uint32_t near_bit = 0u;               // NEAR: hit bit of the nearest inner child of the group just produced
    int sp = 0;
    bool has_ray = false, exhausted = false, done = true;
    // closest: the culling bound = t of the best hit so far, +inf while there is none. shadow: the ray's tmax.
    float tmin = 0.0f, best_t = 0.0f;
    bool blocked = false;                 // shadow: an opaque blocker was found
    RaySlab rs;
    rs = make_ray_slab(V3(0), V3(0));
    // shadow rays through non-opaque geometry: (t, instance, primitive, 1 - alpha)
    float ft[MAX_FACTORS], ff[MAX_FACTORS];
    int fi[MAX_FACTORS], fp[MAX_FACTORS];

#define COLD(w) s_cold[(w) * TRACE_BLOCK + tid]
#define COLDF(w) reinterpret_cast<float*>(s_cold)[(w) * TRACE_BLOCK + tid]
#define COLD_AT(w, l) s_cold[(w) * TRACE_BLOCK + (tid & ~31) + (l)]
    // The thread's column of the shared-memory stack as a 32-bit shared address, accessed with explicit ld/st.shared:
    // the generic-pointer form cost an address rebuild from two special-register reads (S2R) plus spills on every push
    // and pop. In the shadow kernel the value is made opaque so that it stays in a register (measured: shadow 5.67 ->
    // 5.44 ms); the closest-hit kernel is better off letting the compiler rematerialise it (11.50 -> 11.26 ms).
    uint32_t stack_base = (uint32_t)__cvta_generic_to_shared(&s_stack[tid]);
    if(ANY && !FUSED) asm volatile("mov.u32 %0, %0;" : "+r"(stack_base));
#define PUSH(v) do { if(sp < SHORT_STACK) { const uint2 v_ = (v); asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(stack_base + (uint32_t)sp * (uint32_t)(TRACE_BLOCK * sizeof(uint2))), "r"(v_.x), "r"(v_.y) : "memory"); } \
                     else lstack[sp - SHORT_STACK] = (v); sp++; } while(0)
#define POP(dst) do { sp--; if(sp < SHORT_STACK) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(dst.x), "=r"(dst.y) : "r"(stack_base + (uint32_t)sp * (uint32_t)(TRACE_BLOCK * sizeof(uint2))) : "memory"); \
                      else dst = lstack[sp - SHORT_STACK]; } while(0)

    for(;;)
    {
        // ---- retire finished rays --------------------------------------------------------------
        // (SPLIT: only on the ray's own lane, and only when no helper lane is still out)
        if(has_ray && done && (!SPLIT || COLD(COLD_SPLIT) == (uint32_t)lane))
        {
            bool retire = true;
            const uint32_t id = COLD(COLD_ID);
            if(!KIND_ANY())
            {
                const bool found = best_t < inf_();
                // sphere lights + the in-t-order any-hit emulation (see closest_hit() in tr_path.cuh)
                int light_id = -1;
                float light_t = 0.0f;
                if(include_lights && P.scene.point_light_count)
                {
                    const float4 d = wf.dir_seed[id];
                    const vec3 org = V3(rs.ox, rs.oy, rs.oz), dir = V3(d.x, d.y, d.z);
                    for(uint32_t i = 0; i < P.scene.point_light_count; ++i)
                    {
                        const tr_point_light& pl = P.scene.point_lights[i];
                        if(pl.radius == 0.0f) continue;
                        float t = sphere_intersection(ld3(pl.pos), pl.radius, org, dir);
                        if(!(t > 0)) continue;
                        if(!(t >= tmin)) continue;
                        if(light_id < 0 || t < light_t) { light_id = (int)i; light_t = t; }
                    }
                }
                if(light_id >= 0 && (!found || light_t < best_t))
                    wf.hits[id] = make_int4(-1, light_id, __float_as_int(light_t), 0);
                else if(!found)
                    wf.hits[id] = make_int4(-1, -1, 0, 0);
                else
                {
                    const int b_inst = (int)COLD(COLD_INST), b_prim = (int)COLD(COLD_PRIM);
                    const float b_u = COLDF(COLD_U), b_v = COLDF(COLD_V);
                    bool accept = true;
                    if(transparency && !(COLD(COLD_FLAGS) & TR_TRI_FLAG_OPAQUE))
                    {
                        // the any-hit shader's random number comes from the payload seed, which it advances
                        uint32_t seed = __float_as_uint(wf.dir_seed[id].w);
                        float alpha_cutoff = (float)pcg(seed) * TR_INV_UINT32_MAX;
                        wf.dir_seed[id].w = __uint_as_float(seed);
                        if(hit_alpha(P.scene, b_inst, b_prim, b_u, b_v) <= alpha_cutoff) accept = false;
                    }
                    if(accept) wf.hits[id] = make_int4(b_inst, b_prim, __float_as_int(b_u), __float_as_int(b_v));
                    else
                    {
                        // ignored by the any-hit test: continue behind it
                        tmin = best_t;
                        best_t = inf_();
                        sp = 0;
                        G = make_uint2(0u, (0x01000000u << rs.octinv) | 1u);   // the root as a one-child group
                        Gt = make_uint2(0u, 0u);
                        near_bit = 0u;
                        done = false;
                        retire = false;
                    }
                }
            }
            else
            {
                float visibility = blocked ? 0.0f : 1.0f;
                const int nf = transparency ? (int)COLD(COLD_NF) : 0;
                if(!blocked && nf > 0)
                {
                    for(int k = 0; k < nf; ++k)
                    {
                        int m = -1;
                        for(int j = 0; j < nf; ++j)
                        {
                            if(fi[j] == -2) continue;
                            if(m < 0 || ft[j] < ft[m] || (ft[j] == ft[m] && (fi[j] < fi[m] || (fi[j] == fi[m] && fp[j] < fp[m])))) m = j;
                        }
                        visibility *= ff[m];
                        fi[m] = -2;
                        if(visibility == 0.0f) break;
                    }
                    if(visibility != 0.0f) visibility *= COLDF(COLD_OVERFLOW);
                }
                NeeJob job;
                load_job(wf, id, (int)(FUSED ? bounce - 1u : bounce), job);
                PathState s;
                const float4 e = wf.diffuse[id], f = wf.reflection[id];
                s.diffuse = V3(e.x, e.y, e.z);
                s.reflection = V3(f.x, f.y, f.z);
                nee_finish(P, job, visibility, s);
                wf.diffuse[id] = make_float4(s.diffuse.x, s.diffuse.y, s.diffuse.z, e.w);
                wf.reflection[id] = make_float4(s.reflection.x, s.reflection.y, s.reflection.z, 0.0f);
            }
            if(retire) has_ray = false;
        }

        // ---- split: idle lanes take over subtrees of rays still in flight (see SPLIT above) ---------
        // Runs in the tail (the queue is empty: `exhausted`), and before that for rays that have already taken
        // `tune_split_steps` node steps -- the few rays of several hundred steps that would otherwise end every launch.
        if(SPLIT && probe && !probe_seen && __ballot_sync(0xFFFFFFFFu, exhausted))
        {
            probe_seen = true;
            if(lane == 0)
            {
                const unsigned long long t = globaltimer_ns();
                atomicMin(probe + 1, t); atomicMax(probe + 2, t);
                const unsigned long long bin = (t - *(volatile unsigned long long*)probe) / P.tune_probe_bin_ns;
                atomicAdd(probe + 72 + (bin < 63ull ? bin : 63ull), 1ull);
            }
        }
        if(SPLIT && P.tune_split_stage && (FUSED || !ANY || !transparency))
        {
            const uint32_t FULL = 0xFFFFFFFFu;
            const bool can_give = has_ray && !done && sp > 0;
            const uint32_t tail = __ballot_sync(FULL, exhausted);
            const uint32_t longs = __ballot_sync(FULL, can_give && (rays >> 24) >= (uint32_t)P.tune_split_steps);
            if(tail | longs | warp_helpers)
            {
                // The stage is rare and the kernel has no register to spare: the hot per-lane state is parked in local
                // memory by hand for its duration (the compiler otherwise spills inside the traversal loop instead).
                volatile uint32_t park[14];
                park[0] = G.x; park[1] = G.y; park[2] = Gt.x; park[3] = Gt.y;
                park[4] = __float_as_uint(rs.idx); park[5] = __float_as_uint(rs.idy); park[6] = __float_as_uint(rs.idz);
                park[7] = __float_as_uint(rs.ox); park[8] = __float_as_uint(rs.oy); park[9] = __float_as_uint(rs.oz);
                park[10] = rs.octinv; park[11] = rs.sel0; park[12] = __float_as_uint(tmin); park[13] = __float_as_uint(best_t);
                const uint32_t below = (1u << lane) - 1u;
                __syncwarp();
                // (1) a finished helper hands its result to the ray's own lane and becomes idle
                uint32_t fin = warp_helpers ? __ballot_sync(FULL, has_ray && done && (COLD(COLD_SPLIT) & 0xFFu) != (uint32_t)lane) : 0u;
                while(fin)
                {
                    const int h = __ffs(fin) - 1;
                    fin &= fin - 1u;
                    warp_helpers--;
                    const int o = (int)(COLD_AT(COLD_SPLIT, h) & 0xFFu);
                    if(FUSED ? (COLD_AT(COLD_KZ, h) >> 2) != 0u : ANY)
                    {
                        const bool hb = __shfl_sync(FULL, blocked ? 1 : 0, h) != 0;
                        if(lane == o && hb) { blocked = true; done = true; }
                    }
                    else
                    {
                        const float ht = __uint_as_float(__shfl_sync(FULL, park[13], h));
                        if(lane == o && ht < inf_())
                        {
                            const int h_inst = (int)COLD_AT(COLD_INST, h), h_prim = (int)COLD_AT(COLD_PRIM, h);
                            const float own_t = __uint_as_float(park[13]);
                            bool better = ht < own_t;
                            if(ht == own_t)
                            {
                                const int b_inst = (int)COLD(COLD_INST), b_prim = (int)COLD(COLD_PRIM);
                                better = h_inst < b_inst || (h_inst == b_inst && h_prim < b_prim);
                            }
                            if(better)
                            {
                                COLD(COLD_INST) = (uint32_t)h_inst; COLD(COLD_PRIM) = (uint32_t)h_prim;
                                COLD(COLD_U) = COLD_AT(COLD_U, h); COLD(COLD_V) = COLD_AT(COLD_V, h); COLD(COLD_FLAGS) = COLD_AT(COLD_FLAGS, h);
                                park[13] = __float_as_uint(ht);
                            }
                        }
                    }
                    if(lane == o) COLD(COLD_SPLIT) -= 0x100u;
                    if(lane == h) has_ray = false;
                    __syncwarp();
                }
                // (2) the lanes working on one ray share what they found so far
                if(warp_helpers)
                {
                    const int own = has_ray ? (int)(COLD(COLD_SPLIT) & 0xFFu) : -1;
                    if(FUSED || ANY)
                    {
                        const bool ob = __shfl_sync(FULL, blocked ? 1 : 0, own < 0 ? lane : own) != 0;
                        if(has_ray && own != lane && ob) done = true;          // the (shadow) ray is blocked: its helpers stop
                    }
                    if(FUSED || !ANY)
                    {
                        uint32_t todo = __ballot_sync(FULL, has_ray && COLD(COLD_SPLIT) >= 0x100u && !KIND_ANY());    // closest-hit rays with helpers out
                        while(todo)
                        {
                            const int o = __ffs(todo) - 1;
                            todo &= todo - 1u;
                            const bool member = own == o;
                            const uint32_t key = member ? park[13] : 0xFFFFFFFFu;     // t > 0 or +inf: ordered like its bits
                            const uint32_t m = __reduce_min_sync(FULL, key);
                            if(m >= 0x7F800000u) continue;
                            const int src = __ffs(__ballot_sync(FULL, member && key == m)) - 1;
                            if(member && key > m)
                            {
                                COLD(COLD_INST) = COLD_AT(COLD_INST, src); COLD(COLD_PRIM) = COLD_AT(COLD_PRIM, src);
                                COLD(COLD_U) = COLD_AT(COLD_U, src); COLD(COLD_V) = COLD_AT(COLD_V, src); COLD(COLD_FLAGS) = COLD_AT(COLD_FLAGS, src);
                                park[13] = m;
                            }
                            __syncwarp();
                        }
                    }
                }
                // (3) the k-th idle lane takes the bottom stack entry (the largest pending subtree) of the k-th giving lane.
                // The ray travels through the taker's own, idle, column of the shared-memory stack.
                const uint32_t idle = __ballot_sync(FULL, !has_ray);
                const uint32_t givers = tail ? __ballot_sync(FULL, has_ray && !done && sp > 0) : longs;
                const int pairs = min(__popc(idle), __popc(givers));
                if(pairs > 0)
                {
                    const bool gives = ((givers >> lane) & 1u) && __popc(givers & below) < pairs;
                    const bool takes = !has_ray && __popc(idle & below) < pairs;
                    if(gives)
                    {
                        uint2* const box = s_stack + (tid & ~31) + __fns(idle, 0u, __popc(givers & below) + 1);
                        box[0 * TRACE_BLOCK] = s_stack[tid];
                        for(int k = 1; k <= 5; ++k) box[k * TRACE_BLOCK] = make_uint2(park[2 + 2 * k], park[3 + 2 * k]);
                        box[6 * TRACE_BLOCK] = make_uint2((uint32_t)lane, 0u);
                        for(int k = 1; k < sp; ++k)
                        {
                            const uint2 v = k < SHORT_STACK ? s_stack[k * TRACE_BLOCK + tid] : lstack[k - SHORT_STACK];
                            if(k - 1 < SHORT_STACK) s_stack[(k - 1) * TRACE_BLOCK + tid] = v;
                            else lstack[k - 1 - SHORT_STACK] = v;
                        }
                        sp--;
                        rays = (rays & 0x00FFFFFFu) | (((uint32_t)P.tune_split_steps >> 1) << 24);    // gives again after half the threshold
                    }
                    __syncwarp();
                    if(takes)
                    {
                        const uint2* const box = s_stack + tid;
                        uint2 v = box[0 * TRACE_BLOCK];
                        park[0] = v.x; park[1] = v.y; park[2] = 0u; park[3] = 0u;
                        for(int k = 1; k <= 5; ++k) { v = box[k * TRACE_BLOCK]; park[2 + 2 * k] = v.x; park[3 + 2 * k] = v.y; }
                        const int src = (int)box[6 * TRACE_BLOCK].x;
                        for(int w = COLD_ID; w <= COLD_FLAGS; ++w) COLD(w) = COLD_AT(w, src);
                        COLD(COLD_SPLIT) = COLD_AT(COLD_SPLIT, src) & 0xFFu;       // helper of the giver's ray
                        near_bit = 0u;
                        sp = 0;
                        blocked = false;
                        done = false;
                        has_ray = true;
                        rays &= 0x00FFFFFFu;
                    }
                    __syncwarp();
                    uint32_t took = __ballot_sync(FULL, takes);
                    warp_helpers += __popc(took);
                    while(took)
                    {
                        const int h = __ffs(took) - 1;
                        took &= took - 1u;
                        if(lane == (int)(COLD_AT(COLD_SPLIT, h) & 0xFFu)) COLD(COLD_SPLIT) += 0x100u;
                    }
                    __syncwarp();
                }
                G = make_uint2(park[0], park[1]); Gt = make_uint2(park[2], park[3]);
                rs.idx = __uint_as_float(park[4]); rs.idy = __uint_as_float(park[5]); rs.idz = __uint_as_float(park[6]);
                rs.ox = __uint_as_float(park[7]); rs.oy = __uint_as_float(park[8]); rs.oz = __uint_as_float(park[9]);
                rs.octinv = park[10]; rs.sel0 = park[11]; tmin = __uint_as_float(park[12]); best_t = __uint_as_float(park[13]);
            }
        }

        // ---- refill idle lanes: one atomicAdd per warp ----------------------------------------
        {
            const bool want = !has_ray && !exhausted;
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, want);
            if(mask)
            {
                const int leader = __ffs(mask) - 1;
                uint32_t base = 0;
                if(lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(mask));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if(want)
                {
                    const uint32_t i = base + __popc(mask & ((1u << lane) - 1u));
                    if(i < n)
                    {
                        const bool any = FUSED ? i >= n_first : ANY;
                        const uint32_t id = (FUSED && any) ? queue2[i - n_first] : queue[i];
                        const float4 o = wf.pos_pdf[id];
                        const vec3 org = V3(o.x, o.y, o.z);
                        vec3 dir;
                        if(any)
                        {
                            const float4 j0 = wf.nee[4 * (size_t)id];
                            dir = V3(j0.x, j0.y, j0.z);
                            best_t = j0.w;
                        }
                        else
                        {
                            const float4 d = wf.dir_seed[id];
                            dir = V3(d.x, d.y, d.z);
                            best_t = inf_();
                        }
                        tmin = ray_tmin;
                        const RayPre rp = precompute_ray(org, dir);
                        COLD(COLD_ID) = id;
                        COLDF(COLD_SX) = rp.Sx; COLDF(COLD_SY) = rp.Sy; COLDF(COLD_SZ) = rp.Sz; COLD(COLD_KZ) = (uint32_t)rp.kz | (FUSED && any ? 4u : 0u);
                        if(ANY && transparency) { COLD(COLD_NF) = 0u; COLDF(COLD_OVERFLOW) = 1.0f; }
                        if(SPLIT) COLD(COLD_SPLIT) = (uint32_t)lane;
                        rs = make_ray_slab(org, dir);
                        sp = 0;
                        blocked = false;
                        has_ray = true;
                        done = false;
                        rays = (rays & 0x00FFFFFFu) + ((FUSED && any) ? 0x1000u : 1u);
                        G = make_uint2(0u, (0x01000000u << rs.octinv) | 1u);   // the root as a one-child group
                        Gt = make_uint2(0u, 0u);
                        near_bit = 0u;
                    }
                    else exhausted = true;
                }
            }
            if(__ballot_sync(0xFFFFFFFFu, has_ray) == 0) break;
        }

        // ---- traverse ---------------------------------------------------------------------------
        // One node step per iteration for the lanes that have an open node group; lanes that popped a
        // (postponed) triangle group take it over. Triangle tests run right after, unless fewer than
        // 1/4 of the lanes of this iteration have any -- then the group is pushed and traversal goes on.
        while(!done)
        {
            if(G.y & 0xFF000000u)
            {
                uint32_t bit = 31u - (uint32_t)__clz((int)G.y);
                if(NEAR && near_bit) bit = near_bit;
                G.y &= ~(1u << bit);
                const uint32_t child = child_of_bit(G, bit, rs.octinv);
                if(G.y & 0xFF000000u) PUSH(G);
                if(STATS) st.nodes++;
                if(SPLIT) rays += 1u << 24;
                if(TOPSM) intersect_node8_at(child < top_count ? s_top + TR_BVH_NODE_VEC4 * child : nodes + TR_BVH_NODE_VEC4 * (size_t)child, rs, tmin, best_t, G, Gt);
                else if(NEAR) intersect_node8_nearest(nodes, child, rs, tmin, best_t, G, Gt, near_bit);
                else intersect_node8(nodes, child, rs, tmin, best_t, G, Gt);
            }
            else
            {
                Gt = G;
                G = make_uint2(0u, 0u);
            }

            const int active = __popc(__activemask());
            while(Gt.y & 0xFFu)
            {
                if(__popc(__activemask()) * P.tune_postpone_div < active && sp < TR_BVH_POSTPONE_CAP)
                {
                    PUSH(Gt);
                    Gt.y = 0u;
                    break;
                }
                const uint32_t tri_index = pop_triangle(Gt, rs.octinv);
                if(STATS) st.tris++;
                const float4* tp = tris + 3 * (size_t)tri_index;
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                RayPre rp;
                const uint32_t kz_word = COLD(COLD_KZ);
                const bool any = FUSED ? (kz_word >> 2) != 0u : ANY;
                rp.kz = (int)(FUSED ? kz_word & 3u : kz_word); rp.Sx = COLDF(COLD_SX); rp.Sy = COLDF(COLD_SY); rp.Sz = COLDF(COLD_SZ);
                rp.org = V3(rs.ox, rs.oy, rs.oz);
                float t, u, v;
                // closest: any t up to +inf qualifies here, the comparison with the best hit follows; shadow: t < tmax
                if(!intersect_triangle(rp, f3(t0), f3(t1), f3(t2), tmin, any ? best_t : inf_(), t, u, v)) continue;
                const int inst = __float_as_int(t0.w), prim = __float_as_int(t1.w);
                const uint32_t flags = __float_as_uint(t2.w);
                if(any)
                {
                    if(FUSED || (flags & TR_TRI_FLAG_OPAQUE))
                    {
                        blocked = true;        // terminate on the first opaque hit
                        done = true;
                        Gt.y = 0u;
                        break;
                    }
                    const float f = 1.0f - hit_alpha(P.scene, inst, prim, u, v);
                    const int nf = (int)COLD(COLD_NF);
                    if(nf < MAX_FACTORS) { ft[nf] = t; ff[nf] = f; fi[nf] = inst; fp[nf] = prim; COLD(COLD_NF) = (uint32_t)(nf + 1); }
                    else COLDF(COLD_OVERFLOW) = COLDF(COLD_OVERFLOW) * f;
                    continue;
                }
                // best_t is +inf while nothing was hit, so `t < best_t` covers the first hit; ties go to the lower ids
                bool better = t < best_t;
                if(t == best_t)
                {
                    const int b_inst = (int)COLD(COLD_INST), b_prim = (int)COLD(COLD_PRIM);
                    better = inst < b_inst || (inst == b_inst && prim < b_prim);
                }
                if(better)
                {
                    COLD(COLD_INST) = (uint32_t)inst; COLD(COLD_PRIM) = (uint32_t)prim;
                    COLDF(COLD_U) = u; COLDF(COLD_V) = v; COLD(COLD_FLAGS) = flags;
                    best_t = t;
                }
            }
            if(done) break;

            if((G.y & 0xFF000000u) == 0u)
            {
                if(sp == 0) { done = true; break; }
                POP(G);
                if(NEAR) near_bit = 0u;
            }

            // too few lanes left busy: go refill
            if(__popc(__activemask()) < P.tune_fetch_threshold) break;
        }
    }
#undef PUSH
#undef POP
#undef COLD
#undef COLDF
#undef COLD_AT
#undef KIND_ANY
    rays &= 0x00FFFFFFu;
    if(FUSED) flush_counters(P.counters, rays & 0xFFFu, rays >> 12, 0, st.nodes, st.tris, 0);
    else flush_counters(P.counters, ANY ? 0 : rays, ANY ? rays : 0, 0, st.nodes, st.tris, ANY ? 1 : 0);
    if(SPLIT && probe && lane == 0)
    {
        const unsigned long long t = globaltimer_ns();
        atomicMax(probe + 3, t);
        const unsigned long long bin = (t - *(volatile unsigned long long*)probe) / P.tune_probe_bin_ns;
        atomicAdd(probe + 8 + (bin < 63ull ? bin : 63ull), 1ull);
    }
    if(SPLIT) probe_end(P);
}

// ---- persistent traversal with a warp-cooperative triangle phase --------------------------------------
// Same rays, same node steps, same predicate and tie-break as k_wf_trace -- the result is identical -- but triangle
// tests no longer run at the few lanes that happen to have a leaf hit in the current node step (measured: 7-8 of 32
// lanes, 36 % of the kernel's warp instructions). A lane only POSTS the triangle group of a node step (8 bytes, up to
// POST_CAP per lane in shared memory). When the warp has posted `tune_coop_groups` groups, or `tune_coop_wait` lanes
// have run out of nodes and wait for their triangles, or no lane has node work left, the whole warp runs a triangle
// phase: the posted groups are expanded into (ray lane, triangle) pairs in a per-warp queue and tested 32 at a time by
// ALL lanes -- lane L tests pair L with the source lane's ray (origin / tmin by shuffle, shear constants from shared
// memory). Closest hit: the pairs of one ray are reduced with shared-memory atomicMin on the bits of t (positive floats
// order like integers), then on the instance id and the primitive id among the ties, so the (t, instance, primitive)
// minimum wins exactly like in the sequential loop. Shadow rays (opaque scenes only): any hit sets the ray's blocked flag.
constexpr int POST_CAP = 4;
enum { COOP_COLD_T = COLD_WORDS, COOP_COLD_WORDS };

template<bool ANY, bool STATS>
__global__ void __launch_bounds__(WF_BLOCK, ANY ? TRACE_MIN_BLOCKS_SHADOW : TRACE_MIN_BLOCKS_CLOSEST) k_wf_trace_coop(const RenderParams P, const tr_wavefront wf, uint32_t bounce,
                                                        const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
                                                        uint32_t* __restrict__ work_counter)
{
    __shared__ uint2 s_stack[SHORT_STACK * WF_BLOCK];
    __shared__ uint32_t s_cold[COOP_COLD_WORDS * WF_BLOCK];
    __shared__ uint2 s_post[POST_CAP * WF_BLOCK];
    __shared__ uint32_t s_pairs[(WF_BLOCK / 32) * 64];
    uint2 lstack[TR_BVH_STACK - SHORT_STACK];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wbase = tid & ~31;
    uint32_t* const pairs = s_pairs + (tid >> 5) * 64;
    const uint32_t lanes_below = (1u << lane) - 1u;
    const uint32_t n = *count;
    const float4* __restrict__ nodes = P.scene.bvh.nodes;
    const float4* __restrict__ tris = P.scene.bvh.tris;
    const bool include_lights = !ANY && !(P.opt.hide_lights && bounce == 0);
    const float ray_tmin = ANY ? P.opt.min_ray_dist : (bounce == 0 ? 0.0f : P.opt.min_ray_dist);
    const bool transparency = P.scene.any_non_opaque != 0;

    TraceStats st; st.nodes = 0; st.tris = 0;
    uint32_t rays = 0;
    uint2 G = make_uint2(0u, 0u);
    int sp = 0, npost = 0;
    bool has_ray = false, exhausted = false, done = true;
    float tmin = 0.0f, best_t = 0.0f;
    bool blocked = false;
    RaySlab rs;
    rs = make_ray_slab(V3(0), V3(0));
    int wposted = 0;                       // warp-uniform: groups posted since the last triangle phase

#define COLD(w) s_cold[(w) * WF_BLOCK + tid]
#define COLDF(w) reinterpret_cast<float*>(s_cold)[(w) * WF_BLOCK + tid]
#define COLD_OF(w, l) s_cold[(w) * WF_BLOCK + wbase + (l)]
    uint32_t stack_base = (uint32_t)__cvta_generic_to_shared(&s_stack[tid]);
    if(ANY) asm volatile("mov.u32 %0, %0;" : "+r"(stack_base));
#define PUSH(v) do { if(sp < SHORT_STACK) { const uint2 v_ = (v); asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(stack_base + (uint32_t)sp * (uint32_t)(WF_BLOCK * sizeof(uint2))), "r"(v_.x), "r"(v_.y) : "memory"); } \
                     else lstack[sp - SHORT_STACK] = (v); sp++; } while(0)
#define POP(dst) do { sp--; if(sp < SHORT_STACK) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(dst.x), "=r"(dst.y) : "r"(stack_base + (uint32_t)sp * (uint32_t)(WF_BLOCK * sizeof(uint2))) : "memory"); \
                      else dst = lstack[sp - SHORT_STACK]; } while(0)

    // One batch of `m` (<= 32) queued pairs, executed by all 32 lanes.
    auto test_batch = [&](int m)
    {
        __syncwarp();
        const bool valid = lane < m;
        const uint32_t e = valid ? pairs[lane] : 0u;
        const int src = (int)(e >> 27);
        const uint32_t tri_index = e & 0x07FFFFFFu;
        RayPre rp;
        rp.org = V3(__shfl_sync(0xFFFFFFFFu, rs.ox, src), __shfl_sync(0xFFFFFFFFu, rs.oy, src), __shfl_sync(0xFFFFFFFFu, rs.oz, src));
        const float tmin_s = __shfl_sync(0xFFFFFFFFu, tmin, src);
        const float tmax_s = ANY ? __shfl_sync(0xFFFFFFFFu, best_t, src) : inf_();
        rp.kz = (int)COLD_OF(COLD_KZ, src);
        rp.Sx = __uint_as_float(COLD_OF(COLD_SX, src)); rp.Sy = __uint_as_float(COLD_OF(COLD_SY, src)); rp.Sz = __uint_as_float(COLD_OF(COLD_SZ, src));
        bool hit = false;
        float t = 0.0f, u = 0.0f, v = 0.0f;
        int inst = 0, prim = 0;
        uint32_t flags = 0u;
        if(valid && !(ANY && COLD_OF(COLD_FLAGS, src)))
        {
            if(STATS) st.tris++;
            const float4* tp = tris + 3 * (size_t)tri_index;
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            hit = intersect_triangle(rp, f3(t0), f3(t1), f3(t2), tmin_s, tmax_s, t, u, v);
            inst = __float_as_int(t0.w); prim = __float_as_int(t1.w); flags = __float_as_uint(t2.w);
        }
        if(ANY)
        {
            if(hit) COLD_OF(COLD_FLAGS, src) = 1u;              // opaque scenes only: every hit blocks
            __syncwarp();
            return;
        }
        // (t, instance, primitive) minimum per ray, the old best included
        const uint32_t old_t = __float_as_uint(best_t);
        const uint32_t key = __float_as_uint(t);
        if(hit) atomicMin(&COLD_OF(COOP_COLD_T, src), key);
        __syncwarp();
        if(COLD(COOP_COLD_T) != old_t) { COLD(COLD_INST) = 0x7FFFFFFFu; COLD(COLD_PRIM) = 0x7FFFFFFFu; }
        __syncwarp();
        const bool cand = hit && key == COLD_OF(COOP_COLD_T, src);
        const uint32_t inst_before = COLD(COLD_INST);
        if(cand) atomicMin(&COLD_OF(COLD_INST, src), (uint32_t)inst);
        __syncwarp();
        if(COLD(COLD_INST) != inst_before) COLD(COLD_PRIM) = 0x7FFFFFFFu;
        __syncwarp();
        const bool cand2 = cand && (uint32_t)inst == COLD_OF(COLD_INST, src);
        if(cand2) atomicMin(&COLD_OF(COLD_PRIM, src), (uint32_t)prim);
        __syncwarp();
        if(cand2 && (uint32_t)prim == COLD_OF(COLD_PRIM, src))
        {
            COLD_OF(COLD_U, src) = __float_as_uint(u); COLD_OF(COLD_V, src) = __float_as_uint(v); COLD_OF(COLD_FLAGS, src) = flags;
        }
        __syncwarp();
        best_t = __uint_as_float(COLD(COOP_COLD_T));
    };

    // The warp-wide triangle phase: expand every posted group into pairs, test them 32 at a time.
    auto triangle_phase = [&]()
    {
        if(!ANY) COLD(COOP_COLD_T) = __float_as_uint(best_t);
        int qcount = 0;
        uint2 Gp = make_uint2(0u, 0u);
        for(;;)
        {
            if(!(Gp.y & 0xFFu) && npost > 0) { npost--; Gp = s_post[npost * WF_BLOCK + tid]; }
            const bool has = (Gp.y & 0xFFu) != 0u;
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, has);
            if(mask == 0u) break;
            if(has)
            {
                const uint32_t tri_index = pop_triangle(Gp, rs.octinv);
                pairs[qcount + __popc(mask & lanes_below)] = ((uint32_t)lane << 27) | tri_index;
            }
            qcount += __popc(mask);
            if(qcount >= 32)
            {
                test_batch(32);
                const uint32_t keep = (lane < qcount - 32) ? pairs[32 + lane] : 0u;
                __syncwarp();
                if(lane < qcount - 32) pairs[lane] = keep;
                qcount -= 32;
            }
        }
        if(qcount) test_batch(qcount);
        if(ANY && has_ray && COLD(COLD_FLAGS)) { blocked = true; done = true; sp = 0; G = make_uint2(0u, 0u); }
    };

    for(;;)
    {
        // ---- retire finished rays (no posted triangles left) -----------------------------------------
        if(has_ray && done && npost == 0)
        {
            bool retire = true;
            const uint32_t id = COLD(COLD_ID);
            if(!ANY)
            {
                const bool found = best_t < inf_();
                int light_id = -1;
                float light_t = 0.0f;
                if(include_lights && P.scene.point_light_count)
                {
                    const float4 d = wf.dir_seed[id];
                    const vec3 org = V3(rs.ox, rs.oy, rs.oz), dir = V3(d.x, d.y, d.z);
                    for(uint32_t i = 0; i < P.scene.point_light_count; ++i)
                    {
                        const tr_point_light& pl = P.scene.point_lights[i];
                        if(pl.radius == 0.0f) continue;
                        float t = sphere_intersection(ld3(pl.pos), pl.radius, org, dir);
                        if(!(t > 0)) continue;
                        if(!(t >= tmin)) continue;
                        if(light_id < 0 || t < light_t) { light_id = (int)i; light_t = t; }
                    }
                }
                if(light_id >= 0 && (!found || light_t < best_t))
                    wf.hits[id] = make_int4(-1, light_id, __float_as_int(light_t), 0);
                else if(!found)
                    wf.hits[id] = make_int4(-1, -1, 0, 0);
                else
                {
                    const int b_inst = (int)COLD(COLD_INST), b_prim = (int)COLD(COLD_PRIM);
                    const float b_u = COLDF(COLD_U), b_v = COLDF(COLD_V);
                    bool accept = true;
                    if(transparency && !(COLD(COLD_FLAGS) & TR_TRI_FLAG_OPAQUE))
                    {
                        uint32_t seed = __float_as_uint(wf.dir_seed[id].w);
                        float alpha_cutoff = (float)pcg(seed) * TR_INV_UINT32_MAX;
                        wf.dir_seed[id].w = __uint_as_float(seed);
                        if(hit_alpha(P.scene, b_inst, b_prim, b_u, b_v) <= alpha_cutoff) accept = false;
                    }
                    if(accept) wf.hits[id] = make_int4(b_inst, b_prim, __float_as_int(b_u), __float_as_int(b_v));
                    else
                    {
                        tmin = best_t;
                        best_t = inf_();
                        sp = 0;
                        G = make_uint2(0u, (0x01000000u << rs.octinv) | 1u);
                        done = false;
                        retire = false;
                    }
                }
            }
            else
            {
                const float visibility = blocked ? 0.0f : 1.0f;
                NeeJob job;
                load_job(wf, id, (int)bounce, job);
                PathState s;
                const float4 e = wf.diffuse[id], f = wf.reflection[id];
                s.diffuse = V3(e.x, e.y, e.z);
                s.reflection = V3(f.x, f.y, f.z);
                nee_finish(P, job, visibility, s);
                wf.diffuse[id] = make_float4(s.diffuse.x, s.diffuse.y, s.diffuse.z, e.w);
                wf.reflection[id] = make_float4(s.reflection.x, s.reflection.y, s.reflection.z, 0.0f);
            }
            if(retire) has_ray = false;
        }

        // ---- refill idle lanes: one atomicAdd per warp ----------------------------------------
        {
            const bool want = !has_ray && !exhausted;
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, want);
            if(mask)
            {
                const int leader = __ffs(mask) - 1;
                uint32_t base = 0;
                if(lane == leader) base = atomicAdd(work_counter, (uint32_t)__popc(mask));
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                if(want)
                {
                    const uint32_t i = base + __popc(mask & lanes_below);
                    if(i < n)
                    {
                        const uint32_t id = queue[i];
                        const float4 o = wf.pos_pdf[id];
                        const vec3 org = V3(o.x, o.y, o.z);
                        vec3 dir;
                        if(ANY)
                        {
                            const float4 j0 = wf.nee[4 * (size_t)id];
                            dir = V3(j0.x, j0.y, j0.z);
                            best_t = j0.w;
                        }
                        else
                        {
                            const float4 d = wf.dir_seed[id];
                            dir = V3(d.x, d.y, d.z);
                            best_t = inf_();
                        }
                        tmin = ray_tmin;
                        const RayPre rp = precompute_ray(org, dir);
                        COLD(COLD_ID) = id;
                        COLDF(COLD_SX) = rp.Sx; COLDF(COLD_SY) = rp.Sy; COLDF(COLD_SZ) = rp.Sz; COLD(COLD_KZ) = (uint32_t)rp.kz;
                        COLD(COLD_FLAGS) = 0u;
                        if(!ANY) { COLD(COLD_INST) = 0x7FFFFFFFu; COLD(COLD_PRIM) = 0x7FFFFFFFu; }
                        rs = make_ray_slab(org, dir);
                        sp = 0;
                        blocked = false;
                        has_ray = true;
                        done = false;
                        rays++;
                        G = make_uint2(0u, (0x01000000u << rs.octinv) | 1u);
                    }
                    else exhausted = true;
                }
            }
            if(__ballot_sync(0xFFFFFFFFu, has_ray) == 0) break;
        }

        // ---- traverse: warp-uniform loop, one node step per lane and iteration ------------------------------
        for(;;)
        {
            bool posted = false;
            if(has_ray && !done)
            {
                uint2 Gt = make_uint2(0u, 0u);
                if(G.y & 0xFF000000u)
                {
                    const uint32_t bit = 31u - (uint32_t)__clz((int)G.y);
                    G.y &= ~(1u << bit);
                    const uint32_t child = child_of_bit(G, bit, rs.octinv);
                    if(G.y & 0xFF000000u) PUSH(G);
                    if(STATS) st.nodes++;
                    intersect_node8(nodes, child, rs, tmin, best_t, G, Gt);
                }
                if(Gt.y & 0xFFu)
                {
                    if(npost < POST_CAP) { s_post[npost * WF_BLOCK + tid] = Gt; npost++; posted = true; }
                    else
                    {
                        // the lane's post list is full: test this group right here (rare)
                        while(Gt.y & 0xFFu)
                        {
                            const uint32_t tri_index = pop_triangle(Gt, rs.octinv);
                            if(STATS) st.tris++;
                            const float4* tp = tris + 3 * (size_t)tri_index;
                            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                            RayPre rp;
                            rp.kz = (int)COLD(COLD_KZ); rp.Sx = COLDF(COLD_SX); rp.Sy = COLDF(COLD_SY); rp.Sz = COLDF(COLD_SZ);
                            rp.org = V3(rs.ox, rs.oy, rs.oz);
                            float t, u, v;
                            if(!intersect_triangle(rp, f3(t0), f3(t1), f3(t2), tmin, ANY ? best_t : inf_(), t, u, v)) continue;
                            if(ANY) { COLD(COLD_FLAGS) = 1u; blocked = true; done = true; sp = 0; G = make_uint2(0u, 0u); break; }
                            const int inst = __float_as_int(t0.w), prim = __float_as_int(t1.w);
                            bool better = t < best_t;
                            if(t == best_t)
                            {
                                const int b_inst = (int)COLD(COLD_INST), b_prim = (int)COLD(COLD_PRIM);
                                better = inst < b_inst || (inst == b_inst && prim < b_prim);
                            }
                            if(better)
                            {
                                COLD(COLD_INST) = (uint32_t)inst; COLD(COLD_PRIM) = (uint32_t)prim;
                                COLDF(COLD_U) = u; COLDF(COLD_V) = v; COLD(COLD_FLAGS) = __float_as_uint(t2.w);
                                best_t = t;
                            }
                        }
                    }
                }
                if(!done && (G.y & 0xFF000000u) == 0u)
                {
                    if(sp == 0) done = true;
                    else POP(G);
                }
            }
            wposted += __popc(__ballot_sync(0xFFFFFFFFu, posted));
            const uint32_t working = __ballot_sync(0xFFFFFFFFu, has_ray && !done);
            const uint32_t waiting = __ballot_sync(0xFFFFFFFFu, has_ray && done && npost > 0);
            if(wposted > 0 && (wposted >= P.tune_coop_groups || __popc(waiting) >= P.tune_coop_wait || working == 0u))
            {
                triangle_phase();
                wposted = 0;
            }
            // too few lanes left busy: go retire and refill
            if(__popc(__ballot_sync(0xFFFFFFFFu, has_ray && !(done && npost == 0))) < P.tune_fetch_threshold) break;
        }
    }
#undef PUSH
#undef POP
#undef COLD
#undef COLDF
#undef COLD_OF
    flush_counters(P.counters, ANY ? 0 : rays, ANY ? rays : 0, 0, st.nodes, st.tris, ANY ? 1 : 0);
}

// SPEC = 1: the option permutation of the reference's CLI defaults with the sobol-owen sampler on a scene without
// textures, sphere lights, environment map or alpha-tested geometry (BASELINE configs[2..4]) is compiled in: the
// option fields of the launch parameters are overwritten with the constants the host has checked them against, so the
// compiler folds every `if(P.opt...)` -- what the reference does with one shader permutation per option set
// (src/path_tracer_stage.cc:45-81, src/rt_common.cc:6-79). Same code, same arithmetic, same result.
template<int SPEC>
__global__ void __launch_bounds__(WF_BLOCK, SHADE_MIN_BLOCKS) k_wf_shade(const RenderParams P_in, const tr_wavefront wf, uint32_t bounce, uint32_t sample_in_pass,
                                                        const uint32_t* __restrict__ queue, const uint32_t* __restrict__ count,
                                                        uint32_t* __restrict__ next_queue, uint32_t* __restrict__ next_count,
                                                        uint32_t* __restrict__ shadow_queue, uint32_t* __restrict__ shadow_count)
{
    // Sobol byte-chunk tables of this bounce's two dimension groups (2b: NEE, 2b+1: BSDF) in shared memory
    RenderParams P = P_in;
    probe_begin(P);
    if(SPEC)
    {
        P.opt.local_sampler = TR_SAMPLER_SOBOL_OWEN; P.opt.film = TR_FILM_POINT; P.opt.mis_mode = TR_MIS_POWER_HEURISTIC;
        P.opt.bounce_mode = TR_BOUNCE_MATERIAL; P.opt.tri_light_mode = TR_TRI_LIGHT_SOLID_ANGLE;
        P.opt.russian_roulette_delta = 0.0f; P.opt.indirect_clamping = 0.0f; P.opt.regularization_gamma = 0.0f;
        P.opt.use_shadow_terminator_fix = 0; P.opt.use_white_albedo_on_first_bounce = 0; P.opt.hide_lights = 0;
        P.opt.transparent_background = 0; P.opt.depth_of_field = 0;
        P.opt.sampling_weight_point_lights = 1.0f; P.opt.sampling_weight_directional_lights = 1.0f;
        P.opt.sampling_weight_envmap = 1.0f; P.opt.sampling_weight_emissive_triangles = 1.0f;
        P.scene.point_light_count = 0; P.scene.environment_proj = -1; P.scene.any_non_opaque = 0; P.scene.has_textures = 0;
        P.tune_tma = 1;
    }
    __shared__ __align__(128) uint4 s_chunks[2 * 1024];
    __shared__ uint64_t s_mbar;
    SobolTables tables;
    tables.base = P.sobol_chunks;
    tables.first_group = 0;
    const bool stage_tables = P.opt.local_sampler != TR_SAMPLER_UNIFORM_RANDOM && 2 * bounce + 1 < P.frame.max_sobol_bounces;
    if(P.tune_tma & 1)
    {
        // one 32 KB bulk copy through the TMA unit instead of 2048 LDG.128 + STS.128 spread over the block
        if(threadIdx.x == 0) mbar_init(&s_mbar, 1);
        __syncthreads();
        if(stage_tables)
        {
            if(threadIdx.x == 0) tma_load_bulk(s_chunks, P.sobol_chunks + (size_t)2 * bounce * 1024, 2 * 1024 * sizeof(uint4), &s_mbar);
            mbar_wait(&s_mbar, 0);
            tables.base = s_chunks;
            tables.first_group = 2 * bounce;
        }
    }
    else
    {
        if(stage_tables)
        {
            for(int k = threadIdx.x; k < 2 * 1024; k += blockDim.x) s_chunks[k] = P.sobol_chunks[(size_t)2 * bounce * 1024 + k];
            tables.base = s_chunks;
            tables.first_group = 2 * bounce;
        }
        __syncthreads();
    }
    const uint32_t n = *count;
    const uint32_t rounded = (n + 31u) & ~31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    unsigned long long culled = 0;
    // The kernel is latency bound (4 warps per scheduler at 125 registers, a third of the stall samples on the state loads
    // below): the records of the path this thread shades NEXT are prefetched one iteration ahead (its id was loaded two
    // iterations ahead). tune_shade_prefetch: 0 off, 1 into L2, 2 into L1.
    const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t id_next = first + stride < n ? queue[first + stride] : 0u;
    for(uint32_t i = first; i < rounded; i += stride)
    {
        bool alive = false, shadow = false;
        uint32_t id = 0;
        if(P.tune_shade_prefetch)
        {
            const uint32_t id_pf = id_next;
            if(i + 2u * stride < n) id_next = queue[i + 2u * stride];
            if(i + stride < n)
            {
                if(P.tune_shade_prefetch == 1)
                {
                    prefetch_l2(wf.hits + id_pf); prefetch_l2(wf.pos_pdf + id_pf); prefetch_l2(wf.dir_seed + id_pf); prefetch_l2(wf.rs_seed + id_pf);
                    if(bounce != 0) { prefetch_l2(wf.atten_reg + id_pf); prefetch_l2(wf.lobes + id_pf); prefetch_l2(wf.diffuse + id_pf); prefetch_l2(wf.reflection + id_pf); }
                }
                else
                {
                    prefetch_l1(wf.hits + id_pf); prefetch_l1(wf.pos_pdf + id_pf); prefetch_l1(wf.dir_seed + id_pf); prefetch_l1(wf.rs_seed + id_pf);
                    if(bounce != 0) { prefetch_l1(wf.atten_reg + id_pf); prefetch_l1(wf.lobes + id_pf); prefetch_l1(wf.diffuse + id_pf); prefetch_l1(wf.reflection + id_pf); }
                }
            }
        }
        if(i < n)
        {
            id = queue[i];
            PathState s;
            if(bounce == 0)
            {
                // a fresh path (path_begin): k_wf_generate stores only the ray
                const float4 a = wf.pos_pdf[id], b = wf.dir_seed[id];
                s.pos = V3(a.x, a.y, a.z); s.bsdf_pdf = a.w;
                s.view = V3(b.x, b.y, b.z); s.random_seed = __float_as_uint(b.w);
                s.attenuation = V3(1); s.regularization = 1.0f;
                s.primary_lobes = LOBES(0, 0, 0, 1);
                s.diffuse = V3(0); s.reflection = V3(0); s.aux_alpha = 0.0f;
                s.fh_emission = V3(0); s.fh_albedo = V4(0, 0, 0, 0); s.fh_metallic = 1.0f;
            }
            else load_state(wf, id, s);
            const int4 hr = wf.hits[id];
            Hit h;
            h.instance_id = hr.x; h.primitive_id = hr.y; h.u = __int_as_float(hr.z); h.v = __int_as_float(hr.w); h.t = h.u; h.flags = 0;
            const uint4 sr = wf.rs_seed[id];
            uvec4 ss, rs;
            ss.x = sr.x; ss.y = sr.y; ss.z = sr.z; ss.w = sr.w;
            rs = ss;
            NeeJob job;
            alive = shade_bounce(P, tables, bounce, h, s, ss, rs, job);
            if(job.bounce >= 0)
            {
                if(nee_shadow_ray_is_moot(job)) culled++;
                else if(job.needs_shadow_ray) { store_job(wf, id, job); shadow = true; }
                else nee_finish(P, job, 1.0f, s);
            }
            // only what somebody reads again: the ray and the attenuation if the path goes on or a shadow ray starts from this
            // hit, the primary lobes once (they are set at bounce 0), the two accumulators always
            if(alive || shadow)
            {
                store_ray(wf, id, s);
                wf.atten_reg[id] = make_float4(s.attenuation.x, s.attenuation.y, s.attenuation.z, s.regularization);
            }
            if(bounce == 0)
                wf.lobes[id] = make_float4(s.primary_lobes.transmission, s.primary_lobes.diffuse,
                                           s.primary_lobes.dielectric_reflection, s.primary_lobes.metallic_reflection);
            wf.diffuse[id] = make_float4(s.diffuse.x, s.diffuse.y, s.diffuse.z, s.aux_alpha);
            wf.reflection[id] = make_float4(s.reflection.x, s.reflection.y, s.reflection.z, 0.0f);
            if(bounce == 0) store_first_hit(wf, id, s);
            if(P.opt.local_sampler != TR_SAMPLER_SOBOL_OWEN) wf.rs_seed[id] = make_uint4(rs.x, rs.y, rs.z, rs.w);
        }
        warp_push(next_queue, next_count, alive, id);
        warp_push(shadow_queue, shadow_count, shadow, id);
    }
    for(int off = 16; off > 0; off >>= 1) culled += __shfl_xor_sync(0xFFFFFFFFu, culled, off);
    if((threadIdx.x & 31) == 0 && culled) atomicAdd(&P.counters->culled_shadow_rays, culled);
    probe_end(P);
}

__global__ void __launch_bounds__(WF_BLOCK) k_wf_finish(const RenderParams P, const tr_wavefront wf, uint32_t sample_in_pass,
                                                         uint32_t base_slots, uint32_t batch)
{
    probe_begin(P);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t spp = (uint32_t)P.opt.samples_per_pass;
    uint2* const diffuse_layer = reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_DIFFUSE]);
    uint2* const reflection_layer = reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_REFLECTION]);
    const bool light_layers = diffuse_layer || reflection_layer;
    // the first pass of a frame continues from what the previous frame left (see RenderParams::color_prev)
    const bool frame_start = P.frame.previous_samples == 0;
    const uint2* const diffuse_prev = reinterpret_cast<const uint2*>(P.gbuffer_prev[TR_GBUFFER_DIFFUSE]);
    const uint2* const reflection_prev = reinterpret_cast<const uint2*>(P.gbuffer_prev[TR_GBUFFER_REFLECTION]);
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < base_slots; slot += stride)
    {
        uint32_t px, py, view, widx;
        if(!map_slot(P.frame, slot, px, py, view, widx)) continue;
        float4 o = make_float4(0, 0, 0, 0);
        // accumulate the batched passes in pass order (gbuffer.glsl:18-28 is an order-dependent running mean)
        for(uint32_t pass = 0; pass < batch; ++pass)
        {
            const uint32_t id = pass * base_slots + slot;
            PathState s;
            const float4 e = wf.diffuse[id], f = wf.reflection[id];
            s.diffuse = V3(e.x, e.y, e.z);
            s.reflection = V3(f.x, f.y, f.z);
            load_first_hit(wf, id, s);
            vec3 c = path_color(P, s);
            vec4 dif = V4(s.diffuse, e.w), ref = V4(s.reflection, e.w);
            if(spp > 1)   // batch == 1 here: the samples of one pass are summed first (path_tracer.rgen:81-126)
            {
                vec3 sum = V3(0);
                vec4 sum_d = V4(0, 0, 0, 0), sum_r = V4(0, 0, 0, 0);
                if(sample_in_pass != 0)
                {
                    const float4 q = wf.sum_color[slot];
                    sum = V3(q.x, q.y, q.z);
                    if(light_layers)
                    {
                        const float4 qd = wf.sum_color[base_slots + slot], qr = wf.sum_color[2 * base_slots + slot];
                        sum_d = V4(qd.x, qd.y, qd.z, qd.w); sum_r = V4(qr.x, qr.y, qr.z, qr.w);
                    }
                }
                sum += c;
                sum_d = sum_d + dif; sum_r = sum_r + ref;
                if(sample_in_pass + 1 < spp)
                {
                    wf.sum_color[slot] = make_float4(sum.x, sum.y, sum.z, 0.0f);
                    if(light_layers)
                    {
                        wf.sum_color[base_slots + slot] = make_float4(sum_d.x, sum_d.y, sum_d.z, sum_d.w);
                        wf.sum_color[2 * base_slots + slot] = make_float4(sum_r.x, sum_r.y, sum_r.z, sum_r.w);
                    }
                    continue;
                }
                c = sum / (float)spp;
                dif = sum_d / (float)spp; ref = sum_r / (float)spp;
            }
            const float alpha = P.opt.transparent_background ? s.fh_albedo.w : 1.0f;
            o = accumulate(P, c, alpha, widx, pass);
            if(light_layers)
            {
                const uint32_t prev_samples = P.frame.samples_accumulated + P.frame.previous_samples + pass * spp;
                const bool from_prev = frame_start && pass == 0;
                if(diffuse_layer) accumulate_rgba16f(diffuse_layer, from_prev ? diffuse_prev : diffuse_layer, widx, dif, spp, prev_samples);
                if(reflection_layer) accumulate_rgba16f(reflection_layer, from_prev ? reflection_prev : reflection_layer, widx, ref, spp, prev_samples);
            }
        }
        if(spp > 1 && sample_in_pass + 1 < spp) continue;
        store_to_peers(P, o, px, py, view);
    }
    probe_end(P);
}

// The G-buffer layers only the first sample writes (write_all_outputs, path_tracer.glsl:547-563): recomputed
// from the bounce-0 hit records, between the bounce-0 traversal and its shade pass (which overwrites the ray).
// Launched only while no sample has been accumulated and a layer is present -- never on the steady-state path.
__global__ void __launch_bounds__(WF_BLOCK) k_wf_first_sample_layers(const RenderParams P, const tr_wavefront wf, uint32_t base_slots)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < base_slots; slot += stride)
    {
        uint32_t px, py, view, widx;
        if(!map_slot(P.frame, slot, px, py, view, widx)) continue;
        const float4 o = wf.pos_pdf[slot], d = wf.dir_seed[slot];
        const int4 hr = wf.hits[slot];
        Hit h;
        h.instance_id = hr.x; h.primitive_id = hr.y; h.u = __int_as_float(hr.z); h.v = __int_as_float(hr.w); h.t = h.u; h.flags = 0;
        HitInfo hi;
        get_intersection_info(P, h, V3(o.x, o.y, o.z), V3(d.x, d.y, d.z), hi);
        write_first_sample_layers(P, widx, view, hi);
    }
}

// where a per-pass kernel reads the running mean of a light layer: the first pass of a frame continues from what the
// previous frame left (RenderParams::gbuffer_prev), later passes from what this frame wrote
TRD const uint2* light_layer_prev(const RenderParams& P, int layer)
{
    return reinterpret_cast<const uint2*>(P.frame.previous_samples == 0 ? P.gbuffer_prev[layer] : P.gbuffer[layer]);
}

// ---- megakernel ----------------------------------------------------------------------------------
template<bool STATS>
__global__ void __launch_bounds__(WF_BLOCK) k_megakernel(const RenderParams P, uint32_t total_slots)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    TraceStats st; st.nodes = 0; st.tris = 0;
    TraceStats sst; sst.nodes = 0; sst.tris = 0;
    unsigned long long closest = 0, shadow = 0, paths = 0, culled = 0;
    SobolTables tables;
    tables.base = P.sobol_chunks;
    tables.first_group = 0;
    const uint32_t spp = (uint32_t)P.opt.samples_per_pass;
    const bool first_layers = (P.gbuffer_mask & ~((1u << TR_GBUFFER_COLOR) | (1u << TR_GBUFFER_DIFFUSE) | (1u << TR_GBUFFER_REFLECTION))) != 0;
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total_slots; slot += stride)
    {
        uint32_t px, py, view, widx;
        if(!map_slot(P.frame, slot, px, py, view, widx)) continue;
        vec3 sum = V3(0);
        vec4 sum_d = V4(0, 0, 0, 0), sum_r = V4(0, 0, 0, 0);
        float alpha = 1.0f;
        const uint32_t prev_samples = P.frame.samples_accumulated + P.frame.previous_samples;
        for(uint32_t i = 0; i < spp; ++i)
        {
            PathState s;
            uvec4 ss, rs;
            path_begin(P, px, py, view, i, s, ss, rs);
            paths++;
            for(uint32_t bounce = 0; bounce < (uint32_t)P.opt.max_ray_depth; ++bounce)
            {
                Hit h;
                closest_hit<STATS>(P, s.pos, bounce == 0 ? 0.0f : P.opt.min_ray_dist, s.view,
                                   !(P.opt.hide_lights && bounce == 0), s.random_seed, h, st);
                closest++;
                if(bounce == 0 && first_layers && prev_samples == 0 && i + 1 == spp)
                {
                    HitInfo hi;
                    get_intersection_info(P, h, s.pos, s.view, hi);
                    write_first_sample_layers(P, widx, view, hi);
                }
                NeeJob job;
                const bool alive = shade_bounce(P, tables, bounce, h, s, ss, rs, job);
                if(job.bounce >= 0 && nee_shadow_ray_is_moot(job)) culled++;
                else if(job.bounce >= 0)
                {
                    float vis = 1.0f;
                    if(job.needs_shadow_ray)
                    {
                        vis = shadow_visibility<STATS>(P, s.pos, P.opt.min_ray_dist, job.dir, job.tmax, sst);
                        shadow++;
                    }
                    nee_finish(P, job, vis, s);
                }
                if(!alive) break;
            }
            sum += path_color(P, s);
            sum_d = sum_d + V4(s.diffuse, s.aux_alpha);
            sum_r = sum_r + V4(s.reflection, s.aux_alpha);
            alpha = P.opt.transparent_background ? s.fh_albedo.w : 1.0f;
        }
        const float4 o = accumulate(P, sum / (float)spp, alpha, widx);
        if(P.gbuffer[TR_GBUFFER_DIFFUSE])
            accumulate_rgba16f(reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_DIFFUSE]), light_layer_prev(P, TR_GBUFFER_DIFFUSE), widx, sum_d / (float)spp, spp, prev_samples);
        if(P.gbuffer[TR_GBUFFER_REFLECTION])
            accumulate_rgba16f(reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_REFLECTION]), light_layer_prev(P, TR_GBUFFER_REFLECTION), widx, sum_r / (float)spp, spp, prev_samples);
        store_to_peers(P, o, px, py, view);
    }
    flush_counters(P.counters, closest, shadow, paths, st.nodes, st.tris, 0);
    flush_counters(P.counters, 0, 0, 0, sst.nodes, sst.tris, 1);
    for(int off = 16; off > 0; off >>= 1) culled += __shfl_xor_sync(0xFFFFFFFFu, culled, off);
    if((threadIdx.x & 31) == 0 && culled) atomicAdd(&P.counters->culled_shadow_rays, culled);
}

// ---- direct-light estimator: direct_stage (src/direct_stage.cc:29-126) + shader/direct.rgen:50-132 ----------------
// One camera ray per pass (sphere lights masked out), then samples_per_pass next-event samples at the first hit.
// The sampler is keyed by previous_samples alone (direct.rgen:106-113); MIS is off (tr_pt_create forces
// mis_mode = DISABLED for this estimator, as direct_stage defines no MIS macro). main() divides colour and
// diffuse (alpha included) by SAMPLES_PER_PASS but not reflection -- kept.
template<bool STATS>
__global__ void __launch_bounds__(WF_BLOCK) k_direct(const RenderParams P, uint32_t total_slots)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    TraceStats st; st.nodes = 0; st.tris = 0;
    TraceStats sst; sst.nodes = 0; sst.tris = 0;
    unsigned long long closest = 0, shadow = 0, paths = 0;
    SobolTables tables;
    tables.base = P.sobol_chunks;
    tables.first_group = 0;
    const uint32_t spp = (uint32_t)P.opt.samples_per_pass;
    const bool first_layers = (P.gbuffer_mask & ~((1u << TR_GBUFFER_COLOR) | (1u << TR_GBUFFER_DIFFUSE) | (1u << TR_GBUFFER_REFLECTION))) != 0;
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total_slots; slot += stride)
    {
        uint32_t px, py, view, widx;
        if(!map_slot(P.frame, slot, px, py, view, widx)) continue;
        const uint32_t prev_samples = P.frame.samples_accumulated + P.frame.previous_samples;
        PathState s;
        uvec4 ss, rs;
        path_begin(P, px, py, view, 0, s, ss, rs);
        paths++;
        Hit h;
        closest_hit<STATS>(P, s.pos, 0.0f, s.view, false, s.random_seed, h, st);
        closest++;
        HitInfo hi;
        get_intersection_info(P, h, s.pos, s.view, hi);
        if(first_layers && prev_samples == 0) write_first_sample_layers(P, widx, view, hi);
        vec3 color = (hi.light + hi.mat.emission) * (float)spp;
        float a = 0.0f;
        if(hi.surface)
        {
            const mat3 tbn = create_tangent_space(hi.mapped_normal);
            const vec3 shading_view = view_to_tangent_space(s.view, tbn);
            for(uint32_t i = 0; i < spp; ++i)
            {
                // next_event_estimation (path_tracer.glsl:302-344); the draw is made even when no light class is on
                const uvec4 nee_rand = generate_ray_sample_uint(P, tables, ss, rs, i);
                bsdf_lobes lobes = LOBES(0, 0, 0, 0);
                vec3 radiance = V3(0);
                if(nee_enabled(P))
                {
                    vec3 out_dir;
                    float out_length = 0.0f;
                    float light_pdf;
                    vec3 contrib = sample_explicit_light(P, nee_rand, hi.pos, out_dir, out_length, light_pdf);
                    const vec3 shading_light = mul(out_dir, tbn);
                    const float bsdf_pdf = material_bsdf_pdf(P.opt.bounce_mode, shading_light, shading_view, hi.mat, lobes);
                    correct_lobes_for_normal_map(out_dir, hi.hard_normal, lobes);
                    if(contrib.x > 0.0001f || contrib.y > 0.0001f || contrib.z > 0.0001f)
                    {
                        contrib *= shadow_visibility<STATS>(P, hi.pos, P.opt.min_ray_dist, out_dir, out_length, sst);
                        shadow++;
                    }
                    contrib /= nee_mis_pdf(P, light_pdf, bsdf_pdf);
                    radiance = contrib;
                }
                color += radiance * modulate_bsdf(xyz(hi.mat.albedo), lobes);
                add_demodulated(s, radiance, lobes.diffuse + lobes.transmission, lobes.dielectric_reflection + lobes.metallic_reflection);
                a = 1.0f / length(hi.pos - s.pos);
            }
        }
        const float alpha = P.opt.transparent_background ? hi.mat.albedo.w : 1.0f;
        const float4 o = accumulate(P, color / (float)spp, alpha, widx);
        if(P.gbuffer[TR_GBUFFER_DIFFUSE])
            accumulate_rgba16f(reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_DIFFUSE]), light_layer_prev(P, TR_GBUFFER_DIFFUSE), widx, V4(s.diffuse, a) / (float)spp, spp, prev_samples);
        if(P.gbuffer[TR_GBUFFER_REFLECTION])
            accumulate_rgba16f(reinterpret_cast<uint2*>(P.gbuffer[TR_GBUFFER_REFLECTION]), light_layer_prev(P, TR_GBUFFER_REFLECTION), widx, V4(s.reflection, a), spp, prev_samples);
        store_to_peers(P, o, px, py, view);
    }
    flush_counters(P.counters, closest, shadow, paths, st.nodes, st.tris, 0);
    flush_counters(P.counters, 0, 0, 0, sst.nodes, sst.tris, 1);
}

// ---- feature (AOV) tracer: shader/rt_feature.rgen:18-47, rt_feature.rahit, rt_feature.rchit ------
__global__ void __launch_bounds__(WF_BLOCK) k_feature(const RenderParams P, int feature, float4 default_value, uint32_t total_slots)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    TraceStats st; st.nodes = 0; st.tris = 0;
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total_slots; slot += stride)
    {
        uint32_t px, py, view, widx;
        if(!map_slot(P.frame, slot, px, py, view, widx)) continue;
        const tr_camera_data& cam = P.cameras[view].current;
        vec3 origin, dir;
        get_screen_camera_ray(P, cam, px, py, V2(0, 0), V2(0.5f, 0.5f), origin, dir);
        const vec3 o = V3(cam.origin[0], cam.origin[1], cam.origin[2]);
        vec4 data = V4(default_value.x, default_value.y, default_value.z, default_value.w);
        float cur_tmin = P.opt.min_ray_dist;
        for(;;)
        {
            Hit h;
            if(!traverse<false, false>(P.scene.bvh, o, dir, cur_tmin, inf_(), h, st, [](int, int, float, float, float) {})) break;
            if(!(h.flags & TR_TRI_FLAG_OPAQUE) && hit_alpha(P.scene, h.instance_id, h.primitive_id, h.u, h.v) <= 0.0001f)
            {
                cur_tmin = h.t;
                continue;
            }
            const DeviceInstance& inst = P.scene.instances[h.instance_id];
            float pdf;
            SurfaceVertex v = get_interpolated_vertex(P, inst, dir, h.u, h.v, h.primitive_id, o, pdf);
            sampled_material mat = sample_material(P, inst, v);
            switch(feature)
            {
            default:
            case TR_FEATURE_ALBEDO: data = mat.albedo; break;
            case TR_FEATURE_WORLD_NORMAL: data = V4(v.mapped_normal, 1); break;
            case TR_FEATURE_VIEW_NORMAL: data = V4(xyz(mul4(cam.view, V4(v.mapped_normal, 0))), 1); break;
            case TR_FEATURE_WORLD_POS: data = V4(v.pos, 1); break;
            case TR_FEATURE_VIEW_POS: data = mul4(cam.view, V4(v.pos, 1)); break;
            case TR_FEATURE_DISTANCE: data = V4(h.t, h.t, h.t, 1); break;
            case TR_FEATURE_INSTANCE_ID: data = V4((float)h.instance_id, (float)h.primitive_id, 0, 1); break;
            }
            break;
        }
        P.color[widx] = make_float4(data.x, data.y, data.z, data.w);
    }
}

// ---- aux G-buffer layers across devices -----------------------------------------------------------
// The reference transfers and stitches every gbuffer_target entry, not only the colour (src/rt_renderer.cc:370-397,
// src/stitch_stage.cc:156-171). Here one kernel at the end of the frame copies this rank's texels of every enabled aux
// layer to their un-shuffled position in every peer's full-image arena (NVLink stores). The running-mean layers
// (diffuse / reflection) take the stitch blend like the colour; the layers only the first sample writes are copied as
// they are. `src[l]` is where this rank's current texels of layer l live, `src_full[l]` says whether that is indexed by
// full-image pixel (the primary's in-place targets and carried-over arena halves) or by the rank's own write index.
struct AuxStitchSources
{
    const void* src[TR_GBUFFER_LAYER_COUNT];
    int src_full[TR_GBUFFER_LAYER_COUNT];
};

template<typename T> TRD void copy_texel_to_peers(const RenderParams& P, int layer, const void* src, size_t src_index, size_t dst)
{
    const T v = reinterpret_cast<const T*>(src)[src_index];
    for(int k = 0; k < P.peer_count; ++k)
    {
        T* out = peer_texels<T>(P, k, P.full_cur, layer);
        if(static_cast<const void*>(out) == src) continue;    // rendered in place
        out[dst] = v;
    }
}

__global__ void __launch_bounds__(WF_BLOCK) k_stitch_aux_layers(const RenderParams P, const AuxStitchSources S, uint32_t base_slots)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for(uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < base_slots; slot += stride)
    {
        uint32_t px, py, view, widx;
        if(!map_slot(P.frame, slot, px, py, view, widx)) continue;
        const size_t dst = (size_t)view * P.frame.size_x * P.frame.size_y + (size_t)py * P.frame.size_x + px;
        for(int l = TR_GBUFFER_DIFFUSE; l <= TR_GBUFFER_REFLECTION; ++l)
        {
            if(!S.src[l]) continue;
            const size_t si = S.src_full[l] ? dst : (size_t)widx;
            if(P.stitch_blend < 1.0f)
            {
                const vec4 v = unpack_rgba16f(reinterpret_cast<const uint2*>(S.src[l])[si]);
                for(int k = 0; k < P.peer_count; ++k)
                {
                    const int from = P.peer_self >= 0 ? P.peer_self : k;
                    const vec4 old = unpack_rgba16f(peer_texels<uint2>(P, from, P.full_prev, l)[dst]);
                    peer_texels<uint2>(P, k, P.full_cur, l)[dst] = pack_rgba16f(mix_(old, v, P.stitch_blend));
                }
            }
            else copy_texel_to_peers<uint2>(P, l, S.src[l], si, dst);
        }
        if(S.src[TR_GBUFFER_ALBEDO]) copy_texel_to_peers<uint2>(P, TR_GBUFFER_ALBEDO, S.src[TR_GBUFFER_ALBEDO], S.src_full[TR_GBUFFER_ALBEDO] ? dst : (size_t)widx, dst);
        if(S.src[TR_GBUFFER_MATERIAL]) copy_texel_to_peers<uint32_t>(P, TR_GBUFFER_MATERIAL, S.src[TR_GBUFFER_MATERIAL], S.src_full[TR_GBUFFER_MATERIAL] ? dst : (size_t)widx, dst);
        if(S.src[TR_GBUFFER_NORMAL]) copy_texel_to_peers<uint32_t>(P, TR_GBUFFER_NORMAL, S.src[TR_GBUFFER_NORMAL], S.src_full[TR_GBUFFER_NORMAL] ? dst : (size_t)widx, dst);
        if(S.src[TR_GBUFFER_POS]) copy_texel_to_peers<float4>(P, TR_GBUFFER_POS, S.src[TR_GBUFFER_POS], S.src_full[TR_GBUFFER_POS] ? dst : (size_t)widx, dst);
        if(S.src[TR_GBUFFER_SCREEN_MOTION]) copy_texel_to_peers<float4>(P, TR_GBUFFER_SCREEN_MOTION, S.src[TR_GBUFFER_SCREEN_MOTION], S.src_full[TR_GBUFFER_SCREEN_MOTION] ? dst : (size_t)widx, dst);
        if(S.src[TR_GBUFFER_INSTANCE_ID]) copy_texel_to_peers<int32_t>(P, TR_GBUFFER_INSTANCE_ID, S.src[TR_GBUFFER_INSTANCE_ID], S.src_full[TR_GBUFFER_INSTANCE_ID] ? dst : (size_t)widx, dst);
    }
}

// ---- test hooks ----------------------------------------------------------------------------------
struct OutHit { int inst, prim; float u, v, t; };

__global__ void k_trace_rays(const RenderParams P, const float* __restrict__ rays, uint32_t n, OutHit* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float* r = rays + 8 * (size_t)i;
    Hit h;
    TraceStats st; st.nodes = 0; st.tris = 0;
    bool ok = traverse<false, false>(P.scene.bvh, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], r[7], h, st,
                                     [](int, int, float, float, float) {});
    OutHit o;
    if(ok) { o.inst = h.instance_id; o.prim = h.primitive_id; o.u = h.u; o.v = h.v; o.t = h.t; }
    else { o.inst = -1; o.prim = -1; o.u = 0; o.v = 0; o.t = 0; }
    out[i] = o;
}

__global__ void k_sample_test(const RenderParams P, const uint32_t* __restrict__ coords, uint32_t n, int draw_index, uint32_t* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    local_sampler ls = init_local_sampler(coords[4 * i], coords[4 * i + 1], coords[4 * i + 2], coords[4 * i + 3],
                                          P.frame.sample_counter, P.frame.rng_seed, P.opt.local_sampler);
    uvec4 r;
    if(draw_index < 0) r = pcg4d(ls.rs_seed);
    else
    {
        SobolTables tables;
        tables.base = P.sobol_chunks;
        tables.first_group = 0;
        r = generate_ray_sample_uint(P, tables, ls.ss_seed, ls.rs_seed, (uint32_t)draw_index);
    }
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// stitch_shuffled_strips.comp:21-63 for one source strip
__global__ void k_unshuffle(const FrameParams f, const float4* __restrict__ strip, uint32_t index, uint32_t count, float4* __restrict__ full,
                            float blend_ratio)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t view = blockIdx.y;
    if(i >= count) return;
    const uint32_t j = permute_region_id(f, index + i);
    if(j >= f.size_x * f.size_y) return;
    float4 o = strip[(size_t)view * count + i];
    float4* dst = full + (size_t)view * f.size_x * f.size_y + j;
    if(blend_ratio < 1.0f)
    {
        const float4 q = *dst;
        const vec4 m = mix_(V4(q.x, q.y, q.z, q.w), V4(o.x, o.y, o.z, o.w), blend_ratio);
        o = make_float4(m.x, m.y, m.z, m.w);
    }
    *dst = o;
}

} // namespace trd

using namespace trd;

cudaError_t tr_render_init_constants()
{
    std::vector<uint4> chunks(8 * 4 * 256);
    for(int g = 0; g < 8; ++g)
        for(int c = 0; c < 4; ++c)
            for(int v = 0; v < 256; ++v)
            {
                uint4 x = make_uint4(0, 0, 0, 0);
                for(int j = 0; j < 8; ++j)
                    if((v >> j) & 1)
                    {
                        const unsigned int* row = TR_SOBOL_TABLE[g * 32 + c * 8 + j];
                        x.x ^= row[0]; x.y ^= row[1]; x.z ^= row[2]; x.w ^= row[3];
                    }
                chunks[(size_t)(g * 4 + c) * 256 + v] = x;
            }
    return cudaMemcpyToSymbol(g_sobol_chunks, chunks.data(), chunks.size() * sizeof(uint4));
}

const uint4* tr_render_sobol_chunks()
{
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_sobol_chunks) == cudaSuccess ? (const uint4*)p : nullptr;
}

static int grid_for(int sm_count, int blocks_per_sm, uint32_t work)
{
    long long need = ((long long)work + WF_BLOCK - 1) / WF_BLOCK;
    long long cap = (long long)sm_count * blocks_per_sm;
    long long g = need < cap ? need : cap;
    return (int)(g < 1 ? 1 : g);
}

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

cudaError_t tr_render_pass_wavefront(cudaStream_t stream, const RenderParams& P_in, const tr_wavefront& wf, int sm_count, uint64_t* launches,
                                     tr_launch_timer* timer, uint32_t batch, const tr_overlap* overlap, int phase)
{
    // per-launch timing needs one stream; TR_TUNE_OVERLAP=0 switches the second stream off for A/B runs
    static const int k_overlap = env_int("TR_TUNE_OVERLAP", 1);
    const bool two_streams = overlap && overlap->aux && !timer && k_overlap;
    RenderParams P = P_in;
    static const int k_fetch = env_int("TR_TUNE_FETCH", FETCH_THRESHOLD);
    static const int k_postpone = env_int("TR_TUNE_POSTPONE", 4);
    static const int k_blocks = env_int("TR_TUNE_BLOCKS", TRACE_MIN_BLOCKS_CLOSEST);
    static const int k_blocks_shadow = env_int("TR_TUNE_BLOCKS_SHADOW", TRACE_MIN_BLOCKS_SHADOW);
    P.tune_fetch_threshold = k_fetch;
    P.tune_postpone_div = k_postpone;
    // warp-cooperative triangle phase (k_wf_trace_coop): TR_TUNE_COOP bit 0 = closest-hit kernel, bit 1 = shadow kernel.
    // The pair queue packs the triangle index into 27 bits; the shadow variant needs an all-opaque scene.
    // TMA staging (tr_tma.cuh): bit 0 = the shade kernel's Sobol chunk tables, bit 1 = the top of the node array in the
    // traversal kernels
    static const int k_tma = env_int("TR_TUNE_TMA", 1);
    P.tune_tma = k_tma;
    static const int k_nearest_early = env_int("TR_TUNE_NEAREST", 0) || env_int("TR_TUNE_COOP", 0);
    // closest-hit kernel: descend into the inner child with the smallest entry distance first (TR_TUNE_NEAREST)
    static const int k_nearest = env_int("TR_TUNE_NEAREST", 0);
    // tail splitting (k_wf_trace<..., SPLIT>): bit 0 = closest-hit kernel, bit 1 = shadow kernel
    static const int k_split = env_int("TR_TUNE_SPLIT", 0);
    // fused traversal (k_wf_trace<..., FUSED>): closest-hit rays of bounce b + 1 and shadow rays of bounce b in one launch;
    // bit 1: with the tail-splitting stage compiled in
    static const int k_fuse = env_int("TR_TUNE_FUSE", 0);
    const bool fused = k_fuse && !timer && !P.stats && !P.scene.any_non_opaque && !(P.tune_tma & 2) && !k_nearest_early;
    // TR_TUNE_SPLIT_STAGE=0 keeps the SPLIT kernels but never runs their tail stage (A/B of the stage itself)
    static const int k_split_stage = env_int("TR_TUNE_SPLIT_STAGE", 1);
    P.tune_split_stage = k_split_stage;
    static const int k_split_steps = env_int("TR_TUNE_SPLIT_STEPS", 96);
    P.tune_split_steps = k_split_steps < 2 ? 2 : (k_split_steps > 255 ? 255 : k_split_steps);
    // TR_DEBUG_DRAIN=1 (development aid): per-launch start / queue-empty / exit times of the traversal kernels on stderr
    static const int k_probe = env_int("TR_DEBUG_DRAIN", 0);
    static unsigned long long* probe_dev = nullptr;
    static unsigned long long probe_init[136 * 2 * 16 + 2 * 128];
    static const char* probe_names[128];
    int probe_launch = 0;
    static const int k_probe_bin = env_int("TR_DEBUG_DRAIN_BIN_US", 16);
    P.tune_probe_bin_ns = (unsigned long long)k_probe_bin * 1000ull;
    P.probe = nullptr;
    if(k_probe && (uint32_t)P.opt.max_ray_depth <= 16u)
    {
        if(!probe_dev && cudaMalloc(&probe_dev, sizeof(probe_init)) != cudaSuccess) probe_dev = nullptr;
        if(probe_dev)
        {
            for(int k = 0; k < 2 * 16; ++k) { for(int j = 0; j < 136; ++j) probe_init[136 * k + j] = 0ull; probe_init[136 * k] = probe_init[136 * k + 1] = ~0ull; }
            for(int k = 0; k < 128; ++k) { probe_init[136 * 2 * 16 + 2 * k] = ~0ull; probe_init[136 * 2 * 16 + 2 * k + 1] = 0ull; }
            cudaMemcpyAsync(probe_dev, probe_init, sizeof(probe_init), cudaMemcpyHostToDevice, stream);
            P.probe = probe_dev;
        }
    }
    static const int k_coop = env_int("TR_TUNE_COOP", 0);
    static const int k_coop_groups = env_int("TR_TUNE_COOP_GROUPS", 24);
    static const int k_coop_wait = env_int("TR_TUNE_COOP_WAIT", 6);
    P.tune_coop_groups = k_coop_groups;
    P.tune_coop_wait = k_coop_wait;
    static const int k_shade_spec = env_int("TR_TUNE_SHADE_SPEC", 1);
    static const int k_shade_prefetch = env_int("TR_TUNE_SHADE_PREFETCH", 1);
    P.tune_shade_prefetch = k_shade_prefetch;
    const tr_pt_options& o = P.opt;
    const bool shade_spec = k_shade_spec && (P.tune_tma & 1) &&
        o.local_sampler == TR_SAMPLER_SOBOL_OWEN && o.film == TR_FILM_POINT && o.mis_mode == TR_MIS_POWER_HEURISTIC &&
        o.bounce_mode == TR_BOUNCE_MATERIAL && o.tri_light_mode == TR_TRI_LIGHT_SOLID_ANGLE && o.russian_roulette_delta == 0.0f &&
        o.indirect_clamping == 0.0f && o.regularization_gamma == 0.0f && !o.use_shadow_terminator_fix && !o.use_white_albedo_on_first_bounce &&
        !o.hide_lights && !o.transparent_background && !o.depth_of_field && o.sampling_weight_point_lights == 1.0f &&
        o.sampling_weight_directional_lights == 1.0f && o.sampling_weight_envmap == 1.0f && o.sampling_weight_emissive_triangles == 1.0f &&
        P.scene.point_light_count == 0 && P.scene.environment_proj < 0 && !P.scene.any_non_opaque && !P.scene.has_textures;
    const bool coop_ok = P.scene.bvh.tri_count < (1u << 27);
    const bool coop_closest = coop_ok && (k_coop & 1);
    const bool coop_shadow = coop_ok && (k_coop & 2) && !P.scene.any_non_opaque;
#define MARK(cls) do { if(timer) timer->mark(stream, cls); } while(0)
#define SLOT(name) do { if(P.probe && probe_launch < 127) { probe_names[probe_launch] = name; P.probe_slot = probe_launch++; } } while(0)
    const uint32_t base = P.frame.ray_count_x * P.frame.ray_count_y * P.frame.views;
    if(base == 0) return cudaSuccess;
    const uint32_t D = (uint32_t)P.opt.max_ray_depth;
    const uint32_t spp = (uint32_t)P.opt.samples_per_pass;
    if(spp > 1) batch = 1;
    const uint32_t total = base * batch;
    const int grid_full = grid_for(sm_count, 16, total);
    const int grid_finish = grid_for(sm_count, 16, base);
    const int grid_persist = sm_count * k_blocks;            // persistent grids: resident blocks per SM = the launch bounds
    const int grid_persist_shadow = sm_count * k_blocks_shadow;
    static const int k_shade_blocks = env_int("TR_TUNE_SHADE_BLOCKS", 4);
    const int grid_shade = sm_count * k_shade_blocks;   // 128 registers/thread: 4 resident blocks per SM
    // layers written only while prev_samples == 0 (pass 0 = the first `base` slots of a batch)
    const bool first_sample_layers = (P.gbuffer_mask & ~((1u << TR_GBUFFER_COLOR) | (1u << TR_GBUFFER_DIFFUSE) | (1u << TR_GBUFFER_REFLECTION))) != 0 &&
                                     P.frame.samples_accumulated + P.frame.previous_samples == 0;
    if(phase != TR_PHASE_ALL && spp != 1) return cudaErrorInvalidValue;
    if(phase == TR_PHASE_FINISH)
    {
        // the second half of a frame in flight: only the accumulation (and the peer stores), on the caller's stream
        k_wf_finish<<<grid_finish, WF_BLOCK, 0, stream>>>(P, wf, 0, base, batch);
        ++*launches;
        return cudaGetLastError();
    }
    for(uint32_t i = 0; i < spp; ++i)
    {
        cudaError_t e = cudaMemsetAsync(wf.counts, 0, sizeof(uint32_t) * (4 * D + 2), stream);
        if(e != cudaSuccess) return e;
        uint32_t* active_counts = wf.counts;           // [D + 1]
        uint32_t* shadow_counts = wf.counts + D + 1;   // [D]
        uint32_t* work_counters = wf.counts + 2 * D + 1; // [2 D] dynamic-fetch cursors
        bool shadow_pending = false;
        MARK(TR_KCLASS_OTHER);
        SLOT("generate");
        k_wf_generate<<<grid_full, WF_BLOCK, 0, stream>>>(P, wf, i, base, total, wf.queue[0], active_counts);
        ++*launches;
        for(uint32_t b = 0; b < D; ++b)
        {
            uint32_t* q_in = wf.queue[b & 1];
            uint32_t* q_out = wf.queue[(b + 1) & 1];
            MARK(TR_KCLASS_CLOSEST);
            SLOT(fused && b > 0 ? "fused" : "closest");
            if(fused && b > 0)
            {
                // closest(b) + shadow(b - 1): one work list, one drain phase
                if(k_fuse & 2) k_wf_trace<false, false, false, false, true, true><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b,
                                                                                                                         wf.shadow_queue, shadow_counts + b - 1);
                else k_wf_trace<false, false, false, false, false, true><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b,
                                                                                                               wf.shadow_queue, shadow_counts + b - 1);
            }
            else if(coop_closest)
            {
                if(P.stats) k_wf_trace_coop<false, true><<<grid_persist, WF_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
                else k_wf_trace_coop<false, false><<<grid_persist, WF_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
            }
            else if(P.stats) k_wf_trace<false, true><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
            else if(P.tune_tma & 2) k_wf_trace<false, false, true><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
            else if(k_nearest) k_wf_trace<false, false, false, true><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
            else if(k_split & 1) k_wf_trace<false, false, false, false, true><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
            else k_wf_trace<false, false><<<grid_persist, TRACE_BLOCK, 0, stream>>>(P, wf, b, q_in, active_counts + b, work_counters + 2 * b);
            if(b == 0 && first_sample_layers && i + 1 == spp)
            {
                MARK(TR_KCLASS_OTHER);
                k_wf_first_sample_layers<<<grid_finish, WF_BLOCK, 0, stream>>>(P, wf, base);
                ++*launches;
            }
            MARK(TR_KCLASS_SHADE);
            if(shadow_pending)
            {
                cudaError_t e2 = cudaStreamWaitEvent(stream, overlap->shadowed, 0);
                if(e2 != cudaSuccess) return e2;
                shadow_pending = false;
            }
            SLOT("shade");
            if(shade_spec) k_wf_shade<1><<<grid_shade, WF_BLOCK, 0, stream>>>(P, wf, b, i, q_in, active_counts + b, q_out, active_counts + b + 1,
                                                                             wf.shadow_queue, shadow_counts + b);
            else k_wf_shade<0><<<grid_shade, WF_BLOCK, 0, stream>>>(P, wf, b, i, q_in, active_counts + b, q_out, active_counts + b + 1,
                                                                   wf.shadow_queue, shadow_counts + b);
            *launches += 2;
            if(b + 1 < D && !fused)
            {
                MARK(TR_KCLASS_SHADOW);
                SLOT("shadow");
                // shadow(b) only touches the NEE job and the path's accumulators; closest(b + 1) only the ray and the hit
                // record: with a second stream they run side by side and shade(b + 1) waits for both
                cudaStream_t ss = stream;
                if(two_streams)
                {
                    cudaError_t e2 = cudaEventRecord(overlap->shaded, stream);
                    if(e2 == cudaSuccess) e2 = cudaStreamWaitEvent(overlap->aux, overlap->shaded, 0);
                    if(e2 != cudaSuccess) return e2;
                    ss = overlap->aux;
                }
                if(coop_shadow)
                {
                    if(P.stats) k_wf_trace_coop<true, true><<<grid_persist_shadow, WF_BLOCK, 0, ss>>>(P, wf, b, wf.shadow_queue, shadow_counts + b, work_counters + 2 * b + 1);
                    else k_wf_trace_coop<true, false><<<grid_persist_shadow, WF_BLOCK, 0, ss>>>(P, wf, b, wf.shadow_queue, shadow_counts + b, work_counters + 2 * b + 1);
                }
                else if(P.stats) k_wf_trace<true, true><<<grid_persist_shadow, TRACE_BLOCK, 0, ss>>>(P, wf, b, wf.shadow_queue, shadow_counts + b, work_counters + 2 * b + 1);
                else if(P.tune_tma & 2) k_wf_trace<true, false, true><<<grid_persist_shadow, TRACE_BLOCK, 0, ss>>>(P, wf, b, wf.shadow_queue, shadow_counts + b, work_counters + 2 * b + 1);
                else if(k_split & 2) k_wf_trace<true, false, false, false, true><<<grid_persist_shadow, TRACE_BLOCK, 0, ss>>>(P, wf, b, wf.shadow_queue, shadow_counts + b, work_counters + 2 * b + 1);
                else k_wf_trace<true, false><<<grid_persist_shadow, TRACE_BLOCK, 0, ss>>>(P, wf, b, wf.shadow_queue, shadow_counts + b, work_counters + 2 * b + 1);
                ++*launches;
                if(two_streams)
                {
                    cudaError_t e2 = cudaEventRecord(overlap->shadowed, overlap->aux);
                    if(e2 != cudaSuccess) return e2;
                    shadow_pending = true;
                }
            }
        }
        if(shadow_pending)      // max_ray_depth == 1 never gets here with a pending shadow pass; kept for safety
        {
            cudaError_t e2 = cudaStreamWaitEvent(stream, overlap->shadowed, 0);
            if(e2 != cudaSuccess) return e2;
        }
        if(phase == TR_PHASE_TRACE) break;
        MARK(TR_KCLASS_OTHER);
        SLOT("finish");
        k_wf_finish<<<grid_finish, WF_BLOCK, 0, stream>>>(P, wf, i, base, batch);
        ++*launches;
    }
    MARK(TR_KCLASS_END);
#undef MARK
    if(P.probe)
    {
        static unsigned long long got[136 * 2 * 16 + 2 * 128];
        cudaStreamSynchronize(stream);
        if(overlap && overlap->aux) cudaStreamSynchronize(overlap->aux);
        cudaMemcpy(got, P.probe, sizeof(got), cudaMemcpyDeviceToHost);
        for(uint32_t b = 0; b < D; ++b)
            for(int kind = 0; kind < 2; ++kind)
            {
                const unsigned long long* g = got + 136 * (2 * b + kind);
                if(g[3] == 0ull) continue;
                fprintf(stderr, "drain: bounce %u %s  run %.1f us  first-empty at %.1f  last-empty at %.1f  (tail %.1f us)\n", b, kind ? "shadow " : "closest",
                        (g[3] - g[0]) * 1e-3, (g[1] - g[0]) * 1e-3, (g[2] - g[0]) * 1e-3, (g[3] - g[1]) * 1e-3);
                fprintf(stderr, "   warps leaving per %d us:", k_probe_bin);
                for(int j = 0; j < 64; ++j) if(g[8 + j]) fprintf(stderr, " [%d]%llu", j * k_probe_bin, g[8 + j]);
                fprintf(stderr, "\n");
            }
        const unsigned long long* tl = got + 136 * 2 * 16;
        unsigned long long t0 = ~0ull;
        for(int k = 0; k < probe_launch; ++k) if(tl[2 * k] < t0) t0 = tl[2 * k];
        for(int k = 0; k < probe_launch; ++k)
            if(tl[2 * k + 1]) fprintf(stderr, "timeline: %-9s start %8.1f  end %8.1f  (%7.1f us)\n", probe_names[k], (tl[2 * k] - t0) * 1e-3, (tl[2 * k + 1] - t0) * 1e-3,
                                      (tl[2 * k + 1] - tl[2 * k]) * 1e-3);
    }
#undef SLOT
    return cudaGetLastError();
}

cudaError_t tr_render_pass_megakernel(cudaStream_t stream, const RenderParams& P, int sm_count, uint64_t* launches,
                                      tr_launch_timer* timer)
{
    const uint32_t total = P.frame.ray_count_x * P.frame.ray_count_y * P.frame.views;
    if(total == 0) return cudaSuccess;
    const int grid = grid_for(sm_count, 8, total);
    if(timer) timer->mark(stream, TR_KCLASS_OTHER);
    if(P.stats) k_megakernel<true><<<grid, WF_BLOCK, 0, stream>>>(P, total);
    else k_megakernel<false><<<grid, WF_BLOCK, 0, stream>>>(P, total);
    if(timer) timer->mark(stream, TR_KCLASS_END);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t tr_render_pass_direct(cudaStream_t stream, const RenderParams& P, int sm_count, uint64_t* launches,
                                  tr_launch_timer* timer)
{
    const uint32_t total = P.frame.ray_count_x * P.frame.ray_count_y * P.frame.views;
    if(total == 0) return cudaSuccess;
    const int grid = grid_for(sm_count, 8, total);
    if(timer) timer->mark(stream, TR_KCLASS_OTHER);
    if(P.stats) k_direct<true><<<grid, WF_BLOCK, 0, stream>>>(P, total);
    else k_direct<false><<<grid, WF_BLOCK, 0, stream>>>(P, total);
    if(timer) timer->mark(stream, TR_KCLASS_END);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t tr_render_stitch_aux_layers(cudaStream_t stream, const RenderParams& P, const void* const src[TR_GBUFFER_LAYER_COUNT],
                                        const int src_full[TR_GBUFFER_LAYER_COUNT], int sm_count)
{
    const uint32_t base = P.frame.ray_count_x * P.frame.ray_count_y * P.frame.views;
    if(base == 0 || P.peer_count == 0) return cudaSuccess;
    AuxStitchSources S;
    bool any = false;
    for(int l = 0; l < TR_GBUFFER_LAYER_COUNT; ++l)
    {
        S.src[l] = l == TR_GBUFFER_COLOR ? nullptr : src[l];
        S.src_full[l] = src_full[l];
        any = any || S.src[l];
    }
    if(!any) return cudaSuccess;
    k_stitch_aux_layers<<<grid_for(sm_count, 16, base), WF_BLOCK, 0, stream>>>(P, S, base);
    return cudaGetLastError();
}

cudaError_t tr_render_feature(cudaStream_t stream, const RenderParams& P, int feature, const float dv[4])
{
    const uint32_t total = P.frame.ray_count_x * P.frame.ray_count_y * P.frame.views;
    if(total == 0) return cudaSuccess;
    k_feature<<<(total + WF_BLOCK - 1) / WF_BLOCK, WF_BLOCK, 0, stream>>>(P, feature, make_float4(dv[0], dv[1], dv[2], dv[3]), total);
    return cudaGetLastError();
}

cudaError_t tr_render_trace_rays(cudaStream_t stream, const RenderParams& P, const float* rays_dev, uint32_t n, void* hits_dev)
{
    if(n == 0) return cudaSuccess;
    k_trace_rays<<<(n + 127) / 128, 128, 0, stream>>>(P, rays_dev, n, (OutHit*)hits_dev);
    return cudaGetLastError();
}

cudaError_t tr_render_sample_test(cudaStream_t stream, const RenderParams& P, const uint32_t* coords_dev, uint32_t n, int draw_index, uint32_t* out_dev)
{
    if(n == 0) return cudaSuccess;
    k_sample_test<<<(n + 127) / 128, 128, 0, stream>>>(P, coords_dev, n, draw_index, out_dev);
    return cudaGetLastError();
}

// stitch_scanline.comp:21-50 for one source device: row ly of device `index` of `count` is row ly * count + index.
__global__ void k_unscanline(const FrameParams f, const float4* __restrict__ rows, uint32_t index, uint32_t count, uint32_t row_count,
                             float4* __restrict__ full, float blend_ratio)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t view = blockIdx.y;
    if(i >= row_count * f.size_x) return;
    const uint32_t ly = i / f.size_x, x = i - ly * f.size_x;
    float4 o = rows[(size_t)view * row_count * f.size_x + i];
    float4* dst = full + (size_t)view * f.size_x * f.size_y + (size_t)(ly * count + index) * f.size_x + x;
    if(blend_ratio < 1.0f)
    {
        const float4 q = *dst;
        const vec4 m = mix_(V4(q.x, q.y, q.z, q.w), V4(o.x, o.y, o.z, o.w), blend_ratio);
        o = make_float4(m.x, m.y, m.z, m.w);
    }
    *dst = o;
}

cudaError_t tr_render_unscanline(cudaStream_t stream, const FrameParams& f, const float4* rows, uint32_t index, uint32_t count, float4* full_image,
                                 float blend_ratio)
{
    if(count == 0 || index >= count || f.views == 0 || f.size_y <= index) return cudaSuccess;
    const uint32_t row_count = (f.size_y - index + count - 1) / count;
    dim3 grid((row_count * f.size_x + 255) / 256, f.views);
    k_unscanline<<<grid, 256, 0, stream>>>(f, rows, index, count, row_count, full_image, blend_ratio);
    return cudaGetLastError();
}

cudaError_t tr_render_unshuffle(cudaStream_t stream, const FrameParams& f, const float4* strip, uint32_t index, uint32_t count, float4* full_image,
                                float blend_ratio)
{
    if(count == 0 || f.views == 0) return cudaSuccess;
    dim3 grid((count + 255) / 256, f.views);
    k_unshuffle<<<grid, 256, 0, stream>>>(f, strip, index, count, full_image, blend_ratio);
    return cudaGetLastError();
}
