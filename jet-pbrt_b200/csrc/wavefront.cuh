// wavefront.cuh -- the five wavefront stages of the B200 path tracer (sm_100a).
//
//   generate  : FSampler::GetCameraSample + FCamera::GenerateRay      (integrator.cc:100-101, camera.h:52-58)
//   extend    : FScene::Intersect, closest hit                         (integrator.cc:327, scene.cc:25-33)
//   shade     : Le / env, BSDF build, NEE light sampling + BSDF eval,  (integrator.cc:328-399)
//               BSDF sampling, Russian roulette, throughput update
//   connect   : FScene::Occluded for the NEE shadow rays + contribution (integrator.cc:367-370)
//   accumulate: radiance sums -> film; finalize = Clamp01(mean)         (integrator.cc:102-108, film.h:64-68)
//
// A path is a 48-byte record (3 x float4, SoA) that is rewritten, compacted, once per bounce:
//   o = (origin.xyz, pixel)   d = (dir.xyz, sample | bounce << 24 | specular << 31)   b = (beta.rgb, -)
// Queues ARE the records: survivors of `shade` are written contiguously into the other half of a
// ping-pong buffer at positions handed out by one warp-aggregated atomicAdd per warp
// (__ballot_sync + __popc), so every stage reads and writes fully coalesced 16-byte lanes.
// Shadow rays go to static 48-byte slots (origin+tmax, dir, contribution+pixel), one per (vertex, light).
// All radiance goes straight to the float film with RED.ADD.F32 (no per-path radiance state).
//
// The same stages serve the reference's other integrators as MODES of the pipeline:
//   FPathIntegratorRecursive (integrator.cc:233-307): the same estimator and draws -> the same kernels;
//   FWhittedIntegrator (integrator.cc:115-220): k_logic<true> / k_shade<KIND, true> -- emission at every vertex,
//     NEE at every non-delta vertex, continuation through SPECULAR lobes only, a mirror spawning TWO rays
//     (bsdf.h:282 subset match; the path record's b.w carries the vertex's number in the ray tree);
//   FDebugIntegrator (integrator.h:44-58): generate -> extend -> k_debug (hit normal as colour).
//
// Kernels are persistent: grid = SMs x resident blocks, each warp pulls batches of 32 queue items
// from a device-side work counter, and queue lengths are read from device memory -- the host
// never reads a count back, so a whole pass is one uninterrupted stream of launches.
#pragma once

#include <cuda_runtime.h>

#include "bsdf.cuh"
#include "dev_scene.h"
#include "dmath.cuh"
#include "intersect.cuh"
#include "light.cuh"
#include "rng.cuh"

namespace jpbrt {

// per-iteration device counters: queue lengths and the work cursors of the persistent kernels
enum { CNT_RAYS = 0, CNT_UNUSED = 1, CNT_W_EXTEND = 2, CNT_W_SHADE = 3, CNT_W_CONNECT = 4,
       CNT_Q0 = 5 /* 4 kinds */, CNT_WQ0 = 9 /* 4 kinds */, CNT_KINDS = 13 };
enum {
    ST_SAMPLES = 0, ST_EXT_RAYS, ST_SHADOW_RAYS, ST_VERTICES, ST_BOX, ST_PRIM, ST_SH_BOX, ST_SH_PRIM, ST_INVALID, ST_DROPPED,
    ST_STACK_DROPPED,  // far children lost to a full traversal stack (intersect.cuh: kTraversalStack)
    ST_NEE_DROPPED,    // light samples that found no shadow slot (pool smaller than vertices x lights)
    ST_NODE_FETCH, ST_PRIM_FETCH, ST_SH_NODE_FETCH, ST_SH_PRIM_FETCH,  // COUNT variants: distinct records per warp step
    ST_COUNT
};

// The only values that change from one wavefront to the next.  They live in device memory (written by
// k_set_args) so that the launch sequence of a wavefront is IDENTICAL every time and can be replayed as one
// CUDA graph: a step is then a single host call instead of ~90 launches (bench.py measured the host taking
// 10-70 ms to queue a step when the box's CPUs are busy).
struct PassArgs {
    uint32_t k0, k1;   // sampler key (seed)
    int sample_begin;  // first sample index of this wavefront
    int n_paths;       // pixels x samples of this wavefront
    int pixel_begin;   // the wavefront covers pixel_order[pixel_begin .. pixel_begin + pixel_count): a BAND of the frame
    int pixel_count;   //   whose film (12 bytes per pixel) stays L2-resident while its samples accumulate
};

__global__ void k_set_args(PassArgs* dst, PassArgs v) { *dst = v; }

struct WfParams {
    DevScene sc;
    float4* ray_o[2];
    float4* ray_d[2];
    float4* ray_b[2];
    float2* hit;
    float4* sh_o;
    float4* sh_d;
    float4* sh_c;
    int* kind_queue;     // [NUM_KINDS][queue_capacity] indices into the current ray buffer
    int queue_capacity;
    int* counters;  // [CNT_KINDS][counter_stride]
    int counter_stride;
    float* film;
    unsigned long long* stats;
    const PassArgs* args;
    int npix;
    int blocks_per_bounce;
    int shadow_capacity;
    int refill_min;  // ray replacement threshold of the traversal kernels (idle lanes per warp)
    int min_inner;   // the node phase of a warp ends when fewer lanes than this are still at inner nodes (intersect.cuh)
    // Ray reordering (option "sort_rays"): the rays a bounce emits are binned by (origin cell, direction octant) while they
    // are appended, and k_extend of the next bounce walks them in bin order through `perm` (see k_permute).
    unsigned* sort_bins;   // [1 << sort_key_bits] per-bin counters; null = reordering off
    unsigned* sort_start;  // exclusive prefix sums of the counters
    uint2* sort_kr;        // (bin, rank within the bin) of every queued ray
    int* perm;             // sorted position -> queue index
    int sort_cell_bits;    // origin grid: 2^bits cells per axis
    int sort_dir;          // 1: the direction octant is the low 3 bits of the key
    float sort_min[3], sort_scale[3];
};

constexpr int kBlock = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Wavefront state (path / hit / shadow records, kind queues) is written once and read once or twice, GBs per pass:
// it is loaded and stored with the streaming (evict-first) cache policy so that it does not push the scene's
// nodes and primitives -- and the film band -- out of L1/L2.  JPB_STREAM_HINTS=0 builds the plain-policy variant.
#ifndef JPB_STREAM_HINTS
#define JPB_STREAM_HINTS 1
#endif
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
#if JPB_STREAM_HINTS
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, const T& v) {
#if JPB_STREAM_HINTS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// accumulate: film[pixel] += c   (radiance sums; FFilm::AddColor happens at finalize)
__device__ __forceinline__ void film_add(const WfParams& p, int pixel, const f3& c) {
    if (!(isfinite(c.x) && isfinite(c.y) && isfinite(c.z))) {  // the reference only logs these (integrator.cc:104)
        atomicAdd(p.stats + ST_INVALID, 1ull);
        return;
    }
    float* px = p.film + 3 * (size_t)pixel;
    if (c.x != 0.f) atomicAdd(px + 0, c.x);
    if (c.y != 0.f) atomicAdd(px + 1, c.y);
    if (c.z != 0.f) atomicAdd(px + 2, c.z);
}

__device__ __forceinline__ void warp_stat_add(unsigned long long* dst, unsigned v) {
    v = __reduce_add_sync(kFull, v);
    if (lane_id() == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

// ---------------------------------------------------------------------------------------------
// generate
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_generate(const __grid_constant__ WfParams p) {
    const DevCamera& cam = p.sc.cam;
    const PassArgs a = *p.args;
    const RngKey key{a.k0, a.k1};
    const int n_paths = a.n_paths;
    const f3 pos = mk3(cam.pos[0], cam.pos[1], cam.pos[2]);
    const f3 front = mk3(cam.front[0], cam.front[1], cam.front[2]);
    const f3 right = mk3(cam.right[0], cam.right[1], cam.right[2]);
    const f3 up = mk3(cam.up[0], cam.up[1], cam.up[2]);
    const int W = p.sc.width;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_paths; i += gridDim.x * blockDim.x) {
        const int pixel = __ldg(p.sc.pixel_order + a.pixel_begin + i % a.pixel_count);
        const int sample = a.sample_begin + i / a.pixel_count;
        const int x = pixel % W, y = pixel / W;
        const float4 u = rng_block(key, (uint32_t)pixel, (uint32_t)sample, 0u);
        const float fx = (float)x + u.x, fy = (float)y + u.y;  // sampler.h:152
        const f3 dir = front + right * (fx / cam.res_x - 0.5f) + up * (0.5f - fy / cam.res_y);  // camera.h:54-55
        const f3 d = normalize(dir);
        st_stream(&p.ray_o[0][i], make_float4(pos.x, pos.y, pos.z, __int_as_float(pixel)));
        st_stream(&p.ray_d[0][i], make_float4(d.x, d.y, d.z, __int_as_float(sample & 0xffffff)));
        st_stream(&p.ray_b[0][i], make_float4(1.f, 1.f, 1.f, 0.f));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.counters[CNT_RAYS * p.counter_stride + 0] = n_paths;
        atomicAdd(p.stats + ST_SAMPLES, (unsigned long long)n_paths);
    }
}

// ---------------------------------------------------------------------------------------------
// extend: closest hit for every ray of iteration `it`
// ---------------------------------------------------------------------------------------------
template <bool PERM>
struct ExtendIO {
    const float4* ro;
    const float4* rd;
    float2* hit;
    const int* perm;  // PERM: queue position -> ray index (rays are walked in bin order, results land at the ray's own index)
    __device__ __forceinline__ bool load(int j, f3& o, f3& d, float& tmin, float& tmax) const {
        const int i = PERM ? __ldg(&perm[j]) : j;
        o = mk3(ld_stream(&ro[i]));
        d = mk3(ld_stream(&rd[i]));
        tmin = JPBRT_RAY_TMIN;                 // FRay default min_t, geometry.h:395
        tmax = __int_as_float(0x7f800000);     // kInfinity
        return true;
    }
    __device__ __forceinline__ void store(int j, int slot, float t) const {
        const int i = PERM ? __ldg(&perm[j]) : j;
        st_stream(&hit[i], make_float2(t, __int_as_float(slot)));
    }
};

// MINB = resident blocks per SM the register allocation must allow (5: 48 registers, 6: 40; measured on B200:
// 6 is 1-3 % faster on all scenes, 8 = 32 registers spills and is 15-25 % slower).  The 6-block kernels also drop the
// full-stack test of every push (intersect.cuh: GUARD) and are launched only for trees of verified depth (c_api.cu: fast6).
template <bool COUNT, int MINB, bool PERM = false, bool QN = false>
__global__ void __launch_bounds__(kBlock, MINB) k_extend(const __grid_constant__ WfParams p, int it) {
    const int n = min(p.counters[CNT_RAYS * p.counter_stride + it], p.queue_capacity);
    int* work = p.counters + CNT_W_EXTEND * p.counter_stride + it;
    const int buf = it & 1;
    TravCounts cnt;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(p.stats + ST_EXT_RAYS, (unsigned long long)n);
    ExtendIO<PERM> io{p.ray_o[buf], p.ray_d[buf], p.hit, p.perm};
    traverse_queue<false, COUNT, COUNT || MINB < 6, QN>(p.sc, n, work, io, p.refill_min, p.min_inner, cnt, p.stats + ST_STACK_DROPPED);
    if (COUNT) {
        warp_stat_add(p.stats + ST_BOX, cnt.box);
        warp_stat_add(p.stats + ST_PRIM, cnt.prim);
        warp_stat_add(p.stats + ST_NODE_FETCH, cnt.node_fetch);
        warp_stat_add(p.stats + ST_PRIM_FETCH, cnt.prim_fetch);
    }
}

// ---------------------------------------------------------------------------------------------
// connect: any-hit for every shadow ray of iteration `it`; unoccluded -> film += contribution
// ---------------------------------------------------------------------------------------------
// Shadow rays live in STATIC slots: vertex v of the iteration (v = its position in the concatenated
// lambert | conductor | dielectric kind queues) and non-black light k own slot k * n_vertices + v, so
// k_shade needs no atomics to emit them and the slots of one light are contiguous (coalesced writes,
// and coherent rays for k_connect).  A slot whose sample cannot contribute holds tmax < 0.
// Both halves of a shadow record are fetched together (round 1 fetched the direction only after testing o.w), and the
// contribution record carries its pixel, so that delivery is one load.  Measured on B200 (profiles/ab/
// r02_ab_connect_io_shade_prefetch.log): dependent vs parallel fetch, an L2 prefetch of the contribution when the ray is
// taken, and a cp.async copy of it into shared memory (no global load at delivery) are all within +-0.3 % on k_connect --
// the memory latency of the ray records is hidden by the other 47 warps of the SM; the simplest form stays.
struct ConnectIO {
    const WfParams* p;
    unsigned* n_traced;
    __device__ __forceinline__ bool load(int i, f3& o, f3& d, float& tmin, float& tmax) const {
        const float4 so = ld_stream(&p->sh_o[i]);
        const float4 sd = ld_stream(&p->sh_d[i]);
        if (so.w < 0.f) return false;
        o = mk3(so);
        d = mk3(sd);
        tmin = JPBRT_RAY_TMIN;  // scene.h:38
        tmax = so.w;            // dist - 0.001
        ++*n_traced;
        return true;
    }
    __device__ __forceinline__ void store(int i, int slot, float) const {
        if (slot < 0) {  // unoccluded: integrator.cc:367-370
            const float4 c = ld_stream(&p->sh_c[i]);
            film_add(*p, __float_as_int(c.w), mk3(c));
        }
    }
};

__device__ __forceinline__ int nee_vertex_count(const WfParams& p, int it) {
    return p.counters[(CNT_Q0 + 0) * p.counter_stride + it] + p.counters[(CNT_Q0 + 1) * p.counter_stride + it] +
           p.counters[(CNT_Q0 + 2) * p.counter_stride + it];
}

template <bool COUNT, int MINB, bool QN = false>
__global__ void __launch_bounds__(kBlock, MINB) k_connect(const __grid_constant__ WfParams p, int it) {
    const long long slots = (long long)nee_vertex_count(p, it) * p.sc.n_nee_lights;
    const int n = (int)(slots < p.shadow_capacity ? slots : p.shadow_capacity);
    int* work = p.counters + CNT_W_CONNECT * p.counter_stride + it;
    TravCounts cnt;
    unsigned traced = 0;
    ConnectIO io{&p, &traced};
    traverse_queue<true, COUNT, COUNT || MINB < 6, QN>(p.sc, n, work, io, p.refill_min, p.min_inner, cnt, p.stats + ST_STACK_DROPPED);
    warp_stat_add(p.stats + ST_SHADOW_RAYS, traced);
    if (COUNT) {
        warp_stat_add(p.stats + ST_SH_BOX, cnt.box);
        warp_stat_add(p.stats + ST_SH_PRIM, cnt.prim);
        warp_stat_add(p.stats + ST_SH_NODE_FETCH, cnt.node_fetch);
        warp_stat_add(p.stats + ST_SH_PRIM_FETCH, cnt.prim_fetch);
    }
}

// ---------------------------------------------------------------------------------------------
// shade = logic + one material kernel per BSDF kind
//
// A single shade kernel for every material is a 7,000-instruction megakernel whose warps run ~6 of 32
// lanes and stall on instruction fetch (profiles/r01_ncu_bunny_v0.md).  Instead:
//   k_logic       : emission / environment, depth termination, null-material pass-through, and
//                   CLASSIFICATION of every surviving vertex by the BSDF its material builds
//                   (the plastic lobe pick included) into one index queue per kind;
//   k_shade<KIND> : NEE + BSDF sampling + roulette for ONE kind: every lane of a warp runs the same
//                   BSDF code, and each kernel's code is a fraction of the megakernel's.
// ---------------------------------------------------------------------------------------------
enum { KIND_LAMBERT = 0, KIND_MF_CONDUCTOR = 1, KIND_MF_DIELECTRIC = 2, KIND_DELTA = 3, NUM_KINDS = 4 };

__device__ __forceinline__ int bsdf_kind_of(int k) {
    return k == K_LAMBERT ? KIND_LAMBERT : k == K_MICROFACET_CONDUCTOR ? KIND_MF_CONDUCTOR : k == K_MICROFACET_DIELECTRIC ? KIND_MF_DIELECTRIC : KIND_DELTA;
}

// Reordering key of a ray: Morton code of its origin's cell in a 2^bits-per-axis grid over the scene's bounds, then
// (optionally) the octant of its direction.  Rays of one bin start in the same region and head the same way: a warp of
// them walks the same nodes for longer than 32 rays in emission order do.
__device__ __forceinline__ unsigned spread_bits3(unsigned v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ unsigned ray_sort_key(const WfParams& p, const float4& o, const float4& d) {
    const float hi = (float)((1 << p.sort_cell_bits) - 1);
    const unsigned cx = (unsigned)fminf(fmaxf((o.x - p.sort_min[0]) * p.sort_scale[0], 0.f), hi);
    const unsigned cy = (unsigned)fminf(fmaxf((o.y - p.sort_min[1]) * p.sort_scale[1], 0.f), hi);
    const unsigned cz = (unsigned)fminf(fmaxf((o.z - p.sort_min[2]) * p.sort_scale[2], 0.f), hi);
    unsigned key = spread_bits3(cx) | (spread_bits3(cy) << 1) | (spread_bits3(cz) << 2);
    if (p.sort_dir) key = (key << 3) | (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
    return key;
}

// Warp-aggregated append of the survivors' records to the next iteration's ray queue.
// GUARD: the Whitted mode's ray tree can outgrow the pool (a mirror spawns two rays); rays past the end of the
// queue are dropped and counted (stats.invalid_contributions) -- consumers clamp the queue length to the capacity.
template <bool GUARD = false>
__device__ __forceinline__ void append_next(const WfParams& p, int* next_count, int nbuf, bool alive, const float4& no, const float4& nd,
                                            const float4& nbeta) {
    const unsigned mask = __ballot_sync(kFull, alive);
    if (mask) {
        int wbase = 0;
        if (lane_id() == 0) wbase = atomicAdd(next_count, __popc(mask));
        wbase = __shfl_sync(kFull, wbase, 0);
        if (alive) {
            const int dst = wbase + __popc(mask & ((1u << lane_id()) - 1));
            if (GUARD && dst >= p.queue_capacity) {
                atomicAdd(p.stats + ST_DROPPED, 1ull);
                return;
            }
            st_stream(&p.ray_o[nbuf][dst], no);
            st_stream(&p.ray_d[nbuf][dst], nd);
            st_stream(&p.ray_b[nbuf][dst], nbeta);
            if (!GUARD && p.sort_bins) {  // reordering: count the ray into its bin; the count it found is its rank there
                const unsigned key = ray_sort_key(p, no, nd);
                st_stream(&p.sort_kr[dst], make_uint2(key, atomicAdd(p.sort_bins + key, 1u)));
            }
        }
    }
}

// Reordering, second half: after the bins' counters were prefix-summed, ray i belongs at sorted position
// start[bin] + rank.  k_extend then reads perm[] in order and gathers the rays.
__global__ void __launch_bounds__(kBlock) k_permute(const __grid_constant__ WfParams p, int it) {
    const int n = min(p.counters[CNT_RAYS * p.counter_stride + it], p.queue_capacity);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 kr = ld_stream(&p.sort_kr[i]);
        p.perm[__ldg(p.sort_start + kr.x) + kr.y] = i;
    }
}

// (measured on B200, profiles/ab/r01_ab_chunks.log: 2 -> 4 -> 8 chunks per fetch = shade stage 6.28 -> 5.73 -> 5.50 ms on Cornell,
// 3.36 -> 3.02 -> 2.84 ms on the bunny scene: the same-address atomics of the cursors are a real cost there)
#ifndef JPB_CHUNKS_PER_FETCH
#define JPB_CHUNKS_PER_FETCH 8
#endif
#ifndef JPB_GUIDED_FETCH
#define JPB_GUIDED_FETCH 64  // 0: the fetch size is fixed per launch (round 1); N: chunks = left / (warps * N), re-evaluated at every fetch
#endif
constexpr int kChunksPerFetch = JPB_CHUNKS_PER_FETCH;  // work is fetched up to 256 items at a time: fewer atomics on the hot queue cursors

// Items per fetch: 256 while much of the queue is left, shrinking to 32 as it drains (guided self-scheduling: `left` is what
// remained when this warp last looked at the cursor), so that the big fetches that keep the cursor's atomics cheap are not
// what the last warps of a launch are still working through -- a 256-vertex fetch is 70-100 us of shading (16 lights: ~1 ms),
// and every launch used to end with the chip waiting for a few of them.  Late, small iterations (left small from the
// start) spread over the whole chip the same way.
// Measured on B200 (profiles/ab/r02_ab_guided_fetch.log, shade stage, fixed -> guided): bunny 7.48 -> 7.37 ms, glossy (16 lights
// per vertex) 23.25 -> 22.41 ms, and 21.74 ms when the fetches shrink twice as early (`heavy`: scenes with >= 8 lights to sample).
__device__ __forceinline__ int chunks_for(int left, bool heavy = false) {
    const int total_warps = (gridDim.x * blockDim.x) >> 5;
    const int per_warp = JPB_GUIDED_FETCH ? (heavy ? 2 * JPB_GUIDED_FETCH : JPB_GUIDED_FETCH) : 64;
    return max(1, min(kChunksPerFetch, left / (total_warps * per_warp)));
}

__device__ __forceinline__ int warp_fetch_n(int* counter, int amount) {
    int base = 0;
    if (lane_id() == 0) base = atomicAdd(counter, amount);
    return __shfl_sync(kFull, base, 0);
}

// (k_logic waits on memory -- ncu: IPC 1.6, long-scoreboard stalls -- yet prefetching all chunks of a fetch before the first is
// looked at measured +-0, as did 3 or 6 resident blocks: profiles/ab/r02_ab_logic.log)
#ifndef JPB_LOGIC_MIN_BLOCKS
#define JPB_LOGIC_MIN_BLOCKS 4  // 64 registers; 5-6 blocks (48 / 40 registers) measured equal (profiles/ab/r01_ab_shade.log)
#endif
template <bool WHITTED>
__global__ void __launch_bounds__(kBlock, JPB_LOGIC_MIN_BLOCKS) k_logic(const __grid_constant__ WfParams p, int it) {
    const DevScene& sc = p.sc;
    const int n = min(p.counters[CNT_RAYS * p.counter_stride + it], p.queue_capacity);
    int* work = p.counters + CNT_W_SHADE * p.counter_stride + it;
    int* next_count = p.counters + CNT_RAYS * p.counter_stride + it + 1;
    const int buf = it & 1, nbuf = buf ^ 1;
    int nchunks = chunks_for(n);
    const RngKey key{p.args->k0, p.args->k1};
    for (;;) {
        const int base = warp_fetch_n(work, 32 * nchunks);
        if (base >= n) break;
        const int fetched = nchunks;
        if (JPB_GUIDED_FETCH) nchunks = chunks_for(n - base);  // the NEXT fetch of this warp
        int kinds[kChunksPerFetch];
        unsigned masks[kChunksPerFetch][NUM_KINDS];
#pragma unroll
        for (int c = 0; c < kChunksPerFetch; ++c) {
            const int i = base + 32 * c + lane_id();
            bool alive = false;
            int kind = -1;
            float4 no = make_float4(0, 0, 0, 0), nd = no, nbeta = no;
            if (c < fetched && i < n) {
                const float4 rd = ld_stream(&p.ray_d[buf][i]);
                const float2 h = ld_stream(&p.hit[i]);
                const int fl = __float_as_int(rd.w);
                const int bounce = (fl >> 24) & 0x7f;
                const bool specular = fl < 0;
                const int slot = __float_as_int(h.y);
                const bool add_emission = WHITTED || (bounce == 0) || specular;  // integrator.cc:328; Whitted: always (:124,141)
                const bool in_depth = WHITTED || bounce < sc.max_depth;           // Whitted bounds the depth where it spawns (:157)
                if (slot < 0) {
                    if (add_emission && sc.n_inf_lights > 0) {
                        const int pixel = __float_as_int(ld_stream(&p.ray_o[buf][i]).w);
                        const f3 beta = mk3(ld_stream(&p.ray_b[buf][i]));
                        for (int k = 0; k < sc.n_inf_lights; ++k)  // integrator.cc:334-335
                            film_add(p, pixel, cmul(beta, mk3(ldg4(sc.lights + (size_t)sc.inf_lights[k] * kLightStride))));
                    }
                } else {
                    const int2 ml = __ldg(reinterpret_cast<const int2*>(sc.slot_ml) + slot);
                    const bool emits = add_emission && ml.y >= 0 && !(WHITTED && ml.x < 0);  // Whitted: null material returns first (:136)
                    const bool pass = in_depth && ml.x < 0;
                    if (emits || pass) {
                        const float4 ro = ld_stream(&p.ray_o[buf][i]);
                        const float4 rb = ld_stream(&p.ray_b[buf][i]);
                        const f3 o = mk3(ro), d = mk3(rd);
                        const f3 P = o + h.x * d;  // FRay::operator(), geometry.h:413-417
                        if (emits) {
                            const f3 N = hit_normal(sc, slot, P, d);
                            const f3 Le = emitted(sc, ml.y, N, -d);
                            if (!is_black(Le)) film_add(p, __float_as_int(ro.w), cmul(mk3(rb), Le));  // integrator.cc:331
                        }
                        if (pass) {
                            // null material: the ray continues unchanged, the bounce is not counted (integrator.cc:349-353)
                            alive = true;
                            no = make_float4(P.x, P.y, P.z, ro.w);
                            nd = rd;
                            nbeta = rb;
                        }
                    }
                    if (in_depth && ml.x >= 0) {  // integrator.cc:340,348
                        const Float4* mat = sc.materials + (size_t)ml.x * kMaterialStride;
                        const int type = __float_as_int(ldg4(mat).w);
                        if (type == MAT_PLASTIC) {  // the lobe pick is the first number of the bounce's block (material.cc:14)
                            const int pixel = __float_as_int(ld_stream(&p.ray_o[buf][i]).w);
                            const uint32_t blk = 1u + (uint32_t)bounce * (uint32_t)p.blocks_per_bounce;
                            const uint32_t node = WHITTED ? __float_as_uint(ld_stream(&p.ray_b[buf][i]).w) : 0u;
                            const float4 u0 = rng_block(key, (uint32_t)pixel, (uint32_t)(fl & 0xffffff), blk, node);
                            kind = (u0.x < ldg4(mat + 2).y) ? KIND_LAMBERT : KIND_MF_DIELECTRIC;
                        } else {
                            kind = type == MAT_MATTE ? KIND_LAMBERT : type == MAT_METAL ? KIND_MF_CONDUCTOR : KIND_DELTA;
                        }
                    }
                }
            }
            append_next<WHITTED>(p, next_count, nbuf, alive, no, nd, nbeta);
            kinds[c] = kind;
#pragma unroll
            for (int k = 0; k < NUM_KINDS; ++k) masks[c][k] = __ballot_sync(kFull, kind == k);
        }
        // one atomicAdd per kind for all 128 items; the four are issued back to back
        int totals[NUM_KINDS], bases[NUM_KINDS];
#pragma unroll
        for (int k = 0; k < NUM_KINDS; ++k) {
            totals[k] = 0;
#pragma unroll
            for (int c = 0; c < kChunksPerFetch; ++c) totals[k] += __popc(masks[c][k]);
            bases[k] = 0;
            if (lane_id() == 0 && totals[k]) bases[k] = atomicAdd(p.counters + (CNT_Q0 + k) * p.counter_stride + it, totals[k]);
        }
#pragma unroll
        for (int k = 0; k < NUM_KINDS; ++k) bases[k] = __shfl_sync(kFull, bases[k], 0);
#pragma unroll
        for (int c = 0; c < kChunksPerFetch; ++c) {
#pragma unroll
            for (int k = 0; k < NUM_KINDS; ++k) {
                if (kinds[c] == k)
                    st_stream(&p.kind_queue[(size_t)k * p.queue_capacity + bases[k] + __popc(masks[c][k] & ((1u << lane_id()) - 1))], base + 32 * c + lane_id());
                bases[k] += __popc(masks[c][k]);
            }
        }
    }
}

// Resident blocks per SM.  Round 1: 2 (111 registers) -> 3 (80) measured 5-9 % faster, 3 -> 4 (64 registers) another 3 %
// (profiles/ab/r01_ab_shade.log).  Round 2, with the next chunk's records prefetched and the light sample's direction computed
// once, latency is less of the story and the 64-register build's spills more: 3 blocks (80 registers, a third of the spill
// traffic) is 1-4 % faster than 4 on all three scenes, 5 blocks (48 registers) 5-7 % slower; two lights per trip of the NEE loop
// (ILP across lights) is +-3 % either way (profiles/ab/r02_ab_shade_ilp_blocks.log).
#ifndef JPB_SHADE_MIN_BLOCKS
#define JPB_SHADE_MIN_BLOCKS 3
#endif
#ifndef JPB_LIGHT_UNROLL
#define JPB_LIGHT_UNROLL 1
#endif
constexpr int kLightUnroll = JPB_LIGHT_UNROLL;  // lights per trip of k_shade's NEE loop (A/B builds)
// The next chunk's path records are prefetched to L2 while the current chunk is shaded: shade stage -3.8 % (bunny scene),
// -3.5 % (Cornell), -0.6 % (glossy); profiles/ab/r02_ab_connect_io_shade_prefetch.log
#ifndef JPB_SHADE_PREFETCH
#define JPB_SHADE_PREFETCH 1
#endif
template <int KIND, bool WHITTED = false>
__global__ void __launch_bounds__(kBlock, JPB_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ WfParams p, int it) {
    const DevScene& sc = p.sc;
    const int n = p.counters[(CNT_Q0 + KIND) * p.counter_stride + it];
    int* work = p.counters + (CNT_WQ0 + KIND) * p.counter_stride + it;
    int* next_count = p.counters + CNT_RAYS * p.counter_stride + it + 1;
    const int* __restrict__ queue = p.kind_queue + (size_t)KIND * p.queue_capacity;
    const int buf = it & 1, nbuf = buf ^ 1;
    // static shadow slots (see ConnectIO): this kind's vertices start at `kind_base` of the iteration's vertex list
    const int n_vertices = nee_vertex_count(p, it);
    int kind_base = 0;
#pragma unroll
    for (int k = 0; k < NUM_KINDS; ++k)
        if (k < KIND) kind_base += p.counters[(CNT_Q0 + k) * p.counter_stride + it];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(p.stats + ST_VERTICES, (unsigned long long)n);
    const bool heavy = KIND != KIND_DELTA && sc.n_nee_lights >= 8;
    int next_chunks = chunks_for(n, heavy);
    const RngKey key{p.args->k0, p.args->k1};
    for (;;) {
        const int nchunks = next_chunks;
        const int fetch_base = warp_fetch_n(work, 32 * nchunks);
        if (fetch_base >= n) break;
        if (JPB_GUIDED_FETCH) next_chunks = chunks_for(n - fetch_base, heavy);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            if (fetch_base + 32 * c >= n) break;
            const int qi = fetch_base + 32 * c + lane_id();
#if JPB_SHADE_PREFETCH
            // the NEXT chunk's records on their way to L2 while this chunk is shaded (queue -> index -> four records is a chain of
            // dependent DRAM round trips otherwise)
            if (c + 1 < nchunks && qi + 32 < n) {
                const int in = ld_stream(&queue[qi + 32]);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.ray_o[buf][in]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.ray_d[buf][in]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.ray_b[buf][in]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(&p.hit[in]));
            }
#endif
            bool alive = false, alive2 = false;  // alive2: a mirror's second ray in the Whitted mode
            // The continuation ray's record is DEFINED only where the path survives (and zeroed, late, where it does not): a
            // zero-initialisation up here kept 16 words alive across the whole light loop -- in local memory at 64 registers
            // (ncu source view, round 2: 16 STL + 16 LDL per vertex for values nobody reads).
            float4 no, nd, nbeta, nbeta2;
            if (qi < n) {
                const int i = ld_stream(&queue[qi]);
                const float4 ro = ld_stream(&p.ray_o[buf][i]);
                const float4 rd = ld_stream(&p.ray_d[buf][i]);
                const float4 rb = ld_stream(&p.ray_b[buf][i]);
                const float2 h = ld_stream(&p.hit[i]);
                const f3 o = mk3(ro), d = mk3(rd);
                f3 beta = mk3(rb);
                const int pixel = __float_as_int(ro.w);
                const int fl = __float_as_int(rd.w);
                const int sample = fl & 0xffffff;
                const int bounce = (fl >> 24) & 0x7f;
                const int slot = __float_as_int(h.y);
                const f3 P = o + h.x * d;  // FRay::operator(), geometry.h:413-417
                const f3 N = hit_normal(sc, slot, P, d);
                const f3 wo = -d;
                const int2 ml = __ldg(reinterpret_cast<const int2*>(sc.slot_ml) + slot);
                const uint32_t blk = 1u + (uint32_t)bounce * (uint32_t)p.blocks_per_bounce;
                const uint32_t node = WHITTED ? __float_as_uint(rb.w) : 0u;
                const float4 u0 = rng_block(key, (uint32_t)pixel, (uint32_t)sample, blk, node);
                Bsdf bsdf = make_bsdf(sc.materials + (size_t)ml.x * kMaterialStride, u0.x);
                if (KIND == KIND_LAMBERT) bsdf.kind = K_LAMBERT;  // known at compile time: the other BSDFs' code is pruned
                if (KIND == KIND_MF_CONDUCTOR) bsdf.kind = K_MICROFACET_CONDUCTOR;
                if (KIND == KIND_MF_DIELECTRIC) bsdf.kind = K_MICROFACET_DIELECTRIC;
                if (KIND == KIND_DELTA && bsdf.kind != K_SPECULAR) bsdf.kind = K_FRESNEL_SPECULAR;
                const Frame frame = hit_frame(sc, slot, N);
                const f3 wo_l = to_local(frame, wo);
                // Lambda(wo) of the Smith term is the same for every light of the vertex and for the sampled direction: once
                // (bit-identical values; ~20 % fewer instructions per light in the microfacet kernels)
                const float lambda_o = (KIND == KIND_MF_CONDUCTOR || KIND == KIND_MF_DIELECTRIC) ? tr_lambda(bsdf.ax, bsdf.ay, wo_l) : 0.f;
                if (KIND != KIND_DELTA) {  // integrator.cc:357-372
                    float4 lu = make_float4(0, 0, 0, 0);
                    int lu_block = -1;
#pragma unroll kLightUnroll
                    for (int k = 0; k < sc.n_nee_lights; ++k) {  // black lights are skipped (integrator.cc:362), not sampled
                        const int j = __ldg(sc.nee_lights + k);
                        if ((k >> 1) != lu_block) {  // the k-th non-black light's pair: block k/2, words 2*(k%2)  (rng.cuh)
                            lu_block = k >> 1;
                            lu = rng_block(key, (uint32_t)pixel, (uint32_t)sample, blk + 1u + (uint32_t)lu_block, node);
                        }
                        const float ux = (k & 1) ? lu.z : lu.x, uy = (k & 1) ? lu.w : lu.y;
                        const long long si = (long long)k * n_vertices + kind_base + qi;  // this (vertex, light)'s slot
                        const bool fits = si < p.shadow_capacity;
                        const LightSample ls = sample_light(sc, j, P, N, ux, uy);
                        bool valid = !(is_black(ls.Li) || ls.pdf == 0.f);
                        f3 f = mk3(0, 0, 0);
                        if (valid) {
                            f = bsdf_eval_local(bsdf, wo_l, to_local(frame, ls.wi), lambda_o);
                            valid = !is_black(f);
                        }
                        if (valid && fits) {
                            // FScene::Occluded(isect, ls.pos): scene.h:36-47
                            f3 sdir = ls.wi;       // area and point lights: the light sample already holds
                            float dist = ls.dist;  //   normalize(pos - P) and |pos - P| (light.cuh)
                            if (!(dist > 0.f)) {   // environment / directional lights: pos = P + wi * 2R, normalised again as the reference does
                                const f3 v = ls.pos - P;
                                dist = length(v);
                                sdir = v / dist;
                            }
                            const f3 contrib = cmul(cmul(beta, f), ls.Li) * absdot(ls.wi, N) / ls.pdf;  // integrator.cc:369
                            // a degenerate distance (tmax <= 0 or NaN) can hit nothing: the sample is unoccluded
                            const float tmax = dist - 0.001f;
                            if (tmax > 0.f) {
                                st_stream(&p.sh_o[si], make_float4(P.x, P.y, P.z, tmax));
                                st_stream(&p.sh_d[si], make_float4(sdir.x, sdir.y, sdir.z, 0.f));
                                st_stream(&p.sh_c[si], make_float4(contrib.x, contrib.y, contrib.z, ro.w));  // .w: the pixel
                            } else {
                                st_stream(&p.sh_o[si], make_float4(0.f, 0.f, 0.f, -1.f));
                                film_add(p, pixel, contrib);
                            }
                        } else if (fits) {
                            st_stream(&p.sh_o[si], make_float4(0.f, 0.f, 0.f, -1.f));
                        } else if (valid) {
                            atomicAdd(p.stats + ST_NEE_DROPPED, 1ull);  // no shadow slot left: the sample is lost, and counted
                        }
                    }
                }
                if (!WHITTED) {
                    BsdfSample bs = bsdf_sample_local(bsdf, wo_l, u0.y, u0.z, lambda_o);  // integrator.cc:375
                    bs.wi = to_world(frame, bs.wi);                                          // bsdf.h:296-302
                    if (!(is_black(bs.f) || bs.pdf == 0.f)) {
                        const bool spec = (bs.flags & BSDF_SPECULAR) != 0;
                        bool survive = true;
                        if (bounce >= JPBRT_RR_START_BOUNCE) {  // integrator.cc:383-393
                            const float q = std_max(JPBRT_RR_QMIN, 1 - max_component(bs.f));
                            if (u0.w < q) survive = false;
                            else beta = cmul(beta, bs.f * absdot(bs.wi, N) / (bs.pdf * (1 - q)));
                        } else {
                            beta = cmul(beta, bs.f * absdot(bs.wi, N) / bs.pdf);  // integrator.cc:397
                        }
                        // A non-specular path that would arrive at bounce == maxDepth can add nothing there (no
                        // emission, integrator.cc:328; loop ends, :340): do not trace it.
                        if (survive && (spec || bounce + 1 < sc.max_depth)) {
                            alive = true;
                            no = make_float4(P.x, P.y, P.z, ro.w);
                            nd = make_float4(bs.wi.x, bs.wi.y, bs.wi.z,
                                             __int_as_float(sample | ((bounce + 1) << 24) | (spec ? (int)0x80000000 : 0)));
                            nbeta = make_float4(beta.x, beta.y, beta.z, 0.f);
                        }
                    }
                } else if (KIND == KIND_DELTA && bounce + 1 < sc.max_depth) {
                    // integrator.cc:157-163: SpecularReflect / SpecularTransmit / SpecularReflectAndTransmit each sample the
                    // BSDF whose flags are a SUBSET of theirs (bsdf.h:282): a mirror matches the first and the third (two
                    // rays), FFresnelSpecular only the third; no BSDF of the reference is Specular|Transmission alone.
                    BsdfSample bs = bsdf_sample_local(bsdf, wo_l, u0.y, u0.z);
                    bs.wi = to_world(frame, bs.wi);
                    if (!(is_black(bs.f) || bs.pdf == 0.f)) {
                        beta = cmul(beta, bs.f * absdot(bs.wi, N) / bs.pdf);  // integrator.cc:181
                        const bool mirror = bsdf.kind == K_SPECULAR;
                        alive = true;
                        alive2 = mirror;
                        no = make_float4(P.x, P.y, P.z, ro.w);
                        nd = make_float4(bs.wi.x, bs.wi.y, bs.wi.z, __int_as_float(sample | ((bounce + 1) << 24) | (int)0x80000000));
                        nbeta = make_float4(beta.x, beta.y, beta.z, __uint_as_float(3u * node + (mirror ? 1u : 3u)));
                        nbeta2 = make_float4(beta.x, beta.y, beta.z, __uint_as_float(3u * node + 3u));
                    }
                }
            }
            if (!alive) no = nd = nbeta = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!alive2) nbeta2 = make_float4(0.f, 0.f, 0.f, 0.f);
            append_next<WHITTED>(p, next_count, nbuf, alive, no, nd, nbeta);
            if (WHITTED && KIND == KIND_DELTA) append_next<true>(p, next_count, nbuf, alive2, no, nd, nbeta2);
        }
    }
}

// FDebugIntegrator::Li (integrator.h:47-57): the hit normal as a colour (negative components are summed as they
// are and clamped by Clamp01 at finalize, exactly like DoRender's L += Li * ratio).
__global__ void __launch_bounds__(kBlock) k_debug(const __grid_constant__ WfParams p) {
    const int n = min(p.counters[CNT_RAYS * p.counter_stride + 0], p.queue_capacity);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float2 h = ld_stream(&p.hit[i]);
        const int slot = __float_as_int(h.y);
        if (slot < 0) continue;
        const float4 ro = ld_stream(&p.ray_o[0][i]);
        const f3 d = mk3(ld_stream(&p.ray_d[0][i]));
        const f3 P = mk3(ro) + h.x * d;
        film_add(p, __float_as_int(ro.w), hit_normal(p.sc, slot, P, d));
    }
}

// Paths still queued after the last iteration (only possible with null-material chains) are dropped and counted.
__global__ void k_count_dropped(const __grid_constant__ WfParams p, int it) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int n = p.counters[CNT_RAYS * p.counter_stride + it];
        if (n > 0) atomicAdd(p.stats + ST_DROPPED, (unsigned long long)n);
    }
}

// ---------------------------------------------------------------------------------------------
// accumulate / finalize: out = Clamp01(sum * (1/spp))   (integrator.cc:89,102,108; film.h:22-23)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_finalize(const float* __restrict__ film, float* __restrict__ out, size_t n, float ratio) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = film[i] * ratio;
        out[i] = clampf(v, 0.f, 1.f);
    }
}

}  // namespace jpbrt
