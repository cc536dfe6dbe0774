// c_api.cu -- implementation of the C ABI in include/jetpbrt_b200.h (device side).
//
// jpbrt_upload_scene / jpbrt_render_pass / jpbrt_read_film replace, for the hot path, the
// reference's FScene::Preprocess + FIntegrator::Render/DoRender + FFilmView::AddColor
// (scene.cc:11-23, integrator.cc:35-111, film.h:64-68).  There is no CPU fallback in this file:
// every entry point that computes launches CUDA kernels and reports JPBRT_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jetpbrt_b200.h"
#include "scene_flatten.h"
#include "nccl_dl.h"
#include "bsdf_ex.cuh"
#include "bvh_build.cuh"
#include "wavefront.cuh"
#include "unit_shade.cuh"

#include <cub/device/device_scan.cuh>

using namespace jpbrt;

// resident blocks per SM the default traversal kernels are compiled for (6 = 40 registers; A/B builds: -DJPB_TRAV_MINB=7)
#ifndef JPB_TRAV_MINB
#define JPB_TRAV_MINB 6
#endif
constexpr int kTravMinBlocks = JPB_TRAV_MINB;

namespace {

thread_local std::string g_last_error;

int set_error(jpbrt_ctx* ctx, int code, const char* fmt, ...);

#define CU_CHECK(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return set_error(ctx, JPBRT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <typename T>
struct DevBuf {
    T* ptr = nullptr;
    size_t count = 0;
    cudaError_t Alloc(size_t n) {
        Free();
        if (n == 0) n = 1;
        cudaError_t e = cudaMalloc(&ptr, n * sizeof(T));
        if (e == cudaSuccess) count = n; else ptr = nullptr;
        return e;
    }
    cudaError_t Upload(const T* host, size_t n, cudaStream_t s) {
        if (n == 0) return cudaSuccess;
        return cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void Free() { if (ptr) cudaFree(ptr); ptr = nullptr; count = 0; }
    ~DevBuf() { Free(); }
};

struct StageEvent {
    cudaEvent_t a, b;
    int stage;
};

}  // namespace

struct jpbrt_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::string error;
    HostScene hs;
    // scene on the device
    DevBuf<Float4> nodes, qnodes, slots, slot_nrm, materials, lights, slot_frame;
    DevBuf<Int2> slot_ml;
    DevBuf<int> inf_lights, prim_slot, nee_lights, pixel_order;
    DevScene dsc{};
    // wavefront state
    long long paths_in_flight = 0;  // capacity of the path pool
    long long pool_limit = 0;       // default upper bound of the pool, fixed at the first pass
    bool pool_explicit = false;     // the pool was sized by the "paths_in_flight" option
    DevBuf<float4> ray_o[2], ray_d[2], ray_b[2], sh_o, sh_d, sh_c;
    DevBuf<float2> hit;
    DevBuf<int> kind_queue;
    // ray reordering (option "sort_rays"): bin counters, their prefix sums, (bin, rank) per ray, the permutation, cub's scratch
    DevBuf<unsigned> sort_bins, sort_start;
    DevBuf<uint2> sort_kr;
    DevBuf<int> perm;
    DevBuf<unsigned char> scan_tmp;
    DevBuf<int> counters;
    int counter_stride = 0;
    int n_iters = 0;
    DevBuf<float> film;
    DevBuf<float> film_final;  // Clamp01(mean) staging buffer of jpbrt_read_film (no cudaMalloc/cudaFree per call)
    DevBuf<unsigned long long> dstats;
    DevBuf<PassArgs> pass_args;
    // one wavefront (generate + all iterations) captured as a CUDA graph; rebuilt when an option that changes the
    // launch sequence changes
    cudaGraphExec_t wave_graph = nullptr;
    long long wave_graph_key = -1;
    unsigned long long wave_graph_launches = 0;
    bool opt_use_graph = true;
    // options
    long long opt_paths_in_flight = 0;
    bool opt_stage_timing = false;
    bool opt_count_traversal = false;
    int opt_refill_min = -1;  // < 0: automatic (16, or 20 for trees of at most 64 nodes)
    int opt_min_inner = -1;  // < 0: automatic (8, or 4 for trees of at most 64 nodes); 0 or 1: wait for every lane
    long long opt_band_pixels = 0;  // pixels per band of a wavefront (0 = default 2^20); >= the frame: no banding
    int opt_integrator = JPBRT_INTEGRATOR_PATH;  // jpbrt_integrator: which FIntegrator::Li the passes evaluate
    bool has_mirror = false;                     // Whitted traces a mirror vertex twice: the ray tree can grow
    int opt_node_format = -1;  // -1 automatic, 0 64-byte float nodes, 1 32-byte quantised nodes (if the scene uploaded them)
    bool has_qnodes = false;   // the quantised copy of the tree is resident (scenes of at most kQNodesMaxNodes nodes)
    int opt_sort_rays = 0;    // 0 off; else cell bits per axis (1..6) + 16 x (direction octant in the key)
    int opt_trav_blocks = 6;  // resident blocks per SM the traversal kernels are compiled for: 6 (40 registers, default) or 5 (48)
    unsigned kinds_present = 0;  // bit k set: some material of the scene can build BSDF kind k
    // launch geometry
    int grid_generate = 0, grid_extend = 0, grid_extend_c = 0, grid_extend6 = 0, grid_connect6 = 0, grid_logic = 0, grid_logic_w = 0, grid_debug = 0, grid_shade[4] = {0, 0, 0, 0}, grid_shade_w[4] = {0, 0, 0, 0}, grid_connect = 0, grid_connect_c = 0, grid_finalize = 0;
    // multi-GPU: this context's rank in an NCCL communicator (jpbrt_comm_init); the film reduce runs on `stream`
    ncclComm_t comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    bool film_reduced = false;  // the film was reduced since the last pass: jpbrt_read_film must not reduce it again
    // host-side accounting
    unsigned long long kernel_launches = 0;
    double ms_stage[6] = {0, 0, 0, 0, 0, 0};  // generate, extend, shade, connect, finalize, reduce
    std::vector<StageEvent> pending_events;
    std::vector<cudaEvent_t> free_events;
    std::vector<void*> pinned;  // host scene arrays registered with cudaHostRegister
};

namespace {

int set_error(jpbrt_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->error = buf;
    return code;
}

int select_device(jpbrt_ctx* ctx, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return set_error(ctx, JPBRT_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return set_error(ctx, JPBRT_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    CU_CHECK(ctx, cudaSetDevice(device));
    return 0;
}

cudaEvent_t get_event(jpbrt_ctx* c) {
    if (!c->free_events.empty()) { cudaEvent_t e = c->free_events.back(); c->free_events.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct StageTimer {
    jpbrt_ctx* c;
    int stage;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(jpbrt_ctx* ctx, int st) : c(ctx), stage(st) {
        if (c->opt_stage_timing) { a = get_event(c); b = get_event(c); cudaEventRecord(a, c->stream); }
    }
    ~StageTimer() {
        if (c->opt_stage_timing) { cudaEventRecord(b, c->stream); c->pending_events.push_back(StageEvent{a, b, stage}); }
    }
};

void drain_events(jpbrt_ctx* c) {
    for (auto& ev : c->pending_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) c->ms_stage[ev.stage] += ms;
        c->free_events.push_back(ev.a);
        c->free_events.push_back(ev.b);
    }
    c->pending_events.clear();
}

// Which node format the production traversal kernels walk.  Quantised 32-byte nodes halve the node loads -- what the L1 pipe,
// the traversal kernels' bound, spends most of its time on -- at the price of 12 PRMT per node step and boxes rounded outwards by
// up to 3 grid cells.  Measured (profiles/ab/r02_ab_qnodes.log): bunny scene (10 k nodes) k_connect -5.9 %; trees of 15 / 33
// nodes (Cornell, glossy) +5-9 % slower; 5 M triangles k_extend -4 %, but there the fatter boxes also admit primitives whose
// (reference-exact, float-sloppy at t ~ 1000) edge tests accept rays that miss their bounds by more than the float nodes'
// padding -- 0.17 % of camera rays find a different first hit than with the float nodes.  Automatic = trees of 1,024 to 2^20 nodes.
constexpr int kQNodesMinNodes = 1024;  // (kQNodesMaxNodes: scene_flatten.h -- larger trees get no quantised copy at all)
static bool use_qnodes(const jpbrt_ctx* c) {
    if (!c->has_qnodes || c->opt_node_format == 0) return false;
    if (c->opt_node_format == 1) return true;
    return (int)(c->hs.nodes.size() / kNodeStride) >= kQNodesMinNodes;
}

template <typename K>
int occupancy_grid(jpbrt_ctx* c, K kernel) {
    int per_sm = 0;
    // A/B hook: JPBRT_CARVEOUT = preferred shared-memory share of the unified L1 / shared array in percent (0 = all of it to L1).
    // Measured (profiles/ab/r02_ab_carveout.log): unset == 0 (the kernels use no shared memory to speak of and already get the
    // whole array as L1); 25 % costs the bunny scene's traversal 1 %, 50 % costs 3-4 %.
    if (const char* v = getenv("JPBRT_CARVEOUT")) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(v)) != cudaSuccess) cudaGetLastError();
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm <= 0) per_sm = 1;
    return c->sm_count * per_sm;
}

// Page-lock the flattened host arrays so that (re)uploads are true asynchronous DMA transfers.
template <typename V>
void pin_vector(jpbrt_ctx* c, V& v) {
    if (v.empty()) return;
    if (cudaHostRegister(v.data(), v.size() * sizeof(typename V::value_type), cudaHostRegisterDefault) == cudaSuccess) c->pinned.push_back(v.data());
    else cudaGetLastError();
}

// slot_frame[s] = FFrame(normal of slot s) (geometry.h:344-377) for flat shapes, derived ON THE DEVICE from the uploaded
// normals: the same IEEE expressions the host used to evaluate (make_frame, dmath.cuh; this file is compiled without FMA
// contraction), so the values are bit-identical -- and 48 of the 244 bytes per primitive never cross PCIe.
__global__ void __launch_bounds__(kBlock) k_make_frames(const Float4* __restrict__ slot_nrm, Float4* __restrict__ slot_frame, int n) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const float4 nr = ldg4(slot_nrm + s);
        const int type = __float_as_int(nr.w) & ((1 << kTypeBits) - 1);
        float4 f0 = make_float4(0, 0, 0, 0), f1 = f0, f2 = f0;  // a sphere's frame depends on the hit point (hit_frame)
        if (type != SHAPE_SPHERE) {
            const Frame f = make_frame(mk3(nr));
            f0 = make_float4(f.s.x, f.s.y, f.s.z, 0.f);
            f1 = make_float4(f.t.x, f.t.y, f.t.z, 0.f);
            f2 = make_float4(f.n.x, f.n.y, f.n.z, 0.f);
        }
        float4* out = reinterpret_cast<float4*>(slot_frame + (size_t)s * kFrameStride);
        out[0] = f0; out[1] = f1; out[2] = f2;
    }
}

int upload_arrays(jpbrt_ctx* c, size_t* bytes) {
    HostScene& hs = c->hs;
    CU_CHECK(c, c->nodes.Upload(hs.nodes.data(), hs.nodes.size(), c->stream));
    if (c->has_qnodes) CU_CHECK(c, c->qnodes.Upload(hs.qnodes.data(), hs.qnodes.size(), c->stream));
    CU_CHECK(c, c->slots.Upload(hs.slots.data(), hs.slots.size(), c->stream));
    CU_CHECK(c, c->slot_nrm.Upload(hs.slot_nrm.data(), hs.slot_nrm.size(), c->stream));
    CU_CHECK(c, c->slot_ml.Upload(hs.slot_ml.data(), hs.slot_ml.size(), c->stream));
    CU_CHECK(c, c->materials.Upload(hs.materials.data(), hs.materials.size(), c->stream));
    CU_CHECK(c, c->lights.Upload(hs.lights.data(), hs.lights.size(), c->stream));
    CU_CHECK(c, c->inf_lights.Upload(hs.inf_lights.data(), hs.inf_lights.size(), c->stream));
    CU_CHECK(c, c->prim_slot.Upload(hs.prim_slot.data(), hs.prim_slot.size(), c->stream));
    {
        const int n = (int)hs.slot_nrm.size();
        k_make_frames<<<std::min(148 * 8, (n + kBlock - 1) / kBlock), kBlock, 0, c->stream>>>(c->slot_nrm.ptr, c->slot_frame.ptr, n);
        CU_CHECK(c, cudaGetLastError());
        c->kernel_launches++;
    }
    CU_CHECK(c, c->nee_lights.Upload(hs.nee_lights.data(), hs.nee_lights.size(), c->stream));
    if (bytes) *bytes = hs.Bytes() + (c->has_qnodes ? hs.qnodes.size() * sizeof(Float4) : 0);
    return 0;
}

// Bytes of wavefront state per path in flight: 2 x 48 B ray records + 8 B hit + 4 kind-queue indices + one
// 48 B shadow slot per non-black light.
size_t bytes_per_path(int n_nee_lights) { return 2 * 48 + 8 + 4 * NUM_KINDS + (size_t)48 * std::max(1, n_nee_lights); }

void free_pool(jpbrt_ctx* c) {
    for (int b = 0; b < 2; ++b) { c->ray_o[b].Free(); c->ray_d[b].Free(); c->ray_b[b].Free(); }
    c->hit.Free(); c->kind_queue.Free(); c->sh_o.Free(); c->sh_d.Free(); c->sh_c.Free();
    c->sort_kr.Free(); c->perm.Free();
}

// Size the path pool for a pass of `paths_needed` camera paths.  An explicit "paths_in_flight" option is honoured
// exactly; otherwise the pool grows on demand up to the default limit and is never shrunk.
int ensure_pool(jpbrt_ctx* c, long long paths_needed) {
    const long long npix = (long long)c->hs.width * c->hs.height;
    const int n_lights = std::max(1, (int)c->hs.nee_lights.size());  // one shadow slot per (vertex, non-black light)
    long long want;
    if (c->opt_paths_in_flight > 0) {
        want = c->opt_paths_in_flight;
        c->pool_explicit = true;
    } else {
        // Default limit: 2^26 paths (measured on B200: 8 M -> 32 M paths in flight is +13 % on the bunny scene, +6 % on
        // Cornell, 32 M -> 52 M -- the 50-spp pass as ONE wavefront instead of two -- another +3.5 %: longer launches,
        // shorter tails; 14.5 GB of the 180 GB at two lights), but never more than a quarter of the device's free memory
        // as it was at the first pass.  A pass that needs fewer paths gets a pool of its own size.
        if (c->pool_limit == 0) {
            c->pool_limit = 1ll << 26;
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                c->pool_limit = std::min(c->pool_limit, std::max((long long)(free_b / 4 / bytes_per_path(n_lights)), npix));
        }
        want = std::min(c->pool_limit, std::max(paths_needed, npix));
        if (!c->pool_explicit && c->paths_in_flight >= want) return 0;  // large enough already
        c->pool_explicit = false;
    }
    // whole samples only: the pool holds k full-frame samples
    long long k = std::max(1ll, want / npix);
    long long cap = k * npix;
    if (cap > 0x7fff0000ll) return set_error(c, JPBRT_ERR_UNSUPPORTED, "film too large for one wavefront (%lld pixels)", npix);
    long long shadow_cap = cap * n_lights;
    if (shadow_cap > 0x7fff0000ll) {  // keep queue indices in int range
        k = std::max(1ll, 0x7fff0000ll / (npix * n_lights));
        cap = k * npix;
        shadow_cap = cap * n_lights;
        if (shadow_cap > 0x7fff0000ll) return set_error(c, JPBRT_ERR_UNSUPPORTED, "pixels x lights too large for one wavefront");
    }
    if (cap == c->paths_in_flight) return 0;
    if (c->wave_graph) { cudaGraphExecDestroy(c->wave_graph); c->wave_graph = nullptr; }  // it holds the old buffers' addresses
    // The pool is reallocated as a whole: while that is under way the context owns NO pool (capacity 0), and a failed
    // allocation releases everything, so that the next pass starts from scratch instead of trusting a stale capacity.
    c->paths_in_flight = 0;
    free_pool(c);
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        if ((e = c->ray_o[b].Alloc(cap)) == cudaSuccess && (e = c->ray_d[b].Alloc(cap)) == cudaSuccess) e = c->ray_b[b].Alloc(cap);
    }
    if (e == cudaSuccess) e = c->hit.Alloc(cap);
    if (e == cudaSuccess) e = c->kind_queue.Alloc((size_t)cap * NUM_KINDS);
    if (e == cudaSuccess) e = c->sh_o.Alloc(shadow_cap);
    if (e == cudaSuccess) e = c->sh_d.Alloc(shadow_cap);
    if (e == cudaSuccess) e = c->sh_c.Alloc(shadow_cap);
    if (e != cudaSuccess) {
        free_pool(c);
        cudaGetLastError();
        return set_error(c, JPBRT_ERR_CUDA, "allocating a pool of %lld paths failed: %s", cap, cudaGetErrorString(e));
    }
    c->paths_in_flight = cap;
    return 0;
}

// Buffers of the ray reordering ("sort_rays"): sized with the pool, allocated on first use.
constexpr int kSortMaxKeyBits = 21;  // 6 cell bits per axis + 3 direction bits
int ensure_sort_buffers(jpbrt_ctx* c) {
    if (c->opt_sort_rays == 0) return 0;
    const size_t cap = (size_t)c->paths_in_flight;
    if (c->sort_kr.ptr && c->sort_kr.count == cap) return 0;
    if (c->wave_graph) { cudaGraphExecDestroy(c->wave_graph); c->wave_graph = nullptr; }
    const size_t nbins = (size_t)1 << kSortMaxKeyBits;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (int)nbins, c->stream);
    cudaError_t e;
    if ((e = c->sort_kr.Alloc(cap)) != cudaSuccess || (e = c->perm.Alloc(cap)) != cudaSuccess ||
        (e = c->sort_bins.Alloc(nbins)) != cudaSuccess || (e = c->sort_start.Alloc(nbins)) != cudaSuccess ||
        (e = c->scan_tmp.Alloc(tmp_bytes)) != cudaSuccess) {
        c->sort_kr.Free(); c->perm.Free(); c->sort_bins.Free(); c->sort_start.Free(); c->scan_tmp.Free();
        cudaGetLastError();
        return set_error(c, JPBRT_ERR_CUDA, "allocating the reordering buffers failed: %s", cudaGetErrorString(e));
    }
    CU_CHECK(c, cudaMemsetAsync(c->sort_bins.ptr, 0, nbins * sizeof(unsigned), c->stream));
    return 0;
}

WfParams make_params(jpbrt_ctx* c) {
    WfParams p{};
    p.sc = c->dsc;
    for (int b = 0; b < 2; ++b) { p.ray_o[b] = c->ray_o[b].ptr; p.ray_d[b] = c->ray_d[b].ptr; p.ray_b[b] = c->ray_b[b].ptr; }
    p.hit = c->hit.ptr;
    p.kind_queue = c->kind_queue.ptr;
    p.queue_capacity = (int)c->paths_in_flight;
    p.sh_o = c->sh_o.ptr;
    p.sh_d = c->sh_d.ptr;
    p.sh_c = c->sh_c.ptr;
    p.counters = c->counters.ptr;
    p.counter_stride = c->counter_stride;
    p.film = c->film.ptr;
    p.stats = c->dstats.ptr;
    p.args = c->pass_args.ptr;
    p.npix = c->hs.width * c->hs.height;
    p.blocks_per_bounce = rng_blocks_per_bounce(c->dsc.n_nee_lights);  // black lights own no sampler dimensions: nothing ever reads them
    p.shadow_capacity = (int)std::min<size_t>(c->sh_o.count, 0x7fffffff);
    // B200 sweep (profiles/r01_sweep_min_inner.txt): trees of thousands of nodes want min_inner 8 / refill 16 (bunny scene +14 %,
    // 5 M triangles +22 % over waiting for every lane); the 16- and 33-node trees of Cornell / glossy want 4 / 20 (+1 %).
    const bool tiny_tree = c->dsc.n_nodes <= 64;
    p.refill_min = c->opt_refill_min > 0 ? c->opt_refill_min : (tiny_tree ? 20 : 16);
    p.min_inner = c->opt_min_inner >= 0 ? c->opt_min_inner : (tiny_tree ? 4 : 8);
    const bool sorting = c->opt_sort_rays != 0 && c->sort_kr.ptr && c->hs.bvh_depth <= kMaxBvhDepth && (c->opt_integrator == JPBRT_INTEGRATOR_PATH || c->opt_integrator == JPBRT_INTEGRATOR_PATH_RECURSIVE);
    if (sorting) {
        p.sort_bins = c->sort_bins.ptr;
        p.sort_start = c->sort_start.ptr;
        p.sort_kr = c->sort_kr.ptr;
        p.perm = c->perm.ptr;
        p.sort_cell_bits = c->opt_sort_rays & 15;
        p.sort_dir = (c->opt_sort_rays >> 4) & 1;
        for (int a = 0; a < 3; ++a) {
            const float ext = c->hs.world_max[a] - c->hs.world_min[a];
            p.sort_min[a] = c->hs.world_min[a];
            p.sort_scale[a] = ext > 0.f ? (float)(1 << p.sort_cell_bits) / ext : 0.f;
        }
    }
    return p;
}

}  // namespace

// Second half of an upload: the context's HostScene is flattened; allocate, copy and size everything on `device`.
// On failure the context is destroyed.
static int finish_upload(jpbrt_ctx* c, int device) {
    int rc = select_device(nullptr, device);
    if (rc != 0) { delete c; return rc; }
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    auto fail = [&](int code) { jpbrt_destroy(c); return code; };
    if (!c->stream && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(set_error(nullptr, JPBRT_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError())));
    HostScene& hs = c->hs;
    cudaError_t e = cudaSuccess;
    // The quantised copy of the tree rides along for scenes whose trees are neither tiny nor huge (see use_qnodes()).
    c->has_qnodes = hs.nodes.size() / kNodeStride <= (size_t)kQNodesMaxNodes && !hs.qnodes.empty();
    if ((e = c->nodes.Alloc(hs.nodes.size())) != cudaSuccess || (e = c->qnodes.Alloc(c->has_qnodes ? hs.qnodes.size() : 1)) != cudaSuccess || (e = c->slots.Alloc(hs.slots.size())) != cudaSuccess ||
        (e = c->slot_nrm.Alloc(hs.slot_nrm.size())) != cudaSuccess || (e = c->slot_ml.Alloc(hs.slot_ml.size())) != cudaSuccess ||
        (e = c->materials.Alloc(hs.materials.size())) != cudaSuccess || (e = c->lights.Alloc(hs.lights.size())) != cudaSuccess ||
        (e = c->inf_lights.Alloc(hs.inf_lights.size())) != cudaSuccess || (e = c->prim_slot.Alloc(hs.prim_slot.size())) != cudaSuccess ||
        (e = c->slot_frame.Alloc(hs.slot_nrm.size() * kFrameStride)) != cudaSuccess || (e = c->nee_lights.Alloc(hs.nee_lights.size())) != cudaSuccess ||
        (e = c->pixel_order.Alloc(hs.pixel_order.size())) != cudaSuccess ||
        (e = c->film.Alloc((size_t)hs.width * hs.height * 3)) != cudaSuccess ||
        (e = c->film_final.Alloc((size_t)hs.width * hs.height * 3)) != cudaSuccess || (e = c->dstats.Alloc(ST_COUNT)) != cudaSuccess ||
        (e = c->pass_args.Alloc(1)) != cudaSuccess)
        return fail(set_error(nullptr, JPBRT_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)));
    pin_vector(c, hs.nodes);
    if (c->has_qnodes) pin_vector(c, hs.qnodes);
    pin_vector(c, hs.slots); pin_vector(c, hs.slot_nrm); pin_vector(c, hs.slot_ml);
    pin_vector(c, hs.materials); pin_vector(c, hs.lights); pin_vector(c, hs.inf_lights); pin_vector(c, hs.prim_slot);
    pin_vector(c, hs.nee_lights);
    rc = upload_arrays(c, nullptr);
    if (rc == 0 && c->pixel_order.Upload(hs.pixel_order.data(), hs.pixel_order.size(), c->stream) != cudaSuccess)
        rc = set_error(c, JPBRT_ERR_CUDA, "pixel order upload failed");
    if (rc != 0) { g_last_error = c->error; return fail(rc); }
    DevScene& d = c->dsc;
    d.pixel_order = c->pixel_order.ptr;
    d.nodes = c->nodes.ptr; d.qnodes = c->qnodes.ptr; d.slots = c->slots.ptr;
    for (int a = 0; a < 3; ++a) { d.q_origin[a] = hs.q_origin[a]; d.q_cell[a] = hs.q_cell[a]; } d.slot_nrm = c->slot_nrm.ptr; d.slot_ml = c->slot_ml.ptr;
    d.materials = c->materials.ptr; d.lights = c->lights.ptr; d.inf_lights = c->inf_lights.ptr; d.prim_slot = c->prim_slot.ptr;
    d.slot_frame = c->slot_frame.ptr; d.nee_lights = c->nee_lights.ptr;
    d.n_nee_lights = (int)hs.nee_lights.size();
    d.n_nodes = (int)(hs.nodes.size() / kNodeStride);
    d.n_slots = (int)hs.slot_nrm.size();
    d.n_materials = (int)(hs.materials.size() / kMaterialStride);
    d.n_lights = (int)(hs.lights.size() / kLightStride);
    d.n_inf_lights = (int)hs.inf_lights.size();
    d.n_prims = hs.n_prims;
    d.max_depth = hs.max_depth;
    d.max_leaf_prims = hs.max_leaf_prims;
    d.width = hs.width;
    d.height = hs.height;
    d.world_radius = hs.world_radius;
    d.cam = hs.cam;
    for (const Int2& ml : hs.slot_ml) {  // BSDF kinds that materials bound to primitives can build
        if (ml.x < 0) continue;
        int type;
        memcpy(&type, &hs.materials[(size_t)ml.x * kMaterialStride].w, 4);
        c->kinds_present |= type == JPBRT_MAT_MATTE ? 1u : type == JPBRT_MAT_METAL ? 2u : type == JPBRT_MAT_PLASTIC ? (1u | 4u) : 8u;
        if (type == JPBRT_MAT_MIRROR) c->has_mirror = true;
    }
    // iterations: bounces 0..max_depth, plus slack for null-material pass-through vertices
    c->n_iters = hs.max_depth + 1 + (hs.has_null_material ? 16 : 0);
    c->counter_stride = c->n_iters + 2;
    if ((e = c->counters.Alloc((size_t)CNT_KINDS * c->counter_stride)) != cudaSuccess)
        return fail(set_error(nullptr, JPBRT_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)));
    c->grid_generate = occupancy_grid(c, k_generate);
    c->grid_extend = occupancy_grid(c, k_extend<false, 5>);
    c->grid_extend6 = occupancy_grid(c, k_extend<false, kTravMinBlocks>);
    c->grid_extend_c = occupancy_grid(c, k_extend<true, 5>);
    c->grid_logic = occupancy_grid(c, k_logic<false>);
    c->grid_logic_w = occupancy_grid(c, k_logic<true>);
    c->grid_debug = occupancy_grid(c, k_debug);
    c->grid_shade[0] = occupancy_grid(c, k_shade<0>);
    c->grid_shade[1] = occupancy_grid(c, k_shade<1>);
    c->grid_shade[2] = occupancy_grid(c, k_shade<2>);
    c->grid_shade[3] = occupancy_grid(c, k_shade<3>);
    c->grid_shade_w[0] = occupancy_grid(c, k_shade<0, true>);
    c->grid_shade_w[1] = occupancy_grid(c, k_shade<1, true>);
    c->grid_shade_w[2] = occupancy_grid(c, k_shade<2, true>);
    c->grid_shade_w[3] = occupancy_grid(c, k_shade<3, true>);
    c->grid_connect = occupancy_grid(c, k_connect<false, 5>);
    c->grid_connect6 = occupancy_grid(c, k_connect<false, kTravMinBlocks>);
    c->grid_connect_c = occupancy_grid(c, k_connect<true, 5>);
    c->grid_finalize = occupancy_grid(c, k_finalize);
    rc = jpbrt_clear_film(c);
    if (rc == 0) rc = jpbrt_reset_stats(c);
    if (rc != 0) { g_last_error = c->error; return fail(rc); }
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess)
        return fail(set_error(nullptr, JPBRT_ERR_CUDA, "scene upload failed: %s", cudaGetErrorString(e)));
    return 0;
}


extern "C" {

int jpbrt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

long long jpbrt_debug_flatten(const jpbrt_scene_desc* desc, int what, void* out, long long capacity) {
    HostScene hs;
    std::string err;
    int rc = FlattenScene(desc, &hs, &err);
    if (rc != 0) return set_error(nullptr, rc, "%s", err.c_str());
    const void* src = nullptr;
    long long words = 0;
    switch (what) {
    case 0: src = hs.nodes.data(); words = (long long)hs.nodes.size() * 4; break;
    case 1: src = hs.slots.data(); words = (long long)hs.slots.size() * 4; break;
    case 2: src = hs.slot_nrm.data(); words = (long long)hs.slot_nrm.size() * 4; break;
    case 3: src = hs.slot_ml.data(); words = (long long)hs.slot_ml.size() * 2; break;
    case 4: src = hs.prim_slot.data(); words = (long long)hs.prim_slot.size(); break;
    case 5: src = hs.materials.data(); words = (long long)hs.materials.size() * 4; break;
    case 6: src = hs.lights.data(); words = (long long)hs.lights.size() * 4; break;
    case 7: src = hs.qnodes.data(); words = (long long)hs.qnodes.size() * 4; break;
    case 8: {  // the quantised nodes' grid: origin xyz, cell xyz
        float grid[6] = {hs.q_origin[0], hs.q_origin[1], hs.q_origin[2], hs.q_cell[0], hs.q_cell[1], hs.q_cell[2]};
        if (out && capacity > 0) memcpy(out, grid, (size_t)std::min(6ll, capacity) * 4);
        return 6;
    }
    default: return set_error(nullptr, JPBRT_ERR_INVALID, "unknown table %d", what);
    }
    if (out && capacity > 0) memcpy(out, src, (size_t)std::min(words, capacity) * 4);
    return words;
}

long long jpbrt_debug_ctx_table(jpbrt_ctx* c, int what, void* out, long long capacity) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    const HostScene& hs = c->hs;
    const void* src = nullptr;
    long long words = 0;
    switch (what) {
    case 0: src = hs.nodes.data(); words = (long long)hs.nodes.size() * 4; break;
    case 1: src = hs.slots.data(); words = (long long)hs.slots.size() * 4; break;
    case 2: src = hs.slot_nrm.data(); words = (long long)hs.slot_nrm.size() * 4; break;
    case 3: src = hs.slot_ml.data(); words = (long long)hs.slot_ml.size() * 2; break;
    case 4: src = hs.prim_slot.data(); words = (long long)hs.prim_slot.size(); break;
    default: return set_error(c, JPBRT_ERR_INVALID, "unknown table %d", what);
    }
    if (out && capacity > 0) memcpy(out, src, (size_t)std::min(words, capacity) * 4);
    return words;
}

const char* jpbrt_version(void) { return "jet-pbrt_b200 0.1 (sm_100a wavefront path tracer)"; }

const char* jpbrt_last_error(const jpbrt_ctx* ctx) { return ctx ? ctx->error.c_str() : g_last_error.c_str(); }

int jpbrt_upload_scene(const jpbrt_scene_desc* desc, int device, jpbrt_ctx** out_ctx) {
    unsigned flags = 0;
    if (const char* v = getenv("JPBRT_BVH_BUILDER")) flags |= (!strcmp(v, "gpu") || !strcmp(v, "lbvh")) ? JPBRT_UPLOAD_GPU_BVH : 0u;
    return jpbrt_upload_scene_ex(desc, device, flags, out_ctx);
}

int jpbrt_upload_scene_ex(const jpbrt_scene_desc* desc, int device, unsigned flags, jpbrt_ctx** out_ctx) {
    if (!out_ctx) return set_error(nullptr, JPBRT_ERR_INVALID, "out_ctx is null");
    *out_ctx = nullptr;
    if (flags & ~(unsigned)JPBRT_UPLOAD_GPU_BVH) return set_error(nullptr, JPBRT_ERR_INVALID, "unknown upload flags 0x%x", flags);
    jpbrt_ctx* c = new jpbrt_ctx();
    std::string err;
    int rc;
    if (flags & JPBRT_UPLOAD_GPU_BVH) {
        // the BVH is built on the device (csrc/bvh_build.cuh): the device comes first
        rc = select_device(nullptr, device);
        if (rc != 0) { delete c; return rc; }
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            rc = set_error(nullptr, JPBRT_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            delete c;
            return rc;
        }
        rc = FlattenScene(desc, &c->hs, &err, lbvh::build_on_device, (void*)c->stream);
        if (rc != 0) { set_error(nullptr, rc, "%s", err.c_str()); cudaStreamDestroy(c->stream); delete c; return rc; }
    } else {
        // validate and flatten first: a malformed description is reported even where no device exists
        rc = FlattenScene(desc, &c->hs, &err);
        if (rc != 0) { set_error(nullptr, rc, "%s", err.c_str()); delete c; return rc; }
        rc = select_device(nullptr, device);
        if (rc != 0) { delete c; return rc; }
    }
    rc = finish_upload(c, device);
    if (rc != 0) return rc;
    *out_ctx = c;
    return 0;
}

int jpbrt_reupload_scene(jpbrt_ctx* c, size_t* bytes) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    CU_CHECK(c, cudaSetDevice(c->device));
    return upload_arrays(c, bytes);
}

int jpbrt_clear_film(jpbrt_ctx* c) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    CU_CHECK(c, cudaSetDevice(c->device));
    CU_CHECK(c, cudaMemsetAsync(c->film.ptr, 0, c->film.count * sizeof(float), c->stream));
    c->film_reduced = false;  // a cleared film is a new partial sum: the next read reduces again (also on a rank that then renders nothing)
    return 0;
}

int jpbrt_reset_stats(jpbrt_ctx* c) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    CU_CHECK(c, cudaSetDevice(c->device));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    drain_events(c);
    CU_CHECK(c, cudaMemsetAsync(c->dstats.ptr, 0, ST_COUNT * sizeof(unsigned long long), c->stream));
    c->kernel_launches = 0;
    for (double& m : c->ms_stage) m = 0;
    return 0;
}

int jpbrt_set_option(jpbrt_ctx* c, const char* name, long long value) {
    if (!c || !name) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    if (!strcmp(name, "paths_in_flight")) { c->opt_paths_in_flight = value; return 0; }
    if (!strcmp(name, "stage_timing")) { c->opt_stage_timing = value != 0; return 0; }
    if (!strcmp(name, "count_traversal")) { c->opt_count_traversal = value != 0; return 0; }
    if (!strcmp(name, "trav_blocks")) { c->opt_trav_blocks = value >= 6 ? 6 : 5; return 0; }
    if (!strcmp(name, "sort_rays")) {
        if (value != 0 && ((value & 15) < 1 || (value & 15) > 6 || (value >> 5) != 0)) return set_error(c, JPBRT_ERR_INVALID, "sort_rays: cell bits 1..6 (+16 for the direction octant), got %lld", value);
        c->opt_sort_rays = (int)value;
        return 0;
    }
    if (!strcmp(name, "use_graph")) { c->opt_use_graph = value != 0; return 0; }
    if (!strcmp(name, "band_pixels")) { c->opt_band_pixels = std::max(0ll, value); return 0; }
    if (!strcmp(name, "integrator")) {
        if (value < JPBRT_INTEGRATOR_PATH || value > JPBRT_INTEGRATOR_DEBUG) return set_error(c, JPBRT_ERR_INVALID, "unknown integrator %lld", value);
        c->opt_integrator = (int)value;
        return 0;
    }
    if (!strcmp(name, "min_inner")) { c->opt_min_inner = (int)std::max(-1ll, std::min(32ll, value)); return 0; }
    if (!strcmp(name, "node_format")) { c->opt_node_format = (int)std::max(-1ll, std::min(1ll, value)); return 0; }
    if (!strcmp(name, "refill_min")) { c->opt_refill_min = (int)std::max(-1ll, std::min(32ll, value)); return 0; }
    return set_error(c, JPBRT_ERR_INVALID, "unknown option '%s'", name);
}

// Queue (or capture) the launches of ONE wavefront on the context's stream: generate, then for every bounce
// extend -> logic -> shade<kind>... -> connect.  Everything that differs between wavefronts is in PassArgs.
static int queue_wavefront(jpbrt_ctx* c, bool count) {
    WfParams p = make_params(c);
    CU_CHECK(c, cudaMemsetAsync(c->counters.ptr, 0, c->counters.count * sizeof(int), c->stream));
    if (p.sort_bins)  // (a null-material pass-through at the last bounce can leave counts behind)
        CU_CHECK(c, cudaMemsetAsync(c->sort_bins.ptr, 0, sizeof(unsigned) << (3 * p.sort_cell_bits + 3 * p.sort_dir), c->stream));
    {
        StageTimer t(c, 0);
        k_generate<<<c->grid_generate, kBlock, 0, c->stream>>>(p);
        c->kernel_launches++;
    }
    const bool whitted = c->opt_integrator == JPBRT_INTEGRATOR_WHITTED;
    const bool debug = c->opt_integrator == JPBRT_INTEGRATOR_DEBUG;
    const bool sorting = p.sort_bins != nullptr;
    // the 6-block traversal kernels run WITHOUT the full-stack test of every push (intersect.cuh: GUARD): only for trees whose
    // depth the uploader verified; a deeper tree (test hook) takes the guarded 5-block kernels, which count what they lose
    const bool fast6 = c->opt_trav_blocks >= 6 && c->hs.bvh_depth <= kMaxBvhDepth;
    const bool qn = fast6 && use_qnodes(c);  // (the counting, 5-block and reordering variants walk the float nodes)
    for (int it = 0; it < (debug ? 1 : c->n_iters); ++it) {
        if (sorting && it > 0) {
            // reordering: bins were counted while bounce it-1 appended its rays; prefix-sum them, place every ray, clear the bins
            StageTimer t(c, 1);
            size_t tmp_bytes = c->scan_tmp.count;
            CU_CHECK(c, cub::DeviceScan::ExclusiveSum(c->scan_tmp.ptr, tmp_bytes, c->sort_bins.ptr, c->sort_start.ptr,
                                                      1 << (3 * p.sort_cell_bits + 3 * p.sort_dir), c->stream));
            k_permute<<<c->grid_generate, kBlock, 0, c->stream>>>(p, it);
            CU_CHECK(c, cudaMemsetAsync(c->sort_bins.ptr, 0, sizeof(unsigned) << (3 * p.sort_cell_bits + 3 * p.sort_dir), c->stream));
            c->kernel_launches += 2;
        }
        {
            StageTimer t(c, 1);
            if (count) k_extend<true, 5><<<c->grid_extend_c, kBlock, 0, c->stream>>>(p, it);
            else if (sorting && it > 0 && fast6) k_extend<false, kTravMinBlocks, true><<<c->grid_extend6, kBlock, 0, c->stream>>>(p, it);
            else if (qn) k_extend<false, kTravMinBlocks, false, true><<<c->grid_extend6, kBlock, 0, c->stream>>>(p, it);
            else if (fast6) k_extend<false, kTravMinBlocks><<<c->grid_extend6, kBlock, 0, c->stream>>>(p, it);
            else k_extend<false, 5><<<c->grid_extend, kBlock, 0, c->stream>>>(p, it);
            c->kernel_launches++;
        }
        if (debug) {  // FDebugIntegrator: the hit normal is the radiance (integrator.h:47-57)
            StageTimer t(c, 2);
            k_debug<<<c->grid_debug, kBlock, 0, c->stream>>>(p);
            c->kernel_launches++;
            break;
        }
        if (whitted) {  // FWhittedIntegrator (integrator.cc:115-220): every vertex is shaded, only specular lobes continue
            {
                StageTimer t(c, 2);
                k_logic<true><<<c->grid_logic_w, kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 1) k_shade<0, true><<<c->grid_shade_w[0], kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 2) k_shade<1, true><<<c->grid_shade_w[1], kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 4) k_shade<2, true><<<c->grid_shade_w[2], kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 8) k_shade<3, true><<<c->grid_shade_w[3], kBlock, 0, c->stream>>>(p, it);
                c->kernel_launches += 1 + __builtin_popcount(c->kinds_present);
            }
            StageTimer t(c, 3);
            if (count) k_connect<true, 5><<<c->grid_connect_c, kBlock, 0, c->stream>>>(p, it);
            else if (qn) k_connect<false, kTravMinBlocks, true><<<c->grid_connect6, kBlock, 0, c->stream>>>(p, it);
            else if (fast6) k_connect<false, kTravMinBlocks><<<c->grid_connect6, kBlock, 0, c->stream>>>(p, it);
            else k_connect<false, 5><<<c->grid_connect, kBlock, 0, c->stream>>>(p, it);
            c->kernel_launches++;
            continue;
        }
        {
            StageTimer t(c, 2);
            k_logic<false><<<c->grid_logic, kBlock, 0, c->stream>>>(p, it);
            if (it < c->n_iters - 1 || c->hs.has_null_material) {  // at bounce == maxDepth nothing is left to shade
                if (c->kinds_present & 1) k_shade<0><<<c->grid_shade[0], kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 2) k_shade<1><<<c->grid_shade[1], kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 4) k_shade<2><<<c->grid_shade[2], kBlock, 0, c->stream>>>(p, it);
                if (c->kinds_present & 8) k_shade<3><<<c->grid_shade[3], kBlock, 0, c->stream>>>(p, it);
                c->kernel_launches += __builtin_popcount(c->kinds_present);
            }
            c->kernel_launches++;
        }
        if (it < c->n_iters - 1 || c->hs.has_null_material) {  // no NEE at bounce == maxDepth (integrator.cc:340)
            StageTimer t(c, 3);
            if (count) k_connect<true, 5><<<c->grid_connect_c, kBlock, 0, c->stream>>>(p, it);
            else if (qn) k_connect<false, kTravMinBlocks, true><<<c->grid_connect6, kBlock, 0, c->stream>>>(p, it);
            else if (fast6) k_connect<false, kTravMinBlocks><<<c->grid_connect6, kBlock, 0, c->stream>>>(p, it);
            else k_connect<false, 5><<<c->grid_connect, kBlock, 0, c->stream>>>(p, it);
            c->kernel_launches++;
        }
    }
    if (c->hs.has_null_material && !debug) {
        k_count_dropped<<<1, 32, 0, c->stream>>>(p, c->n_iters);
        c->kernel_launches++;
    }
    return 0;
}

int jpbrt_render_pass(jpbrt_ctx* c, int sample_begin, int sample_count, uint64_t seed) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    if (sample_begin < 0 || sample_count < 0 || (long long)sample_begin + sample_count > 0xffffff)
        return set_error(c, JPBRT_ERR_INVALID, "sample range [%d, %d) outside [0, 2^24)", sample_begin, sample_begin + sample_count);
    CU_CHECK(c, cudaSetDevice(c->device));
    // Whitted numbers the vertices of its ray tree in 32 bits (children 3n+1 / 3n+3 key the sampler): past depth 20 the ids wrap
    // and sampler streams would correlate.  Refuse instead of rendering a subtly wrong image.
    if (c->opt_integrator == JPBRT_INTEGRATOR_WHITTED && c->has_mirror && c->hs.max_depth > 20)
        return set_error(c, JPBRT_ERR_UNSUPPORTED, "Whitted integrator with mirrors supports max_depth <= 20 (scene has %d)", c->hs.max_depth);
    // Whitted traces a mirror vertex twice (bsdf.h:282): leave room for the ray tree, x2 per mirror bounce, at most x8
    const int tree_growth = (c->opt_integrator == JPBRT_INTEGRATOR_WHITTED && c->has_mirror) ? 1 << std::min(3, std::max(0, c->hs.max_depth - 1)) : 1;
    c->film_reduced = false;
    int rc = ensure_pool(c, (long long)c->hs.width * c->hs.height * std::max(1, sample_count) * tree_growth);
    if (rc == 0) rc = ensure_sort_buffers(c);
    if (rc != 0) return rc;
    const long long npix = (long long)c->hs.width * c->hs.height;
    // Bands: a wavefront covers a contiguous range of the Morton pixel order (a compact region of the frame) times as
    // many samples as the pool holds, not the whole frame times a few samples.  Every radiance contribution is a float
    // atomic on the film; a band's film (<= 12 MB) stays in the 126 MB L2 next to the streamed wavefront state, a 4K
    // frame's (99.5 MB) does not -- measured on the 3840x2160 Cornell config: k_connect 222 -> see DESIGN.md.
    const long long band_target = c->opt_band_pixels > 0 ? c->opt_band_pixels : (1ll << 20);
    const int n_bands = (int)std::max(1ll, (npix + band_target - 1) / band_target);
    const long long band_pixels = (npix + n_bands - 1) / n_bands;
    int chunk_max = (int)std::max(1ll, c->paths_in_flight / band_pixels);
    chunk_max = std::max(1, chunk_max / tree_growth);
    // equal wavefronts: 50 spp with room for 32 run as 25 + 25, not 32 + 18 (short wavefronts are less efficient)
    const int n_waves = (sample_count + chunk_max - 1) / std::max(1, chunk_max);
    const int chunk = n_waves > 0 ? (sample_count + n_waves - 1) / n_waves : chunk_max;
    const bool count = c->opt_count_traversal;
    // Graph replay needs a launch sequence that never changes: no per-launch events, no counting variant.
    const bool use_graph = c->opt_use_graph && !c->opt_stage_timing && !count;
    const long long graph_key = (((((long long)c->opt_integrator * 16 + c->opt_trav_blocks) * 64 + (c->opt_refill_min + 1)) * 64 + (c->opt_min_inner + 1)) * 64 + c->opt_sort_rays) * 4 + (c->opt_node_format + 1);
    if (use_graph && (c->wave_graph == nullptr || c->wave_graph_key != graph_key)) {
        if (c->wave_graph) { cudaGraphExecDestroy(c->wave_graph); c->wave_graph = nullptr; }
        cudaGraph_t graph = nullptr;
        CU_CHECK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const unsigned long long before = c->kernel_launches;
        int qrc = queue_wavefront(c, false);
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        c->wave_graph_launches = c->kernel_launches - before;
        c->kernel_launches = before;
        if (qrc != 0) { if (graph) cudaGraphDestroy(graph); return qrc; }
        if (ce != cudaSuccess) return set_error(c, JPBRT_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&c->wave_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { c->wave_graph = nullptr; return set_error(c, JPBRT_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce)); }
        c->wave_graph_key = graph_key;
    }
    for (int band = 0; band < n_bands; ++band) {
        const long long pix0 = band * band_pixels, pix1 = std::min(npix, pix0 + band_pixels);
        if (pix1 <= pix0) break;
        for (int done = 0; done < sample_count;) {
            const int spp = std::min(chunk, sample_count - done);
            PassArgs a;
            a.k0 = (uint32_t)seed;
            a.k1 = (uint32_t)(seed >> 32);
            a.sample_begin = sample_begin + done;
            a.pixel_begin = (int)pix0;
            a.pixel_count = (int)(pix1 - pix0);
            a.n_paths = (int)((pix1 - pix0) * spp);
            k_set_args<<<1, 1, 0, c->stream>>>(c->pass_args.ptr, a);
            c->kernel_launches++;
            if (use_graph) {
                CU_CHECK(c, cudaGraphLaunch(c->wave_graph, c->stream));
                c->kernel_launches += c->wave_graph_launches;
            } else {
                int qrc = queue_wavefront(c, count);
                if (qrc != 0) return qrc;
            }
            CU_CHECK(c, cudaGetLastError());
            done += spp;
        }
    }
    return 0;
}

int jpbrt_finalize_film_device(jpbrt_ctx* c, void* out_device, int spp_total) {
    if (!c || !out_device || spp_total <= 0) return set_error(c, JPBRT_ERR_INVALID, "bad argument to finalize");
    CU_CHECK(c, cudaSetDevice(c->device));
    StageTimer t(c, 4);
    const float ratio = 1.0f / (float)spp_total;  // integrator.cc:89
    k_finalize<<<c->grid_finalize, kBlock, 0, c->stream>>>(c->film.ptr, (float*)out_device, c->film.count, ratio);
    c->kernel_launches++;
    CU_CHECK(c, cudaGetLastError());
    return 0;
}

// The ONE collective of the path: the raw float32 film sums of all ranks are added onto rank 0 with a single
// ncclReduce over NVLink, queued on the context's stream right behind the rank's last wavefront.  The other ranks'
// films are cleared afterwards, so that "sum over ranks = total" keeps holding if more passes follow.
static int reduce_film(jpbrt_ctx* c, bool grouped = false) {
    if (c->comm_size <= 1 || c->film_reduced) return 0;
    NcclApi& n = nccl();
    {
        StageTimer t(c, 5);
        ncclResult_t r = n.Reduce(c->film.ptr, c->film.ptr, c->film.count, ncclFloat32, ncclSum, 0, c->comm, c->stream);
        if (r != ncclSuccess) return set_error(c, JPBRT_ERR_CUDA, "ncclReduce failed: %s", n.GetErrorString(r));
    }
    // (inside an ncclGroup the reduce is only queued at ncclGroupEnd: a memset issued here would overtake it)
    if (c->comm_rank != 0 && !grouped) CU_CHECK(c, cudaMemsetAsync(c->film.ptr, 0, c->film.count * sizeof(float), c->stream));
    c->film_reduced = true;
    return 0;
}

int jpbrt_read_film(jpbrt_ctx* c, float* rgb, int spp_total, int finalize) {
    if (!c) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    CU_CHECK(c, cudaSetDevice(c->device));
    if (c->comm_size > 1) {  // multi-GPU: the film is complete only on rank 0, after the reduce
        int rc = reduce_film(c);
        if (rc != 0) return rc;
        if (c->comm_rank != 0) {
            CU_CHECK(c, cudaStreamSynchronize(c->stream));
            drain_events(c);
            return 0;
        }
    }
    if (!rgb) return set_error(c, JPBRT_ERR_INVALID, "rgb is null");
    const size_t n = c->film.count;
    if (finalize) {
        if (spp_total <= 0) return set_error(c, JPBRT_ERR_INVALID, "spp_total must be positive");
        int rc = jpbrt_finalize_film_device(c, c->film_final.ptr, spp_total);
        if (rc != 0) return rc;
        CU_CHECK(c, cudaMemcpyAsync(rgb, c->film_final.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
    } else {
        CU_CHECK(c, cudaMemcpyAsync(rgb, c->film.ptr, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    drain_events(c);
    return 0;
}

void* jpbrt_film_device_ptr(jpbrt_ctx* c) { return c ? (void*)c->film.ptr : nullptr; }
size_t jpbrt_film_num_floats(const jpbrt_ctx* c) { return c ? c->film.count : 0; }
void* jpbrt_stream(jpbrt_ctx* c) { return c ? (void*)c->stream : nullptr; }

int jpbrt_synchronize(jpbrt_ctx* c) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    CU_CHECK(c, cudaSetDevice(c->device));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    drain_events(c);
    return 0;
}

int jpbrt_get_stats(jpbrt_ctx* c, jpbrt_stats* out) {
    if (!c || !out) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    CU_CHECK(c, cudaSetDevice(c->device));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    drain_events(c);
    unsigned long long h[ST_COUNT];
    CU_CHECK(c, cudaMemcpy(h, c->dstats.ptr, sizeof(h), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->samples = h[ST_SAMPLES];
    out->extension_rays = h[ST_EXT_RAYS];
    out->shadow_rays = h[ST_SHADOW_RAYS];
    out->shaded_vertices = h[ST_VERTICES];
    out->box_tests = h[ST_BOX];
    out->prim_tests = h[ST_PRIM];
    out->shadow_box_tests = h[ST_SH_BOX];
    out->shadow_prim_tests = h[ST_SH_PRIM];
    out->node_fetches = h[ST_NODE_FETCH];
    out->prim_fetches = h[ST_PRIM_FETCH];
    out->shadow_node_fetches = h[ST_SH_NODE_FETCH];
    out->shadow_prim_fetches = h[ST_SH_PRIM_FETCH];
    out->node_bytes = (c->opt_trav_blocks >= 6 && c->hs.bvh_depth <= kMaxBvhDepth && use_qnodes(c)) ? 32 : 64;
    out->invalid_contributions = h[ST_INVALID] + h[ST_DROPPED] + h[ST_STACK_DROPPED] + h[ST_NEE_DROPPED];
    out->dropped_rays = h[ST_DROPPED];
    out->stack_overflows = h[ST_STACK_DROPPED];
    out->nee_dropped = h[ST_NEE_DROPPED];
    out->bvh_depth = (uint64_t)c->hs.bvh_depth;
    out->kernel_launches = c->kernel_launches;
    out->ms_generate = c->ms_stage[0];
    out->ms_extend = c->ms_stage[1];
    out->ms_shade = c->ms_stage[2];
    out->ms_connect = c->ms_stage[3];
    out->ms_finalize = c->ms_stage[4];
    out->ms_reduce = c->ms_stage[5];
    out->n_nodes = c->dsc.n_nodes;
    out->n_prim_slots = c->dsc.n_slots;
    out->scene_bytes = c->hs.Bytes();
    out->bvh_build_seconds = c->hs.bvh_build_seconds;
    out->bvh_builder = (uint64_t)c->hs.bvh_builder;
    out->bvh_device_seconds = c->hs.bvh_device_seconds;
    out->paths_in_flight = (uint64_t)c->paths_in_flight;
    return 0;
}

void jpbrt_destroy(jpbrt_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& ev : c->pending_events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto& ev : c->free_events) cudaEventDestroy(ev);
    if (c->wave_graph) cudaGraphExecDestroy(c->wave_graph);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (void* p : c->pinned) cudaHostUnregister(p);
    delete c;
}

int jpbrt_render(const jpbrt_scene_desc* desc, int spp, uint64_t seed, int device, float* rgb, double* seconds_out) {
    return jpbrt_render_integrator(desc, JPBRT_INTEGRATOR_PATH, spp, seed, device, rgb, seconds_out);
}

int jpbrt_render_integrator(const jpbrt_scene_desc* desc, int integrator, int spp, uint64_t seed, int device, float* rgb,
                            double* seconds_out) {
    if (spp <= 0 || !rgb) return set_error(nullptr, JPBRT_ERR_INVALID, "spp must be positive and rgb non-null");
    jpbrt_ctx* c = nullptr;
    int rc = jpbrt_upload_scene(desc, device, &c);
    if (rc != 0) return rc;
    rc = jpbrt_set_option(c, "integrator", integrator);
    if (rc != 0) { g_last_error = c->error; jpbrt_destroy(c); return rc; }
    auto t0 = std::chrono::steady_clock::now();
    rc = jpbrt_render_pass(c, 0, spp, seed);
    if (rc == 0) rc = jpbrt_read_film(c, rgb, spp, 1);
    auto t1 = std::chrono::steady_clock::now();
    if (rc == 0) {  // the image is in rgb; but a one-shot caller never looks at the stats, so lost rays must not pass silently
        jpbrt_stats st;
        if (jpbrt_get_stats(c, &st) == 0 && (st.dropped_rays || st.stack_overflows || st.nee_dropped))
            rc = set_error(c, JPBRT_ERR_UNSUPPORTED, "render completed but lost work: %llu rays dropped (ray tree outgrew the path pool), %llu traversal-stack "
                           "overflows, %llu light samples without a shadow slot -- the image is too dark; use passes of fewer samples or a larger paths_in_flight",
                           (unsigned long long)st.dropped_rays, (unsigned long long)st.stack_overflows, (unsigned long long)st.nee_dropped);
    }
    if (rc != 0) g_last_error = c->error;
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    jpbrt_destroy(c);
    return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// Multi-GPU: scene replicated, samples partitioned, one NCCL reduce of the film.
// ------------------------------------------------------------------------------------------------------------------
int jpbrt_comm_unique_id(void* id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId)) return set_error(nullptr, JPBRT_ERR_INVALID, "id buffer must hold JPBRT_COMM_ID_BYTES bytes");
    NcclApi& n = nccl();
    if (!n.Load()) return set_error(nullptr, JPBRT_ERR_UNSUPPORTED, "%s", n.error.c_str());
    ncclUniqueId uid;
    ncclResult_t r = n.GetUniqueId(&uid);
    if (r != ncclSuccess) return set_error(nullptr, JPBRT_ERR_CUDA, "ncclGetUniqueId failed: %s", n.GetErrorString(r));
    memset(id, 0, bytes);
    memcpy(id, &uid, sizeof(uid));
    return 0;
}

int jpbrt_comm_init(jpbrt_ctx* c, const void* id, size_t bytes, int rank, int nranks) {
    if (!c || !id || bytes < sizeof(ncclUniqueId)) return set_error(c, JPBRT_ERR_INVALID, "null argument / short id");
    if (nranks < 1 || rank < 0 || rank >= nranks) return set_error(c, JPBRT_ERR_INVALID, "rank %d outside [0, %d)", rank, nranks);
    if (c->comm) return set_error(c, JPBRT_ERR_INVALID, "the context already has a communicator");
    NcclApi& n = nccl();
    if (!n.Load()) return set_error(c, JPBRT_ERR_UNSUPPORTED, "%s", n.error.c_str());
    CU_CHECK(c, cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclResult_t r = n.CommInitRank(&c->comm, nranks, uid, rank);
    if (r != ncclSuccess) { c->comm = nullptr; return set_error(c, JPBRT_ERR_CUDA, "ncclCommInitRank failed: %s", n.GetErrorString(r)); }
    c->comm_rank = rank;
    c->comm_size = nranks;
    return 0;
}

int jpbrt_comm_rank(const jpbrt_ctx* c) { return c ? c->comm_rank : 0; }
int jpbrt_comm_size(const jpbrt_ctx* c) { return c ? c->comm_size : 1; }

int jpbrt_reduce_film(jpbrt_ctx* c) {
    if (!c) return set_error(nullptr, JPBRT_ERR_INVALID, "ctx is null");
    CU_CHECK(c, cudaSetDevice(c->device));
    return reduce_film(c);
}

void jpbrt_sample_partition(int spp_total, int rank, int nranks, int* begin, int* count) {
    // as even as possible, contiguous: rank r renders [begin, begin + count) of every pixel
    const int base = nranks > 0 ? spp_total / nranks : spp_total, rem = nranks > 0 ? spp_total % nranks : 0;
    if (begin) *begin = rank * base + std::min(rank, rem);
    if (count) *count = base + (rank < rem ? 1 : 0);
}

// One process, `ngpus` devices: FIntegrator::Render(scene, sampler, film, numthreads) with GPUs for threads.
int jpbrt_render_multi(const jpbrt_scene_desc* desc, int integrator, int spp, uint64_t seed, int ngpus, float* rgb, double* seconds_out,
                       double* reduce_ms_out) {
    if (spp <= 0 || !rgb || ngpus < 1) return set_error(nullptr, JPBRT_ERR_INVALID, "spp and ngpus must be positive and rgb non-null");
    if (ngpus == 1) {
        if (reduce_ms_out) *reduce_ms_out = 0;
        return jpbrt_render_integrator(desc, integrator, spp, seed, 0, rgb, seconds_out);
    }
    if (ngpus > jpbrt_device_count()) return set_error(nullptr, JPBRT_ERR_INVALID, "%d GPUs requested, %d visible", ngpus, jpbrt_device_count());
    NcclApi& n = nccl();
    if (!n.Load()) return set_error(nullptr, JPBRT_ERR_UNSUPPORTED, "%s", n.error.c_str());
    // flatten (and build the BVH) once; every device gets a copy
    HostScene hs;
    std::string err;
    int rc = FlattenScene(desc, &hs, &err);
    if (rc != 0) return set_error(nullptr, rc, "%s", err.c_str());
    std::vector<jpbrt_ctx*> ctx(ngpus, nullptr);
    auto cleanup = [&]() { for (jpbrt_ctx* c : ctx) jpbrt_destroy(c); };
    for (int d = 0; d < ngpus; ++d) {
        jpbrt_ctx* c = new jpbrt_ctx();
        c->hs = hs;
        rc = finish_upload(c, d);
        if (rc != 0) { cleanup(); return rc; }
        ctx[d] = c;
        if ((rc = jpbrt_set_option(c, "integrator", integrator)) != 0) { g_last_error = c->error; cleanup(); return rc; }
    }
    std::vector<int> devs(ngpus);
    std::vector<ncclComm_t> comms(ngpus);
    for (int d = 0; d < ngpus; ++d) devs[d] = d;
    ncclResult_t r = n.CommInitAll(comms.data(), ngpus, devs.data());
    if (r != ncclSuccess) { cleanup(); return set_error(nullptr, JPBRT_ERR_CUDA, "ncclCommInitAll failed: %s", n.GetErrorString(r)); }
    for (int d = 0; d < ngpus; ++d) { ctx[d]->comm = comms[d]; ctx[d]->comm_rank = d; ctx[d]->comm_size = ngpus; }
    auto t0 = std::chrono::steady_clock::now();
    for (int d = 0; d < ngpus && rc == 0; ++d) {  // asynchronous: every device works on its own stream
        int begin, count;
        jpbrt_sample_partition(spp, d, ngpus, &begin, &count);
        if (count > 0) rc = jpbrt_render_pass(ctx[d], begin, count, seed);
        if (rc != 0) g_last_error = ctx[d]->error;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (rc == 0) {
        cudaSetDevice(0);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx[0]->stream);
        n.GroupStart();  // one thread drives several communicators: the calls must be grouped
        for (int d = 0; d < ngpus && rc == 0; ++d) {
            cudaSetDevice(d);
            rc = reduce_film(ctx[d], true);
            if (rc != 0) g_last_error = ctx[d]->error;
        }
        n.GroupEnd();
        cudaSetDevice(0);
        cudaEventRecord(e1, ctx[0]->stream);
    }
    if (rc == 0) {
        for (int d = 1; d < ngpus; ++d) jpbrt_synchronize(ctx[d]);
        rc = jpbrt_read_film(ctx[0], rgb, spp, 1);
        if (rc != 0) g_last_error = ctx[0]->error;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (reduce_ms_out) {
        float ms = 0;
        if (rc == 0 && e0 && e1) { cudaSetDevice(0); cudaEventElapsedTime(&ms, e0, e1); }
        *reduce_ms_out = ms;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cleanup();
    return rc;
}

int jpbrt_scene_info(jpbrt_ctx* c, float* out7) {
    if (!c || !out7) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    for (int a = 0; a < 3; ++a) { out7[a] = c->hs.world_min[a]; out7[3 + a] = c->hs.world_max[a]; }
    out7[6] = c->hs.world_radius;
    return 0;
}

}  // extern "C"

// =================================================================================================
// Unit kernels: the stage device functions on caller-supplied arrays (parity tests).
// =================================================================================================
namespace {

__global__ void k_unit_intersect_shape(DevScene sc, int n, const float* rays8, int* hit, float* t, float* pos3, float* nrm3) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* q = rays8 + 8 * i;
        const f3 o = mk3(q[0], q[1], q[2]), d = mk3(q[3], q[4], q[5]);
        float tmax = q[7];
        bool h = intersect_slot(sc.slots, sc.slot_nrm, o, d, q[6], tmax);
        hit[i] = h ? 1 : 0;
        t[i] = h ? tmax : 0.f;
        f3 P = mk3(0, 0, 0), N = mk3(0, 0, 0);
        if (h) { P = o + tmax * d; N = hit_normal(sc, 0, P, d); }
        st3(pos3, i, P);
        st3(nrm3, i, N);
    }
}

// The unit entry points run the SAME warp-cooperative traversal as k_extend / k_connect.
struct UnitRayIO {
    const float* rays8;
    int* prim;
    float* t;
    float* pos3;
    float* nrm3;
    const DevScene* sc;
    __device__ __forceinline__ bool load(int i, f3& o, f3& d, float& tmin, float& tmax) const {
        const float* q = rays8 + 8 * (size_t)i;
        o = mk3(q[0], q[1], q[2]);
        d = mk3(q[3], q[4], q[5]);
        tmin = q[6];
        tmax = q[7];
        return true;
    }
    __device__ __forceinline__ void store(int i, int slot, float tmax) const {
        const float* q = rays8 + 8 * (size_t)i;
        const f3 o = mk3(q[0], q[1], q[2]), d = mk3(q[3], q[4], q[5]);
        f3 P = mk3(0, 0, 0), N = mk3(0, 0, 0);
        int pi = -1;
        if (slot >= 0) {
            P = o + tmax * d;
            N = hit_normal(*sc, slot, P, d);
            pi = __float_as_int(ldg4(sc->slot_nrm + slot).w) >> kTypeBits;
        }
        prim[i] = pi;
        t[i] = slot >= 0 ? tmax : 0.f;
        if (pos3) st3(pos3, i, P);
        if (nrm3) st3(nrm3, i, N);
    }
};

__global__ void __launch_bounds__(kBlock) k_unit_scene_intersect(DevScene sc, int n, int* work, const float* rays8, int* prim, float* t, float* pos3, float* nrm3) {
    TravCounts cnt;
    UnitRayIO io{rays8, prim, t, pos3, nrm3, &sc};
    traverse_queue<false, false>(sc, n, work, io, 8, 8, cnt);
}

struct UnitOccIO {
    const float* pos3;
    const float* target3;
    int* occ;
    __device__ __forceinline__ bool load(int i, f3& o, f3& d, float& tmin, float& tmax) const {
        const f3 P = ld3(pos3, i), T = ld3(target3, i);
        const f3 v = T - P;  // scene.h:44-47: Normalize(target - pos), Distance(pos, target)
        const float dist = length(v);
        o = P;
        d = v / dist;
        tmin = JPBRT_RAY_TMIN;
        tmax = dist - 0.001f;
        return true;
    }
    __device__ __forceinline__ void store(int i, int slot, float) const { occ[i] = slot >= 0 ? 1 : 0; }
};

__global__ void __launch_bounds__(kBlock) k_unit_scene_occluded(DevScene sc, int n, int* work, const float* pos3, const float* target3, int* occ) {
    TravCounts cnt;
    UnitOccIO io{pos3, target3, occ};
    traverse_queue<true, false>(sc, n, work, io, 8, 8, cnt);
}

__global__ void k_unit_bsdf_ex(jpbrt_bsdf_desc desc, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2,
                               float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags) {
    const BsdfEx b = make_bsdf_ex(desc);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Frame fr = make_frame(ld3(nrm3, i));
        const f3 wo = to_local(fr, ld3(wo3, i)), wi = to_local(fr, ld3(wi3, i));
        st3(f_eval3, i, bsdf_ex_eval_local(b, wo, wi));
        pdf_eval[i] = bsdf_ex_pdf_local(b, wo, wi);
        BsdfSample s = bsdf_ex_sample_local(b, fr, wo, u2[2 * i], u2[2 * i + 1]);
        st3(s_wi3, i, to_world(fr, s.wi));
        st3(s_f3, i, s.f);
        s_pdf[i] = s.pdf;
        s_flags[i] = s.flags;
    }
}

__global__ void k_unit_emitted(DevScene sc, int n, const int* prim, const float* nrm3, const float* wo3, float* Le3) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        f3 Le = mk3(0, 0, 0);
        int pi = prim[i];
        if (pi >= 0 && pi < sc.n_prims) {
            int slot = sc.prim_slot[pi];
            Le = emitted(sc, sc.slot_ml[slot].y, ld3(nrm3, i), ld3(wo3, i));
        }
        st3(Le3, i, Le);
    }
}

__global__ void k_unit_generate_rays(DevScene sc, int n, const float* posfilm2, float* o3, float* d3) {
    const DevCamera& cam = sc.cam;
    const f3 pos = mk3(cam.pos[0], cam.pos[1], cam.pos[2]);
    const f3 front = mk3(cam.front[0], cam.front[1], cam.front[2]);
    const f3 right = mk3(cam.right[0], cam.right[1], cam.right[2]);
    const f3 up = mk3(cam.up[0], cam.up[1], cam.up[2]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float fx = posfilm2[2 * i], fy = posfilm2[2 * i + 1];
        const f3 dir = front + right * (fx / cam.res_x - 0.5f) + up * (0.5f - fy / cam.res_y);
        st3(o3, i, pos);
        st3(d3, i, normalize(dir));
    }
}

__global__ void k_unit_rng_block(int n, const uint32_t* pixel, const uint32_t* sample, const uint32_t* block, RngKey key, float* out4) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 r = rng_block(key, pixel[i], sample[i], block[i]);
        out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
    }
}

__global__ void k_unit_philox_raw(int n, const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(ctr4[4 * i], ctr4[4 * i + 1], ctr4[4 * i + 2], ctr4[4 * i + 3], key2[2 * i], key2[2 * i + 1]);
        out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
    }
}

// host <-> device marshalling for the unit entry points
struct Arena {
    std::vector<void*> ptrs;
    cudaError_t err = cudaSuccess;
    template <typename T>
    T* In(const T* host, size_t n) {
        if (!host || n == 0) return nullptr;
        T* d = Out<T>(n);
        if (d && err == cudaSuccess) err = cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice);
        return d;
    }
    template <typename T>
    T* Out(size_t n) {
        T* d = nullptr;
        if (err != cudaSuccess) return nullptr;
        err = cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T));
        if (err == cudaSuccess) ptrs.push_back(d); else d = nullptr;
        return d;
    }
    template <typename T>
    void Back(T* host, const T* dev, size_t n) {
        if (host && dev && err == cudaSuccess) err = cudaMemcpy(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost);
    }
    ~Arena() { for (void* p : ptrs) cudaFree(p); }
};

int unit_grid(int n) { return std::max(1, std::min((n + kBlock - 1) / kBlock, 148 * 8)); }

int finish_unit(jpbrt_ctx* c, Arena& a) {
    if (a.err == cudaSuccess) a.err = cudaGetLastError();
    if (a.err == cudaSuccess) a.err = cudaDeviceSynchronize();
    if (a.err != cudaSuccess) return set_error(c, JPBRT_ERR_CUDA, "unit kernel failed: %s", cudaGetErrorString(a.err));
    return 0;
}

}  // namespace

extern "C" {

int jpbrt_unit_intersect_shape(const jpbrt_shape* shape, int device, int n, const float* rays8, int* hit, float* t, float* pos3, float* nrm3) {
    if (!shape || n < 0 || !rays8 || !hit || !t || !pos3 || !nrm3) return set_error(nullptr, JPBRT_ERR_INVALID, "null argument");
    Float4 q[4], nr;
    float bmn[3], bmx[3];
    if (!MakeSlot(*shape, 0, q, &nr, bmn, bmx)) return set_error(nullptr, JPBRT_ERR_INVALID, "unknown shape type");
    int rc = select_device(nullptr, device);
    if (rc != 0) return rc;
    Arena a;
    DevScene sc{};
    sc.slots = a.In(q, 4);
    sc.slot_nrm = a.In(&nr, 1);
    sc.n_slots = 1;
    const float* d_rays = a.In(rays8, (size_t)n * 8);
    int* d_hit = a.Out<int>(n);
    float* d_t = a.Out<float>(n);
    float* d_pos = a.Out<float>((size_t)n * 3);
    float* d_nrm = a.Out<float>((size_t)n * 3);
    if (a.err == cudaSuccess && n > 0) k_unit_intersect_shape<<<unit_grid(n), kBlock>>>(sc, n, d_rays, d_hit, d_t, d_pos, d_nrm);
    rc = finish_unit(nullptr, a);
    if (rc != 0) return rc;
    a.Back(hit, d_hit, n); a.Back(t, d_t, n); a.Back(pos3, d_pos, (size_t)n * 3); a.Back(nrm3, d_nrm, (size_t)n * 3);
    return a.err == cudaSuccess ? 0 : set_error(nullptr, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_scene_intersect(jpbrt_ctx* c, int n, const float* rays8, int* prim, float* t, float* pos3, float* nrm3) {
    if (!c || n < 0 || !rays8 || !prim || !t) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    CU_CHECK(c, cudaSetDevice(c->device));
    Arena a;
    const float* d_rays = a.In(rays8, (size_t)n * 8);
    int* d_prim = a.Out<int>(n);
    float* d_t = a.Out<float>(n);
    float* d_pos = pos3 ? a.Out<float>((size_t)n * 3) : nullptr;
    float* d_nrm = nrm3 ? a.Out<float>((size_t)n * 3) : nullptr;
    int* d_work = a.Out<int>(1);
    if (a.err == cudaSuccess) a.err = cudaMemset(d_work, 0, sizeof(int));
    if (a.err == cudaSuccess && n > 0) k_unit_scene_intersect<<<unit_grid(n), kBlock>>>(c->dsc, n, d_work, d_rays, d_prim, d_t, d_pos, d_nrm);
    int rc = finish_unit(c, a);
    if (rc != 0) return rc;
    a.Back(prim, d_prim, n); a.Back(t, d_t, n); a.Back(pos3, d_pos, (size_t)n * 3); a.Back(nrm3, d_nrm, (size_t)n * 3);
    return a.err == cudaSuccess ? 0 : set_error(c, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_scene_occluded(jpbrt_ctx* c, int n, const float* pos3, const float* target3, int* occluded) {
    if (!c || n < 0 || !pos3 || !target3 || !occluded) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    CU_CHECK(c, cudaSetDevice(c->device));
    Arena a;
    const float* d_pos = a.In(pos3, (size_t)n * 3);
    const float* d_tgt = a.In(target3, (size_t)n * 3);
    int* d_occ = a.Out<int>(n);
    int* d_work = a.Out<int>(1);
    if (a.err == cudaSuccess) a.err = cudaMemset(d_work, 0, sizeof(int));
    if (a.err == cudaSuccess && n > 0) k_unit_scene_occluded<<<unit_grid(n), kBlock>>>(c->dsc, n, d_work, d_pos, d_tgt, d_occ);
    int rc = finish_unit(c, a);
    if (rc != 0) return rc;
    a.Back(occluded, d_occ, n);
    return a.err == cudaSuccess ? 0 : set_error(c, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_bsdf(const jpbrt_material* mat, int device, int n, const float* nrm3, const float* wo3, const float* wi3,
                    const float* u2, const float* ulobe, float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3,
                    float* s_pdf, int* s_flags, int* is_delta) {
    if (!mat || n < 0 || !nrm3 || !wo3 || !wi3 || !u2) return set_error(nullptr, JPBRT_ERR_INVALID, "null argument");
    Float4 m[3];
    if (!MakeMaterial(*mat, m)) return set_error(nullptr, JPBRT_ERR_INVALID, "unknown material type");
    int rc = select_device(nullptr, device);
    if (rc != 0) return rc;
    Arena a;
    const Float4* d_m = a.In(m, 3);
    const float *d_n = a.In(nrm3, (size_t)n * 3), *d_wo = a.In(wo3, (size_t)n * 3), *d_wi = a.In(wi3, (size_t)n * 3);
    const float *d_u = a.In(u2, (size_t)n * 2), *d_ul = a.In(ulobe, (size_t)n);
    float *d_fe = a.Out<float>((size_t)n * 3), *d_pe = a.Out<float>(n), *d_swi = a.Out<float>((size_t)n * 3);
    float *d_sf = a.Out<float>((size_t)n * 3), *d_sp = a.Out<float>(n);
    int *d_fl = a.Out<int>(n), *d_dl = a.Out<int>(n);
    if (a.err == cudaSuccess && n > 0)
        k_unit_bsdf<<<unit_grid(n), kBlock>>>(d_m, n, d_n, d_wo, d_wi, d_u, d_ul, d_fe, d_pe, d_swi, d_sf, d_sp, d_fl, d_dl);
    rc = finish_unit(nullptr, a);
    if (rc != 0) return rc;
    a.Back(f_eval3, d_fe, (size_t)n * 3); a.Back(pdf_eval, d_pe, n); a.Back(s_wi3, d_swi, (size_t)n * 3);
    a.Back(s_f3, d_sf, (size_t)n * 3); a.Back(s_pdf, d_sp, n); a.Back(s_flags, d_fl, n); a.Back(is_delta, d_dl, n);
    return a.err == cudaSuccess ? 0 : set_error(nullptr, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_bsdf_ex(const jpbrt_bsdf_desc* desc, int device, int n, const float* nrm3, const float* wo3, const float* wi3,
                       const float* u2, float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags) {
    if (!desc || n < 0 || !nrm3 || !wo3 || !wi3 || !u2) return set_error(nullptr, JPBRT_ERR_INVALID, "null argument");
    if (desc->kind < JPBRT_BSDF_PHONG || desc->kind > JPBRT_BSDF_MICROFACET_TRANSMISSION || desc->distribution < 0 ||
        desc->distribution > JPBRT_DIST_TROWBRIDGE_REITZ || desc->fresnel < 0 || desc->fresnel > JPBRT_FRESNEL_CONDUCTOR)
        return set_error(nullptr, JPBRT_ERR_INVALID, "unknown BSDF kind / distribution / Fresnel");
    int rc = select_device(nullptr, device);
    if (rc != 0) return rc;
    Arena a;
    const float *d_n = a.In(nrm3, (size_t)n * 3), *d_wo = a.In(wo3, (size_t)n * 3), *d_wi = a.In(wi3, (size_t)n * 3);
    const float* d_u = a.In(u2, (size_t)n * 2);
    float *d_fe = a.Out<float>((size_t)n * 3), *d_pe = a.Out<float>(n), *d_swi = a.Out<float>((size_t)n * 3);
    float *d_sf = a.Out<float>((size_t)n * 3), *d_sp = a.Out<float>(n);
    int* d_fl = a.Out<int>(n);
    if (a.err == cudaSuccess && n > 0)
        k_unit_bsdf_ex<<<unit_grid(n), kBlock>>>(*desc, n, d_n, d_wo, d_wi, d_u, d_fe, d_pe, d_swi, d_sf, d_sp, d_fl);
    rc = finish_unit(nullptr, a);
    if (rc != 0) return rc;
    a.Back(f_eval3, d_fe, (size_t)n * 3); a.Back(pdf_eval, d_pe, n); a.Back(s_wi3, d_swi, (size_t)n * 3);
    a.Back(s_f3, d_sf, (size_t)n * 3); a.Back(s_pdf, d_sp, n); a.Back(s_flags, d_fl, n);
    return a.err == cudaSuccess ? 0 : set_error(nullptr, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_light_sample(jpbrt_ctx* c, int light, int n, const float* pos3, const float* nrm3, const float* u2,
                            float* lpos3, float* wi3, float* pdf, float* Li3) {
    if (!c || n < 0 || !pos3 || !nrm3 || !u2) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    if (light < 0 || light >= c->dsc.n_lights) return set_error(c, JPBRT_ERR_INVALID, "light index out of range");
    CU_CHECK(c, cudaSetDevice(c->device));
    Arena a;
    const float *d_p = a.In(pos3, (size_t)n * 3), *d_n = a.In(nrm3, (size_t)n * 3), *d_u = a.In(u2, (size_t)n * 2);
    float *d_lp = a.Out<float>((size_t)n * 3), *d_wi = a.Out<float>((size_t)n * 3), *d_pdf = a.Out<float>(n), *d_li = a.Out<float>((size_t)n * 3);
    if (a.err == cudaSuccess && n > 0) k_unit_light_sample<<<unit_grid(n), kBlock>>>(c->dsc, light, n, d_p, d_n, d_u, d_lp, d_wi, d_pdf, d_li);
    int rc = finish_unit(c, a);
    if (rc != 0) return rc;
    a.Back(lpos3, d_lp, (size_t)n * 3); a.Back(wi3, d_wi, (size_t)n * 3); a.Back(pdf, d_pdf, n); a.Back(Li3, d_li, (size_t)n * 3);
    return a.err == cudaSuccess ? 0 : set_error(c, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_emitted(jpbrt_ctx* c, int n, const int* prim, const float* nrm3, const float* wo3, float* Le3) {
    if (!c || n < 0 || !prim || !nrm3 || !wo3 || !Le3) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    CU_CHECK(c, cudaSetDevice(c->device));
    Arena a;
    const int* d_prim = a.In(prim, n);
    const float *d_n = a.In(nrm3, (size_t)n * 3), *d_wo = a.In(wo3, (size_t)n * 3);
    float* d_le = a.Out<float>((size_t)n * 3);
    if (a.err == cudaSuccess && n > 0) k_unit_emitted<<<unit_grid(n), kBlock>>>(c->dsc, n, d_prim, d_n, d_wo, d_le);
    int rc = finish_unit(c, a);
    if (rc != 0) return rc;
    a.Back(Le3, d_le, (size_t)n * 3);
    return a.err == cudaSuccess ? 0 : set_error(c, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_generate_rays(jpbrt_ctx* c, int n, const float* posfilm2, float* o3, float* d3) {
    if (!c || n < 0 || !posfilm2 || !o3 || !d3) return set_error(c, JPBRT_ERR_INVALID, "null argument");
    CU_CHECK(c, cudaSetDevice(c->device));
    Arena a;
    const float* d_pf = a.In(posfilm2, (size_t)n * 2);
    float *d_o = a.Out<float>((size_t)n * 3), *d_d = a.Out<float>((size_t)n * 3);
    if (a.err == cudaSuccess && n > 0) k_unit_generate_rays<<<unit_grid(n), kBlock>>>(c->dsc, n, d_pf, d_o, d_d);
    int rc = finish_unit(c, a);
    if (rc != 0) return rc;
    a.Back(o3, d_o, (size_t)n * 3); a.Back(d3, d_d, (size_t)n * 3);
    return a.err == cudaSuccess ? 0 : set_error(c, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_rng_block(int device, int n, const uint32_t* pixel, const uint32_t* sample, const uint32_t* block, uint64_t seed, float* out4) {
    if (n < 0 || !pixel || !sample || !block || !out4) return set_error(nullptr, JPBRT_ERR_INVALID, "null argument");
    int rc = select_device(nullptr, device);
    if (rc != 0) return rc;
    Arena a;
    const uint32_t *d_p = a.In(pixel, n), *d_s = a.In(sample, n), *d_b = a.In(block, n);
    float* d_o = a.Out<float>((size_t)n * 4);
    RngKey key{(uint32_t)seed, (uint32_t)(seed >> 32)};
    if (a.err == cudaSuccess && n > 0) k_unit_rng_block<<<unit_grid(n), kBlock>>>(n, d_p, d_s, d_b, key, d_o);
    rc = finish_unit(nullptr, a);
    if (rc != 0) return rc;
    a.Back(out4, d_o, (size_t)n * 4);
    return a.err == cudaSuccess ? 0 : set_error(nullptr, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

int jpbrt_unit_philox_raw(int device, int n, const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
    if (n < 0 || !ctr4 || !key2 || !out4) return set_error(nullptr, JPBRT_ERR_INVALID, "null argument");
    int rc = select_device(nullptr, device);
    if (rc != 0) return rc;
    Arena a;
    const uint32_t *d_c = a.In(ctr4, (size_t)n * 4), *d_k = a.In(key2, (size_t)n * 2);
    uint32_t* d_o = a.Out<uint32_t>((size_t)n * 4);
    if (a.err == cudaSuccess && n > 0) k_unit_philox_raw<<<unit_grid(n), kBlock>>>(n, d_c, d_k, d_o);
    rc = finish_unit(nullptr, a);
    if (rc != 0) return rc;
    a.Back(out4, d_o, (size_t)n * 4);
    return a.err == cudaSuccess ? 0 : set_error(nullptr, JPBRT_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(a.err));
}

}  // extern "C"
