// peaks_l2.cu -- the ceilings the traversal kernels are measured against, MEASURED on the GPU the bench runs on.
//
// MEASURED_PEAKS.json (driver-written) holds a streaming HBM copy and a bf16 GEMM.  Neither bounds a BVH walk: the
// scenes of four of the five configs are cache-resident, and every lane fetches its OWN 64-byte node or primitive
// record.  This micro-benchmark measures what the memory system can deliver for exactly that access pattern --
//   gather : every lane loads one 64-byte record (two 256-bit loads, the traversal kernel's own instruction) at an
//            independent pseudo-random index of a working set of W bytes; W from 32 KB (L1-resident per SM) through
//            4 MB / 32 MB (L2-resident) to 1 GB (DRAM);  "uniform": all 32 lanes of a warp load the SAME record
//            (the best case of lanes sharing a node) --
// plus the plain ceilings:
//   l2_stream : coalesced 128-bit loads over a 48 MB set that bypass L1 (ld.global.cg);
//   ffma      : warp-instruction issue rate of independent FFMAs (FP32 pipe), which is also the issue ceiling
//               (4 schedulers x 1 instruction per clock per SM).
// Output: ONE JSON object on stdout.  Built by `make -C jet-pbrt_b200 peaks` -> jet-pbrt_b200/build/peaks_l2;
// bench.py runs it before its timed region and quotes the numbers in `roofline`.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (x);                                                                  \
        if (e__ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int kBlock = 256;

__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// Every lane gathers `iters` x 4 independent records of 64 bytes (two 256-bit loads: the float BVH node, a triangle slot) or,
// HALF, of 32 bytes (one 256-bit load: the quantised node).  UNIFORM: the index depends on the warp only.
template <bool UNIFORM, bool HALF = false>
__global__ void __launch_bounds__(kBlock, 6) k_gather(const float4* __restrict__ recs, unsigned mask, int iters, float* sink) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned s = hash32((UNIFORM ? (tid >> 5) : tid) * 2654435761u + 12345u);
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        float4 a[4], b[4], c[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s = s * 1664525u + 1013904223u;
            const float4* p = recs + (size_t)((s >> 7) & mask) * (HALF ? 2 : 4);
            ldg256(p, a[u], b[u]);
            if (!HALF) ldg256(p + 2, c[u], d[u]);
            else c[u] = d[u] = a[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += a[u].x + b[u].y + c[u].z + d[u].w;
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void __launch_bounds__(kBlock, 6) k_l2_stream(const float4* __restrict__ buf, size_t n4, int reps, float* sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldcg(buf + i);
            acc += v.x + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}

__global__ void __launch_bounds__(kBlock, 6) k_ffma(int iters, float* sink, float seed) {
    float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const float m = 1.0000001f, c = 1e-9f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
            a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
        }
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f) *sink = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();  // warm-up
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int main(int argc, char** argv) {
    int dev = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
    const int grid = sms * 6;
    float* sink;
    CK(cudaMalloc(&sink, 4));
    const size_t max_bytes = (size_t)1 << 30;
    float4* buf;
    CK(cudaMalloc(&buf, max_bytes));
    {  // fill with something that is not all zeros
        std::vector<float> h(1 << 20, 1.5f);
        for (size_t off = 0; off < max_bytes; off += h.size() * 4) CK(cudaMemcpy((char*)buf + off, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %.0f", prop.name, sms, clock_khz / 1e3);

    // ---- gather of 64-byte records
    const size_t sizes[] = {(size_t)32 << 10, (size_t)256 << 10, (size_t)1 << 20, (size_t)4 << 20, (size_t)32 << 20, (size_t)256 << 20, (size_t)1 << 30};
    printf(", \"gather64\": [");
    bool first = true;
    for (size_t w : sizes) {
        const unsigned mask = (unsigned)(w / 64 - 1);
        const int iters = w <= ((size_t)32 << 20) ? 64 : 16;
        const double recs = (double)grid * kBlock * iters * 4;
        const double ms_d = time_ms([&] { k_gather<false><<<grid, kBlock>>>(buf, mask, iters, sink); }, 5);
        const double ms_u = time_ms([&] { k_gather<true><<<grid, kBlock>>>(buf, mask, iters, sink); }, 5);
        // the same working set as 32-byte records (twice as many of them)
        const double ms_h = time_ms([&] { k_gather<false, true><<<grid, kBlock>>>(buf, (unsigned)(w / 32 - 1), iters, sink); }, 5);
        printf("%s{\"working_set_bytes\": %zu, \"divergent_gbs\": %.1f, \"divergent_grecords_s\": %.2f, \"uniform_gbs\": %.1f, \"divergent32_gbs\": %.1f, \"divergent32_grecords_s\": %.2f}", first ? "" : ", ",
               w, recs * 64 / (ms_d * 1e-3) / 1e9, recs / (ms_d * 1e-3) / 1e9, recs * 64 / (ms_u * 1e-3) / 1e9, recs * 32 / (ms_h * 1e-3) / 1e9, recs / (ms_h * 1e-3) / 1e9);
        first = false;
    }
    printf("]");

    // ---- L2 streaming read (48 MB set, L1 bypassed)
    {
        const size_t bytes = (size_t)48 << 20, n4 = bytes / 16;
        const int reps = 20;
        const double ms = time_ms([&] { k_l2_stream<<<grid, kBlock>>>(buf, n4, reps, sink); }, 5);
        printf(", \"l2_stream_gbs\": %.1f", (double)bytes * reps / (ms * 1e-3) / 1e9);
    }
    // ---- HBM streaming read (1 GB set)
    {
        const size_t n4 = max_bytes / 16;
        const double ms = time_ms([&] { k_l2_stream<<<grid, kBlock>>>(buf, n4, 1, sink); }, 5);
        printf(", \"hbm_stream_read_gbs\": %.1f", (double)max_bytes / (ms * 1e-3) / 1e9);
    }
    // ---- FP32 issue
    {
        const int iters = 2048;
        const double ms = time_ms([&] { k_ffma<<<grid, kBlock>>>(iters, sink, 1.0f); }, 5);
        const double warp_inst = (double)grid * (kBlock / 32) * iters * 16 * 8;
        printf(", \"ffma_warp_ginst_s\": %.1f, \"ffma_tflops\": %.2f, \"issue_peak_warp_ginst_s_at_max_clock\": %.1f", warp_inst / (ms * 1e-3) / 1e9,
               warp_inst * 64 / (ms * 1e-3) / 1e12, sms * 4.0 * clock_khz * 1e3 / 1e9);
    }
    printf("}\n");
    cudaFree(buf);
    cudaFree(sink);
    return 0;
}
