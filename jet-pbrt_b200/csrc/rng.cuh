// rng.cuh -- counter-based sampler: replaces FRNG/FRandomSampler (sampler.h:16-54,130-156).
//
// The reference draws from one sequential mt19937_64 stream per band task, which cannot be
// parallelised per sample.  Here every uniform number is a pure function of
// (seed, pixel, sample index, block) through Philox4x32-10 (Salmon et al., SC'11), so a sample's
// path does not depend on which GPU, pass or thread traces it.
//
// Dimension schedule, 4 numbers per block, B = 1 + ceil(nNee / 2) where nNee counts the lights whose
// colour is not black (a black light can never contribute, integrator.cc:362: its numbers would be drawn and thrown away):
//   block 0                       : film jitter (x, y, -, -)            integrator.cc:100
//   block 1 + b*B                 : bounce b: (lobe, bsdf.x, bsdf.y, rr) material.cc:14, integrator.cc:375,386
//   block 1 + b*B + 1 + k/2       : bounce b: the k-th non-black light's (u.x, u.y) at words 2*(k%2)  integrator.cc:361
// (Cornell: [black env, tri, tri] -> both triangle lights share ONE block: two Philox evaluations per vertex instead of three.)
#pragma once

#include <stdint.h>

#include <cuda_runtime.h>

namespace jpbrt {

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

struct RngKey {
    uint32_t k0, k1;
};

// `branch` (4th counter word) is 0 for a path; the Whitted mode numbers the vertices of its ray TREE with it
// (children of node n: 3n+1 SpecularReflect, 3n+2 SpecularTransmit, 3n+3 SpecularReflectAndTransmit).
__device__ __forceinline__ float4 rng_block(const RngKey& key, uint32_t pixel, uint32_t sample, uint32_t block, uint32_t branch = 0u) {
    uint4 r = philox4x32_10(pixel, sample, block, branch, key.k0, key.k1);
    return make_float4(u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w));
}

__host__ __device__ __forceinline__ int rng_blocks_per_bounce(int n_nee_lights) { return 1 + (n_nee_lights + 1) / 2; }

}  // namespace jpbrt
