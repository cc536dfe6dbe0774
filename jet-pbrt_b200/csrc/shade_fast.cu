// shade_fast.cu -- k_logic and the LAMBERT shade kernel built a second time with relaxed arithmetic: FMA contraction on,
// division and square root by reciprocal approximation (nvcc -fmad=true -prec-div=false -prec-sqrt=false, set for THIS
// file by the Makefile).  See shade_fast.h for why only these two.
//
// The shared headers put everything in `namespace jpbrt`; this translation unit renames that namespace so that its
// kernels and device functions are distinct symbols from the exact build's in csrc/c_api.cu.
#include "../../include/jetpbrt_scene.h"  // (constants; wavefront.cuh gets them through its includer in c_api.cu)
#define jpbrt jpbrt_fast_impl
#include "wavefront.cuh"
#include "unit_shade.cuh"
#undef jpbrt

#include "shade_fast.h"

namespace jpbrt_shade_fast {

using namespace jpbrt_fast_impl;

template <typename K>
static int blocks_per_sm(K kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm <= 0) per_sm = 1;
    return per_sm;
}

int occupancy_logic() { return blocks_per_sm(k_logic<false>); }

int occupancy_shade_lambert() { return blocks_per_sm(k_shade<KIND_LAMBERT>); }

void launch_logic(const void* wf_params, int it, int grid, cudaStream_t stream) {
    k_logic<false><<<grid, kBlock, 0, stream>>>(*static_cast<const WfParams*>(wf_params), it);
}

void launch_shade_lambert(const void* wf_params, int it, int grid, cudaStream_t stream) {
    k_shade<KIND_LAMBERT><<<grid, kBlock, 0, stream>>>(*static_cast<const WfParams*>(wf_params), it);
}

void launch_unit_bsdf(int grid, const void* mat, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2, const float* ulobe,
                      float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags, int* is_delta) {
    k_unit_bsdf<<<grid, kBlock>>>(static_cast<const Float4*>(mat), n, nrm3, wo3, wi3, u2, ulobe, f_eval3, pdf_eval, s_wi3, s_f3, s_pdf, s_flags, is_delta);
}

void launch_unit_light_sample(int grid, const void* dev_scene, int light, int n, const float* pos3, const float* nrm3, const float* u2, float* lpos3,
                              float* wi3, float* pdf, float* Li3) {
    k_unit_light_sample<<<grid, kBlock>>>(*static_cast<const DevScene*>(dev_scene), light, n, pos3, nrm3, u2, lpos3, wi3, pdf, Li3);
}

}  // namespace jpbrt_shade_fast
