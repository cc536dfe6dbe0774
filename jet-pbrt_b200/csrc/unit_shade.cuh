// unit_shade.cuh -- the shade stage's device functions (BSDF build / Evalf / Pdf / Sample, light sampling) on
// caller-supplied arrays, for the parity tests (jpbrt_unit_bsdf, jpbrt_unit_light_sample in csrc/c_api.cu).
#pragma once

#include "bsdf.cuh"
#include "dev_scene.h"
#include "dmath.cuh"
#include "light.cuh"

namespace jpbrt {

__device__ __forceinline__ f3 ld3(const float* p, int i) { return mk3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void st3(float* p, int i, const f3& v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

__global__ void k_unit_bsdf(const Float4* mat, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2,
                            const float* ulobe, float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf,
                            int* s_flags, int* is_delta) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Bsdf b = make_bsdf(mat, ulobe ? ulobe[i] : 0.f);
        const Frame fr = make_frame(ld3(nrm3, i));
        const f3 wo = to_local(fr, ld3(wo3, i)), wi = to_local(fr, ld3(wi3, i));
        st3(f_eval3, i, bsdf_eval_local(b, wo, wi));
        pdf_eval[i] = bsdf_pdf_local(b, wo, wi);
        BsdfSample s = bsdf_sample_local(b, wo, u2[2 * i], u2[2 * i + 1]);
        st3(s_wi3, i, to_world(fr, s.wi));
        st3(s_f3, i, s.f);
        s_pdf[i] = s.pdf;
        s_flags[i] = s.flags;
        is_delta[i] = bsdf_is_delta(b) ? 1 : 0;
    }
}

__global__ void k_unit_light_sample(DevScene sc, int light, int n, const float* pos3, const float* nrm3, const float* u2,
                                    float* lpos3, float* wi3, float* pdf, float* Li3) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        LightSample s = sample_light(sc, light, ld3(pos3, i), ld3(nrm3, i), u2[2 * i], u2[2 * i + 1]);
        st3(lpos3, i, s.pos);
        st3(wi3, i, s.wi);
        pdf[i] = s.pdf;
        st3(Li3, i, s.Li);
    }
}

}  // namespace jpbrt
