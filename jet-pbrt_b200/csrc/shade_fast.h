// shade_fast.h -- host-callable launchers of the part of the shade stage that is also built with RELAXED arithmetic
// (csrc/shade_fast.cu): k_logic and the Lambert kernel k_shade<KIND_LAMBERT>.
//
// The hot path's default build reproduces the reference's float expressions operation for operation (-fmad=false, IEEE
// division): bit-comparable values, at the price of ~25 % more instructions in k_shade<Lambert> (202 of its
// 1,534 instructions are the three IEEE divisions of every `vector / scalar`).  north_star's bar for f / pdf is 1e-5
// relative.  Measured on B200 with the WHOLE stage relaxed (2^18 inputs per material against the reference's CPU BSDFs):
// Lambert sampling / pdf, light sampling and the throughput arithmetic stay within 3e-6 -- but the microfacet and Fresnel
// expressions are cancellation-prone (1 - cos^2, tan^2, eta^2 (1 - cos^2), the visible-normal slopes): relaxed rounding
// moved their sampled directions by > 1e-5 for 1-2 % of the inputs (worst 1.7e-2) and 5 % of Cornell's pixels by > 1e-3.
// So option "shade_math" = 1 relaxes exactly the two kernels that keep the bar -- every vertex whose BSDF is a microfacet
// or a delta lobe is shaded by the exact build in both modes -- and generate / extend / connect / finalize, everything
// that decides WHICH primitive is hit, are exact in both modes as well.
// Parameters travel as untyped bytes: both translation units compile the same WfParams / DevScene definitions.
#pragma once

#include <cuda_runtime.h>

namespace jpbrt_shade_fast {

int occupancy_logic();           // resident blocks per SM of the relaxed k_logic
int occupancy_shade_lambert();   //   ... and k_shade<KIND_LAMBERT>
void launch_logic(const void* wf_params, int it, int grid, cudaStream_t stream);
void launch_shade_lambert(const void* wf_params, int it, int grid, cudaStream_t stream);
// unit kernels (parity tests of the relaxed build): same arguments as k_unit_bsdf / k_unit_light_sample.  The BSDF one is
// only meaningful for materials whose BSDF is Lambert -- the only BSDF code the relaxed build ever runs in a render.
void launch_unit_bsdf(int grid, const void* mat, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2, const float* ulobe,
                      float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags, int* is_delta);
void launch_unit_light_sample(int grid, const void* dev_scene, int light, int n, const float* pos3, const float* nrm3, const float* u2, float* lpos3,
                              float* wi3, float* pdf, float* Li3);

}  // namespace jpbrt_shade_fast
