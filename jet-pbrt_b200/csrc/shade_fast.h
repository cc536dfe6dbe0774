// shade_fast.h -- host-callable launchers of the shade stage compiled with RELAXED arithmetic (csrc/shade_fast.cu).
//
// The hot path's default build reproduces the reference's float expressions operation for operation (-fmad=false, IEEE
// division): bit-comparable BSDF values, at the price of ~25 % more instructions in k_shade.  north_star's
// bar for f / pdf is 1e-5 relative, which FMA contraction and reciprocal-multiply division meet with room to spare.
// Option "shade_math" = 1 routes k_logic / k_shade<KIND> of the path integrator through this second build; generate,
// extend, connect and finalize -- everything that decides WHICH primitive is hit -- stay exact in both modes.
// Parameters travel as untyped bytes: both translation units compile the same WfParams / DevScene definitions.
#pragma once

#include <cuda_runtime.h>

namespace jpbrt_shade_fast {

int occupancy_logic();        // resident blocks per SM of the fast k_logic
int occupancy_shade(int kind);
void launch_logic(const void* wf_params, int it, int grid, cudaStream_t stream);
void launch_shade(int kind, const void* wf_params, int it, int grid, cudaStream_t stream);
// unit kernels (parity tests of the fast build): same arguments as k_unit_bsdf / k_unit_light_sample
void launch_unit_bsdf(int grid, const void* mat, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2, const float* ulobe,
                      float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags, int* is_delta);
void launch_unit_light_sample(int grid, const void* dev_scene, int light, int n, const float* pos3, const float* nrm3, const float* u2, float* lpos3,
                              float* wi3, float* pdf, float* Li3);

}  // namespace jpbrt_shade_fast
