/* oracle_api.h -- C API implemented TWICE, with identical signatures:
 *
 *   jref_*  : oracle/ref_harness.cc -- drives the UNMODIFIED reference (compiled in place from
 *             /root/reference/src by oracle/Makefile into oracle/_ref/libjetpbrt_ref.so).
 *   jorc_*  : oracle/pt_oracle.cc   -- our CPU restatement of the same algorithm
 *             (oracle/libjetpbrt_oracle.so).  Travels to the GPU box even without _ref.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries.  The product never does.
 *
 * All arrays are caller-owned, row-major, tightly packed float32/int32.
 * "rays8" = n x 8 floats: origin.xyz, dir.xyz, min_t, max_t   (FRay, geometry.h:384-420).
 */
#ifndef JPBRT_ORACLE_API_H
#define JPBRT_ORACLE_API_H

#include "../include/jetpbrt_scene.h"

#ifndef ORC
#error "define ORC(name) to jref_##name or jorc_##name before including oracle_api.h"
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ORC(scene) ORC(scene);

/* Build the scene from the neutral description; runs the equivalent of FScene::Preprocess
 * (scene.cc:11-23): world bound, light preprocess, BVH build.  NULL on error. */
ORC(scene)* ORC(scene_create)(const jpbrt_scene_desc* desc);
void        ORC(scene_destroy)(ORC(scene)* s);

/* FShape::Intersect on one shape (shape.h:291-327 / 399-435 / 487-526 / 199-221). */
int ORC(intersect_shape)(const jpbrt_shape* shape, int n, const float* rays8,
                         int* hit, float* t, float* pos3, float* nrm3);

/* FScene::Intersect (scene.cc:25-33): closest hit through the BVH.  prim = index into
 * desc->primitives (creation order) or -1. */
int ORC(scene_intersect)(ORC(scene)* s, int n, const float* rays8,
                         int* prim, float* t, float* pos3, float* nrm3);

/* FScene::Occluded(isect, target) (scene.h:36-47). */
int ORC(scene_occluded)(ORC(scene)* s, int n, const float* pos3, const float* target3, int* occluded);

/* material->Scattering(isect{normal}, sampler{GetFloat()=ulobe}) then, on the returned BSDF,
 * Evalf(wo,wi), Pdf(wo,wi) and Sample(wo,u)  (material.h/.cc, bsdf.h:285-302).
 * All directions are WORLD space; nrm3 is the hit normal the shading frame is built from. */
int ORC(bsdf)(const jpbrt_material* mat, int n,
              const float* nrm3, const float* wo3, const float* wi3, const float* u2, const float* ulobe,
              float* f_eval3, float* pdf_eval,
              float* s_wi3, float* s_f3, float* s_pdf, int* s_flags, int* is_delta);

/* The BSDF classes no material builds (jpbrt_bsdf_desc): constructed directly with FFrame(nrm) and the described
 * distribution / Fresnel objects, then Evalf(wo,wi), Pdf(wo,wi), Sample(wo,u) as above. */
int ORC(bsdf_ex)(const jpbrt_bsdf_desc* desc, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2,
                 float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags);

/* light->Sample_Li(isect{position,normal}, u) for desc->lights[light]  (light.h). */
int ORC(light_sample)(ORC(scene)* s, int light, int n, const float* pos3, const float* nrm3, const float* u2,
                      float* lpos3, float* wi3, float* pdf, float* Li3);

/* primitive->GetLe(isect) = arealight->L(...) (primitive.h:60-63, light.h:234-238). */
int ORC(emitted)(ORC(scene)* s, int n, const int* prim, const float* nrm3, const float* wo3, float* Le3);

/* FCamera::GenerateRay (camera.h:52-58). */
int ORC(generate_rays)(ORC(scene)* s, int n, const float* posfilm2, float* o3, float* d3);

/* FIntegrator::Render with FPathIntegratorIteration(desc->max_depth) and an FRandomSampler(spp)
 * (integrator.cc:35-111,316-403; main.cc:149-156).  film = W*H*3 floats, row 0 = top, receives
 * Clamp01(mean) exactly as FFilm::AddColor does on a cleared film.  seed < 0 keeps the
 * reference's 1234.  Returns wall seconds of Render(), or < 0 on error. */
double ORC(render)(ORC(scene)* s, int spp, int numthreads, int seed, float* film);

/* Same, with the reference's other integrators (main.cc:151-154):
 * mode 0 FPathIntegratorIteration, 1 FPathIntegratorRecursive (integrator.cc:233-307),
 * 2 FWhittedIntegrator (integrator.cc:115-220), 3 FDebugIntegrator (integrator.h:44-58). */
double ORC(render_mode)(ORC(scene)* s, int mode, int spp, int numthreads, int seed, float* film);

/* Scene facts: out[0..2]=world min, [3..5]=world max, [6]=environment-light worldRadius. */
int ORC(scene_info)(ORC(scene)* s, float* out7);

#ifdef __cplusplus
}
#endif
#endif
