/* b200_integrator.h -- the reference-side binding: the file a maintainer of JettHuang/jet-pbrt adds next to
 * src/integrator.h to send the hot path to libjetpbrt_b200.so.
 *
 * It is compiled WITH the reference (it includes the reference's own headers) -- in this repository by
 * oracle/Makefile's `shim` target, against /root/reference/src in place, into oracle/_ref/libjetpbrt_refshim.so, and
 * tests/test_gpu_reference_shim.py renders through it into an FFilm and saves with the reference's own
 * FFilm::SaveAsImage.
 *
 * FIntegrator::Render is public and NON-virtual (integrator.h:32) and the reference's scene objects keep their data
 * private (FSphere shape.h:658-661; materials, lights, camera `protected`), so the binding is a class with the same
 * Render() signature that carries the neutral description (include/jetpbrt_scene.h) recorded where the scene is built:
 * main.cc:154-156 changes by one type name.
 */
#pragma once

#include <vector>

#include "film.h"        /* reference: FFilm, FColor */
#include "integrator.h"  /* reference: FIntegrator (for the signature being mirrored), FScene, FSampler */
#include "jetpbrt_b200.h"

namespace pbrt {

class FB200PathIntegrator {
public:
    /* maxDepth as FPathIntegratorIteration(int) (integrator.h:108-121); ngpus > 1 renders with that many devices of
     * this process where the reference passes numthreads (jpbrt_render_multi). */
    FB200PathIntegrator(int inMaxDepth, const jpbrt_scene_desc* inDesc, int inDevice = 0, int inNumGpus = 1, uint64_t inSeed = 1234)
        : maxDepth(inMaxDepth), desc(*inDesc), device(inDevice), ngpus(inNumGpus), seed(inSeed) { desc.max_depth = inMaxDepth; }

    /* Same signature and contract as FIntegrator::Render (integrator.h:32, integrator.cc:35-80): blocks, ADDS
     * Clamp01(mean radiance) onto the caller's film (film.h:64-68), prints the seconds.  `scene` is unused (the
     * description stands in for it) and may be null; `numthreads` is the reference's host-thread count and is ignored. */
    void Render(const FScene* /*scene*/, FSampler* sampler, FFilm* film, int /*numthreads*/ = 1) const {
        const int spp = sampler->GetSamplesPerPixel();  /* sampler.h:76-79 */
        if (desc.camera.width != film->Width() || desc.camera.height != film->Height()) {
            PBRT_ERROR("%s", "FB200PathIntegrator: film resolution differs from the camera's\n");
            return;
        }
        std::vector<float> rgb((size_t)film->Width() * film->Height() * 3);
        double seconds = 0, reduce_ms = 0;
        PBRT_PRINT("%s", "start rendering ...\n");  /* integrator.cc:44 */
        const int rc = ngpus > 1 ? jpbrt_render_multi(&desc, JPBRT_INTEGRATOR_PATH, spp, seed, ngpus, rgb.data(), &seconds, &reduce_ms)
                                 : jpbrt_render(&desc, spp, seed, device, rgb.data(), &seconds);
        if (rc != 0) {
            PBRT_ERROR("B200 render failed: %s\n", jpbrt_last_error(nullptr));
            return;
        }
        for (int y = 0; y < film->Height(); ++y)  /* FFilmView::AddColor(x, y, Clamp01(L)), integrator.cc:108 */
            for (int x = 0; x < film->Width(); ++x) {
                const float* p = &rgb[3 * ((size_t)y * film->Width() + x)];
                film->AddColor(x, y, FColor(p[0], p[1], p[2]));
            }
        PBRT_PRINT("%s", "finish rendering ...\n");                            /* integrator.cc:78 */
        PBRT_PRINT("FIntegrator::Render used %f seconds.\n", (float)seconds);  /* integrator.cc:79 */
    }

private:
    int maxDepth;
    jpbrt_scene_desc desc;  /* shallow copy: the arrays stay owned by the caller (jetpbrt::Scene), as FScene owns the reference's */
    int device, ngpus;
    uint64_t seed;
};

}  /* namespace pbrt */
