// render.cc -- see render.h.
#include "render.h"

#include <algorithm>
#include <chrono>
#include <cstdio>

#include "../../include/jetpbrt_b200.h"

namespace jetpbrt {

bool Film::SaveAsImage(const std::string& filename, EImageType type) const {
    return jpbrt_save_image(filename.c_str(), (int)type, width_, height_, pixels_.data()) == JPBRT_OK;
}

bool Integrator::Render(Scene* scene, int spp, Film* film, int device, uint64_t seed) const {
    if (maxDepth_ >= 0) scene->SetMaxDepth(maxDepth_);
    const jpbrt_scene_desc* desc = scene->Desc();
    if (desc->camera.width != film->Width() || desc->camera.height != film->Height()) {
        fprintf(stdout, "film resolution does not match the camera's\n");
        return false;
    }
    auto t0 = std::chrono::steady_clock::now();
    fprintf(stdout, "start rendering ...\n");  // integrator.cc:44
    jpbrt_ctx* ctx = nullptr;
    int rc = jpbrt_upload_scene(desc, device, &ctx);
    std::vector<float> tmp((size_t)film->Width() * film->Height() * 3);
    if (rc == 0) rc = jpbrt_set_option(ctx, "integrator", kind_);
    if (rc == 0) rc = jpbrt_render_pass(ctx, 0, spp, seed);
    if (rc == 0) rc = jpbrt_read_film(ctx, tmp.data(), spp, 1);
    if (rc != 0) {
        fprintf(stdout, "render failed: %s\n", jpbrt_last_error(ctx));
        jpbrt_destroy(ctx);
        return false;
    }
    float* dst = film->Data();
    for (size_t i = 0; i < tmp.size(); ++i) dst[i] += tmp[i];  // FFilm::AddColor, film.h:64-68
    jpbrt_stats st;
    jpbrt_get_stats(ctx, &st);
    jpbrt_destroy(ctx);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stdout, "finish rendering ...\n");                                   // integrator.cc:78
    fprintf(stdout, "FIntegrator::Render used %f seconds.\n", (float)sec);        // integrator.cc:79
    fprintf(stdout, "  (%llu samples, %llu extension + %llu shadow rays, BVH build %.3f s)\n",
            (unsigned long long)st.samples, (unsigned long long)st.extension_rays, (unsigned long long)st.shadow_rays,
            st.bvh_build_seconds);
    return true;
}

bool Integrator::RenderMultiGpu(Scene* scene, int spp, Film* film, int ngpus, uint64_t seed) const {
    if (ngpus <= 1) return Render(scene, spp, film, 0, seed);
    if (maxDepth_ >= 0) scene->SetMaxDepth(maxDepth_);
    const jpbrt_scene_desc* desc = scene->Desc();
    if (desc->camera.width != film->Width() || desc->camera.height != film->Height()) {
        fprintf(stdout, "film resolution does not match the camera's\n");
        return false;
    }
    fprintf(stdout, "start rendering ...\n");  // integrator.cc:44
    std::vector<float> tmp((size_t)film->Width() * film->Height() * 3);
    double sec = 0, reduce_ms = 0;
    int rc = jpbrt_render_multi(desc, kind_, spp, seed, ngpus, tmp.data(), &sec, &reduce_ms);
    if (rc != 0) {
        fprintf(stdout, "render failed: %s\n", jpbrt_last_error(nullptr));
        return false;
    }
    float* dst = film->Data();
    for (size_t i = 0; i < tmp.size(); ++i) dst[i] += tmp[i];  // FFilm::AddColor, film.h:64-68
    fprintf(stdout, "finish rendering ...\n");                                   // integrator.cc:78
    fprintf(stdout, "FIntegrator::Render used %f seconds.\n", (float)sec);        // integrator.cc:79
    fprintf(stdout, "  (%d GPUs, samples partitioned, NCCL film reduce %.3f ms)\n", ngpus, reduce_ms);
    return true;
}

}  // namespace jetpbrt
