// ref_shim.cc -- TEST INFRASTRUCTURE: compiles include/b200_integrator.h (the reference-side binding of INTEGRATION.md)
// against the UNMODIFIED reference and drives it the way main.cc:149-160 drives an integrator: FFilm + FRandomSampler +
// Render + FFilm::SaveAsImage.  Linked with the reference's objects AND libjetpbrt_b200.so into
// oracle/_ref/libjetpbrt_refshim.so (oracle/Makefile, target `shim`).  The product never loads it.
#include "pbrt.h"
#include "sampler.h"

#include "../include/b200_integrator.h"

#include <cstring>

using namespace pbrt;

extern "C" {

// main.cc:149-160 with FB200PathIntegrator in place of FPathIntegratorIteration.  Returns 0, fills film_out (w*h*3, the
// FFilm's pixels after Render) and, if basename is given, writes <basename>.<ext> with the reference's own writer.
int jshim_render(const jpbrt_scene_desc* desc, int max_depth, int spp, int device, int ngpus, unsigned long long seed,
                 const char* basename, int image_kind, float* film_out) {
    const int w = desc->camera.width, h = desc->camera.height;
    FFilm film(w, h);                                   // main.cc:116
    std::shared_ptr<FSampler> sampler = std::make_shared<FRandomSampler>(spp);  // main.cc:149
    FB200PathIntegrator integrator(max_depth, desc, device, ngpus, seed);       // main.cc:154
    integrator.Render(nullptr, sampler.get(), &film, 16);                       // main.cc:156
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const FColor& c = film(x, y);
            float* o = film_out + 3 * ((size_t)y * w + x);
            o[0] = c.r; o[1] = c.g; o[2] = c.b;
        }
    if (basename && *basename) {
        const EImageType t = image_kind == 0 ? EImageType::PPM : image_kind == 1 ? EImageType::BMP : EImageType::HDR;
        if (!film.SaveAsImage(basename, t)) return -1;  // main.cc:160
    }
    return 0;
}

}  // extern "C"
