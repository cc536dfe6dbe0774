// render.h -- host-side mirror of the reference's render entry point for the B200 path.
//
//   reference (integrator.h:25-41, film.h:27-94, main.cc:149-160)     here
//   FFilm film(w, h)                                                   jetpbrt::Film film(w, h)
//   FRandomSampler sampler(spp)                                        (spp argument; sampler is counter based)
//   FPathIntegratorIteration integrator(5)                             jetpbrt::PathIntegratorIteration integrator(5)
//   FDebugIntegrator / FWhittedIntegrator(5) / FPathIntegratorRecursive(5)   jetpbrt::DebugIntegrator / WhittedIntegrator(5) / ...
//   integrator.Render(scene, sampler, &film, 16)                       integrator.Render(scene, spp, &film, device)
//     (numthreads = 16 host threads, integrator.cc:53-74)                integrator.Render(scene, spp, &film, ngpus): GPUs for threads
//   film.SaveAsImage(name, EImageType::BMP)                            film.SaveAsImage(name, EImageType::BMP)
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "scene.h"

namespace jetpbrt {

enum class EImageType { PPM = 0, BMP = 1, HDR = 2 };  // film.h:15-20

class Film {
public:
    Film(int w, int h) : width_(w), height_(h), pixels_((size_t)w * h * 3, 0.f) {}
    int Width() const { return width_; }
    int Height() const { return height_; }
    float* Data() { return pixels_.data(); }
    const float* Data() const { return pixels_.data(); }
    void Clear() { std::fill(pixels_.begin(), pixels_.end(), 0.f); }
    bool SaveAsImage(const std::string& filename, EImageType type) const;  // film.cc:13-43

private:
    int width_, height_;
    std::vector<float> pixels_;  // row-major RGB, row 0 = top (film.h:50-56)
};

// FIntegrator (integrator.h:25-41): Render() is the public entry; the subclass picks the estimator.
class Integrator {
public:
    virtual ~Integrator() {}
    // Blocks until the film is filled, like FIntegrator::Render (integrator.cc:35-80); adds
    // Clamp01(mean radiance) onto the film (film.h:64-68).  Returns false and prints the C-ABI
    // error on failure (the reference's Render returns void and cannot fail).
    bool Render(Scene* scene, int spp, Film* film, int device = 0, uint64_t seed = 1234) const;
    // The same with `ngpus` devices (0 .. ngpus-1) of this process where the reference has `numthreads` host threads
    // (integrator.h:32): the samples of every pixel are partitioned across the devices, the partial films are summed
    // with one NCCL reduce (jpbrt_render_multi).
    bool RenderMultiGpu(Scene* scene, int spp, Film* film, int ngpus, uint64_t seed = 1234) const;

protected:
    Integrator(int kind, int maxDepth) : kind_(kind), maxDepth_(maxDepth) {}
    int kind_;      // jpbrt_integrator
    int maxDepth_;  // < 0: keep the scene's
};

class PathIntegratorIteration : public Integrator {  // integrator.h:108-121 -- the hot path
public:
    explicit PathIntegratorIteration(int maxDepth) : Integrator(0, maxDepth) {}
};
class PathIntegratorRecursive : public Integrator {  // integrator.h:87-105
public:
    explicit PathIntegratorRecursive(int maxDepth) : Integrator(1, maxDepth) {}
};
class WhittedIntegrator : public Integrator {  // integrator.h:62-84
public:
    explicit WhittedIntegrator(int maxDepth) : Integrator(2, maxDepth) {}
};
class DebugIntegrator : public Integrator {  // integrator.h:44-58
public:
    DebugIntegrator() : Integrator(3, -1) {}
};

}  // namespace jetpbrt
