// main.cc -- command line with the reference's arguments (main.cc:113-163): jetpbrt sceneid spp
//   sceneid 0 = cornell box (create_cornellbox_scene), 1 = bunny scene (create_bunny_scene),
//   2 = large mesh (C3), 3 = glossy / 16 lights (C4).  Defaults: 1024 x 1024, 50 spp, depth 5, BMP.
// Extra optional arguments: width height device integrator  (integrator: 0 path -- what main.cc:154 ships with --
// 1 recursive path, 2 Whitted, 3 debug: the alternatives main.cc:151-153 keeps commented out), and `--gpus N` anywhere
// on the line: render with N devices of this box (where the reference passes numthreads = 16, main.cc:156).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "render.h"

using namespace jetpbrt;

int main(int argc, char* argv[]) {
    int width = 1024, height = 1024;
    int samples_per_pixel = 50;
    printf("pbrt.exe  sceneid   spp\n");
    int ngpus = 1;
    for (int i = 1; i + 1 < argc; ++i) {  // strip "--gpus N"
        if (!strcmp(argv[i], "--gpus")) {
            ngpus = atoi(argv[i + 1]);
            for (int j = i; j + 2 < argc; ++j) argv[j] = argv[j + 2];
            argc -= 2;
            break;
        }
    }
    if (argc < 2) return 0;
    int sceneId = atoi(argv[1]);
    if (argc > 2) { int spp = atoi(argv[2]); if (spp > 0) samples_per_pixel = spp; }
    if (argc > 4) { width = atoi(argv[3]); height = atoi(argv[4]); }
    int device = argc > 5 ? atoi(argv[5]) : 0;
    static const char* names[] = {"cornell", "bunny", "large", "glossy"};
    if (sceneId < 0 || sceneId > 3 || width <= 0 || height <= 0) return 0;
    std::unique_ptr<Scene> scene(MakeBuiltinScene(names[sceneId], width, height, 1.f));
    if (!scene) return 0;
    printf("current scene: %s\n", scene->Name().c_str());
    Film film(width, height);
    int depth = scene->Desc()->max_depth;
    int kind = argc > 6 ? atoi(argv[6]) : 0;
    std::unique_ptr<Integrator> integrator;
    switch (kind) {
    case 1: integrator.reset(new PathIntegratorRecursive(depth)); break;
    case 2: integrator.reset(new WhittedIntegrator(depth)); break;
    case 3: integrator.reset(new DebugIntegrator()); break;
    default: integrator.reset(new PathIntegratorIteration(depth)); break;
    }
    const bool ok = ngpus > 1 ? integrator->RenderMultiGpu(scene.get(), samples_per_pixel, &film, ngpus)
                              : integrator->Render(scene.get(), samples_per_pixel, &film, device);
    if (!ok) return 1;
    char fullname[256];
    snprintf(fullname, sizeof(fullname), "%s_%d", scene->Name().c_str(), samples_per_pixel);
    film.SaveAsImage(fullname, EImageType::BMP);
    return 0;
}
