// ref_harness.cc -- C harness around the UNMODIFIED reference (JettHuang/jet-pbrt).
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).  Compiled together with the reference's
// own translation units, in place from /root/reference/src, by oracle/Makefile; the result
// (oracle/_ref/libjetpbrt_ref.so) is git-ignored and is the parity anchor for everything else:
// the CPU restatement (oracle/pt_oracle.cc) is pinned against it, golden fixtures under
// tests/golden/ are generated from it, and bench.py --impl reference times it.
//
// Nothing here re-implements reference arithmetic: every function below only marshals arrays
// into reference objects and calls the reference's public methods.  Scenes are rebuilt from the
// neutral description through FScene::Create* exactly the way main.cc:13-111 and
// scene.cc:49-97 do it.

#include "pbrt.h"
#include "light.h"
#include "integrator.h"
#include "microfacet.h"

#include <chrono>
#include <ctime>
#include <unordered_map>

#define ORC(name) jref_##name
#include "oracle_api.h"

// ---- what pbrt.cc would have provided (it needs <Windows.h>, so it is not compiled) ----------
namespace pbrt {
static bool g_ref_verbose = false;
void log_print_fmt_only(const char* fmt) { if (g_ref_verbose) fputs(fmt, stdout); }
static double now_us() {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
double appInitTiming() { return now_us() * 1e-6; }
double appSeconds() { return now_us() * 1e-6; }
double appMicroSeconds() { return now_us(); }
int64_t appCycles() { return (int64_t)now_us(); }
}  // namespace pbrt

using namespace pbrt;

namespace {

inline FVector3 V3(const float* p) { return FVector3(p[0], p[1], p[2]); }
inline FColor C3(const float* p) { return FColor(p[0], p[1], p[2]); }
inline void put3(float* o, const FVector3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
inline void putcol(float* o, const FColor& v) { o[0] = v.r; o[1] = v.g; o[2] = v.b; }

std::shared_ptr<FShape> make_shape(const jpbrt_shape& s) {
    switch (s.type) {
    case JPBRT_SHAPE_TRIANGLE: {
        FVector3 p0 = V3(s.p[0]), p1 = V3(s.p[1]), p2 = V3(s.p[2]);  // third vertex is a non-const ref (shape.h:280)
        FVector2 uv;
        return std::make_shared<FTriangle>(p0, p1, p2, uv, uv, uv, s.flip_normal != 0);
    }
    case JPBRT_SHAPE_RECTANGLE:
        return std::make_shared<FRectangle>(V3(s.p[0]), V3(s.p[1]), V3(s.p[2]), V3(s.p[3]), s.flip_normal != 0);
    case JPBRT_SHAPE_SPHERE:
        return std::make_shared<FSphere>(V3(s.p[0]), s.p[1][0]);
    case JPBRT_SHAPE_DISK:
        return std::make_shared<FDisk>(V3(s.p[0]), V3(s.p[1]), s.p[2][0]);
    }
    return nullptr;
}

std::shared_ptr<FMaterial> make_material(const jpbrt_material& m) {
    switch (m.type) {
    case JPBRT_MAT_MATTE:   return std::make_shared<FMatteMaterial>(C3(m.a));
    case JPBRT_MAT_MIRROR:  return std::make_shared<FMirrorMaterial>(C3(m.a));
    case JPBRT_MAT_GLASS:   return std::make_shared<FGlassMaterial>(m.f0, C3(m.a), C3(m.b));
    case JPBRT_MAT_PLASTIC: return std::make_shared<FPlasticMaterial>(C3(m.a), C3(m.b), m.f0, m.remap_roughness != 0);
    case JPBRT_MAT_METAL:   return std::make_shared<FMetalMaterial>(C3(m.a), C3(m.b), m.f0, m.f1, m.remap_roughness != 0);
    }
    return nullptr;
}

// FRandomSampler whose stream can be reseeded; Clone() keeps the reference's quirk that every
// band task restarts from the SAME seed (sampler.h:135-138, integrator.cc:66).
class SeededRandomSampler : public FRandomSampler {
public:
    SeededRandomSampler(int spp, int seed) : FRandomSampler(spp), seed_(seed) { rng = FRNG(seed); }
    std::unique_ptr<FSampler> Clone() override { return std::make_unique<SeededRandomSampler>(samples_per_pixel, seed_); }
private:
    int seed_;
};

// Sampler that hands a caller-chosen number to material->Scattering (plastic lobe pick, material.cc:14).
class FixedSampler : public FSampler {
public:
    explicit FixedSampler(Float v) : FSampler(1), v_(v) {}
    void Set(Float v) { v_ = v; }
    std::unique_ptr<FSampler> Clone() override { return std::make_unique<FixedSampler>(v_); }
    Float GetFloat() override { return v_; }
    FFloat2 GetFloat2() override { return FFloat2(v_, v_); }
    FCameraSample GetCameraSample(const FPoint2& p) override { FCameraSample s; s.posfilm = p; return s; }
private:
    Float v_;
};

}  // namespace

struct jref_scene {
    std::shared_ptr<FScene> scene;
    std::unordered_map<const FPrimitive*, int> prim_index;
    std::vector<FLight*> lights;  // creation order == desc->lights order
    int max_depth = 5;
    int width = 0, height = 0;
};

extern "C" {

void jref_set_verbose(int v) { pbrt::g_ref_verbose = v != 0; }

jref_scene* jref_scene_create(const jpbrt_scene_desc* d) {
    if (!d || d->n_primitives <= 0) return nullptr;
    auto* out = new jref_scene();
    out->max_depth = d->max_depth;
    out->width = d->camera.width;
    out->height = d->camera.height;
    auto scene = std::make_shared<FScene>(d->name ? d->name : "scene");
    out->scene = scene;

    const jpbrt_camera& c = d->camera;
    scene->CreateCamera<FCamera>(V3(c.pos), V3(c.front), V3(c.up), (Float)c.vfov_deg,
                                 FVector2((Float)c.width, (Float)c.height));

    std::vector<std::shared_ptr<FShape>> shapes(d->n_shapes);
    for (int i = 0; i < d->n_shapes; ++i) {
        shapes[i] = make_shape(d->shapes[i]);
        if (!shapes[i]) { delete out; return nullptr; }
        scene->shapes.push_back(shapes[i]);  // what CreateShape / CreateTriangleMesh do (scene.h:75-82, scene.cc:56-60)
    }
    std::vector<std::shared_ptr<FMaterial>> mats(d->n_materials);
    for (int i = 0; i < d->n_materials; ++i) {
        mats[i] = make_material(d->materials[i]);
        if (!mats[i]) { delete out; return nullptr; }
        scene->materials.push_back(mats[i]);
    }
    // Lights in creation order (scene.h:93-107).
    for (int i = 0; i < d->n_lights; ++i) {
        const jpbrt_light& l = d->lights[i];
        std::shared_ptr<FLight> lp;
        switch (l.type) {
        case JPBRT_LIGHT_ENVIRONMENT: lp = scene->CreateLight<FEnvironmentLight>(V3(l.pos), 1, C3(l.color)); break;
        case JPBRT_LIGHT_AREA:
            if (l.shape < 0 || l.shape >= d->n_shapes) { delete out; return nullptr; }
            lp = scene->CreateLight<FAreaLight>(FPoint3(0, 0, 0), 1, C3(l.color), (const FShape*)shapes[l.shape].get());
            break;
        case JPBRT_LIGHT_POINT:     lp = scene->CreateLight<FPointLight>(V3(l.pos), 1, C3(l.color)); break;
        case JPBRT_LIGHT_DIRECTION: lp = scene->CreateLight<FDirectionLight>(V3(l.pos), 1, C3(l.color), V3(l.dir)); break;
        default: delete out; return nullptr;
        }
        out->lights.push_back(lp.get());
    }
    // Primitives in creation order (scene.h:109-118).
    for (int i = 0; i < d->n_primitives; ++i) {
        const jpbrt_primitive& p = d->primitives[i];
        const FShape* sh = shapes[p.shape].get();
        const FMaterial* mt = p.material >= 0 ? mats[p.material].get() : nullptr;
        const FAreaLight* al = p.light >= 0 ? static_cast<const FAreaLight*>(out->lights[p.light]) : nullptr;
        auto prim = scene->CreatePrimitive(sh, mt, al);
        out->prim_index[prim.get()] = i;
    }
    // The BVH build draws its split axes from rand() (pbrt.h:106-120, bvh.h:61), which the
    // reference never seeds; a fresh process therefore always sees the srand(1) sequence.
    srand(1);
    scene->Preprocess();
    return out;
}

void jref_scene_destroy(jref_scene* s) { delete s; }

int jref_intersect_shape(const jpbrt_shape* shape, int n, const float* r, int* hit, float* t, float* pos3, float* nrm3) {
    auto sh = make_shape(*shape);
    if (!sh) return -1;
    for (int i = 0; i < n; ++i) {
        const float* q = r + 8 * i;
        FRay ray(V3(q), V3(q + 3), q[6], q[7]);
        FIntersection isect;
        bool h = sh->Intersect(ray, isect);
        hit[i] = h ? 1 : 0;
        t[i] = h ? ray.MaxT() : 0.f;
        put3(pos3 + 3 * i, h ? isect.position : FVector3());
        put3(nrm3 + 3 * i, h ? isect.normal : FVector3());
    }
    return 0;
}

int jref_scene_intersect(jref_scene* s, int n, const float* r, int* prim, float* t, float* pos3, float* nrm3) {
    for (int i = 0; i < n; ++i) {
        const float* q = r + 8 * i;
        FRay ray(V3(q), V3(q + 3), q[6], q[7]);
        FIntersection isect;
        bool h = s->scene->Intersect(ray, isect);
        prim[i] = h ? s->prim_index[isect.primitive] : -1;
        t[i] = h ? ray.MaxT() : 0.f;
        if (pos3) put3(pos3 + 3 * i, h ? isect.position : FVector3());
        if (nrm3) put3(nrm3 + 3 * i, h ? isect.normal : FVector3());
    }
    return 0;
}

int jref_scene_occluded(jref_scene* s, int n, const float* pos3, const float* target3, int* occluded) {
    for (int i = 0; i < n; ++i) {
        FIntersection isect(V3(pos3 + 3 * i), FVector3(0, 0, 1), FVector3(0, 0, 1));
        occluded[i] = s->scene->Occluded(isect, V3(target3 + 3 * i)) ? 1 : 0;
    }
    return 0;
}

int jref_bsdf(const jpbrt_material* mat, int n, const float* nrm3, const float* wo3, const float* wi3,
              const float* u2, const float* ulobe, float* f_eval3, float* pdf_eval,
              float* s_wi3, float* s_f3, float* s_pdf, int* s_flags, int* is_delta) {
    auto m = make_material(*mat);
    if (!m) return -1;
    FixedSampler sampler(0.f);
    for (int i = 0; i < n; ++i) {
        FIntersection isect(FVector3(0, 0, 0), V3(nrm3 + 3 * i), V3(wo3 + 3 * i));
        sampler.Set(ulobe ? ulobe[i] : 0.f);
        std::unique_ptr<FBSDF> bsdf = m->Scattering(isect, &sampler);
        FVector3 wo = V3(wo3 + 3 * i), wi = V3(wi3 + 3 * i);
        putcol(f_eval3 + 3 * i, bsdf->Evalf(wo, wi));
        pdf_eval[i] = bsdf->Pdf(wo, wi);
        FBSDFSample bs = bsdf->Sample(wo, FFloat2(u2[2 * i], u2[2 * i + 1]));
        put3(s_wi3 + 3 * i, bs.wi);
        putcol(s_f3 + 3 * i, bs.f);
        s_pdf[i] = bs.pdf;
        s_flags[i] = bs.ebsdf;
        is_delta[i] = bsdf->IsDelta() ? 1 : 0;
    }
    return 0;
}

int jref_bsdf_ex(const jpbrt_bsdf_desc* d, int n, const float* nrm3, const float* wo3, const float* wi3, const float* u2,
                 float* f_eval3, float* pdf_eval, float* s_wi3, float* s_f3, float* s_pdf, int* s_flags) {
    if (!d) return -1;
    for (int i = 0; i < n; ++i) {
        FFrame frame(V3(nrm3 + 3 * i));
        std::unique_ptr<FBSDF> bsdf;
        auto make_dist = [&]() -> MicrofacetDistribution* {  // owned (and deleted) by the BSDF, bsdf.cc:29-33,80-83
            if (d->distribution == JPBRT_DIST_BECKMANN) return new BeckmannDistribution(d->alphax, d->alphay, d->sample_visible_area != 0);
            return new TrowbridgeReitzDistribution(d->alphax, d->alphay, d->sample_visible_area != 0);
        };
        switch (d->kind) {
        case JPBRT_BSDF_PHONG: bsdf = std::make_unique<FPhongSpecularReflection>(frame, C3(d->color), d->exponent); break;
        case JPBRT_BSDF_MICROFACET_REFLECTION: {
            Fresnel* fr = nullptr;
            if (d->fresnel == JPBRT_FRESNEL_DIELECTRIC) fr = new FresnelDielectric(d->eta_a, d->eta_b);
            else if (d->fresnel == JPBRT_FRESNEL_CONDUCTOR) fr = new FresnelConductor(C3(d->c_eta_i), C3(d->c_eta_t), C3(d->c_k));
            else fr = new FresnelNoOp();
            bsdf = std::make_unique<FMicrofacetReflection>(frame, C3(d->color), make_dist(), fr);
            break;
        }
        case JPBRT_BSDF_MICROFACET_TRANSMISSION:
            bsdf = std::make_unique<FMicrofacetTransmission>(frame, C3(d->color), make_dist(), d->eta_a, d->eta_b);
            break;
        default: return -1;
        }
        FVector3 wo = V3(wo3 + 3 * i), wi = V3(wi3 + 3 * i);
        putcol(f_eval3 + 3 * i, bsdf->Evalf(wo, wi));
        pdf_eval[i] = bsdf->Pdf(wo, wi);
        FBSDFSample smp = bsdf->Sample(wo, FFloat2(u2[2 * i], u2[2 * i + 1]));
        put3(s_wi3 + 3 * i, smp.wi);
        putcol(s_f3 + 3 * i, smp.f);
        s_pdf[i] = smp.pdf;
        s_flags[i] = smp.ebsdf;
    }
    return 0;
}

int jref_light_sample(jref_scene* s, int light, int n, const float* pos3, const float* nrm3, const float* u2,
                      float* lpos3, float* wi3, float* pdf, float* Li3) {
    if (light < 0 || light >= (int)s->lights.size()) return -1;
    FLight* L = s->lights[light];
    for (int i = 0; i < n; ++i) {
        FIntersection isect(V3(pos3 + 3 * i), V3(nrm3 + 3 * i), FVector3(0, 0, 1));
        FLightSample ls = L->Sample_Li(isect, FFloat2(u2[2 * i], u2[2 * i + 1]));
        put3(lpos3 + 3 * i, ls.pos);
        put3(wi3 + 3 * i, ls.wi);
        pdf[i] = ls.pdf;
        putcol(Li3 + 3 * i, ls.Li);
    }
    return 0;
}

int jref_emitted(jref_scene* s, int n, const int* prim, const float* nrm3, const float* wo3, float* Le3) {
    std::vector<const FPrimitive*> by_index(s->prim_index.size(), nullptr);
    for (auto& kv : s->prim_index) by_index[kv.second] = kv.first;
    for (int i = 0; i < n; ++i) {
        FIntersection isect(FVector3(0, 0, 0), V3(nrm3 + 3 * i), V3(wo3 + 3 * i));
        isect.primitive = (prim[i] >= 0 && prim[i] < (int)by_index.size()) ? by_index[prim[i]] : nullptr;
        putcol(Le3 + 3 * i, isect.Le());
    }
    return 0;
}

int jref_generate_rays(jref_scene* s, int n, const float* posfilm2, float* o3, float* d3) {
    const FCamera* cam = s->scene->Camera();
    for (int i = 0; i < n; ++i) {
        FCameraSample cs;
        cs.posfilm = FPoint2(posfilm2[2 * i], posfilm2[2 * i + 1]);
        FRay ray = cam->GenerateRay(cs);
        put3(o3 + 3 * i, ray.Origin());
        put3(d3 + 3 * i, ray.Dir());
    }
    return 0;
}

double jref_render_mode(jref_scene* s, int mode, int spp, int numthreads, int seed, float* film_out) {
    if (!s || spp <= 0 || !film_out || mode < 0 || mode > 3) return -1.0;
    FFilm film(s->width, s->height);
    film.Clear();
    std::shared_ptr<FSampler> sampler;
    if (seed < 0) sampler = std::make_shared<FRandomSampler>(spp);   // main.cc:149
    else          sampler = std::make_shared<SeededRandomSampler>(spp, seed);
    std::unique_ptr<FIntegrator> integrator;                          // main.cc:151-154
    switch (mode) {
    case 1: integrator = std::make_unique<FPathIntegratorRecursive>(s->max_depth); break;
    case 2: integrator = std::make_unique<FWhittedIntegrator>(s->max_depth); break;
    case 3: integrator = std::make_unique<FDebugIntegrator>(); break;
    default: integrator = std::make_unique<FPathIntegratorIteration>(s->max_depth); break;
    }
    auto t0 = std::chrono::steady_clock::now();
    integrator->Render(s->scene.get(), sampler.get(), &film, numthreads);  // main.cc:156
    auto t1 = std::chrono::steady_clock::now();
    for (int y = 0; y < s->height; ++y)
        for (int x = 0; x < s->width; ++x) {
            const FColor& c = film(x, y);
            float* o = film_out + 3 * ((size_t)y * s->width + x);
            o[0] = c.r; o[1] = c.g; o[2] = c.b;
        }
    return std::chrono::duration<double>(t1 - t0).count();
}

double jref_render(jref_scene* s, int spp, int numthreads, int seed, float* film_out) {
    return jref_render_mode(s, 0, spp, numthreads, seed, film_out);
}

int jref_scene_info(jref_scene* s, float* out7) {
    FBounds3 b = s->scene->WorldBound();
    put3(out7, b._min);
    put3(out7 + 3, b._max);
    FPoint3 c; Float r;
    b.BoundingSphere(c, r);  // what FEnvironmentLight::Preprocess stores (light.cc:26-33)
    out7[6] = r;
    return 0;
}

// ---- the callers and data formats either side of the path, for pinning the product's writers ----

// FFilm::AddColor + FFilm::SaveAsImage (film.h:64-68, film.cc:11-188): kind 0 PPM, 1 BMP, 2 HDR; basename without extension.
int jref_save_image(const char* basename, int kind, int width, int height, const float* rgb) {
    FFilm film(width, height);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const float* p = rgb + 3 * ((size_t)y * width + x);
            film.AddColor(x, y, FColor(p[0], p[1], p[2]));
        }
    const EImageType t = kind == 0 ? EImageType::PPM : kind == 1 ? EImageType::BMP : EImageType::HDR;
    return film.SaveAsImage(basename, t) ? 0 : -1;
}

// LoadTriangleMesh (shape.cc:23-68, through external/obj_loader.h): 9 floats per triangle (p0, p1, p2) + its stored normal.
// Returns the triangle count (also when it exceeds `capacity`), or -1 when the load fails.
int jref_load_obj(const char* filename, int flip_normal, int flip_handedness, const float* offset3, float scale,
                  float* tris9, float* normals3, int capacity) {
    std::vector<std::shared_ptr<FTriangle>> mesh;
    if (!LoadTriangleMesh(filename, mesh, flip_normal != 0, flip_handedness != 0, V3(offset3), scale)) return -1;
    for (int i = 0; i < (int)mesh.size() && i < capacity; ++i) {
        put3(tris9 + 9 * i, mesh[i]->p0); put3(tris9 + 9 * i + 3, mesh[i]->p1); put3(tris9 + 9 * i + 6, mesh[i]->p2);
        if (normals3) put3(normals3 + 3 * i, mesh[i]->normal);
    }
    return (int)mesh.size();
}

}  // extern "C"
