/* jetpbrt_scene.h -- neutral, POD scene description shared by every consumer of the hot path.
 *
 * The reference (JettHuang/jet-pbrt) describes a scene by calling FScene::Create* factory
 * templates (reference src/scene.h:66-124) on heap objects whose data members are private or
 * protected.  The B200 path, the CPU oracle restatement (oracle/) and
 * the compiled-reference harness (oracle/_ref) all consume THIS description instead; the
 * harness rebuilds the reference's own objects from it through the reference's public API.
 *
 * Everything is plain C: fixed-size structs, host-owned arrays, no pointers into reference
 * objects.  Lists keep the reference's CREATION ORDER, which is observable:
 *   - lights[]      : FScene::Lights() order = NEE loop order (integrator.cc:359-371)
 *   - primitives[]  : shadow_primitives order = BVH input order / tie-break order (bvh.h:99-100)
 */
#ifndef JETPBRT_SCENE_H
#define JETPBRT_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- camera: FCamera ctor arguments (camera.h:36) ---- */
typedef struct jpbrt_camera {
    float pos[3];
    float front[3];   /* need not be unit; FCamera normalises it (camera.h:38) */
    float up[3];
    float vfov_deg;   /* the reference's "ifov" in degrees (camera.h:44) */
    int   width;      /* film resolution (film.h:30) */
    int   height;
} jpbrt_camera;

/* ---- shapes (shape.h) ---- */
enum {
    JPBRT_SHAPE_TRIANGLE  = 0,  /* FTriangle  shape.h:277  p[0..2] = p0,p1,p2 */
    JPBRT_SHAPE_RECTANGLE = 1,  /* FRectangle shape.h:380  p[0..3] = p0,p1,p2,p3 */
    JPBRT_SHAPE_SPHERE    = 2,  /* FSphere    shape.h:476  p[0] = centre, p[1][0] = radius */
    JPBRT_SHAPE_DISK      = 3   /* FDisk      shape.h:189  p[0] = position, p[1] = normal, p[2][0] = radius */
};

typedef struct jpbrt_shape {
    int   type;
    int   flip_normal;  /* FTriangle/FRectangle ctor flag (shape.h:280,383); ignored otherwise */
    float p[4][3];
} jpbrt_shape;

/* ---- materials (material.h) ---- */
enum {
    JPBRT_MAT_MATTE   = 0,  /* FMatteMaterial   material.h:27   a = diffuseColor */
    JPBRT_MAT_MIRROR  = 1,  /* FMirrorMaterial  material.h:45   a = specularColor */
    JPBRT_MAT_GLASS   = 2,  /* FGlassMaterial   material.h:63   f0 = eta, a = Kr, b = Kt */
    JPBRT_MAT_PLASTIC = 3,  /* FPlasticMaterial material.h:85   a = Kd, b = Ks, f0 = roughness, remap */
    JPBRT_MAT_METAL   = 4   /* FMetalMaterial   material.h:113  a = eta, b = k, f0 = uRough, f1 = vRough, remap */
};

typedef struct jpbrt_material {
    int   type;
    int   remap_roughness;
    float a[3];
    float b[3];
    float f0;
    float f1;
} jpbrt_material;

/* ---- lights (light.h) ---- */
enum {
    JPBRT_LIGHT_ENVIRONMENT = 0,  /* FEnvironmentLight light.h:248  color = radiance */
    JPBRT_LIGHT_AREA        = 1,  /* FAreaLight        light.h:183  color = radiance, shape = index into shapes[] */
    JPBRT_LIGHT_POINT       = 2,  /* FPointLight       light.h:81   color = intensity, pos */
    JPBRT_LIGHT_DIRECTION   = 3   /* FDirectionLight   light.h:136  color = irradiance, dir */
};

typedef struct jpbrt_light {
    int   type;
    int   shape;      /* area lights only, else -1 */
    float color[3];
    float pos[3];
    float dir[3];
} jpbrt_light;

/* ---- primitives: FPrimitive{shape, material, arealight} (primitive.h:23-25) ---- */
typedef struct jpbrt_primitive {
    int shape;     /* index into shapes[] */
    int material;  /* index into materials[], -1 = null material (pass-through, integrator.cc:349-353) */
    int light;     /* index into lights[] of the FAreaLight bound to this primitive, -1 = none */
} jpbrt_primitive;

/* ---- whole scene ---- */
typedef struct jpbrt_scene_desc {
    jpbrt_camera            camera;
    int                     max_depth;      /* FPathIntegratorIteration(maxDepth), main.cc:154 (default 5) */
    int                     n_shapes;
    int                     n_materials;
    int                     n_lights;
    int                     n_primitives;
    const jpbrt_shape*      shapes;
    const jpbrt_material*   materials;
    const jpbrt_light*      lights;
    const jpbrt_primitive*  primitives;
    const char*             name;           /* FScene name (scene.h:27), used for "<name>_<spp>" output files */
} jpbrt_scene_desc;

/* ---- BSDFs the reference implements but no FMaterial builds: reachable only through the
 * unit entry point jpbrt_unit_bsdf_ex, constructed exactly as their C++ constructors are. ---- */
typedef enum jpbrt_bsdf_kind {
    JPBRT_BSDF_PHONG = 0,                    /* FPhongSpecularReflection(frame, Ks, exponent)            bsdf.h:557-633 */
    JPBRT_BSDF_MICROFACET_REFLECTION = 1,    /* FMicrofacetReflection(frame, R, distribution, fresnel)   bsdf.cc:35-78 */
    JPBRT_BSDF_MICROFACET_TRANSMISSION = 2   /* FMicrofacetTransmission(frame, T, distribution, etaA, etaB) bsdf.cc:80-145 */
} jpbrt_bsdf_kind;
typedef enum jpbrt_distribution {
    JPBRT_DIST_BECKMANN = 0,                 /* BeckmannDistribution(alphax, alphay, samplevis)          microfacet.cc:11-254 */
    JPBRT_DIST_TROWBRIDGE_REITZ = 1          /* TrowbridgeReitzDistribution(alphax, alphay, samplevis)   microfacet.cc:181-357 */
} jpbrt_distribution;
typedef enum jpbrt_fresnel {
    JPBRT_FRESNEL_NOOP = 0,                  /* FresnelNoOp                       bsdf.h:666-669 */
    JPBRT_FRESNEL_DIELECTRIC = 1,            /* FresnelDielectric(eta_a, eta_b)   bsdf.h:656-664 */
    JPBRT_FRESNEL_CONDUCTOR = 2              /* FresnelConductor(c_eta_i, c_eta_t, c_k) bsdf.h:644-654 */
} jpbrt_fresnel;
typedef struct jpbrt_bsdf_desc {
    int   kind;                 /* jpbrt_bsdf_kind */
    int   distribution;         /* jpbrt_distribution (microfacet kinds) */
    int   sample_visible_area;  /* the distributions' samplevis flag */
    int   fresnel;              /* jpbrt_fresnel (microfacet reflection) */
    float color[3];             /* Ks | R | T */
    float exponent;             /* Phong */
    float alphax, alphay;
    float eta_a, eta_b;         /* dielectric Fresnel (etaI, etaT) | transmission (etaA, etaB) */
    float c_eta_i[3], c_eta_t[3], c_k[3];
} jpbrt_bsdf_desc;

/* Constants the reference hard-codes on the hot path (kept as constants, not knobs, so that the
 * compiled reference, the oracle and the CUDA path can never disagree about them). */
#define JPBRT_RAY_TMIN        0.001f  /* geometry.h:395,399 ; scene.h:38 */
#define JPBRT_RR_START_BOUNCE 3       /* integrator.cc:383 */
#define JPBRT_RR_QMIN         0.05f   /* integrator.cc:385 */
#define JPBRT_THINNESS        0.01f   /* geometry.h:299 */

#ifdef __cplusplus
}
#endif
#endif /* JETPBRT_SCENE_H */
