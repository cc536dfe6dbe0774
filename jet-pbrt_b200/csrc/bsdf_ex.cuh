// bsdf_ex.cuh -- the BSDF classes of the reference that no FMaterial builds, behind the unit entry
// point jpbrt_unit_bsdf_ex.  No scene can reach them; they exist so that the whole BSDF vocabulary of the reference
// (bsdf.h, bsdf.cc, microfacet.cc) has a device-side counterpart checked against the CPU classes.
//
// Organisation (round 2: rebuilt around compile-time composition instead of one switch per function):
//
//     slope distribution  Ndf<BECKMANN> | Ndf<TROWBRIDGE_REITZ>     D, Lambda, full-distribution and visible-normal sampling
//     Fresnel term        FresnelUnit | FresnelDielectricTerm | FresnelConductorTerm
//     lobe                RoughMirror<NDF, FRESNEL> | RoughGlass<NDF> | PhongLobe       eval / pdf / sample
//     visit_bsdf_ex()     builds the ONE composition a jpbrt_bsdf_desc names and hands it to a generic functor,
//
// so that every combination is its own straight-line code (no per-call distribution / Fresnel switches inside the
// formulas).  What could NOT be chosen freely is the order of the floating-point operations inside the formulas: the
// parity bar is 1e-5 RELATIVE on values as small as exp(-40), and the visible-normal Beckmann sample is by definition
// the iterate at which a particular safeguarded Newton iteration stops (|residual| < 1e-5), not the exact inverse of the
// CDF -- an exact solver misses it by up to 7e-3 (measured).  Those expressions therefore follow the reference's
// evaluation order and are marked "order fixed by parity"; their constants are the published ones (Giles' erfinv,
// Abramowitz-Stegun 7.1.26, the rational Lambda fit and initial-guess fit of Heitz / pbrt-v3, which the reference uses).
#pragma once

#include "../../include/jetpbrt_scene.h"
#include "bsdf.cuh"

namespace jpbrt {

// ---------------------------------------------------------------------------------------------------------------------
// Scalar building blocks
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ float horner_desc(const float (&coef)[N], float x) {  // coef[0] x^(N-1) + ... + coef[N-1]
    float acc = coef[0];
#pragma unroll
    for (int i = 1; i < N; ++i) acc = coef[i] + acc * x;
    return acc;
}

// erf^-1 on (-1, 1): Giles' two-branch single-precision polynomial in w = -log(1 - x^2)  (reference: microfacet.cc:11-40)
__device__ __forceinline__ float inverse_erf(float x) {
    const float kCentral[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                               -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
    const float kTail[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                            -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
    x = clampf(x, -.99999f, .99999f);
    const float w = -logf((1 - x) * (1 + x));
    const float poly = (w < 5) ? horner_desc(kCentral, w - 2.5f) : horner_desc(kTail, sqrtf(w) - 3);
    return poly * x;
}

// erf by Abramowitz & Stegun 7.1.26 (reference: microfacet.cc:42-63); order fixed by parity
__device__ __forceinline__ float erf_as7126(float x) {
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    const float ax = fabsf(x);
    const float t = 1 / (1 + p * ax);
    const float y = 1 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * expf(-ax * ax);
    return x < 0 ? -y : y;
}

__device__ __forceinline__ f3 unit_from_spherical(float sin_t, float cos_t, float phi) {  // geometry.h:203-209
    return mk3(sin_t * cosf(phi), sin_t * sinf(phi), cos_t);
}

struct Roughness {
    float x, y;  // alpha_x, alpha_y, clamped to >= 1e-3 like the reference's constructors (microfacet.h:58-59,79-80)
    __device__ __forceinline__ bool isotropic() const { return x == y; }
};

// Azimuth of a full-distribution sample of an ANISOTROPIC distribution (shared by both NDFs: microfacet.cc:214-218,337-339)
__device__ __forceinline__ float anisotropic_phi(const Roughness& a, float u1) {
    float phi = atanf(a.y / a.x * tanf(2 * JPB_PI * u1 + 0.5f * JPB_PI));
    if (u1 > 0.5f) phi += JPB_PI;
    return phi;
}
// cos^2 phi / ax^2 + sin^2 phi / ay^2: the inverse squared roughness seen along azimuth phi
__device__ __forceinline__ float inv_roughness2_along(const Roughness& a, float phi) {
    const float s = sinf(phi), c = cosf(phi);
    return c * c / (a.x * a.x) + s * s / (a.y * a.y);
}

// ---------------------------------------------------------------------------------------------------------------------
// Slope distributions
// ---------------------------------------------------------------------------------------------------------------------
template <int KIND>
struct Ndf;

template <>
struct Ndf<JPBRT_DIST_TROWBRIDGE_REITZ> {  // GGX: the hot path's own functions (bsdf.cuh), bit-exact with the reference
    Roughness a;
    __device__ __forceinline__ float D(const f3& m) const { return tr_D(a.x, a.y, m); }
    __device__ __forceinline__ float lambda(const f3& w) const { return tr_lambda(a.x, a.y, w); }
    __device__ __forceinline__ f3 sample_visible(const f3& wo, float u0, float u1) const { return tr_sample_wh(a.x, a.y, wo, u0, u1); }
    // cos(theta) of a full-distribution sample, q = 1 / alpha^2(phi); order fixed by parity (microfacet.cc:326-349)
    __device__ __forceinline__ float full_cos_theta(float q, float u0) const {
        const float alpha2 = 1 / q;
        return 1 / sqrtf(1 + alpha2 * u0 / (1 - u0));
    }
    __device__ __forceinline__ float full_cos_theta_iso(float u0) const { return 1 / sqrtf(1 + a.x * a.x * u0 / (1.0f - u0)); }
};

template <>
struct Ndf<JPBRT_DIST_BECKMANN> {
    Roughness a;
    // exp(-tan^2 (cos^2phi/ax^2 + sin^2phi/ay^2)) / (pi ax ay cos^4); order fixed by parity (microfacet.cc:172-179)
    __device__ __forceinline__ float D(const f3& m) const {
        const float t2 = tan2_theta(m);
        if (isinf(t2)) return 0.f;
        const float c2 = cos2_theta(m);
        return expf(-t2 * (cos2_phi(m) / (a.x * a.x) + sin2_phi(m) / (a.y * a.y))) / (JPB_PI * a.x * a.y * (c2 * c2));
    }
    // rational fit of the Smith Lambda in a = 1 / (alpha(phi) |tan theta|)   (microfacet.cc:191-200)
    __device__ __forceinline__ float lambda(const f3& w) const {
        const float t = fabsf(tan_theta(w));
        if (isinf(t)) return 0.f;
        const float alpha = sqrtf(cos2_phi(w) * a.x * a.x + sin2_phi(w) * a.y * a.y);
        const float inv = 1 / (alpha * t);
        return inv >= 1.6f ? 0.f : (1 - 1.259f * inv + 0.396f * inv * inv) / (3.535f * inv + 2.181f * inv * inv);
    }
    __device__ __forceinline__ float full_cos_theta(float q, float u0) const { return 1 / sqrtf(1 + (-logf(1 - u0) / q)); }  // (microfacet.cc:204-238)
    __device__ __forceinline__ float full_cos_theta_iso(float u0) const { return 1 / sqrtf(1 + (-a.x * a.x * logf(1 - u0))); }

    // Slopes of a visible normal for unit roughness and incidence cosine `mu` (microfacet.cc:66-144).  The x slope solves
    // N (1 + b + tan/sqrt(pi) exp(-erfinv(b)^2)) = u for b = erf(slope) in [-1, erf(cot)].  The RESULT the reference
    // returns is the iterate at which its bracketed Newton scheme stops (|residual| < 1e-5, at most 9 updates, started
    // from a fitted guess): reproducing it means running that scheme -- every line below is order fixed by parity.
    static __device__ __forceinline__ void unit_slopes(float mu, float u0, float u1, float& sx, float& sy) {
        if (mu > .9999f) {  // normal incidence: the slope distribution is radially symmetric
            const float r = sqrtf(-logf(1.0f - u0));
            const float s = sinf(2 * JPB_PI * u1), c = cosf(2 * JPB_PI * u1);
            sx = r * c;
            sy = r * s;
            return;
        }
        const float tan_i = sqrtf(std_max(0.f, 1.f - mu * mu)) / mu, cot_i = 1 / tan_i;
        const float target = std_max(u0, 1e-6f);
        const float inv_sqrt_pi = 1.f / sqrtf(JPB_PI);
        float lo = -1, hi = erf_as7126(cot_i);  // bracket of b
        const float angle = acosf(mu);
        const float guess_exponent = 1 + angle * (-0.876f + angle * (0.4265f - 0.0594f * angle));
        float b = hi - (1 + hi) * powf(1 - target, guess_exponent);
        const float norm = 1 / (1 + hi + inv_sqrt_pi * tan_i * expf(-cot_i * cot_i));
        for (int update = 1; update < 10; ++update) {
            if (!(b >= lo && b <= hi)) b = 0.5f * (lo + hi);  // Newton left the bracket: bisect
            const float slope = inverse_erf(b);
            const float residual = norm * (1 + b + inv_sqrt_pi * tan_i * expf(-slope * slope)) - target;
            if (fabsf(residual) < 1e-5f) break;
            const float gradient = norm * (1 - slope * tan_i);
            if (residual > 0) hi = b; else lo = b;
            b -= residual / gradient;
        }
        sx = inverse_erf(b);
        sy = inverse_erf(2.0f * std_max(u1, 1e-6f) - 1.0f);
    }
    // stretch -> sample unit slopes -> rotate -> unstretch -> normal   (microfacet.cc:146-170,240-254)
    __device__ __forceinline__ f3 sample_visible(const f3& wo, float u0, float u1) const {
        const bool below = wo.z < 0;
        const f3 w = below ? -wo : wo;
        const f3 stretched = normalize(mk3(a.x * w.x, a.y * w.y, w.z));
        float sx, sy;
        unit_slopes(stretched.z, u0, u1, sx, sy);
        const float cp = cos_phi(stretched), sp = sin_phi(stretched);
        const float rx = cp * sx - sp * sy, ry = sp * sx + cp * sy;
        const f3 m = normalize(mk3(-(a.x * rx), -(a.y * ry), 1.f));
        return below ? -m : m;
    }
};

// What both distributions share (microfacet.h:22-30, microfacet.cc:359-365)
template <class NDF>
struct Microsurface {
    NDF ndf;
    bool visible;  // sample / pdf of the VISIBLE normals (samplevis) or of the whole distribution
    __device__ __forceinline__ float D(const f3& m) const { return ndf.D(m); }
    __device__ __forceinline__ float G1(const f3& w) const { return 1 / (1 + ndf.lambda(w)); }
    __device__ __forceinline__ float G(const f3& wo, const f3& wi) const { return 1 / (1 + ndf.lambda(wo) + ndf.lambda(wi)); }
    __device__ __forceinline__ float pdf(const f3& wo, const f3& m) const {
        return visible ? D(m) * G1(wo) * absdot(wo, m) / fabsf(wo.z) : D(m) * fabsf(m.z);
    }
    __device__ __forceinline__ f3 sample(const f3& wo, float u0, float u1) const {
        if (visible) return ndf.sample_visible(wo, u0, u1);
        float cos_t, phi;
        if (ndf.a.isotropic()) {
            cos_t = ndf.full_cos_theta_iso(u0);
            phi = (2 * JPB_PI) * u1;
        } else {
            phi = anisotropic_phi(ndf.a, u1);
            cos_t = ndf.full_cos_theta(inv_roughness2_along(ndf.a, phi), u0);
        }
        const f3 m = unit_from_spherical(sqrtf(std_max(0.f, 1.f - cos_t * cos_t)), cos_t, phi);
        return same_hemisphere(wo, m) ? m : -m;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Fresnel terms (bsdf.h:646-669, bsdf.cc:15-24)
// ---------------------------------------------------------------------------------------------------------------------
struct FresnelUnit {
    __device__ __forceinline__ f3 operator()(float) const { return splat(1.f); }
};
struct FresnelDielectricTerm {
    float eta_outside, eta_inside;
    __device__ __forceinline__ f3 operator()(float cos_i) const { return splat(fresnel_dielectric(cos_i, eta_outside, eta_inside)); }
};
struct FresnelConductorTerm {
    f3 eta, k;  // already relative to the incident medium (bsdf.h:178-179)
    __device__ __forceinline__ f3 operator()(float cos_i) const { return fresnel_conductor(fabsf(cos_i), eta, k); }
};

// ---------------------------------------------------------------------------------------------------------------------
// Lobes.  All directions are in the shading frame; eval returns f, not f |cos|.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ BsdfSample no_sample() {
    BsdfSample s;
    s.f = mk3(0, 0, 0);
    s.wi = mk3(0, 0, 1);
    s.pdf = 0;
    s.flags = 0;
    return s;
}

template <class NDF, class FRESNEL>
struct RoughMirror {  // FMicrofacetReflection, bsdf.cc:35-78
    f3 tint;
    Microsurface<NDF> surf;
    FRESNEL fresnel;
    __device__ __forceinline__ f3 eval(const f3& wo, const f3& wi) const {
        const float co = fabsf(wo.z), ci = fabsf(wi.z);
        f3 m = wi + wo;
        if (ci == 0 || co == 0 || (m.x == 0 && m.y == 0 && m.z == 0)) return mk3(0, 0, 0);
        m = normalize(m);
        const f3 F = fresnel(dot(wi, face_forward(m, mk3(0, 0, 1))));
        return cmul(tint * surf.D(m) * surf.G(wo, wi), F) / (4 * ci * co);
    }
    __device__ __forceinline__ float pdf(const f3& wo, const f3& wi) const {
        if (!same_hemisphere(wo, wi)) return 0.f;
        const f3 m = normalize(wo + wi);
        return surf.pdf(wo, m) / (4 * dot(wo, m));
    }
    __device__ __forceinline__ BsdfSample sample(const Frame&, const f3& wo, float u0, float u1) const {
        BsdfSample s = no_sample();
        if (wo.z == 0) return s;
        const f3 m = surf.sample(wo, u0, u1);
        if (dot(wo, m) < 0) return s;
        const f3 wi = reflect(wo, m);
        if (!same_hemisphere(wo, wi)) return s;
        s.wi = wi;
        s.f = eval(wo, wi);
        s.pdf = surf.pdf(wo, m) / (4 * dot(wo, m));
        s.flags = BSDF_REFLECTION | BSDF_GLOSSY;
        return s;
    }
};

template <class NDF>
struct RoughGlass {  // FMicrofacetTransmission, bsdf.cc:80-145
    f3 tint;
    Microsurface<NDF> surf;
    float eta_outside, eta_inside;
    // index ratio seen by a ray arriving along wo, "transmitted over incident"
    __device__ __forceinline__ float ratio(const f3& wo) const { return wo.z > 0 ? (eta_inside / eta_outside) : (eta_outside / eta_inside); }
    // ... and "incident over transmitted", as refract() wants it (its own division: 1 / ratio rounds differently)
    __device__ __forceinline__ float inverse_ratio(const f3& wo) const { return wo.z > 0 ? (eta_outside / eta_inside) : (eta_inside / eta_outside); }
    // generalised half vector of a refraction pair; false where the pair is not a refraction through it
    __device__ __forceinline__ bool half_vector(const f3& wo, const f3& wi, float eta, bool flip_up, f3& m) const {
        m = normalize(wo + wi * eta);
        if (flip_up && m.z < 0) m = -m;
        return !(dot(wo, m) * dot(wi, m) > 0);
    }
    __device__ __forceinline__ f3 eval(const f3& wo, const f3& wi) const {
        if (same_hemisphere(wo, wi) || wi.z == 0 || wo.z == 0) return mk3(0, 0, 0);
        const float eta = ratio(wo);
        f3 m;
        if (!half_vector(wo, wi, eta, true, m)) return mk3(0, 0, 0);
        const float om = dot(wo, m), im = dot(wi, m);
        const f3 F = splat(fresnel_dielectric(om, eta_outside, eta_inside));
        const float jac = om + eta * im;
        const float inv_eta = 1 / eta;
        // order fixed by parity (bsdf.cc:104-110)
        return cmul(splat(1.f) - F, tint) *
               fabsf(surf.D(m) * surf.G(wo, wi) * eta * eta * fabsf(im) * fabsf(om) * inv_eta * inv_eta / (wi.z * wo.z * jac * jac));
    }
    __device__ __forceinline__ float pdf(const f3& wo, const f3& wi) const {
        if (same_hemisphere(wo, wi)) return 0.f;
        const float eta = ratio(wo);
        f3 m;
        if (!half_vector(wo, wi, eta, false, m)) return 0.f;
        const float jac = dot(wo, m) + eta * dot(wi, m);
        return surf.pdf(wo, m) * fabsf((eta * eta * dot(wi, m)) / (jac * jac));
    }
    // The reference's Sample_Local evaluates the WORLD-space Pdf() on its local vectors (bsdf.cc:140): the pair goes
    // through ToLocal() a second time.  Kept -- that is what a caller of the reference observes.
    __device__ __forceinline__ BsdfSample sample(const Frame& frame, const f3& wo, float u0, float u1) const {
        BsdfSample s = no_sample();
        if (wo.z == 0) return s;
        const f3 m = surf.sample(wo, u0, u1);
        if (dot(wo, m) < 0) return s;
        f3 wi;
        if (!refract(wo, m, inverse_ratio(wo), &wi)) return s;
        s.wi = wi;
        s.pdf = pdf(to_local(frame, wo), to_local(frame, wi));
        s.f = eval(wo, wi);
        s.flags = BSDF_TRANSMISSION | BSDF_GLOSSY;
        return s;
    }
};

struct PhongLobe {  // FPhongSpecularReflection, bsdf.h:557-633
    f3 tint;
    float exponent;
    static __device__ __forceinline__ f3 mirror_dir(const f3& wo) { return reflect(wo, mk3(0, 0, 1)); }
    __device__ __forceinline__ f3 eval(const f3& wo, const f3& wi) const {
        if (!same_hemisphere(wo, wi)) return mk3(0, 0, 0);
        const f3 scale = tint * (exponent + 2.f) * (1.0f / JPB_2PI);
        return scale * powf(dot(mirror_dir(wo), wi), exponent);
    }
    __device__ __forceinline__ float pdf(const f3& wo, const f3& wi) const {
        return (exponent + 1) * powf(std_max(0.f, dot(mirror_dir(wo), wi)), exponent) * (1.0f / JPB_2PI);
    }
    __device__ __forceinline__ BsdfSample sample(const Frame&, const f3& wo, float u0, float u1) const {
        BsdfSample s = no_sample();
        const float cos_a = powf(u1, 1.f / (exponent + 1));  // power-cosine lobe about the mirror direction
        const f3 about_mirror = unit_from_spherical(sqrtf(1.f - cos_a * cos_a), cos_a, 2 * JPB_PI * u0);
        s.wi = to_world(make_frame(mirror_dir(wo)), about_mirror);
        if (wo.z < 0) s.wi.z *= -1;
        s.f = eval(wo, s.wi);
        s.pdf = pdf(wo, s.wi);
        s.flags = BSDF_REFLECTION | BSDF_GLOSSY;
        return s;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Descriptor -> composition
// ---------------------------------------------------------------------------------------------------------------------
struct BsdfEx {
    jpbrt_bsdf_desc d;
};
__device__ __forceinline__ BsdfEx make_bsdf_ex(const jpbrt_bsdf_desc& d) { return BsdfEx{d}; }

template <class NDF, class FN>
__device__ __forceinline__ auto visit_rough(const jpbrt_bsdf_desc& d, const FN& fn) {
    const f3 tint = mk3(d.color[0], d.color[1], d.color[2]);
    const Microsurface<NDF> surf{NDF{Roughness{std_max(0.001f, d.alphax), std_max(0.001f, d.alphay)}}, d.sample_visible_area != 0};
    if (d.kind == JPBRT_BSDF_MICROFACET_TRANSMISSION) return fn(RoughGlass<NDF>{tint, surf, d.eta_a, d.eta_b});
    if (d.fresnel == JPBRT_FRESNEL_CONDUCTOR) {
        const f3 etai = mk3(d.c_eta_i[0], d.c_eta_i[1], d.c_eta_i[2]);
        const FresnelConductorTerm fr{cdiv(mk3(d.c_eta_t[0], d.c_eta_t[1], d.c_eta_t[2]), etai), cdiv(mk3(d.c_k[0], d.c_k[1], d.c_k[2]), etai)};
        return fn(RoughMirror<NDF, FresnelConductorTerm>{tint, surf, fr});
    }
    if (d.fresnel == JPBRT_FRESNEL_DIELECTRIC) return fn(RoughMirror<NDF, FresnelDielectricTerm>{tint, surf, FresnelDielectricTerm{d.eta_a, d.eta_b}});
    return fn(RoughMirror<NDF, FresnelUnit>{tint, surf, FresnelUnit{}});
}

// Calls fn(lobe) with the lobe the descriptor names; every branch instantiates fn for one concrete composition.
template <class FN>
__device__ __forceinline__ auto visit_bsdf_ex(const jpbrt_bsdf_desc& d, const FN& fn) {
    if (d.kind == JPBRT_BSDF_PHONG) return fn(PhongLobe{mk3(d.color[0], d.color[1], d.color[2]), d.exponent});
    if (d.distribution == JPBRT_DIST_BECKMANN) return visit_rough<Ndf<JPBRT_DIST_BECKMANN>>(d, fn);
    return visit_rough<Ndf<JPBRT_DIST_TROWBRIDGE_REITZ>>(d, fn);
}

__device__ __forceinline__ f3 bsdf_ex_eval_local(const BsdfEx& b, const f3& wo, const f3& wi) {
    return visit_bsdf_ex(b.d, [&](const auto& lobe) { return lobe.eval(wo, wi); });
}
__device__ __forceinline__ float bsdf_ex_pdf_local(const BsdfEx& b, const f3& wo, const f3& wi) {
    return visit_bsdf_ex(b.d, [&](const auto& lobe) { return lobe.pdf(wo, wi); });
}
__device__ __forceinline__ BsdfSample bsdf_ex_sample_local(const BsdfEx& b, const Frame& frame, const f3& wo, float u0, float u1) {
    return visit_bsdf_ex(b.d, [&](const auto& lobe) { return lobe.sample(frame, wo, u0, u1); });
}

}  // namespace jpbrt
