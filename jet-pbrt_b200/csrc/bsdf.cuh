// bsdf.cuh -- device restatement of the scattering code the default materials reach:
//   material->Scattering        material.h:27-137, material.cc:12-43
//   FBSDF::Evalf/Pdf/Sample     bsdf.h:285-302
//   FLambertionReflection       bsdf.h:336-385          FSpecularReflection  bsdf.h:394-435
//   FFresnelSpecular            bsdf.h:455-552          FMicrofacetReflection bsdf.cc:29-78
//   fresnel_dielectric/conductor bsdf.h:91-122,174-197  TrowbridgeReitz (visible normals) microfacet.cc:181-365
//   concentric/cosine sampling  sampling.h:25-64
// The reference heap-allocates a polymorphic BSDF (+ Fresnel + distribution) per hit; here a BSDF
// is a small tagged struct in registers.  Expressions keep the reference's evaluation order.
#pragma once

#include "dev_scene.h"
#include "dmath.cuh"

namespace jpbrt {

enum { BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_SPECULAR = 4, BSDF_DIFFUSE = 8, BSDF_GLOSSY = 16 };  // bsdf.h:208-219
enum { MAT_MATTE = 0, MAT_MIRROR = 1, MAT_GLASS = 2, MAT_PLASTIC = 3, MAT_METAL = 4 };
enum { K_LAMBERT = 0, K_SPECULAR = 1, K_FRESNEL_SPECULAR = 2, K_MICROFACET_CONDUCTOR = 3, K_MICROFACET_DIELECTRIC = 4 };

struct Bsdf {
    int kind;
    f3 R;          // albedo | reflectance | microfacet R
    f3 T;          // glass transmittance | conductor k
    f3 eta3;       // conductor eta (etaT / etaI with etaI = 1)
    float eta_i, eta_t;
    float ax, ay;  // Trowbridge-Reitz alpha (already clamped >= 0.001, microfacet.h:73-74)
};

struct BsdfSample {
    f3 f;
    f3 wi;
    float pdf;
    int flags;
};

__device__ __forceinline__ bool bsdf_is_delta(const Bsdf& b) { return b.kind == K_SPECULAR || b.kind == K_FRESNEL_SPECULAR; }

// material.cc / material.h Scattering(); `lobe_u` is the plastic material's sampler->GetFloat().
__device__ __forceinline__ Bsdf make_bsdf(const Float4* __restrict__ mat, float lobe_u) {
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mat));
    const float4 m1 = __ldg(reinterpret_cast<const float4*>(mat + 1));
    const float4 m2 = __ldg(reinterpret_cast<const float4*>(mat + 2));
    const int type = __float_as_int(m0.w);
    Bsdf b;
    b.R = mk3(m0);
    b.T = mk3(m1);
    b.eta3 = mk3(0, 0, 0);
    b.eta_i = 1.f;
    b.eta_t = 1.f;
    b.ax = b.ay = 1.f;
    switch (type) {
    case MAT_MATTE: b.kind = K_LAMBERT; break;
    case MAT_MIRROR: b.kind = K_SPECULAR; break;
    case MAT_GLASS: b.kind = K_FRESNEL_SPECULAR; b.eta_i = 1.f; b.eta_t = m1.w; break;  // material.h:72-75
    case MAT_PLASTIC:                                                                  // material.cc:12-29
        if (lobe_u < m2.y) { b.kind = K_LAMBERT; }                                     // R = Kd / Qd (precomputed)
        else { b.kind = K_MICROFACET_DIELECTRIC; b.R = b.T; b.eta_i = 1.5f; b.eta_t = 1.f; b.ax = b.ay = m1.w; }
        break;
    default:                                                                           // MAT_METAL, material.cc:31-43
        b.kind = K_MICROFACET_CONDUCTOR;
        b.eta3 = mk3(m0);
        b.R = mk3(1, 1, 1);
        b.ax = m1.w;
        b.ay = m2.x;
        break;
    }
    return b;
}

// ---- local-frame trigonometry, bsdf.h:17-60 ------------------------------------------------------
__device__ __forceinline__ float cos2_theta(const f3& w) { return w.z * w.z; }
__device__ __forceinline__ bool same_hemisphere(const f3& w, const f3& wp) { return w.z * wp.z > 0; }
__device__ __forceinline__ float sin2_theta(const f3& w) { return std_max(0.f, 1.f - cos2_theta(w)); }
__device__ __forceinline__ float sin_theta(const f3& w) { return sqrtf(sin2_theta(w)); }
__device__ __forceinline__ float tan_theta(const f3& w) { return sin_theta(w) / w.z; }
__device__ __forceinline__ float tan2_theta(const f3& w) { return sin2_theta(w) / cos2_theta(w); }
__device__ __forceinline__ float cos_phi(const f3& w) { float s = sin_theta(w); return (s == 0) ? 1.f : clampf(w.x / s, -1.f, 1.f); }
__device__ __forceinline__ float sin_phi(const f3& w) { float s = sin_theta(w); return (s == 0) ? 0.f : clampf(w.y / s, -1.f, 1.f); }
__device__ __forceinline__ float cos2_phi(const f3& w) { return cos_phi(w) * cos_phi(w); }
__device__ __forceinline__ float sin2_phi(const f3& w) { return sin_phi(w) * sin_phi(w); }
__device__ __forceinline__ f3 face_forward(const f3& v, const f3& v2) { return (dot(v, v2) < 0) ? -v : v; }
__device__ __forceinline__ f3 reflect(const f3& wo, const f3& n) { return -wo + (2 * dot(wo, n)) * n; }  // bsdf.h:62-67

__device__ __forceinline__ bool refract(const f3& wi, const f3& n, float eta, f3* wt) {  // bsdf.h:70-88
    float cos_i = dot(n, wi);
    float sin2_i = std_max(0.f, 1 - cos_i * cos_i);
    float sin2_t = eta * eta * sin2_i;
    if (sin2_t >= 1) return false;
    float cos_t = sqrtf(1 - sin2_t);
    *wt = eta * -wi + (eta * cos_i - cos_t) * n;
    return true;
}

__device__ __forceinline__ float fresnel_dielectric(float cos_i, float eta_i, float eta_t) {  // bsdf.h:91-122
    cos_i = clampf(cos_i, -1.f, 1.f);
    bool entering = cos_i > 0.f;
    if (!entering) { float t = eta_i; eta_i = eta_t; eta_t = t; cos_i = fabsf(cos_i); }
    float sin_i = sqrtf(std_max(0.f, 1 - cos_i * cos_i));
    float sin_t = eta_i / eta_t * sin_i;
    if (sin_t >= 1) return 1;
    float cos_t = sqrtf(std_max(0.f, 1 - sin_t * sin_t));
    float r_para = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
    float r_perp = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
    return (r_para * r_para + r_perp * r_perp) / 2;
}

// bsdf.h:174-197 with etai = (1,1,1): eta = etat / 1, etak = k / 1 (exact).
__device__ __forceinline__ f3 fresnel_conductor(float cosI, const f3& eta, const f3& etak) {
    cosI = clampf(cosI, -1.f, 1.f);
    float cosI2 = cosI * cosI;
    float sinI2 = 1 - cosI2;
    f3 eta2 = cmul(eta, eta);
    f3 etak2 = cmul(etak, etak);
    f3 t0 = eta2 - etak2 - splat(sinI2);
    f3 a2plusb2 = csqrt(cmul(t0, t0) + cmul(4 * eta2, etak2));
    f3 t1 = a2plusb2 + splat(cosI2);
    f3 a = csqrt(0.5f * (a2plusb2 + t0));
    f3 t2 = (2.f * cosI) * a;
    f3 Rs = cdiv(t1 - t2, t1 + t2);
    f3 t3 = cosI2 * a2plusb2 + splat(sinI2 * sinI2);
    f3 t4 = t2 * sinI2;
    f3 Rp = cdiv(cmul(Rs, t3 - t4), t3 + t4);
    return 0.5f * (Rp + Rs);
}

// ---- TrowbridgeReitzDistribution, microfacet.cc:181-189,202-210,359-365 ----------------------------
__device__ __forceinline__ float tr_D(float ax, float ay, const f3& wh) {
    float tan2 = tan2_theta(wh);
    if (isinf(tan2)) return 0.f;
    const float cos4 = cos2_theta(wh) * cos2_theta(wh);
    float e = (cos2_phi(wh) / (ax * ax) + sin2_phi(wh) / (ay * ay)) * tan2;
    return 1 / (JPB_PI * ax * ay * cos4 * (1 + e) * (1 + e));
}
__device__ __forceinline__ float tr_lambda(float ax, float ay, const f3& w) {
    float abs_tan = fabsf(tan_theta(w));
    if (isinf(abs_tan)) return 0.f;
    float alpha = sqrtf(cos2_phi(w) * ax * ax + sin2_phi(w) * ay * ay);
    float a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
    return (-1 + sqrtf(1.f + a2t2)) / 2;
}
__device__ __forceinline__ float tr_G1(float ax, float ay, const f3& w) { return 1 / (1 + tr_lambda(ax, ay, w)); }
__device__ __forceinline__ float tr_G(float ax, float ay, const f3& wo, const f3& wi) {
    return 1 / (1 + tr_lambda(ax, ay, wo) + tr_lambda(ax, ay, wi));
}
__device__ __forceinline__ float tr_pdf(float ax, float ay, const f3& wo, const f3& wh) {  // sampleVisibleArea = true
    return tr_D(ax, ay, wh) * tr_G1(ax, ay, wo) * absdot(wo, wh) / fabsf(wo.z);
}
// The same with Lambda(wo) supplied by the caller: at a shading vertex wo is fixed while the lights (and the sampled
// direction) vary, and Lambda(wo) -- two divisions and four square roots of exact arithmetic -- is the same number every
// time.  The expressions and their order are unchanged, so the values are bit-identical to the forms above.
__device__ __forceinline__ float tr_G(float ax, float ay, float lambda_o, const f3& wi) { return 1 / (1 + lambda_o + tr_lambda(ax, ay, wi)); }
__device__ __forceinline__ float tr_pdf(float ax, float ay, float lambda_o, const f3& wo, const f3& wh) {
    return tr_D(ax, ay, wh) * (1 / (1 + lambda_o)) * absdot(wo, wh) / fabsf(wo.z);
}

// TrowbridgeReitzSample11, microfacet.cc:256-303.  The reference's normal-incidence branch compares
// against a double literal and calls the double-precision ::cos/::sin; both are kept.
// The normal-incidence branch in double precision, out of line: it is taken for < 1 % of the samples but
// its two double-precision trigonometric expansions are several hundred instructions.
__device__ __noinline__ void tr_sample11_normal_incidence(float U1, float U2, float* slope_x, float* slope_y) {
    float r = sqrtf(U1 / (1 - U1));
    float phi = 6.28318530718f * U2;
    *slope_x = (float)((double)r * cos((double)phi));
    *slope_y = (float)((double)r * sin((double)phi));
}

__device__ __forceinline__ void tr_sample11(float cosTheta, float U1, float U2, float* slope_x, float* slope_y) {
    if ((double)cosTheta > .9999) {
        tr_sample11_normal_incidence(U1, U2, slope_x, slope_y);
        return;
    }
    float sinTheta = sqrtf(std_max(0.f, 1.f - cosTheta * cosTheta));
    float tanTheta = sinTheta / cosTheta;
    float a = 1 / tanTheta;
    float G1 = 2 / (1 + sqrtf(1.f + 1.f / (a * a)));
    float A = 2 * U1 / G1 - 1;
    float tmp = 1.f / (A * A - 1.f);
    if (tmp > 1e10f) tmp = 1e10f;
    float B = tanTheta;
    float D = sqrtf(std_max(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
    float slope_x_1 = B * tmp - D;
    float slope_x_2 = B * tmp + D;
    *slope_x = (A < 0 || slope_x_2 > 1.f / tanTheta) ? slope_x_1 : slope_x_2;
    float S;
    if (U2 > 0.5f) { S = 1.f; U2 = 2.f * (U2 - .5f); }
    else { S = -1.f; U2 = 2.f * (.5f - U2); }
    float z = (U2 * (U2 * (U2 * 0.27385f - 0.73369f) + 0.46341f)) /
              (U2 * (U2 * (U2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
    *slope_y = S * z * sqrtf(1.f + *slope_x * *slope_x);
}

__device__ __forceinline__ f3 tr_sample_wh(float ax, float ay, const f3& wo, float u0, float u1) {  // microfacet.cc:305-357
    bool flip = wo.z < 0;
    f3 wi = flip ? -wo : wo;
    f3 wiS = normalize(mk3(ax * wi.x, ay * wi.y, wi.z));
    float sx, sy;
    tr_sample11(wiS.z, u0, u1, &sx, &sy);
    float tmp = cos_phi(wiS) * sx - sin_phi(wiS) * sy;
    sy = sin_phi(wiS) * sx + cos_phi(wiS) * sy;
    sx = tmp;
    sx = ax * sx;
    sy = ay * sy;
    f3 wh = normalize(mk3(-sx, -sy, 1.f));
    if (flip) wh = -wh;
    return wh;
}

// ---- sampling.h:25-64 ------------------------------------------------------------------------------
__device__ __forceinline__ void concentric_disk_sample(float ux, float uy, float* px, float* py) {
    ux = ux * 2.f - 1;
    uy = uy * 2.f - 1;
    if (ux == 0 && uy == 0) { *px = 0; *py = 0; return; }
    float radius, theta;
    if (fabsf(ux) > fabsf(uy)) { radius = ux; theta = JPB_PI_OVER_4 * (uy / ux); }
    else { radius = uy; theta = JPB_PI_OVER_2 - JPB_PI_OVER_4 * (ux / uy); }
    *px = jp_cosf(theta) * radius;
    *py = jp_sinf(theta) * radius;
}
__device__ __forceinline__ f3 cosine_hemisphere_sample(float ux, float uy) {
    float px, py;
    concentric_disk_sample(ux, uy, &px, &py);
    float z = sqrtf(std_max(0.f, 1 - px * px - py * py));
    return mk3(px, py, z);
}

// ---- Evalf_Local / Pdf_Local / Sample_Local --------------------------------------------------------
__device__ __forceinline__ f3 bsdf_fresnel(const Bsdf& b, float cosI) {  // bsdf.cc:15-24
    if (b.kind == K_MICROFACET_CONDUCTOR) return fresnel_conductor(fabsf(cosI), b.eta3, b.T);
    return splat(fresnel_dielectric(cosI, b.eta_i, b.eta_t));
}

// FMicrofacetReflection::Evalf_Local, bsdf.cc:35-50.  Out of line: it is needed by NEE and by Sample_Local.
template <bool CONDUCTOR>
__device__ __noinline__ f3 microfacet_eval(const Bsdf& b, const f3& wo, const f3& wi, float lambda_o) {
    float cosO = fabsf(wo.z), cosI = fabsf(wi.z);
    f3 wh = wi + wo;
    if (cosI == 0 || cosO == 0) return mk3(0, 0, 0);
    if (wh.x == 0 && wh.y == 0 && wh.z == 0) return mk3(0, 0, 0);
    wh = normalize(wh);
    const float cosF = dot(wi, face_forward(wh, mk3(0, 0, 1)));
    f3 F = CONDUCTOR ? fresnel_conductor(fabsf(cosF), b.eta3, b.T) : splat(fresnel_dielectric(cosF, b.eta_i, b.eta_t));  // bsdf.cc:15-24
    return cmul(b.R * tr_D(b.ax, b.ay, wh) * tr_G(b.ax, b.ay, lambda_o, wi), F) / (4 * cosI * cosO);
}

// Lambda(wo) of a microfacet BSDF (0 for the others): computed once per shading vertex, see tr_G above.
__device__ __forceinline__ float bsdf_lambda_o(const Bsdf& b, const f3& wo) {
    return (b.kind == K_MICROFACET_CONDUCTOR || b.kind == K_MICROFACET_DIELECTRIC) ? tr_lambda(b.ax, b.ay, wo) : 0.f;
}

__device__ __forceinline__ f3 bsdf_eval_local(const Bsdf& b, const f3& wo, const f3& wi, float lambda_o) {
    if (b.kind == K_LAMBERT) {  // bsdf.h:347-355
        if (!same_hemisphere(wo, wi)) return mk3(0, 0, 0);
        return b.R * JPB_INV_PI;
    }
    if (b.kind == K_MICROFACET_CONDUCTOR) return microfacet_eval<true>(b, wo, wi, lambda_o);
    if (b.kind == K_MICROFACET_DIELECTRIC) return microfacet_eval<false>(b, wo, wi, lambda_o);
    return mk3(0, 0, 0);  // delta BSDFs, bsdf.h:405-408,468-471
}
__device__ __forceinline__ f3 bsdf_eval_local(const Bsdf& b, const f3& wo, const f3& wi) { return bsdf_eval_local(b, wo, wi, bsdf_lambda_o(b, wo)); }

__device__ __forceinline__ float bsdf_pdf_local(const Bsdf& b, const f3& wo, const f3& wi) {
    if (b.kind == K_LAMBERT) return same_hemisphere(wo, wi) ? fabsf(wi.z) * JPB_INV_PI : 0.f;  // bsdf.h:357-360
    if (b.kind >= K_MICROFACET_CONDUCTOR) {  // bsdf.cc:52-57
        if (!same_hemisphere(wo, wi)) return 0.f;
        f3 wh = normalize(wo + wi);
        return tr_pdf(b.ax, b.ay, wo, wh) / (4 * dot(wo, wh));
    }
    return 0.f;
}

__device__ __forceinline__ BsdfSample bsdf_sample_local(const Bsdf& b, const f3& wo, float u0, float u1, float lambda_o) {
    BsdfSample s;
    s.f = mk3(0, 0, 0);
    s.wi = mk3(0, 0, 1);
    s.pdf = 0.f;
    s.flags = 0;
    if (b.kind == K_LAMBERT) {  // bsdf.h:362-377
        s.wi = cosine_hemisphere_sample(u0, u1);
        if (wo.z < 0) s.wi.z *= -1;
        s.f = bsdf_eval_local(b, wo, s.wi);
        s.pdf = bsdf_pdf_local(b, wo, s.wi);
        s.flags = BSDF_REFLECTION | BSDF_DIFFUSE;
        return s;
    }
    if (b.kind == K_SPECULAR) {  // bsdf.h:415-430
        s.wi = mk3(-wo.x, -wo.y, wo.z);
        s.f = b.R / fabsf(s.wi.z);
        s.pdf = 1;
        s.flags = BSDF_REFLECTION | BSDF_SPECULAR;
        return s;
    }
    if (b.kind == K_FRESNEL_SPECULAR) {  // bsdf.h:478-540
        if (wo.z == 0.f) return s;
        float F = fresnel_dielectric(wo.z, b.eta_i, b.eta_t);
        if (u0 < F) {
            s.wi = mk3(-wo.x, -wo.y, wo.z);
            s.pdf = F;
            s.f = (b.R * F) / fabsf(s.wi.z);
            s.flags = BSDF_REFLECTION | BSDF_SPECULAR;
        } else {
            bool entering = wo.z > 0;
            f3 wn = entering ? mk3(0, 0, 1) : mk3(-0.f, -0.f, -1.f);
            float etaI = entering ? b.eta_i : b.eta_t;
            float etaT = entering ? b.eta_t : b.eta_i;
            f3 wt;
            if (refract(wo, wn, etaI / etaT, &wt)) {
                s.wi = wt;
                f3 ft = b.T * (1 - F);
                ft = ft * ((etaI * etaI) / (etaT * etaT));
                s.pdf = 1 - F;
                s.f = ft / fabsf(s.wi.z);
                s.flags = BSDF_TRANSMISSION | BSDF_SPECULAR;
            }
        }
        return s;
    }
    // microfacet reflection, bsdf.cc:59-78
    if (wo.z == 0) return s;
    f3 wh = tr_sample_wh(b.ax, b.ay, wo, u0, u1);
    if (dot(wo, wh) < 0) return s;
    f3 wi = reflect(wo, wh);
    if (!same_hemisphere(wo, wi)) return s;
    s.wi = wi;
    s.f = bsdf_eval_local(b, wo, wi, lambda_o);
    s.pdf = tr_pdf(b.ax, b.ay, lambda_o, wo, wh) / (4 * dot(wo, wh));
    s.flags = BSDF_REFLECTION | BSDF_GLOSSY;
    return s;
}
__device__ __forceinline__ BsdfSample bsdf_sample_local(const Bsdf& b, const f3& wo, float u0, float u1) {
    return bsdf_sample_local(b, wo, u0, u1, bsdf_lambda_o(b, wo));
}

}  // namespace jpbrt
