// bsdf_ex.cuh -- the BSDF classes of the reference that no FMaterial builds, as device
// functions behind the unit entry point jpbrt_unit_bsdf_ex:
//   BeckmannDistribution (D, Lambda, Sample_wh with BeckmannSample11 / ErfInv / Erf)        microfacet.cc:11-254
//   full-distribution sampling (samplevis = false) of both distributions                   microfacet.cc:204-238,326-349
//   FMicrofacetReflection with FresnelNoOp / FresnelDielectric / FresnelConductor           bsdf.cc:35-78
//   FMicrofacetTransmission                                                                 bsdf.cc:80-145
//   FPhongSpecularReflection                                                                bsdf.h:557-633
// Same expressions, same order, -fmad=false.  Unlike the hot path's BSDFs these lean on expf / logf / powf / acosf /
// tanf / atanf, whose CUDA and glibc results differ in the last ulp, and BeckmannSample11 is a Newton iteration that
// stops on |value| < 1e-5: parity is 1e-5 relative except where that tolerance itself decides (tests flag those).
#pragma once

#include "../../include/jetpbrt_scene.h"
#include "bsdf.cuh"

namespace jpbrt {

#define JPB_INV_2PI (1.0f / JPB_2PI) /* pbrt.h:45 */

__device__ __forceinline__ float erf_inv(float x) {  // microfacet.cc:11-40
    float w, p;
    x = clampf(x, -.99999f, .99999f);
    w = -logf((1 - x) * (1 + x));
    if (w < 5) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = 3.43273939e-07f + p * w;
        p = -3.5233877e-06f + p * w;
        p = -4.39150654e-06f + p * w;
        p = 0.00021858087f + p * w;
        p = -0.00125372503f + p * w;
        p = -0.00417768164f + p * w;
        p = 0.246640727f + p * w;
        p = 1.50140941f + p * w;
    } else {
        w = sqrtf(w) - 3;
        p = -0.000200214257f;
        p = 0.000100950558f + p * w;
        p = 0.00134934322f + p * w;
        p = -0.00367342844f + p * w;
        p = 0.00573950773f + p * w;
        p = -0.0076224613f + p * w;
        p = 0.00943887047f + p * w;
        p = 1.00167406f + p * w;
        p = 2.83297682f + p * w;
    }
    return p * x;
}

__device__ __forceinline__ float erf_as(float x) {  // microfacet.cc:42-63 (Abramowitz & Stegun 7.1.26)
    const float a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f, a5 = 1.061405429f, p = 0.3275911f;
    int sign = 1;
    if (x < 0) sign = -1;
    x = fabsf(x);
    float t = 1 / (1 + p * x);
    float y = 1 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * expf(-x * x);
    return sign * y;
}

struct DistEx {
    int type;
    bool vis;
    float ax, ay;
};

__device__ __forceinline__ float dist_D(const DistEx& d, const f3& wh) {
    if (d.type == JPBRT_DIST_TROWBRIDGE_REITZ) return tr_D(d.ax, d.ay, wh);
    float tan2 = tan2_theta(wh);  // microfacet.cc:172-179
    if (isinf(tan2)) return 0.f;
    float cos4 = cos2_theta(wh) * cos2_theta(wh);
    return expf(-tan2 * (cos2_phi(wh) / (d.ax * d.ax) + sin2_phi(wh) / (d.ay * d.ay))) / (JPB_PI * d.ax * d.ay * cos4);
}
__device__ __forceinline__ float dist_lambda(const DistEx& d, const f3& w) {
    if (d.type == JPBRT_DIST_TROWBRIDGE_REITZ) return tr_lambda(d.ax, d.ay, w);
    float abs_tan = fabsf(tan_theta(w));  // microfacet.cc:191-200
    if (isinf(abs_tan)) return 0.f;
    float alpha = sqrtf(cos2_phi(w) * d.ax * d.ax + sin2_phi(w) * d.ay * d.ay);
    float a = 1 / (alpha * abs_tan);
    if (a >= 1.6f) return 0.f;
    return (1 - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
__device__ __forceinline__ float dist_G1(const DistEx& d, const f3& w) { return 1 / (1 + dist_lambda(d, w)); }
__device__ __forceinline__ float dist_G(const DistEx& d, const f3& wo, const f3& wi) { return 1 / (1 + dist_lambda(d, wo) + dist_lambda(d, wi)); }
__device__ __forceinline__ float dist_pdf(const DistEx& d, const f3& wo, const f3& wh) {  // microfacet.cc:359-365
    if (d.vis) return dist_D(d, wh) * dist_G1(d, wo) * absdot(wo, wh) / fabsf(wo.z);
    return dist_D(d, wh) * fabsf(wh.z);
}

__device__ __forceinline__ void beckmann_sample11(float cosThetaI, float U1, float U2, float* slope_x, float* slope_y) {  // microfacet.cc:66-144
    if (cosThetaI > .9999f) {
        float r = sqrtf(-logf(1.0f - U1));
        float sinPhi = sinf(2 * JPB_PI * U2);
        float cosPhi = cosf(2 * JPB_PI * U2);
        *slope_x = r * cosPhi;
        *slope_y = r * sinPhi;
        return;
    }
    float sinThetaI = sqrtf(std_max(0.f, 1.f - cosThetaI * cosThetaI));
    float tanThetaI = sinThetaI / cosThetaI;
    float cotThetaI = 1 / tanThetaI;
    float a = -1, c = erf_as(cotThetaI);
    float sample_x = std_max(U1, 1e-6f);
    float thetaI = acosf(cosThetaI);
    float fit = 1 + thetaI * (-0.876f + thetaI * (0.4265f - 0.0594f * thetaI));
    float b = c - (1 + c) * powf(1 - sample_x, fit);
    const float SQRT_PI_INV = 1.f / sqrtf(JPB_PI);
    float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * expf(-cotThetaI * cotThetaI));
    int it = 0;
    while (++it < 10) {
        if (!(b >= a && b <= c)) b = 0.5f * (a + c);
        float invErf = erf_inv(b);
        float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * expf(-invErf * invErf)) - sample_x;
        float derivative = normalization * (1 - invErf * tanThetaI);
        if (fabsf(value) < 1e-5f) break;
        if (value > 0) c = b;
        else a = b;
        b -= value / derivative;
    }
    *slope_x = erf_inv(b);
    *slope_y = erf_inv(2.0f * std_max(U2, 1e-6f) - 1.0f);
}

__device__ __forceinline__ f3 dist_sample_wh(const DistEx& d, const f3& wo, float u0, float u1) {
    if (!d.vis) {  // full distribution of normals
        float cosTheta, phi;
        if (d.type == JPBRT_DIST_BECKMANN) {  // microfacet.cc:204-238
            float tan2Theta;
            if (d.ax == d.ay) {
                float logSample = logf(1 - u0);
                tan2Theta = -d.ax * d.ax * logSample;
                phi = u1 * 2 * JPB_PI;
            } else {
                float logSample = logf(1 - u0);
                phi = atanf(d.ay / d.ax * tanf(2 * JPB_PI * u1 + 0.5f * JPB_PI));
                if (u1 > 0.5f) phi += JPB_PI;
                float sinPhi = sinf(phi), cosPhi = cosf(phi);
                float ax2 = d.ax * d.ax, ay2 = d.ay * d.ay;
                tan2Theta = -logSample / (cosPhi * cosPhi / ax2 + sinPhi * sinPhi / ay2);
            }
            cosTheta = 1 / sqrtf(1 + tan2Theta);
        } else {  // microfacet.cc:326-349
            cosTheta = 0;
            phi = (2 * JPB_PI) * u1;
            if (d.ax == d.ay) {
                float tanTheta2 = d.ax * d.ax * u0 / (1.0f - u0);
                cosTheta = 1 / sqrtf(1 + tanTheta2);
            } else {
                phi = atanf(d.ay / d.ax * tanf(2 * JPB_PI * u1 + .5f * JPB_PI));
                if (u1 > .5f) phi += JPB_PI;
                float sinPhi = sinf(phi), cosPhi = cosf(phi);
                const float ax2 = d.ax * d.ax, ay2 = d.ay * d.ay;
                const float alpha2 = 1 / (cosPhi * cosPhi / ax2 + sinPhi * sinPhi / ay2);
                float tanTheta2 = alpha2 * u0 / (1 - u0);
                cosTheta = 1 / sqrtf(1 + tanTheta2);
            }
        }
        float sinTheta = sqrtf(std_max(0.f, 1.f - cosTheta * cosTheta));
        f3 wh = mk3(sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta);  // Spherical_2_Direction, geometry.h:203-209
        if (!same_hemisphere(wo, wh)) wh = -wh;
        return wh;
    }
    if (d.type == JPBRT_DIST_TROWBRIDGE_REITZ) return tr_sample_wh(d.ax, d.ay, wo, u0, u1);
    bool flip = wo.z < 0;  // microfacet.cc:240-254, BeckmannSample :146-170
    f3 wi = flip ? -wo : wo;
    f3 wiS = normalize(mk3(d.ax * wi.x, d.ay * wi.y, wi.z));
    float sx, sy;
    beckmann_sample11(wiS.z, u0, u1, &sx, &sy);
    float tmp = cos_phi(wiS) * sx - sin_phi(wiS) * sy;
    sy = sin_phi(wiS) * sx + cos_phi(wiS) * sy;
    sx = tmp;
    sx = d.ax * sx;
    sy = d.ay * sy;
    f3 wh = normalize(mk3(-sx, -sy, 1.f));
    if (flip) wh = -wh;
    return wh;
}

struct BsdfEx {
    jpbrt_bsdf_desc d;
    DistEx dist;
    f3 color;
};

__device__ __forceinline__ BsdfEx make_bsdf_ex(const jpbrt_bsdf_desc& d) {
    BsdfEx b;
    b.d = d;
    b.dist.type = d.distribution;
    b.dist.vis = d.sample_visible_area != 0;
    b.dist.ax = std_max(0.001f, d.alphax);  // microfacet.h:58-59,79-80
    b.dist.ay = std_max(0.001f, d.alphay);
    b.color = mk3(d.color[0], d.color[1], d.color[2]);
    return b;
}

__device__ __forceinline__ f3 bsdf_ex_fresnel(const BsdfEx& b, float cosI) {  // bsdf.cc:15-24, bsdf.h:666-669
    if (b.d.fresnel == JPBRT_FRESNEL_CONDUCTOR) {
        const f3 etai = mk3(b.d.c_eta_i[0], b.d.c_eta_i[1], b.d.c_eta_i[2]);
        const f3 eta = cdiv(mk3(b.d.c_eta_t[0], b.d.c_eta_t[1], b.d.c_eta_t[2]), etai);  // bsdf.h:178-179
        const f3 etak = cdiv(mk3(b.d.c_k[0], b.d.c_k[1], b.d.c_k[2]), etai);
        return fresnel_conductor(fabsf(cosI), eta, etak);
    }
    if (b.d.fresnel == JPBRT_FRESNEL_DIELECTRIC) return splat(fresnel_dielectric(cosI, b.d.eta_a, b.d.eta_b));
    return splat(1.f);
}

__device__ __forceinline__ f3 bsdf_ex_eval_local(const BsdfEx& b, const f3& wo, const f3& wi) {
    switch (b.d.kind) {
    case JPBRT_BSDF_PHONG: {  // bsdf.h:570-581
        if (!same_hemisphere(wo, wi)) return mk3(0, 0, 0);
        const f3 wr = reflect(wo, mk3(0, 0, 1));
        const float cos_alpha = dot(wr, wi);
        const f3 rho = b.color * (b.d.exponent + 2.f) * JPB_INV_2PI;
        return rho * powf(cos_alpha, b.d.exponent);
    }
    case JPBRT_BSDF_MICROFACET_REFLECTION: {  // bsdf.cc:35-50
        float cosO = fabsf(wo.z), cosI = fabsf(wi.z);
        f3 wh = wi + wo;
        if (cosI == 0 || cosO == 0) return mk3(0, 0, 0);
        if (wh.x == 0 && wh.y == 0 && wh.z == 0) return mk3(0, 0, 0);
        wh = normalize(wh);
        f3 F = bsdf_ex_fresnel(b, dot(wi, face_forward(wh, mk3(0, 0, 1))));
        return cmul(b.color * dist_D(b.dist, wh) * dist_G(b.dist, wo, wi), F) / (4 * cosI * cosO);
    }
    default: {  // FMicrofacetTransmission, bsdf.cc:85-111
        if (same_hemisphere(wo, wi)) return mk3(0, 0, 0);
        float cosO = wo.z, cosI = wi.z;
        if (cosI == 0 || cosO == 0) return mk3(0, 0, 0);
        float eta = wo.z > 0 ? (b.d.eta_b / b.d.eta_a) : (b.d.eta_a / b.d.eta_b);
        f3 wh = normalize(wo + wi * eta);
        if (wh.z < 0) wh = -wh;
        if (dot(wo, wh) * dot(wi, wh) > 0) return mk3(0, 0, 0);
        f3 F = splat(fresnel_dielectric(dot(wo, wh), b.d.eta_a, b.d.eta_b));
        float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
        float factor = (1 / eta);
        return cmul(splat(1.f) - F, b.color) *
               fabsf(dist_D(b.dist, wh) * dist_G(b.dist, wo, wi) * eta * eta * absdot(wi, wh) * absdot(wo, wh) * factor * factor /
                     (cosI * cosO * sqrtDenom * sqrtDenom));
    }
    }
}

__device__ __forceinline__ float bsdf_ex_pdf_local(const BsdfEx& b, const f3& wo, const f3& wi) {
    switch (b.d.kind) {
    case JPBRT_BSDF_PHONG: {  // bsdf.h:583-589, 621-625
        const f3 wr = reflect(wo, mk3(0, 0, 1));
        const float cosTheta = std_max(0.f, dot(wr, wi));
        return (b.d.exponent + 1) * powf(cosTheta, b.d.exponent) * JPB_INV_2PI;
    }
    case JPBRT_BSDF_MICROFACET_REFLECTION: {  // bsdf.cc:52-57
        if (!same_hemisphere(wo, wi)) return 0.f;
        f3 wh = normalize(wo + wi);
        return dist_pdf(b.dist, wo, wh) / (4 * dot(wo, wh));
    }
    default: {  // bsdf.cc:113-126
        if (same_hemisphere(wo, wi)) return 0.f;
        float eta = wo.z > 0 ? (b.d.eta_b / b.d.eta_a) : (b.d.eta_a / b.d.eta_b);
        f3 wh = normalize(wo + wi * eta);
        if (dot(wo, wh) * dot(wi, wh) > 0) return 0.f;
        float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
        float dwh_dwi = fabsf((eta * eta * dot(wi, wh)) / (sqrtDenom * sqrtDenom));
        return dist_pdf(b.dist, wo, wh) * dwh_dwi;
    }
    }
}

// `frame` is needed by the transmission lobe only: the reference's Sample_Local calls the WORLD-space Pdf() on its
// local vectors (bsdf.cc:140), i.e. the directions go through ToLocal() a second time.  Kept.
__device__ __forceinline__ BsdfSample bsdf_ex_sample_local(const BsdfEx& b, const Frame& frame, const f3& wo, float u0, float u1) {
    BsdfSample s;
    s.f = mk3(0, 0, 0);
    s.wi = mk3(0, 0, 1);
    s.pdf = 0;
    s.flags = 0;
    switch (b.d.kind) {
    case JPBRT_BSDF_PHONG: {  // bsdf.h:591-619
        const float phi = 2 * JPB_PI * u0;
        const float cos_t = powf(u1, 1.f / (b.d.exponent + 1));
        const float sin_t = sqrtf(1.f - cos_t * cos_t);
        const f3 l = mk3(cosf(phi) * sin_t, sinf(phi) * sin_t, cos_t);
        const f3 wr = reflect(wo, mk3(0, 0, 1));
        const Frame fr = make_frame(wr);
        s.wi = to_world(fr, l);
        if (wo.z < 0) s.wi.z *= -1;
        s.f = bsdf_ex_eval_local(b, wo, s.wi);
        s.pdf = bsdf_ex_pdf_local(b, wo, s.wi);
        s.flags = BSDF_REFLECTION | BSDF_GLOSSY;
        return s;
    }
    case JPBRT_BSDF_MICROFACET_REFLECTION: {  // bsdf.cc:59-78
        if (wo.z == 0) return s;
        f3 wh = dist_sample_wh(b.dist, wo, u0, u1);
        if (dot(wo, wh) < 0) return s;
        f3 wi = reflect(wo, wh);
        if (!same_hemisphere(wo, wi)) return s;
        s.wi = wi;
        s.f = bsdf_ex_eval_local(b, wo, wi);
        s.pdf = dist_pdf(b.dist, wo, wh) / (4 * dot(wo, wh));
        s.flags = BSDF_REFLECTION | BSDF_GLOSSY;
        return s;
    }
    default: {  // bsdf.cc:128-145
        if (wo.z == 0) return s;
        f3 wh = dist_sample_wh(b.dist, wo, u0, u1);
        if (dot(wo, wh) < 0) return s;
        f3 wi;
        float eta = wo.z > 0 ? (b.d.eta_a / b.d.eta_b) : (b.d.eta_b / b.d.eta_a);
        if (!refract(wo, wh, eta, &wi)) return s;
        s.wi = wi;
        s.pdf = bsdf_ex_pdf_local(b, to_local(frame, wo), to_local(frame, wi));
        s.f = bsdf_ex_eval_local(b, wo, wi);
        s.flags = BSDF_TRANSMISSION | BSDF_GLOSSY;
        return s;
    }
    }
}

}  // namespace jpbrt
