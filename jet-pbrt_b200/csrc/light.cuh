// light.cuh -- device restatement of FLight::Sample_Li and emission:
//   FAreaLight::Sample_Li / L        light.h:199-216, 234-238
//   FShape::SampleDirection          shape.h:124-145      FSphere::SampleDirection shape.h:564-644
//   F{Triangle,Rectangle,Sphere,Disk}::SamplePosition  shape.h:353-363, 459-467, 549-561, 257-268
//   FEnvironmentLight::Sample_Li     light.h:265-287      FPointLight light.h:95-124   FDirectionLight light.h:153-162
#pragma once

#include "bsdf.cuh"
#include "dev_scene.h"
#include "dmath.cuh"

namespace jpbrt {

enum { LIGHT_ENV = 0, LIGHT_AREA = 1, LIGHT_POINT = 2, LIGHT_DIR = 3 };

struct LightSample {
    f3 pos, wi, Li;
    float pdf;
    // dist > 0: wi IS (pos - P) / dist with dist = length(pos - P), bit for bit -- the direction and distance of the shadow ray
    // (FScene::Occluded recomputes exactly these, scene.h:36-47), so that k_shade does not normalise the same vector again
    float dist;
};

__device__ __forceinline__ f3 uniform_sphere_sample(float ux, float uy) {  // sampling.h:80-87
    float z = 1 - 2 * ux;
    float radius = sqrtf(std_max(0.f, 1.f - z * z));
    float phi = JPB_2PI * uy;
    return mk3(radius * jp_cosf(phi), radius * jp_sinf(phi), z);
}

// FAreaLight::L, light.h:234-238
__device__ __forceinline__ f3 area_L(const f3& radiance, const f3& light_normal, const f3& wo) {
    return (dot(light_normal, wo) > 0.f) ? radiance : mk3(0, 0, 0);
}

__device__ __forceinline__ LightSample sample_light(const DevScene& sc, int li, const f3& P, const f3& N, float ux, float uy) {
    const Float4* L = sc.lights + (size_t)li * kLightStride;
    const float4 l0 = ldg4(L);
    const int tagv = __float_as_int(l0.w);
    const int type = tagv & 0xff, shape_type = tagv >> 8;
    const f3 color = mk3(l0);
    LightSample s;
    s.pos = mk3(0, 0, 0);
    s.wi = mk3(0, 0, 0);
    s.Li = mk3(0, 0, 0);
    s.pdf = 0.f;
    s.dist = 0.f;
    if (type == LIGHT_ENV) {  // light.h:265-287
        float theta = uy * JPB_PI, phi = ux * 2 * JPB_PI;
        float cosTheta = jp_cosf(theta), sinTheta = jp_sinf(theta);
        float sinPhi = jp_sinf(phi), cosPhi = jp_cosf(phi);
        s.wi = mk3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        s.pos = P + s.wi * 2 * sc.world_radius;
        if (sinTheta != 0) s.pdf = 1 / (2 * JPB_PI * JPB_PI * sinTheta);
        s.Li = color;
        return s;
    }
    if (type == LIGHT_POINT) {  // light.h:95-124
        const f3 lp = mk3(ldg4(L + 1));
        s.pos = lp;
        const f3 w = lp - P;
        const float d2 = length2(w);
        s.dist = sqrtf(d2);
        s.wi = w / s.dist;  // normalize(lp - P)
        s.pdf = 1.f;
        s.Li = color / d2;
        return s;
    }
    if (type == LIGHT_DIR) {  // light.h:153-162
        const f3 wd = mk3(ldg4(L + 1));
        s.wi = -wd;
        s.pos = P + s.wi * 2 * sc.world_radius;
        s.pdf = 1.f;
        s.Li = color;
        return s;
    }
    // ---- area light: shape->SampleDirection(isect, u, &pdf) ----
    const float4 l1 = ldg4(L + 1), l2 = ldg4(L + 2);
    const f3 p0 = mk3(l1);
    const float inv_area = l1.w;  // 1 / Area()
    f3 lpos, lnrm;
    float pdf;
    // The reference normalises (lpos - P) in SampleDirection, again in Sample_Li and a third time in FScene::Occluded: the same
    // expression on the same operands -- computed once here (wn, dist, dist2) and reused, values unchanged.
    f3 wn = mk3(0, 0, 0);
    float dist = 0.f, dist2 = -1.f;  // dist2 < 0: (lpos - P) not looked at yet
    if (shape_type == SHAPE_SPHERE) {
        const float radius = l2.w;
        const f3 dcp = P - p0;
        if (length2(dcp) <= radius * radius) {  // inside or on the sphere: shape.h:567-588
            f3 dir = uniform_sphere_sample(ux, uy);
            lpos = p0 + radius * dir;
            lnrm = normalize(dir);
            pdf = inv_area;
            const f3 w = lpos - P;
            dist2 = length2(w);
            if (dist2 == 0) pdf = 0;
            else {
                dist = sqrtf(dist2);
                wn = w / dist;  // normalize(w)
                pdf *= dist2 / absdot(N, -wn);
            }
            if (isinf(pdf)) pdf = 0;
        } else {  // cone sampling: shape.h:606-643
            float dist = length(dcp);
            float inv_dist = 1 / dist;
            float sin_theta_max = radius * inv_dist;
            float sin_theta_max_sq = sin_theta_max * sin_theta_max;
            float inv_sin_theta_max = 1 / sin_theta_max;
            float cos_theta_max = sqrtf(std_max(0.f, 1 - sin_theta_max_sq));
            float cos_theta = (cos_theta_max - 1) * ux + 1;
            float sin_theta_sq = 1 - cos_theta * cos_theta;
            if (sin_theta_max_sq < 0.00068523f) {
                sin_theta_sq = sin_theta_max_sq * ux;
                cos_theta = sqrtf(1 - sin_theta_sq);
            }
            float cos_alpha = sin_theta_sq * inv_sin_theta_max +
                              cos_theta * sqrtf(std_max(0.f, 1.f - sin_theta_sq * inv_sin_theta_max * inv_sin_theta_max));
            float sin_alpha = sqrtf(std_max(0.f, 1.f - cos_alpha * cos_alpha));
            float phi = uy * 2 * JPB_PI;
            Frame fr = make_frame((p0 - P) * inv_dist);
            f3 wn = (sin_alpha * jp_cosf(phi)) * (-fr.s) + (sin_alpha * jp_sinf(phi)) * (-fr.t) + cos_alpha * (-fr.n);
            lpos = p0 + radius * wn;
            lnrm = wn;
            pdf = 1 / (2 * JPB_PI * (1 - cos_theta_max));
        }
    } else {
        // SamplePosition, then FShape::SampleDirection's area -> solid-angle conversion (shape.h:124-145)
        if (shape_type == SHAPE_TRI) {
            const f3 p1 = mk3(l2), p2 = mk3(ldg4(L + 3));
            float su0 = sqrtf(ux);
            float bx = 1 - su0, by = uy * su0;
            lpos = bx * p0 + by * p1 + (1 - bx - by) * p2;
            lnrm = mk3(ldg4(L + 4));
        } else if (shape_type == SHAPE_RECT) {
            const f3 p1 = mk3(l2), p2 = mk3(ldg4(L + 3));
            lpos = p1 + (p0 - p1) * ux + (p2 - p1) * uy;
            lnrm = mk3(ldg4(L + 4));
        } else {  // disk
            lnrm = mk3(ldg4(L + 4));
            Frame fr = make_frame(lnrm);
            float sx, sy;
            concentric_disk_sample(ux, uy, &sx, &sy);
            lpos = p0 + l2.w * (fr.s * sx + fr.t * sy);
        }
        pdf = inv_area;
        const f3 w = lpos - P;
        dist2 = length2(w);
        if (dist2 == 0) pdf = 0;
        else {
            dist = sqrtf(dist2);
            wn = w / dist;  // normalize(w)
            pdf *= dist2 / absdot(lnrm, -wn);
            if (isinf(pdf)) pdf = 0;
        }
    }
    // FAreaLight::Sample_Li, light.h:199-216
    s.pdf = pdf;
    s.pos = lpos;
    if (dist2 < 0.f) {  // (cone sampling of a sphere light: not normalised above)
        const f3 w = lpos - P;
        dist2 = length2(w);
        if (dist2 != 0) {
            dist = sqrtf(dist2);
            wn = w / dist;
        }
    }
    if (pdf == 0 || dist2 == 0) {
        s.Li = mk3(0, 0, 0);
    } else {
        s.wi = wn;
        s.dist = dist;
        s.Li = area_L(color, lnrm, -s.wi);
    }
    return s;
}

// FIntersection::Le -> FPrimitive::GetLe -> FAreaLight::L, shape.cc:17-20, primitive.h:60-63
__device__ __forceinline__ f3 emitted(const DevScene& sc, int light, const f3& N, const f3& wo) {
    if (light < 0) return mk3(0, 0, 0);
    const f3 radiance = mk3(ldg4(sc.lights + (size_t)light * kLightStride));
    return area_L(radiance, N, wo);
}

}  // namespace jpbrt
