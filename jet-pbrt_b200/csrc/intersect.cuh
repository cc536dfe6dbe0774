// intersect.cuh -- primitive tests and BVH traversal: the device restatement of
// FShape::Intersect (shape.h:199-221,291-327,399-435,487-526), FPrimitive::Intersect
// (primitive.h:39-48) and FScene::Intersect/Occluded (scene.cc:25-33, scene.h:36-47).
//
// Primitive tests keep the reference's exact float expressions (compiled with -fmad=false): the
// set of accepted (ray, primitive) pairs and every accepted t are bit-identical to the CPU code.
// The BVH only prunes: boxes are padded and tested conservatively (reciprocal-multiply slabs
// widened by 2*gamma(3)), traversal is front-to-back with early-out, and an any-hit variant
// serves shadow rays -- none of which can change a hit, only the tie-break between primitives
// that report EXACTLY equal t.
#pragma once

#include "dev_scene.h"
#include "dmath.cuh"

namespace jpbrt {

enum { SHAPE_TRI = 0, SHAPE_RECT = 1, SHAPE_SPHERE = 2, SHAPE_DISK = 3 };

__device__ __forceinline__ float4 ldg4(const Float4* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// One primitive slot against one ray.  Returns true and shrinks tmax on an accepted hit.
// `nrm_lookup` is the slot's (normal, tag) record, fetched only after the edge tests pass.
__device__ __forceinline__ bool intersect_slot(const Float4* __restrict__ slot, const Float4* __restrict__ nrm_rec,
                                               const f3& o, const f3& d, float tmin, float& tmax) {
    const float4 q0 = ldg4(slot);
    const int type = __float_as_int(q0.w) & ((1 << kTypeBits) - 1);
    const f3 p0 = mk3(q0);
    if (type == SHAPE_TRI) {  // shape.h:291-327
        const f3 p1 = mk3(ldg4(slot + 1)), p2 = mk3(ldg4(slot + 2));
        const f3 oa = p0 - o, ob = p1 - o, oc = p2 - o;
        const f3 v0 = cross(oc, ob), v1 = cross(ob, oa), v2 = cross(oa, oc);
        const float v0d = dot(v0, d), v1d = dot(v1, d), v2d = dot(v2, d);
        if (((v0d < 0) && (v1d < 0) && (v2d < 0)) || ((v0d >= 0) && (v1d >= 0) && (v2d >= 0))) {
            const f3 n = mk3(ldg4(nrm_rec));
            const float dist = dot(n, oa) / dot(n, d);
            if ((dist > tmin) && (dist < tmax)) { tmax = dist; return true; }
        }
        return false;
    }
    if (type == SHAPE_RECT) {  // shape.h:399-435
        const f3 p1 = mk3(ldg4(slot + 1)), p2 = mk3(ldg4(slot + 2)), p3 = mk3(ldg4(slot + 3));
        const f3 oa = p0 - o, ob = p1 - o, oc = p2 - o, od = p3 - o;
        const f3 v0 = cross(oc, ob), v1 = cross(ob, oa), v2 = cross(oa, od), v3 = cross(od, oc);
        const float v0d = dot(v0, d), v1d = dot(v1, d), v2d = dot(v2, d), v3d = dot(v3, d);
        if (((v0d < 0) && (v1d < 0) && (v2d < 0) && (v3d < 0)) || ((v0d >= 0) && (v1d >= 0) && (v2d >= 0) && (v3d >= 0))) {
            const f3 n = mk3(ldg4(nrm_rec));
            const float dist = dot(n, oa) / dot(n, d);
            if ((dist > tmin) && (dist < tmax)) { tmax = dist; return true; }
        }
        return false;
    }
    if (type == SHAPE_SPHERE) {  // shape.h:487-526
        const float radius = ldg4(slot + 1).x;
        const f3 oc = o - p0;
        const float a = length2(d);
        const float half_b = dot(oc, d);
        const float c = length2(oc) - radius * radius;
        const float disc = half_b * half_b - a * c;
        if (disc > 0.0f) {
            const float root = sqrtf(disc);
            float time;
            const float root1 = (-half_b - root) / a;
            if (root1 < tmax && root1 > tmin) time = root1;
            else {
                const float root2 = (-half_b + root) / a;
                if (root2 < tmax && root2 > tmin) time = root2;
                else return false;
            }
            tmax = time;
            return true;
        }
        return false;
    }
    {  // SHAPE_DISK, shape.h:199-221; isEqual(x, 0) is |x| <= eps * max(1, |x|)  (pbrt.h:97-104)
        const float4 q1 = ldg4(slot + 1);
        const f3 n = mk3(q1);
        const float radius = q1.w;
        const float dn = dot(d, n);
        if (fabsf(dn) <= 1.1920928955078125e-07f * std_max(1.0f, std_max(fabsf(dn), 0.0f))) return false;
        const f3 op = p0 - o;
        const float dist = dot(n, op) / dot(n, d);
        if ((dist > tmin) && (dist < tmax)) {
            const f3 hit = o + dist * d;
            if (length(p0 - hit) <= radius) { tmax = dist; return true; }
        }
        return false;
    }
}

// Hit-record normal as FIntersection::normal would hold it (shape.h:320,427,517,214).
__device__ __forceinline__ f3 hit_normal(const DevScene& sc, int slot, const f3& pos, const f3& d) {
    const float4 nr = ldg4(sc.slot_nrm + slot);
    const int type = __float_as_int(nr.w) & ((1 << kTypeBits) - 1);
    f3 n = mk3(nr);
    if (type == SHAPE_RECT) return (dot(n, d) <= 0) ? n : -n;
    if (type == SHAPE_SPHERE) return normalize(pos - mk3(ldg4(sc.slots + (size_t)slot * kSlotStride)));
    return n;
}

constexpr int kTraversalStack = 64;

// Closest-hit (ANY_HIT = false) or any-hit (ANY_HIT = true) traversal.
// Returns the hit slot (or -1); tmax is shrunk to the hit distance.
template <bool ANY_HIT, bool COUNT>
__device__ __forceinline__ int traverse(const DevScene& sc, const f3& o, const f3& d, float tmin, float& tmax,
                                        unsigned& n_box, unsigned& n_prim) {
    const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const float widen = 1.0000004f;  // 1 + 2*gamma(3): pbrt's conservative slab bound
    int stack[kTraversalStack];
    int sp = 0;
    int cur = 0;
    int hit_slot = -1;
    const Float4* __restrict__ nodes = sc.nodes;
    for (;;) {
        if (cur >= 0) {
            const Float4* np = nodes + (size_t)cur * kNodeStride;
            const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2), n3 = ldg4(np + 3);
            if (COUNT) n_box += 2;
            // left box: min = (n0.x n0.y n0.z), max = (n0.w n1.x n1.y)
            float a0 = (n0.x - o.x) * inv.x, a1 = (n0.w - o.x) * inv.x;
            float b0 = (n0.y - o.y) * inv.y, b1 = (n1.x - o.y) * inv.y;
            float c0 = (n0.z - o.z) * inv.z, c1 = (n1.y - o.z) * inv.z;
            const float ltn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), tmin));
            const float ltf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), tmax));
            // right box: min = (n1.z n1.w n2.x), max = (n2.y n2.z n2.w)
            a0 = (n1.z - o.x) * inv.x; a1 = (n2.y - o.x) * inv.x;
            b0 = (n1.w - o.y) * inv.y; b1 = (n2.z - o.y) * inv.y;
            c0 = (n2.x - o.z) * inv.z; c1 = (n2.w - o.z) * inv.z;
            const float rtn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), tmin));
            const float rtf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), tmax));
            const bool hl = ltn <= ltf * widen;
            const bool hr = rtn <= rtf * widen;
            const int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
            if (hl && hr) {
                const bool left_first = ltn <= rtn;
                cur = left_first ? cl : cr;
                if (sp < kTraversalStack) stack[sp++] = left_first ? cr : cl;
            } else if (hl) {
                cur = cl;
            } else if (hr) {
                cur = cr;
            } else {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        } else {
            const int bits = ~cur;
            const int first = bits >> kLeafCountBits;
            const int cnt = bits & ((1 << kLeafCountBits) - 1);
            for (int k = 0; k < cnt; ++k) {
                const int s = first + k;
                if (COUNT) n_prim += 1;
                if (intersect_slot(sc.slots + (size_t)s * kSlotStride, sc.slot_nrm + s, o, d, tmin, tmax)) {
                    hit_slot = s;
                    if (ANY_HIT) return hit_slot;
                }
            }
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
    return hit_slot;
}

}  // namespace jpbrt
