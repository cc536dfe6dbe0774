// intersect.cuh -- primitive tests and BVH traversal: the device restatement of
// FShape::Intersect (shape.h:199-221,291-327,399-435,487-526), FPrimitive::Intersect
// (primitive.h:39-48) and FScene::Intersect/Occluded (scene.cc:25-33, scene.h:36-47).
//
// Primitive tests keep the reference's exact float expressions (compiled with -fmad=false): the
// set of accepted (ray, primitive) pairs and every accepted t are bit-identical to the CPU code.
// The BVH only prunes: boxes are padded and tested conservatively (reciprocal-multiply slabs
// widened by 2*gamma(3); or, for mid-size trees, 32-byte nodes quantised to a 16-bit grid and rounded
// outwards), traversal is front-to-back with early-out, and an any-hit variant serves shadow rays --
// none of which can change a hit, only the tie-break between primitives that report EXACTLY equal t
// (SURVEY.md Appendix A.6).
#pragma once

#include "dev_scene.h"
#include "dmath.cuh"

namespace jpbrt {

enum { SHAPE_TRI = 0, SHAPE_RECT = 1, SHAPE_SPHERE = 2, SHAPE_DISK = 3 };

__device__ __forceinline__ float4 ldg4(const Float4* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 256-bit read-only load (sm_100: LDG.E.256): two float4 of a 32-byte-aligned record in ONE load instruction.
// A 64-byte BVH node is then 2 loads instead of 4, a triangle slot 2 instead of 3.
#ifndef JPB_LDG256
#define JPB_LDG256 1
#endif
// JPB_LDG256: 0 off, 1 nodes and slots, 2 nodes only, 3 slots only (A/B builds)
template <bool WIDE>
__device__ __forceinline__ void ldg8(const Float4* p, float4& a, float4& b) {
    if (WIDE) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
    } else {
        a = ldg4(p);
        b = ldg4(p + 1);
    }
}
constexpr bool kWideNodeLoads = JPB_LDG256 == 1 || JPB_LDG256 == 2;
constexpr bool kWideSlotLoads = JPB_LDG256 == 1 || JPB_LDG256 == 3;

// One primitive slot against one ray.  Returns true and shrinks tmax on an accepted hit.
// `nrm_lookup` is the slot's (normal, tag) record, fetched only after the edge tests pass.
__device__ __forceinline__ bool intersect_slot(const Float4* __restrict__ slot, const Float4* __restrict__ nrm_rec,
                                               const f3& o, const f3& d, float tmin, float& tmax) {
    float4 q0, q1, q2, q3;
    ldg8<kWideSlotLoads>(slot, q0, q1);
    const int type = __float_as_int(q0.w) & ((1 << kTypeBits) - 1);
    const f3 p0 = mk3(q0);
    if (type == SHAPE_TRI) {  // shape.h:291-327
        ldg8<kWideSlotLoads>(slot + 2, q2, q3);
        const f3 p1 = mk3(q1), p2 = mk3(q2);
        const f3 oa = p0 - o, ob = p1 - o, oc = p2 - o;
        const f3 v0 = cross(oc, ob), v1 = cross(ob, oa), v2 = cross(oa, oc);
        const float v0d = dot(v0, d), v1d = dot(v1, d), v2d = dot(v2, d);
        if (((v0d < 0) && (v1d < 0) && (v2d < 0)) || ((v0d >= 0) && (v1d >= 0) && (v2d >= 0))) {
            const f3 n = mk3(q3);  // the stored normal: a copy rides in the slot's fourth quarter (scene_flatten.cc MakeSlot)
            const float dist = dot(n, oa) / dot(n, d);
            if ((dist > tmin) && (dist < tmax)) { tmax = dist; return true; }
        }
        return false;
    }
    if (type == SHAPE_RECT) {  // shape.h:399-435
        ldg8<kWideSlotLoads>(slot + 2, q2, q3);
        const f3 p1 = mk3(q1), p2 = mk3(q2), p3 = mk3(q3);
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
        const float radius = q1.x;
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

// FFrame(isect.normal) (geometry.h:344-377).  For flat shapes it was computed once at upload with the same
// expressions (slot_frame); a rectangle whose normal was flipped towards the ray gets (s, -t, -n), which is
// exactly FFrame(-n): cross(-n, tmp) = -t and cross(-t, -n) = s, and negation is exact.
__device__ __forceinline__ Frame hit_frame(const DevScene& sc, int slot, const f3& N) {
    const float4 nr = ldg4(sc.slot_nrm + slot);
    const int type = __float_as_int(nr.w) & ((1 << kTypeBits) - 1);
    if (type == SHAPE_SPHERE) return make_frame(N);
    const Float4* fp = sc.slot_frame + (size_t)slot * kFrameStride;
    Frame f;
    f.s = mk3(ldg4(fp));
    f.t = mk3(ldg4(fp + 1));
    f.n = mk3(ldg4(fp + 2));
    const bool flipped = (type == SHAPE_RECT) && (N.x != nr.x || N.y != nr.y || N.z != nr.z);
    if (flipped) { f.t = -f.t; f.n = -f.n; }
    return f;
}

constexpr int kTraversalStack = 64;     // deepest node stack a ray can use (entries, the sentinel included)
constexpr int kTravDone = 0x7fffffff;  // `cur` value of a lane whose stack is empty (or that found an any-hit)
constexpr int kTravBlock = 256;        // threads per block of every kernel that traverses (wavefront.cuh: kBlock)

// Where the node stack lives: LOCAL memory, interleaved per lane by the hardware (a converged push is one L1 wavefront).
// Round 2 built the obvious alternative -- the first 8 / 12 / 16 entries of every lane's stack in SHARED memory, [entry][thread]
// (conflict-free at any mix of depths), deeper entries spilling to local -- and measured it 4-5 % SLOWER on every scene
// (profiles/ab/r02_ab_stack.log: bunny k_extend 13.09 vs 12.52 ms, k_connect 12.25 vs 11.84): two more address instructions
// per access and 48-96 KB per SM taken from L1.  (The variant lived behind -DJPB_SMEM_STACK until the pointer stack below.)
//
// GUARD (template parameter of the walk): test every push against the end of the stack and count what does not fit.  The
// uploader bounds the depth of every tree it installs (c_api.cu: kMaxBvhDepth < kTraversalStack), and a walk's stack never
// holds more entries than the tree has levels, so the production kernels run unguarded; the COUNT variants, the 5-block
// variants and any scene installed with a deeper tree (test hook JPBRT_TEST_ALLOW_DEEP_BVH) keep the guard and its counter.

// QUANTISED NODES (template parameter QN of the walk): a 32-byte node -- twelve 16-bit plane indices on a grid over the (padded) scene bounds + the two
// child references -- fetched with ONE 256-bit load instead of two.  The planes are rebuilt without a conversion instruction:
// PRMT puts the 16 bits under the exponent of 2^23 (float 8388608 + q), and the ray's slab constants absorb the offset:
// t = (8388608 + q) * (cell * inv) + ((origin - o) * inv - 8388608 * cell * inv).  The cancellation costs up to one cell of
// plane position, so the uploader rounds every box outwards by 2 more cells (scene_flatten.cc); boxes only prune, hits are
// the primitive tests', so the images cannot change -- only how many boxes a ray enters.
// Measured on B200 (profiles/ab/r02_ab_qnodes.log): bunny scene k_connect -5.9 %, k_extend +-0; 5 M triangles k_extend -4 %,
// k_connect +0.4 %; Cornell / glossy (trees of 15 / 33 nodes that live in L1 whatever their size) +5-9 % SLOWER: the decode's 12
// PRMT are pure cost there.  QN is therefore a template parameter of the walk and the library picks per scene (c_api.cu).

// Per-lane traversal state.  The BVH walk is a state machine so that a warp can (a) run the inner-node
// step and the leaf step in separate, converged phases (while-while traversal) and (b) hand a finished
// lane a new ray while the other lanes keep walking (ray replacement) -- see traverse_queue() below.
#ifndef JPB_QN_FAST_RCP
#define JPB_QN_FAST_RCP 1
#endif
#ifndef JPB_QN_SELECT
#define JPB_QN_SELECT 1  // quantised walk: near / far planes picked by a per-ray PRMT selector (0: both distances + min / max, A/B)
#endif
struct Trav {
    f3 o, d, inv, oi;
    float tmin, tmax;
    int cur, hit;
    int* top;  // one past the newest entry of this lane's node stack
#if JPB_QN_SELECT
    unsigned selx, sely, selz;  // quantised walk: PRMT selector of the NEAR plane of each axis (low half = min if the ray runs up the axis)
#endif
};  // the node stack is separate (shared + local arrays) so that these scalars stay in registers

// This lane's node stack.  Entry 0 is a sentinel (kTravDone) that is never overwritten: popping an empty stack yields
// "done" without a compare.
// The stack position is a POINTER into the lane's local array (round 2: with an index the compiler spent an LEA per push and
// per pop and branched around the push; with the pointer the push is one predicated STL + IADD: 54 instead of 62 instructions
// per node step -- for 0.3 % of the kernels' time, which is how the L1 pipe, not the issue rate, was found to be their bound).
// (Caching the newest entry in a register, so that a push followed by a pop touches no memory: measured 3-6 % SLOWER -- the
// extra register costs spills at 40 registers; profiles/ab/r02_ab_normal_in_slot_tos.log.)
struct TravStack {
    int* lm;  // the lane's stack, kTraversalStack entries of local memory
    unsigned long long* dropped;  // stats counter of pushes that found the stack full (never in the five configs)
    __device__ __forceinline__ void init() const { lm[0] = kTravDone; }
    template <bool GUARD>
    __device__ __forceinline__ void push(int*& top, int v) const {
        if (!GUARD || top < lm + kTraversalStack) *top++ = v;
        else if (dropped) atomicAdd(dropped, 1ull);  // the far child is lost: counted, reported as invalid_contributions / stack_overflows
    }
    __device__ __forceinline__ int pop(int*& top) const { return *--top; }
};

// One axis of the quantised walk's slab constants.  plane(q) = origin + q * cell;  t(q) = (8388608 + q) * cinv + oiq  with
// cinv = cell * inv and oiq = (origin - o) * inv - 8388608 * cinv.  If 8388608 * cinv leaves the float range (a direction
// component of ~1e-34, or a scene spanning 30 orders of magnitude) both constants become NaN: FMNMX ignores NaN operands, so the
// axis drops out of the slab test -- conservative -- instead of producing planes at -inf that would reject every box.
__device__ __forceinline__ void quant_axis(float origin, float cell, float o, float inv, float& cinv, float& oiq) {
    cinv = cell * inv;
    oiq = __fmaf_rn(origin - o, inv, -8388608.f * cinv);
    if (!(fabsf(oiq) <= 3.4028234664e38f)) cinv = oiq = __int_as_float(0x7fc00000);
}

template <bool QN = false>
__device__ __forceinline__ void trav_init(const DevScene& sc, Trav& t, const f3& o, const f3& d, float tmin, float tmax) {
    t.o = o;
    t.d = d;
    if (QN) {
        // (the slab constants of the quantised walk feed box tests with two cells of slack: the hardware's approximate reciprocal
        //  -- one MUFU instead of the ~12-instruction IEEE division with its slow-path branch, three times per ray -- is exact
        //  enough by four orders of magnitude; a flushed denormal or zero component gives +-inf and the axis drops out below.
        //  Measured: k_extend -0.9 %, k_connect -1.0 % on the bunny scene, film unchanged; profiles/ab/r02_ab_qn_fast_rcp_farsel.log)
#if JPB_QN_FAST_RCP
        f3 inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(d.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(d.y));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.z) : "f"(d.z));
#else
        const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
#endif
        quant_axis(sc.q_origin[0], sc.q_cell[0], o.x, inv.x, t.inv.x, t.oi.x);
        quant_axis(sc.q_origin[1], sc.q_cell[1], o.y, inv.y, t.inv.y, t.oi.y);
        quant_axis(sc.q_origin[2], sc.q_cell[2], o.z, inv.z, t.inv.z, t.oi.z);
#if JPB_QN_SELECT
        // the plane a ray meets first on an axis is the box's min if it runs up the axis, its max if it runs down
        t.selx = d.x < 0.f ? 0x7632u : 0x7610u;
        t.sely = d.y < 0.f ? 0x7632u : 0x7610u;
        t.selz = d.z < 0.f ? 0x7632u : 0x7610u;
#endif
    } else {
        const f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        t.inv = inv;
        // slab distances as one explicit FMA each: bound * inv - o * inv.  Its rounding error (about one ulp of
        // |o * inv|) is covered by the padding baked into every stored box (4e-6 * the largest ray-origin
        // coordinate, scene_flatten.cc) plus the 2*gamma(3) widening of tfar.
        t.oi = mk3(-(o.x * t.inv.x), -(o.y * t.inv.y), -(o.z * t.inv.z));
    }
    t.tmin = tmin;
    t.tmax = tmax;
    t.cur = 0;  // root (always an inner node, scene_flatten.cc)
    t.hit = -1;  // (t.top: set by the caller, which owns the stack)
}

__device__ __forceinline__ bool trav_at_inner(const Trav& t) { return (unsigned)t.cur < (unsigned)kTravDone; }
__device__ __forceinline__ bool trav_at_leaf(const Trav& t) { return t.cur < 0; }
__device__ __forceinline__ bool trav_done(const Trav& t) { return t.cur == kTravDone; }

// One inner node: fetch 64 bytes, test both child boxes, descend into the nearer hit child.  Written without
// divergent branches: the push (both children hit) and the pop (none hit) are short predicated sequences.
// Counters of the COUNT kernel variants (bench.py's roofline): per-lane tests as SURVEY.md 8(d) accounts them, and
// DISTINCT records per warp step -- lanes of a warp that sit on the same node (or primitive) share one fetch, so the
// distinct counts are what the memory system actually has to deliver.
struct TravCounts {
    unsigned box = 0, prim = 0;            // box tests (2 per node step) and primitive tests, per lane
    unsigned node_fetch = 0, prim_fetch = 0;  // distinct nodes / primitive records fetched per warp step (counted on one lane)
};

// Occlusion queries (k_connect) descend LEFT-first instead of near-first: FScene::Occluded only asks whether anything is hit
// (scene.h:36-47), so any order gives the same boolean, and skipping the distance compare + child ordering measured
// k_connect -5.8 % on the bunny scene, -5.6 % on 5 M triangles, -1.7 % Cornell, -1.1 % glossy (profiles/ab/r02_ab_anyhit.log).
#ifndef JPB_ANYHIT_UNORDERED
#define JPB_ANYHIT_UNORDERED 1
#endif
template <bool COUNT, bool ANY_HIT = false, bool GUARD = true, bool QN = false>
__device__ __forceinline__ void trav_node_step(const DevScene& sc, Trav& t, const TravStack& stk, TravCounts& cnt, unsigned step_mask) {
    const float widen = 1.0000004f;  // 1 + 2*gamma(3): pbrt's conservative slab bound
    int cl, cr;
    float ltn, ltf, rtn, rtf;
    bool hl, hr;
    if (QN) {
        float4 q0, q1;
        ldg8<true>(sc.qnodes + (size_t)t.cur * kQNodeStride, q0, q1);
        const unsigned w0 = __float_as_uint(q0.x), w1 = __float_as_uint(q0.y), w2 = __float_as_uint(q0.z);  // left box: x, y, z (min | max << 16)
        const unsigned w3 = __float_as_uint(q0.w), w4 = __float_as_uint(q1.x), w5 = __float_as_uint(q1.y);  // right box
        const unsigned two23 = 0x4B000000u;  // 8388608.f: a 16-bit q in its low mantissa bits reads 8388608 + q
        cl = __float_as_int(q1.z);
        cr = __float_as_int(q1.w);
        if (COUNT) {
            cnt.box += 2;
            const unsigned same = __match_any_sync(step_mask, t.cur);
            if ((threadIdx.x & 31) == __ffs(same) - 1) cnt.node_fetch += 1;
        }
#if JPB_QN_SELECT
        // near and far plane of every axis straight out of the node's words: 6 PRMT + 6 FFMA + 4 FMNMX per box
        // (prmt.b32 spelled in PTX: __byte_perm masks its selector with 0x7777 first -- one LOP3 per axis and step)
        auto plane = [](unsigned w, unsigned sel) {
            unsigned r;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(sel));
            return __uint_as_float(r);
        };
        // (the far planes' selectors: derived here -- keeping them in three more registers measured +-0)
        const unsigned fx = t.selx ^ 0x22u, fy = t.sely ^ 0x22u, fz = t.selz ^ 0x22u;
        ltn = fmaxf(fmaxf(__fmaf_rn(plane(w0, t.selx), t.inv.x, t.oi.x),
                          __fmaf_rn(plane(w1, t.sely), t.inv.y, t.oi.y)),
                    fmaxf(__fmaf_rn(plane(w2, t.selz), t.inv.z, t.oi.z), t.tmin));
        ltf = fminf(fminf(__fmaf_rn(plane(w0, fx), t.inv.x, t.oi.x),
                          __fmaf_rn(plane(w1, fy), t.inv.y, t.oi.y)),
                    fminf(__fmaf_rn(plane(w2, fz), t.inv.z, t.oi.z), t.tmax));
        rtn = fmaxf(fmaxf(__fmaf_rn(plane(w3, t.selx), t.inv.x, t.oi.x),
                          __fmaf_rn(plane(w4, t.sely), t.inv.y, t.oi.y)),
                    fmaxf(__fmaf_rn(plane(w5, t.selz), t.inv.z, t.oi.z), t.tmin));
        rtf = fminf(fminf(__fmaf_rn(plane(w3, fx), t.inv.x, t.oi.x),
                          __fmaf_rn(plane(w4, fy), t.inv.y, t.oi.y)),
                    fminf(__fmaf_rn(plane(w5, fz), t.inv.z, t.oi.z), t.tmax));
        // (no widening of tfar: the quantised boxes are rounded outwards by 2 cells more than the arithmetic can lose)
        hl = ltn <= ltf;
        hr = rtn <= rtf;
#else
        const float lx0 = __uint_as_float(__byte_perm(w0, two23, 0x7610)), lx1 = __uint_as_float(__byte_perm(w0, two23, 0x7632));
        const float ly0 = __uint_as_float(__byte_perm(w1, two23, 0x7610)), ly1 = __uint_as_float(__byte_perm(w1, two23, 0x7632));
        const float lz0 = __uint_as_float(__byte_perm(w2, two23, 0x7610)), lz1 = __uint_as_float(__byte_perm(w2, two23, 0x7632));
        const float rx0 = __uint_as_float(__byte_perm(w3, two23, 0x7610)), rx1 = __uint_as_float(__byte_perm(w3, two23, 0x7632));
        const float ry0 = __uint_as_float(__byte_perm(w4, two23, 0x7610)), ry1 = __uint_as_float(__byte_perm(w4, two23, 0x7632));
        const float rz0 = __uint_as_float(__byte_perm(w5, two23, 0x7610)), rz1 = __uint_as_float(__byte_perm(w5, two23, 0x7632));
        float a0 = __fmaf_rn(lx0, t.inv.x, t.oi.x), a1 = __fmaf_rn(lx1, t.inv.x, t.oi.x);
        float b0 = __fmaf_rn(ly0, t.inv.y, t.oi.y), b1 = __fmaf_rn(ly1, t.inv.y, t.oi.y);
        float c0 = __fmaf_rn(lz0, t.inv.z, t.oi.z), c1 = __fmaf_rn(lz1, t.inv.z, t.oi.z);
        ltn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), t.tmin));
        ltf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), t.tmax));
        a0 = __fmaf_rn(rx0, t.inv.x, t.oi.x); a1 = __fmaf_rn(rx1, t.inv.x, t.oi.x);
        b0 = __fmaf_rn(ry0, t.inv.y, t.oi.y); b1 = __fmaf_rn(ry1, t.inv.y, t.oi.y);
        c0 = __fmaf_rn(rz0, t.inv.z, t.oi.z); c1 = __fmaf_rn(rz1, t.inv.z, t.oi.z);
        rtn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), t.tmin));
        rtf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), t.tmax));
        hl = ltn <= ltf * widen;
        hr = rtn <= rtf * widen;
#endif
    } else {
        const Float4* np = sc.nodes + (size_t)t.cur * kNodeStride;
        float4 n0, n1, n2, n3;
        ldg8<kWideNodeLoads>(np, n0, n1);
        ldg8<kWideNodeLoads>(np + 2, n2, n3);
        cl = __float_as_int(n3.x);
        cr = __float_as_int(n3.y);
        if (COUNT) {
            cnt.box += 2;
            const unsigned same = __match_any_sync(step_mask, t.cur);  // the lanes of this step that sit on the same node
            if ((threadIdx.x & 31) == __ffs(same) - 1) cnt.node_fetch += 1;
        }
        // left box: min = (n0.x n0.y n0.z), max = (n0.w n1.x n1.y)
        float a0 = __fmaf_rn(n0.x, t.inv.x, t.oi.x), a1 = __fmaf_rn(n0.w, t.inv.x, t.oi.x);
        float b0 = __fmaf_rn(n0.y, t.inv.y, t.oi.y), b1 = __fmaf_rn(n1.x, t.inv.y, t.oi.y);
        float c0 = __fmaf_rn(n0.z, t.inv.z, t.oi.z), c1 = __fmaf_rn(n1.y, t.inv.z, t.oi.z);
        ltn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), t.tmin));
        ltf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), t.tmax));
        // right box: min = (n1.z n1.w n2.x), max = (n2.y n2.z n2.w)
        a0 = __fmaf_rn(n1.z, t.inv.x, t.oi.x); a1 = __fmaf_rn(n2.y, t.inv.x, t.oi.x);
        b0 = __fmaf_rn(n1.w, t.inv.y, t.oi.y); b1 = __fmaf_rn(n2.z, t.inv.y, t.oi.y);
        c0 = __fmaf_rn(n2.x, t.inv.z, t.oi.z); c1 = __fmaf_rn(n2.w, t.inv.z, t.oi.z);
        rtn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), t.tmin));
        rtf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), t.tmax));
        hl = ltn <= ltf * widen;
        hr = rtn <= rtf * widen;
    }
    const bool both = hl && hr;
    // the nearer child first, left on ties (the reference's order, bvh.h:99-100): right_first = both ? ltn > rtn : hr, spelled
    // as one predicate expression (4 fewer instructions per step than the select chain the ternary compiled to)
    const bool right_first = (JPB_ANYHIT_UNORDERED && ANY_HIT) ? !hl  // occlusion only: any order finds the same boolean
                                                               : (hr & (!hl | (ltn > rtn)));
    const int near = right_first ? cr : cl;
    const int far = right_first ? cl : cr;
    if (both) stk.push<GUARD>(t.top, far);  // (prefetching the far child measured 3-6 % slower)
    t.cur = (hl || hr) ? near : stk.pop(t.top);
}

// One leaf: test its (<= 4) primitives, then pop.
template <bool ANY_HIT, bool COUNT>
__device__ __forceinline__ void trav_leaf_step(const DevScene& sc, Trav& t, const TravStack& stk, TravCounts& cnt, unsigned leaf_mask) {
    const int bits = ~t.cur;
    const int first = bits >> kLeafCountBits;
    const int n = bits & ((1 << kLeafCountBits) - 1);
    if (COUNT) {
        // the counting variant keeps the lanes of the phase together (no early return) so that it can count, per
        // primitive round, how many DISTINCT records the warp fetches
        bool found = false;
        for (int k = 0; k < kMaxLeafPrims; ++k) {
            const bool active = k < n && !(ANY_HIT && found);
            const unsigned m = __ballot_sync(leaf_mask, active);
            if (active) {
                const int s = first + k;
                cnt.prim += 1;
                const unsigned same = __match_any_sync(m, s);
                if ((threadIdx.x & 31) == __ffs(same) - 1) cnt.prim_fetch += 1;
                if (intersect_slot(sc.slots + (size_t)s * kSlotStride, sc.slot_nrm + s, t.o, t.d, t.tmin, t.tmax)) { t.hit = s; found = true; }
            }
        }
        t.cur = (ANY_HIT && found) ? kTravDone : stk.pop(t.top);
        return;
    }
    for (int k = 0; k < n; ++k) {
        const int s = first + k;
        if (intersect_slot(sc.slots + (size_t)s * kSlotStride, sc.slot_nrm + s, t.o, t.d, t.tmin, t.tmax)) {
            t.hit = s;
            if (ANY_HIT) { t.cur = kTravDone; return; }
        }
    }
    t.cur = stk.pop(t.top);
}


// The leaf phase with the primitive tests SPREAD OVER THE WARP (JPB_LEAF_SHARE, round 2).  In the per-lane form above a lane
// tests its own leaf's 1-4 primitives one after the other: ncu shows ~11 lanes in the first test and ~8 in the later ones,
// ~2.1 rounds per phase.  Here every (ray, primitive) pair of the phase is an ITEM; the items are laid out in a per-warp
// shared-memory list (offsets from three ballots: a leaf holds <= 4 primitives) and lane j takes item j -- fetching the
// owner's ray by shuffle -- so that ~24 items run as ONE round on ~24 lanes.  Every pair is tested with the reference's
// arithmetic against the owner's tmax at the start of the phase, the results go back through the same list, and the
// owner folds them in primitive order with the reference's strict `t < tmax`: exactly what the sequential loop accepts
// (a later primitive wins only if it is strictly closer), hence the same hit record, ties included.
// Called by ALL lanes of the warp, converged.  items: this warp's 32 * kMaxLeafPrims words of shared memory.
// MEASURED AND REJECTED (profiles/ab/r02_ab_leafshare.log): bit-identical films, but k_extend +17 % (bunny) / +32 % (Cornell),
// k_connect +21 % / +40 %: eight shuffles, two __syncwarp and the list traffic per phase cost more than the ~1.1 primitive
// rounds they save -- the sequential per-lane loop stays.  Kept buildable (-DJPB_LEAF_SHARE=1) as the record of the experiment.
#ifndef JPB_LEAF_SHARE
#define JPB_LEAF_SHARE 0
#endif
template <bool ANY_HIT>
__device__ __forceinline__ void trav_leaf_phase_shared(const DevScene& sc, Trav& t, const TravStack& stk, unsigned* items) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool at_leaf = trav_at_leaf(t);
    int first = 0, n = 0;
    if (at_leaf) {
        const int bits = ~t.cur;
        first = bits >> kLeafCountBits;
        n = bits & ((1 << kLeafCountBits) - 1);
    }
    // exclusive prefix sum of n over the lanes, n in 0..4, from the ballots of its three bits
    const unsigned b0 = __ballot_sync(full, n & 1), b1 = __ballot_sync(full, n & 2), b2 = __ballot_sync(full, n & 4);
    const unsigned below = (1u << lane) - 1;
    const int off = __popc(b0 & below) + 2 * __popc(b1 & below) + 4 * __popc(b2 & below);
    const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    for (int k = 0; k < n; ++k) items[off + k] = ((unsigned)lane << 27) | (unsigned)(first + k);
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        const bool mine = j < total;
        const unsigned item = mine ? items[j] : ((unsigned)lane << 27);
        const int owner = (int)(item >> 27), slot = (int)(item & 0x7ffffffu);
        f3 o, d;
        o.x = __shfl_sync(full, t.o.x, owner); o.y = __shfl_sync(full, t.o.y, owner); o.z = __shfl_sync(full, t.o.z, owner);
        d.x = __shfl_sync(full, t.d.x, owner); d.y = __shfl_sync(full, t.d.y, owner); d.z = __shfl_sync(full, t.d.z, owner);
        const float tmin = __shfl_sync(full, t.tmin, owner);
        float tmax = __shfl_sync(full, t.tmax, owner);
        if (mine) {
            const bool hit = intersect_slot(sc.slots + (size_t)slot * kSlotStride, sc.slot_nrm + slot, o, d, tmin, tmax);
            items[j] = hit ? __float_as_uint(tmax) : 0x7f800000u;  // the accepted t, or +inf
        }
    }
    __syncwarp();
    if (at_leaf) {
        bool found = false;
        for (int k = 0; k < n; ++k) {  // in primitive order, strict <: the sequential loop's acceptance (shape.h:318)
            const float r = __uint_as_float(items[off + k]);
            if (r < t.tmax) { t.tmax = r; t.hit = first + k; found = true; }
        }
        t.cur = (ANY_HIT && found) ? kTravDone : stk.pop(t.top);
    }
    __syncwarp();  // the list is rewritten by the next phase
}

// Warp-cooperative traversal of a whole ray queue.
//   io.load(i, o, d, tmin, tmax)    fetch ray i (false: slot i holds no ray)   io.store(i, slot, t)   deliver its result
// Every lane owns one ray at a time.  Each trip round the outer loop has three converged phases:
//   refill : idle lanes take the next rays from the queue (one atomicAdd per warp) once at least
//            `refill_min` lanes are idle -- finished rays are REPLACED instead of idling until the
//            slowest ray of the warp is done;
//   nodes  : lanes at an inner node step until (almost, see min_inner) every active lane sits on a leaf or is done;
//   leaves : lanes at a leaf test its primitives and pop; finished lanes store their result.
// Invariant: a lane without a ray (idx < 0) has t.cur == kTravDone, so "at an inner node" / "at a leaf" need no idx test.
// Must be called by every thread of a kTravBlock-thread block (static shared memory).
// Node steps per warp vote of the node phase.  The votes that decide whether the phase goes on (any lane at an inner node?
// fewer than min_inner? any lane at a leaf?) are ~15 of the ~85 instructions of a step; taking two steps per vote -- a lane
// that reaches a leaf on the first sits out the second -- measured +1.5-2 % on all scenes (bunny 1,543 -> 1,568 Msamples/s at
// 48 spp, Cornell 1,076 -> 1,090, glossy 186 -> 190), three steps per vote gives the gain back (profiles/ab/r02_ab_unroll.log).
// SPECULATIVE WALK (Aila & Laine's postponed leaf: a lane that reaches a leaf stashes it and keeps stepping through inner nodes
// until its next leaf, so that more lanes are busy in both phases) -- built in round 2, measured, removed: k_extend +4 %
// (bunny), +6 % (Cornell), +8 % (5 M triangles) SLOWER, k_connect +2-5 %, glossy -1.6 % (profiles/ab/r02_ab_speculative_lean.log):
// the nodes walked before the postponed leaf's hit shortens the ray are wasted, and the extra state costs spills at 40 registers.
// EARLY REFILL (cutting a node phase short to deliver results and take new rays once 8 / 12 / 16 / 20 / 24 lanes have finished
// inside it -- most rays end in a node phase): measured and removed, k_extend +1-5 %, k_connect +1-4 % slower at every threshold
// (profiles/ab/r02_ab_early_refill.log): a refill costs the warp more issue slots than the idle lanes it fills give back.
// CLAIM-AHEAD (ray indices handed out from 32-ray chunks claimed one chunk ahead, so that no warp waits for the round trip of the
// work cursor's atomicAdd -- ncu attributes 5-6 % (bunny) to 19 % (Cornell k_connect) of the stall samples to the shuffle behind
// it): measured and removed, k_extend +1-3 %, k_connect +0.7-2.5 % SLOWER (profiles/ab/r02_ab_claim_ahead.log).  The kernels are
// bound by issue slots and the L1 pipe, not by latency: while one warp waits for its claim the SM's other 47 issue, and the
// chunk bookkeeping's instructions and spills cost more than the wait.  (The same holds for the ray records' DRAM latency: see
// ConnectIO in wavefront.cuh.)
#ifndef JPB_NODE_UNROLL
#define JPB_NODE_UNROLL 2
#endif
template <bool ANY_HIT, bool COUNT, bool GUARD = true, bool QN = false, typename IO>
__device__ __forceinline__ void traverse_queue(const DevScene& sc, int n, int* work, const IO& io, int refill_min, int min_inner,
                                               TravCounts& cnt, unsigned long long* dropped = nullptr) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    __shared__ unsigned s_items[JPB_LEAF_SHARE ? kTravBlock * kMaxLeafPrims : 1];  // per warp: 32 * kMaxLeafPrims leaf-phase items
    const bool share_leaves = JPB_LEAF_SHARE && !COUNT && sc.n_slots < (1 << 27) && sc.max_leaf_prims <= kMaxLeafPrims;
    int l_stack[kTraversalStack];
    const TravStack stk{l_stack, dropped};
    stk.init();
    Trav t;
    t.cur = kTravDone;
    t.top = l_stack + 1;
    int idx = -1;
    bool exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(full, idx < 0);
        if (!exhausted && (idle == full || __popc(idle) >= refill_min)) {
            const int cnt = __popc(idle);
            int base = 0;
            if (lane == 0) base = atomicAdd(work, cnt);
            base = __shfl_sync(full, base, 0);
            if (idx < 0) {
                const int i = base + __popc(idle & ((1u << lane) - 1));
                if (i < n) {
                    f3 o, d;
                    float tmin, tmax;
                    if (io.load(i, o, d, tmin, tmax)) {
                        trav_init<QN>(sc, t, o, d, tmin, tmax);
                        t.top = l_stack + 1;  // above the sentinel
                        idx = i;
                    }
                }
            }
            if (base + cnt >= n) exhausted = true;
        }
        if (__ballot_sync(full, idx >= 0) == 0) {
            if (exhausted) break;
            continue;
        }
        // Node phase.  Waiting for EVERY lane to reach a leaf costs ~17 node steps per phase at ~11 active lanes (ncu source
        // counters, bunny scene): the slowest ray of 32 sets the pace.  The phase therefore ends as soon as fewer than
        // `min_inner` lanes are still walking inner nodes and some lane has a leaf to test; the stragglers resume after the
        // (short) leaf phase.  Measured on B200: min_inner 8 = +12 % on the bunny scene, +20 % on the 5 M-triangle scene.
        for (;;) {
            const unsigned m_inner = __ballot_sync(full, trav_at_inner(t));
            if (m_inner == 0) break;
            if (__popc(m_inner) < min_inner && __any_sync(full, trav_at_leaf(t))) break;
#pragma unroll
            for (int u = 0; u < JPB_NODE_UNROLL; ++u) {
                const unsigned step_mask = COUNT ? (u == 0 ? m_inner : __ballot_sync(full, trav_at_inner(t))) : 0u;
                if (trav_at_inner(t)) trav_node_step<COUNT, ANY_HIT, GUARD, QN>(sc, t, stk, cnt, step_mask);
            }
        }
        const unsigned leaf_mask = COUNT ? __ballot_sync(full, trav_at_leaf(t)) : 0u;
        if (share_leaves) {
            if (__any_sync(full, trav_at_leaf(t))) trav_leaf_phase_shared<ANY_HIT>(sc, t, stk, s_items + (threadIdx.x >> 5) * (32 * kMaxLeafPrims));
        } else if (trav_at_leaf(t)) {
            trav_leaf_step<ANY_HIT, COUNT>(sc, t, stk, cnt, leaf_mask);
        }
        if (idx >= 0 && trav_done(t)) {
            io.store(idx, t.hit, t.tmax);
            idx = -1;
        }
    }
}

}  // namespace jpbrt
