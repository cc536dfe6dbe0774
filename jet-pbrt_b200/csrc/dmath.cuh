// dmath.cuh -- float3 / colour arithmetic for the wavefront kernels.
//
// Every operator spells out the reference's evaluation order (geometry.h:65-159, color.h:13-69)
// and the translation unit is compiled with -fmad=false, so no multiply-add is contracted and
// results match the reference's x86-64 (no-FMA) floats bit for bit wherever only + - * / sqrt are
// involved.  Where contraction is wanted (conservative slab tests), code calls __fmaf_rn itself.
#pragma once

#include <cuda_runtime.h>

namespace jpbrt {

struct f3 {
    float x, y, z;
};

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 mk3(const float4& v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ f3 operator+(const f3& a, const f3& b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(const f3& a, const f3& b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(const f3& a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(const f3& a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, const f3& a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator/(const f3& a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
// component-wise (FColor * FColor, FColor / FColor)
__device__ __forceinline__ f3 cmul(const f3& a, const f3& b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 cdiv(const f3& a, const f3& b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ f3 csqrt(const f3& a) { return mk3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
__device__ __forceinline__ f3 splat(float v) { return mk3(v, v, v); }

__device__ __forceinline__ float dot(const f3& a, const f3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float absdot(const f3& a, const f3& b) { return fabsf(dot(a, b)); }
__device__ __forceinline__ f3 cross(const f3& a, const f3& b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length2(const f3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ float length(const f3& a) { return sqrtf(length2(a)); }
__device__ __forceinline__ f3 normalize(const f3& a) { return a / length(a); }
__device__ __forceinline__ bool is_black(const f3& c) { return c.x == 0.f && c.y == 0.f && c.z == 0.f; }
__device__ __forceinline__ float max_component(const f3& c) {  // std::max(r, std::max(g, b)), color.h:42-45
    float gb = (c.y < c.z) ? c.z : c.y;
    return (c.x < gb) ? gb : c.x;
}

// std::max / std::min semantics (second operand wins only on strict compare; NaN-asymmetric),
// used where the reference's NaN behaviour is observable.
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }

template <typename T>
__device__ __forceinline__ T clampf(T v, T lo, T hi) {  // pbrt.h:75-83
    if (v < lo) return lo;
    else if (v > hi) return hi;
    else return v;
}

// pbrt.h:39-46 (float-rounded, as the reference's constexpr initialisers produce them)
#define JPB_PI 3.14159274101257324219f          /* (float)3.14159265358979323846 */
#define JPB_2PI (2.0f * JPB_PI)
#define JPB_PI_OVER_2 (JPB_PI / 2.0f)
#define JPB_PI_OVER_4 (JPB_PI / 4.0f)
#define JPB_INV_PI (1.0f / JPB_PI)

// sinf / cosf are ~150 SASS instructions each once inlined (range reduction + Payne-Hanek slow path).
// One shared, out-of-line copy per kernel keeps the material kernels inside the instruction cache
// (profiles/: k_shade was 63 % stalled on instruction fetch when everything was inlined).
__device__ __noinline__ float jp_sinf(float x) { return sinf(x); }
__device__ __noinline__ float jp_cosf(float x) { return cosf(x); }

// FFrame(n): geometry.h:344-377.  n is re-normalised exactly as the reference's ctor does.
struct Frame {
    f3 s, t, n;
};
__device__ __forceinline__ Frame make_frame(const f3& nn) {
    Frame f;
    f.n = normalize(nn);
    f3 tmp = (fabsf(f.n.x) > 0.99f) ? mk3(0, 1, 0) : mk3(1, 0, 0);
    f.t = normalize(cross(f.n, tmp));
    f.s = normalize(cross(f.t, f.n));
    return f;
}
__device__ __forceinline__ f3 to_local(const Frame& f, const f3& w) { return mk3(dot(f.s, w), dot(f.t, w), dot(f.n, w)); }
__device__ __forceinline__ f3 to_world(const Frame& f, const f3& l) { return f.s * l.x + f.t * l.y + f.n * l.z; }

}  // namespace jpbrt
