// dev_scene.h -- the flattened scene layout shared by the host uploader and the CUDA kernels.
//
// Everything the wavefront kernels read is a 16-byte-aligned array of float4 (SoA by role), so
// every fetch is one LDG.128 (DESIGN.md "Data layout in HBM"):
//
//   nodes     4 x float4 / node  (64 B)  two child boxes + two child references
//   slots     4 x float4 / slot  (64 B)  primitive geometry in BVH leaf order
//   slot_nrm  1 x float4 / slot          stored normal + tag, read on accepted hits and by shade
//   slot_ml   1 x int2   / slot          (material, light) of the primitive
//   slot_frame 3 x float4 / slot         the shading frame FFrame(normal) = (s, t, n') of flat shapes
//   materials 3 x float4 / material
//   lights    6 x float4 / light
#pragma once

#include <stdint.h>

namespace jpbrt {

// ---- node: n0 = (Lmin.x Lmin.y Lmin.z Lmax.x) n1 = (Lmax.y Lmax.z Rmin.x Rmin.y)
//            n2 = (Rmin.z Rmax.x Rmax.y Rmax.z) n3 = (int left, int right, -, -)
// child reference >= 0: inner node index;  < 0: leaf, ~ref = (first_slot << 4) | count
constexpr int kNodeStride = 4;
// ---- quantised node (32 B): q0 = (Lmin.x|Lmax.x<<16, Lmin.y|Lmax.y<<16, Lmin.z|Lmax.z<<16, Rmin.x|Rmax.x<<16)
//                             q1 = (Rmin.y|Rmax.y<<16, Rmin.z|Rmax.z<<16, int left, int right);  plane = q_origin + q * q_cell
constexpr int kQNodeStride = 2;
constexpr int kLeafCountBits = 4;
constexpr int kMaxLeafPrims = 4;

// ---- slot: q0.w carries the tag = type | (primitive_index << 2)
//   triangle : q0 = p0, q1 = p1, q2 = p2, q3 = stored normal (a copy of slot_nrm: no third fetch for an accepted candidate)
//   rectangle: q0 = p0, q1 = p1, q2 = p2, q3 = p3
//   sphere   : q0 = centre, q1.x = radius
//   disk     : q0 = position, q1 = normal(xyz) radius(w)
constexpr int kSlotStride = 4;
constexpr int kTypeBits = 2;

// ---- material: m0 = (a.rgb, type)  m1 = (b.rgb, f0)  m2 = (f1, Qd, -, -)
//   matte   a = albedo                      mirror  a = reflectance
//   glass   a = Kr, b = Kt, f0 = eta
//   plastic a = Kd/Qd, b = Ks/(1-Qd), f0 = alpha (clamped, remapped), m2.y = Qd
//   metal   a = eta, b = k, f0 = alpha_x, m2.x = alpha_y
constexpr int kMaterialStride = 3;

// ---- light: l0 = (color.rgb, type | shape_type << 8)  l1 = (p0 | centre | position | dir, 1/area)
//             l2 = (p1, radius)  l3 = (p2, -)  l4 = (stored normal, -)  l5 = spare
constexpr int kLightStride = 6;

// ---- shading frame of triangles / rectangles / disks, precomputed with FFrame's expressions
// (geometry.h:344-377): f0 = s, f1 = t, f2 = n' = Normalize(stored normal).  Spheres compute theirs per hit.
constexpr int kFrameStride = 3;

struct DevCamera {
    float pos[3];
    float front[3];
    float right[3];
    float up[3];
    float res_x, res_y;
};

struct Float4 { float x, y, z, w; };
struct Int2 { int x, y; };

// Passed to kernels by value (pointers are device pointers).
struct DevScene {
    const Float4* nodes;
    const Float4* qnodes;  // the same tree with 32-byte quantised nodes (intersect.cuh: JPB_QNODES)
    const Float4* slots;
    const Float4* slot_nrm;
    const Int2*   slot_ml;
    const Float4* materials;
    const Float4* lights;
    const Float4* slot_frame;
    const int*    nee_lights;  // indices of the lights whose colour is not black (the others can never contribute)
    const int*    inf_lights;
    const int*    prim_slot;  // primitive index -> slot
    const int*    pixel_order;  // path i of a wavefront renders pixel pixel_order[i % npix]: 8x4 tiles in Morton order
    int n_nodes, n_slots, n_materials, n_lights, n_inf_lights, n_prims, n_nee_lights;
    int max_depth;
    int max_leaf_prims;  // most primitives any leaf holds (<= kMaxLeafPrims unless JPBRT_BVH_LEAF asked for more)
    int width, height;
    float q_origin[3], q_cell[3];  // the quantised nodes' grid
    float world_radius;  // FEnvironmentLight::worldRadius (light.cc:26-33)
    DevCamera cam;
};

}  // namespace jpbrt
