// scene_flatten.h -- host side of jpbrt_upload_scene: neutral description -> flattened arrays.
//
// Replaces, for the device path, what the reference does in the shape constructors
// (shape.h:280-289,383-393,479-485: stored normals, CheckThinness'd bounds), FScene::Preprocess
// (scene.cc:11-23: world bound, light preprocess, BVH build) and the material constructors
// (material.h:94-98: plastic Qd).  All derived quantities are computed with the reference's
// float expressions on the host (g++, no FMA contraction), so the device consumes bit-identical
// normals, areas and radii.
#pragma once

#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/jetpbrt_scene.h"
#include "dev_scene.h"

namespace jpbrt {

// std::vector whose resize() leaves trivially-constructible elements UNINITIALISED: the per-primitive arrays of a 5 M-triangle
// scene are 0.9 GB that every element of is written exactly once by the flattener's (parallel) loops; value-initialising them
// first was a serial 0.9 GB memset -- a fifth of the upload's host time.
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    using std::allocator<T>::allocator;
    template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
    template <class U, class... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
template <class T>
using BigVec = std::vector<T, DefaultInitAllocator<T>>;

struct HostScene {
    BigVec<Float4> qnodes;         // the same tree, 32-byte quantised nodes (dev_scene.h kQNodeStride) on the grid below
    float q_origin[3]{}, q_cell[3]{};
    BigVec<Float4> nodes, slots, slot_nrm;  // per node / per primitive (the shading frames of flat shapes are derived from
                                            // slot_nrm on the device: c_api.cu k_make_frames)
    std::vector<Float4> materials, lights;
    BigVec<Int2> slot_ml;
    BigVec<int> prim_slot;
    std::vector<int> inf_lights, nee_lights, pixel_order;
    DevCamera cam{};
    float world_min[3]{}, world_max[3]{};
    float world_radius = 0;
    int max_depth = 5;
    int width = 0, height = 0;
    int n_prims = 0;
    bool has_null_material = false;
    double bvh_build_seconds = 0;
    int bvh_builder = 0;  // 0 host binned SAH, 1 GPU LBVH, 2 host object-median rebuild (the SAH/LBVH tree was too deep)
    int max_leaf_prims = 0;  // most primitives any leaf holds
    int bvh_depth = 0;    // levels of the tree (root = 1); at most kMaxBvhDepth
    double bvh_device_seconds = 0;  // GPU builder only: upload of the boxes + kernels + download of the tree
    size_t Bytes() const {
        return (nodes.size() + slots.size() + slot_nrm.size() + materials.size() + lights.size()) * sizeof(Float4) +
               slot_ml.size() * sizeof(Int2) + (inf_lights.size() + prim_slot.size() + nee_lights.size()) * sizeof(int);  // pixel_order is film state, not scene
    }
};

// Deepest tree the traversal kernels can walk without losing a subtree: their node stack holds 64 entries, one of them
// the sentinel (csrc/intersect.cuh), and a ray has at most one pending subtree per level below the root.
constexpr int kMaxBvhDepth = 62;

// Trees of more nodes than this get no quantised copy (HostScene::qnodes stays empty): see use_qnodes() in c_api.cu.
constexpr int kQNodesMaxNodes = 1 << 20;

// BVH topology handed to the flattener by an external builder (the GPU LBVH builder, csrc/bvh_build.cuh): a binary
// radix tree over the primitives in `order`.  Inner node i covers order[first..last]; a child reference >= 0 is an
// inner node, < 0 is the single primitive at position ~ref of `order`.  Boxes are UNPADDED primitive-bounds unions.
struct BuiltNode {
    float mn[3], mx[3];
    int left, right;
    int first, last;
};
struct BuiltBvh {
    std::vector<BuiltNode> nodes;  // n_prims - 1 inner nodes, root = 0
    std::vector<int> order;        // sorted position -> primitive index
    double seconds = 0;            // device time + transfers
};
// prim_boxes: 6 floats per primitive (min xyz, max xyz).  Returns false to fall back to the host builder.
typedef bool (*BvhBuildFn)(void* user, const float* prim_boxes, int n_prims, BuiltBvh* out, std::string* err);

// Returns 0 or a negative jpbrt_status; *err receives a message.  With `build` the BVH topology comes from that
// builder (leaves are cut where a subtree holds <= kMaxLeafPrims primitives); without, from the host's binned-SAH builder.
int FlattenScene(const jpbrt_scene_desc* desc, HostScene* out, std::string* err, BvhBuildFn build = nullptr, void* build_user = nullptr);

// One shape -> one slot (+ normal/tag), as used by the unit kernels.  Returns false on a bad type.
bool MakeSlot(const jpbrt_shape& shape, int prim_index, Float4 slot[4], Float4* nrm, float bounds_min[3], float bounds_max[3]);

// One material -> kMaterialStride float4.
bool MakeMaterial(const jpbrt_material& m, Float4 out[3]);

}  // namespace jpbrt
