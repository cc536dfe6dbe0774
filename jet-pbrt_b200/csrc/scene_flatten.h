// scene_flatten.h -- host side of jpbrt_upload_scene: neutral description -> flattened arrays.
//
// Replaces, for the device path, what the reference does in the shape constructors
// (shape.h:280-289,383-393,479-485: stored normals, CheckThinness'd bounds), FScene::Preprocess
// (scene.cc:11-23: world bound, light preprocess, BVH build) and the material constructors
// (material.h:94-98: plastic Qd).  All derived quantities are computed with the reference's
// float expressions on the host (g++, no FMA contraction), so the device consumes bit-identical
// normals, areas and radii.
#pragma once

#include <string>
#include <vector>

#include "../../include/jetpbrt_scene.h"
#include "dev_scene.h"

namespace jpbrt {

struct HostScene {
    std::vector<Float4> nodes, slots, slot_nrm, materials, lights, slot_frame;
    std::vector<Int2> slot_ml;
    std::vector<int> inf_lights, prim_slot, nee_lights, pixel_order;
    DevCamera cam{};
    float world_min[3]{}, world_max[3]{};
    float world_radius = 0;
    int max_depth = 5;
    int width = 0, height = 0;
    int n_prims = 0;
    bool has_null_material = false;
    double bvh_build_seconds = 0;
    size_t Bytes() const {
        return (nodes.size() + slots.size() + slot_nrm.size() + materials.size() + lights.size() + slot_frame.size()) * sizeof(Float4) +
               slot_ml.size() * sizeof(Int2) + (inf_lights.size() + prim_slot.size() + nee_lights.size()) * sizeof(int);  // pixel_order is film state, not scene
    }
};

// Returns 0 or a negative jpbrt_status; *err receives a message.
int FlattenScene(const jpbrt_scene_desc* desc, HostScene* out, std::string* err);

// One shape -> one slot (+ normal/tag), as used by the unit kernels.  Returns false on a bad type.
bool MakeSlot(const jpbrt_shape& shape, int prim_index, Float4 slot[4], Float4* nrm, float bounds_min[3], float bounds_max[3]);

// One material -> kMaterialStride float4.
bool MakeMaterial(const jpbrt_material& m, Float4 out[3]);

}  // namespace jpbrt
