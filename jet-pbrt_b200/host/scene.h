// scene.h -- host-side scene description builder (C++), the mirror of the reference's FScene
// factory API (reference src/scene.h:66-124, scene.cc:49-97) for the B200 path.
//
// The reference builds a graph of heap objects; here every Create* call appends a POD record to
// a jpbrt_scene_desc (include/jetpbrt_scene.h).  Names and argument meaning follow the
// reference so that a scene written against FScene reads the same here:
//
//   reference (main.cc:13-62)                         this file
//   scene->CreateCamera<FCamera>(pos,front,up,fov,res) Scene::CreateCamera(pos,front,up,fov,w,h)
//   scene->CreateLight<FEnvironmentLight>(p,1,rad)     Scene::CreateEnvironmentLight(rad)
//   scene->CreateMaterial<FMatteMaterial>(c)           Scene::CreateMatteMaterial(c)
//   scene->CreateTriangleMesh(file,flipN,flipH,off,s)  Scene::CreateTriangleMesh(file,flipN,flipH,off,s)
//   scene->CreateAreaLights(1,rad,shapes,mat)          Scene::CreateAreaLights(rad,shapes,mat)
//   scene->CreatePrimitives(mesh,mat)                  Scene::CreatePrimitives(mesh,mat)
//   scene->Preprocess()                                (done by jpbrt_upload_scene: bounds, BVH, upload)
#pragma once

#include <string>
#include <vector>

#include "../../include/jetpbrt_scene.h"

namespace jetpbrt {

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};

class Scene {
public:
    explicit Scene(const std::string& name) : name_(name) {}

    // ---- camera (camera.h:36) ----
    void CreateCamera(Vec3 pos, Vec3 front, Vec3 up, float vfov_deg, int width, int height);
    void SetMaxDepth(int d) { max_depth_ = d; }

    // ---- lights (scene.h:93-107); return the light index (creation order) ----
    int CreateEnvironmentLight(Vec3 radiance);
    int CreatePointLight(Vec3 pos, Vec3 intensity);
    int CreateDirectionLight(Vec3 irradiance, Vec3 dir);

    // ---- materials (scene.h:84-91); return the material index ----
    int CreateMatteMaterial(Vec3 diffuse);
    int CreateMirrorMaterial(Vec3 specular);
    int CreateGlassMaterial(float eta, Vec3 kr, Vec3 kt);
    int CreatePlasticMaterial(Vec3 kd, Vec3 ks, float roughness, bool remap);
    int CreateMetalMaterial(Vec3 eta, Vec3 k, float urough, float vrough, bool remap);

    // ---- shapes (scene.h:75-82, shape.cc:76-95); return the shape index ----
    int CreateTriangle(Vec3 p0, Vec3 p1, Vec3 p2, bool flip_normal);
    int CreateRectangle(Vec3 p0, Vec3 p1, Vec3 p2, Vec3 p3, bool flip_normal);
    int CreateRectangleXY(float x0, float x1, float y0, float y1, float z, bool flip_normal = false);
    int CreateRectangleXZ(float x0, float x1, float z0, float z1, float y, bool flip_normal = false);
    int CreateRectangleYZ(float y0, float y1, float z0, float z1, float x, bool flip_normal = false);
    int CreateSphere(Vec3 center, float radius);
    int CreateDisk(Vec3 pos, Vec3 normal, float radius);

    // Triangle soup -> FTriangle list with LoadTriangleMesh's transform order (shape.cc:48-62):
    // z negated if flip_handedness, then * scale, then + offset.  `tris` = 9 floats per triangle.
    std::vector<int> CreateTriangleMeshFromSoup(const std::vector<float>& tris, bool flip_normal,
                                                bool flip_handedness, Vec3 offset = Vec3(0, 0, 0), float scale = 1.f);
    // Wavefront OBJ file (scene.cc:49-64).  Returns an empty list on load failure, like the reference.
    std::vector<int> CreateTriangleMesh(const std::string& filename, bool flip_normal = false,
                                        bool flip_handedness = false, Vec3 offset = Vec3(0, 0, 0), float scale = 1.f);

    // ---- primitives (scene.h:109-118, scene.cc:66-97) ----
    int CreatePrimitive(int shape, int material, int light);
    std::vector<int> CreatePrimitives(const std::vector<int>& shapes, int material);
    int CreateAreaLight(Vec3 radiance, int shape, int material);                       // light + its primitive
    std::vector<int> CreateAreaLights(Vec3 radiance, const std::vector<int>& shapes, int material);  // one light PER shape

    // The description; pointers stay valid until the next Create* call or destruction.
    const jpbrt_scene_desc* Desc();

    int NumPrimitives() const { return (int)primitives_.size(); }
    const std::string& Name() const { return name_; }

private:
    std::string name_;
    jpbrt_camera camera_{};
    int max_depth_ = 5;
    std::vector<jpbrt_shape> shapes_;
    std::vector<jpbrt_material> materials_;
    std::vector<jpbrt_light> lights_;
    std::vector<jpbrt_primitive> primitives_;
    jpbrt_scene_desc desc_{};
};

// Minimal OBJ reader: emits one (9 float) triangle per face, polygons are fan-triangulated.
bool LoadObjTriangles(const std::string& filename, std::vector<float>* tris, std::string* err);

// ---- built-in scenes: the five BASELINE.json configs ----
// The reference's OBJ assets are absent, so geometry is procedural.
Scene* MakeCornellBoxScene(int width, int height);                 // C1 / C5  (main.cc:13-62)
Scene* MakeBunnyScene(int width, int height, int mesh_lon, int mesh_lat);  // C2 (main.cc:64-111), mesh stand-in
Scene* MakeLargeMeshScene(int width, int height, int grid_n, int sphere_n);  // C3
Scene* MakeGlossyLightsScene(int width, int height);               // C4
Scene* MakeBuiltinScene(const std::string& name, int width, int height, float scale_param);

}  // namespace jetpbrt
