// scene.cc -- Scene builder + OBJ reader.  See scene.h for the mapping to the reference API.
#include "scene.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace jetpbrt {

static void set3(float* d, Vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

void Scene::CreateCamera(Vec3 pos, Vec3 front, Vec3 up, float vfov_deg, int width, int height) {
    set3(camera_.pos, pos);
    set3(camera_.front, front);
    set3(camera_.up, up);
    camera_.vfov_deg = vfov_deg;
    camera_.width = width;
    camera_.height = height;
}

int Scene::CreateEnvironmentLight(Vec3 radiance) {
    jpbrt_light l{};
    l.type = JPBRT_LIGHT_ENVIRONMENT;
    l.shape = -1;
    set3(l.color, radiance);
    lights_.push_back(l);
    return (int)lights_.size() - 1;
}

int Scene::CreatePointLight(Vec3 pos, Vec3 intensity) {
    jpbrt_light l{};
    l.type = JPBRT_LIGHT_POINT;
    l.shape = -1;
    set3(l.color, intensity);
    set3(l.pos, pos);
    lights_.push_back(l);
    return (int)lights_.size() - 1;
}

int Scene::CreateDirectionLight(Vec3 irradiance, Vec3 dir) {
    jpbrt_light l{};
    l.type = JPBRT_LIGHT_DIRECTION;
    l.shape = -1;
    set3(l.color, irradiance);
    set3(l.dir, dir);
    lights_.push_back(l);
    return (int)lights_.size() - 1;
}

int Scene::CreateMatteMaterial(Vec3 diffuse) {
    jpbrt_material m{};
    m.type = JPBRT_MAT_MATTE;
    set3(m.a, diffuse);
    materials_.push_back(m);
    return (int)materials_.size() - 1;
}

int Scene::CreateMirrorMaterial(Vec3 specular) {
    jpbrt_material m{};
    m.type = JPBRT_MAT_MIRROR;
    set3(m.a, specular);
    materials_.push_back(m);
    return (int)materials_.size() - 1;
}

int Scene::CreateGlassMaterial(float eta, Vec3 kr, Vec3 kt) {
    jpbrt_material m{};
    m.type = JPBRT_MAT_GLASS;
    m.f0 = eta;
    set3(m.a, kr);
    set3(m.b, kt);
    materials_.push_back(m);
    return (int)materials_.size() - 1;
}

int Scene::CreatePlasticMaterial(Vec3 kd, Vec3 ks, float roughness, bool remap) {
    jpbrt_material m{};
    m.type = JPBRT_MAT_PLASTIC;
    m.remap_roughness = remap ? 1 : 0;
    set3(m.a, kd);
    set3(m.b, ks);
    m.f0 = roughness;
    materials_.push_back(m);
    return (int)materials_.size() - 1;
}

int Scene::CreateMetalMaterial(Vec3 eta, Vec3 k, float urough, float vrough, bool remap) {
    jpbrt_material m{};
    m.type = JPBRT_MAT_METAL;
    m.remap_roughness = remap ? 1 : 0;
    set3(m.a, eta);
    set3(m.b, k);
    m.f0 = urough;
    m.f1 = vrough;
    materials_.push_back(m);
    return (int)materials_.size() - 1;
}

int Scene::CreateTriangle(Vec3 p0, Vec3 p1, Vec3 p2, bool flip_normal) {
    jpbrt_shape s{};
    s.type = JPBRT_SHAPE_TRIANGLE;
    s.flip_normal = flip_normal ? 1 : 0;
    set3(s.p[0], p0);
    set3(s.p[1], p1);
    set3(s.p[2], p2);
    shapes_.push_back(s);
    return (int)shapes_.size() - 1;
}

int Scene::CreateRectangle(Vec3 p0, Vec3 p1, Vec3 p2, Vec3 p3, bool flip_normal) {
    jpbrt_shape s{};
    s.type = JPBRT_SHAPE_RECTANGLE;
    s.flip_normal = flip_normal ? 1 : 0;
    set3(s.p[0], p0);
    set3(s.p[1], p1);
    set3(s.p[2], p2);
    set3(s.p[3], p3);
    shapes_.push_back(s);
    return (int)shapes_.size() - 1;
}

// Corner order as FRectangle::FromXY/XZ/YZ (shape.cc:76-95).
int Scene::CreateRectangleXY(float x0, float x1, float y0, float y1, float z, bool flip) {
    return CreateRectangle(Vec3(x0, y0, z), Vec3(x1, y0, z), Vec3(x1, y1, z), Vec3(x0, y1, z), flip);
}
int Scene::CreateRectangleXZ(float x0, float x1, float z0, float z1, float y, bool flip) {
    return CreateRectangle(Vec3(x0, y, z0), Vec3(x0, y, z1), Vec3(x1, y, z1), Vec3(x1, y, z0), flip);
}
int Scene::CreateRectangleYZ(float y0, float y1, float z0, float z1, float x, bool flip) {
    return CreateRectangle(Vec3(x, y0, z0), Vec3(x, y1, z0), Vec3(x, y1, z1), Vec3(x, y0, z1), flip);
}

int Scene::CreateSphere(Vec3 center, float radius) {
    jpbrt_shape s{};
    s.type = JPBRT_SHAPE_SPHERE;
    set3(s.p[0], center);
    s.p[1][0] = radius;
    shapes_.push_back(s);
    return (int)shapes_.size() - 1;
}

int Scene::CreateDisk(Vec3 pos, Vec3 normal, float radius) {
    jpbrt_shape s{};
    s.type = JPBRT_SHAPE_DISK;
    set3(s.p[0], pos);
    set3(s.p[1], normal);
    s.p[2][0] = radius;
    shapes_.push_back(s);
    return (int)shapes_.size() - 1;
}

std::vector<int> Scene::CreateTriangleMeshFromSoup(const std::vector<float>& tris, bool flip_normal,
                                                   bool flip_handedness, Vec3 offset, float scale) {
    std::vector<int> out;
    size_t n = tris.size() / 9;
    out.reserve(n);
    shapes_.reserve(shapes_.size() + n);
    for (size_t i = 0; i < n; ++i) {
        Vec3 v[3];
        for (int k = 0; k < 3; ++k) {
            const float* p = &tris[9 * i + 3 * k];
            v[k] = Vec3(p[0], p[1], p[2]);
            if (flip_handedness) v[k].z = -v[k].z;
            v[k].x *= scale; v[k].y *= scale; v[k].z *= scale;
            v[k].x += offset.x; v[k].y += offset.y; v[k].z += offset.z;
        }
        out.push_back(CreateTriangle(v[0], v[1], v[2], flip_normal));
    }
    return out;
}

std::vector<int> Scene::CreateTriangleMesh(const std::string& filename, bool flip_normal, bool flip_handedness,
                                           Vec3 offset, float scale) {
    std::vector<float> tris;
    std::string err;
    if (!LoadObjTriangles(filename, &tris, &err)) {
        fprintf(stdout, "load triangle mesh failed. %s (%s)\n", filename.c_str(), err.c_str());  // shape.cc:30
        return {};
    }
    return CreateTriangleMeshFromSoup(tris, flip_normal, flip_handedness, offset, scale);
}

int Scene::CreatePrimitive(int shape, int material, int light) {
    jpbrt_primitive p{shape, material, light};
    primitives_.push_back(p);
    return (int)primitives_.size() - 1;
}

std::vector<int> Scene::CreatePrimitives(const std::vector<int>& shapes, int material) {
    std::vector<int> out;
    out.reserve(shapes.size());
    primitives_.reserve(primitives_.size() + shapes.size());
    for (int s : shapes) out.push_back(CreatePrimitive(s, material, -1));
    return out;
}

int Scene::CreateAreaLight(Vec3 radiance, int shape, int material) {
    jpbrt_light l{};
    l.type = JPBRT_LIGHT_AREA;
    l.shape = shape;
    set3(l.color, radiance);
    lights_.push_back(l);
    int li = (int)lights_.size() - 1;
    CreatePrimitive(shape, material, li);
    return li;
}

std::vector<int> Scene::CreateAreaLights(Vec3 radiance, const std::vector<int>& shapes, int material) {
    std::vector<int> out;
    for (int s : shapes) out.push_back(CreateAreaLight(radiance, s, material));
    return out;
}

const jpbrt_scene_desc* Scene::Desc() {
    desc_.camera = camera_;
    desc_.max_depth = max_depth_;
    desc_.n_shapes = (int)shapes_.size();
    desc_.n_materials = (int)materials_.size();
    desc_.n_lights = (int)lights_.size();
    desc_.n_primitives = (int)primitives_.size();
    desc_.shapes = shapes_.data();
    desc_.materials = materials_.data();
    desc_.lights = lights_.data();
    desc_.primitives = primitives_.data();
    desc_.name = name_.c_str();
    return &desc_;
}

// ---------------------------------------------------------------------------------------------
// OBJ reader.  The reference goes through external/obj_loader.h and consumes the emitted vertex
// list in triples (shape.cc:36-62).  This reader keeps what that path observes: positions only,
// one triangle per 3-vertex face; faces with more vertices are fan-triangulated.  Path
// separators are normalised so the reference's "scene\\bunny\\bunny.obj" spelling also works.
bool LoadObjTriangles(const std::string& filename_in, std::vector<float>* tris, std::string* err) {
    std::string filename = filename_in;
    for (char& c : filename) if (c == '\\') c = '/';
    std::ifstream in(filename);
    if (!in) { if (err) *err = "cannot open"; return false; }
    std::vector<float> pos;
    std::string line;
    tris->clear();
    while (std::getline(in, line)) {
        if (line.size() < 2) continue;
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
            float x, y, z;
            if (sscanf(line.c_str() + 1, "%f %f %f", &x, &y, &z) == 3) { pos.push_back(x); pos.push_back(y); pos.push_back(z); }
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            std::istringstream ss(line.substr(1));
            std::string tok;
            std::vector<int> idx;
            while (ss >> tok) {
                int v = atoi(tok.c_str());  // "v", "v/vt", "v//vn", "v/vt/vn": atoi stops at '/'
                int nv = (int)(pos.size() / 3);
                if (v < 0) v = nv + v; else v = v - 1;
                if (v < 0 || v >= nv) { if (err) *err = "face index out of range"; return false; }
                idx.push_back(v);
            }
            for (size_t k = 2; k < idx.size(); ++k) {
                int tri[3] = {idx[0], idx[k - 1], idx[k]};
                for (int t : tri) for (int c = 0; c < 3; ++c) tris->push_back(pos[3 * t + c]);
            }
        }
    }
    if (tris->empty()) { if (err) *err = "no faces"; return false; }
    return true;
}

}  // namespace jetpbrt
