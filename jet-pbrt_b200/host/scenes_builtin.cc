// scenes_builtin.cc -- the five BASELINE.json configurations as procedural scenes.
//
// The reference's scene files ("scene\\cornellbox\\*.obj", "scene\\bunny\\bunny.obj",
// main.cc:34-53,94-106) are not in its repository, so geometry is generated here and pushed
// through the SAME load transform the reference applies (flip handedness, scale, offset,
// flip_normal -- shape.cc:48-62).  Cameras, materials, radiances and creation order follow
// main.cc:13-111 literally.  SURVEY.md Appendix D / 8(d) define the stand-ins.
#include <cmath>
#include <cstring>

#include "scene.h"

namespace jetpbrt {

namespace {

// main.cc:35 / :75 -- evaluated in float, left to right, exactly as the reference's expression.
Vec3 CornellRadiance() {
    float a[3] = {0.747f + 0.058f, 0.747f + 0.258f, 0.747f};
    float b[3] = {0.740f + 0.287f, 0.740f + 0.160f, 0.740f};
    float c[3] = {0.737f + 0.642f, 0.737f + 0.159f, 0.737f};
    float r[3];
    for (int i = 0; i < 3; ++i) r[i] = (a[i] * 8.0f + b[i] * 15.6f) + c[i] * 18.4f;
    return Vec3(r[0], r[1], r[2]);
}

// main.cc:22,73 pass Normalize(lookat - lookfrom): v / sqrt(x*x + y*y + z*z) (geometry.h:108-111).
Vec3 LookDir(Vec3 from, Vec3 at) {
    Vec3 v(at.x - from.x, at.y - from.y, at.z - from.z);
    float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return Vec3(v.x / len, v.y / len, v.z / len);
}

void PushTri(std::vector<float>& t, const float* a, const float* b, const float* c) {
    t.insert(t.end(), a, a + 3);
    t.insert(t.end(), b, b + 3);
    t.insert(t.end(), c, c + 3);
}

// quad (a,b,c,d) -> triangles (a,b,c),(a,c,d)
void PushQuad(std::vector<float>& t, const float q[4][3]) {
    PushTri(t, q[0], q[1], q[2]);
    PushTri(t, q[0], q[2], q[3]);
}

// Closed UV sphere with optional smooth radial displacement; triangles wound so the geometric
// normal (p1-p0)x(p2-p0) points away from `c` (standard OBJ outward winding).
void MakeBlob(std::vector<float>& t, const float c[3], float radius, int lon, int lat, float bump) {
    const float kPi = 3.14159265358979323846f;
    auto P = [&](int i, int j, float* o) {
        float th = kPi * (float)i / (float)lat;
        float ph = 2.f * kPi * (float)(j % lon) / (float)lon;
        float s = std::sin(th);
        float r = radius * (1.f + bump * (0.55f * std::sin(3.f * ph + 1.f) * std::sin(2.f * th) * s +
                                          0.45f * std::cos(5.f * ph) * s * s * std::cos(3.f * th)));
        o[0] = c[0] + r * s * std::cos(ph);
        o[1] = c[1] + r * std::cos(th);
        o[2] = c[2] + r * s * std::sin(ph);
    };
    auto Emit = [&](const float* a, const float* b, const float* d) {
        float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        float e2[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
        float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        float g[3] = {(a[0] + b[0] + d[0]) / 3.f - c[0], (a[1] + b[1] + d[1]) / 3.f - c[1], (a[2] + b[2] + d[2]) / 3.f - c[2]};
        if (n[0] * g[0] + n[1] * g[1] + n[2] * g[2] >= 0) PushTri(t, a, b, d);
        else PushTri(t, a, d, b);
    };
    for (int i = 0; i < lat; ++i)
        for (int j = 0; j < lon; ++j) {
            float A[3], B[3], C[3], D[3];
            P(i, j, A); P(i + 1, j, B); P(i + 1, j + 1, C); P(i, j + 1, D);
            if (i == 0) Emit(A, B, C);                 // top cap: A == D (pole)
            else if (i == lat - 1) Emit(A, B, D);      // bottom cap: B == C (pole)
            else { Emit(A, B, C); Emit(A, C, D); }
        }
}

}  // namespace

// ---- C1 / C5: create_cornellbox_scene (main.cc:13-62) ------------------------------------------
Scene* MakeCornellBoxScene(int width, int height) {
    Scene* s = new Scene("cornell_box_scene");
    const Vec3 lookfrom(278, 273, 960), lookat(278, 273, 0);
    s->CreateCamera(lookfrom, LookDir(lookfrom, lookat), Vec3(0, 1, 0), 60.f, width, height);
    s->SetMaxDepth(5);
    s->CreateEnvironmentLight(Vec3(0, 0, 0));

    int red = s->CreateMatteMaterial(Vec3(0.63f, 0.065f, 0.05f));
    int green = s->CreateMatteMaterial(Vec3(0.14f, 0.45f, 0.091f));
    int white = s->CreateMatteMaterial(Vec3(0.725f, 0.71f, 0.68f));
    int golden = s->CreateMetalMaterial(Vec3(0.18f, 0.15f, 0.81f), Vec3(0.11f, 0.11f, 0.11f), 0.2f, 0.2f, false);
    int mat_light = s->CreateMatteMaterial(Vec3(0.65f, 0.65f, 0.65f));

    // Canonical Cornell-box data, in the OBJ files' (pre-flip) space.
    static const float light[4][3] = {{343, 548.7f, 227}, {343, 548.7f, 332}, {213, 548.7f, 332}, {213, 548.7f, 227}};
    static const float floor_[4][3] = {{552.8f, 0, 0}, {0, 0, 0}, {0, 0, 559.2f}, {549.6f, 0, 559.2f}};
    static const float ceil_[4][3] = {{556, 548.8f, 0}, {556, 548.8f, 559.2f}, {0, 548.8f, 559.2f}, {0, 548.8f, 0}};
    static const float back[4][3] = {{549.6f, 0, 559.2f}, {0, 0, 559.2f}, {0, 548.8f, 559.2f}, {556, 548.8f, 559.2f}};
    static const float left[4][3] = {{552.8f, 0, 0}, {549.6f, 0, 559.2f}, {556, 548.8f, 559.2f}, {556, 548.8f, 0}};
    static const float right[4][3] = {{0, 0, 559.2f}, {0, 0, 0}, {0, 548.8f, 0}, {0, 548.8f, 559.2f}};
    static const float shortbox[5][4][3] = {
        {{130, 165, 65}, {82, 165, 225}, {240, 165, 272}, {290, 165, 114}},
        {{290, 0, 114}, {290, 165, 114}, {240, 165, 272}, {240, 0, 272}},
        {{130, 0, 65}, {130, 165, 65}, {290, 165, 114}, {290, 0, 114}},
        {{82, 0, 225}, {82, 165, 225}, {130, 165, 65}, {130, 0, 65}},
        {{240, 0, 272}, {240, 165, 272}, {82, 165, 225}, {82, 0, 225}}};
    static const float tallbox[5][4][3] = {
        {{423, 330, 247}, {265, 330, 296}, {314, 330, 456}, {472, 330, 406}},
        {{423, 0, 247}, {423, 330, 247}, {472, 330, 406}, {472, 0, 406}},
        {{472, 0, 406}, {472, 330, 406}, {314, 330, 456}, {314, 0, 456}},
        {{314, 0, 456}, {314, 330, 456}, {265, 330, 296}, {265, 0, 296}},
        {{265, 0, 296}, {265, 330, 296}, {423, 330, 247}, {423, 0, 247}}};

    std::vector<float> t;
    PushQuad(t, light);
    auto shape_light = s->CreateTriangleMeshFromSoup(t, true, true);
    s->CreateAreaLights(CornellRadiance(), shape_light, mat_light);  // one light per triangle (scene.cc:79-89)

    t.clear(); PushQuad(t, floor_); PushQuad(t, ceil_); PushQuad(t, back);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), white);
    t.clear(); for (auto& q : shortbox) PushQuad(t, q);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), white);
    t.clear(); for (auto& q : tallbox) PushQuad(t, q);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), golden);
    t.clear(); PushQuad(t, left);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), red);
    t.clear(); PushQuad(t, right);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), green);

    // main.cc:56-58 also creates a glass material and an FSphere that is never bound to a primitive.
    s->CreateGlassMaterial(1.5f, Vec3(0.98f, 0.98f, 0.98f), Vec3(0.98f, 0.98f, 0.98f));
    s->CreateSphere(Vec3(273, 273, 150), 60.f);
    return s;
}

// ---- C2: create_bunny_scene (main.cc:64-111) ---------------------------------------------------
Scene* MakeBunnyScene(int width, int height, int mesh_lon, int mesh_lat) {
    Scene* s = new Scene("bunny_scene");
    const Vec3 lookfrom(-300, 300, -300);
    s->CreateCamera(lookfrom, LookDir(lookfrom, Vec3(0, 0, 0)), Vec3(0, 1, 0), 60.f, width, height);
    s->SetMaxDepth(5);
    s->CreateEnvironmentLight(Vec3(0.1f, 0.1f, 0.5f));

    int red = s->CreateMatteMaterial(Vec3(0.63f, 0.065f, 0.05f));
    int green = s->CreateMatteMaterial(Vec3(0.14f, 0.45f, 0.091f));
    s->CreateMatteMaterial(Vec3(0.725f, 0.71f, 0.68f));  // "white": created, unused (main.cc:80)
    int mat_light = s->CreateMatteMaterial(Vec3(0.65f, 0.65f, 0.65f));

    int shape_light = s->CreateRectangleXZ(-100, 100, -100, 100, 350, true);
    s->CreateAreaLight(CornellRadiance(), shape_light, mat_light);
    int floor_ = s->CreateRectangleXZ(-200, 200, -200, 200, 0);
    s->CreatePrimitive(floor_, green, -1);

    // Stand-in for bunny.obj: a bumpy closed blob in OBJ space (about the size and place of the
    // Stanford bunny: ~0.14 units tall, resting near y = 0), 2*lon*(lat-1) triangles; the default
    // 72 x 36 gives 5,040 per instance.
    std::vector<float> mesh;
    const float c[3] = {0.f, 0.078f, 0.f};
    MakeBlob(mesh, c, 0.062f, mesh_lon, mesh_lat, 0.22f);

    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(mesh, true, true, Vec3(0, 0, 0), 500.f), red);
    int plastic = s->CreatePlasticMaterial(Vec3(0.35f, 0.12f, 0.48f),
                                           Vec3(1.f - 0.35f, 1.f - 0.12f, 1.f - 0.48f), 0.1f, false);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(mesh, true, true, Vec3(-100, 0, -100), 500.f), plastic);
    int golden = s->CreateMetalMaterial(Vec3(0.18f, 0.15f, 0.81f), Vec3(0.11f, 0.11f, 0.11f), 0.2f, 0.2f, false);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(mesh, true, true, Vec3(0, 0, -100), 500.f), golden);
    int glass = s->CreateGlassMaterial(1.5f, Vec3(0.98f, 0.98f, 0.98f), Vec3(0.98f, 0.98f, 0.98f));
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(mesh, true, true, Vec3(-100, 0, 0), 500.f), glass);
    return s;
}

// ---- C3: ~5 M triangles, traversal / memory-bound stress ------------------------
Scene* MakeLargeMeshScene(int width, int height, int grid_n, int sphere_n) {
    Scene* s = new Scene("large_mesh_scene");
    const Vec3 lookfrom(-620, 420, -620), lookat(0, 60, 0);
    s->CreateCamera(lookfrom, LookDir(lookfrom, lookat), Vec3(0, 1, 0), 60.f, width, height);
    s->SetMaxDepth(8);
    s->CreateEnvironmentLight(Vec3(0.1f, 0.1f, 0.5f));
    int ground = s->CreateMatteMaterial(Vec3(0.55f, 0.5f, 0.42f));
    int mat_light = s->CreateMatteMaterial(Vec3(0.65f, 0.65f, 0.65f));
    int golden = s->CreateMetalMaterial(Vec3(0.18f, 0.15f, 0.81f), Vec3(0.11f, 0.11f, 0.11f), 0.2f, 0.2f, false);

    int shape_light = s->CreateRectangleXZ(-150, 150, -150, 150, 600, true);
    s->CreateAreaLight(CornellRadiance(), shape_light, mat_light);

    auto H = [](float x, float z) {
        return 20.f * (0.5f * std::sin(0.013f * x + 0.7f) * std::cos(0.017f * z) +
                       0.3f * std::sin(0.041f * x + 0.029f * z) + 0.2f * std::sin(0.11f * x) * std::sin(0.13f * z));
    };
    std::vector<float> t;
    t.reserve((size_t)grid_n * grid_n * 18);
    const float lo = -500.f, step = 1000.f / (float)grid_n;
    for (int i = 0; i < grid_n; ++i)
        for (int j = 0; j < grid_n; ++j) {
            float x0 = lo + step * i, x1 = lo + step * (i + 1), z0 = lo + step * j, z1 = lo + step * (j + 1);
            float a[3] = {x0, H(x0, z0), z0}, b[3] = {x0, H(x0, z1), z1}, c[3] = {x1, H(x1, z1), z1}, d[3] = {x1, H(x1, z0), z0};
            PushTri(t, a, b, c);  // (b-a)x(c-a) has +y: normals face up
            PushTri(t, a, c, d);
        }
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, false, false), ground);
    t.clear();
    t.shrink_to_fit();
    const float c[3] = {0, 200, 0};
    MakeBlob(t, c, 150.f, sphere_n, sphere_n, 0.f);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, false, false), golden);
    return s;
}

// ---- C4: glossy room, 16 area lights, depth 16 (shade divergence + shadow-ray stress) ----------
Scene* MakeGlossyLightsScene(int width, int height) {
    Scene* s = new Scene("glossy_lights_scene");
    const Vec3 lookfrom(278, 273, 960), lookat(278, 273, 0);
    s->CreateCamera(lookfrom, LookDir(lookfrom, lookat), Vec3(0, 1, 0), 60.f, width, height);
    s->SetMaxDepth(16);
    s->CreateEnvironmentLight(Vec3(0, 0, 0));
    int red = s->CreateMatteMaterial(Vec3(0.63f, 0.065f, 0.05f));
    int green = s->CreateMatteMaterial(Vec3(0.14f, 0.45f, 0.091f));
    int white = s->CreateMatteMaterial(Vec3(0.725f, 0.71f, 0.68f));
    int mat_light = s->CreateMatteMaterial(Vec3(0.65f, 0.65f, 0.65f));
    int floor_pl = s->CreatePlasticMaterial(Vec3(0.3f, 0.3f, 0.32f), Vec3(0.7f, 0.7f, 0.68f), 0.1f, false);
    int back_pl = s->CreatePlasticMaterial(Vec3(0.2f, 0.35f, 0.5f), Vec3(0.8f, 0.65f, 0.5f), 0.4f, false);
    int metal_a = s->CreateMetalMaterial(Vec3(0.18f, 0.15f, 0.81f), Vec3(0.11f, 0.11f, 0.11f), 0.05f, 0.05f, false);
    int metal_b = s->CreateMetalMaterial(Vec3(0.2f, 0.92f, 1.1f), Vec3(3.9f, 2.45f, 2.14f), 0.2f, 0.2f, false);
    int metal_c = s->CreateMetalMaterial(Vec3(0.18f, 0.15f, 0.81f), Vec3(0.11f, 0.11f, 0.11f), 0.4f, 0.1f, false);
    int glass = s->CreateGlassMaterial(1.5f, Vec3(0.98f, 0.98f, 0.98f), Vec3(0.98f, 0.98f, 0.98f));

    // 4 x 4 grid of 40 x 40 downward-facing rectangle lights just under the ceiling; radiance is
    // scaled so that the total emitted power equals the single Cornell light's (130 x 105).
    Vec3 rad = CornellRadiance();
    const float k = (130.f * 105.f) / (16.f * 40.f * 40.f);
    rad = Vec3(rad.x * k, rad.y * k, rad.z * k);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float cx = 98.f + 120.f * i, cz = -(100.f + 120.f * j);
            int sh = s->CreateRectangleXZ(cx - 20, cx + 20, cz - 20, cz + 20, 548.f, true);
            s->CreateAreaLight(rad, sh, mat_light);
        }

    // room (already in world space: z in [-559.2, 0], open towards the camera)
    s->CreatePrimitive(s->CreateRectangleXZ(0, 556, -559.2f, 0, 0), floor_pl, -1);
    s->CreatePrimitive(s->CreateRectangleXZ(0, 556, -559.2f, 0, 548.8f), white, -1);
    s->CreatePrimitive(s->CreateRectangleXY(0, 556, 0, 548.8f, -559.2f), back_pl, -1);
    s->CreatePrimitive(s->CreateRectangleYZ(0, 548.8f, -559.2f, 0, 556), red, -1);
    s->CreatePrimitive(s->CreateRectangleYZ(0, 548.8f, -559.2f, 0, 0), green, -1);

    static const float shortbox[5][4][3] = {
        {{130, 165, 65}, {82, 165, 225}, {240, 165, 272}, {290, 165, 114}},
        {{290, 0, 114}, {290, 165, 114}, {240, 165, 272}, {240, 0, 272}},
        {{130, 0, 65}, {130, 165, 65}, {290, 165, 114}, {290, 0, 114}},
        {{82, 0, 225}, {82, 165, 225}, {130, 165, 65}, {130, 0, 65}},
        {{240, 0, 272}, {240, 165, 272}, {82, 165, 225}, {82, 0, 225}}};
    static const float tallbox[5][4][3] = {
        {{423, 330, 247}, {265, 330, 296}, {314, 330, 456}, {472, 330, 406}},
        {{423, 0, 247}, {423, 330, 247}, {472, 330, 406}, {472, 0, 406}},
        {{472, 0, 406}, {472, 330, 406}, {314, 330, 456}, {314, 0, 456}},
        {{314, 0, 456}, {314, 330, 456}, {265, 330, 296}, {265, 0, 296}},
        {{265, 0, 296}, {265, 330, 296}, {423, 330, 247}, {423, 0, 247}}};
    std::vector<float> t;
    for (auto& q : shortbox) PushQuad(t, q);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), metal_a);
    t.clear();
    for (auto& q : tallbox) PushQuad(t, q);
    s->CreatePrimitives(s->CreateTriangleMeshFromSoup(t, true, true), metal_b);

    s->CreatePrimitive(s->CreateSphere(Vec3(186, 225, -168), 60.f), metal_c, -1);  // on the short box
    s->CreatePrimitive(s->CreateSphere(Vec3(420, 70, -120), 70.f), glass, -1);
    return s;
}

Scene* MakeBuiltinScene(const std::string& name, int width, int height, float scale) {
    if (scale <= 0) scale = 1.f;
    if (name == "cornell" || name == "cornell_box_scene") return MakeCornellBoxScene(width, height);
    if (name == "bunny" || name == "bunny_scene") {
        int lon = (int)std::lround(72 * scale), lat = (int)std::lround(36 * scale);
        return MakeBunnyScene(width, height, lon < 3 ? 3 : lon, lat < 2 ? 2 : lat);
    }
    if (name == "large" || name == "large_mesh_scene") {
        int g = (int)std::lround(1500 * scale), sp = (int)std::lround(500 * scale);
        return MakeLargeMeshScene(width, height, g < 2 ? 2 : g, sp < 3 ? 3 : sp);
    }
    if (name == "glossy" || name == "glossy_lights_scene") return MakeGlossyLightsScene(width, height);
    return nullptr;
}

}  // namespace jetpbrt
