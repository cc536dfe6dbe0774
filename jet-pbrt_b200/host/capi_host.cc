// capi_host.cc -- host-only entry points of the C ABI: built-in scenes and the film writers.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "../../include/jetpbrt_b200.h"
#include "scene.h"

struct jpbrt_scene {
    std::unique_ptr<jetpbrt::Scene> scene;
};

namespace {

inline float Clamp01(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
// gamma_encoding, film.h:24
inline uint8_t GammaEncode(float x) { return (uint8_t)(std::pow(Clamp01(x), (float)(1 / 2.2)) * 255.0); }

// FFilm::SaveAsPPM, film.cc:45-60.  The reference streams gamma_encoding()'s uint8_t with operator<<, which writes each
// value as ONE RAW BYTE (a char), under a "P3" (text) header.  `text` = false reproduces those bytes exactly (the drop-in
// behaviour, byte-identical to the reference's file); `text` = true writes the numbers as decimal text, i.e. the
// well-formed P3 file the header promises (image kind 3).
bool SavePPM(const std::string& fn, int w, int h, const float* rgb, bool text) {
    std::ofstream f(fn, std::ios::binary);
    if (!f) return false;
    f << "P3\n" << w << " " << h << "\n255\n";
    for (int i = 0; i < w * h; ++i) {
        const uint8_t r = GammaEncode(rgb[3 * i]), g = GammaEncode(rgb[3 * i + 1]), b = GammaEncode(rgb[3 * i + 2]);
        if (text) f << (int)r << "  " << (int)g << "  " << (int)b << "\n";
        else f << (char)r << "  " << (char)g << "  " << (char)b << "\n";
    }
    return (bool)f;
}

// FFilm::SaveAsBMP, film.cc:62-144: 24-bit BGR, gamma 1/2.2, rows bottom-up.
bool SaveBMP(const std::string& fn, int w, int h, const float* rgb) {
    std::ofstream f(fn, std::ios::binary);
    if (!f) return false;
    const uint32_t line = ((uint32_t)w * 3 + 3) & ~3u;
    const uint32_t image = line * (uint32_t)h;
    uint8_t hdr[54] = {0};
    auto put16 = [&](int off, uint16_t v) { memcpy(hdr + off, &v, 2); };
    auto put32 = [&](int off, uint32_t v) { memcpy(hdr + off, &v, 4); };
    put16(0, 0x4d42);
    put32(2, 54 + image);
    put32(10, 54);
    put32(14, 40);
    put32(18, (uint32_t)w);
    put32(22, (uint32_t)h);
    put16(26, 1);
    put16(28, 24);
    f.write((const char*)hdr, 54);
    std::vector<uint8_t> row(line, 0);
    for (int y = h - 1; y >= 0; --y) {
        for (int x = 0; x < w; ++x) {
            const float* p = rgb + 3 * ((size_t)y * w + x);
            row[3 * x + 0] = GammaEncode(p[2]);
            row[3 * x + 1] = GammaEncode(p[1]);
            row[3 * x + 2] = GammaEncode(p[0]);
        }
        f.write((const char*)row.data(), line);
    }
    return (bool)f;
}

// FFilm::SaveAsHDR, film.cc:147-188: flat (non-RLE) RGBE.
bool SaveHDR(const std::string& fn, int w, int h, const float* rgb) {
    std::ofstream f(fn, std::ios::binary);
    if (!f) return false;
    f << "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " << h << " +X " << w << "\n";
    for (int i = 0; i < w * h; ++i) {
        uint8_t rgbe[4] = {0, 0, 0, 0};
        const float* c = rgb + 3 * (size_t)i;
        float v = std::max({c[0], c[1], c[2]});
        if (v >= 1e-32f) {
            int e;
            float m = float(std::frexp(v, &e) * 256.f / v);
            rgbe[0] = uint8_t(c[0] * m);
            rgbe[1] = uint8_t(c[1] * m);
            rgbe[2] = uint8_t(c[2] * m);
            rgbe[3] = uint8_t(e + 128);
        }
        f.write((const char*)rgbe, 4);
    }
    return (bool)f;
}

}  // namespace

extern "C" {

jpbrt_scene* jpbrt_scene_builtin(const char* name, int width, int height, float scale) {
    if (!name || width <= 0 || height <= 0) return nullptr;
    jetpbrt::Scene* s = jetpbrt::MakeBuiltinScene(name, width, height, scale);
    if (!s) return nullptr;
    jpbrt_scene* h = new jpbrt_scene();
    h->scene.reset(s);
    return h;
}

const jpbrt_scene_desc* jpbrt_scene_get_desc(jpbrt_scene* s) { return s ? s->scene->Desc() : nullptr; }

void jpbrt_scene_free(jpbrt_scene* s) { delete s; }

long long jpbrt_load_obj_triangles(const char* filename, int flip_handedness, const float* offset3, float scale, float* tris9,
                                   long long capacity) {
    if (!filename) return JPBRT_ERR_INVALID;
    std::vector<float> tris;
    std::string err;
    if (!jetpbrt::LoadObjTriangles(filename, &tris, &err)) return JPBRT_ERR_IO;
    const long long n = (long long)(tris.size() / 9);
    const float off[3] = {offset3 ? offset3[0] : 0.f, offset3 ? offset3[1] : 0.f, offset3 ? offset3[2] : 0.f};
    for (long long i = 0; i < n && i < capacity && tris9; ++i)
        for (int v = 0; v < 3; ++v) {
            float p[3] = {tris[9 * i + 3 * v], tris[9 * i + 3 * v + 1], tris[9 * i + 3 * v + 2]};
            if (flip_handedness) p[2] = -p[2];                    // LoadTriangleMesh's order, shape.cc:48-62:
            for (int c = 0; c < 3; ++c) p[c] = p[c] * scale;     //   negate z, then scale,
            for (int c = 0; c < 3; ++c) p[c] = p[c] + off[c];    //   then offset
            for (int c = 0; c < 3; ++c) tris9[9 * i + 3 * v + c] = p[c];
        }
    return n;
}

int jpbrt_save_image(const char* basename, int kind, int width, int height, const float* rgb) {
    if (!basename || !rgb || width <= 0 || height <= 0) return JPBRT_ERR_INVALID;
    std::string base(basename);
    bool ok = false;
    switch (kind) {  // EImageType, film.h:15-20
    case 0: ok = SavePPM(base + ".ppm", width, height, rgb, false); break;
    case 3: ok = SavePPM(base + ".ppm", width, height, rgb, true); break;
    case 1: ok = SaveBMP(base + ".bmp", width, height, rgb); break;
    case 2: ok = SaveHDR(base + ".hdr", width, height, rgb); break;
    default: return JPBRT_ERR_INVALID;
    }
    return ok ? JPBRT_OK : JPBRT_ERR_IO;
}

}  // extern "C"
