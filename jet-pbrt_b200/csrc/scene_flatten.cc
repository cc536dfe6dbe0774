// scene_flatten.cc -- see scene_flatten.h.  Host only (compiled by g++, -O2, no FMA contraction).
#include "scene_flatten.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <future>
#include <limits>
#include <memory>
#include <new>
#include <thread>

#include "../../include/jetpbrt_b200.h"

namespace jpbrt {

namespace {

// Minimal float3 with the reference's evaluation order (geometry.h:65-159).
struct H3 {
    float x, y, z;
    H3() : x(0), y(0), z(0) {}
    H3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit H3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    H3 operator+(const H3& v) const { return H3(x + v.x, y + v.y, z + v.z); }
    H3 operator-(const H3& v) const { return H3(x - v.x, y - v.y, z - v.z); }
    H3 operator-() const { return H3(-x, -y, -z); }
    H3 operator*(float s) const { return H3(x * s, y * s, z * s); }
    H3 operator/(float s) const { return H3(x / s, y / s, z / s); }
    float Length() const { return std::sqrt(x * x + y * y + z * z); }
    H3 Normalize() const { return *this / Length(); }
    H3 Cross(const H3& v) const { return H3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }
};

constexpr float kPi = (float)3.14159265358979323846;

struct Box {
    float mn[3], mx[3];
    struct Uninitialised {};
    explicit Box(Uninitialised) {}  // (for arrays whose every element is assigned before it is read)
    Box() {
        for (int a = 0; a < 3; ++a) { mn[a] = std::numeric_limits<float>::max(); mx[a] = std::numeric_limits<float>::lowest(); }
    }
    void Add(const H3& p) {
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    void Add(const Box& b) {
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], b.mn[a]); mx[a] = std::max(mx[a], b.mx[a]); }
    }
    void CheckThinness(float th = JPBRT_THINNESS) {  // geometry.h:299-304
        for (int a = 0; a < 3; ++a)
            if (mn[a] == mx[a]) { mn[a] -= th; mx[a] += th; }
    }
    float HalfArea() const {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

// FFrame(n) tangent vectors (geometry.h:344-377), needed for the disk's bounds (shape.h:239-252).
void FrameST(const H3& nn, H3* s, H3* t, H3* nout = nullptr) {
    H3 n = nn.Normalize();
    H3 tmp = (std::fabs(n.x) > 0.99f) ? H3(0, 1, 0) : H3(1, 0, 0);
    *t = n.Cross(tmp).Normalize();
    *s = t->Cross(n).Normalize();
    if (nout) *nout = n;
}

inline float IntAsFloat(int v) { float f; memcpy(&f, &v, 4); return f; }

}  // namespace

bool MakeSlot(const jpbrt_shape& sh, int prim_index, Float4 q[4], Float4* nrm, float bmin[3], float bmax[3]) {
    const int tag = (sh.type & ((1 << kTypeBits) - 1)) | (prim_index << kTypeBits);
    const float tagf = IntAsFloat(tag);
    for (int i = 0; i < 4; ++i) q[i] = Float4{0, 0, 0, 0};
    H3 p0(sh.p[0]), p1(sh.p[1]), p2(sh.p[2]), p3(sh.p[3]);
    Box b;
    H3 n;
    switch (sh.type) {
    case JPBRT_SHAPE_TRIANGLE:  // shape.h:280-289, 339-349
        n = (p1 - p0).Cross(p2 - p0).Normalize();
        if (sh.flip_normal) n = -n;
        b.Add(p0); b.Add(p1); b.Add(p2);
        b.CheckThinness();
        q[0] = Float4{p0.x, p0.y, p0.z, tagf};
        q[1] = Float4{p1.x, p1.y, p1.z, 0};
        q[2] = Float4{p2.x, p2.y, p2.z, 0};
        q[3] = Float4{n.x, n.y, n.z, 0};  // the stored normal rides in the record's spare quarter: the second 256-bit load of the
                                          // triangle test brings it along, and an accepted candidate needs no third fetch
        break;
    case JPBRT_SHAPE_RECTANGLE:  // shape.h:383-393, 449-455
        n = (p1 - p0).Cross(p2 - p0).Normalize();
        if (sh.flip_normal) n = -n;
        b.Add(p0); b.Add(p1); b.Add(p2); b.Add(p3);
        b.CheckThinness();
        q[0] = Float4{p0.x, p0.y, p0.z, tagf};
        q[1] = Float4{p1.x, p1.y, p1.z, 0};
        q[2] = Float4{p2.x, p2.y, p2.z, 0};
        q[3] = Float4{p3.x, p3.y, p3.z, 0};
        break;
    case JPBRT_SHAPE_SPHERE: {  // shape.h:479-485, 541-545
        float r = sh.p[1][0];
        H3 half(r, r, r);
        b.Add(p0 + half); b.Add(p0 - half);
        q[0] = Float4{p0.x, p0.y, p0.z, tagf};
        q[1] = Float4{r, 0, 0, 0};
        n = H3(0, 0, 0);
        break;
    }
    case JPBRT_SHAPE_DISK: {  // shape.h:192-197, 239-252
        n = p1.Normalize();
        float r = sh.p[2][0];
        H3 s, t;
        FrameST(n, &s, &t);
        H3 rb = s * r, rt = t * r;
        b.Add(p0 + rb + rt); b.Add(p0 + rb - rt); b.Add(p0 - rb - rt); b.Add(p0 - rb + rt);
        b.CheckThinness();
        q[0] = Float4{p0.x, p0.y, p0.z, tagf};
        q[1] = Float4{n.x, n.y, n.z, r};
        break;
    }
    default:
        return false;
    }
    *nrm = Float4{n.x, n.y, n.z, tagf};
    for (int a = 0; a < 3; ++a) { bmin[a] = b.mn[a]; bmax[a] = b.mx[a]; }
    return true;
}

static float Luminance(const float* c) { return 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2]; }  // color.h:47-50

static float RoughnessToAlpha(float roughness) {  // microfacet.h:87-92
    roughness = std::max(roughness, (float)1e-3);
    float x = std::log(roughness);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

bool MakeMaterial(const jpbrt_material& m, Float4 out[3]) {
    out[0] = Float4{m.a[0], m.a[1], m.a[2], IntAsFloat(m.type)};
    out[1] = Float4{m.b[0], m.b[1], m.b[2], m.f0};
    out[2] = Float4{m.f1, 0, 0, 0};
    switch (m.type) {
    case JPBRT_MAT_MATTE:
    case JPBRT_MAT_MIRROR:
    case JPBRT_MAT_GLASS:
        return true;
    case JPBRT_MAT_PLASTIC: {  // material.h:94-98, material.cc:12-29
        float Ld = Luminance(m.a), Ls = Luminance(m.b);
        float L = Ld + Ls;
        float Qd = Ld / L;
        float rough = m.f0;
        if (m.remap_roughness) rough = RoughnessToAlpha(rough);
        float alpha = std::max(float(0.001), rough);  // microfacet.h:73-74
        float om = 1 - Qd;
        out[0] = Float4{m.a[0] / Qd, m.a[1] / Qd, m.a[2] / Qd, IntAsFloat(m.type)};
        out[1] = Float4{m.b[0] / om, m.b[1] / om, m.b[2] / om, alpha};
        out[2] = Float4{alpha, Qd, 0, 0};
        return true;
    }
    case JPBRT_MAT_METAL: {  // material.cc:31-43
        float ur = m.f0, vr = m.f1;
        if (m.remap_roughness) { ur = RoughnessToAlpha(ur); vr = RoughnessToAlpha(vr); }
        out[1].w = std::max(float(0.001), ur);
        out[2].x = std::max(float(0.001), vr);
        return true;
    }
    }
    return false;
}

// =================================================================================================
// BVH: binned-SAH top-down build on the host, flattened depth-first with the children's boxes
// stored in the parent.  Topology is deliberately NOT the reference's (bvh.h:59-92: random axis,
// median split, no ordering): hits are defined by the primitive tests, the tree only prunes.
// =================================================================================================
namespace {

struct TmpNode {
    Box box;
    int left, right;   // inner
    int first, count;  // leaf when count > 0
    // No initialisation: the builder sizes its node array for the worst case (2 N) up front and every node it hands out is
    // filled in completely by Build(); constructing 10 M empty nodes first was 0.3 s of a 5 M-primitive build.
    TmpNode() : box(Box::Uninitialised{}) {}
};

// Build knobs; overridable for experiments through the environment: JPBRT_BVH_LEAF (max primitives per
// leaf, <= 15) and JPBRT_BVH_TRAV (cost of a traversal step in primitive-test units x 100).  Defaults from
// the B200 sweep in profiles/r01_bvh_sweep.txt: leaf 4 / cost 1.0 (Cornell traversal 11 % faster than with
// cost 2.0, the bunny scene is flat within 3 % over leaf 2-8 x cost 0.5-4).
static int EnvInt(const char* name, int def) {
    const char* v = getenv(name);
    return v ? atoi(v) : def;
}

// fn(begin, end, part) over [0, n) cut into at most `parts` contiguous pieces, one std::thread each (the caller's included).
// Every use below combines the pieces' results with min / max / integer sums only, so the outcome does not depend on the cut.
template <typename F>
static void ParallelFor(size_t n, int parts, F&& fn) {
    parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, parts), n >> 14));
    if (parts <= 1) { fn((size_t)0, n, 0); return; }
    const size_t chunk = (n + parts - 1) / parts;
    std::vector<std::thread> th;
    th.reserve(parts - 1);
    for (int t = 1; t < parts; ++t)
        th.emplace_back([&fn, t, chunk, n]() { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk), t); });
    fn((size_t)0, std::min(n, chunk), 0);
    for (auto& x : th) x.join();
}

struct Builder {
    int max_leaf = std::max(1, std::min(15, EnvInt("JPBRT_BVH_LEAF", kMaxLeafPrims)));
    float trav_cost = EnvInt("JPBRT_BVH_TRAV", 100) * 0.01f;
    // Split search by set size (B200 sweep, profiles/r01_bvh_quality_sweep.txt): <= 1,024 primitives: 16 bins (an exact sweep
    // changes nothing on the bunny scene, helps Cornell by 4 % and costs the glossy scene 15 %); 1,025 .. sweep_hi: exact
    // sweep SAH (bunny scene: box tests per ray 34.3 -> 28.1, k_extend -6 %, for 70 ms of build); above: bins_big bins.
    int bins_big = EnvInt("JPBRT_BVH_BINS_BIG", 256);
    int sweep_hi = EnvInt("JPBRT_BVH_SWEEP_HI", 65536);  // (scenes of more than 2^18 primitives: bins only -- the sweep doubles a 5 M build for < 2 %)
    int sweep_max = EnvInt("JPBRT_BVH_SWEEP", 0);  // experiments: also sweep sets of at most this many primitives
    const std::vector<Box>& pb;
    BigVec<float> cx, cy, cz;
    BigVec<int> idx;
    std::vector<TmpNode> nodes;
    std::atomic<int> next{0};
    std::atomic<int> threads_left;
    int nthreads;  // host threads the build may use: subtree tasks (threads_left) and data-parallel passes over big sets
    bool force_median = false;  // object-median splits only: the fallback for trees SAH leaves deeper than the traversal stack
    bool chain_test = EnvInt("JPBRT_TEST_CHAIN_BVH", 0) != 0;  // TEST HOOK: one primitive peeled off per level (a tree as deep
                                                               // as the scene is large), to exercise the kernels' stack-overflow counter

    explicit Builder(const std::vector<Box>& prim_boxes, int nthreads_) : pb(prim_boxes), threads_left(nthreads_), nthreads(nthreads_) {
        size_t n = pb.size();
        cx.resize(n); cy.resize(n); cz.resize(n); idx.resize(n);
        ParallelFor(n, nthreads, [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) {
                cx[i] = 0.5f * (pb[i].mn[0] + pb[i].mx[0]);
                cy[i] = 0.5f * (pb[i].mn[1] + pb[i].mx[1]);
                cz[i] = 0.5f * (pb[i].mn[2] + pb[i].mx[2]);
                idx[i] = (int)i;
            }
        });
        nodes.resize(std::max<size_t>(2 * n, 2));
        if (n > (size_t)1 << 18) sweep_hi = 0;
    }
    float C(int axis, int i) const { return axis == 0 ? cx[i] : (axis == 1 ? cy[i] : cz[i]); }
    int Alloc() { return next.fetch_add(1); }

    // Sets of at least kParallelSet primitives (the top few levels of a multi-million-primitive tree, where there are fewer
    // subtrees than host threads) are bounded and binned by several threads at once: a share of the build's threads
    // proportional to the set's share of the scene.
    static constexpr int kParallelSet = 1 << 18;
    int WidthFor(int count) const {
        if (count < kParallelSet) return 1;
        return (int)std::max<long long>(1, (long long)nthreads * count / (long long)std::max<size_t>(1, pb.size()));
    }

    void Build(int node, int first, int last, const Box* known_box = nullptr, const Box* known_cbox = nullptr) {
        TmpNode& N = nodes[node];
        Box box, cbox;
        int count = last - first;
        if (known_box) {  // the parent's bins already hold this set's bounds
            box = *known_box;
            cbox = *known_cbox;
        } else if (const int width = WidthFor(count); width > 1) {
            std::vector<Box> pbx(width), pcb(width);
            ParallelFor((size_t)count, width, [&](size_t b, size_t e, int t) {
                Box bx, cb;
                for (size_t i = first + b; i < first + e; ++i) {
                    int p = idx[i];
                    bx.Add(pb[p]);
                    cb.Add(H3(cx[p], cy[p], cz[p]));
                }
                pbx[t] = bx;
                pcb[t] = cb;
            });
            for (int t = 0; t < width; ++t) { box.Add(pbx[t]); cbox.Add(pcb[t]); }
        } else {
            for (int i = first; i < last; ++i) {
                int p = idx[i];
                box.Add(pb[p]);
                cbox.Add(H3(cx[p], cy[p], cz[p]));
            }
        }
        N.box = box;
        if (chain_test && !force_median) {
            if (count == 1) { N.first = first; N.count = 1; return; }
            int l = Alloc(), r = Alloc();
            N.left = l; N.right = r; N.count = 0;
            Build(l, first, last - 1);
            Build(r, last - 1, last);
            return;
        }
        if (count <= max_leaf) {
            // a small set becomes a leaf unless SAH says that splitting it is clearly cheaper
            bool leaf = true;
            if (count > 1 && !force_median) {
                float best = BestSplitCost(first, last, box, cbox, nullptr, nullptr);
                leaf = !(best < (float)count);
            }
            if (leaf) { N.first = first; N.count = count; return; }
        }
        int axis = -1, mid = -1;
        ChildBounds kids;
        if (!force_median) BestSplitCost(first, last, box, cbox, &axis, &mid, &kids);
        if (mid <= first || mid >= last) {  // degenerate (coincident centroids): median by index
            int a = 0;
            float ext = -1;
            for (int k = 0; k < 3; ++k) { float e = cbox.mx[k] - cbox.mn[k]; if (e > ext) { ext = e; a = k; } }
            mid = first + count / 2;
            std::nth_element(idx.begin() + first, idx.begin() + mid, idx.begin() + last,
                             [&](int l, int r) { return C(a, l) < C(a, r); });
            kids.known = false;
        }
        int l = Alloc(), r = Alloc();
        N.left = l;
        N.right = r;
        N.count = 0;
        if (count > 1 << 15 && threads_left.fetch_sub(1) > 0) {
            const Box *lb = kids.known ? &kids.box[0] : nullptr, *lc = kids.known ? &kids.cbox[0] : nullptr;
            const Box *rb = kids.known ? &kids.box[1] : nullptr, *rc = kids.known ? &kids.cbox[1] : nullptr;
            auto fut = std::async(std::launch::async, [this, l, first, mid, lb, lc]() { Build(l, first, mid, lb, lc); });
            Build(r, mid, last, rb, rc);
            fut.get();
            threads_left.fetch_add(1);
        } else {
            if (count > 1 << 15) threads_left.fetch_add(1);
            const Box *lb = kids.known ? &kids.box[0] : nullptr, *lc = kids.known ? &kids.cbox[0] : nullptr;
            const Box *rb = kids.known ? &kids.box[1] : nullptr, *rc = kids.known ? &kids.cbox[1] : nullptr;
            Build(l, first, mid, lb, lc);
            Build(r, mid, last, rb, rc);
        }
    }

    // Exact SAH for small sets: every split position of the centroid-sorted order on every axis.
    float SweepSplitCost(int first, int last, const Box& box, int* out_axis, int* out_mid) {
        const int count = last - first;
        float best = std::numeric_limits<float>::infinity();
        int best_axis = -1, best_k = -1;
        const float parent_area = std::max(box.HalfArea(), 1e-30f);
        std::vector<int> order(idx.begin() + first, idx.begin() + last), best_order;
        std::vector<float> right_area(count);
        for (int a = 0; a < 3; ++a) {
            std::sort(order.begin(), order.end(), [&](int l, int r) { return C(a, l) < C(a, r) || (C(a, l) == C(a, r) && l < r); });
            Box acc;
            for (int k = count - 1; k > 0; --k) { acc.Add(pb[order[k]]); right_area[k] = acc.HalfArea(); }
            Box accl;
            for (int k = 1; k < count; ++k) {  // left = order[0..k), right = order[k..count)
                accl.Add(pb[order[k - 1]]);
                float cost = trav_cost + (accl.HalfArea() * k + right_area[k] * (count - k)) / parent_area;
                if (cost < best) { best = cost; best_axis = a; best_k = k; }
            }
            if (out_axis && best_axis == a) best_order = order;  // one copy per winning axis (not per improvement)
        }
        if (out_axis) {
            if (best_axis >= 0) {
                std::copy(best_order.begin(), best_order.end(), idx.begin() + first);
                *out_axis = best_axis;
                *out_mid = first + best_k;
            } else {
                *out_axis = -1;
                *out_mid = -1;
            }
        }
        return best;
    }

    // Bounds of the two sides of a split, handed to the children so that they need not walk their primitives for them.
    struct ChildBounds {
        bool known = false;
        Box box[2], cbox[2];
    };

    // Binned SAH over the three axes.  Returns the best cost (in primitive-test units, traversal
    // step = 1); if axis/mid are given, partitions idx[first,last) and reports the split -- and, through `kids`, the
    // primitive and centroid bounds of both sides, which are unions of the winning axis's bins.
    // ONE pass over the set fills the bins of all three axes (the pass is a gather of 36 bytes per primitive from three
    // arrays: memory-bound, and it used to run once per axis after a separate pass for the bounds).
    float BestSplitCost(int first, int last, const Box& box, const Box& cbox, int* out_axis, int* out_mid, ChildBounds* kids = nullptr) {
        constexpr int kMaxBins = 256;
        const int NB = (last - first) > 1024 ? std::min(kMaxBins, std::max(2, bins_big)) : 16;
        int count = last - first;
        if (count <= sweep_max || (count > 1024 && count <= sweep_hi)) return SweepSplitCost(first, last, box, out_axis, out_mid);
        float best = std::numeric_limits<float>::infinity();
        int best_axis = -1, best_bin = -1;
        float parent_area = std::max(box.HalfArea(), 1e-30f);
        const bool want_kids = kids != nullptr && out_axis != nullptr;
        struct Bins {
            alignas(Box) unsigned char bb_raw[3][sizeof(Box) * kMaxBins];  // primitive bounds per (axis, bin)
            alignas(Box) unsigned char cb_raw[3][sizeof(Box) * kMaxBins];  // centroid bounds per (axis, bin): only with want_kids
            int bc[3][kMaxBins];
            Box* bb(int a) { return reinterpret_cast<Box*>(bb_raw[a]); }
            Box* cb(int a) { return reinterpret_cast<Box*>(cb_raw[a]); }
            void Init(int nb, bool centroids) {  // (only the bins in use: most calls are for a few primitives and 16 bins)
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < nb; ++b) {
                        new (&bb(a)[b]) Box();
                        if (centroids) new (&cb(a)[b]) Box();
                        bc[a][b] = 0;
                    }
            }
        };
        float lo3[3], scale3[3];
        bool use[3];
        for (int a = 0; a < 3; ++a) {
            lo3[a] = cbox.mn[a];
            const float ext = cbox.mx[a] - cbox.mn[a];
            use[a] = ext > 0;
            scale3[a] = use[a] ? (float)NB / ext : 0.f;
        }
        auto fill = [&](Bins& B, size_t b0, size_t e0) {
            for (size_t i = b0; i < e0; ++i) {
                const int p = idx[i];
                const Box& pbx = pb[p];
                const H3 c(cx[p], cy[p], cz[p]);
                const float cc[3] = {c.x, c.y, c.z};
                for (int a = 0; a < 3; ++a) {
                    if (!use[a]) continue;
                    const int b = std::min(NB - 1, std::max(0, (int)((cc[a] - lo3[a]) * scale3[a])));
                    B.bb(a)[b].Add(pbx);
                    if (want_kids) B.cb(a)[b].Add(c);
                    B.bc[a][b]++;
                }
            }
        };
        Bins bins;
        bins.Init(NB, want_kids);
        if (const int width = WidthFor(count); width > 1) {  // big sets: several threads, each into its own bins
            std::vector<std::unique_ptr<Bins>> part(width);
            ParallelFor((size_t)count, width, [&](size_t b0, size_t e0, int t) {
                part[t].reset(new Bins);
                part[t]->Init(NB, want_kids);
                fill(*part[t], first + b0, first + e0);
            });
            for (int t = 0; t < width; ++t) {
                if (!part[t]) continue;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < NB; ++b)
                        if (part[t]->bc[a][b]) {
                            bins.bb(a)[b].Add(part[t]->bb(a)[b]);
                            if (want_kids) bins.cb(a)[b].Add(part[t]->cb(a)[b]);
                            bins.bc[a][b] += part[t]->bc[a][b];
                        }
            }
        } else {
            fill(bins, (size_t)first, (size_t)last);
        }
        for (int a = 0; a < 3; ++a) {
            if (!use[a]) continue;
            const Box* bb = bins.bb(a);
            const int* bc = bins.bc[a];
            float right_area[kMaxBins];
            int right_cnt[kMaxBins];
            Box acc;
            int c = 0;
            for (int b = NB - 1; b > 0; --b) {
                acc.Add(bb[b]);
                c += bc[b];
                right_area[b] = c ? acc.HalfArea() : 0.f;
                right_cnt[b] = c;
            }
            Box accl;
            int cl = 0;
            for (int b = 0; b < NB - 1; ++b) {
                accl.Add(bb[b]);
                cl += bc[b];
                if (cl == 0 || right_cnt[b + 1] == 0) continue;
                float cost = trav_cost + (accl.HalfArea() * cl + right_area[b + 1] * right_cnt[b + 1]) / parent_area;
                if (cost < best) { best = cost; best_axis = a; best_bin = b; }
            }
        }
        if (out_axis && best_axis >= 0) {
            int a = best_axis;
            const float lo = lo3[a], scale = scale3[a];
            auto it = std::partition(idx.begin() + first, idx.begin() + last, [&](int p) {
                int b = std::min(NB - 1, std::max(0, (int)((C(a, p) - lo) * scale)));
                return b <= best_bin;
            });
            *out_axis = a;
            *out_mid = (int)(it - idx.begin());
            if (want_kids) {
                for (int b = 0; b < NB; ++b) {
                    if (!bins.bc[a][b]) continue;
                    const int side = b <= best_bin ? 0 : 1;
                    kids->box[side].Add(bins.bb(a)[b]);
                    kids->cbox[side].Add(bins.cb(a)[b]);
                }
                kids->known = true;
            }
        } else if (out_axis) {
            *out_axis = -1;
            *out_mid = -1;
        }
        (void)count;
        return best;
    }
};

int LeafRef(int first, int count) { return ~((first << kLeafCountBits) | count); }

// Insertion-based optimisation of a built tree (Bittner, Hapala, Havran 2013, "Fast insertion-based optimization of
// bounding volume hierarchies"): every node in turn, largest boxes first, is cut out of the tree together with its parent
// and re-inserted where it increases the tree's surface-area cost least (branch-and-bound search over the tree).  The
// leaves -- the primitive ranges -- are untouched, so the slot order and every leaf reference stay valid; only the inner
// topology and the boxes change.  Returns the number of nodes that moved.
struct Reinserter {
    std::vector<TmpNode>& nodes;
    int& root;
    std::vector<int> parent;
    explicit Reinserter(std::vector<TmpNode>& n, int& r) : nodes(n), root(r), parent(n.size(), -1) {}

    static float Area(const Box& b) { return b.HalfArea(); }
    static Box Union(const Box& a, const Box& b) { Box u = a; u.Add(b); return u; }
    bool IsLeaf(int i) const { return nodes[i].count > 0; }

    void LinkParents(int r) {
        std::vector<int> st{r};
        parent[r] = -1;
        while (!st.empty()) {
            int i = st.back();
            st.pop_back();
            if (IsLeaf(i)) continue;
            parent[nodes[i].left] = i;
            parent[nodes[i].right] = i;
            st.push_back(nodes[i].left);
            st.push_back(nodes[i].right);
        }
    }
    void Refit(int i) {  // i and its ancestors
        for (; i >= 0; i = parent[i]) nodes[i].box = Union(nodes[nodes[i].left].box, nodes[nodes[i].right].box);
    }
    int Depth(int r) const {
        int best = 0;
        std::vector<std::pair<int, int>> st{{r, 0}};
        while (!st.empty()) {
            auto [i, d] = st.back();
            st.pop_back();
            best = std::max(best, d);
            if (!IsLeaf(i)) { st.push_back({nodes[i].left, d + 1}); st.push_back({nodes[i].right, d + 1}); }
        }
        return best;
    }
    double Cost(int r) const {  // sum of inner-node areas (the leaves' share does not change)
        double c = 0;
        std::vector<int> st{r};
        while (!st.empty()) {
            int i = st.back();
            st.pop_back();
            if (IsLeaf(i)) continue;
            c += Area(nodes[i].box);
            st.push_back(nodes[i].left);
            st.push_back(nodes[i].right);
        }
        return c;
    }

    // Search budget of a pass, in heap pops.  The branch-and-bound search prunes by box area; where boxes overlap massively
    // (10^5 concentric primitives: 73 s of "optimisation" measured on the GPU box) it cannot prune, so the pass stops
    // re-inserting once it has spent 256 pops per node -- an ordinary scene uses a tenth of that.
    mutable long long work = 0;
    long long budget = 0;

    // Best node X to pair L with: minimises Area(X u L) + the area added to X's ancestors.
    int FindBest(int L) const {
        const Box& lb = nodes[L].box;
        const float larea = Area(lb);
        struct Item { float induced; int node; };
        auto cmp = [](const Item& a, const Item& b) { return a.induced > b.induced; };
        std::vector<Item> heap{{0.f, root}};
        float best_cost = std::numeric_limits<float>::infinity();
        int best = root;
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            Item it = heap.back();
            heap.pop_back();
            ++work;
            if (it.induced + larea >= best_cost) break;  // every remaining candidate costs at least this much
            const float direct = Area(Union(nodes[it.node].box, lb));
            const float total = it.induced + direct;
            if (total < best_cost) { best_cost = total; best = it.node; }
            if (IsLeaf(it.node)) continue;
            const float child_induced = total - Area(nodes[it.node].box);  // what this node's box would grow by
            if (child_induced + larea < best_cost) {
                heap.push_back({child_induced, nodes[it.node].left});
                std::push_heap(heap.begin(), heap.end(), cmp);
                heap.push_back({child_induced, nodes[it.node].right});
                std::push_heap(heap.begin(), heap.end(), cmp);
            }
        }
        return best;
    }

    int Pass() {
        std::vector<int> order;
        for (int i = 0; i < (int)nodes.size(); ++i)
            if (i != root && parent[i] >= 0 && parent[i] != root) order.push_back(i);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return Area(nodes[a].box) > Area(nodes[b].box); });
        int moved = 0;
        if (budget == 0) budget = 256ll * (long long)order.size();
        for (int L : order) {
            if (work > budget) break;  // out of search budget: keep what was gained so far
            const int P = parent[L];
            if (P < 0 || P == root) continue;  // the tree changed under us
            const int G = parent[P];
            const int S = nodes[P].left == L ? nodes[P].right : nodes[P].left;
            // cut L and P out: S takes P's place under G
            (nodes[G].left == P ? nodes[G].left : nodes[G].right) = S;
            parent[S] = G;
            Refit(G);
            const int X = FindBest(L);
            // P becomes the parent of (X, L) where X was
            const int XP = parent[X];
            nodes[P].left = X;
            nodes[P].right = L;
            nodes[P].count = 0;
            parent[X] = P;
            parent[L] = P;
            parent[P] = XP;
            if (XP < 0) root = P;
            else (nodes[XP].left == X ? nodes[XP].left : nodes[XP].right) = P;
            Refit(P);
            if (X != S) ++moved;
        }
        return moved;
    }
};

}  // namespace

int FlattenScene(const jpbrt_scene_desc* d, HostScene* out, std::string* err, BvhBuildFn build, void* build_user) {
    auto fail = [&](int code, const char* msg) { if (err) *err = msg; return code; };
    if (!d || !out) return fail(JPBRT_ERR_INVALID, "null scene description");
    if (d->n_primitives <= 0 || !d->primitives || !d->shapes) return fail(JPBRT_ERR_INVALID, "scene has no primitives");
    if (d->camera.width <= 0 || d->camera.height <= 0) return fail(JPBRT_ERR_INVALID, "film resolution must be positive");
    if (d->max_depth < 0 || d->max_depth > 126) return fail(JPBRT_ERR_INVALID, "max_depth out of range [0,126]");
    if ((long long)d->n_primitives >= (1ll << (31 - kLeafCountBits))) return fail(JPBRT_ERR_UNSUPPORTED, "too many primitives");

    HostScene& hs = *out;
    hs = HostScene();
    // JPBRT_FLATTEN_TIMING=1: phase times on stderr (where a 5 M-primitive upload spends its host seconds)
    const bool timing = EnvInt("JPBRT_FLATTEN_TIMING", 0) != 0;
    auto t_phase = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[flatten] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_phase).count());
        t_phase = now;
    };
    hs.max_depth = d->max_depth;
    hs.width = d->camera.width;
    hs.height = d->camera.height;
    hs.n_prims = d->n_primitives;

    {  // FCamera ctor, camera.h:36-49
        const jpbrt_camera& c = d->camera;
        H3 pos(c.pos), front = H3(c.front).Normalize(), up = H3(c.up).Normalize();
        float res_x = (float)c.width, res_y = (float)c.height;
        float tan_fov = std::tan(((c.vfov_deg * kPi) / (float)180) / 2);
        float aspect = res_x / res_y;
        H3 right = up.Cross(front).Normalize() * (tan_fov * aspect);
        H3 up2 = front.Cross(right).Normalize() * tan_fov;
        float* dst[4] = {hs.cam.pos, hs.cam.front, hs.cam.right, hs.cam.up};
        H3 src[4] = {pos, front, right, up2};
        for (int i = 0; i < 4; ++i) { dst[i][0] = src[i].x; dst[i][1] = src[i].y; dst[i][2] = src[i].z; }
        hs.cam.res_x = res_x;
        hs.cam.res_y = res_y;
    }

    {   // Pixel visiting order of a wavefront: a warp's 32 consecutive paths cover one 8 x 4 pixel tile and tiles
        // follow a Morton curve, so the rays a warp traces -- and the surface points they hit -- are neighbours
        // (the reference walks rows, integrator.cc:90-92; the order is unobservable in the film).
        const int W = hs.width, H = hs.height;
        auto spread = [](uint32_t v) {  // interleave 16 bits with zeros
            v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f;
            v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v;
        };
        std::vector<std::pair<uint64_t, int>> keyed((size_t)W * H);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                uint64_t tile = (uint64_t)(spread((uint32_t)(x >> 3)) | (spread((uint32_t)(y >> 2)) << 1));
                keyed[(size_t)y * W + x] = {(tile << 5) | (uint64_t)(((y & 3) << 3) | (x & 7)), y * W + x};
            }
        std::sort(keyed.begin(), keyed.end());
        hs.pixel_order.resize(keyed.size());
        for (size_t i = 0; i < keyed.size(); ++i) hs.pixel_order[i] = keyed[i].second;
    }

    // materials
    hs.materials.resize((size_t)d->n_materials * kMaterialStride);
    for (int i = 0; i < d->n_materials; ++i)
        if (!MakeMaterial(d->materials[i], &hs.materials[(size_t)i * kMaterialStride])) return fail(JPBRT_ERR_INVALID, "unknown material type");

    // primitives -> slots (pre-BVH order), bounds
    const int N = d->n_primitives;
    BigVec<Float4> pre_slots((size_t)N * kSlotStride), pre_nrm(N);  // (uninitialised: every element is written below)
    std::vector<Box> boxes(N);
    Box world;
    const int nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    {
        std::vector<Box> part_world(nthreads);
        std::atomic<int> bad{0}, null_material{0};  // bad: the first kind of invalid reference any thread met
        ParallelFor((size_t)N, nthreads, [&](size_t b, size_t e, int t) {
            Box w;
            for (size_t i = b; i < e; ++i) {
                const jpbrt_primitive& p = d->primitives[i];
                int code = 0;
                if (p.shape < 0 || p.shape >= d->n_shapes) code = 1;
                else if (p.material >= d->n_materials) code = 2;
                else if (p.light >= d->n_lights) code = 3;
                else if (!MakeSlot(d->shapes[p.shape], (int)i, &pre_slots[i * kSlotStride], &pre_nrm[i], boxes[i].mn, boxes[i].mx)) code = 4;
                if (code) { int none = 0; bad.compare_exchange_strong(none, code); return; }
                if (p.material < 0) null_material.store(1, std::memory_order_relaxed);
                w.Add(boxes[i]);  // FScene::CalculateWorldBound, scene.cc:35-45
            }
            part_world[t] = w;
        });
        switch (bad.load()) {
        case 1: return fail(JPBRT_ERR_INVALID, "primitive references a missing shape");
        case 2: return fail(JPBRT_ERR_INVALID, "primitive references a missing material");
        case 3: return fail(JPBRT_ERR_INVALID, "primitive references a missing light");
        case 4: return fail(JPBRT_ERR_INVALID, "unknown shape type");
        }
        if (null_material.load()) hs.has_null_material = true;
        for (const Box& w : part_world) world.Add(w);
    }
    for (int a = 0; a < 3; ++a) { hs.world_min[a] = world.mn[a]; hs.world_max[a] = world.mx[a]; }
    phase("primitive records + boxes");
    {  // FBounds3::BoundingSphere (geometry.h:307-311) as used by FEnvironmentLight::Preprocess (light.cc:26-33)
        H3 mn(world.mn), mx(world.mx);
        H3 c = mn + (mx - mn) * 0.5f;  // Lerp(u, v, t) = u + t * (v - u), geometry.h:137
        bool inside = c.x >= mn.x && c.x <= mx.x && c.y >= mn.y && c.y <= mx.y && c.z >= mn.z && c.z <= mx.z;
        hs.world_radius = inside ? (c - mx).Length() : 0.f;
    }

    // lights
    hs.lights.assign((size_t)d->n_lights * kLightStride, Float4{0, 0, 0, 0});
    for (int i = 0; i < d->n_lights; ++i) {
        const jpbrt_light& l = d->lights[i];
        // a light whose colour is black yields Li == black for every sample and is skipped by
        // integrator.cc:362-364; it keeps its place in the sampler's dimension schedule
        if (!(l.color[0] == 0.f && l.color[1] == 0.f && l.color[2] == 0.f)) hs.nee_lights.push_back(i);
        Float4* o = &hs.lights[(size_t)i * kLightStride];
        int shape_type = 0;
        switch (l.type) {
        case JPBRT_LIGHT_ENVIRONMENT:
            hs.inf_lights.push_back(i);  // scene.h:101-104
            break;
        case JPBRT_LIGHT_AREA: {
            if (l.shape < 0 || l.shape >= d->n_shapes) return fail(JPBRT_ERR_INVALID, "area light references a missing shape");
            const jpbrt_shape& sh = d->shapes[l.shape];
            shape_type = sh.type;
            Float4 q[4], nrm;
            float bmn[3], bmx[3];
            if (!MakeSlot(sh, 0, q, &nrm, bmn, bmx)) return fail(JPBRT_ERR_INVALID, "unknown shape type");
            H3 p0(sh.p[0]), p1(sh.p[1]), p2(sh.p[2]);
            float area = 0, radius = 0;
            switch (sh.type) {
            case JPBRT_SHAPE_TRIANGLE: area = 0.5f * (p1 - p0).Cross(p2 - p0).Length(); break;                  // shape.h:351
            case JPBRT_SHAPE_RECTANGLE: area = (p0 - p1).Cross(p2 - p1).Length(); break;                       // shape.h:457
            case JPBRT_SHAPE_SPHERE: radius = sh.p[1][0]; area = 4 * kPi * (radius * radius); break;           // shape.h:547
            case JPBRT_SHAPE_DISK: radius = sh.p[2][0]; area = kPi * radius * radius; break;                   // shape.h:254
            }
            o[1] = Float4{p0.x, p0.y, p0.z, 1 / area};
            o[2] = Float4{p1.x, p1.y, p1.z, radius};
            o[3] = Float4{p2.x, p2.y, p2.z, 0};
            o[4] = Float4{nrm.x, nrm.y, nrm.z, 0};
            break;
        }
        case JPBRT_LIGHT_POINT:
            o[1] = Float4{l.pos[0], l.pos[1], l.pos[2], 0};
            break;
        case JPBRT_LIGHT_DIRECTION: {
            H3 dn = H3(l.dir).Normalize();  // worldDir(Normalize(worlddir)), light.h:142
            o[1] = Float4{dn.x, dn.y, dn.z, 0};
            break;
        }
        default:
            return fail(JPBRT_ERR_INVALID, "unknown light type");
        }
        o[0] = Float4{l.color[0], l.color[1], l.color[2], IntAsFloat(l.type | (shape_type << 8))};
    }

    // BVH
    auto t0 = std::chrono::steady_clock::now();
    // Conservative padding of every stored box: the reference's edge tests round, so a primitive
    // can report a hit a few ulps outside its own bounds (DESIGN.md "Conservative boxes").
    float maxabs = 0;
    for (int a = 0; a < 3; ++a) maxabs = std::max(maxabs, std::max(std::fabs(world.mn[a]), std::fabs(world.mx[a])));
    // ray origins are the camera position or points on surfaces: include the camera in the magnitude
    for (int a = 0; a < 3; ++a) maxabs = std::max(maxabs, std::fabs(d->camera.pos[a]));
    const float pad = 4e-6f * maxabs + 1e-30f;
    // (Scenes small enough for the insertion-based optimisation below are built by ONE thread: the pass visits nodes in an order
    // that depends on their temporary ids, which concurrent subtree tasks hand out in timing order -- the same scene must
    // flatten to the same tree every time.)
    Builder bld(boxes, N <= (1 << 18) ? 1 : nthreads);
    phase("builder setup (centroids)");
    int root = -1;
    if (build && N >= 2) {
        // external (GPU) builder: a binary radix tree over the Morton-sorted primitives; cut it into our leaves here
        static_assert(sizeof(Box) == 6 * sizeof(float), "Box must be 6 packed floats");
        BuiltBvh built;
        std::string berr;
        if (build(build_user, reinterpret_cast<const float*>(boxes.data()), N, &built, &berr) && (int)built.order.size() == N &&
            (int)built.nodes.size() == N - 1) {
            bld.idx.assign(built.order.begin(), built.order.end());
            bld.nodes.clear();
            bld.nodes.reserve((size_t)2 * N);
            struct Todo { int built_ref; int tmp; int depth; };
            int max_depth = 0;
            std::vector<Todo> todo;
            auto new_tmp = [&]() { bld.nodes.emplace_back(); return (int)bld.nodes.size() - 1; };
            root = new_tmp();
            todo.push_back(Todo{0, root, 0});
            while (!todo.empty()) {
                Todo t = todo.back();
                todo.pop_back();
                max_depth = std::max(max_depth, t.depth);
                if (t.built_ref < 0) {  // a single primitive
                    const int pos = ~t.built_ref;
                    TmpNode& T = bld.nodes[t.tmp];
                    T.box = boxes[built.order[pos]];
                    T.first = pos;
                    T.count = 1;
                    continue;
                }
                const BuiltNode& b = built.nodes[t.built_ref];
                for (int a = 0; a < 3; ++a) { bld.nodes[t.tmp].box.mn[a] = b.mn[a]; bld.nodes[t.tmp].box.mx[a] = b.mx[a]; }
                const int count = b.last - b.first + 1;
                if (count <= bld.max_leaf && count <= kMaxLeafPrims) {
                    bld.nodes[t.tmp].first = b.first;
                    bld.nodes[t.tmp].count = count;
                    continue;
                }
                const int l = new_tmp(), r = new_tmp();
                bld.nodes[t.tmp].left = l;
                bld.nodes[t.tmp].right = r;
                bld.nodes[t.tmp].count = 0;
                todo.push_back(Todo{b.right, r, t.depth + 1});
                todo.push_back(Todo{b.left, l, t.depth + 1});
            }
            hs.bvh_builder = 1;
            hs.bvh_device_seconds = built.seconds;
            if (max_depth > 60) {  // deeper than the traversal stack (pathological duplicate centroids): use the SAH builder
                hs.bvh_builder = 0;
                root = -1;
                bld.nodes.clear();
        bld.nodes.resize(std::max<size_t>(2 * (size_t)N, 2));
                for (int i = 0; i < N; ++i) bld.idx[i] = i;
                bld.next = 0;
            }
        } else if (!berr.empty()) {
            return fail(JPBRT_ERR_CUDA, berr.c_str());
        }
    }
    if (root < 0) {
        root = bld.Alloc();
        bld.Build(root, 0, N);
        bld.nodes.resize((size_t)bld.next.load());  // the nodes handed out (the rest of the 2 N were never written)
    }
    phase("tree build");
    // Insertion-based optimisation of the topology: two passes for scenes of 1,025 .. 2^17 primitives (JPBRT_BVH_REINSERT
    // overrides the number of passes for any scene up to 2^18).  Measured on B200 (profiles/ab/r01_ab_reinsert.log): the
    // bunny scene's surface-area cost drops 6.98 -> 5.48 and the step is 2.7 % faster for ~0.1 s more build; Cornell
    // gains 3 %, but the glossy scene (33 nodes, 16 lights' worth of shadow rays) LOSES 3 % -- the heuristic is not the
    // kernel's cost for such tiny trees, so they keep the plain SAH tree.  Kept only if it lowers the tree's surface-area
    // cost and the tree stays shallower than the traversal stack.
    const int default_passes = (N > 1024 && N <= (1 << 17)) ? 2 : 0;
    if (const int passes = EnvInt("JPBRT_BVH_REINSERT", default_passes); passes > 0 && N <= (1 << 18) && bld.nodes[root].count == 0) {
        std::vector<TmpNode> backup = bld.nodes;
        const int root_backup = root;
        Reinserter re(bld.nodes, root);
        re.LinkParents(root);
        const double before = re.Cost(root);
        for (int p = 0; p < passes; ++p)
            if (re.Pass() == 0) break;
        if (!(re.Cost(root) < before) || re.Depth(root) > 56) {
            bld.nodes = backup;
            root = root_backup;
        }
    }

    // The traversal kernels keep at most kMaxBvhDepth pending subtrees per ray (csrc/intersect.cuh: kTraversalStack minus the
    // sentinel); a deeper tree would silently lose far children.  SAH on skewed input (long chains of nested or
    // near-coincident primitives) can exceed that: such a tree is rebuilt with object-median splits, whose depth is
    // ceil(log2(N / leaf)) <= 31.  The kernels additionally COUNT any push that finds the stack full (stack_overflows).
    // One pass over the finished topology: depth, inner nodes per subtree (the layout below places every node from these
    // counts alone) and the largest leaf.  The top levels fan out over host threads.
    std::vector<int> inner_below;  // per tmp node: inner nodes in its subtree, itself included (0 for a leaf)
    std::atomic<int> largest_leaf{0};
    std::function<int(int, int)> measure = [&](int n, int level) -> int {  // returns the subtree's depth in levels
        const TmpNode& T = bld.nodes[n];
        if (T.count > 0) {
            int seen = largest_leaf.load(std::memory_order_relaxed);
            while (T.count > seen && !largest_leaf.compare_exchange_weak(seen, T.count)) {}
            inner_below[n] = 0;
            return 1;
        }
        int dl, dr;
        if (level < 4 && nthreads > 1) {
            auto fut = std::async(std::launch::async, [&measure, &T, level]() { return measure(T.left, level + 1); });
            dr = measure(T.right, level + 1);
            dl = fut.get();
        } else {
            dl = measure(T.left, level + 1);
            dr = measure(T.right, level + 1);
        }
        inner_below[n] = 1 + inner_below[T.left] + inner_below[T.right];
        return 1 + std::max(dl, dr);
    };
    auto tree_depth = [&](int r) {
        inner_below.assign(bld.nodes.size(), 0);
        largest_leaf.store(0);
        return measure(r, 0);
    };
    hs.bvh_depth = tree_depth(root);
    const int depth_limit = std::max(1, std::min(kMaxBvhDepth, EnvInt("JPBRT_BVH_MAX_DEPTH", kMaxBvhDepth)));  // (lowered by tests)
    const bool allow_deep = EnvInt("JPBRT_TEST_ALLOW_DEEP_BVH", 0) != 0;  // tests of the kernels' overflow counter only
    if (hs.bvh_depth > depth_limit && !allow_deep) {
        bld.nodes.clear();
        bld.nodes.resize(std::max<size_t>(2 * (size_t)N, 2));
        for (int i = 0; i < N; ++i) bld.idx[i] = i;
        bld.next = 0;
        bld.force_median = true;
        root = bld.Alloc();
        bld.Build(root, 0, N);
        bld.nodes.resize((size_t)bld.next.load());
        hs.bvh_depth = tree_depth(root);
        hs.bvh_builder = 2;
        if (hs.bvh_depth > kMaxBvhDepth) return fail(JPBRT_ERR_UNSUPPORTED, "BVH deeper than the traversal stack even with median splits");
    }

    phase("reinsertion / depth check");
    // flatten: inner nodes depth-first, leaves reference idx ranges (== slot ranges)
    auto& fn = hs.nodes;
    fn.clear();
    auto box_of = [&](int tmp, float mn[3], float mx[3]) {
        const Box& b = bld.nodes[tmp].box;
        for (int a = 0; a < 3; ++a) { mn[a] = b.mn[a] - pad; mx[a] = b.mx[a] + pad; }
    };
    auto emit_inner = [&]() { int i = (int)(fn.size() / kNodeStride); fn.resize(fn.size() + kNodeStride); return i; };
    const float inf = std::numeric_limits<float>::infinity();
    if (bld.nodes[root].count > 0) {
        // root is a leaf: wrap it in an inner node whose right child is an empty, never-hit box
        int f = emit_inner();
        float mn[3], mx[3];
        box_of(root, mn, mx);
        fn[f * kNodeStride + 0] = Float4{mn[0], mn[1], mn[2], mx[0]};
        fn[f * kNodeStride + 1] = Float4{mx[1], mx[2], inf, inf};
        fn[f * kNodeStride + 2] = Float4{inf, -inf, -inf, -inf};
        fn[f * kNodeStride + 3] = Float4{IntAsFloat(LeafRef(bld.nodes[root].first, bld.nodes[root].count)), IntAsFloat(LeafRef(0, 0)), 0, 0};
        hs.max_leaf_prims = bld.nodes[root].count;
    } else {
        // The order is that of a depth-first walk that hands both children of a node their indices when the node is visited
        // (siblings adjacent, the left subtree before the right): the node visited after A earlier allocations puts its inner
        // children at 1 + A (and 2 + A).  A follows from the subtree counts -- A(left) = A + c, A(right) = A + c + (inner
        // nodes below the left child), c = the node's inner children -- so subtrees are laid out independently, in parallel.
        fn.resize((size_t)inner_below[root] * kNodeStride);  // (uninitialised: the layout writes every node)
        hs.max_leaf_prims = largest_leaf.load();
        std::function<void(int, int, int, int)> layout = [&](int tmp, int flat, int A, int level) {
            for (;;) {
                const TmpNode& T = bld.nodes[tmp];
                const int kids[2] = {T.left, T.right};
                const bool inner[2] = {bld.nodes[kids[0]].count == 0, bld.nodes[kids[1]].count == 0};
                const int child_flat[2] = {1 + A, 1 + A + (inner[0] ? 1 : 0)};
                int refs[2];
                float mn[2][3], mx[2][3];
                for (int k = 0; k < 2; ++k) {
                    const TmpNode& K = bld.nodes[kids[k]];
                    box_of(kids[k], mn[k], mx[k]);
                    refs[k] = inner[k] ? child_flat[k] : LeafRef(K.first, K.count);
                }
                fn[(size_t)flat * kNodeStride + 0] = Float4{mn[0][0], mn[0][1], mn[0][2], mx[0][0]};
                fn[(size_t)flat * kNodeStride + 1] = Float4{mx[0][1], mx[0][2], mn[1][0], mn[1][1]};
                fn[(size_t)flat * kNodeStride + 2] = Float4{mn[1][2], mx[1][0], mx[1][1], mx[1][2]};
                fn[(size_t)flat * kNodeStride + 3] = Float4{IntAsFloat(refs[0]), IntAsFloat(refs[1]), 0, 0};
                const int c = (inner[0] ? 1 : 0) + (inner[1] ? 1 : 0);
                const int A_left = A + c, A_right = A + c + (inner[0] ? inner_below[kids[0]] - 1 : 0);
                if (inner[0] && inner[1]) {
                    if (level < 4 && nthreads > 1) {
                        auto fut = std::async(std::launch::async, [&layout, k = kids[0], f = child_flat[0], A_left, level]() { layout(k, f, A_left, level + 1); });
                        layout(kids[1], child_flat[1], A_right, level + 1);
                        fut.get();
                        return;
                    }
                    layout(kids[0], child_flat[0], A_left, level + 1);
                    tmp = kids[1]; flat = child_flat[1]; A = A_right; ++level;  // (tail call)
                } else if (inner[0]) {
                    tmp = kids[0]; flat = child_flat[0]; A = A_left; ++level;
                } else if (inner[1]) {
                    tmp = kids[1]; flat = child_flat[1]; A = A_right; ++level;
                } else {
                    return;
                }
            }
        };
        layout(root, 0, 0, 0);
    }
    {
        // Quantised copy of the tree (intersect.cuh JPB_QNODES): every plane of every (padded) child box as a 16-bit index on a
        // grid over the padded scene bounds, rounded OUTWARDS and then moved out by kSlack more cells -- the kernels rebuild a
        // plane as (2^23 + q) * (cell / d) + const, whose cancellation can misplace it by up to one cell.
        constexpr int kSlack = 2;
        constexpr double kCells = 65520.0;  // planes of scene geometry fall in [4, 65524]; with rounding + slack in [1, 65527]
        double org[3], cell[3];
        for (int a = 0; a < 3; ++a) {
            const double lo = (double)world.mn[a] - (double)pad, hi = (double)world.mx[a] + (double)pad;
            const float cf = (float)(std::max(hi - lo, 1e-30) / kCells);
            const float of = (float)(lo - 4.0 * (double)cf);
            hs.q_cell[a] = cf;   // the kernels use exactly these floats: the indices below are computed against them
            hs.q_origin[a] = of;
            org[a] = (double)of;
            cell[a] = (double)cf;
        }
        const size_t n_nodes = fn.size() / kNodeStride;
        hs.qnodes.clear();
        if (n_nodes <= (size_t)kQNodesMaxNodes) hs.qnodes.resize(n_nodes * kQNodeStride);
        auto q_lo = [&](float v, int a) -> unsigned {
            if (!(v > -std::numeric_limits<float>::infinity())) return 0u;
            if (v == std::numeric_limits<float>::infinity()) return 65535u;
            const double q = std::floor(((double)v - org[a]) / cell[a]) - kSlack;
            return (unsigned)std::min(65535.0, std::max(0.0, q));
        };
        auto q_hi = [&](float v, int a) -> unsigned {
            if (!(v < std::numeric_limits<float>::infinity())) return 65535u;
            if (v == -std::numeric_limits<float>::infinity()) return 0u;
            const double q = std::ceil(((double)v - org[a]) / cell[a]) + kSlack;
            return (unsigned)std::min(65535.0, std::max(0.0, q));
        };
        auto as_float = [](unsigned u) { float f; memcpy(&f, &u, 4); return f; };
        ParallelFor(hs.qnodes.empty() ? 0 : n_nodes, nthreads, [&](size_t b, size_t e, int) {
            for (size_t i = b; i < e; ++i) {
                const Float4* n = &fn[i * kNodeStride];
                // n0 = (Lmin.xyz, Lmax.x) n1 = (Lmax.yz, Rmin.xy) n2 = (Rmin.z, Rmax.xyz) n3 = (left, right, -, -)
                const unsigned lmn[3] = {q_lo(n[0].x, 0), q_lo(n[0].y, 1), q_lo(n[0].z, 2)};
                const unsigned lmx[3] = {q_hi(n[0].w, 0), q_hi(n[1].x, 1), q_hi(n[1].y, 2)};
                const unsigned rmn[3] = {q_lo(n[1].z, 0), q_lo(n[1].w, 1), q_lo(n[2].x, 2)};
                const unsigned rmx[3] = {q_hi(n[2].y, 0), q_hi(n[2].z, 1), q_hi(n[2].w, 2)};
                Float4* q = &hs.qnodes[i * kQNodeStride];
                // one word per (box, axis): min in the low half, max in the high half -- the kernel picks the ray's near and far
                // plane of an axis out of one word with a per-ray PRMT selector (no min / max of the two distances)
                q[0] = Float4{as_float(lmn[0] | (lmx[0] << 16)), as_float(lmn[1] | (lmx[1] << 16)), as_float(lmn[2] | (lmx[2] << 16)),
                              as_float(rmn[0] | (rmx[0] << 16))};
                q[1] = Float4{as_float(rmn[1] | (rmx[1] << 16)), as_float(rmn[2] | (rmx[2] << 16)), n[3].x, n[3].y};
            }
        });
    }
    hs.bvh_build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    phase("node layout");

    // slots in leaf order
    hs.slots.resize((size_t)N * kSlotStride);
    hs.slot_nrm.resize(N);
    hs.slot_ml.resize(N);
    hs.prim_slot.resize(N);
    ParallelFor((size_t)N, nthreads, [&](size_t b, size_t e, int) {
        for (size_t s = b; s < e; ++s) {
            const int p = bld.idx[s];
            for (int k = 0; k < kSlotStride; ++k) hs.slots[s * kSlotStride + k] = pre_slots[(size_t)p * kSlotStride + k];
            hs.slot_nrm[s] = pre_nrm[p];
            hs.slot_ml[s] = Int2{d->primitives[p].material, d->primitives[p].light};
            hs.prim_slot[p] = (int)s;
        }
    });
    phase("slots in leaf order");
    return 0;
}

}  // namespace jpbrt
