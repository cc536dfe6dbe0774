"""Parity of the CUDA path (through the C ABI) with the CPU checker, on identical inputs.

Checker = the compiled reference (oracle/_ref) when present, else the pinned restatement; the
committed golden vectors (outputs of the reference) are checked as well.

Bars (BASELINE.json north_star):
  * integer / index results (hit flag, primitive id, BSDF flags, occlusion): exact, except
    explicitly FLAGGED grazing cases, whose count is bounded and reported;
  * t, position, normal, camera rays, RNG: bit-exact (the kernels keep the reference's float
    expression order and are compiled with -fmad=false);
  * f, pdf, sampled directions, light samples: within REL_TOL = 1e-5 relative, ill-conditioned
    (grazing) strata flagged a priori and held to GRAZING_TOL.
"""
import zlib
from pathlib import Path

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
GRAZING_TOL = 2e-2
GOLD = np.load(Path(__file__).parent / "golden" / "ref_golden.npz")


def vec_rel(a, b):
    """Error of vector quantities relative to the vector's magnitude (rows)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.maximum(np.maximum(np.abs(a).max(axis=-1), np.abs(b).max(axis=-1)), 1e-30)
    return np.abs(a - b).max(axis=-1) / scale


def test_rng_blocks_bit_exact(pkg, port, gpu):
    rng = np.random.default_rng(1)
    n = 4096
    pix = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    smp = rng.integers(0, 2**24, n).astype(np.uint32)
    blk = rng.integers(0, 300, n).astype(np.uint32)
    seed = 0xDEADBEEF12345678
    g = pkg.unit_rng_block(pix, smp, blk, seed)
    c = np.stack([port.philox_block(int(a), int(b), int(d), seed) for a, b, d in zip(pix, smp, blk)])
    assert np.array_equal(g, c)
    assert (g >= 0).all() and (g < 1).all() and abs(g.mean() - 0.5) < 0.01


def test_philox_known_answer_vectors(pkg, gpu):
    """The CUDA generator against the published Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    ctr = np.array([k[0] for k in common.PHILOX_KAT], np.uint32)
    key = np.array([k[1] for k in common.PHILOX_KAT], np.uint32)
    want = np.array([k[2] for k in common.PHILOX_KAT], np.uint32)
    assert np.array_equal(pkg.unit_philox_raw(ctr, key), want)
    blk = pkg.unit_rng_block([0], [0], [0], 0)[0]
    assert np.array_equal(blk, (want[0] >> 8).astype(np.float32) / np.float32(1 << 24))


@pytest.mark.parametrize("name", ["tri", "tri_flip_big", "rect", "rect_xz_flip", "sphere", "disk"])
def test_shape_intersect_bit_exact(pkg, checker, gpu, name):
    sh = common.shapes(pkg)[name]
    # golden vectors of the reference
    hit, t, pos, nrm = pkg.unit_intersect_shape(sh, GOLD[f"shape_{name}_rays"])
    assert np.array_equal(hit, GOLD[f"shape_{name}_hit"])
    assert np.array_equal(t, GOLD[f"shape_{name}_t"]) and np.array_equal(pos, GOLD[f"shape_{name}_pos"])
    assert np.array_equal(nrm, GOLD[f"shape_{name}_nrm"])
    # a large seeded set against the live checker
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    pts = np.array([[sh.p[i][k] for k in range(3)] for i in range(4)], np.float64)
    c = pts[0] if sh.type >= 2 else pts[:3 + (sh.type == 1)].mean(0)
    r = float(sh.p[1][0]) if sh.type == 2 else float(sh.p[2][0]) if sh.type == 3 else np.linalg.norm(pts[:3] - c, axis=1).max()
    rays = common.shape_rays(rng, 1 << 18, c, r)
    rays[::7, 6] = 0.5 * r   # vary tmin / tmax so the range test is exercised
    rays[::5, 7] = 3.0 * r
    g = pkg.unit_intersect_shape(sh, rays)
    k = checker.intersect_shape(sh, rays)
    assert 0.05 < g[0].mean() < 0.95
    for a, b in zip(g, k):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["matte", "mirror", "glass", "plastic", "plastic_remap", "metal", "metal_aniso_remap"])
def test_bsdf_parity(pkg, checker, gpu, name):
    m = common.materials(pkg)[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    for source in ("golden", "live"):
        if source == "golden":
            i = [GOLD[f"bsdf_{name}_in_{k}"] for k in ("nrm", "wo", "wi", "u2", "ul")]
            want = {k: GOLD[f"bsdf_{name}_{k}"] for k in ("f_eval", "pdf_eval", "s_wi", "s_f", "s_pdf", "s_flags", "is_delta")}
        else:
            i = common.bsdf_inputs(rng, 1 << 18)
            want = checker.bsdf(m, *i)
        got = pkg.unit_bsdf(m, *i)
        nrm, wo = i[0].astype(np.float64), i[1].astype(np.float64)
        assert np.array_equal(got["is_delta"], want["is_delta"])
        flags_differ = got["s_flags"] != want["s_flags"]
        assert flags_differ.mean() <= 1e-5, f"{flags_differ.sum()} sampled lobes differ"
        ok = ~flags_differ
        # a-priori conditioning: cos of the sampled / evaluated directions against the normal
        cos_s = np.abs((want["s_wi"].astype(np.float64) * nrm).sum(1))
        cos_o = np.abs((wo * nrm).sum(1))
        cos_i = np.abs((i[2].astype(np.float64) * nrm).sum(1))
        # cosine-hemisphere sampling computes z = sqrt(1 - x^2 - y^2): a 1-ulp difference between the
        # host's and CUDA's cosf/sinf is amplified by 1 / (2 z^2), so z < 0.2 is the flagged stratum
        graz_s = (cos_s < 0.2) | (cos_o < 0.05)
        graz_e = (cos_i < 0.05) | (cos_o < 0.05)
        checks = [("f_eval", vec_rel(got["f_eval"], want["f_eval"]), graz_e),
                  ("pdf_eval", common.rel_err(got["pdf_eval"], want["pdf_eval"]), graz_e),
                  ("s_wi", vec_rel(got["s_wi"], want["s_wi"]), graz_s),
                  ("s_f", vec_rel(got["s_f"], want["s_f"]), graz_s),
                  ("s_pdf", common.rel_err(got["s_pdf"], want["s_pdf"]), graz_s)]
        for key, err, graz in checks:
            assert np.isfinite(err[ok]).all() or not np.isfinite(want[key]).all(), key
            e = np.nan_to_num(err, nan=0.0)
            assert e[ok & ~graz].max(initial=0) <= REL_TOL, (name, source, key, float(e[ok & ~graz].max()))
            assert e[ok & graz].max(initial=0) <= GRAZING_TOL, (name, source, key, "grazing", float(e[ok & graz].max()))


SCENES = [("cornell", 1.0), ("bunny", 1.0), ("glossy", 1.0), ("large", 0.3)]


@pytest.fixture(scope="module", params=SCENES, ids=[s[0] for s in SCENES])
def scene_pair(request, pkg, checker, port, gpu):
    name, scale = request.param
    sc = pkg.HostScene.builtin(name, 256, 256, scale)
    ctx = pkg.Context(sc)
    yield name, sc, ctx, checker.scene(sc), port.scene(sc)
    ctx.close()


def edge_grazing(sc, prim, rays):
    """A-priori test, in float64, that a ray passes a triangle/rectangle within the ROUNDING-ERROR band of
    one of its edges: some edge function v_i . d (shape.h:300-306) is smaller than the worst-case float32
    error of its own evaluation.  There the reference's sign test is decided by rounding noise."""
    out = np.zeros(len(prim), bool)
    eps = 2.0 ** -24
    for j, (p, r) in enumerate(zip(prim, rays.astype(np.float64))):
        if p < 0:
            continue
        sh = sc.d.shapes[sc.d.primitives[int(p)].shape]
        if sh.type > 1:
            continue
        k = 3 if sh.type == 0 else 4
        v = np.array([[sh.p[i][c] for c in range(3)] for i in range(k)], np.float64) - r[:3]
        d = r[3:6]
        pairs = [(2, 1), (1, 0), (0, 2)] if k == 3 else [(2, 1), (1, 0), (0, 3), (3, 2)]
        for a, b in pairs:
            val = np.dot(np.cross(v[a], v[b]), d)
            bound = 16 * eps * (np.abs(np.outer(v[a], v[b])).sum() * np.abs(d).max())
            if abs(val) <= bound:
                out[j] = True
    return out


def check_hits(name, sc, tag, ctx, ks, ps, rays, report):
    """prim id exact except FLAGGED grazing cases; t / position / normal bit-exact wherever the primitive agrees.

    A mismatch must be one of (DESIGN.md "Flagged cases"):
      tie    -- both found a hit at (relatively) equal t on different primitives (shared edge / coplanar);
      refbox -- the GPU's answer equals the closest hit over ALL primitives without any boxes: the
                reference's unpadded BVH boxes culled a hit its own primitive test accepts;
      edge   -- the primitive one side reports and the other does not is hit within the float32 rounding
                band of one of its edges: the reference's sign test is rounding noise there, and whether
                such a hit is reached depends on BVH topology (its own random split axes, bvh.h:61).
    """
    pg, tg, posg, ng = ctx.unit_scene_intersect(rays)
    pk, tk, posk, nk = ks.intersect(rays)
    same = pg == pk
    assert np.array_equal(tg[same], tk[same]) and np.array_equal(ng[same], nk[same]) and np.array_equal(posg[same], posk[same])
    mism = np.flatnonzero(~same)
    kinds = {"tie": 0, "refbox": 0, "edge": 0, "UNEXPLAINED": 0}
    if len(mism):
        pb, tb, _, _ = ps.intersect_brute(rays[mism])
        eg = edge_grazing(sc, pg[mism], rays[mism]) | edge_grazing(sc, pk[mism], rays[mism])
        for j, i in enumerate(mism):
            if pg[i] >= 0 and pk[i] >= 0 and abs(tg[i] - tk[i]) <= 1e-5 * max(abs(tg[i]), abs(tk[i])):
                kinds["tie"] += 1
            elif pg[i] == pb[j] and tg[i] == tb[j]:
                kinds["refbox"] += 1
            elif eg[j]:
                kinds["edge"] += 1
            else:
                kinds["UNEXPLAINED"] += 1
    report.append((name, tag, len(rays), kinds))
    assert kinds["UNEXPLAINED"] == 0, (name, tag, kinds)
    assert len(mism) / len(rays) <= (2e-2 if name == "large" else 2e-4), (name, tag, kinds)
    return pk, posk, nk


def test_scene_closest_hit_and_occlusion(scene_pair, pkg):
    name, sc, ctx, ks, ps = scene_pair
    rng = np.random.default_rng(0xC0FFEE)
    report = []
    assert np.array_equal(ctx.scene_info(), ks.info())
    # golden camera rays of the reference (only at the golden configuration's scale)
    raysA, pf = common.camera_rays(ks, rng, 1 << 17, 256, 256)
    o, d = ctx.unit_generate_rays(pf)
    assert np.array_equal(np.concatenate([o, d], 1), raysA[:, :6]), "generate: camera rays differ"
    check_hits(name, sc, "A:camera", ctx, ks, ps, raysA, report)
    raysB, P, N = common.secondary_rays(ks, raysA, rng)
    check_hits(name, sc, "B:surface", ctx, ks, ps, raysB, report)
    raysC = common.bbox_rays(ks.info(), rng, 1 << 17)
    check_hits(name, sc, "C:bbox", ctx, ks, ps, raysC, report)
    # Occluded(): targets at random distances along the secondary rays, plus light sample points
    tgt = (P + raysB[:, 3:6] * rng.uniform(0.5, 900, (len(P), 1))).astype(np.float32)
    og, ok = ctx.unit_scene_occluded(P, tgt), ks.occluded(P, tgt)
    report.append((name, "occluded", len(P), int((og != ok).sum())))
    assert (og != ok).mean() <= (2e-2 if name == "large" else 2e-4), int((og != ok).sum())
    print("flagged grazing cases:", report)


def test_scene_golden_hits(pkg, gpu):
    for name in ("cornell", "bunny", "glossy", "large"):
        scale, w, h, spp = GOLD[f"scene_{name}_cfg"]
        sc = pkg.HostScene.builtin(name, int(w), int(h), float(scale))
        ctx = pkg.Context(sc)
        g = lambda k: GOLD[f"scene_{name}_{k}"]  # noqa: E731
        prim, t, pos, nrm = ctx.unit_scene_intersect(g("rays"))
        same = prim == g("prim")
        assert same.mean() >= 0.999 and np.array_equal(t[same], g("t")[same]) and np.array_equal(nrm[same], g("nrm")[same])
        prim2, t2, _, nrm2 = ctx.unit_scene_intersect(g("rays2"))
        same2 = prim2 == g("prim2")
        assert same2.mean() >= 0.999 and np.array_equal(t2[same2], g("t2")[same2])
        occ = ctx.unit_scene_occluded(g("rays2")[:, :3], g("occ_tgt"))
        assert (occ != g("occ")).mean() <= 1e-3
        assert np.array_equal(ctx.unit_emitted(g("prim"), g("nrm"), -g("rays")[:, 3:6]), g("Le"))
        ctx.close()


def test_light_sampling_and_emission(scene_pair, pkg):
    name, sc, ctx, ks, ps = scene_pair
    rng = np.random.default_rng(7)
    rays, _ = common.camera_rays(ks, rng, 1 << 15, 256, 256)
    prim, t, pos, nrm = ks.intersect(rays)
    m = prim >= 0
    P, N = pos[m], nrm[m]
    for li in range(sc.d.n_lights):
        u2 = rng.uniform(0, 1, (len(P), 2)).astype(np.float32)
        lpos, wi, pdf, Li = ctx.unit_light_sample(li, P, N, u2)
        kpos, kwi, kpdf, kLi = ks.light_sample(li, P, N, u2)
        assert np.array_equal(Li, kLi), (name, li)
        assert vec_rel(lpos, kpos).max() <= REL_TOL and vec_rel(wi, kwi).max() <= REL_TOL
        assert common.rel_err(pdf, kpdf).max() <= 2e-5, (name, li, float(common.rel_err(pdf, kpdf).max()))
    wo = -rays[:, 3:6]
    assert np.array_equal(ctx.unit_emitted(prim, nrm, wo), ks.emitted(prim, nrm, wo))


def test_sphere_and_delta_lights(pkg, checker, gpu):
    """Light kinds the built-in scenes do not use: sphere area light (cone + inside sampling), disk, point, direction."""
    cam = pkg.Camera((0, 2, 12), (0, -0.1, -1), (0, 1, 0), 60.0, 64, 64)
    S, M, L, Pm = pkg.Shape, pkg.Material, pkg.Light, pkg.Primitive
    z3 = (0.0, 0.0, 0.0)
    shapes = [S(pkg.SHAPE_SPHERE, 0, ((0, 4, 0), (1.5, 0, 0), z3, z3)),
              S(pkg.SHAPE_DISK, 0, ((3, 3, 1), (0.2, -1, 0.1), (1.2, 0, 0), z3)),
              S(pkg.SHAPE_RECTANGLE, 0, ((-8, 0, -8), (-8, 0, 8), (8, 0, 8), (8, 0, -8))),
              S(pkg.SHAPE_SPHERE, 0, ((-2, 1, 1), (1.0, 0, 0), z3, z3))]
    mats = [M(pkg.MAT_MATTE, 0, (.6, .6, .6), z3, 0, 0), M(pkg.MAT_MIRROR, 0, (.9, .9, .9), z3, 0, 0)]
    lights = [L(pkg.LIGHT_ENVIRONMENT, -1, (.05, .05, .1), z3, z3), L(pkg.LIGHT_AREA, 0, (10, 9, 8), z3, z3),
              L(pkg.LIGHT_AREA, 1, (5, 5, 5), z3, z3), L(pkg.LIGHT_POINT, -1, (30, 30, 30), (4, 5, -2), z3),
              L(pkg.LIGHT_DIRECTION, -1, (1, 1, 1), z3, (0.3, -1, 0.2))]
    prims = [Pm(0, 0, 1), Pm(1, 0, 2), Pm(2, 0, -1), Pm(3, 1, -1)]
    sc = pkg.HostScene.from_arrays(cam, shapes, mats, lights, prims, max_depth=4, name="lights")
    ctx, ks = pkg.Context(sc), checker.scene(sc)
    rng = np.random.default_rng(3)
    n = 1 << 15
    P = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    P[: n // 8] = (np.array([0, 4, 0]) + common.unit_vectors(rng, n // 8) * rng.uniform(0, 1.5, (n // 8, 1))).astype(np.float32)  # inside the sphere light
    N = common.unit_vectors(rng, n)
    for li in range(5):
        u2 = rng.uniform(0, 1, (n, 2)).astype(np.float32)
        g, k = ctx.unit_light_sample(li, P, N, u2), ks.light_sample(li, P, N, u2)
        assert np.array_equal(g[3], k[3]), li
        assert vec_rel(g[0], k[0]).max() <= REL_TOL and vec_rel(g[1], k[1]).max() <= 5 * REL_TOL, li
        e = np.nan_to_num(common.rel_err(g[2], k[2]))
        assert np.quantile(e, 0.999) <= 1e-4 and e.max() <= GRAZING_TOL, (li, float(e.max()))
    rays, _ = common.camera_rays(ks, rng, 1 << 15, 64, 64)
    for a, b in zip(ctx.unit_scene_intersect(rays), ks.intersect(rays)):
        assert np.array_equal(a, b)
    ctx.close()


def test_full_size_five_million_triangle_scene_hits(pkg, checker, port, gpu):
    """BASELINE config C3 at FULL size (4,999,001 triangles): closest hits and occlusion against the checker
    (its BVH is the reference's own median-split tree, built here in ~12 s)."""
    sc = pkg.HostScene.builtin("large", 256, 256, 1.0)
    ctx, ks, ps = pkg.Context(sc), checker.scene(sc), port.scene(sc)
    rng = np.random.default_rng(31337)
    report = []
    raysA, _ = common.camera_rays(ks, rng, 1 << 15, 256, 256)
    check_hits("large", sc, "A:camera", ctx, ks, ps, raysA, report)
    raysB, P, N = common.secondary_rays(ks, raysA, rng)
    check_hits("large", sc, "B:surface", ctx, ks, ps, raysB, report)
    tgt = (P + raysB[:, 3:6] * rng.uniform(0.5, 1500, (len(P), 1))).astype(np.float32)
    og, ok = ctx.unit_scene_occluded(P, tgt), ks.occluded(P, tgt)
    report.append(("large-full", "occluded", len(P), int((og != ok).sum())))
    assert (og != ok).mean() <= 2e-2
    print("flagged grazing cases (full-size C3):", report)
    ctx.close()


@pytest.mark.parametrize("name", ["phong", "phong_soft", "refl_beckmann_visible_dielectric", "refl_beckmann_visible_aniso_conductor",
                                  "refl_beckmann_full", "refl_beckmann_full_aniso", "refl_tr_full", "refl_tr_full_aniso", "trans_tr",
                                  "trans_beckmann_aniso"])
def test_unbuilt_bsdf_classes_parity(pkg, checker, gpu, name):
    """jpbrt_unit_bsdf_ex against the reference's classes no material builds.  Evalf / Pdf: 1e-5
    relative everywhere.  Sampled values inherit the last-ulp differences of expf / logf / powf / acosf / tanf between CUDA
    and glibc through exp(-tan^2/alpha^2), pow(., exponent) and BeckmannSample11's Newton iteration (which stops on
    |value| < 1e-5): the direction stays within 1e-4, f and pdf within 1e-5 for >= 99 % of the samples, 2e-3 for 99.99 % and 2e-2 for all
    (measured on B200: worst single sample of 2^17, 5.7e-3)."""
    d = common.bsdf_ex_cases(pkg)[name]
    g = np.load(Path(__file__).parent / "golden" / "ref_golden_bsdf_ex.npz")
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 7)
    for source in ("golden", "live"):
        if source == "golden":
            i = [g[k] for k in ("nrm", "wo", "wi", "u2")]
            want = {k: g[f"{name}_{k}"] for k in ("f_eval", "pdf_eval", "s_wi", "s_f", "s_pdf", "s_flags")}
        else:
            i = common.bsdf_inputs(rng, 1 << 17)[:4]
            want = checker.bsdf_ex(d, *i)
        got = pkg.unit_bsdf_ex(d, *i)
        assert (got["s_flags"] != want["s_flags"]).mean() <= 1e-5
        ok = got["s_flags"] == want["s_flags"]
        for key in ("f_eval", "pdf_eval", "s_wi", "s_f", "s_pdf"):
            assert np.array_equal(np.isnan(got[key]), np.isnan(want[key])), (name, key)
            e = vec_rel(got[key], want[key]) if got[key].ndim == 2 else common.rel_err(got[key], want[key])
            e = np.nan_to_num(e, nan=0.0)[ok]
            if key in ("f_eval", "pdf_eval"):
                assert e.max(initial=0) <= REL_TOL, (name, source, key, float(e.max()))
            elif key == "s_wi":
                assert e.max(initial=0) <= 1e-4 and np.quantile(e, 0.999) <= REL_TOL, (name, source, key, float(e.max()))
            else:
                assert e.max(initial=0) <= 2e-2 and np.quantile(e, 0.9999) <= 2e-3 and np.quantile(e, 0.99) <= REL_TOL, \
                    (name, source, key, float(e.max()), float(np.quantile(e, 0.9999)), float(np.quantile(e, 0.99)))
