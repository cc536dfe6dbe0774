"""Option "shade_math" = 1: the shade stage (k_logic, k_shade<KIND>) built with FMA contraction and reciprocal-multiply
division (csrc/shade_fast.cu) -- an OPTION, not the default, and this file says exactly how far it is from the reference.

Measured on B200 (profiles/ab/r02_ab_shade_math.log): shade stage -8 % (bunny) / -12 % (Cornell) / -27 % (glossy), frames +2.4 /
+5 / +12 %.  Accuracy against the reference's CPU functions on 2^18 seeded inputs per material: Lambert and the delta lobes
stay within 3e-6; the microfacet / Fresnel expressions are cancellation-prone (1 - cos^2, eta^2 (1 - cos^2), tan^2), and
there a changed rounding moves f / pdf by up to 3e-4 relative -- >= 99 % of the samples meet north_star's 1e-5 (plastic:
99.9th percentile 1.2e-5), the tail does not.  That is why the DEFAULT build keeps the reference's expression order bit for bit (tests/test_gpu_parity.py)
and this build is offered for throughput where 1e-3 per-value agreement is enough.  Images: statistically
indistinguishable from the reference (same bars as the exact build); pixel-for-pixel within 1e-3.
Intersections are untouched: which primitive a ray hits cannot depend on the option."""
import zlib

import numpy as np
import pytest

import common
from test_gpu_parity import GRAZING_TOL, REL_TOL, vec_rel

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fast_default(pkg):
    pkg.set_default_option("shade_math", 1)
    yield
    pkg.set_default_option("shade_math", 0)


@pytest.mark.parametrize("name", ["matte", "mirror", "glass", "plastic", "plastic_remap", "metal", "metal_aniso_remap"])
def test_fast_bsdf_error_against_the_reference(pkg, checker, gpu, fast_default, name):
    m = common.materials(pkg)[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 11)
    i = common.bsdf_inputs(rng, 1 << 18)
    want = checker.bsdf(m, *i)
    got = pkg.unit_bsdf(m, *i)
    nrm, wo = i[0].astype(np.float64), i[1].astype(np.float64)
    assert np.array_equal(got["is_delta"], want["is_delta"])
    flags_differ = got["s_flags"] != want["s_flags"]   # a discrete decision (u < F, hemisphere test) sitting within rounding of its threshold
    assert flags_differ.mean() <= 2e-5, f"{flags_differ.sum()} sampled lobes differ"
    ok = ~flags_differ
    cos_s = np.abs((want["s_wi"].astype(np.float64) * nrm).sum(1))
    cos_o = np.abs((wo * nrm).sum(1))
    cos_i = np.abs((i[2].astype(np.float64) * nrm).sum(1))
    graz_s = (cos_s < 0.2) | (cos_o < 0.05)
    graz_e = (cos_i < 0.05) | (cos_o < 0.05)
    worst = {}
    for key, err, graz in [("f_eval", vec_rel(got["f_eval"], want["f_eval"]), graz_e), ("pdf_eval", common.rel_err(got["pdf_eval"], want["pdf_eval"]), graz_e),
                           ("s_wi", vec_rel(got["s_wi"], want["s_wi"]), graz_s), ("s_f", vec_rel(got["s_f"], want["s_f"]), graz_s),
                           ("s_pdf", common.rel_err(got["s_pdf"], want["s_pdf"]), graz_s)]:
        e = np.nan_to_num(err, nan=0.0)
        sel = e[ok & ~graz]
        worst[key] = (float(sel.max(initial=0)), float((sel > REL_TOL).mean()) if len(sel) else 0.0)
        q99, q999 = (float(np.quantile(sel, 0.99)), float(np.quantile(sel, 0.999))) if len(sel) else (0.0, 0.0)
        worst[key] += (q999,)
        assert q99 <= REL_TOL, (name, key, q99)            # 99 % within north_star's 1e-5 ...
        assert q999 <= 1e-4, (name, key, q999)             # ... 99.9 % within 1e-4 ...
        assert sel.max(initial=0) <= 2e-3, (name, key, worst[key])  # ... and the cancellation-prone tail within 2e-3
        assert e[ok & graz].max(initial=0) <= 2 * GRAZING_TOL, (name, key, "grazing", float(e[ok & graz].max()))
    print(name, "fast-math (worst relative error, fraction beyond 1e-5, 99.9th percentile) outside the grazing strata:", worst)


@pytest.mark.parametrize("name,scale", [("cornell", 1.0), ("bunny", 1.0), ("glossy", 1.0)])
def test_fast_light_sampling_within_tolerance(pkg, checker, gpu, name, scale):
    sc = pkg.HostScene.builtin(name, 128, 128, scale)
    ctx, ks = pkg.Context(sc), checker.scene(sc)
    ctx.set_option("shade_math", 1)
    rng = np.random.default_rng(17)
    rays, _ = common.camera_rays(ks, rng, 1 << 14, 128, 128)
    prim, t, pos, nrm = ks.intersect(rays)
    P, N = pos[prim >= 0], nrm[prim >= 0]
    for li in range(sc.d.n_lights):
        u2 = rng.uniform(0, 1, (len(P), 2)).astype(np.float32)
        lpos, wi, pdf, Li = ctx.unit_light_sample(li, P, N, u2)
        kpos, kwi, kpdf, kLi = ks.light_sample(li, P, N, u2)
        lit = (Li != 0).any(axis=1) == (kLi != 0).any(axis=1)   # one-sided emission: n_l . (-wi) > 0 sits on rounding at grazing angles
        assert (~lit).mean() <= 1e-3, (name, li, float((~lit).mean()))
        assert vec_rel(Li[lit], kLi[lit]).max(initial=0) <= REL_TOL
        assert vec_rel(lpos, kpos).max() <= REL_TOL and vec_rel(wi, kwi).max() <= REL_TOL
        e = common.rel_err(pdf[lit], kpdf[lit])
        assert np.quantile(e, 0.999) <= 3e-5 and e.max() <= 2e-3, (name, li, float(e.max()))
    ctx.close()


@pytest.mark.parametrize("name,scale,res,spp", [("cornell", 1.0, 160, 3), ("bunny", 0.5, 160, 3), ("glossy", 1.0, 96, 2)])
def test_fast_same_path_images(pkg, port, gpu, name, scale, res, spp):
    sc = pkg.HostScene.builtin(name, res, res, scale)
    ctx = pkg.Context(sc)
    ctx.render_pass(0, spp, seed=2024)
    exact = ctx.read_film(finalize=False)
    st_exact = ctx.stats()
    ctx.clear_film(); ctx.reset_stats()
    ctx.set_option("shade_math", 1)
    ctx.render_pass(0, spp, seed=2024)
    g = ctx.read_film(finalize=False)
    st = ctx.stats()
    ctx.close()
    c, _, cnt = port.scene(sc).render_counter(0, spp, 2024, numthreads=16, counters=True)
    assert np.isfinite(g).all() and st["invalid_contributions"] == 0 and st["samples"] == res * res * spp
    # relaxed arithmetic moves microfacet values by up to ~1e-4: pixel for pixel the image agrees to 1e-3, not to the exact
    # build's 1e-4; only a discrete decision within rounding of its threshold changes a path
    bad = (np.abs(g - c) > 1e-3 * np.maximum(np.abs(c), 1.0)).any(axis=2)
    assert bad.mean() <= 1.5e-2, f"{name}: {bad.mean():.4%} of pixels differ from the same-path oracle by more than 1e-3"
    assert abs(g.mean() - c.mean()) <= 2e-3 * c.mean()
    assert abs(st["shaded_vertices"] - cnt["vertices"]) <= 3e-3 * cnt["vertices"]
    assert abs(st["shadow_rays"] - cnt["shadow_rays"]) <= 3e-3 * cnt["shadow_rays"]
    bad_vs_exact = (np.abs(g - exact) > 1e-4 * np.maximum(np.abs(exact), 1.0)).any(axis=2)
    print(name, "fast vs oracle pixels off", float(bad.mean()), "| fast vs exact build pixels off", float(bad_vs_exact.mean()),
          "| vertices", st["shaded_vertices"], st_exact["shaded_vertices"])


def test_fast_statistical_parity_with_reference_sampler(pkg, checker, gpu):
    res, spp = 128, 16
    sc = pkg.HostScene.builtin("bunny", res, res, 0.5)
    ks = checker.scene(sc)
    cpu = [ks.render(spp, 16, seed=s)[0] for s in (1234, 4321, 777, 31337)]
    ctx = pkg.Context(sc)
    ctx.set_option("shade_math", 1)
    ctx.render_pass(0, spp, seed=5)
    g = ctx.read_film(spp_total=spp)
    ctx.close()
    relmse = lambda a, b: float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))  # noqa: E731
    cc = float(np.mean([relmse(cpu[i], cpu[j]) for i in range(4) for j in range(i + 1, 4)]))
    gc = float(np.mean([relmse(g, c) for c in cpu]))
    assert gc <= 1.25 * cc, (gc, cc)
    for ch in range(3):
        m = np.array([c[..., ch].mean() for c in cpu], np.float64)
        assert abs(g[..., ch].mean() - m.mean()) <= 0.005 * m.mean() + 4 * m.std(ddof=1)
