"""the reference's other integrators as modes of the wavefront pipeline (option "integrator").

FPathIntegratorRecursive (integrator.cc:233-307), FWhittedIntegrator (integrator.cc:115-220) and FDebugIntegrator
(integrator.h:44-58), all through the C ABI, against
  * the CPU restatement driven by the SAME counter-based sampler (same rays, pixel by pixel) -- the restatement itself
    is pinned bit-for-bit to the compiled reference in tests/test_oracle_pin.py, and
  * the reference's own FRandomSampler render at equal spp (statistical bar of SURVEY.md 8d).
"""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def relmse(g, c):
    return float(np.mean((g - c) ** 2 / (c ** 2 + 1e-2)))


def scenes(pkg):
    return {"specular": lambda: common.specular_scene(pkg, 128), "specular_d2": lambda: common.specular_scene(pkg, 96, max_depth=2),
            "bunny": lambda: pkg.HostScene.builtin("bunny", 128, 128, 0.5), "cornell": lambda: pkg.HostScene.builtin("cornell", 128, 128)}


@pytest.mark.parametrize("mode", ["path_recursive", "whitted", "debug"])
@pytest.mark.parametrize("scene", ["specular", "specular_d2", "bunny", "cornell"])
def test_same_rays_as_counter_oracle(pkg, port, orc_mod, gpu, mode, scene):
    sc = scenes(pkg)[scene]()
    spp = 3
    ctx = pkg.Context(sc)
    ctx.set_option("integrator", pkg.INTEGRATORS[mode])
    ctx.render_pass(0, spp, seed=99)
    g = ctx.read_film(finalize=False)
    st = ctx.stats()
    c, _, cnt = port.scene(sc).render_counter(0, spp, 99, numthreads=16, counters=True, mode=orc_mod.INTEGRATORS[mode])
    assert np.isfinite(g).all() and st["invalid_contributions"] == 0
    tol = 1e-4 if mode != "debug" else 2e-5
    bad = (np.abs(g - c) > tol * np.maximum(np.abs(c), 1.0)).any(axis=2)
    assert bad.mean() <= 1e-2, f"{mode}/{scene}: {bad.mean():.4%} of pixels differ from the same-ray oracle"
    assert abs(g.mean() - c.mean()) <= 2e-3 * abs(c.mean()) + 1e-6
    assert np.abs(c).max() > 0
    if mode == "whitted":  # the ray TREE is the same: every vertex and every shadow ray of the recursion is there
        assert abs(st["shaded_vertices"] - cnt["vertices"]) <= 2e-3 * cnt["vertices"]
        assert abs(st["shadow_rays"] - cnt["shadow_rays"]) <= 2e-3 * cnt["shadow_rays"]
        assert abs(st["extension_rays"] - cnt["ext_rays"]) <= 2e-3 * cnt["ext_rays"]
    if mode == "debug":
        assert st["extension_rays"] == cnt["ext_rays"] == sc.d.camera.width * sc.d.camera.height * spp and st["shadow_rays"] == 0
    ctx.close()


def test_whitted_mirror_is_traced_twice(pkg, port, orc_mod, gpu):
    """Reference behaviour kept: Specular|Reflection matches SpecularReflect AND SpecularReflectAndTransmit (bsdf.h:282)."""
    sc = common.specular_scene(pkg, 96)
    ctx = pkg.Context(sc)
    ctx.set_option("integrator", pkg.INTEGRATORS["whitted"])
    ctx.render_pass(0, 2, seed=5)
    st = ctx.stats()
    _, _, cnt = port.scene(sc).render_counter(0, 2, 5, numthreads=8, counters=True, mode=2)
    assert st["extension_rays"] == cnt["ext_rays"] > 1.2 * st["samples"]
    ctx.close()


@pytest.mark.parametrize("mode,spp", [("whitted", 16), ("path_recursive", 16), ("debug", 4)])
def test_statistical_parity_with_reference_sampler(pkg, checker, orc_mod, gpu, mode, spp):
    sc = common.specular_scene(pkg, 112)
    ks = checker.scene(sc)
    m = orc_mod.INTEGRATORS[mode]
    cpu = [ks.render(spp, 16, seed=s, mode=m)[0] for s in (1234, 4321, 777, 31337)]
    hi, _ = ks.render(spp * 8, 16, seed=999, mode=m)
    g, _ = pkg.render(sc, spp, seed=5, integrator=mode)
    assert np.isfinite(g).all() and g.min() >= 0 and g.max() <= 1
    cpu_vs_hi = float(np.mean([relmse(c, hi) for c in cpu]))
    gpu_vs_hi = relmse(g, hi)
    assert gpu_vs_hi <= 1.25 * cpu_vs_hi + 1e-7, (mode, gpu_vs_hi, cpu_vs_hi)
    for ch in range(3):
        mm = np.array([c[..., ch].mean() for c in cpu], np.float64)
        assert abs(g[..., ch].mean() - mm.mean()) <= 0.005 * mm.mean() + 4 * mm.std(ddof=1), (mode, ch)
    print(mode, "relMSE vs 8x-spp reference image: gpu", gpu_vs_hi, "cpu", cpu_vs_hi)


def test_modes_are_pass_split_invariant_and_switchable(pkg, gpu):
    sc = common.specular_scene(pkg, 96)
    ctx = pkg.Context(sc)
    films = {}
    for mode in ("path", "whitted", "debug", "path_recursive"):
        ctx.set_option("integrator", pkg.INTEGRATORS[mode])
        ctx.clear_film()
        ctx.render_pass(0, 6, seed=3)
        whole = ctx.read_film(finalize=False)
        ctx.clear_film()
        ctx.set_option("paths_in_flight", 96 * 96 * 2)
        ctx.render_pass(0, 2, seed=3)
        ctx.render_pass(2, 4, seed=3)
        parts = ctx.read_film(finalize=False)
        ctx.set_option("paths_in_flight", 0)
        np.testing.assert_allclose(parts, whole, rtol=3e-5, atol=2e-6)
        films[mode] = whole
    np.testing.assert_allclose(films["path"], films["path_recursive"], rtol=3e-5, atol=2e-6)  # same estimator, same draws
    assert not np.allclose(films["path"], films["whitted"]) and not np.allclose(films["whitted"], films["debug"])
    assert ctx.stats()["invalid_contributions"] == 0
    with pytest.raises(pkg.JpbrtError):
        ctx.set_option("integrator", 7)
    ctx.close()


def test_whitted_tree_overflow_is_counted_not_fatal(pkg, gpu):
    """A pool with room for one sample per pixel and a mirror-mirror scene: rays beyond the pool are dropped and counted."""
    sc = common.specular_scene(pkg, 64, max_depth=9)
    ctx = pkg.Context(sc)
    ctx.set_option("integrator", pkg.INTEGRATORS["whitted"])
    ctx.render_pass(0, 2, seed=1)
    full = ctx.read_film(finalize=False)
    assert ctx.stats()["invalid_contributions"] == 0
    ctx.clear_film(); ctx.reset_stats()
    ctx.set_option("paths_in_flight", 64 * 64)
    ctx.render_pass(0, 2, seed=1)
    small = ctx.read_film(finalize=False)
    st = ctx.stats()
    assert np.isfinite(small).all()
    if st["invalid_contributions"] == 0:
        np.testing.assert_allclose(small, full, rtol=3e-5, atol=2e-6)
    else:
        assert (small <= full + 1e-4 * np.maximum(full, 1)).all()  # only non-negative contributions can be missing
    ctx.close()


def test_whitted_refuses_depths_whose_ray_tree_ids_would_wrap(pkg, gpu):
    """The Whitted mode numbers ray-tree vertices in 32 bits (3n+1 / 3n+3); beyond depth 20 the ids wrap and sampler streams
    would correlate.  The pass is refused with a message instead of rendering a subtly wrong image; the path
    integrator, which has no ray tree, takes the same scene at the same depth."""
    sc = common.specular_scene(pkg, 32, max_depth=24)
    ctx = pkg.Context(sc)
    ctx.set_option("integrator", pkg.INTEGRATORS["whitted"])
    with pytest.raises(pkg.JpbrtError, match="max_depth <= 20"):
        ctx.render_pass(0, 1, seed=1)
    ctx.set_option("integrator", pkg.INTEGRATORS["path"])
    ctx.render_pass(0, 1, seed=1)
    assert np.isfinite(ctx.read_film(finalize=False)).all() and ctx.stats()["invalid_contributions"] == 0
    ctx.close()
