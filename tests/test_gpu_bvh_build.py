"""the BVH built ON THE GPU (linear BVH, csrc/bvh_build.cuh) instead of the host's binned-SAH build.

The tree only prunes: every (ray, primitive) test is the reference's arithmetic, so a device-built tree must give the
same hits, the same occlusion and the same images as the host-built one (and as the CPU checker), and it must be a
valid partition of the primitives whose boxes nest.
"""
import numpy as np
import pytest

import common
from test_gpu_parity import check_hits
from test_host_bvh import check_tree

pytestmark = pytest.mark.gpu


def build(pkg, name):
    if name == "specular":
        return common.specular_scene(pkg, 96)
    scale = {"bunny": 0.5, "large": 0.05}.get(name, 1.0)
    return pkg.HostScene.builtin(name, 96, 96, scale)


@pytest.mark.parametrize("name", ["cornell", "bunny", "glossy", "large", "specular"])
def test_device_built_tree_is_valid_and_gives_the_same_hits(pkg, checker, port, gpu, name):
    sc = build(pkg, name)
    host, dev = pkg.Context(sc, gpu_bvh=False), pkg.Context(sc, gpu_bvh=True)
    assert host.stats()["bvh_builder"] == 0 and dev.stats()["bvh_builder"] == 1
    nodes = dev.table("nodes").reshape(-1, 16)
    refs = nodes[:, 12:14].copy().view(np.int32)
    depth = check_tree(nodes, refs, dev.table("slots").reshape(-1, 4, 4), dev.table("slot_nrm").reshape(-1, 4), dev.table("prim_slot"),
                       sc.d.n_primitives)
    ks = checker.scene(sc)
    rng = np.random.default_rng(77)
    raysA, _ = common.camera_rays(ks, rng, 1 << 15, 96, 96)
    raysB, P, N = common.secondary_rays(ks, raysA, rng)
    raysC = common.bbox_rays(ks.info(), rng, 1 << 15)
    flagged = 0
    # the device-built tree against the REFERENCE's own hits (not just against the host-built tree): prim id exact except
    # flagged ties / grazing cases, t / position / normal bit-exact
    report = []
    ps = port.scene(sc)
    for tag, rays in (("A", raysA), ("B", raysB), ("C", raysC)):
        check_hits(name, sc, "lbvh:" + tag, dev, ks, ps, rays, report)
    for rays in (raysA, raysB, raysC):
        h, d = host.unit_scene_intersect(rays), dev.unit_scene_intersect(rays)
        same = h[0] == d[0]
        flagged += int((~same).sum())
        assert (~same).mean() <= 2e-4, (name, float((~same).mean()))  # grazing-edge cases depend on the boxes entered
        for a, b in zip(h[1:], d[1:]):
            assert np.array_equal(a[same], b[same])
    tgt = (P + raysB[:, 3:6] * rng.uniform(0.5, 500, (len(P), 1))).astype(np.float32)
    assert (host.unit_scene_occluded(P, tgt) != dev.unit_scene_occluded(P, tgt)).mean() <= 2e-4
    assert (ks.occluded(P, tgt) != dev.unit_scene_occluded(P, tgt)).mean() <= (2e-2 if name == "large" else 2e-4)
    print("LBVH vs reference:", report)
    print(name, "LBVH depth", depth, "nodes", len(nodes), "vs SAH nodes", len(host.table("nodes")) // 16, "flagged", flagged,
          "build s: gpu %.4f host %.4f" % (dev.stats()["bvh_build_seconds"], host.stats()["bvh_build_seconds"]))
    host.close(); dev.close()


@pytest.mark.parametrize("name", ["bunny", "specular"])
def test_device_built_tree_renders_the_same_image(pkg, port, gpu, name):
    sc = build(pkg, name)
    dev = pkg.Context(sc, gpu_bvh=True)
    dev.render_pass(0, 3, seed=12)
    g = dev.read_film(finalize=False)
    c, _ = port.scene(sc).render_counter(0, 3, 12, numthreads=16)
    bad = (np.abs(g - c) > 1e-4 * np.maximum(np.abs(c), 1.0)).any(axis=2)
    assert bad.mean() <= 1e-2 and abs(g.mean() - c.mean()) <= 2e-3 * c.mean()
    assert dev.stats()["invalid_contributions"] == 0
    dev.close()


def test_tiny_scene_falls_back_to_the_host_builder(pkg, gpu):
    cam = pkg.Camera((0, 0, 5), (0, 0, -1), (0, 1, 0), 60.0, 8, 8)
    sh = pkg.Shape(pkg.SHAPE_SPHERE, 0, ((0, 0, 0), (1, 0, 0), (0, 0, 0), (0, 0, 0)))
    sc = pkg.HostScene.from_arrays(cam, [sh], [pkg.Material(pkg.MAT_MATTE, 0, (.5, .5, .5), (0, 0, 0), 0, 0)],
                                   [pkg.Light(pkg.LIGHT_ENVIRONMENT, -1, (1, 1, 1), (0, 0, 0), (0, 0, 0))], [pkg.Primitive(0, 0, -1)])
    ctx = pkg.Context(sc, gpu_bvh=True)
    assert ctx.stats()["bvh_builder"] == 0
    ctx.render_pass(0, 1, seed=1)
    assert ctx.read_film(finalize=False).max() > 0
    ctx.close()
