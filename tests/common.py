"""Shared input generators for the parity tests (ray sets A/B/C, BSDF set)."""
import numpy as np


def unit_vectors(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-30)


def make_rays(o, d, tmin=0.001, tmax=np.inf):
    n = len(o)
    return np.concatenate([o, d, np.full((n, 1), tmin), np.full((n, 1), tmax)], axis=1).astype(np.float32)


def camera_rays(scene_oracle, rng, n, width, height):
    """Set A: camera rays with random film jitter."""
    pf = np.stack([rng.uniform(0, width, n), rng.uniform(0, height, n)], 1).astype(np.float32)
    o, d = scene_oracle.generate_rays(pf)
    return make_rays(o, d), pf


def secondary_rays(scene_oracle, rays, rng):
    """Set B: origins ON surfaces (closest hits of `rays`), uniform random directions: exercises tmin."""
    prim, t, pos, nrm = scene_oracle.intersect(rays)
    m = prim >= 0
    P, N = pos[m], nrm[m]
    d = unit_vectors(rng, len(P))
    return make_rays(P, d), P, N


def bbox_rays(info7, rng, n):
    """Set C: uniform origins in the (slightly enlarged) world box, uniform directions."""
    lo, hi = info7[:3], info7[3:6]
    c, e = 0.5 * (lo + hi), 0.5 * (hi - lo) * 1.2 + 1.0
    o = (c + rng.uniform(-1, 1, (n, 3)) * e).astype(np.float32)
    return make_rays(o, unit_vectors(rng, n))


def bsdf_inputs(rng, n):
    """BSDF set: wo, wi uniform on the sphere (both hemispheres) plus exact-axis and grazing strata."""
    nrm = unit_vectors(rng, n)
    wo = unit_vectors(rng, n)
    wi = unit_vectors(rng, n)
    k = n // 16
    nrm[:k] = np.array([0, 0, 1], np.float32)          # exact axes
    nrm[k:2 * k] = np.array([1, 0, 0], np.float32)     # |n.x| > 0.99 branch of FFrame
    wo[2 * k:3 * k] = nrm[2 * k:3 * k]                 # normal incidence (TrowbridgeReitzSample11 special case)
    # grazing wo: almost perpendicular to n
    g = slice(3 * k, 4 * k)
    t = np.cross(nrm[g], unit_vectors(rng, k))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    wo[g] = (t + 1e-4 * nrm[g]).astype(np.float32)
    wo[g] /= np.linalg.norm(wo[g], axis=1, keepdims=True)
    u2 = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    ul = rng.uniform(0, 1, n).astype(np.float32)
    return nrm, wo.astype(np.float32), wi, u2, ul


def materials(pkg):
    M = pkg.Material
    return {
        "matte": M(pkg.MAT_MATTE, 0, (0.5, 0.4, 0.3), (0, 0, 0), 0, 0),
        "mirror": M(pkg.MAT_MIRROR, 0, (0.9, 0.8, 0.7), (0, 0, 0), 0, 0),
        "glass": M(pkg.MAT_GLASS, 0, (0.98, 0.98, 0.98), (0.98, 0.98, 0.98), 1.5, 0),
        "plastic": M(pkg.MAT_PLASTIC, 0, (0.35, 0.12, 0.48), (0.65, 0.88, 0.52), 0.1, 0),
        "plastic_remap": M(pkg.MAT_PLASTIC, 1, (0.3, 0.3, 0.32), (0.7, 0.7, 0.68), 0.3, 0),
        "metal": M(pkg.MAT_METAL, 0, (0.18, 0.15, 0.81), (0.11, 0.11, 0.11), 0.2, 0.2),
        "metal_aniso_remap": M(pkg.MAT_METAL, 1, (0.2, 0.92, 1.1), (3.9, 2.45, 2.14), 0.4, 0.1),
    }


def shapes(pkg):
    S = pkg.Shape

    def mk(t, flip, pts):
        p = [list(map(float, q)) for q in pts] + [[0.0, 0.0, 0.0]] * (4 - len(pts))
        return S(t, flip, tuple(tuple(q) for q in p))

    return {
        "tri": mk(pkg.SHAPE_TRIANGLE, 0, [(0, 0, 0), (1, 0, 0), (0, 1, 0)]),
        "tri_flip_big": mk(pkg.SHAPE_TRIANGLE, 1, [(343, 548.7, -227), (343, 548.7, -332), (213, 548.7, -332)]),
        "rect": mk(pkg.SHAPE_RECTANGLE, 0, [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]),
        "rect_xz_flip": mk(pkg.SHAPE_RECTANGLE, 1, [(-100, 350, -100), (-100, 350, 100), (100, 350, 100), (100, 350, -100)]),
        "sphere": mk(pkg.SHAPE_SPHERE, 0, [(0.5, -0.25, 2.0), (1.5, 0, 0)]),
        "disk": mk(pkg.SHAPE_DISK, 0, [(0.2, 0.1, -0.3), (0.3, 1.0, -0.2), (0.8, 0, 0)]),
    }


def shape_rays(rng, n, center, radius):
    """Rays aimed near a shape: origins on a shell around it, targets jittered around it (hits, misses, grazing)."""
    o = (center + unit_vectors(rng, n) * radius * rng.uniform(0.2, 6, (n, 1))).astype(np.float32)
    tgt = center + rng.normal(size=(n, 3)) * radius * 0.8
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return make_rays(o, d.astype(np.float32))


def specular_scene(pkg, res=64, max_depth=5):
    """A room for the Whitted / recursive / debug integrators: mirror sphere, mirror wall facing
    it (mirror-mirror chains: the reference traces every mirror vertex twice), glass sphere, plastic box top, matte
    floor, a null-material pane the rays pass through, a rectangle area light, a point light and a dim environment."""
    cam = pkg.Camera((0, 3, 11), (0, -0.12, -1), (0, 1, 0), 60.0, res, res)
    S, M, L, Pm = pkg.Shape, pkg.Material, pkg.Light, pkg.Primitive
    z3 = (0.0, 0.0, 0.0)
    shapes = [S(pkg.SHAPE_RECTANGLE, 1, ((-1.5, 7.5, -1.5), (-1.5, 7.5, 1.5), (1.5, 7.5, 1.5), (1.5, 7.5, -1.5))),   # 0 light
              S(pkg.SHAPE_RECTANGLE, 0, ((-8, 0, -8), (-8, 0, 8), (8, 0, 8), (8, 0, -8))),                           # 1 floor
              S(pkg.SHAPE_SPHERE, 0, ((-2.2, 1.5, 0), (1.5, 0, 0), z3, z3)),                                          # 2 mirror ball
              S(pkg.SHAPE_SPHERE, 0, ((2.0, 1.2, 1.5), (1.2, 0, 0), z3, z3)),                                         # 3 glass ball
              S(pkg.SHAPE_RECTANGLE, 0, ((-6, 0, -4), (-6, 6, -4), (4, 6, -5), (4, 0, -5))),                          # 4 mirror wall
              S(pkg.SHAPE_RECTANGLE, 0, ((3.5, 0.8, -2), (3.5, 0.8, 0), (5.5, 0.8, 0), (5.5, 0.8, -2))),             # 5 plastic slab
              S(pkg.SHAPE_RECTANGLE, 0, ((-1, 0, 5), (-1, 4, 5), (1, 4, 5), (1, 0, 5))),                              # 6 null pane
              S(pkg.SHAPE_TRIANGLE, 0, ((5, 0, -4.5), (7, 0, -3), (6, 4, -4), z3))]                                   # 7 metal fin
    mats = [M(pkg.MAT_MATTE, 0, (.6, .55, .5), z3, 0, 0), M(pkg.MAT_MIRROR, 0, (.9, .85, .8), z3, 0, 0),
            M(pkg.MAT_GLASS, 0, (.95, .95, .95), (.95, .95, .95), 1.5, 0), M(pkg.MAT_PLASTIC, 0, (.3, .1, .4), (.7, .9, .6), 0.15, 0),
            M(pkg.MAT_METAL, 0, (.18, .15, .81), (.11, .11, .11), 0.2, 0.2)]
    lights = [L(pkg.LIGHT_ENVIRONMENT, -1, (.08, .08, .2), z3, z3), L(pkg.LIGHT_AREA, 0, (14, 13, 12), z3, z3),
              L(pkg.LIGHT_POINT, -1, (20, 20, 20), (5, 6, 4), z3)]
    prims = [Pm(0, 0, 1), Pm(1, 0, -1), Pm(2, 1, -1), Pm(3, 2, -1), Pm(4, 1, -1), Pm(5, 3, -1), Pm(6, -1, -1), Pm(7, 4, -1)]
    return pkg.HostScene.from_arrays(cam, shapes, mats, lights, prims, max_depth=max_depth, name="specular")


def bsdf_ex_cases(pkg):
    """The BSDF classes no material of the reference builds (jpbrt_bsdf_desc)."""
    def D(**k):
        d = pkg.BsdfDesc()
        for a, v in k.items():
            setattr(d, a, (pkg.C.c_float * 3)(*v) if isinstance(v, (tuple, list)) else v)
        return d
    P, MR, MT = pkg.BSDF_PHONG, pkg.BSDF_MICROFACET_REFLECTION, pkg.BSDF_MICROFACET_TRANSMISSION
    BK, TR = pkg.DIST_BECKMANN, pkg.DIST_TROWBRIDGE_REITZ
    return {
        "phong": D(kind=P, color=(.8, .7, .6), exponent=20.0),
        "phong_soft": D(kind=P, color=(.5, .5, .5), exponent=3.0),
        "refl_beckmann_visible_dielectric": D(kind=MR, distribution=BK, sample_visible_area=1, fresnel=pkg.FRESNEL_DIELECTRIC, color=(.9, .9, .9),
                                              alphax=.2, alphay=.2, eta_a=1.0, eta_b=1.5),
        "refl_beckmann_visible_aniso_conductor": D(kind=MR, distribution=BK, sample_visible_area=1, fresnel=pkg.FRESNEL_CONDUCTOR, color=(1, 1, 1),
                                                   alphax=.3, alphay=.1, c_eta_i=(1, 1, 1), c_eta_t=(.2, .92, 1.1), c_k=(3.9, 2.45, 2.14)),
        "refl_beckmann_full": D(kind=MR, distribution=BK, sample_visible_area=0, fresnel=pkg.FRESNEL_NOOP, color=(1, 1, 1), alphax=.25, alphay=.25),
        "refl_beckmann_full_aniso": D(kind=MR, distribution=BK, sample_visible_area=0, fresnel=pkg.FRESNEL_NOOP, color=(1, 1, 1), alphax=.4, alphay=.15),
        "refl_tr_full": D(kind=MR, distribution=TR, sample_visible_area=0, fresnel=pkg.FRESNEL_DIELECTRIC, color=(1, 1, 1), alphax=.3, alphay=.3,
                          eta_a=1.0, eta_b=1.5),
        "refl_tr_full_aniso": D(kind=MR, distribution=TR, sample_visible_area=0, fresnel=pkg.FRESNEL_NOOP, color=(1, 1, 1), alphax=.3, alphay=.12),
        "trans_tr": D(kind=MT, distribution=TR, sample_visible_area=1, color=(.95, .95, .95), alphax=.2, alphay=.2, eta_a=1.0, eta_b=1.5),
        "trans_beckmann_aniso": D(kind=MT, distribution=BK, sample_visible_area=1, color=(.9, .95, .9), alphax=.15, alphay=.3, eta_a=1.0, eta_b=1.33),
    }


# Philox4x32-10 known-answer vectors (Random123 kat_vectors: counter[4], key[2] -> output[4])
PHILOX_KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def skewed_scene(pkg, n=1100, ratio=1.03, res=32):
    """Corner-anchored triangles of geometrically growing size: exact-sweep SAH peels the largest one off at every level,
    i.e. builds a chain one primitive per level -- deeper than the traversal stack unless the uploader rebuilds it."""
    cam = pkg.Camera((3.0, 2.0, 40.0), (-0.05, -0.03, -1.0), (0, 1, 0), 60.0, res, res)
    S, M, L, Pm = pkg.Shape, pkg.Material, pkg.Light, pkg.Primitive
    z3 = (0.0, 0.0, 0.0)
    shapes = []
    for k in range(n):
        s = float(np.float32(0.5 * ratio ** k))
        shapes.append(S(pkg.SHAPE_TRIANGLE, 0, ((0, 0, -0.001 * k), (s, 0, -0.001 * k), (0, s, -0.001 * k - 0.5 * s), z3)))
    mats = [M(pkg.MAT_MATTE, 0, (.7, .6, .5), z3, 0, 0)]
    lights = [L(pkg.LIGHT_ENVIRONMENT, -1, (.6, .7, .9), z3, z3)]
    prims = [Pm(k, 0, -1) for k in range(n)]
    return pkg.HostScene.from_arrays(cam, shapes, mats, lights, prims, max_depth=3, name="skewed")


def coincident_scene(pkg, n=100000, res=32):
    """n triangles with IDENTICAL centroids (concentric, growing): no builder can separate them by position."""
    cam = pkg.Camera((0.0, 0.0, 30.0), (0, 0, -1.0), (0, 1, 0), 60.0, res, res)
    S, M, L, Pm = pkg.Shape, pkg.Material, pkg.Light, pkg.Primitive
    z3 = (0.0, 0.0, 0.0)
    shapes = []
    for k in range(n):
        s = 1.0 + (k % 1000) * 0.004
        shapes.append(S(pkg.SHAPE_TRIANGLE, 0, ((-s, -s, 0.0), (2 * s, -s, 0.0), (-s, 2 * s, 0.0), z3)))  # centroid (0, 0, 0) for every s
    mats = [M(pkg.MAT_MATTE, 0, (.7, .6, .5), z3, 0, 0)]
    lights = [L(pkg.LIGHT_ENVIRONMENT, -1, (.6, .7, .9), z3, z3)]
    prims = [Pm(k, 0, -1) for k in range(n)]
    return pkg.HostScene.from_arrays(cam, shapes, mats, lights, prims, max_depth=2, name="coincident")


def chain_scene(pkg, n=110, res=16):
    """Small triangles across the x axis at x = 2^k: with an exact SAH sweep (JPBRT_BVH_SWEEP) every split peels off the
    farthest one -- a chain of ~n levels.  A ray up the axis enters every box near-child (the rest) first and pushes one
    far leaf per level."""
    cam = pkg.Camera((-1.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0, 1, 0), 20.0, res, res)
    S, M, L, Pm = pkg.Shape, pkg.Material, pkg.Light, pkg.Primitive
    z3 = (0.0, 0.0, 0.0)
    shapes = []
    for k in range(n):
        x, h = float(2.0 ** k), 0.25 * float(2.0 ** k)
        shapes.append(S(pkg.SHAPE_TRIANGLE, 0, ((x, -h, -h), (x, h, -h), (x, 0.0, h), z3)))
    mats = [M(pkg.MAT_MATTE, 0, (.7, .6, .5), z3, 0, 0)]
    lights = [L(pkg.LIGHT_ENVIRONMENT, -1, (.6, .7, .9), z3, z3)]
    prims = [Pm(k, 0, -1) for k in range(n)]
    return pkg.HostScene.from_arrays(cam, shapes, mats, lights, prims, max_depth=2, name="chain")
