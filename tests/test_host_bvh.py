"""Host logic of jpbrt_upload_scene without a GPU: the flattened BVH and tables (csrc/scene_flatten.cc)."""
import numpy as np
import pytest

import common


def decode(pkg, sc):
    nodes = pkg.debug_flatten(sc, "nodes").reshape(-1, 16)
    refs = nodes[:, 12:14].copy().view(np.int32)
    slots = pkg.debug_flatten(sc, "slots").reshape(-1, 4, 4)
    nrm = pkg.debug_flatten(sc, "slot_nrm").reshape(-1, 4)
    return nodes, refs, slots, nrm


def slot_bounds(slots, nrm):
    """Axis-aligned bounds of every slot's geometry, from the slot records themselves."""
    tag = nrm[:, 3].copy().view(np.int32)
    typ = tag & 3
    n = len(slots)
    lo = np.empty((n, 3)); hi = np.empty((n, 3))
    for i in range(n):
        q = slots[i]
        if typ[i] == 0:
            p = q[:3, :3]
        elif typ[i] == 1:
            p = q[:4, :3]
        elif typ[i] == 2:
            r = q[1, 0]
            p = np.stack([q[0, :3] - r, q[0, :3] + r])
        else:
            r = q[1, 3]
            p = np.stack([q[0, :3] - r, q[0, :3] + r])
        lo[i], hi[i] = p.min(0), p.max(0)
    return lo, hi, tag >> 2


def check_tree(nodes, refs, slots, nrm, prim_slot, n_prims):
    """The flattened BVH is a partition of the primitives into leaves of <= 4 whose boxes nest (any builder)."""
    assert len(slots) == n_prims == len(nrm)
    lo, hi, prim_of_slot = slot_bounds(slots, nrm)
    assert sorted(prim_of_slot.tolist()) == list(range(n_prims))  # every primitive in exactly one slot
    assert np.array_equal(prim_of_slot[prim_slot], np.arange(n_prims))
    seen = np.zeros(n_prims, int)
    visited = np.zeros(len(nodes), int)
    max_depth = 0
    todo = [(0, -np.full(3, 1e30), np.full(3, 1e30), 0)]
    while todo:
        ref, blo, bhi, depth = todo.pop()
        assert depth < 64, "tree deeper than the traversal stack"
        max_depth = max(max_depth, depth)
        if ref < 0:
            bits = ~ref
            first, cnt = bits >> 4, bits & 15
            assert cnt <= 4
            for s in range(first, first + cnt):
                seen[s] += 1
                assert (lo[s] >= blo - 1e-6).all() and (hi[s] <= bhi + 1e-6).all(), "leaf box does not contain its primitive"
            continue
        visited[ref] += 1
        nd = nodes[ref]
        lmin, lmax = nd[[0, 1, 2]], nd[[3, 4, 5]]
        rmin, rmax = nd[[6, 7, 8]], nd[[9, 10, 11]]
        for (cmin, cmax) in ((lmin, lmax), (rmin, rmax)):
            if np.isfinite(cmin).all():
                assert (cmin >= blo - 1e-6).all() and (cmax <= bhi + 1e-6).all(), "child box escapes its parent"
        todo.append((int(refs[ref, 0]), lmin, lmax, depth + 1))
        todo.append((int(refs[ref, 1]), rmin, rmax, depth + 1))
    assert (seen == 1).all(), "a primitive is missing from, or duplicated in, the leaves"
    assert (visited == 1).all(), "an inner node is unreachable or shared"
    return max_depth


@pytest.mark.parametrize("name,scale", [("cornell", 1.0), ("bunny", 0.3), ("glossy", 1.0), ("large", 0.03)])
def test_bvh_is_a_partition_with_containing_boxes(pkg, name, scale):
    sc = pkg.HostScene.builtin(name, 32, 32, scale)
    nodes, refs, slots, nrm = decode(pkg, sc)
    check_tree(nodes, refs, slots, nrm, pkg.debug_flatten(sc, "prim_slot"), sc.d.n_primitives)


def test_single_primitive_scene_gets_a_wrapped_root(pkg):
    cam = pkg.Camera((0, 0, 5), (0, 0, -1), (0, 1, 0), 60.0, 8, 8)
    sh = pkg.Shape(pkg.SHAPE_SPHERE, 0, ((0, 0, 0), (1, 0, 0), (0, 0, 0), (0, 0, 0)))
    sc = pkg.HostScene.from_arrays(cam, [sh], [pkg.Material(pkg.MAT_MATTE, 0, (.5, .5, .5), (0, 0, 0), 0, 0)],
                                   [pkg.Light(pkg.LIGHT_ENVIRONMENT, -1, (1, 1, 1), (0, 0, 0), (0, 0, 0))],
                                   [pkg.Primitive(0, 0, -1)])
    nodes, refs, slots, nrm = decode(pkg, sc)
    assert len(nodes) == 1 and refs[0, 0] == ~((0 << 4) | 1) and refs[0, 1] == ~0
    assert np.isinf(nodes[0, 6:12]).all()  # the empty right box can never be hit


def test_material_and_light_tables(pkg, port):
    sc = pkg.HostScene.builtin("bunny", 16, 16, 0.1)
    mats = pkg.debug_flatten(sc, "materials").reshape(-1, 3, 4)
    d = sc.d
    for i in range(d.n_materials):
        m = d.materials[i]
        assert mats[i, 0, 3:4].copy().view(np.int32)[0] == m.type
        if m.type == pkg.MAT_PLASTIC:  # material.h:94-98: Qd = lum(Kd) / (lum(Kd) + lum(Ks))
            f = np.float32
            lum = lambda c: f(0.212671) * f(c[0]) + f(0.715160) * f(c[1]) + f(0.072169) * f(c[2])  # noqa: E731
            Ld, Ls = lum(m.a), lum(m.b)
            Qd = Ld / (Ld + Ls)
            assert mats[i, 2, 1] == Qd
            assert np.array_equal(mats[i, 0, :3], np.array(list(m.a), np.float32) / Qd)
            assert np.array_equal(mats[i, 1, :3], np.array(list(m.b), np.float32) / (f(1) - Qd))
            assert mats[i, 1, 3] == np.float32(0.1)
    lights = pkg.debug_flatten(sc, "lights").reshape(-1, 6, 4)
    tag = lights[1, 0, 3:4].copy().view(np.int32)[0]
    assert tag & 0xff == pkg.LIGHT_AREA and tag >> 8 == pkg.SHAPE_RECTANGLE
    assert lights[1, 1, 3] == np.float32(1) / np.float32(200 * 200)  # 1 / Area() of the 200 x 200 light (shape.h:457)
    assert np.array_equal(lights[1, 4, :3], np.array([0, -1, 0], np.float32))  # flipped normal faces down
    # environment radius equals the oracle's (light.cc:26-33)
    info = port.scene(sc).info()
    assert info[6] > 0


def sah_cost(nodes, refs):
    """Surface-area-heuristic cost of the flattened tree: sum over inner nodes of area(child) / area(root) x (1 per inner
    child visit, primitive count per leaf)."""
    def half_area(mn, mx):
        d = np.maximum(mx - mn, 0)
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]
    root_lo = np.minimum(nodes[0, [0, 1, 2]], nodes[0, [6, 7, 8]])
    root_hi = np.maximum(nodes[0, [3, 4, 5]], nodes[0, [9, 10, 11]])
    root_area = half_area(root_lo, root_hi)
    cost = 1.0
    for i in range(len(nodes)):
        for k, (lo, hi) in enumerate((([0, 1, 2], [3, 4, 5]), ([6, 7, 8], [9, 10, 11]))):
            mn, mx = nodes[i, lo], nodes[i, hi]
            if not np.isfinite(mn).all():
                continue
            r = int(refs[i, k])
            cost += half_area(mn, mx) / root_area * (1.0 if r >= 0 else float(~r & 15))
    return cost


def test_sweep_sah_builds_a_cheaper_tree_than_16_bins(pkg, monkeypatch):
    """The builder's exact sweep for sets of 1,025..65,536 primitives (B200: bunny scene 34.3 -> 28.3 box tests per ray)
    must not lose to plain 16-bin SAH by the heuristic's own measure."""
    sc = pkg.HostScene.builtin("bunny", 32, 32, 1.0)
    nodes, refs, _, _ = decode(pkg, sc)
    monkeypatch.setenv("JPBRT_BVH_SWEEP_HI", "0")
    monkeypatch.setenv("JPBRT_BVH_BINS_BIG", "16")
    nodes16, refs16, _, _ = decode(pkg, sc)
    c_sweep, c_bins = sah_cost(nodes, refs), sah_cost(nodes16, refs16)
    assert c_sweep < c_bins, (c_sweep, c_bins)
    print("SAH cost: sweep", c_sweep, "16 bins", c_bins)


def test_reinsertion_lowers_the_cost_and_keeps_the_tree_valid(pkg, monkeypatch):
    """Insertion-based optimisation (default for 1,025..131,072 primitives): a valid partition with a lower SAH cost."""
    sc = pkg.HostScene.builtin("bunny", 32, 32, 0.5)
    monkeypatch.setenv("JPBRT_BVH_REINSERT", "0")
    n0, r0, s0, m0 = decode(pkg, sc)
    monkeypatch.setenv("JPBRT_BVH_REINSERT", "2")
    n2, r2, s2, m2 = decode(pkg, sc)
    depth = check_tree(n2, r2, s2, m2, pkg.debug_flatten(sc, "prim_slot"), sc.d.n_primitives)
    assert np.array_equal(s0, s2) and np.array_equal(m0, m2)  # leaves and slot order untouched
    assert sah_cost(n2, r2) < 0.95 * sah_cost(n0, r0) and depth < 56
    monkeypatch.delenv("JPBRT_BVH_REINSERT")
    nd, rd, _, _ = decode(pkg, sc)
    assert np.array_equal(nd.view(np.int32), n2.view(np.int32))  # the default for a scene of this size (bit compare: child refs are ints)


def test_tree_deeper_than_the_traversal_stack_is_rebuilt(pkg, monkeypatch):
    """The kernels' 64-entry node stack would silently lose subtrees of a tree deeper than 62 levels (the host
    SAH build had no depth check).  The uploader measures the depth of whatever its builder produced and rebuilds with
    object-median splits when it exceeds the limit; the limit is lowered here to force that path on a skewed scene
    (nested, corner-anchored triangles of geometrically growing size)."""
    sc = common.skewed_scene(pkg)
    nodes, refs, slots, nrm = decode(pkg, sc)
    natural = check_tree(nodes, refs, slots, nrm, pkg.debug_flatten(sc, "prim_slot"), sc.d.n_primitives)
    assert 12 < natural + 1 <= 62, natural  # (check_tree counts edges below the root)
    monkeypatch.setenv("JPBRT_BVH_MAX_DEPTH", "12")
    nodes, refs, slots, nrm = decode(pkg, sc)
    rebuilt = check_tree(nodes, refs, slots, nrm, pkg.debug_flatten(sc, "prim_slot"), sc.d.n_primitives)
    assert rebuilt + 1 <= 12, rebuilt  # 1,100 primitives, leaves of <= 4: ceil(log2(275)) + 1 = 10 levels


def test_coincident_centroids_build_a_shallow_valid_tree(pkg):
    sc = common.coincident_scene(pkg, n=20000)
    nodes, refs, slots, nrm = decode(pkg, sc)
    depth = check_tree(nodes, refs, slots, nrm, pkg.debug_flatten(sc, "prim_slot"), sc.d.n_primitives)
    assert depth <= 20, depth


@pytest.mark.parametrize("name,scale", [("cornell", 1.0), ("bunny", 1.0), ("glossy", 1.0), ("large", 0.1)])
def test_quantised_nodes_contain_the_float_nodes(pkg, name, scale):
    """The 32-byte nodes the production kernels walk mid-size trees through (csrc/intersect.cuh, QN): every plane index, turned
    back into a coordinate with the grid the kernels use, lies OUTSIDE the float box by at least two cells (the slack the
    kernels' 2^23-offset arithmetic may consume) and by at most four; child references are the float nodes' own."""
    sc = pkg.HostScene.builtin(name, 64, 64, scale)
    nodes = pkg.debug_flatten(sc, "nodes").reshape(-1, 16)
    q = pkg.debug_flatten(sc, "qnodes").reshape(-1, 8)
    grid = pkg.debug_flatten(sc, "qgrid").astype(np.float64)
    origin, cell = grid[:3], grid[3:]
    assert len(q) == len(nodes) and (cell > 0).all()
    assert np.array_equal(q[:, 6:8], nodes[:, 12:14].view(np.uint32))
    planes = q[:, :6].astype(np.int64)
    lo, hi = planes & 0xffff, planes >> 16                       # per (box, axis): L.x L.y L.z R.x R.y R.z
    fmin = np.concatenate([nodes[:, 0:3], nodes[:, 6:9]], axis=1).astype(np.float64)     # Lmin.xyz, Rmin.xyz
    fmax = np.concatenate([nodes[:, 3:6], nodes[:, 9:12]], axis=1).astype(np.float64)    # Lmax.xyz, Rmax.xyz
    o6, c6 = np.tile(origin, 2), np.tile(cell, 2)
    finite = np.isfinite(fmin) & np.isfinite(fmax)               # (the wrapped root of a one-leaf scene has an infinite dummy box)
    below = (fmin - (o6 + lo * c6)) / c6
    above = ((o6 + hi * c6) - fmax) / c6
    assert (below[finite] >= 2.0 - 1e-6).all() and (below[finite] <= 4.0 + 1e-6).all(), (below[finite].min(), below[finite].max())
    assert (above[finite] >= 2.0 - 1e-6).all() and (above[finite] <= 4.0 + 1e-6).all(), (above[finite].min(), above[finite].max())
    assert lo.min() >= 0 and hi.max() <= 65535 and (lo[finite] < hi[finite]).all()


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64, one rounding to float32 at the end
    (the float64 addition's own rounding is 2^29 times finer)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


@pytest.mark.parametrize("name,scale", [("bunny", 1.0), ("large", 0.1), ("cornell", 1.0)])
def test_quantised_slab_arithmetic_never_rejects_a_box_the_float_test_enters(pkg, name, scale):
    """The kernels' quantised slab test, restated in numpy float32 (csrc/intersect.cuh: quant_axis + the QN branch of
    trav_node_step: t = (2^23 + q) * (cell * inv) + ((origin - o) * inv - 2^23 * cell * inv), near / far plane picked by the
    sign of the direction, no widening), run on the scene's real nodes with rays aimed at them -- many nearly parallel to
    an axis, the hard case for the 2^23 cancellation: it must enter every box the exact test on the float box enters."""
    f32 = np.float32
    sc = pkg.HostScene.builtin(name, 64, 64, scale)
    nodes = pkg.debug_flatten(sc, "nodes").reshape(-1, 16)
    q = pkg.debug_flatten(sc, "qnodes").reshape(-1, 8)
    grid = pkg.debug_flatten(sc, "qgrid")
    origin, cell = grid[:3].astype(f32), grid[3:].astype(f32)
    rng = np.random.default_rng(11)
    n = 400_000
    pick = rng.integers(0, len(nodes), n)
    bmin, bmax = nodes[pick, 0:3].astype(np.float64), nodes[pick, 3:6].astype(np.float64)   # the LEFT child box of a random node
    ok = np.isfinite(bmin).all(axis=1) & np.isfinite(bmax).all(axis=1)
    planes = q[pick, 0:3].astype(np.int64)
    qlo, qhi = (planes & 0xffff).astype(f32), (planes >> 16).astype(f32)
    wmin, wmax = nodes[:, 0:3][np.isfinite(nodes[:, 0:3]).all(axis=1)].min(axis=0), nodes[:, 3:6][np.isfinite(nodes[:, 3:6]).all(axis=1)].max(axis=0)
    o = rng.uniform(wmin, wmax, (n, 3))
    ext = np.maximum(bmax - bmin, 1e-3)
    tgt = 0.5 * (bmin + bmax) + rng.uniform(-0.7, 0.7, (n, 3)) * ext                       # inside or just outside the box
    par = rng.random(n) < 0.4                                                              # nearly axis-parallel rays
    k = rng.integers(0, 3, n)
    o[par, k[par]] = tgt[par, k[par]] + rng.uniform(-1e-5, 1e-5, par.sum()) * ext[par, k[par]]
    o = o.astype(f32)
    d = tgt - o.astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(f32)
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        inv = (f32(1) / d).astype(f32)   # (the kernels use the hardware's approximate reciprocal: 1 ulp more, four orders below the slack)
        cinv = (cell * inv).astype(f32)
        oiq = _fma32((origin - o).astype(f32), inv, (f32(-8388608.0) * cinv).astype(f32))
        dead = ~(np.abs(oiq) <= f32(3.4028234664e38))                                      # the axis drops out (NaN constants)
        t_lo = _fma32((f32(8388608.0) + qlo).astype(f32), cinv, oiq)
        t_hi = _fma32((f32(8388608.0) + qhi).astype(f32), cinv, oiq)
        near = np.where(d < 0, t_hi, t_lo)
        far = np.where(d < 0, t_lo, t_hi)
        near[dead], far[dead] = -np.inf, np.inf
        tn = np.maximum(near.max(axis=1), f32(0.001))
        tf = far.min(axis=1)
        hit_q = tn <= tf
        t0 = (bmin - o) / d.astype(np.float64)
        t1 = (bmax - o) / d.astype(np.float64)
        degenerate = ~np.isfinite(t0).all(axis=1) | ~np.isfinite(t1).all(axis=1)
        hit_exact = np.maximum(np.minimum(t0, t1).max(axis=1), 0.001) <= np.maximum(t0, t1).min(axis=1)
    use = ok & ~degenerate
    assert hit_exact[use].sum() > 0.2 * n
    missed = use & hit_exact & ~hit_q
    assert missed.sum() == 0, f"{missed.sum()} boxes entered by the exact test are rejected by the quantised arithmetic"


def test_the_same_scene_flattens_to_the_same_tables(pkg):
    """50 k primitives: large enough for concurrent subtree tasks, small enough for the insertion-based optimisation pass, whose
    visiting order depends on the temporary node ids -- which used to be handed out in thread-timing order (sibling leaves came
    out swapped from one upload to the next).  Such scenes are now built by one thread; bigger ones skip the pass."""
    sc = pkg.HostScene.builtin("large", 64, 64, 0.1)
    first = {t: pkg.debug_flatten(sc, t).tobytes() for t in ("nodes", "qnodes", "slots", "prim_slot")}
    for _ in range(3):
        for t, ref in first.items():
            assert pkg.debug_flatten(sc, t).tobytes() == ref, t
