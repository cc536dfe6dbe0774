"""Generate tests/golden/ref_golden.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container, where /root/reference is mounted:
    make -C oracle ref && python tests/golden/make_golden.py
The reference ships no golden vectors; these are its own outputs on seeded inputs, so
that the pin of oracle/pt_oracle.cc and of the CUDA path survives on machines without the reference.
Inputs are stored beside the outputs; nothing is regenerated at test time.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge  # noqa: E402
import common  # noqa: E402

pkg = ge.load_package()
orc = ge.load_oracle()
ref = orc.Oracle("ref")
rng = np.random.default_rng(20261018)
G = {}

# 1. single-shape intersection
for name, sh in common.shapes(pkg).items():
    pts = np.array([[sh.p[i][k] for k in range(3)] for i in range(4)], np.float64)
    if sh.type in (pkg.SHAPE_SPHERE,):
        c, r = pts[0], float(sh.p[1][0])
    elif sh.type == pkg.SHAPE_DISK:
        c, r = pts[0], float(sh.p[2][0])
    else:
        k = 3 if sh.type == pkg.SHAPE_TRIANGLE else 4
        c = pts[:k].mean(0)
        r = np.linalg.norm(pts[:k] - c, axis=1).max()
    rays = common.shape_rays(rng, 768, c, r)
    hit, t, pos, nrm = ref.intersect_shape(sh, rays)
    G[f"shape_{name}_rays"] = rays
    G[f"shape_{name}_hit"], G[f"shape_{name}_t"], G[f"shape_{name}_pos"], G[f"shape_{name}_nrm"] = hit, t, pos, nrm

# 2. BSDFs
for name, m in common.materials(pkg).items():
    nrm, wo, wi, u2, ul = common.bsdf_inputs(rng, 768)
    r = ref.bsdf(m, nrm, wo, wi, u2, ul)
    for k, v in dict(nrm=nrm, wo=wo, wi=wi, u2=u2, ul=ul).items():
        G[f"bsdf_{name}_in_{k}"] = v
    for k, v in r.items():
        G[f"bsdf_{name}_{k}"] = v

# 3. scenes: camera rays, closest hits, occlusion, light samples, emission, small films
SCENES = [("cornell", 1.0, 40, 40, 4), ("bunny", 0.2, 40, 40, 4), ("glossy", 1.0, 32, 32, 2), ("large", 0.02, 32, 32, 2)]
for name, scale, w, h, spp in SCENES:
    sc = pkg.HostScene.builtin(name, w, h, scale)
    rs = ref.scene(sc)
    G[f"scene_{name}_cfg"] = np.array([scale, w, h, spp], np.float64)
    G[f"scene_{name}_info"] = rs.info()
    rays, pf = common.camera_rays(rs, rng, 1024, w, h)
    G[f"scene_{name}_posfilm"] = pf
    G[f"scene_{name}_rays"] = rays
    prim, t, pos, nrm = rs.intersect(rays)
    G[f"scene_{name}_prim"], G[f"scene_{name}_t"], G[f"scene_{name}_pos"], G[f"scene_{name}_nrm"] = prim, t, pos, nrm
    rays2, P, N = common.secondary_rays(rs, rays, rng)
    prim2, t2, pos2, nrm2 = rs.intersect(rays2)
    G[f"scene_{name}_rays2"] = rays2
    G[f"scene_{name}_prim2"], G[f"scene_{name}_t2"], G[f"scene_{name}_nrm2"] = prim2, t2, nrm2
    tgt = (P + rays2[:, 3:6] * rng.uniform(1, 900, (len(P), 1))).astype(np.float32)
    G[f"scene_{name}_occ_tgt"] = tgt
    G[f"scene_{name}_occ"] = rs.occluded(P, tgt)
    G[f"scene_{name}_Le"] = rs.emitted(prim, nrm, -rays[:, 3:6])
    for li in range(min(sc.d.n_lights, 3)):
        u2 = rng.uniform(0, 1, (len(P), 2)).astype(np.float32)
        lpos, wi, pdf, Li = rs.light_sample(li, P, N, u2)
        G[f"scene_{name}_light{li}_u2"] = u2
        G[f"scene_{name}_light{li}_lpos"], G[f"scene_{name}_light{li}_wi"] = lpos, wi
        G[f"scene_{name}_light{li}_pdf"], G[f"scene_{name}_light{li}_Li"] = pdf, Li
    film, _ = rs.render(spp, 4)          # FIntegrator::Render, 4 threads, seed 1234
    G[f"scene_{name}_film"] = film
    film0, _ = rs.render(1, 0)           # numthreads < 1: inline path, one stream over the image
    G[f"scene_{name}_film_inline"] = film0

out = Path(__file__).resolve().parent / "ref_golden.npz"
np.savez_compressed(out, **G)
print("wrote", out, out.stat().st_size, "bytes,", len(G), "arrays")
