"""Unit queries of the loaded library on a fixed ray set, saved for comparison between two builds:
python scripts/dump_queries.py <scene> <out.npz>   (camera rays, secondary rays from their hits, shadow segments to a point above)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
pkg = ge.load_package()
name, out = sys.argv[1], sys.argv[2]
sc = pkg.HostScene.builtin(name, 1024, 1024, 1.0)
ctx = pkg.Context(sc)
rng = np.random.default_rng(7)
n = 1 << 20
cam = sc.d.camera
pos = np.array([cam.pos[0], cam.pos[1], cam.pos[2]], np.float32)
# camera-like rays: from the camera position towards random points of the scene's bounds
st = ctx.stats()
tgt = rng.uniform(-500, 500, (n, 3)).astype(np.float32) if name == "large" else rng.uniform(-250, 400, (n, 3)).astype(np.float32)
d = tgt - pos; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([np.tile(pos, (n, 1)), d, np.full((n, 1), 0.001), np.full((n, 1), np.inf)], 1).astype(np.float32)
prim, t, p, nr = ctx.unit_scene_intersect(rays)
hit = prim >= 0
# secondary rays from the hit points, random directions
d2 = rng.normal(size=(n, 3)); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
rays2 = np.concatenate([np.where(hit[:, None], p, pos), d2, np.full((n, 1), 0.001), np.full((n, 1), np.inf)], 1).astype(np.float32)
prim2, t2, p2, _ = ctx.unit_scene_intersect(rays2)
# shadow segments from the hit points to random points high above
top = np.stack([rng.uniform(-150, 150, n), np.full(n, 600.0 if name == "large" else 350.0), rng.uniform(-150, 150, n)], 1).astype(np.float32)
occ = ctx.unit_scene_occluded(np.where(hit[:, None], p, pos), top)
occ2 = ctx.unit_scene_occluded(np.where((prim2 >= 0)[:, None], p2, pos), top)
np.savez(out, prim=prim, t=t, prim2=prim2, t2=t2, occ=occ, occ2=occ2)
print(name, "hits", int(hit.sum()), "hits2", int((prim2 >= 0).sum()), "occ", int(occ.sum()), "occ2", int(occ2.sum()))
