"""First-contact GPU probe: parity spot checks + timings.  Writes gpurun_out/probe.json."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle()
out = {}
port = orc.Oracle("port")
ref = orc.Oracle("ref") if orc.have("ref") else port
print("devices", pkg.device_count())
rng = np.random.default_rng(1)

# 1 rng
pix = rng.integers(0, 2**20, 64).astype(np.uint32); smp = rng.integers(0, 4096, 64).astype(np.uint32); blk = rng.integers(0, 40, 64).astype(np.uint32)
g = pkg.unit_rng_block(pix, smp, blk, 0x1234567890ab)
c = np.stack([port.philox_block(int(a), int(b), int(d), 0x1234567890ab) for a, b, d in zip(pix, smp, blk)])
out["rng_equal"] = bool(np.array_equal(g, c)); print("rng equal", out["rng_equal"])

def rel(a, b):
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-30)

for name, scale in [("cornell", 1), ("bunny", 1), ("glossy", 1)]:
    sc = pkg.HostScene.builtin(name, 256, 256, scale)
    ctx = pkg.Context(sc)
    rs = ref.scene(sc); ps = port.scene(sc)
    n = 1 << 16
    pf = np.stack([rng.uniform(0, 256, n), rng.uniform(0, 256, n)], 1).astype(np.float32)
    o_g, d_g = ctx.unit_generate_rays(pf); o_r, d_r = rs.generate_rays(pf)
    out[f"{name}_gen_equal"] = bool(np.array_equal(d_g, d_r))
    rays = np.concatenate([o_r, d_r, np.full((n, 1), 0.001, np.float32), np.full((n, 1), np.inf, np.float32)], 1).astype(np.float32)
    pg, tg, posg, ng = ctx.unit_scene_intersect(rays)
    pr, tr, posr, nr = rs.intersect(rays)
    pb, tb, _, _ = ps.intersect_brute(rays)
    same = pg == pr
    out[f"{name}_cam_prim_mismatch"] = int((~same).sum())
    out[f"{name}_cam_t_equal_when_same"] = bool(np.array_equal(tg[same], tr[same]))
    out[f"{name}_cam_brute_mismatch"] = int((pg != pb).sum())
    out[f"{name}_cam_nrm_equal"] = bool(np.array_equal(ng[same], nr[same]))
    # secondary rays: from hit points in random directions
    hitm = pr >= 0
    P = posr[hitm]; m = len(P)
    dirs = rng.normal(size=(m, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rays2 = np.concatenate([P, dirs, np.full((m, 1), 0.001), np.full((m, 1), np.inf)], 1).astype(np.float32)
    pg2, tg2, _, ng2 = ctx.unit_scene_intersect(rays2); pr2, tr2, _, nr2 = rs.intersect(rays2); pb2, tb2, _, _ = ps.intersect_brute(rays2)
    same2 = pg2 == pr2
    out[f"{name}_sec_n"] = int(m)
    out[f"{name}_sec_prim_mismatch"] = int((~same2).sum())
    out[f"{name}_sec_t_equal_when_same"] = bool(np.array_equal(tg2[same2], tr2[same2]))
    out[f"{name}_sec_brute_mismatch"] = int((pg2 != pb2).sum())
    out[f"{name}_sec_ref_vs_brute_mismatch"] = int((pr2 != pb2).sum())
    # occlusion
    tgt = P + dirs * rng.uniform(1, 800, (m, 1))
    og = ctx.unit_scene_occluded(P, tgt); orr = rs.occluded(P, tgt)
    out[f"{name}_occ_mismatch"] = int((og != orr).sum())
    # lights
    nl = sc.d.n_lights
    N = nr[hitm]
    for li in range(min(nl, 3)):
        u2 = rng.uniform(0, 1, (m, 2)).astype(np.float32)
        lg = ctx.unit_light_sample(li, P, N, u2); lr = rs.light_sample(li, P, N, u2)
        out[f"{name}_light{li}_maxrel"] = [float(rel(a, b).max()) for a, b in zip(lg, lr)]
    # image parity vs counter oracle
    ctx.render_pass(0, 2, seed=99)
    fg = ctx.read_film(finalize=False)
    fc, _ = ps.render_counter(0, 2, 99, numthreads=16)
    bad = (np.abs(fg - fc) > 1e-4 * np.maximum(np.abs(fc), 1.0)).any(axis=2)
    out[f"{name}_img_bad_frac"] = float(bad.mean()); out[f"{name}_img_mean"] = [float(fg.mean()), float(fc.mean())]
    st = ctx.stats(); out[f"{name}_stats_small"] = {k: st[k] for k in ("samples", "extension_rays", "shadow_rays", "invalid_contributions", "kernel_launches")}
    print(name, {k: v for k, v in out.items() if k.startswith(name)})
    ctx.close()

# bsdf
mats = {
 "matte": pkg.Material(pkg.MAT_MATTE, 0, (0.5, 0.4, 0.3), (0, 0, 0), 0, 0),
 "mirror": pkg.Material(pkg.MAT_MIRROR, 0, (0.9, 0.8, 0.7), (0, 0, 0), 0, 0),
 "glass": pkg.Material(pkg.MAT_GLASS, 0, (0.98, 0.98, 0.98), (0.98, 0.98, 0.98), 1.5, 0),
 "plastic": pkg.Material(pkg.MAT_PLASTIC, 0, (0.35, 0.12, 0.48), (0.65, 0.88, 0.52), 0.1, 0),
 "metal": pkg.Material(pkg.MAT_METAL, 0, (0.18, 0.15, 0.81), (0.11, 0.11, 0.11), 0.2, 0.2),
 "metal_aniso_remap": pkg.Material(pkg.MAT_METAL, 1, (0.2, 0.92, 1.1), (3.9, 2.45, 2.14), 0.4, 0.1),
}
n = 1 << 16
def sph(n):
    v = rng.normal(size=(n, 3)); return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
for mn, m in mats.items():
    nrm, wo, wi = sph(n), sph(n), sph(n)
    u2 = rng.uniform(0, 1, (n, 2)).astype(np.float32); ul = rng.uniform(0, 1, n).astype(np.float32)
    g = pkg.unit_bsdf(m, nrm, wo, wi, u2, ul); r = ref.bsdf(m, nrm, wo, wi, u2, ul)
    res = {}
    for k in g:
        if g[k].dtype == np.int32: res[k] = int((g[k] != r[k]).sum())
        else:
            e = rel(g[k], r[k]); res[k] = [float(np.nanmax(e)), float((e > 1e-5).mean())]
    out[f"bsdf_{mn}"] = res; print("bsdf", mn, res)

# timings
for name, res_, spp in [("cornell", 1024, 16), ("bunny", 1024, 16), ("glossy", 1024, 4)]:
    sc = pkg.HostScene.builtin(name, res_, res_, 1)
    t0 = time.time(); ctx = pkg.Context(sc); t_up = time.time() - t0
    ctx.set_option("stage_timing", 1)
    ctx.render_pass(0, spp, 1); ctx.synchronize(); ctx.clear_film()
    t0 = time.time(); ctx.render_pass(0, spp, 1); ctx.synchronize(); dt = time.time() - t0
    st = ctx.stats()
    out[f"time_{name}"] = dict(upload_s=t_up, render_s=dt, msamples_s=res_ * res_ * spp / dt / 1e6, mrays_s=(st["extension_rays"] + st["shadow_rays"]) / dt / 1e6, stats=st)
    print("time", name, out[f"time_{name}"])
    ctx.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "probe.json", "w"), indent=1, default=str)
