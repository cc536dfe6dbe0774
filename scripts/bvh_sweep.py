"""BVH build-parameter sweep: box/prim tests per ray and stage times.  Usage: python scripts/bvh_sweep.py scene"""
import os, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
spp = 8 if name != "glossy" else 2
for leaf in (2, 4, 6, 8):
    for trav in (50, 100, 200, 400):
        os.environ["JPBRT_BVH_LEAF"] = str(leaf); os.environ["JPBRT_BVH_TRAV"] = str(trav)
        sc = pkg.HostScene.builtin(name, 1024, 1024, 1.0)
        ctx = pkg.Context(sc)
        ctx.set_option("count_traversal", 1)
        ctx.render_pass(0, 1, 1); ctx.synchronize()
        c = ctx.stats()
        ctx.set_option("count_traversal", 0); ctx.set_option("stage_timing", 1)
        for i in range(2):
            ctx.clear_film(); ctx.reset_stats(); ctx.render_pass(0, spp, 1234); ctx.synchronize()
        st = ctx.stats()
        print(json.dumps({"scene": name, "leaf": leaf, "trav": trav, "nodes": st["n_nodes"],
                          "box/ray": round(c["box_tests"] / c["extension_rays"], 1), "prim/ray": round(c["prim_tests"] / c["extension_rays"], 2),
                          "sh_box/ray": round(c["shadow_box_tests"] / max(1, c["shadow_rays"]), 1), "sh_prim/ray": round(c["shadow_prim_tests"] / max(1, c["shadow_rays"]), 2),
                          "extend": round(st["ms_extend"], 3), "connect": round(st["ms_connect"], 3), "shade": round(st["ms_shade"], 3)}))
        ctx.close()
