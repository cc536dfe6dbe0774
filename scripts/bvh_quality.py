"""SAH quality knobs (env JPBRT_BVH_SWEEP = exact sweep SAH below that set size): box/prim tests per ray and stage times."""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as ge
pkg = ge.load_package()
scenes = (sys.argv[1] if len(sys.argv) > 1 else "cornell,bunny,glossy").split(",")
for name in scenes:
    spp = 16 if name != "glossy" else 4
    for sweep, bins in ((0, 16), (0, 32), (0, 64), (0, 128), (0, 256), (1 << 30, 16)):
        if name == "large" and sweep > 1024:
            continue
        os.environ["JPBRT_BVH_SWEEP"] = str(sweep)
        os.environ["JPBRT_BVH_BINS_BIG"] = str(bins)
        sc = pkg.HostScene.builtin(name, 1024, 1024, 1.0)
        ctx = pkg.Context(sc)
        ctx.set_option("count_traversal", 1)
        ctx.render_pass(0, 1, 1); ctx.synchronize()
        c = ctx.stats()
        ctx.set_option("count_traversal", 0); ctx.set_option("stage_timing", 1)
        for i in range(2):
            ctx.clear_film(); ctx.reset_stats(); ctx.render_pass(0, spp, 1234); ctx.synchronize()
        st = ctx.stats()
        print(json.dumps({"scene": name, "sweep_max": sweep, "bins_big": bins, "nodes": st["n_nodes"], "build_s": round(st["bvh_build_seconds"], 3),
                          "box/ray": round(c["box_tests"] / c["extension_rays"], 1), "prim/ray": round(c["prim_tests"] / c["extension_rays"], 2),
                          "sh_box/ray": round(c["shadow_box_tests"] / max(1, c["shadow_rays"]), 1), "sh_prim/ray": round(c["shadow_prim_tests"] / max(1, c["shadow_rays"]), 2),
                          "extend": round(st["ms_extend"], 3), "connect": round(st["ms_connect"], 3)}))
        ctx.close()
