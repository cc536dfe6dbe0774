"""Host binned-SAH build vs GPU linear-BVH build: build time and render throughput.  python scripts/bvh_builders.py [scenes] [spp]"""
import sys, json, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as ge
pkg = ge.load_package()
scenes = (sys.argv[1] if len(sys.argv) > 1 else "bunny,large").split(",")
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
for name in scenes:
    sc = pkg.HostScene.builtin(name, 1024, 1024, 1.0)
    for gpu_bvh in (False, True):
        t0 = time.time()
        ctx = pkg.Context(sc, gpu_bvh=gpu_bvh)
        upload_s = time.time() - t0
        ctx.set_option("stage_timing", 1)
        ctx.set_option("count_traversal", 1)
        ctx.render_pass(0, 2, 1); ctx.synchronize()
        st = ctx.stats()
        boxes = st["box_tests"] / max(1, st["extension_rays"])
        ctx.set_option("count_traversal", 0)
        for i in range(2):
            ctx.clear_film(); ctx.reset_stats()
            ctx.render_pass(0, spp, 1234); ctx.synchronize()
        st = ctx.stats()
        tot = st["ms_generate"] + st["ms_extend"] + st["ms_shade"] + st["ms_connect"]
        print(json.dumps({"scene": name, "builder": "gpu-lbvh" if st["bvh_builder"] else "host-sah", "n_prims": sc.d.n_primitives,
                          "bvh_build_s": round(st["bvh_build_seconds"], 4), "bvh_device_s": round(st["bvh_device_seconds"], 4), "upload_total_s": round(upload_s, 3), "nodes": st["n_nodes"],
                          "box_tests_per_ext_ray": round(boxes, 1), "Msamples/s": round(1024 * 1024 * spp / tot / 1e3, 1),
                          "extend_ms": round(st["ms_extend"], 2), "connect_ms": round(st["ms_connect"], 2)}))
        ctx.close()
