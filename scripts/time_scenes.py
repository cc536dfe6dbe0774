"""Stage timings for the BASELINE scenes: python scripts/time_scenes.py [opt=value ...] [--scenes a,b] [--spp n]"""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
pkg = ge.load_package()
opts = {}
scenes = ["cornell", "bunny", "glossy"]
scales = {"large": 1.0}
spp = 8
for a in sys.argv[1:]:
    if a.startswith("--scenes="): scenes = a.split("=")[1].split(",")
    elif a.startswith("--spp="): spp = int(a.split("=")[1])
    elif "=" in a:
        k, v = a.split("="); opts[k] = int(v)
# the first process-level renders run before the GPU's clocks have ramped up: warm up on the first scene, untimed
_w = pkg.Context(pkg.HostScene.builtin(scenes[0], 1024, 1024, 1.0))
for i in range(12):
    _w.render_pass(0, 8, 1)
_w.synchronize(); _w.close()
for name in scenes:
    res = 1024
    scale = 1.0
    sc = pkg.HostScene.builtin(name, res, res, scale)
    ctx = pkg.Context(sc)
    for k, v in opts.items(): ctx.set_option(k, v)
    ctx.set_option("stage_timing", 1)
    n = spp if name != "glossy" else max(1, spp // 4)
    for i in range(3):
        ctx.clear_film(); ctx.reset_stats()
        ctx.render_pass(0, n, 1234)
        ctx.synchronize()
    st = ctx.stats()
    tot = st["ms_generate"] + st["ms_extend"] + st["ms_shade"] + st["ms_connect"]
    print(json.dumps({"scene": name, "opts": opts, "spp": n, "ms_total": round(tot, 3), "Msamples/s": round(res * res * n / tot / 1e3, 1),
                      "Mrays/s": round((st["extension_rays"] + st["shadow_rays"]) / tot / 1e3, 1),
                      "extend": round(st["ms_extend"], 3), "shade": round(st["ms_shade"], 3), "connect": round(st["ms_connect"], 3),
                      "film_mean": float(ctx.read_film(finalize=False).mean())}))
    ctx.close()
