"""A/B of the pixel-band wavefronts on the 3840x2160 Cornell config: python scripts/ab_bands.py [spp]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import __graft_entry__ as ge
pkg = ge.load_package()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for name, w, h in [("cornell", 3840, 2160), ("bunny", 3840, 2160), ("bunny", 1024, 1024)]:
    sc = pkg.HostScene.builtin(name, w, h)
    ctx = pkg.Context(sc)
    ctx.set_option("stage_timing", 1)
    for band in (1 << 30, 1 << 21, 1 << 20, 1 << 19, 1 << 18):
        ctx.set_option("band_pixels", band)
        for i in range(2):
            ctx.clear_film(); ctx.reset_stats()
            ctx.render_pass(0, spp, 1234)
            ctx.synchronize()
        st = ctx.stats()
        tot = st["ms_generate"] + st["ms_extend"] + st["ms_shade"] + st["ms_connect"]
        print(json.dumps({"scene": name, "res": [w, h], "band_pixels": band, "spp": spp, "ms_total": round(tot, 2),
                          "Msamples/s": round(w * h * spp / tot / 1e3, 1), "extend": round(st["ms_extend"], 2), "shade": round(st["ms_shade"], 2),
                          "connect": round(st["ms_connect"], 2), "launches": st["kernel_launches"]}))
    ctx.close()
