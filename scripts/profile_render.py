"""Small fixed render for ncu: python scripts/profile_render.py <scene> <spp> [res] [scale]
One warm-up pass, then one profiled pass (each = 1 generate + (depth+1) x (extend, shade) + depth x connect)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
pkg = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
res = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
sc = pkg.HostScene.builtin(name, res, res, scale)
ctx = pkg.Context(sc)
ctx.set_option("stage_timing", 1)
for i in range(2):
    ctx.clear_film(); ctx.reset_stats()
    ctx.render_pass(0, spp, 1234)
    ctx.synchronize()
st = ctx.stats()
print({k: st[k] for k in ("samples", "extension_rays", "shadow_rays", "kernel_launches", "ms_generate", "ms_extend", "ms_shade", "ms_connect")})
ctx.close()
