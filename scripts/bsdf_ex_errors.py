"""Error distribution of jpbrt_unit_bsdf_ex against the CPU checker (exploration for the test tolerances)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge
import common
pkg, orc = ge.load_package(), ge.load_oracle()
chk = orc.Oracle("ref" if orc.have("ref") else "port")
rng = np.random.default_rng(5)
def vrel(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.maximum(np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1)), 1e-30)
for name, d in common.bsdf_ex_cases(pkg).items():
    i = common.bsdf_inputs(rng, 1 << 17)[:4]
    want, got = chk.bsdf_ex(d, *i), pkg.unit_bsdf_ex(d, *i)
    fl = got["s_flags"] != want["s_flags"]
    out = [f"{name:40s} flags!= {int(fl.sum()):4d}"]
    for k in ("f_eval", "pdf_eval", "s_wi", "s_f", "s_pdf"):
        e = vrel(got[k], want[k]) if got[k].ndim == 2 else common.rel_err(got[k], want[k])
        e = np.nan_to_num(e[~fl], nan=0.0)
        nanmis = int((np.isnan(got[k]) != np.isnan(want[k])).sum())
        out.append(f"{k} q50 {np.quantile(e, .5):.1e} q999 {np.quantile(e, .999):.1e} max {e.max():.1e} >1e-5 {(e > 1e-5).mean():.1e} nan!= {nanmis}")
    print(" | ".join(out))
