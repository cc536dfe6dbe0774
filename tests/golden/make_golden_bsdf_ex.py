"""Golden vectors for the BSDF classes no material builds: outputs of the UNMODIFIED reference (oracle/_ref) on seeded
inputs.  Run where /root/reference is mounted and oracle/_ref is built:  python tests/golden/make_golden_bsdf_ex.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge  # noqa: E402
import common  # noqa: E402

pkg, orc = ge.load_package(), ge.load_oracle()
ref = orc.Oracle("ref")
rng = np.random.default_rng(2718)
nrm, wo, wi, u2, _ = common.bsdf_inputs(rng, 2048)
out = dict(nrm=nrm, wo=wo, wi=wi, u2=u2)
for name, d in common.bsdf_ex_cases(pkg).items():
    for k, v in ref.bsdf_ex(d, nrm, wo, wi, u2).items():
        out[f"{name}_{k}"] = v
np.savez_compressed(Path(__file__).parent / "ref_golden_bsdf_ex.npz", **out)
print("wrote", len(out), "arrays")
