"""Golden films of the reference's other integrators (FPathIntegratorRecursive, FWhittedIntegrator, FDebugIntegrator):
outputs of the UNMODIFIED reference (oracle/_ref) for small renders.  python tests/golden/make_golden_modes.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge  # noqa: E402
import common  # noqa: E402

pkg, orc = ge.load_package(), ge.load_oracle()
ref = orc.Oracle("ref")
out = {}
for sname, sc in (("specular", common.specular_scene(pkg, 40)), ("bunny", pkg.HostScene.builtin("bunny", 32, 32, 0.3))):
    for mode, m in orc.INTEGRATORS.items():
        film, _ = ref.scene(sc).render(2, 4, mode=m)
        out[f"{sname}_{mode}"] = film
np.savez_compressed(Path(__file__).parent / "ref_golden_modes.npz", **out)
print("wrote", sorted(out))
