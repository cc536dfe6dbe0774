"""Writes tests/golden/ref_golden_io.npz from the UNMODIFIED reference (oracle/_ref): the bytes FFilm::SaveAsImage
produces for the test films of tests/test_reference_io_pin.py and the triangles LoadTriangleMesh reads from its OBJ.
Run where /root/reference is mounted:  python tests/golden/make_golden_io.py"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as ge  # noqa: E402
import test_reference_io_pin as T  # noqa: E402

ref = ge.load_oracle().Oracle("ref")
out = {}
with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    for w, h in [(8, 6), (64, 3)]:
        for kind, ext in T.EXT.items():
            assert ref.save_image(str(d / "r"), kind, T.film_pattern(w, h, positive=(kind == 2))) == 0
            out[f"{ext}_{w}x{h}"] = np.frombuffer((d / f"r.{ext}").read_bytes(), np.uint8)
    tris, nrm = ref.load_obj(T.obj_file(d), False, True, (10, 20, 30), 2.0)
    out["obj_tris_flip_10_20_30_x2"] = tris
np.savez_compressed(ROOT / "tests" / "golden" / "ref_golden_io.npz", **out)
print({k: v.shape for k, v in out.items()})
