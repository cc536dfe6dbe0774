"""The reference-side binding, BUILT: include/b200_integrator.h -- the FIntegrator::Render-shaped class
of INTEGRATION.md -- is compiled against the unmodified reference (oracle/Makefile `shim`) and driven the way
main.cc:149-160 drives an integrator: FFilm + FRandomSampler(spp) + Render + the reference's own FFilm::SaveAsImage.
What lands in the reference's FFilm must be the C ABI's image, and the file the reference then writes must be the file
jpbrt_save_image writes for those pixels."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHIM = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "libjetpbrt_refshim.so"
EXT = {0: "ppm", 1: "bmp", 2: "hdr"}


@pytest.fixture(scope="module")
def shim(pkg):
    if not SHIM.exists():
        pytest.skip("oracle/_ref/libjetpbrt_refshim.so not built (needs /root/reference at build time)")
    lib = C.CDLL(str(SHIM))
    lib.jshim_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_ulonglong, C.c_char_p, C.c_int, C.POINTER(C.c_float)]
    return lib


@pytest.mark.parametrize("name,w,h,spp", [("cornell", 128, 96, 4), ("bunny", 96, 64, 3)])
def test_render_through_the_reference_side_binding(pkg, shim, gpu, tmp_path, name, w, h, spp):
    sc = pkg.HostScene.builtin(name, w, h, 0.3 if name == "bunny" else 1.0)
    want, _ = pkg.render(sc, spp, seed=1234)
    for kind in (1, 0, 2):
        film = np.zeros((h, w, 3), np.float32)
        rc = shim.jshim_render(C.cast(sc.desc, C.c_void_p), sc.d.max_depth, spp, 0, 1, 1234, str(tmp_path / "ref").encode(), kind,
                               film.ctypes.data_as(C.POINTER(C.c_float)))
        assert rc == 0
        # the FFilm the reference's main() would go on to save == the C ABI's Clamp01(mean) image (float atomics: summation order)
        np.testing.assert_allclose(film, want, rtol=2e-5, atol=2e-6)
        assert film.min() >= 0 and film.max() <= 1 and film.mean() > 0.01
        # ... and the file FFilm::SaveAsImage wrote from it == the file jpbrt_save_image writes from the same pixels
        # (HDR: the reference leaves pixels below 1e-32 uninitialised, film.cc:159-181 -- compared where defined)
        pkg.save_image(str(tmp_path / "ours"), kind, film)
        a, b = (tmp_path / f"ref.{EXT[kind]}").read_bytes(), (tmp_path / f"ours.{EXT[kind]}").read_bytes()
        if kind == 2:
            head = a.index(b"\n", a.index(b"-Y")) + 1
            assert a[:head] == b[:head] and len(a) == len(b)
            pa, pb = np.frombuffer(a[head:], np.uint8).reshape(-1, 4), np.frombuffer(b[head:], np.uint8).reshape(-1, 4)
            defined = film.reshape(-1, 3).max(axis=1) >= 1e-32
            assert np.array_equal(pa[defined], pb[defined]) and not pb[~defined].any()
        else:
            assert a == b, (name, EXT[kind])


def test_binding_adds_onto_the_callers_film_and_checks_the_resolution(pkg, shim, gpu):
    """FIntegrator::Render ADDS to the film it is given (film.h:64-68); a mismatching film is refused, not overrun."""
    sc = pkg.HostScene.builtin("cornell", 64, 48)
    a = np.zeros((48, 64, 3), np.float32)
    f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    assert shim.jshim_render(C.cast(sc.desc, C.c_void_p), 5, 2, 0, 1, 7, None, 1, f(a)) == 0
    b, _ = pkg.render(sc, 2, seed=7)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-6)
