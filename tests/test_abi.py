"""The drop-in boundary without a GPU: the library loads, exports every symbol the header declares,
builds the five BASELINE.json scenes, validates descriptions and fails loudly (never falls back to the
CPU) when no CUDA device exists."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(pkg):
    hdr = (ROOT / "include" / "jetpbrt_b200.h").read_text()
    declared = set(re.findall(r"\b(jpbrt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(pkg.EXPORTS), declared ^ set(pkg.EXPORTS)
    for name in declared:
        assert hasattr(pkg.lib, name), name
    assert "sm_100a" in pkg.version()


def test_struct_layouts_match_header(pkg):
    # sizes implied by include/jetpbrt_scene.h (4-byte fields, natural alignment)
    assert C.sizeof(pkg.Camera) == 48 and C.sizeof(pkg.Shape) == 56 and C.sizeof(pkg.Material) == 40
    assert C.sizeof(pkg.Light) == 44 and C.sizeof(pkg.Primitive) == 12
    assert C.sizeof(pkg.SceneDesc) == 48 + 5 * 4 + 4 + 5 * 8


@pytest.mark.parametrize("name,scale,prims,lights,depth", [
    ("cornell", 1.0, 32, 3, 5),             # 2 light + 6 + 10 + 10 + 2 + 2 triangles; env + 2 triangle lights
    ("bunny", 1.0, 2 + 4 * 5040, 2, 5),     # light rect + floor rect + 4 x 5,040-triangle stand-in mesh
    ("glossy", 1.0, 16 + 5 + 20 + 2, 17, 16),
    ("large", 0.02, 2 * 30 * 30 + 1 + 2 * 10 * 8 + 2 * 10, 2, 8),
])
def test_builtin_scenes(pkg, name, scale, prims, lights, depth):
    sc = pkg.HostScene.builtin(name, 64, 48, scale)
    d = sc.d
    assert (d.n_primitives, d.n_lights, d.max_depth) == (prims, lights, depth)
    assert (d.camera.width, d.camera.height) == (64, 48)
    assert d.lights[0].type == pkg.LIGHT_ENVIRONMENT  # the environment light is created first (main.cc:25,76)


def test_full_size_large_scene_has_five_million_triangles(pkg):
    sc = pkg.HostScene.builtin("large", 8, 8, 1.0)
    assert 4_990_000 <= sc.d.n_primitives <= 5_010_000


def test_cornell_light_radiance_matches_main_cc(pkg):
    sc = pkg.HostScene.builtin("cornell", 8, 8)
    l = sc.d.lights[1]
    f = np.float32
    a = np.array([f(0.747) + f(0.058), f(0.747) + f(0.258), f(0.747)], np.float32)
    b = np.array([f(0.740) + f(0.287), f(0.740) + f(0.160), f(0.740)], np.float32)
    c = np.array([f(0.737) + f(0.642), f(0.737) + f(0.159), f(0.737)], np.float32)
    want = (a * f(8.0) + b * f(15.6)) + c * f(18.4)
    assert np.array_equal(np.array(list(l.color), np.float32), want)
    assert l.type == pkg.LIGHT_AREA and sc.d.primitives[0].light == 1 and sc.d.primitives[1].light == 2


def test_no_gpu_means_error_not_fallback(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    sc = pkg.HostScene.builtin("cornell", 16, 16)
    with pytest.raises(pkg.JpbrtError, match="no CUDA device|no CPU fallback"):
        pkg.Context(sc)
    with pytest.raises(pkg.JpbrtError):
        pkg.render(sc, 1)
    with pytest.raises(pkg.JpbrtError):
        pkg.unit_rng_block([0], [0], [0], 1)


def test_invalid_descriptions_are_rejected(pkg):
    sc = pkg.HostScene.builtin("cornell", 16, 16)
    d = sc.d
    keep = d.n_primitives
    d.n_primitives = 0
    assert pkg.lib.jpbrt_debug_flatten(sc.desc, 0, None, 0) == -1  # JPBRT_ERR_INVALID
    assert b"no primitives" in pkg.lib.jpbrt_last_error(None)
    d.n_primitives = keep
    d.max_depth = 500
    assert pkg.lib.jpbrt_debug_flatten(sc.desc, 0, None, 0) == -1
    d.max_depth = 5
    bad = d.primitives[3].shape
    d.primitives[3].shape = 10**6
    assert pkg.lib.jpbrt_debug_flatten(sc.desc, 0, None, 0) == -1
    d.primitives[3].shape = bad
    assert pkg.lib.jpbrt_debug_flatten(sc.desc, 0, None, 0) > 0
    ctx = C.c_void_p()
    assert pkg.lib.jpbrt_upload_scene(None, 0, C.byref(ctx)) < 0 and not ctx.value
    assert pkg.lib.jpbrt_render_pass(None, 0, 1, 1) == -1
    assert pkg.lib.jpbrt_read_film(None, None, 1, 1) == -1


def test_save_image_formats(pkg, tmp_path):
    h, w = 6, 8  # w*3 is a multiple of 4: identical bytes to the reference's writer (film.cc:62-144)
    film = np.linspace(0, 1.2, h * w * 3, dtype=np.float32).reshape(h, w, 3)
    base = str(tmp_path / "img")
    for kind in (3, 1, 2):  # 3 = the well-formed text P3 (kind 0 reproduces the reference's raw-byte PPM: test_reference_io_pin.py)
        pkg.save_image(base, kind, film)
    bmp = (tmp_path / "img.bmp").read_bytes()
    assert bmp[:2] == b"BM" and len(bmp) == 54 + w * 3 * h
    assert int.from_bytes(bmp[18:22], "little") == w and int.from_bytes(bmp[22:26], "little") == h
    g = lambda x: int((min(max(x, 0.0), 1.0) ** np.float32(1 / 2.2)) * 255.0)  # noqa: E731  gamma_encoding, film.h:24
    last_row_first_px = film[h - 1, 0]  # BMP rows are bottom-up, pixels BGR
    assert list(bmp[54:57]) == [g(last_row_first_px[2]), g(last_row_first_px[1]), g(last_row_first_px[0])]
    ppm = (tmp_path / "img.ppm").read_text().split()
    assert ppm[:4] == ["P3", str(w), str(h), "255"] and len(ppm) == 4 + w * h * 3
    hdr = (tmp_path / "img.hdr").read_bytes()
    assert hdr.startswith(b"#?RADIANCE") and hdr.endswith(bytes(4)) is False
    assert len(hdr.split(b"\n", 4)[4]) == w * h * 4


def test_cli_binary_exists_and_prints_usage(pkg):
    """jet-pbrt_b200/jetpbrt keeps the reference's command line (main.cc:121-124): no arguments -> usage, exit 0."""
    import subprocess

    exe = ROOT / "jet-pbrt_b200" / "jetpbrt"
    assert exe.exists(), "run make -C jet-pbrt_b200"
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=30)
    assert r.returncode == 0 and "pbrt.exe  sceneid   spp" in r.stdout


def test_new_entry_points_validate_before_touching_a_device(pkg):
    """jpbrt_upload_scene_ex / jpbrt_unit_bsdf_ex / jpbrt_render_integrator reject bad arguments on any machine."""
    sc = pkg.HostScene.builtin("cornell", 16, 16)
    ctx = C.c_void_p()
    assert pkg.lib.jpbrt_upload_scene_ex(sc.desc, 0, 0x80, C.byref(ctx)) == -1 and not ctx.value  # unknown flag
    assert b"flags" in pkg.lib.jpbrt_last_error(None)
    assert pkg.lib.jpbrt_upload_scene_ex(sc.desc, 0, 0, None) == -1
    d = pkg.BsdfDesc()
    d.kind = 7
    z = np.zeros(3, np.float32)
    f = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    o3, o1, oi = np.zeros(3, np.float32), np.zeros(1, np.float32), np.zeros(1, np.int32)
    rc = pkg.lib.jpbrt_unit_bsdf_ex(C.byref(d), 0, 1, f(z), f(z), f(z), f(z), f(o3), f(o1), f(o3), f(o3), f(o1), oi.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == -1 and b"BSDF" in pkg.lib.jpbrt_last_error(None)
    out = np.zeros((16, 16, 3), np.float32)
    assert pkg.lib.jpbrt_render_integrator(sc.desc, 0, 0, 1, 0, f(out), None) == -1  # spp must be positive
    assert set(pkg.INTEGRATORS) == {"path", "path_recursive", "whitted", "debug"}
