"""scene_desc.py -- ctypes mirror of include/jetpbrt_scene.h and the host-side scene handle.

CUDA-free: importable (and usable, through libjetpbrt_host.so) in a process that must not load the product's CUDA
library -- `bench.py --impl reference` builds its scene description with it and renders with the reference alone.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
HOST_LIB_PATH = _HERE / "libjetpbrt_host.so"


# ---- include/jetpbrt_scene.h --------------------------------------------------------------------
class Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("front", C.c_float * 3), ("up", C.c_float * 3),
                ("vfov_deg", C.c_float), ("width", C.c_int), ("height", C.c_int)]


class Shape(C.Structure):
    _fields_ = [("type", C.c_int), ("flip_normal", C.c_int), ("p", (C.c_float * 3) * 4)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int), ("remap_roughness", C.c_int), ("a", C.c_float * 3), ("b", C.c_float * 3),
                ("f0", C.c_float), ("f1", C.c_float)]


class BsdfDesc(C.Structure):
    """jpbrt_bsdf_desc (include/jetpbrt_scene.h): the BSDF classes of the reference that no material builds."""
    _fields_ = [("kind", C.c_int), ("distribution", C.c_int), ("sample_visible_area", C.c_int), ("fresnel", C.c_int),
                ("color", C.c_float * 3), ("exponent", C.c_float), ("alphax", C.c_float), ("alphay", C.c_float),
                ("eta_a", C.c_float), ("eta_b", C.c_float), ("c_eta_i", C.c_float * 3), ("c_eta_t", C.c_float * 3), ("c_k", C.c_float * 3)]


BSDF_PHONG, BSDF_MICROFACET_REFLECTION, BSDF_MICROFACET_TRANSMISSION = 0, 1, 2
DIST_BECKMANN, DIST_TROWBRIDGE_REITZ = 0, 1
FRESNEL_NOOP, FRESNEL_DIELECTRIC, FRESNEL_CONDUCTOR = 0, 1, 2


class Light(C.Structure):
    _fields_ = [("type", C.c_int), ("shape", C.c_int), ("color", C.c_float * 3), ("pos", C.c_float * 3),
                ("dir", C.c_float * 3)]


class Primitive(C.Structure):
    _fields_ = [("shape", C.c_int), ("material", C.c_int), ("light", C.c_int)]


class SceneDesc(C.Structure):
    _fields_ = [("camera", Camera), ("max_depth", C.c_int), ("n_shapes", C.c_int), ("n_materials", C.c_int),
                ("n_lights", C.c_int), ("n_primitives", C.c_int),
                ("shapes", C.POINTER(Shape)), ("materials", C.POINTER(Material)), ("lights", C.POINTER(Light)),
                ("primitives", C.POINTER(Primitive)), ("name", C.c_char_p)]


SHAPE_TRIANGLE, SHAPE_RECTANGLE, SHAPE_SPHERE, SHAPE_DISK = 0, 1, 2, 3
UPLOAD_GPU_BVH = 1  # jpbrt_upload_scene_ex flag
MAT_MATTE, MAT_MIRROR, MAT_GLASS, MAT_PLASTIC, MAT_METAL = 0, 1, 2, 3, 4
LIGHT_ENVIRONMENT, LIGHT_AREA, LIGHT_POINT, LIGHT_DIRECTION = 0, 1, 2, 3



class JpbrtError(RuntimeError):
    pass


_scene_lib = None


def bind_scene_lib(lib):
    """Declare the scene entry points of `lib` (libjetpbrt_b200.so or libjetpbrt_host.so) and make HostScene use them."""
    global _scene_lib
    lib.jpbrt_scene_builtin.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_float]
    lib.jpbrt_scene_builtin.restype = C.c_void_p
    lib.jpbrt_scene_get_desc.argtypes = [C.c_void_p]
    lib.jpbrt_scene_get_desc.restype = C.POINTER(SceneDesc)
    lib.jpbrt_scene_free.argtypes = [C.c_void_p]
    lib.jpbrt_scene_free.restype = None
    _scene_lib = lib
    return lib


def load_host_lib():
    """libjetpbrt_host.so: scene description, built-in scenes, OBJ loader, film writers -- host C++ only, no CUDA."""
    if not HOST_LIB_PATH.exists():
        raise ImportError(f"{HOST_LIB_PATH} is missing: build it with `make -C {_HERE}`")
    return bind_scene_lib(C.CDLL(str(HOST_LIB_PATH)))


# ---- scenes --------------------------------------------------------------------------------------
class HostScene:
    """A scene description owned by the host library (jetpbrt::Scene) or built in Python."""

    def __init__(self, handle=None, desc=None, keepalive=None):
        self._handle = handle
        self._desc = desc
        self._keepalive = keepalive

    @classmethod
    def builtin(cls, name: str, width: int, height: int, scale: float = 1.0) -> "HostScene":
        h = _scene_lib.jpbrt_scene_builtin(name.encode(), width, height, scale)
        if not h:
            raise JpbrtError(f"unknown built-in scene {name!r}")
        return cls(handle=h, desc=_scene_lib.jpbrt_scene_get_desc(h))

    @classmethod
    def from_arrays(cls, camera: Camera, shapes, materials, lights, primitives, max_depth=5, name="scene"):
        """Build a description from Python lists of Shape/Material/Light/Primitive structs."""
        sa = (Shape * max(1, len(shapes)))(*shapes)
        ma = (Material * max(1, len(materials)))(*materials)
        la = (Light * max(1, len(lights)))(*lights)
        pa = (Primitive * max(1, len(primitives)))(*primitives)
        nm = name.encode()
        d = SceneDesc(camera, max_depth, len(shapes), len(materials), len(lights), len(primitives),
                      C.cast(sa, C.POINTER(Shape)), C.cast(ma, C.POINTER(Material)), C.cast(la, C.POINTER(Light)),
                      C.cast(pa, C.POINTER(Primitive)), nm)
        return cls(desc=C.pointer(d), keepalive=(sa, ma, la, pa, nm, d))

    @property
    def desc(self):
        return self._desc

    @property
    def d(self) -> SceneDesc:
        return self._desc.contents

    def set_max_depth(self, depth: int):
        self._desc.contents.max_depth = depth

    def set_resolution(self, w: int, h: int):
        self._desc.contents.camera.width = w
        self._desc.contents.camera.height = h

    def close(self):
        if self._handle:
            _scene_lib.jpbrt_scene_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


