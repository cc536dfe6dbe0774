"""jet-pbrt_b200 -- thin ctypes binding of libjetpbrt_b200.so (include/jetpbrt_b200.h).

This package is plumbing for tests, bench.py and multi-GPU launches (torch.distributed); the
product is the shared library: host C++ (scene description, BVH build) + hand-written sm_100a
CUDA (wavefront path tracer).  Nothing here computes: if the library is missing the import
fails loudly, and every compute call fails with an exception when no B200 is present.

The directory name has a hyphen (the reference's name); import it through
``__graft_entry__.load_package()`` which registers it as ``jet_pbrt_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libjetpbrt_b200.so"


from .scene_desc import *  # noqa: F401,F403  (ctypes mirror of include/jetpbrt_scene.h + HostScene)
from . import scene_desc as _scene_desc


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("extension_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shaded_vertices", C.c_uint64), ("box_tests", C.c_uint64), ("prim_tests", C.c_uint64),
                ("shadow_box_tests", C.c_uint64), ("shadow_prim_tests", C.c_uint64),
                ("invalid_contributions", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("ms_generate", C.c_double), ("ms_extend", C.c_double), ("ms_shade", C.c_double),
                ("ms_connect", C.c_double), ("ms_finalize", C.c_double),
                ("n_nodes", C.c_uint64), ("n_prim_slots", C.c_uint64), ("scene_bytes", C.c_uint64),
                ("bvh_build_seconds", C.c_double), ("paths_in_flight", C.c_uint64), ("bvh_builder", C.c_uint64),
                ("bvh_device_seconds", C.c_double), ("dropped_rays", C.c_uint64), ("stack_overflows", C.c_uint64),
                ("nee_dropped", C.c_uint64), ("bvh_depth", C.c_uint64), ("ms_reduce", C.c_double),
                ("node_fetches", C.c_uint64), ("prim_fetches", C.c_uint64), ("shadow_node_fetches", C.c_uint64),
                ("shadow_prim_fetches", C.c_uint64), ("node_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# Every symbol include/jetpbrt_b200.h declares (tests check the library exports them all).
EXPORTS = [
    "jpbrt_upload_scene", "jpbrt_upload_scene_ex", "jpbrt_render_pass", "jpbrt_read_film", "jpbrt_clear_film", "jpbrt_reset_stats", "jpbrt_destroy",
    "jpbrt_last_error", "jpbrt_render", "jpbrt_render_integrator", "jpbrt_film_device_ptr", "jpbrt_film_num_floats", "jpbrt_stream",
    "jpbrt_synchronize", "jpbrt_finalize_film_device", "jpbrt_reupload_scene", "jpbrt_set_option",
    "jpbrt_get_stats", "jpbrt_unit_intersect_shape", "jpbrt_unit_scene_intersect", "jpbrt_unit_scene_occluded",
    "jpbrt_unit_bsdf", "jpbrt_unit_bsdf_ex", "jpbrt_unit_light_sample", "jpbrt_unit_emitted", "jpbrt_unit_generate_rays",
    "jpbrt_unit_rng_block", "jpbrt_unit_philox_raw", "jpbrt_scene_info", "jpbrt_scene_builtin", "jpbrt_scene_get_desc",
    "jpbrt_comm_unique_id", "jpbrt_comm_init", "jpbrt_comm_rank", "jpbrt_comm_size", "jpbrt_reduce_film", "jpbrt_sample_partition",
    "jpbrt_render_multi", "jpbrt_load_obj_triangles", "jpbrt_scene_free", "jpbrt_save_image", "jpbrt_version", "jpbrt_device_count", "jpbrt_debug_flatten", "jpbrt_debug_ctx_table",
]


def _load():
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}` "
                          "(or __graft_entry__.build()); there is no Python/CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    P, I, F = C.c_void_p, C.c_int, C.POINTER(C.c_float)
    IP = C.POINTER(C.c_int)
    lib.jpbrt_version.restype = C.c_char_p
    lib.jpbrt_last_error.restype = C.c_char_p
    lib.jpbrt_last_error.argtypes = [P]
    lib.jpbrt_upload_scene.argtypes = [C.POINTER(SceneDesc), I, C.POINTER(P)]
    lib.jpbrt_upload_scene_ex.argtypes = [C.POINTER(SceneDesc), I, C.c_uint, C.POINTER(P)]
    lib.jpbrt_render_pass.argtypes = [P, I, I, C.c_uint64]
    lib.jpbrt_read_film.argtypes = [P, F, I, I]
    lib.jpbrt_clear_film.argtypes = [P]
    lib.jpbrt_reset_stats.argtypes = [P]
    lib.jpbrt_destroy.argtypes = [P]
    lib.jpbrt_destroy.restype = None
    lib.jpbrt_render.argtypes = [C.POINTER(SceneDesc), I, C.c_uint64, I, F, C.POINTER(C.c_double)]
    lib.jpbrt_render_integrator.argtypes = [C.POINTER(SceneDesc), I, I, C.c_uint64, I, F, C.POINTER(C.c_double)]
    lib.jpbrt_film_device_ptr.argtypes = [P]
    lib.jpbrt_film_device_ptr.restype = P
    lib.jpbrt_film_num_floats.argtypes = [P]
    lib.jpbrt_film_num_floats.restype = C.c_size_t
    lib.jpbrt_stream.argtypes = [P]
    lib.jpbrt_stream.restype = P
    lib.jpbrt_synchronize.argtypes = [P]
    lib.jpbrt_finalize_film_device.argtypes = [P, P, I]
    lib.jpbrt_reupload_scene.argtypes = [P, C.POINTER(C.c_size_t)]
    lib.jpbrt_set_option.argtypes = [P, C.c_char_p, C.c_longlong]
    lib.jpbrt_get_stats.argtypes = [P, C.POINTER(Stats)]
    lib.jpbrt_unit_intersect_shape.argtypes = [C.POINTER(Shape), I, I, F, IP, F, F, F]
    lib.jpbrt_unit_scene_intersect.argtypes = [P, I, F, IP, F, F, F]
    lib.jpbrt_unit_scene_occluded.argtypes = [P, I, F, F, IP]
    lib.jpbrt_unit_bsdf.argtypes = [C.POINTER(Material), I, I, F, F, F, F, F, F, F, F, F, F, IP, IP]
    lib.jpbrt_unit_bsdf_ex.argtypes = [C.POINTER(BsdfDesc), I, I, F, F, F, F, F, F, F, F, F, IP]
    lib.jpbrt_unit_light_sample.argtypes = [P, I, I, F, F, F, F, F, F, F]
    lib.jpbrt_unit_emitted.argtypes = [P, I, IP, F, F, F]
    lib.jpbrt_unit_generate_rays.argtypes = [P, I, F, F, F]
    lib.jpbrt_unit_rng_block.argtypes = [I, I, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                         C.c_uint64, F]
    lib.jpbrt_scene_info.argtypes = [P, F]
    lib.jpbrt_scene_builtin.argtypes = [C.c_char_p, I, I, C.c_float]
    lib.jpbrt_scene_builtin.restype = P
    lib.jpbrt_scene_get_desc.argtypes = [P]
    lib.jpbrt_scene_get_desc.restype = C.POINTER(SceneDesc)
    lib.jpbrt_scene_free.argtypes = [P]
    lib.jpbrt_scene_free.restype = None
    lib.jpbrt_save_image.argtypes = [C.c_char_p, I, I, I, F]
    lib.jpbrt_device_count.restype = I
    lib.jpbrt_debug_flatten.argtypes = [C.POINTER(SceneDesc), I, P, C.c_longlong]
    lib.jpbrt_debug_flatten.restype = C.c_longlong
    lib.jpbrt_debug_ctx_table.argtypes = [P, I, P, C.c_longlong]
    lib.jpbrt_debug_ctx_table.restype = C.c_longlong
    if not hasattr(lib, "jpbrt_render_multi"):  # a round-1 build loaded for an A/B timing (scripts/ab_variants.sh): no round-2 entry points
        return lib
    lib.jpbrt_unit_philox_raw.argtypes = [I, I, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.jpbrt_comm_unique_id.argtypes = [P, C.c_size_t]
    lib.jpbrt_comm_init.argtypes = [P, P, C.c_size_t, I, I]
    lib.jpbrt_comm_rank.argtypes = [P]
    lib.jpbrt_comm_size.argtypes = [P]
    lib.jpbrt_reduce_film.argtypes = [P]
    lib.jpbrt_sample_partition.argtypes = [I, I, I, IP, IP]
    lib.jpbrt_sample_partition.restype = None
    lib.jpbrt_render_multi.argtypes = [C.POINTER(SceneDesc), I, I, C.c_uint64, I, F, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.jpbrt_load_obj_triangles.argtypes = [C.c_char_p, I, F, C.c_float, F, C.c_longlong]
    lib.jpbrt_load_obj_triangles.restype = C.c_longlong
    return lib


lib = _load()
_scene_desc.bind_scene_lib(lib)  # HostScene's built-in scenes come from this library (the same host code as libjetpbrt_host.so)


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _check(rc, ctx=None):
    if rc != 0:
        msg = lib.jpbrt_last_error(ctx)
        raise JpbrtError(f"jetpbrt_b200 error {rc}: {msg.decode() if msg else '?'}")


# ---- device context --------------------------------------------------------------------------------
class Context:
    """jpbrt_ctx: a scene uploaded to one GPU plus its film and wavefront buffers."""

    def __init__(self, scene: HostScene, device: int = 0, gpu_bvh: bool | None = None):
        """gpu_bvh: True = build the BVH on the device (JPBRT_UPLOAD_GPU_BVH), False = host binned SAH,
        None = the library default (host, unless JPBRT_BVH_BUILDER=gpu is set)."""
        self._ctx = C.c_void_p()
        self.scene = scene
        if gpu_bvh is None:
            rc = lib.jpbrt_upload_scene(scene.desc, device, C.byref(self._ctx))
        else:
            rc = lib.jpbrt_upload_scene_ex(scene.desc, device, UPLOAD_GPU_BVH if gpu_bvh else 0, C.byref(self._ctx))
        _check(rc, None)
        self.width = scene.d.camera.width
        self.height = scene.d.camera.height
        self.device = device

    # the render trio
    def render_pass(self, sample_begin: int, sample_count: int, seed: int = 1234):
        _check(lib.jpbrt_render_pass(self._ctx, sample_begin, sample_count, seed), self._ctx)

    def read_film(self, spp_total: int = 1, finalize: bool = True, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.float32)
        _check(lib.jpbrt_read_film(self._ctx, _f(out), spp_total, 1 if finalize else 0), self._ctx)
        return out

    def clear_film(self):
        _check(lib.jpbrt_clear_film(self._ctx), self._ctx)

    def reset_stats(self):
        _check(lib.jpbrt_reset_stats(self._ctx), self._ctx)

    def synchronize(self):
        _check(lib.jpbrt_synchronize(self._ctx), self._ctx)

    def set_option(self, name: str, value: int):
        _check(lib.jpbrt_set_option(self._ctx, name.encode(), value), self._ctx)

    def stats(self) -> dict:
        s = Stats()
        _check(lib.jpbrt_get_stats(self._ctx, C.byref(s)), self._ctx)
        return s.as_dict()

    # multi-GPU (one process per GPU): NCCL communicator inside the library
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        _load_torch_nccl_first()
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
        _check(lib.jpbrt_comm_init(self._ctx, buf, COMM_ID_BYTES, rank, nranks), self._ctx)

    def reduce_film(self):
        _check(lib.jpbrt_reduce_film(self._ctx), self._ctx)

    def read_film_root(self, spp_total: int, out_ptr=None, finalize: bool = True):
        """jpbrt_read_film on a context with a communicator: reduce, then rank 0 finalizes and reads into host memory at
        `out_ptr` (ranks != 0 pass None)."""
        _check(lib.jpbrt_read_film(self._ctx, C.cast(out_ptr, C.POINTER(C.c_float)) if out_ptr else None, spp_total, 1 if finalize else 0), self._ctx)

    def reupload_scene(self) -> int:
        n = C.c_size_t(0)
        _check(lib.jpbrt_reupload_scene(self._ctx, C.byref(n)), self._ctx)
        return n.value

    def film_device_ptr(self) -> int:
        return lib.jpbrt_film_device_ptr(self._ctx)

    def film_num_floats(self) -> int:
        return lib.jpbrt_film_num_floats(self._ctx)

    def stream(self) -> int:
        return lib.jpbrt_stream(self._ctx) or 0

    def finalize_film_device(self, out_ptr: int, spp_total: int):
        _check(lib.jpbrt_finalize_film_device(self._ctx, out_ptr, spp_total), self._ctx)

    def film_tensor(self):
        """The raw-sum device film as a torch tensor aliasing the context's buffer (for NCCL)."""
        import torch

        n = self.film_num_floats()
        ptr = self.film_device_ptr()

        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

        return torch.as_tensor(_Arr(), device=f"cuda:{self.device}")

    def table(self, name: str) -> np.ndarray:
        """The flattened table `name` (nodes, slots, slot_nrm, slot_ml, prim_slot) of the tree this context's builder produced."""
        what, dtype = _TABLES[name]
        n = lib.jpbrt_debug_ctx_table(self._ctx, what, None, 0)
        _check(min(n, 0), self._ctx)
        out = np.empty(n, dtype=dtype)
        lib.jpbrt_debug_ctx_table(self._ctx, what, out.ctypes.data_as(C.c_void_p), n)
        return out

    def scene_info(self) -> np.ndarray:
        o = np.zeros(7, dtype=np.float32)
        _check(lib.jpbrt_scene_info(self._ctx, _f(o)), self._ctx)
        return o

    # unit kernels
    def unit_scene_intersect(self, rays8):
        rays8 = _f32(rays8, (-1, 8))
        n = len(rays8)
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32)
        pos = np.empty((n, 3), np.float32); nrm = np.empty((n, 3), np.float32)
        _check(lib.jpbrt_unit_scene_intersect(self._ctx, n, _f(rays8), _i(prim), _f(t), _f(pos), _f(nrm)), self._ctx)
        return prim, t, pos, nrm

    def unit_scene_occluded(self, pos3, target3):
        pos3 = _f32(pos3, (-1, 3)); target3 = _f32(target3, (-1, 3))
        n = len(pos3)
        occ = np.empty(n, np.int32)
        _check(lib.jpbrt_unit_scene_occluded(self._ctx, n, _f(pos3), _f(target3), _i(occ)), self._ctx)
        return occ

    def unit_light_sample(self, light, pos3, nrm3, u2):
        pos3 = _f32(pos3, (-1, 3)); nrm3 = _f32(nrm3, (-1, 3)); u2 = _f32(u2, (-1, 2))
        n = len(pos3)
        lpos = np.empty((n, 3), np.float32); wi = np.empty((n, 3), np.float32)
        pdf = np.empty(n, np.float32); Li = np.empty((n, 3), np.float32)
        _check(lib.jpbrt_unit_light_sample(self._ctx, light, n, _f(pos3), _f(nrm3), _f(u2), _f(lpos), _f(wi), _f(pdf), _f(Li)), self._ctx)
        return lpos, wi, pdf, Li

    def unit_emitted(self, prim, nrm3, wo3):
        prim = np.ascontiguousarray(prim, np.int32); nrm3 = _f32(nrm3, (-1, 3)); wo3 = _f32(wo3, (-1, 3))
        n = len(prim)
        Le = np.empty((n, 3), np.float32)
        _check(lib.jpbrt_unit_emitted(self._ctx, n, _i(prim), _f(nrm3), _f(wo3), _f(Le)), self._ctx)
        return Le

    def unit_generate_rays(self, posfilm2):
        posfilm2 = _f32(posfilm2, (-1, 2))
        n = len(posfilm2)
        o = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32)
        _check(lib.jpbrt_unit_generate_rays(self._ctx, n, _f(posfilm2), _f(o), _f(d)), self._ctx)
        return o, d

    def close(self):
        if self._ctx:
            lib.jpbrt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unit_intersect_shape(shape: Shape, rays8, device: int = 0):
    rays8 = _f32(rays8, (-1, 8))
    n = len(rays8)
    hit = np.empty(n, np.int32); t = np.empty(n, np.float32)
    pos = np.empty((n, 3), np.float32); nrm = np.empty((n, 3), np.float32)
    _check(lib.jpbrt_unit_intersect_shape(C.byref(shape), device, n, _f(rays8), _i(hit), _f(t), _f(pos), _f(nrm)))
    return hit, t, pos, nrm


def unit_bsdf(mat: Material, nrm3, wo3, wi3, u2, ulobe, device: int = 0):
    nrm3 = _f32(nrm3, (-1, 3)); wo3 = _f32(wo3, (-1, 3)); wi3 = _f32(wi3, (-1, 3)); u2 = _f32(u2, (-1, 2)); ulobe = _f32(ulobe, (-1,))
    n = len(nrm3)
    fe = np.empty((n, 3), np.float32); pe = np.empty(n, np.float32); swi = np.empty((n, 3), np.float32)
    sf = np.empty((n, 3), np.float32); sp = np.empty(n, np.float32); fl = np.empty(n, np.int32); dl = np.empty(n, np.int32)
    _check(lib.jpbrt_unit_bsdf(C.byref(mat), device, n, _f(nrm3), _f(wo3), _f(wi3), _f(u2), _f(ulobe), _f(fe), _f(pe),
                               _f(swi), _f(sf), _f(sp), _i(fl), _i(dl)))
    return dict(f_eval=fe, pdf_eval=pe, s_wi=swi, s_f=sf, s_pdf=sp, s_flags=fl, is_delta=dl)


def unit_bsdf_ex(desc: BsdfDesc, nrm3, wo3, wi3, u2, device: int = 0):
    """jpbrt_unit_bsdf_ex: Evalf / Pdf / Sample of the BSDF classes no material builds."""
    nrm3 = _f32(nrm3, (-1, 3)); wo3 = _f32(wo3, (-1, 3)); wi3 = _f32(wi3, (-1, 3)); u2 = _f32(u2, (-1, 2))
    n = len(nrm3)
    fe = np.empty((n, 3), np.float32); pe = np.empty(n, np.float32); swi = np.empty((n, 3), np.float32)
    sf = np.empty((n, 3), np.float32); sp = np.empty(n, np.float32); fl = np.empty(n, np.int32)
    _check(lib.jpbrt_unit_bsdf_ex(C.byref(desc), device, n, _f(nrm3), _f(wo3), _f(wi3), _f(u2), _f(fe), _f(pe), _f(swi), _f(sf), _f(sp), _i(fl)))
    return dict(f_eval=fe, pdf_eval=pe, s_wi=swi, s_f=sf, s_pdf=sp, s_flags=fl)


def unit_rng_block(pixel, sample, block, seed: int, device: int = 0):
    pixel = np.ascontiguousarray(pixel, np.uint32); sample = np.ascontiguousarray(sample, np.uint32)
    block = np.ascontiguousarray(block, np.uint32)
    n = len(pixel)
    out = np.empty((n, 4), np.float32)
    u32p = C.POINTER(C.c_uint32)
    _check(lib.jpbrt_unit_rng_block(device, n, pixel.ctypes.data_as(u32p), sample.ctypes.data_as(u32p),
                                    block.ctypes.data_as(u32p), seed, _f(out)))
    return out


def unit_philox_raw(ctr4, key2, device: int = 0):
    """Philox4x32-10 on (counter[4], key[2]) rows -> uint32[n, 4] (known-answer tests)."""
    ctr4 = np.ascontiguousarray(ctr4, np.uint32).reshape(-1, 4); key2 = np.ascontiguousarray(key2, np.uint32).reshape(-1, 2)
    out = np.empty_like(ctr4)
    u32p = C.POINTER(C.c_uint32)
    _check(lib.jpbrt_unit_philox_raw(device, len(ctr4), ctr4.ctypes.data_as(u32p), key2.ctypes.data_as(u32p), out.ctypes.data_as(u32p)))
    return out


# jpbrt_integrator (include/jetpbrt_b200.h): the reference's FIntegrator subclasses (main.cc:151-154)
INTEGRATORS = {"path": 0, "path_recursive": 1, "whitted": 2, "debug": 3}


def render(scene: HostScene, spp: int, seed: int = 1234, device: int = 0, integrator: str = "path"):
    """FIntegrator::Render equivalent: returns (film[h,w,3] = clamp01(mean), seconds)."""
    out = np.empty((scene.d.camera.height, scene.d.camera.width, 3), dtype=np.float32)
    sec = C.c_double(0)
    _check(lib.jpbrt_render_integrator(scene.desc, INTEGRATORS[integrator], spp, seed, device, _f(out), C.byref(sec)))
    return out, sec.value


COMM_ID_BYTES = 128


def _load_torch_nccl_first():
    """The library binds NCCL at run time (dlopen libnccl.so.2) and a process gets ONE libnccl.so.2: the first loaded.
    torch ships a newer NCCL than the system's and fails to import against the older one, so when torch is installed it
    must load (its) NCCL before the library's first communicator call does."""
    try:
        import torch  # noqa: F401
    except ImportError:
        pass


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); hand the bytes to the other ranks with any transport."""
    _load_torch_nccl_first()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib.jpbrt_comm_unique_id(buf, COMM_ID_BYTES))
    return buf.raw


def sample_partition(spp_total: int, rank: int, nranks: int):
    b, n = C.c_int(0), C.c_int(0)
    lib.jpbrt_sample_partition(spp_total, rank, nranks, C.byref(b), C.byref(n))
    return b.value, n.value


def render_multi(scene: HostScene, spp: int, ngpus: int, seed: int = 1234, integrator: str = "path"):
    """FIntegrator::Render with `ngpus` devices of this process: returns (film, seconds, reduce_ms)."""
    _load_torch_nccl_first()
    out = np.empty((scene.d.camera.height, scene.d.camera.width, 3), dtype=np.float32)
    sec, red = C.c_double(0), C.c_double(0)
    _check(lib.jpbrt_render_multi(scene.desc, INTEGRATORS[integrator], spp, seed, ngpus, _f(out), C.byref(sec), C.byref(red)))
    return out, sec.value, red.value


def save_image(basename: str, kind: int, film: np.ndarray):
    film = _f32(film)
    h, w = film.shape[0], film.shape[1]
    _check(lib.jpbrt_save_image(basename.encode(), kind, w, h, _f(film)))


def load_obj_triangles(filename: str, flip_handedness=False, offset=(0, 0, 0), scale=1.0):
    """jpbrt_load_obj_triangles: tris[n, 3, 3] as the reference's LoadTriangleMesh delivers them, or None on failure."""
    off = np.asarray(offset, np.float32)
    n = lib.jpbrt_load_obj_triangles(filename.encode(), int(flip_handedness), _f(off), scale, None, 0)
    if n < 0:
        return None
    out = np.empty((n, 3, 3), np.float32)
    lib.jpbrt_load_obj_triangles(filename.encode(), int(flip_handedness), _f(off), scale, _f(out), n)
    return out


def device_count() -> int:
    return lib.jpbrt_device_count()


_TABLES = {"nodes": (0, np.float32), "slots": (1, np.float32), "slot_nrm": (2, np.float32), "slot_ml": (3, np.int32),
           "prim_slot": (4, np.int32), "materials": (5, np.float32), "lights": (6, np.float32), "qnodes": (7, np.uint32),
           "qgrid": (8, np.float32)}


def debug_flatten(scene: HostScene, table: str) -> np.ndarray:
    """Host-only copy of one flattened table (see jpbrt_debug_flatten)."""
    what, dt = _TABLES[table]
    n = lib.jpbrt_debug_flatten(scene.desc, what, None, 0)
    if n < 0:
        _check(int(n))
    out = np.empty(int(n), dtype=dt)
    lib.jpbrt_debug_flatten(scene.desc, what, out.ctypes.data_as(C.c_void_p), n)
    return out


def version() -> str:
    return lib.jpbrt_version().decode()
