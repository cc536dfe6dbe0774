"""oracle_py.py -- ctypes binding of the two CPU checkers (oracle/oracle_api.h).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

    Oracle("ref")   -> oracle/_ref/libjetpbrt_ref.so   (the unmodified reference, compiled in place)
    Oracle("port")  -> oracle/libjetpbrt_oracle.so     (our CPU restatement, oracle/pt_oracle.cc)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REF_LIB = _HERE / "_ref" / "libjetpbrt_ref.so"
PORT_LIB = _HERE / "libjetpbrt_oracle.so"


def have(kind: str) -> bool:
    return (REF_LIB if kind == "ref" else PORT_LIB).exists()


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(shape) if shape is not None else a


INTEGRATORS = {"path": 0, "path_recursive": 1, "whitted": 2, "debug": 3}  # main.cc:151-154


class Oracle:
    def __init__(self, kind: str):
        assert kind in ("ref", "port")
        self.kind = kind
        path = REF_LIB if kind == "ref" else PORT_LIB
        if not path.exists():
            raise FileNotFoundError(f"{path} missing: run `make -C {_HERE}`")
        self.lib = C.CDLL(str(path))
        self.pfx = "jref_" if kind == "ref" else "jorc_"
        P, I, F, IP = C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)
        g = self._fn
        g("scene_create").restype = P
        g("scene_create").argtypes = [P]
        g("scene_destroy").argtypes = [P]
        g("scene_destroy").restype = None
        g("intersect_shape").argtypes = [P, I, F, IP, F, F, F]
        g("scene_intersect").argtypes = [P, I, F, IP, F, F, F]
        g("scene_occluded").argtypes = [P, I, F, F, IP]
        g("bsdf").argtypes = [P, I, F, F, F, F, F, F, F, F, F, F, IP, IP]
        g("bsdf_ex").argtypes = [P, I, F, F, F, F, F, F, F, F, F, IP]
        g("light_sample").argtypes = [P, I, I, F, F, F, F, F, F, F]
        g("emitted").argtypes = [P, I, IP, F, F, F]
        g("generate_rays").argtypes = [P, I, F, F, F]
        g("render").argtypes = [P, I, I, I, F]
        g("render").restype = C.c_double
        g("render_mode").argtypes = [P, I, I, I, I, F]
        g("render_mode").restype = C.c_double
        g("scene_info").argtypes = [P, F]
        if kind == "ref":
            g("save_image").argtypes = [C.c_char_p, I, I, I, F]
            g("load_obj").argtypes = [C.c_char_p, I, I, F, C.c_float, F, F, I]
        if kind == "port":
            g("scene_intersect_brute").argtypes = [P, I, F, IP, F, F, F]
            g("render_counter").argtypes = [P, I, I, C.c_uint64, I, F, C.POINTER(C.c_uint64)]
            g("render_counter").restype = C.c_double
            g("render_counter_mode").argtypes = [P, I, I, I, C.c_uint64, I, F, C.POINTER(C.c_uint64)]
            g("render_counter_mode").restype = C.c_double
            g("render_counted").argtypes = [P, I, I, I, F, C.POINTER(C.c_uint64)]
            g("render_counted").restype = C.c_double
            g("philox_block").argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, F]
            g("philox_raw").argtypes = [C.POINTER(C.c_uint32)] * 3

    def _fn(self, name):
        return getattr(self.lib, self.pfx + name)

    def scene(self, host_scene) -> "OracleScene":
        return OracleScene(self, host_scene)

    def intersect_shape(self, shape, rays8):
        rays8 = _f32(rays8, (-1, 8))
        n = len(rays8)
        hit = np.empty(n, np.int32); t = np.empty(n, np.float32)
        pos = np.empty((n, 3), np.float32); nrm = np.empty((n, 3), np.float32)
        rc = self._fn("intersect_shape")(C.addressof(shape), n, _f(rays8), _i(hit), _f(t), _f(pos), _f(nrm))
        assert rc == 0
        return hit, t, pos, nrm

    def bsdf(self, mat, nrm3, wo3, wi3, u2, ulobe):
        nrm3 = _f32(nrm3, (-1, 3)); wo3 = _f32(wo3, (-1, 3)); wi3 = _f32(wi3, (-1, 3))
        u2 = _f32(u2, (-1, 2)); ulobe = _f32(ulobe, (-1,))
        n = len(nrm3)
        fe = np.empty((n, 3), np.float32); pe = np.empty(n, np.float32); swi = np.empty((n, 3), np.float32)
        sf = np.empty((n, 3), np.float32); sp = np.empty(n, np.float32); fl = np.empty(n, np.int32); dl = np.empty(n, np.int32)
        rc = self._fn("bsdf")(C.addressof(mat), n, _f(nrm3), _f(wo3), _f(wi3), _f(u2), _f(ulobe), _f(fe), _f(pe), _f(swi),
                              _f(sf), _f(sp), _i(fl), _i(dl))
        assert rc == 0
        return dict(f_eval=fe, pdf_eval=pe, s_wi=swi, s_f=sf, s_pdf=sp, s_flags=fl, is_delta=dl)

    def bsdf_ex(self, desc, nrm3, wo3, wi3, u2):
        """The BSDF classes no material builds (jpbrt_bsdf_desc): Evalf / Pdf / Sample in world space."""
        nrm3 = _f32(nrm3, (-1, 3)); wo3 = _f32(wo3, (-1, 3)); wi3 = _f32(wi3, (-1, 3)); u2 = _f32(u2, (-1, 2))
        n = len(nrm3)
        fe = np.empty((n, 3), np.float32); pe = np.empty(n, np.float32); swi = np.empty((n, 3), np.float32)
        sf = np.empty((n, 3), np.float32); sp = np.empty(n, np.float32); fl = np.empty(n, np.int32)
        rc = self._fn("bsdf_ex")(C.addressof(desc), n, _f(nrm3), _f(wo3), _f(wi3), _f(u2), _f(fe), _f(pe), _f(swi), _f(sf), _f(sp), _i(fl))
        assert rc == 0
        return dict(f_eval=fe, pdf_eval=pe, s_wi=swi, s_f=sf, s_pdf=sp, s_flags=fl)

    def philox_block(self, pixel, sample, block, seed):
        o = np.empty(4, np.float32)
        self._fn("philox_block")(pixel, sample, block, seed, _f(o))
        return o

    def save_image(self, basename: str, kind: int, film):
        """(ref only) FFilm::AddColor + FFilm::SaveAsImage of the unmodified reference."""
        film = _f32(film)
        h, w = film.shape[0], film.shape[1]
        return self._fn("save_image")(basename.encode(), kind, w, h, _f(film))

    def load_obj(self, filename: str, flip_normal=False, flip_handedness=False, offset=(0, 0, 0), scale=1.0, capacity=1 << 20):
        """(ref only) LoadTriangleMesh through the vendored obj_loader.h: (tris[n,3,3], normals[n,3]) or None on failure."""
        tris = np.zeros((capacity, 3, 3), np.float32); nrm = np.zeros((capacity, 3), np.float32)
        off = np.asarray(offset, np.float32)
        n = self._fn("load_obj")(filename.encode(), int(flip_normal), int(flip_handedness), _f(off), scale, _f(tris), _f(nrm), capacity)
        return None if n < 0 else (tris[:n].copy(), nrm[:n].copy())

    def philox_raw(self, ctr4, key2):
        ctr4 = np.ascontiguousarray(ctr4, np.uint32); key2 = np.ascontiguousarray(key2, np.uint32)
        o = np.empty(4, np.uint32)
        u32p = C.POINTER(C.c_uint32)
        self._fn("philox_raw")(ctr4.ctypes.data_as(u32p), key2.ctypes.data_as(u32p), o.ctypes.data_as(u32p))
        return o


class OracleScene:
    def __init__(self, oracle: Oracle, host_scene):
        self.o = oracle
        self.host_scene = host_scene  # keeps the description alive
        self.width = host_scene.d.camera.width
        self.height = host_scene.d.camera.height
        self.h = oracle._fn("scene_create")(C.cast(host_scene.desc, C.c_void_p))
        if not self.h:
            raise RuntimeError("oracle scene_create failed")

    def close(self):
        if self.h:
            self.o._fn("scene_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _isect(self, fn, rays8):
        rays8 = _f32(rays8, (-1, 8))
        n = len(rays8)
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32)
        pos = np.empty((n, 3), np.float32); nrm = np.empty((n, 3), np.float32)
        rc = self.o._fn(fn)(self.h, n, _f(rays8), _i(prim), _f(t), _f(pos), _f(nrm))
        assert rc == 0
        return prim, t, pos, nrm

    def intersect(self, rays8):
        return self._isect("scene_intersect", rays8)

    def intersect_brute(self, rays8):
        return self._isect("scene_intersect_brute", rays8)

    def occluded(self, pos3, target3):
        pos3 = _f32(pos3, (-1, 3)); target3 = _f32(target3, (-1, 3))
        n = len(pos3)
        occ = np.empty(n, np.int32)
        assert self.o._fn("scene_occluded")(self.h, n, _f(pos3), _f(target3), _i(occ)) == 0
        return occ

    def light_sample(self, light, pos3, nrm3, u2):
        pos3 = _f32(pos3, (-1, 3)); nrm3 = _f32(nrm3, (-1, 3)); u2 = _f32(u2, (-1, 2))
        n = len(pos3)
        lpos = np.empty((n, 3), np.float32); wi = np.empty((n, 3), np.float32)
        pdf = np.empty(n, np.float32); Li = np.empty((n, 3), np.float32)
        assert self.o._fn("light_sample")(self.h, light, n, _f(pos3), _f(nrm3), _f(u2), _f(lpos), _f(wi), _f(pdf), _f(Li)) == 0
        return lpos, wi, pdf, Li

    def emitted(self, prim, nrm3, wo3):
        prim = np.ascontiguousarray(prim, np.int32); nrm3 = _f32(nrm3, (-1, 3)); wo3 = _f32(wo3, (-1, 3))
        n = len(prim)
        Le = np.empty((n, 3), np.float32)
        assert self.o._fn("emitted")(self.h, n, _i(prim), _f(nrm3), _f(wo3), _f(Le)) == 0
        return Le

    def generate_rays(self, posfilm2):
        posfilm2 = _f32(posfilm2, (-1, 2))
        n = len(posfilm2)
        o = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32)
        assert self.o._fn("generate_rays")(self.h, n, _f(posfilm2), _f(o), _f(d)) == 0
        return o, d

    def render(self, spp, numthreads=8, seed=-1, mode=0):
        """FIntegrator::Render: returns (film[h,w,3] = Clamp01(mean), seconds).
        mode: 0 path (iteration), 1 path (recursive), 2 Whitted, 3 debug -- INTEGRATORS below."""
        film = np.empty((self.height, self.width, 3), np.float32)
        sec = self.o._fn("render_mode")(self.h, mode, spp, numthreads, seed, _f(film))
        assert sec >= 0
        return film, sec

    def render_counter(self, sample_begin, sample_count, seed, numthreads=8, counters=False, mode=0):
        """Port only: raw radiance sums with the B200 path's counter-based sampler."""
        film = np.empty((self.height, self.width, 3), np.float32)
        cnt = (C.c_uint64 * 7)()
        sec = self.o._fn("render_counter_mode")(self.h, mode, sample_begin, sample_count, seed, numthreads, _f(film), cnt if counters else None)
        assert sec >= 0
        if counters:
            keys = ["ext_rays", "shadow_rays", "node_tests", "prim_tests", "samples", "rng_draws", "vertices"]
            return film, sec, dict(zip(keys, list(cnt)))
        return film, sec

    def render_counted(self, spp, numthreads=8, seed=-1):
        film = np.empty((self.height, self.width, 3), np.float32)
        cnt = (C.c_uint64 * 7)()
        sec = self.o._fn("render_counted")(self.h, spp, numthreads, seed, _f(film), cnt)
        keys = ["ext_rays", "shadow_rays", "node_tests", "prim_tests", "samples", "rng_draws", "vertices"]
        return film, sec, dict(zip(keys, list(cnt)))

    def info(self):
        o = np.zeros(7, np.float32)
        self.o._fn("scene_info")(self.h, _f(o))
        return o
