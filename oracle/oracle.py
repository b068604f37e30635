"""oracle.py -- ctypes driver for the two CPU oracles.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package never does.

  * ``PortOracle``  -> oracle/liboracle_port.so : our restatement (oracle/port/port.h)
  * ``RefOracle``   -> oracle/_ref/libref_cb.so : the UNMODIFIED reference sources behind
                        oracle/ref_harness.cpp, with the shared counter-based RNG engine

Both expose the same methods over numpy arrays so tests can diff them call by call.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle_port.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_cb.so")
REF_STOCK_LIB = os.path.join(HERE, "_ref", "libref_stock.so")
REF_BIN = os.path.join(HERE, "_ref", "RayTracer_ref")
REF_BIN_OMP = os.path.join(HERE, "_ref", "RayTracer_ref_omp")
REF_MESHES = os.path.join(HERE, "_ref", "meshes")
REFERENCE_ROOT = "/root/reference"

DOMAIN_PIXEL = 1
DOMAIN_PHOTON = 2

_f = np.float32
_i = np.int32


def build(ref: bool | None = None) -> None:
    """Compile the oracles (make).  ``ref`` defaults to "when /root/reference exists"."""
    if ref is None:
        ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "source"))
    subprocess.run(["make", "-s", "port"] + (["ref"] if ref else []), cwd=HERE, check=True)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class FlatScene:
    """What crosses the Renderer::render seam, as numpy arrays."""

    MAGIC = b"RTSCENE1"

    def __init__(self, pos, nrm, tri, mesh_tri_off, mesh_vtx_off, mats, lights, lights_ctor, cam, w, h):
        self.pos = _c(pos, _f).reshape(-1, 3)
        self.nrm = _c(nrm, _f).reshape(-1, 3)
        self.tri = _c(tri, _i).reshape(-1, 3)
        self.mesh_tri_off = _c(mesh_tri_off, _i)
        self.mesh_vtx_off = _c(mesh_vtx_off, _i)
        self.mats = _c(mats, _f).reshape(-1, 8)
        self.lights = _c(lights, _f).reshape(-1, 21)
        self.lights_ctor = _c(lights_ctor, _f).reshape(-1, 11)
        self.cam = _c(cam, _f).reshape(12)
        self.w, self.h = int(w), int(h)

    V = property(lambda s: s.pos.shape[0])
    T = property(lambda s: s.tri.shape[0])
    M = property(lambda s: s.mats.shape[0])
    L = property(lambda s: s.lights.shape[0])

    def tri_mesh(self):
        out = np.zeros(self.T, _i)
        for m in range(self.M):
            out[self.mesh_tri_off[m]:self.mesh_tri_off[m + 1]] = m
        return out

    def save(self, path):
        with open(path, "wb") as f:
            f.write(self.MAGIC)
            np.array([self.V, self.T, self.M, self.L, self.w, self.h], _i).tofile(f)
            for a in (self.pos, self.nrm, self.tri, self.mesh_tri_off, self.mesh_vtx_off, self.mats, self.lights,
                      self.lights_ctor, self.cam):
                a.tofile(f)

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            assert f.read(8) == cls.MAGIC, "not an .rtscene file"
            V, T, M, L, w, h = np.fromfile(f, _i, 6)
            pos = np.fromfile(f, _f, 3 * V)
            nrm = np.fromfile(f, _f, 3 * V)
            tri = np.fromfile(f, _i, 3 * T)
            mto = np.fromfile(f, _i, M + 1)
            mvo = np.fromfile(f, _i, M + 1)
            mats = np.fromfile(f, _f, 8 * M)
            lights = np.fromfile(f, _f, 21 * L)
            lctor = np.fromfile(f, _f, 11 * L)
            cam = np.fromfile(f, _f, 12)
        return cls(pos, nrm, tri, mto, mvo, mats, lights, lctor, cam, w, h)


class _Oracle:
    prefix = ""

    def __init__(self, lib_path):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(f"{lib_path} missing: run oracle.build() / make -C oracle")
        self.lib = C.CDLL(lib_path)
        self.scene = None
        self.flat = None
        self._declare()

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _declare(self):
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        sigs = {
            "scene_destroy": (None, [vp]),
            "triangle_intersect": (None, [vp, i64, vp, vp]),
            "bsdf": (None, [vp, vp, i64, vp]),
            "light_eval": (None, [vp, i32, vp, i64, vp]),
            "light_sample": (None, [vp, i32, u64, u64, u64, i64, vp]),
            "jitter": (None, [u64, u64, u64, i64, i32, i32, vp]),
            "hsphere": (None, [u64, u64, u64, i64, vp, vp]),
            "rng_words": (None, [u64, u64, u64, i32, vp]),
            "rng_uniform_float": (None, [u64, u64, u64, i32, C.c_float, C.c_float, vp]),
            "rng_uniform_double": (None, [u64, u64, u64, i32, C.c_double, C.c_double, vp]),
            "camera_rays": (None, [vp, vp, vp, i64, vp]),
            "photon_map_create": (vp, [vp, i32, u64, i32, i32]),
            "photon_map_from_list": (vp, [vp, i64]),
            "photon_map_destroy": (None, [vp]),
            "photon_map_size": (i64, [vp]),
            "photon_map_get": (None, [vp, vp, vp]),
            "kdtree_layout": (None, [vp, vp, vp, vp, vp]),
            "render": (i32, [vp, i32, i32, i32, i32, u64, vp] + [i32] * 6 + [vp] * 4),
            "background": (None, [i32, i32, vp]),
            "composite": (None, [i32, i32, i32, vp, vp, vp]),
        }
        for name, (res, args) in sigs.items():
            fn = self._fn(name)
            fn.restype, fn.argtypes = res, args

    # ---- scene ------------------------------------------------------------------------------
    def close(self):
        if self.scene:
            self._fn("scene_destroy")(self.scene)
            self.scene = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pure functions -----------------------------------------------------------------------
    def triangle_intersect(self, packed15):
        a = _c(packed15, _f).reshape(-1, 15)
        flag = np.zeros(len(a), _i)
        uvt = np.zeros((len(a), 3), _f)
        self._fn("triangle_intersect")(_p(a), len(a), _p(flag), _p(uvt))
        return flag, uvt

    def bsdf(self, mat8, n_wi_wo):
        m = _c(mat8, _f)
        a = _c(n_wi_wo, _f).reshape(-1, 9)
        out = np.zeros((len(a), 3), _f)
        self._fn("bsdf")(_p(m), _p(a), len(a), _p(out))
        return out

    def light_eval(self, light, pts):
        a = _c(pts, _f).reshape(-1, 3)
        out = np.zeros_like(a)
        self._fn("light_eval")(self.scene, light, _p(a), len(a), _p(out))
        return out

    def light_sample(self, light, seed, domain, index0, n):
        out = np.zeros((n, 3), _f)
        self._fn("light_sample")(self.scene, light, seed, domain, index0, n, _p(out))
        return out

    def jitter(self, seed, domain, index0, n, sample, nsamples):
        out = np.zeros((n, 2), _f)
        self._fn("jitter")(seed, domain, index0, n, sample, nsamples, _p(out))
        return out

    def hsphere(self, seed, domain, index0, normals):
        a = _c(normals, _f).reshape(-1, 3)
        out = np.zeros_like(a)
        self._fn("hsphere")(seed, domain, index0, len(a), _p(a), _p(out))
        return out

    def rng_words(self, seed, domain, index, n):
        out = np.zeros(n, np.uint32)
        self._fn("rng_words")(seed, domain, index, n, _p(out))
        return out

    def rng_uniform_float(self, seed, domain, index, n, a, b):
        out = np.zeros(n, _f)
        self._fn("rng_uniform_float")(seed, domain, index, n, a, b, _p(out))
        return out

    def rng_uniform_double(self, seed, domain, index, n, a, b):
        out = np.zeros(n, np.float64)
        self._fn("rng_uniform_double")(seed, domain, index, n, a, b, _p(out))
        return out

    def camera_rays(self, xy, shift):
        xy = _c(xy, _i).reshape(-1, 2)
        sh = _c(shift, _f).reshape(-1, 2)
        out = np.zeros((len(xy), 6), _f)
        self._fn("camera_rays")(self.scene, _p(xy), _p(sh), len(xy), _p(out))
        return out

    # ---- photon map ---------------------------------------------------------------------------
    def photon_map_create(self, num_photons, seed, first_path=-1, num_paths=-1):
        return PhotonMapHandle(self, self._fn("photon_map_create")(self.scene, num_photons, seed, first_path,
                                                                    num_paths))

    def photon_map_from_list(self, particles7):
        a = _c(particles7, _f).reshape(-1, 7)
        return PhotonMapHandle(self, self._fn("photon_map_from_list")(_p(a), len(a)))

    # ---- render -------------------------------------------------------------------------------
    def render(self, N, mode, seed, num_photons=0, k=0, photon_map=None, window=None, samples=None,
               want_samples=False, threads=1):
        """Returns dict(sum_rgb[h,w,3], counter[h,w], samples[ns,h,w,3]|None, found|None) over the window."""
        w, h = self.flat.w, self.flat.h
        x0, y0, x1, y1 = window if window else (0, 0, w, h)
        s0, s1 = samples if samples else (0, N)
        ww, wh, ns = x1 - x0, y1 - y0, s1 - s0
        sum_rgb = np.zeros((wh, ww, 3), _f)
        counter = np.zeros((wh, ww), _i)
        smp = np.zeros((ns, wh, ww, 3), _f) if want_samples else None
        fnd = np.zeros((ns, wh, ww), np.int8) if want_samples else None
        pm = photon_map.handle if photon_map is not None else None
        args = [self.scene, N, mode, num_photons, k, seed, pm, x0, y0, x1, y1, s0, s1, _p(smp), _p(fnd), _p(sum_rgb),
                _p(counter)]
        if threads > 1 and self.prefix == "orc_":
            fn = self.lib.orc_render_mt
            fn.restype = C.c_int
            fn.argtypes = self._fn("render").argtypes + [C.c_int]
            rc = fn(*args, threads)
        else:
            rc = self._fn("render")(*args)
        if rc != 0:
            raise RuntimeError("oracle render failed (reference exception: empty tree or k too large)")
        return dict(sum_rgb=sum_rgb, counter=counter, samples=smp, found=fnd)

    def background(self, w, h):
        out = np.zeros((h, w, 3), _f)
        self._fn("background")(w, h, _p(out))
        return out

    def composite(self, N, sum_rgb, counter, background):
        h, w = counter.shape
        img = _c(background, _f).copy()
        self._fn("composite")(w, h, N, _p(_c(sum_rgb, _f)), _p(_c(counter, _i)), _p(img))
        return img


class PhotonMapHandle:
    def __init__(self, oracle, handle):
        self.o, self.handle = oracle, handle

    def __del__(self):
        try:
            if self.handle:
                self.o._fn("photon_map_destroy")(self.handle)
                self.handle = None
        except Exception:
            pass

    def size(self):
        return int(self.o._fn("photon_map_size")(self.handle))

    def get(self):
        n = self.size()
        p = np.zeros((n, 7), _f)
        hist = np.zeros(20, _i)
        self.o._fn("photon_map_get")(self.handle, _p(p), _p(hist))
        return p, hist

    def light_counts(self, num_lights=3):
        """(port only) particles stored per light by the emission that built this map."""
        out = np.zeros(num_lights, np.int64)
        fn = self.o.lib.orc_photon_map_light_counts
        fn.restype, fn.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int]
        fn(self.handle, _p(out), num_lights)
        return out

    def layout(self):
        n = self.size()
        nodes = np.zeros((n, 7), _f)
        left = np.zeros(n, _i)
        right = np.zeros(n, _i)
        root = np.zeros(1, _i)
        self.o._fn("kdtree_layout")(self.handle, _p(nodes), _p(left), _p(right), _p(root))
        return nodes, left, right, int(root[0])

    def knn(self, q3, k, want_index=False):
        q = _c(q3, _f).reshape(-1, 3)
        out = np.zeros((len(q), k, 7), _f)
        visited = np.zeros(len(q), np.int64)
        lib = self.o.lib
        if self.o.prefix == "orc_":
            idx = np.zeros((len(q), k), _i)
            fn = lib.orc_knn
            fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 3
            rc = fn(self.handle, _p(q), len(q), k, _p(out), _p(visited), _p(idx))
        else:
            idx = None
            fn = lib.ref_knn
            fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 2
            rc = fn(self.handle, _p(q), len(q), k, _p(out), _p(visited))
        if rc != 0:
            raise RuntimeError("knn failed (empty tree or k larger than the photon count)")
        return (out, visited, idx) if want_index else (out, visited)


class PortOracle(_Oracle):
    prefix = "orc_"

    def __init__(self, flat: FlatScene | None = None):
        super().__init__(PORT_LIB)
        lib = self.lib
        lib.orc_scene_from_flat.restype = C.c_void_p
        lib.orc_scene_from_flat.argtypes = [C.c_int] * 4 + [C.c_void_p] * 8 + [C.c_int] * 2
        lib.orc_trace.restype = None
        lib.orc_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 5
        lib.orc_trace_mt.restype = None
        lib.orc_trace_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 3 + [C.c_int]
        lib.orc_counters.restype = None
        lib.orc_counters.argtypes = [C.c_void_p, C.c_int]
        if flat is not None:
            self.set_scene(flat)

    def set_scene(self, flat: FlatScene):
        self.close()
        self.flat = flat
        self.scene = self.lib.orc_scene_from_flat(flat.V, flat.T, flat.M, flat.L, _p(flat.pos), _p(flat.nrm),
                                                  _p(flat.tri), _p(flat.mesh_tri_off), _p(flat.mesh_vtx_off),
                                                  _p(flat.mats), _p(flat.lights), _p(flat.cam), flat.w, flat.h)

    def trace(self, rays):
        r = _c(rays, _f).reshape(-1, 6)
        n = len(r)
        hit, mesh, tri3, uvd, tidx = np.zeros(n, _i), np.zeros(n, _i), np.zeros((n, 3), _i), np.zeros((n, 3), _f), \
            np.zeros(n, _i)
        self.lib.orc_trace(self.scene, _p(r), n, _p(hit), _p(mesh), _p(tri3), _p(uvd), _p(tidx))
        return dict(hit=hit, mesh=mesh, tri3=tri3, uvd=uvd, tri_index=tidx)

    def trace_mt(self, rays, threads):
        r = _c(rays, _f).reshape(-1, 6)
        n = len(r)
        hit, tidx, uvd = np.zeros(n, _i), np.zeros(n, _i), np.zeros((n, 3), _f)
        self.lib.orc_trace_mt(self.scene, _p(r), n, _p(hit), _p(tidx), _p(uvd), threads)
        return dict(hit=hit, tri_index=tidx, uvd=uvd)

    def counters(self, reset=False):
        out = np.zeros(3, np.uint64)
        self.lib.orc_counters(_p(out), 1 if reset else 0)
        return dict(rays=int(out[0]), queries=int(out[1]), visits=int(out[2]))


class RefOracle(_Oracle):
    prefix = "ref_"

    def __init__(self, stock_rng=False):
        """stock_rng=True loads the build whose `gen` is the reference's own serial minstd_rand0."""
        super().__init__(REF_STOCK_LIB if stock_rng else REF_LIB)
        lib = self.lib
        lib.ref_scene_create.restype = C.c_void_p
        lib.ref_scene_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        lib.ref_scene_from_flat.restype = C.c_void_p
        lib.ref_scene_from_flat.argtypes = [C.c_int] * 4 + [C.c_void_p] * 7 + [C.c_int] * 2
        lib.ref_scene_counts.restype = None
        lib.ref_scene_counts.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_scene_flatten.restype = None
        lib.ref_scene_flatten.argtypes = [C.c_void_p] * 10
        lib.ref_trace.restype = None
        lib.ref_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        lib.ref_load_off.restype = C.c_int
        lib.ref_load_off.argtypes = [C.c_char_p] + [C.c_void_p] * 4
        lib.ref_save_ppm.restype = None
        lib.ref_save_ppm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_char_p]

    def create_scene(self, w, h, custom_off=None, meshdir=None):
        """The reference's own scene assembly (Main.cpp:165-208); custom_off replaces cube_tri.off."""
        self.close()
        meshdir = meshdir or (os.path.join(REFERENCE_ROOT, "meshes")
                              if os.path.isdir(os.path.join(REFERENCE_ROOT, "meshes")) else REF_MESHES)
        self.scene = self.lib.ref_scene_create(meshdir.encode(), (custom_off or "").encode(), w, h)
        if not self.scene:
            raise RuntimeError("reference scene creation failed (OFF load)")
        self.flat = self.flatten(w, h)
        return self.flat

    def set_scene(self, flat: FlatScene):
        """Hand a flat scene to the reference classes (lights are rebuilt by the reference ctor)."""
        self.close()
        self.scene = self.lib.ref_scene_from_flat(flat.V, flat.T, flat.M, flat.L, _p(flat.pos), _p(flat.nrm),
                                                  _p(flat.tri), _p(flat.mesh_tri_off), _p(flat.mesh_vtx_off),
                                                  _p(flat.mats), _p(flat.lights_ctor), flat.w, flat.h)
        self.flat = flat

    def flatten(self, w, h):
        cnt = np.zeros(4, _i)
        self.lib.ref_scene_counts(self.scene, _p(cnt))
        V, T, M, L = (int(v) for v in cnt)
        pos, nrm, tri = np.zeros((V, 3), _f), np.zeros((V, 3), _f), np.zeros((T, 3), _i)
        mto, mvo = np.zeros(M + 1, _i), np.zeros(M + 1, _i)
        mats, lights, lctor, cam = np.zeros((M, 8), _f), np.zeros((L, 21), _f), np.zeros((L, 11), _f), np.zeros(12, _f)
        self.lib.ref_scene_flatten(self.scene, _p(pos), _p(nrm), _p(tri), _p(mto), _p(mvo), _p(mats), _p(lights),
                                   _p(lctor), _p(cam))
        return FlatScene(pos, nrm, tri, mto, mvo, mats, lights, lctor, cam, w, h)

    def trace(self, rays):
        r = _c(rays, _f).reshape(-1, 6)
        n = len(r)
        hit, mesh, tri3, uvd = np.zeros(n, _i), np.zeros(n, _i), np.zeros((n, 3), _i), np.zeros((n, 3), _f)
        self.lib.ref_trace(self.scene, _p(r), n, _p(hit), _p(mesh), _p(tri3), _p(uvd))
        return dict(hit=hit, mesh=mesh, tri3=tri3, uvd=uvd)

    def load_off(self, path):
        cnt = np.zeros(2, _i)
        if self.lib.ref_load_off(path.encode(), _p(cnt), None, None, None) != 0:
            raise RuntimeError("reference loadOFF threw")
        pos, nrm, tri = np.zeros((cnt[0], 3), _f), np.zeros((cnt[0], 3), _f), np.zeros((cnt[1], 3), _i)
        self.lib.ref_load_off(path.encode(), _p(cnt), _p(pos), _p(nrm), _p(tri))
        return pos, nrm, tri

    def save_pcd(self, particles7, path):
        """PhotonMap::saveToPCD (PhotonMap.h:59-84) on a particle list."""
        a = _c(particles7, _f).reshape(-1, 7)
        self.lib.ref_save_pcd.restype, self.lib.ref_save_pcd.argtypes = None, [C.c_void_p, C.c_int64, C.c_char_p]
        self.lib.ref_save_pcd(_p(a), len(a), path.encode())

    def save_ppm(self, rgb, path):
        h, w, _ = rgb.shape
        self.lib.ref_save_ppm(w, h, _p(_c(rgb, _f)), path.encode())
