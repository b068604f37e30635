"""ray-tracing-engine_b200 -- B200-native render hot path of nikitakaraevv/ray-tracing-engine.

Host-side mirror (Python) of the reference interface for the ONE path this package replaces:

    reference (C++)                                  here
    ------------------------------------------------ ---------------------------------------------
    Scene  (source/Scene.h)                          Scene      flat arrays, `.rtscene` I/O
    Image  (source/Image.h, Image.cpp)               Image      fillBackground(), savePPM()
    Renderer(scene,numRays,mode,rayTracer[,p,k])     Renderer   same argument order and meaning
      .render(image)   (source/Renderer.cpp:203)       .render(image)   -> rt_render (CUDA)
    RayTracer::rayTrace (source/RayTracer.h:27)      Renderer.rayTrace(rays)  batch parity hook
    kdtree::knearest (source/kdtree.h:180)           Renderer.knearest(points, k)

All computation happens in lib/librt_b200.so (hand-written sm_100a kernels behind the C ABI of
include/rt_b200.h).  There is no CPU fallback: without the library or without a GPU the calls raise.
The C++ host (CLI, OFF loader, scene assembly) lives in host/ and is built into bin/RayTracer.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import RT_FLAG_BRUTE_FORCE, RT_FLAG_KNN_EXACT, RtError, rt_params, rt_stats  # noqa: F401

__all__ = ["Scene", "Image", "Renderer", "RtError", "RAYTRACE", "PATHTRACE"]

RAYTRACE, PATHTRACE = 0, 1  # source/Renderer.h:11-12


class Scene:
    """Scene::camera()/lightsources()/meshes() flattened the way they cross the render seam.

    pos, nrm: [V,3] float32; tri: [T,3] int32 GLOBAL vertex ids in scene (mesh, triangle) order;
    mesh_tri_off / mesh_vtx_off: [M+1]; mats: [M,8] {kd, alpha, albedo, F0}; lights: [L,21]
    {position, color, normal, vertical, horizontal, intensity, side, ac, al, aq, factor};
    cam: [12] {position, lowerLeft, horizontal, vertical} for aspect w/h.
    """

    MAGIC = b"RTSCENE1"

    def __init__(self, pos, nrm, tri, mesh_tri_off, mesh_vtx_off, mats, lights, cam, w, h, lights_ctor=None):
        f, i = np.float32, np.int32
        self.pos = np.ascontiguousarray(pos, f).reshape(-1, 3)
        self.nrm = np.ascontiguousarray(nrm, f).reshape(-1, 3)
        self.tri = np.ascontiguousarray(tri, i).reshape(-1, 3)
        self.mesh_tri_off = np.ascontiguousarray(mesh_tri_off, i)
        self.mesh_vtx_off = np.ascontiguousarray(mesh_vtx_off, i)
        self.mats = np.ascontiguousarray(mats, f).reshape(-1, 8)
        self.lights = np.ascontiguousarray(lights, f).reshape(-1, 21)
        self.cam = np.ascontiguousarray(cam, f).reshape(12)
        self.lights_ctor = (np.ascontiguousarray(lights_ctor, f).reshape(-1, 11) if lights_ctor is not None else
                            np.zeros((len(self.lights), 11), f))
        self.w, self.h = int(w), int(h)

    V = property(lambda s: s.pos.shape[0])
    T = property(lambda s: s.tri.shape[0])
    M = property(lambda s: s.mats.shape[0])
    L = property(lambda s: s.lights.shape[0])

    @classmethod
    def load(cls, path):
        with open(path, "rb") as fh:
            if fh.read(8) != cls.MAGIC:
                raise ValueError(f"{path}: not an .rtscene file")
            V, T, M, L, w, h = np.fromfile(fh, np.int32, 6)
            pos = np.fromfile(fh, np.float32, 3 * V)
            nrm = np.fromfile(fh, np.float32, 3 * V)
            tri = np.fromfile(fh, np.int32, 3 * T)
            mto = np.fromfile(fh, np.int32, M + 1)
            mvo = np.fromfile(fh, np.int32, M + 1)
            mats = np.fromfile(fh, np.float32, 8 * M)
            lights = np.fromfile(fh, np.float32, 21 * L)
            lctor = np.fromfile(fh, np.float32, 11 * L)
            cam = np.fromfile(fh, np.float32, 12)
        return cls(pos, nrm, tri, mto, mvo, mats, lights, cam, w, h, lctor)

    def save(self, path):
        with open(path, "wb") as fh:
            fh.write(self.MAGIC)
            np.array([self.V, self.T, self.M, self.L, self.w, self.h], np.int32).tofile(fh)
            for a in (self.pos, self.nrm, self.tri, self.mesh_tri_off, self.mesh_vtx_off, self.mats, self.lights,
                      self.lights_ctor, self.cam):
                a.tofile(fh)

    @classmethod
    def build(cls, width, height, mesh_dir="../meshes", input_off=None, subdivisions=0, cache_dir=None):
        """The scene main() assembles (source/Main.cpp:165-208), built by the C++ host code
        (host/scene_host.cpp through lib/librt_host.so): Cornell box, 3 lights, the two meshes loaded
        from `mesh_dir`; `input_off` ("a.off" or "a.off,b.off,...") replaces cube_tri.off (and cube_tri2.off, then
        appends further meshes); `subdivisions` midpoint-subdivides the first; `cache_dir` enables the binary OFF cache."""
        lib = _host_lib()
        h = lib.rth_scene_create_cached(str(mesh_dir).encode(), (input_off or "").encode(), int(subdivisions),
                                        int(width), int(height), (cache_dir or "").encode())
        if not h:
            raise RuntimeError(lib.rth_last_error().decode(errors="replace"))
        try:
            cnt = np.zeros(4, np.int32)
            lib.rth_scene_counts(h, _capi.ptr(cnt))
            V, T, M, L = (int(v) for v in cnt)
            f, i = np.float32, np.int32
            pos, nrm, tri = np.zeros((V, 3), f), np.zeros((V, 3), f), np.zeros((T, 3), i)
            mto, mvo = np.zeros(M + 1, i), np.zeros(M + 1, i)
            mats, lights, cam = np.zeros((M, 8), f), np.zeros((L, 21), f), np.zeros(12, f)
            lib.rth_scene_get(h, *[_capi.ptr(a) for a in (pos, nrm, tri, mto, mvo, mats, lights, cam)])
        finally:
            lib.rth_scene_destroy(h)
        return cls(pos, nrm, tri, mto, mvo, mats, lights, cam, width, height)

    def with_size(self, width, height):
        """Same geometry, camera rebuilt for another aspect ratio (source/Main.cpp:169-170)."""
        cam = np.zeros(12, np.float32)
        _host_lib().rth_camera(int(width), int(height), _capi.ptr(cam))
        return Scene(self.pos, self.nrm, self.tri, self.mesh_tri_off, self.mesh_vtx_off, self.mats, self.lights, cam,
                     width, height, self.lights_ctor)

    def tri_mesh(self):
        out = np.zeros(self.T, np.int32)
        for m in range(self.M):
            out[self.mesh_tri_off[m]:self.mesh_tri_off[m + 1]] = m
        return out

    def build_bvh_host(self, pad_fraction=0.0):
        """The host BVH builder alone (no device needed): (nodes [n,16] float32, slot_triangle [T], depth).
        Split policy of BVH::from_triangles (source/BVH.h:100-161); layout in csrc/host_build.h."""
        lib = _capi.load()
        cs = self._as_c()
        cnt, depth = np.zeros(1, np.int32), np.zeros(1, np.int32)
        _capi.check(lib.rt_build_bvh_host(C.byref(cs), pad_fraction, None, 0, None, _capi.ptr(cnt), _capi.ptr(depth)))
        nodes, slots = np.zeros((int(cnt[0]), 16), np.float32), np.zeros(self.T, np.int32)
        _capi.check(lib.rt_build_bvh_host(C.byref(cs), pad_fraction, _capi.ptr(nodes), int(cnt[0]), _capi.ptr(slots),
                                          _capi.ptr(cnt), _capi.ptr(depth)))
        return nodes, slots, int(depth[0])

    def _as_c(self):
        s = _capi.rt_scene()
        s.num_vertices, s.num_triangles, s.num_meshes, s.num_lights = self.V, self.T, self.M, self.L
        s.positions, s.normals, s.triangles = _capi.ptr(self.pos), _capi.ptr(self.nrm), _capi.ptr(self.tri)
        s.mesh_first_triangle, s.mesh_first_vertex = _capi.ptr(self.mesh_tri_off), _capi.ptr(self.mesh_vtx_off)
        s.materials, s.lights = _capi.ptr(self.mats), _capi.ptr(self.lights)
        C.memmove(C.byref(s.camera), self.cam.ctypes.data, 48)
        return s


_hostlib = None


def _host_lib():
    """lib/librt_host.so: the C++ host code (OFF loader, scene assembly, camera) shared with bin/RayTracer."""
    global _hostlib
    if _hostlib is None:
        path = os.path.join(_capi.PKG_DIR, "lib", "librt_host.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run make -C ray-tracing-engine_b200")
        lib = C.CDLL(path)
        vp = C.c_void_p
        lib.rth_last_error.restype = C.c_char_p
        lib.rth_scene_create.restype = vp
        lib.rth_scene_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        lib.rth_scene_create_cached.restype = vp
        lib.rth_scene_create_cached.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        lib.rth_save_pcd.argtypes = [C.c_char_p, vp, C.c_int64]
        lib.rth_scene_destroy.argtypes = [vp]
        lib.rth_scene_counts.argtypes = [vp, vp]
        lib.rth_scene_get.argtypes = [vp] * 9
        lib.rth_load_off.restype = C.c_int
        lib.rth_load_off.argtypes = [C.c_char_p, C.c_int, vp, vp, vp, vp]
        lib.rth_camera.argtypes = [C.c_int, C.c_int, vp]
        lib.rth_background.argtypes = [C.c_int, C.c_int, vp]
        lib.rth_save_ppm.argtypes = [C.c_char_p, C.c_int, C.c_int, vp]
        lib.rth_save_ppm_binary.argtypes = [C.c_char_p, C.c_int, C.c_int, vp]
        _hostlib = lib
    return _hostlib


def load_off(path, subdivisions=0):
    """Mesh::loadOFF (source/Mesh.h:57-90) through the C++ host: (positions, normals, triangles)."""
    lib = _host_lib()
    cnt = np.zeros(2, np.int32)
    if lib.rth_load_off(str(path).encode(), int(subdivisions), _capi.ptr(cnt), None, None, None) != 0:
        raise RuntimeError(lib.rth_last_error().decode(errors="replace"))
    pos, nrm, tri = np.zeros((cnt[0], 3), np.float32), np.zeros((cnt[0], 3), np.float32), np.zeros((cnt[1], 3), np.int32)
    lib.rth_load_off(str(path).encode(), int(subdivisions), _capi.ptr(cnt), _capi.ptr(pos), _capi.ptr(nrm),
                     _capi.ptr(tri))
    return pos, nrm, tri


class Image:
    """source/Image.h: row-major RGB float pixels, y = 0 is the top row."""

    def __init__(self, width=64, height=64):
        self.width, self.height = int(width), int(height)
        self.pixels = np.zeros((self.height, self.width, 3), np.float32)

    def fillBackground(self):
        """Image::fillBackground (source/Image.cpp:12-21): vertical mix of two blues, binary32 per op."""
        f = np.float32
        c0, c1 = np.array([0.1, 0.2, 0.8], f), np.array([0.9, 0.9, 1.0], f)
        with np.errstate(all="ignore"):
            alpha = np.clip(np.arange(self.height, dtype=f) / f(self.height - 1), f(0), f(1)).astype(f)
        rows = c0[None, :] * (f(1.0) - alpha)[:, None] + c1[None, :] * alpha[:, None]
        self.pixels[:] = rows[:, None, :]
        return self

    def to8(self):
        """Image::savePPM's quantisation (source/Image.cpp:31-38): unsigned(255.f * v), truncating."""
        return (np.float32(255.0) * self.pixels).astype(np.uint32).astype(np.uint8)

    def savePPM(self, filename):
        """ASCII P3 exactly as source/Image.cpp:23-43 writes it (one line of values, trailing space)."""
        v = (np.float32(255.0) * self.pixels).astype(np.uint32).reshape(-1)
        try:
            fh = open(filename, "w")
        except OSError:
            raise SystemExit(f"Cannot open file {filename}")
        with fh:
            fh.write(f"P3\n{self.width} {self.height}\n255\n")
            fh.write(" ".join(map(str, v.tolist())) + " \n")


def _params(width, height, num_rays, mode, num_photons=0, k=5, seed=1, shard_rank=0, shard_count=1, shard_tile=16,
            sample_first=0, sample_count=0, samples_per_batch=0, bvh_pad=0.0, flags=0):
    p = rt_params()
    p.width, p.height, p.num_rays = int(width), int(height), int(num_rays)
    p.mode = int(mode) if int(mode) in (0, 1) else 0  # source/CommandLine.h:84-87
    p.num_photons, p.k, p.seed = int(num_photons), int(k), int(seed)
    p.shard_rank, p.shard_count, p.shard_tile = int(shard_rank), int(shard_count), int(shard_tile)
    p.sample_first, p.sample_count = int(sample_first), int(sample_count)
    p.samples_per_batch, p.bvh_pad, p.flags = int(samples_per_batch), float(bvh_pad), int(flags)
    return p


class Renderer:
    """source/Renderer.h:19-36 -- Renderer(scene, numRays, mode, rayTracer[, numPhotons, k]).

    `rayTracer` is accepted for signature parity and ignored (RayTracer is stateless,
    source/RayTracer.h:17-21).  Additive keyword arguments select the image size (defaults to the
    scene's camera aspect), the random seed, the GPU and the shard of the pixel grid.
    """

    def __init__(self, scene: Scene, numRays: int, mode: int, rayTracer=None, numPhotons: int = 0, k: int = 5, *,
                 width=None, height=None, seed=1, device=0, shard_rank=0, shard_count=1, shard_tile=16,
                 sample_first=0, sample_count=0, samples_per_batch=0, bvh_pad=0.0, flags=0):
        self.lib = _capi.load()
        self.scene = scene
        self._kw = dict(width=width or scene.w, height=height or scene.h, num_rays=numRays, mode=mode,
                        num_photons=numPhotons, k=k, seed=seed, shard_rank=shard_rank, shard_count=shard_count,
                        shard_tile=shard_tile, sample_first=sample_first, sample_count=sample_count,
                        samples_per_batch=samples_per_batch, bvh_pad=bvh_pad, flags=flags)
        self.params = _params(**self._kw)
        self._ctx = C.c_void_p()
        cs = scene._as_c()
        _capi.check(self.lib.rt_create(C.byref(cs), C.byref(self.params), int(device), C.byref(self._ctx)))

    # ---- lifetime --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.rt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, **kw):
        """Change width/height/num_rays/mode/num_photons/k/seed/shard_*/flags of the live context."""
        self._kw.update(kw)
        self.params = _params(**self._kw)
        _capi.check(self.lib.rt_set_params(self._ctx, C.byref(self.params)))
        return self

    width = property(lambda s: s.params.width)
    height = property(lambda s: s.params.height)

    # ---- the path: Renderer::render --------------------------------------------------------------
    def render(self, image: Image, every: int = 0, on_update=None) -> Image:
        """source/Renderer.cpp:203-272: image holds the background on entry, the render on exit.

        on_update(samples_done, num_rays, pixels[H,W,3]) is the reference's per-pass `update.ppm`
        (Renderer.cpp:262-269), called every `every` sample passes and after the last one."""
        if image.width != self.width or image.height != self.height:
            raise ValueError("image size differs from the renderer's")
        buf = np.ascontiguousarray(image.pixels, np.float32)
        if on_update is not None and every > 0:
            H, W = self.height, self.width

            def trampoline(_user, done, total, ptr):
                on_update(int(done), int(total), np.ctypeslib.as_array(ptr, shape=(H, W, 3)).copy())

            cb = _capi.PROGRESS_FN(trampoline)
            _capi.check(self.lib.rt_render_progressive(self._ctx, _capi.ptr(buf), int(every), cb, None))
        else:
            _capi.check(self.lib.rt_render(self._ctx, _capi.ptr(buf)))
        image.pixels = buf
        return image

    def render_accumulate(self):
        """updateImage / counter of source/Renderer.cpp:254-258 (zero outside this shard)."""
        s = np.zeros((self.height, self.width, 3), np.float32)
        c = np.zeros((self.height, self.width), np.int32)
        _capi.check(self.lib.rt_render_accumulate(self._ctx, _capi.ptr(s), _capi.ptr(c)))
        return s, c

    def render_accumulate_device(self, sum_rgb_ptr: int, counter_ptr: int):
        """Same, into device buffers (e.g. torch tensors' data_ptr()) on this context's GPU."""
        _capi.check(self.lib.rt_render_accumulate_device(self._ctx, C.c_void_p(sum_rgb_ptr), C.c_void_p(counter_ptr)))

    def render_accumulate_packed_device(self, sum_rgbn_ptr: int):
        """The sums and the counter as one [H,W,4] float32 device buffer (a single reduce across GPUs)."""
        _capi.check(self.lib.rt_render_accumulate_packed_device(self._ctx, C.c_void_p(sum_rgbn_ptr)))

    def composite_packed_device(self, num_rays, sum_rgbn_ptr: int, background, out=None):
        """Renderer.cpp:262-265 on the packed DEVICE frame (after the multi-GPU reduce); returns [H,W,3].
        out: an optional caller-owned float32 [H,W,3] host frame (e.g. pinned memory) that receives the result."""
        if out is None:
            out = np.ascontiguousarray(background, np.float32).copy()
        else:
            np.copyto(out, background)
        _capi.check(self.lib.rt_composite_packed_device(self._ctx, int(num_rays), C.c_void_p(sum_rgbn_ptr),
                                                        _capi.ptr(out)))
        return out

    @staticmethod
    def composite(num_rays, sum_rgb, counter, background):
        """source/Renderer.cpp:262-265."""
        h, w = counter.shape
        out = np.ascontiguousarray(background, np.float32).copy()
        _capi.check(_capi.load().rt_composite(w, h, int(num_rays), _capi.ptr(_capi.f32(sum_rgb)),
                                              _capi.ptr(_capi.i32(counter)), _capi.ptr(out)))
        return out

    def composite_device(self, num_rays, sum_rgb_ptr: int, counter_ptr: int, background):
        """Renderer.cpp:262-265 on full-frame DEVICE sums/counters (after the multi-GPU reduce); returns [H,W,3]."""
        out = np.ascontiguousarray(background, np.float32).copy()
        _capi.check(self.lib.rt_composite_device(self._ctx, int(num_rays), C.c_void_p(sum_rgb_ptr),
                                                 C.c_void_p(counter_ptr), _capi.ptr(out)))
        return out

    def render_samples(self, window=None, samples=None):
        """Clamped per-sample colours [ns,h,w,3] and posIntersectionFound [ns,h,w] over a window."""
        x0, y0, x1, y1 = window if window else (0, 0, self.width, self.height)
        s0, s1 = samples if samples else (0, self.params.num_rays)
        rgb = np.zeros((s1 - s0, y1 - y0, x1 - x0, 3), np.float32)
        found = np.zeros((s1 - s0, y1 - y0, x1 - x0), np.uint8)
        _capi.check(self.lib.rt_render_samples(self._ctx, x0, y0, x1, y1, s0, s1, _capi.ptr(rgb), _capi.ptr(found)))
        return rgb, found

    # ---- parity hooks ----------------------------------------------------------------------------
    def rayTrace(self, rays, brute_force=False):
        """RayTracer::rayTrace (source/RayTracer.h:27-53) on [n,6] rays.

        Returns dict(hit, mesh, tri3 (mesh-local vertex triple, like the reference's Vec3i), uvd,
        tri_index (global triangle index))."""
        r = _capi.f32(rays).reshape(-1, 6)
        n = len(r)
        hits = np.zeros(n, np.dtype([("tri", np.int32), ("u", np.float32), ("v", np.float32), ("t", np.float32)]))
        _capi.check(self.lib.rt_trace_rays(self._ctx, _capi.ptr(r), n, _capi.ptr(hits),
                                           RT_FLAG_BRUTE_FORCE if brute_force else 0))
        tri = hits["tri"].astype(np.int32)
        hit = (tri >= 0).astype(np.int32)
        safe = np.where(tri >= 0, tri, 0)
        if self.scene.T:
            mesh = np.where(tri >= 0, self.scene.tri_mesh()[safe], 0).astype(np.int32)
            tri3 = np.where((tri >= 0)[:, None], self.scene.tri[safe] - self.scene.mesh_vtx_off[mesh][:, None], 0)
        else:  # an empty scene: every ray misses
            mesh, tri3 = np.zeros(n, np.int32), np.zeros((n, 3), np.int32)
        uvd = np.stack([hits["u"], hits["v"], hits["t"]], 1).astype(np.float32)
        return dict(hit=hit, mesh=mesh, tri3=tri3.astype(np.int32), uvd=uvd, tri_index=tri)

    def occluded(self, rays, brute_force=False):
        """The boolean use of rayTrace for shadow rays (source/Renderer.cpp:52-55)."""
        r = _capi.f32(rays).reshape(-1, 6)
        out = np.zeros(len(r), np.uint8)
        _capi.check(self.lib.rt_occluded(self._ctx, _capi.ptr(r), len(r), _capi.ptr(out),
                                         RT_FLAG_BRUTE_FORCE if brute_force else 0))
        return out

    def evaluateColorResponse(self, mat8, n_wi_wo):
        """Material::evaluateColorResponse (source/Material.h:25-36) on [n,9] (normal, wi, wo)."""
        m = _capi.rt_material()
        C.memmove(C.byref(m), _capi.f32(mat8).ctypes.data, 32)
        a = _capi.f32(n_wi_wo).reshape(-1, 9)
        out = np.zeros((len(a), 3), np.float32)
        _capi.check(self.lib.rt_eval_bsdf(self._ctx, C.byref(m), _capi.ptr(a), len(a), _capi.ptr(out)))
        return out

    def hsphereUniformSample(self, normals, seed, domain, index0):
        """RayTracer::hsphereUniformSample (source/RayTracer.h:95-107) around [n,3] normals; item i draws from
        stream (seed, domain, index0 + i)."""
        a = _capi.f32(normals).reshape(-1, 3)
        out = np.zeros_like(a)
        _capi.check(self.lib.rt_eval_hsphere(self._ctx, int(seed), int(domain), int(index0), _capi.ptr(a), len(a),
                                             _capi.ptr(out)))
        return out

    # ---- photon map ------------------------------------------------------------------------------
    def photons_per_light(self):
        out = C.c_int32()
        _capi.check(self.lib.rt_photons_per_light(self._ctx, C.byref(out)))
        return out.value

    def emit_photons(self, first_path=0, num_paths=-1):
        """PhotonMap::PhotonMap (source/PhotonMap.h:14-50) for a range of paths of every light.

        Returns (photons[n,7] in (light, path) order, per_light_counts[L], depth_hist[20])."""
        per = self.photons_per_light()
        cap = max(1, per * self.scene.L)
        out = np.zeros((cap, 7), np.float32)
        counts = np.zeros(max(self.scene.L, 1), np.int64)
        hist = np.zeros(20, np.int32)
        _capi.check(self.lib.rt_emit_photons(self._ctx, first_path, num_paths, _capi.ptr(out), cap, _capi.ptr(counts),
                                             _capi.ptr(hist)))
        return out[:int(counts.sum())].copy(), counts[:self.scene.L], hist

    def emit_photons_device(self, first_path, num_paths, out7_ptr: int, capacity: int):
        """rt_emit_photons_device: the stored particles stay on the GPU, compacted in (light, path) order at
        `out7_ptr` ([capacity,7] float32 device memory).  Returns (per_light_counts[L], depth_hist[20])."""
        counts = np.zeros(max(self.scene.L, 1), np.int64)
        hist = np.zeros(20, np.int32)
        _capi.check(self.lib.rt_emit_photons_device(self._ctx, int(first_path), int(num_paths), C.c_void_p(out7_ptr),
                                                    int(capacity), _capi.ptr(counts), _capi.ptr(hist)))
        return counts[:self.scene.L], hist

    def splice_photons_device(self, gathered_ptr: int, world: int, stride: int, counts, out7_ptr: int, capacity: int):
        """rt_splice_photons_device: all-gathered shards -> the single-process (light, path) order, on the device.
        counts: [world, L] particles of light l in rank r's shard.  Returns the total."""
        cnt = np.ascontiguousarray(counts, np.int64).reshape(world, -1)
        total = C.c_int64()
        _capi.check(self.lib.rt_splice_photons_device(self._ctx, C.c_void_p(gathered_ptr), int(world), int(stride),
                                                      _capi.ptr(cnt), C.c_void_p(out7_ptr), int(capacity),
                                                      C.byref(total)))
        return total.value

    def set_photons_device(self, photons7_ptr: int, n: int):
        _capi.check(self.lib.rt_set_photons_device(self._ctx, C.c_void_p(photons7_ptr), int(n)))

    def set_photons(self, photons7):
        a = _capi.f32(photons7).reshape(-1, 7)
        _capi.check(self.lib.rt_set_photons(self._ctx, _capi.ptr(a), len(a)))

    def build_photon_map(self):
        _capi.check(self.lib.rt_build_photon_map(self._ctx))

    def kdtree(self):
        n = C.c_int64()
        _capi.check(self.lib.rt_get_photons(self._ctx, None, 0, C.byref(n)))
        n = n.value
        nodes = np.zeros((n, 7), np.float32)
        left, right = np.zeros(n, np.int32), np.zeros(n, np.int32)
        root = C.c_int32()
        _capi.check(self.lib.rt_get_kdtree(self._ctx, _capi.ptr(nodes), _capi.ptr(left), _capi.ptr(right),
                                           C.byref(root), n))
        return nodes, left, right, root.value

    def knearest(self, points, k):
        """kdtree::knearest (source/kdtree.h:180-195): [n,k] indices into kdtree()[0], reference order."""
        q = _capi.f32(points).reshape(-1, 3)
        out = np.zeros((len(q), k), np.int32)
        _capi.check(self.lib.rt_knn(self._ctx, _capi.ptr(q), len(q), int(k), _capi.ptr(out)))
        return out

    def savePhotonMap(self, filename="pointcloud.pcd"):
        """PhotonMap::saveToPCD (source/PhotonMap.h:59-84).  The reference's own call writes a
        header-only file (it runs before render and on a shadowed member); we write the real map."""
        nodes = self.kdtree()[0] if self.params.num_photons > 0 else np.zeros((0, 7), np.float32)
        save_pcd(filename, nodes)

    # ---- introspection ---------------------------------------------------------------------------
    def stats(self):
        s = rt_stats()
        _capi.check(self.lib.rt_get_stats(self._ctx, C.byref(s)))
        out = {name: getattr(s, name) for name, _ in rt_stats._fields_ if name not in ("kernel_ms", "kernel_count")}
        out["kernel_ms"] = {n: float(s.kernel_ms[i]) for i, n in enumerate(_capi.KERNEL_CLASSES)}
        out["kernel_count"] = {n: int(s.kernel_count[i]) for i, n in enumerate(_capi.KERNEL_CLASSES)}
        return out

    def reset_stats(self):
        _capi.check(self.lib.rt_reset_stats(self._ctx))

    def bvh_slots(self):
        """leaf slot -> global triangle index of this context's BVH"""
        out = np.zeros(self.scene.T, np.int32)
        _capi.check(self.lib.rt_get_bvh_slots(self._ctx, _capi.ptr(out), len(out)))
        return out

    def bvh(self):
        n, d = C.c_int32(), C.c_int32()
        _capi.check(self.lib.rt_get_bvh(self._ctx, None, 0, C.byref(n), C.byref(d)))
        nodes = np.zeros((n.value, 16), np.float32)
        _capi.check(self.lib.rt_get_bvh(self._ctx, _capi.ptr(nodes), n.value, C.byref(n), C.byref(d)))
        return nodes, d.value


def save_pcd(filename, photons7):
    """PhotonMap::saveToPCD (source/PhotonMap.h:59-84) through the C++ host writer (the CLI's own)."""
    a = np.ascontiguousarray(photons7, np.float32).reshape(-1, 7)
    _host_lib().rth_save_pcd(str(filename).encode(), _capi.ptr(a), len(a))


def build_kdtree_host(photons7, canonical=False):
    """The host kd-tree builder alone (no device needed): (nodes [n,7] in kdtree::make_tree's array order, list index
    of every node or None, height).  canonical=True: the tree of the exact k-NN mode (see rt_build_kdtree_host)."""
    a = np.ascontiguousarray(photons7, np.float32).reshape(-1, 7).copy()
    orig = np.zeros(len(a), np.int32)
    h = C.c_int32()
    _capi.check(_capi.load().rt_build_kdtree_host(_capi.ptr(a), len(a), 1 if canonical else 0, _capi.ptr(orig), C.byref(h)))
    return a, (orig if canonical else None), h.value


def device_count():
    return _capi.load().rt_device_count()


def shard_pixels(width, height, shard_rank, shard_count, shard_tile=16):
    """Pixel indices (y*W + x) owned by a shard of the interleaved-tile partition (host only, no GPU)."""
    p = _params(width, height, 1, 0, shard_rank=shard_rank, shard_count=shard_count, shard_tile=shard_tile)
    n = C.c_int64()
    _capi.check(_capi.load().rt_shard_pixels(C.byref(p), None, 0, C.byref(n)))
    out = np.zeros(n.value, np.int32)
    _capi.check(_capi.load().rt_shard_pixels(C.byref(p), _capi.ptr(out), n.value, C.byref(n)))
    return out
