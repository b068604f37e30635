"""ctypes binding of lib/librt_b200.so (the C ABI in include/rt_b200.h).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute entry point returns RT_ERR_NO_DEVICE unless a CUDA device is present, and a missing
shared library raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB selects another build of the same library (A/B runs of compile-time variants); default: the product build
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(PKG_DIR, "lib", "librt_b200.so")

RT_OK = 0
RT_ERR_INVALID, RT_ERR_NO_DEVICE, RT_ERR_CUDA, RT_ERR_EMPTY_TREE, RT_ERR_K_TOO_LARGE, RT_ERR_OOM = -1, -2, -3, -4, -5, -6
RT_FLAG_BRUTE_FORCE = 1
RT_FLAG_KNN_EXACT = 2
RT_MAX_K = 4096
RT_MAX_LIGHTS = 4096
KERNEL_CLASSES = ("raygen", "trace_nearest", "sort", "shade", "trace_any", "combine", "resolve", "emit", "other", "gather")


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rt_b200 error {code}: {msg}")
        self.code = code


class rt_material(C.Structure):
    _fields_ = [("kd", C.c_float), ("alpha", C.c_float), ("albedo", C.c_float * 3), ("f0", C.c_float * 3)]


class rt_light(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("normal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("horizontal", C.c_float * 3), ("intensity", C.c_float),
                ("side", C.c_float), ("ac", C.c_float), ("al", C.c_float), ("aq", C.c_float), ("factor", C.c_float)]


class rt_camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("lower_left", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3)]


class rt_scene(C.Structure):
    _fields_ = [("num_vertices", C.c_int32), ("num_triangles", C.c_int32), ("num_meshes", C.c_int32),
                ("num_lights", C.c_int32), ("positions", C.c_void_p), ("normals", C.c_void_p),
                ("triangles", C.c_void_p), ("mesh_first_triangle", C.c_void_p), ("mesh_first_vertex", C.c_void_p),
                ("materials", C.c_void_p), ("lights", C.c_void_p), ("camera", rt_camera)]


class rt_params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("num_rays", C.c_int32), ("mode", C.c_int32),
                ("num_photons", C.c_int32), ("k", C.c_int32), ("seed", C.c_uint64), ("shard_rank", C.c_int32),
                ("shard_count", C.c_int32), ("shard_tile", C.c_int32), ("sample_first", C.c_int32),
                ("sample_count", C.c_int32), ("samples_per_batch", C.c_int32),
                ("bvh_pad", C.c_float), ("flags", C.c_int32)]


class rt_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("primary_rays", C.c_uint64), ("bounce_rays", C.c_uint64),
                ("shadow_rays", C.c_uint64), ("photon_rays", C.c_uint64), ("knn_queries", C.c_uint64),
                ("samples", C.c_uint64), ("kernel_launches", C.c_uint64), ("device_ms", C.c_double),
                ("trace_ms", C.c_double), ("photon_ms", C.c_double), ("bvh_nodes", C.c_int32),
                ("bvh_depth", C.c_int32), ("photons_stored", C.c_int64), ("create_ms", C.c_double),
                ("bvh_build_ms", C.c_double), ("kd_build_ms", C.c_double), ("kd_visits", C.c_uint64),
                ("kernel_ms", C.c_double * len(KERNEL_CLASSES)), ("kernel_count", C.c_uint64 * len(KERNEL_CLASSES)),
                ("device_ms_total", C.c_double)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float))

# every symbol include/rt_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
SYMBOLS = {
    "rt_last_error": (C.c_char_p, []),
    "rt_device_count": (C.c_int, []),
    "rt_abi_sizes": (C.c_int, [_vp, _i32]),
    "rt_create": (C.c_int, [C.POINTER(rt_scene), C.POINTER(rt_params), C.c_int, C.POINTER(_vp)]),
    "rt_destroy": (C.c_int, [_vp]),
    "rt_set_params": (C.c_int, [_vp, C.POINTER(rt_params)]),
    "rt_render": (C.c_int, [_vp, _vp]),
    "rt_render_progressive": (C.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "rt_render_accumulate": (C.c_int, [_vp, _vp, _vp]),
    "rt_render_accumulate_device": (C.c_int, [_vp, _vp, _vp]),
    "rt_render_accumulate_packed_device": (C.c_int, [_vp, _vp]),
    "rt_composite_packed_device": (C.c_int, [_vp, _i32, _vp, _vp]),
    "rt_composite": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp]),
    "rt_composite_device": (C.c_int, [_vp, _i32, _vp, _vp, _vp]),
    "rt_render_samples": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "rt_trace_rays": (C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    "rt_occluded": (C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    "rt_eval_bsdf": (C.c_int, [_vp, C.POINTER(rt_material), _vp, _i64, _vp]),
    "rt_eval_hsphere": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, _i64, _vp]),
    "rt_photons_per_light": (C.c_int, [_vp, _vp]),
    "rt_emit_photons": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "rt_set_photons": (C.c_int, [_vp, _vp, _i64]),
    "rt_emit_photons_device": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "rt_splice_photons_device": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _i64, _vp]),
    "rt_set_photons_device": (C.c_int, [_vp, _vp, _i64]),
    "rt_build_photon_map": (C.c_int, [_vp]),
    "rt_get_photons": (C.c_int, [_vp, _vp, _i64, _vp]),
    "rt_knn": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "rt_get_kdtree": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64]),
    "rt_shard_pixels": (C.c_int, [C.POINTER(rt_params), _vp, _i64, _vp]),
    "rt_profiler_range": (C.c_int, [C.c_int]),
    "rt_get_stats": (C.c_int, [_vp, C.POINTER(rt_stats)]),
    "rt_reset_stats": (C.c_int, [_vp]),
    "rt_get_bvh": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rt_get_bvh_slots": (C.c_int, [_vp, _vp, _i64]),
    "rt_build_kdtree_host": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "rt_build_bvh_host": (C.c_int, [C.POINTER(rt_scene), C.c_float, _vp, _i64, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the CUDA library; fail loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(or make -C ray-tracing-engine_b200). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc):
    if rc != RT_OK:
        raise RtError(rc, load().rt_last_error().decode(errors="replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
