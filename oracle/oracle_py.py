"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  It consumes the same rptb_scene_desc the CUDA library
consumes (built by rpt_b200.api.FlatScene), so both sides see identical inputs.
PARITY UNPINNED by the reference's own tests (see oracle.cpp header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rpt_b200 import _capi as capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

dp = capi.c_double_p


def build() -> None:
    """Compile the C++ restatement (gcc only; see the Makefile for the flags)."""
    subprocess.check_call(["make", "-s", "oracle"], cwd=os.path.dirname(_HERE))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    L.oracle_scene_create.restype = C.c_void_p
    L.oracle_scene_create.argtypes = [C.POINTER(capi.SceneDesc), C.c_int]
    L.oracle_scene_destroy.argtypes = [C.c_void_p]
    L.oracle_hardware_threads.restype = C.c_int
    L.oracle_render.restype = C.c_int
    L.oracle_render.argtypes = [C.c_void_p, C.POINTER(capi.Camera), C.POINTER(capi.RenderParams), dp,
                                C.POINTER(capi.Stats), C.c_int]
    L.oracle_closest_hit.restype = C.c_int
    L.oracle_closest_hit.argtypes = [C.c_void_p, dp, C.c_uint64, C.c_double, dp, capi.c_i32_p, dp,
                                     C.POINTER(capi.Stats)]
    L.oracle_bsdf.argtypes = [C.POINTER(capi.Material), dp, C.c_uint64, dp]
    L.oracle_sample_f.argtypes = [C.POINTER(capi.Material), dp, C.c_uint64, C.c_uint64, dp, dp]
    L.oracle_illuminate.argtypes = [C.c_void_p, C.c_uint32, dp, C.c_uint64, C.c_uint64, dp, dp, dp]
    L.oracle_build_kdtree.restype = C.c_int
    L.oracle_build_kdtree.argtypes = [dp, C.c_uint64, C.POINTER(capi.KdTreeOut)]
    L.oracle_build_kdtree_boxes.restype = C.c_int
    L.oracle_build_kdtree_boxes.argtypes = [dp, C.c_uint64, C.POINTER(capi.KdTreeOut)]
    L.oracle_shape_bounds.restype = C.c_int
    L.oracle_shape_bounds.argtypes = [C.POINTER(capi.SceneDesc), C.POINTER(capi.Object), dp]
    L.oracle_free_kdtree.argtypes = [C.POINTER(capi.KdTreeOut)]
    L.oracle_hex_color.argtypes = [C.c_uint32, dp]
    L.oracle_color_bytes.argtypes = [dp, capi.c_u8_p]
    L.oracle_film_resolve.argtypes = [dp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, capi.c_u8_p]
    L.oracle_variance.restype = C.c_double
    L.oracle_variance.argtypes = [dp, C.c_uint32, C.c_uint64]
    L.oracle_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, capi.c_u32_p]
    L.oracle_draws.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, dp]
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(dp)


class OracleScene:
    """The oracle's own scene object, built from the same flattened description."""

    def __init__(self, flat, brute_force: bool = False):
        self.flat = flat
        self.handle = C.c_void_p(lib().oracle_scene_create(C.byref(flat.desc), 1 if brute_force else 0))

    def close(self):
        if self.handle:
            lib().oracle_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, camera, params, nthreads: int = 0):
        """Renderer::sample -> (width*height, 3) float64 + stats dict."""
        out = np.empty((params.width * params.height, 3), np.float64)
        stats = capi.Stats()
        cam = camera.to_c() if hasattr(camera, "to_c") else camera
        lib().oracle_render(self.handle, C.byref(cam), C.byref(params), _p(out), C.byref(stats), nthreads)
        return out, stats.as_dict()

    def closest_hit(self, rays: np.ndarray, t_min: float = 1e-12):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        n = rays.shape[0]
        t = np.empty(n, np.float64)
        obj = np.empty(n, np.int32)
        nrm = np.empty((n, 3), np.float64)
        stats = capi.Stats()
        lib().oracle_closest_hit(self.handle, _p(rays), n, t_min, _p(t), obj.ctypes.data_as(capi.c_i32_p), _p(nrm),
                                 C.byref(stats))
        return t, obj, nrm, stats.as_dict()

    def illuminate(self, light_index: int, pos: np.ndarray, seed: int = 0):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        n = pos.shape[0]
        inten = np.empty((n, 3))
        wi = np.empty((n, 3))
        dist = np.empty(n)
        lib().oracle_illuminate(self.handle, light_index, _p(pos), n, seed, _p(inten), _p(wi), _p(dist))
        return inten, wi, dist


def bsdf(material, dirs: np.ndarray) -> np.ndarray:
    dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 9)
    out = np.empty((dirs.shape[0], 3))
    m = material.to_c()
    lib().oracle_bsdf(C.byref(m), _p(dirs), dirs.shape[0], _p(out))
    return out


def sample_f(material, dirs: np.ndarray, seed: int = 0):
    dirs = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 6)
    n = dirs.shape[0]
    wi = np.empty((n, 3))
    pdf = np.empty(n)
    m = material.to_c()
    lib().oracle_sample_f(C.byref(m), _p(dirs), n, seed, _p(wi), _p(pdf))
    return wi, pdf


def shape_bounds(flat, obj_index: int):
    """Bounded::bounding_box of scene object `obj_index` -> (p_min, p_max), or None (Plane)."""
    out = np.empty(6)
    if not lib().oracle_shape_bounds(C.byref(flat.desc), C.byref(flat.objects[obj_index]), _p(out)):
        return None
    return out[:3].copy(), out[3:].copy()


def build_kdtree(tris: np.ndarray, boxes: bool = False):
    """KdTree::new restated literally -> (nodes structured array view, refs, depth, max_leaf).
    boxes=True: `tris` is an (n, 6) array of bounding boxes (p_min, p_max) instead of triangles."""
    tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 6 if boxes else 18)
    out = capi.KdTreeOut()
    if boxes:
        lib().oracle_build_kdtree_boxes(_p(tris), tris.shape[0], C.byref(out))
    else:
        lib().oracle_build_kdtree(_p(tris), tris.shape[0], C.byref(out))
    try:
        n = int(out.nnodes)
        nodes = np.frombuffer(C.string_at(out.nodes, C.sizeof(capi.KdNode) * n), dtype=KDNODE_DTYPE).copy()
        refs = np.ctypeslib.as_array(out.refs, shape=(int(out.nrefs),)).copy() if out.nrefs else np.zeros(0, np.uint32)
        return nodes, refs, int(out.depth), int(out.max_leaf)
    finally:
        lib().oracle_free_kdtree(C.byref(out))


KDNODE_DTYPE = np.dtype([("split", "<f8"), ("kind", "<u4"), ("left", "<u4"), ("right", "<u4"),
                         ("first_ref", "<u4"), ("num_refs", "<u4"), ("_pad", "<u4")])


def hex_color(x: int) -> np.ndarray:
    out = np.empty(3)
    lib().oracle_hex_color(x, _p(out))
    return out


def color_bytes(c) -> list:
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.empty(3, np.uint8)
    lib().oracle_color_bytes(_p(c), out.ctypes.data_as(capi.c_u8_p))
    return [int(v) for v in out]


def film_resolve(sums: np.ndarray, nbatches: int, width: int, height: int, radius: int) -> np.ndarray:
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    out = np.empty((height, width, 3), np.uint8)
    lib().oracle_film_resolve(_p(sums), nbatches, width, height, radius, out.ctypes.data_as(capi.c_u8_p))
    return out


def variance(batches: np.ndarray) -> float:
    batches = np.ascontiguousarray(batches, dtype=np.float64)
    return float(lib().oracle_variance(_p(batches), batches.shape[0], batches.shape[1]))


def philox(seed: int, pixel: int, sample: int, nblocks: int, first_block: int = 0) -> np.ndarray:
    out = np.empty(4 * nblocks, np.uint32)
    lib().oracle_philox(seed, pixel, sample, first_block, nblocks, out.ctypes.data_as(capi.c_u32_p))
    return out


def draws(seed: int, pixel: int, kind: int, count: int, param: int = 0) -> np.ndarray:
    out = np.empty(count * (2 if kind in (2, 3) else 1))
    lib().oracle_draws(seed, pixel, kind, param, count, _p(out))
    return out


def hardware_threads() -> int:
    return int(lib().oracle_hardware_threads())
