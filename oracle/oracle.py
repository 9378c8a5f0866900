"""ctypes front-end of the CPU oracle (oracle/ssrlcv_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import this module.  The product package (ssrlcv_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ssrlcv_b200 import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssrlcv_oracle.so")
_lib = None


class Camera(C.Structure):
    _fields_ = [
        ("cam_pos", C.c_float * 3), ("cam_rot", C.c_float * 3), ("fov", C.c_float * 2), ("foc", C.c_float),
        ("_pad0", C.c_float), ("dpix", C.c_float * 2), ("timeStamp", C.c_int64), ("ecef_offset", C.c_float * 3),
        ("no_rot", C.c_uint8), ("_pad1", C.c_uint8 * 3), ("size", C.c_uint32 * 2),
    ]


class Params(C.Structure):
    _fields_ = [
        ("mask_kind", C.c_int32), ("out_kind", C.c_int32), ("has_seed", C.c_int32), ("trig_mode", C.c_int32),
        ("abs_threshold", C.c_float), ("rel_threshold", C.c_float), ("query_id", C.c_uint32), ("target_id", C.c_uint32),
        ("epsilon", C.c_float), ("delta", C.c_float), ("fundamental", C.c_float * 9),
        ("query_camera", Camera), ("target_projection", C.c_float * 12),
    ]


assert C.sizeof(Camera) == 80


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ssrlcv_oracle.c")
    hdr = os.path.join(_HERE, "ssrlcv_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libssrlcv_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        vp, u64, f32p = C.c_void_p, C.c_uint64, C.POINTER(C.c_float)
        _lib.orc_dist2.restype = C.c_int32
        _lib.orc_dist2.argtypes = [vp, vp]
        _lib.orc_cuda_sinf.restype = C.c_float
        _lib.orc_cuda_sinf.argtypes = [C.c_float]
        _lib.orc_cuda_cosf.restype = C.c_float
        _lib.orc_cuda_cosf.argtypes = [C.c_float]
        _lib.orc_projection_matrix.argtypes = [C.POINTER(Camera), f32p]
        _lib.orc_epipolar_endpoints.argtypes = [C.POINTER(Camera), f32p, f32p, C.c_float, C.c_int, f32p, f32p]
        _lib.orc_seed_distances.argtypes = [vp, u64, vp, u64, vp]
        _lib.orc_ratio_threshold.restype = C.c_float
        _lib.orc_ratio_threshold.argtypes = [C.c_int, C.c_int, C.c_float]
        _lib.orc_match_winners.argtypes = [C.POINTER(Params), vp, u64, vp, u64, vp, vp, vp, vp]
        _lib.orc_generate_matches.restype = u64
        _lib.orc_generate_matches.argtypes = [C.POINTER(Params), vp, u64, vp, u64, vp, vp]
        _lib.orc_refine.restype = u64
        _lib.orc_refine.argtypes = [vp, u64, C.c_float]
        _lib.orc_sort.argtypes = [vp, u64]
        _lib.orc_raw_matches.argtypes = [vp, u64, vp]
        _lib.orc_fuse_nview.restype = u64
        _lib.orc_fuse_nview.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_uint32, vp, vp, vp]
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def camera_struct(cam) -> Camera:
    """cam: numpy scalar/0-d array of dtype types.CAMERA."""
    c = Camera()
    C.memmove(C.byref(c), np.asarray(cam, dtype=T.CAMERA).tobytes(), 80)
    return c


def projection_matrix(cam) -> np.ndarray:
    c = camera_struct(cam)
    P = np.zeros(12, np.float32)
    lib().orc_projection_matrix(C.byref(c), P.ctypes.data_as(C.POINTER(C.c_float)))
    return P


def epipolar_endpoints(qcam, P, loc, delta, trig_mode=0):
    c = camera_struct(qcam)
    P = np.ascontiguousarray(P, np.float32)
    loc = np.ascontiguousarray(loc, np.float32)
    p1 = np.zeros(2, np.float32)
    p2 = np.zeros(2, np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_epipolar_endpoints(C.byref(c), P.ctypes.data_as(fp), loc.ctypes.data_as(fp), float(delta), trig_mode,
                                 p1.ctypes.data_as(fp), p2.ctypes.data_as(fp))
    return p1, p2


def make_params(out_kind=T.OUT_MATCH, mask_kind=T.MASK_NONE, has_seed=False, abs_threshold=np.finfo(np.float32).max,
                rel_threshold=1.0, query_id=0, target_id=1, epsilon=0.0, delta=0.0, fundamental=None,
                query_camera=None, target_camera=None, target_projection=None, trig_mode=0) -> Params:
    p = Params()
    p.mask_kind, p.out_kind, p.has_seed, p.trig_mode = mask_kind, out_kind, int(bool(has_seed)), trig_mode
    p.abs_threshold, p.rel_threshold = float(np.float32(abs_threshold)), float(np.float32(rel_threshold))
    p.query_id, p.target_id = query_id, target_id
    p.epsilon, p.delta = float(np.float32(epsilon)), float(np.float32(delta))
    if fundamental is not None:
        p.fundamental = (C.c_float * 9)(*np.asarray(fundamental, np.float32).reshape(9))
    if query_camera is not None:
        p.query_camera = camera_struct(query_camera)
    if target_projection is None and target_camera is not None:
        target_projection = projection_matrix(target_camera)
    if target_projection is not None:
        p.target_projection = (C.c_float * 12)(*np.asarray(target_projection, np.float32).reshape(12))
    return p


def seed_distances(q: np.ndarray, seed: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q, T.FEATURE)
    seed = np.ascontiguousarray(seed, T.FEATURE)
    out = np.empty(len(q), np.float32)
    lib().orc_seed_distances(_ptr(q), len(q), _ptr(seed), len(seed), _ptr(out))
    return out


def match_winners(p: Params, q, t, seed_dist=None):
    q = np.ascontiguousarray(q, T.FEATURE)
    t = np.ascontiguousarray(t, T.FEATURE)
    bd = np.empty(len(q), np.float32)
    bi = np.empty(len(q), np.int32)
    ok = np.empty(len(q), np.uint8)
    sd = None if seed_dist is None else np.ascontiguousarray(seed_dist, np.float32)
    lib().orc_match_winners(C.byref(p), _ptr(q), len(q), _ptr(t), len(t), None if sd is None else _ptr(sd),
                            _ptr(bd), _ptr(bi), _ptr(ok))
    return bd, bi, ok


def generate_matches(p: Params, q, t, seed_dist=None) -> np.ndarray:
    q = np.ascontiguousarray(q, T.FEATURE)
    t = np.ascontiguousarray(t, T.FEATURE)
    out = np.zeros(len(q), T.OUT_DTYPES[p.out_kind])
    sd = None if seed_dist is None else np.ascontiguousarray(seed_dist, np.float32)
    n = lib().orc_generate_matches(C.byref(p), _ptr(q), len(q), _ptr(t), len(t), None if sd is None else _ptr(sd), _ptr(out))
    return out[:n].copy()


def refine(m: np.ndarray, threshold: float) -> np.ndarray:
    m = np.ascontiguousarray(m, T.DMATCH).copy()
    n = lib().orc_refine(_ptr(m), len(m), float(np.float32(threshold)))
    return m[:n].copy()


def sort(m: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(m, T.DMATCH).copy()
    lib().orc_sort(_ptr(m), len(m))
    return m


def raw_matches(m: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(m, T.DMATCH)
    out = np.zeros(len(m), T.MATCH)
    lib().orc_raw_matches(_ptr(m), len(m), _ptr(out))
    return out


def fuse_nview(features: list, pair_lists: list):
    """features: list of FEATURE arrays per view; pair_lists: list of UINT2_PAIR arrays in reference pair order."""
    feats = [np.ascontiguousarray(f, T.FEATURE) for f in features]
    n_views = len(feats)
    nfeat = np.array([len(f) for f in feats], np.uint64)
    ptrs = (C.c_void_p * n_views)(*[f.ctypes.data for f in feats])
    counts = np.array([len(p) for p in pair_lists], np.uint64)
    pairs = np.ascontiguousarray(np.concatenate(pair_lists) if pair_lists else np.zeros(0, T.UINT2_PAIR), T.UINT2_PAIR)
    nkp = C.c_uint64(0)
    n_mm = lib().orc_fuse_nview(n_views, _ptr(nfeat), ptrs, _ptr(pairs), _ptr(counts), len(pair_lists), None, None, C.byref(nkp))
    mm = np.zeros(n_mm, T.MULTIMATCH)
    kp = np.zeros(nkp.value, T.KEYPOINT)
    lib().orc_fuse_nview(n_views, _ptr(nfeat), ptrs, _ptr(pairs), _ptr(counts), len(pair_lists), _ptr(mm), _ptr(kp), C.byref(nkp))
    return mm, kp
