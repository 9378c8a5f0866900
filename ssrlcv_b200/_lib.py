"""ctypes binding of libssrlcv_match_b200.so (the C ABI in include/ssrlcv_b200.h).

There is no fallback: if the shared library is missing or a call fails, an
exception is raised.  The library itself refuses to create a context without
an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SSRLCV_B200_LIB: another build of the same library (A/B timing of kernel changes); never a different implementation
LIB_PATH = os.environ.get("SSRLCV_B200_LIB") or os.path.join(HERE, "libssrlcv_match_b200.so")

ERRORS = {-1: "EINVAL", -2: "ENODEVICE", -3: "ESEEDSIZE", -4: "ETHRESHOLD", -5: "ENOMEM"}

# every symbol include/ssrlcv_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "ssrlcv_b200_abi_version", "ssrlcv_b200_create", "ssrlcv_b200_destroy", "ssrlcv_b200_last_error",
    "ssrlcv_b200_synchronize", "ssrlcv_b200_launch_count", "ssrlcv_b200_stream", "ssrlcv_b200_tune",
    "ssrlcv_b200_multi_create", "ssrlcv_b200_multi_destroy", "ssrlcv_b200_multi_device_count", "ssrlcv_b200_multi_last_error", "ssrlcv_b200_multi_device_ms",
    "ssrlcv_b200_match_pairs_multi_gpu", "ssrlcv_b200_multi_assignment", "ssrlcv_b200_set_profiling", "ssrlcv_b200_kernel_time",
    "ssrlcv_b200_pack", "ssrlcv_b200_pack_batch", "ssrlcv_b200_pack_host", "ssrlcv_b200_set_alloc", "ssrlcv_b200_set_arrays_of", "ssrlcv_b200_set_commit",
    "ssrlcv_b200_set_size", "ssrlcv_b200_set_free", "ssrlcv_b200_seed_distances", "ssrlcv_b200_match",
    "ssrlcv_b200_match_host", "ssrlcv_b200_match_features", "ssrlcv_b200_match_batch", "ssrlcv_b200_match_candidates", "ssrlcv_b200_match_candidates_seeded", "ssrlcv_b200_match_runner_up", "ssrlcv_b200_finalize_candidates",
    "ssrlcv_b200_set_loc", "ssrlcv_b200_validate", "ssrlcv_b200_refine", "ssrlcv_b200_sort",
    "ssrlcv_b200_raw_matches", "ssrlcv_b200_match_set_2view", "ssrlcv_b200_epipolar_endpoints", "ssrlcv_b200_device_sincos", "ssrlcv_b200_set_download",
]


class MatchParams(C.Structure):
    _fields_ = [
        ("out_kind", C.c_int32), ("mask_kind", C.c_int32),
        ("absolute_threshold", C.c_float), ("relative_threshold", C.c_float),
        ("query_image_id", C.c_uint32), ("target_image_id", C.c_uint32),
        ("epsilon", C.c_float), ("delta", C.c_float),
        ("fundamental", C.c_float * 9),
        ("query_camera", C.c_uint8 * 80), ("target_camera", C.c_uint8 * 80),
        ("compact", C.c_int32), ("reserved", C.c_int32),
    ]


# ssrlcv_b200_tune knobs (include/ssrlcv_b200.h)
TUNE_FOLD, TUNE_COLUMN_MASKS, TUNE_SORTED_TARGETS, TUNE_HOST_CHUNK_QUERY_ROWS, TUNE_HOST_CHUNK_TARGET_ROWS, TUNE_CROSSCHECK_SIMT, TUNE_TIMING_BUFFER, TUNE_DIAG, TUNE_PAIR_UNITS, TUNE_MASK_COOPERATIVE, TUNE_MASK_WALK, TUNE_CELL_SORT, TUNE_SEED_ONLY, TUNE_LEAN_MASKED = range(14)


class SetArrays(C.Structure):
    _fields_ = [("desc", C.c_void_p), ("loc", C.c_void_p), ("sigma_theta", C.c_void_p), ("n", C.c_uint64), ("n_pad", C.c_uint64)]


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ssrlcv_b200 error {code} ({ERRORS.get(code, 'cuda' if code > 0 else '?')}): {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ssrlcv_b200.build` (needs nvcc); "
            "there is no CPU fallback for the matching path"
        )
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32, u32, f32 = C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_float
    pp = C.POINTER(MatchParams)
    pu64 = C.POINTER(C.c_uint64)
    lib.ssrlcv_b200_abi_version.restype = i32
    lib.ssrlcv_b200_create.argtypes = [C.POINTER(vp), i32, vp]
    lib.ssrlcv_b200_destroy.argtypes = [vp]
    lib.ssrlcv_b200_destroy.restype = None
    lib.ssrlcv_b200_last_error.argtypes = [vp]
    lib.ssrlcv_b200_last_error.restype = C.c_char_p
    lib.ssrlcv_b200_synchronize.argtypes = [vp]
    lib.ssrlcv_b200_launch_count.argtypes = [vp]
    lib.ssrlcv_b200_launch_count.restype = u64
    lib.ssrlcv_b200_stream.argtypes = [vp]
    lib.ssrlcv_b200_stream.restype = vp
    lib.ssrlcv_b200_tune.argtypes = [vp, i32, C.c_int64]
    pu32 = C.POINTER(C.c_uint32)
    lib.ssrlcv_b200_multi_create.argtypes = [C.POINTER(vp), i32, C.POINTER(C.c_int)]
    lib.ssrlcv_b200_multi_destroy.argtypes = [vp]
    lib.ssrlcv_b200_multi_destroy.restype = None
    lib.ssrlcv_b200_multi_device_count.argtypes = [vp]
    lib.ssrlcv_b200_multi_last_error.argtypes = [vp]
    lib.ssrlcv_b200_multi_last_error.restype = C.c_char_p
    lib.ssrlcv_b200_multi_device_ms.argtypes = [vp, i32]
    lib.ssrlcv_b200_multi_device_ms.restype = C.c_double
    lib.ssrlcv_b200_match_pairs_multi_gpu.argtypes = [vp, u32, C.POINTER(vp), pu64, vp, u64, u32, pu32, pu32, pp, C.POINTER(vp), pu64]
    lib.ssrlcv_b200_multi_assignment.argtypes = [vp, u32, pu64, u32, pu32, pu32, C.POINTER(C.c_int32)]
    lib.ssrlcv_b200_set_profiling.argtypes = [vp, i32]
    lib.ssrlcv_b200_kernel_time.argtypes = [vp, C.POINTER(C.c_double), pu64]
    lib.ssrlcv_b200_pack.argtypes = [vp, vp, u64, C.POINTER(vp)]
    lib.ssrlcv_b200_pack_batch.argtypes = [vp, u32, C.POINTER(vp), pu64, C.POINTER(vp)]
    lib.ssrlcv_b200_pack_host.argtypes = [vp, vp, u64, C.POINTER(vp)]
    lib.ssrlcv_b200_set_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    lib.ssrlcv_b200_set_arrays_of.argtypes = [vp, C.POINTER(SetArrays)]
    lib.ssrlcv_b200_set_commit.argtypes = [vp, vp]
    lib.ssrlcv_b200_set_size.argtypes = [vp]
    lib.ssrlcv_b200_set_size.restype = u64
    lib.ssrlcv_b200_set_free.argtypes = [vp]
    lib.ssrlcv_b200_set_free.restype = None
    lib.ssrlcv_b200_set_loc.argtypes = [vp]
    lib.ssrlcv_b200_set_loc.restype = vp
    lib.ssrlcv_b200_seed_distances.argtypes = [vp, vp, vp, vp]
    lib.ssrlcv_b200_match.argtypes = [vp, pp, vp, vp, vp, u64, vp, pu64]
    lib.ssrlcv_b200_match_batch.argtypes = [vp, u32, pp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), pu64]
    lib.ssrlcv_b200_match_host.argtypes = [vp, pp, vp, u64, vp, u64, vp, u64, vp, pu64]
    lib.ssrlcv_b200_match_features.argtypes = [vp, pp, vp, u64, vp, u64, vp, u64, vp, vp, pu64]
    lib.ssrlcv_b200_match_candidates.argtypes = [vp, pp, vp, vp, u64, vp]
    lib.ssrlcv_b200_match_candidates_seeded.argtypes = [vp, pp, vp, vp, u64, vp, vp]
    lib.ssrlcv_b200_match_runner_up.argtypes = [vp, pp, vp, vp, vp, vp]
    lib.ssrlcv_b200_finalize_candidates.argtypes = [vp, pp, vp, vp, u32, vp, vp, vp, pu64]
    lib.ssrlcv_b200_validate.argtypes = [vp, i32, vp, u64, pu64]
    lib.ssrlcv_b200_refine.argtypes = [vp, vp, u64, f32, pu64]
    lib.ssrlcv_b200_sort.argtypes = [vp, vp, u64]
    lib.ssrlcv_b200_raw_matches.argtypes = [vp, vp, u64, vp]
    lib.ssrlcv_b200_match_set_2view.argtypes = [vp, vp, u64, vp, vp, C.POINTER(C.c_float)]
    lib.ssrlcv_b200_epipolar_endpoints.argtypes = [vp, pp, vp, vp]
    lib.ssrlcv_b200_device_sincos.argtypes = [vp, vp, u64, vp, vp]
    lib.ssrlcv_b200_set_download.argtypes = [vp, vp, i32, vp, u64, pu64]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ssrlcv_b200_abi_version",):
            fn.restype = i32
    _lib = lib
    return lib


def check(ctx_handle, rc: int) -> None:
    if rc != 0:
        msg = load().ssrlcv_b200_last_error(ctx_handle)
        raise B200Error(rc, msg.decode() if msg else "")
