"""Binary layouts of the reference's records on the matching path.

Every dtype below is byte-compatible with the corresponding C++ struct of the
reference (sizes/offsets measured with nvcc against its headers, see SURVEY.md
section 8a):

  Feature<SIFT_Descriptor>  include/Feature.cuh:30-38,68-94      152 B
  Image::Camera             include/Image.cuh:41-57               80 B
  KeyPoint / Match / DMatch / FeatureMatch<SIFT> / uint2_pair / MultiMatch
                            include/MatchFactory.cuh:23-72        16/40/48/320/16/8 B

Padding bytes are named `_pad*`; the reference leaves them uninitialised, this
framework writes zeros.  Never compare records with memcmp: compare fields.
"""
from __future__ import annotations

import numpy as np

FEATURE = np.dtype(
    {
        "names": ["parent", "_pad", "loc", "sigma", "theta", "values"],
        "formats": ["<i4", "<i4", ("<f4", 2), "<f4", "<f4", ("u1", 128)],
        "offsets": [0, 4, 8, 16, 20, 24],
        "itemsize": 152,
    }
)

CAMERA = np.dtype(
    {
        "names": ["cam_pos", "cam_rot", "fov", "foc", "dpix", "timeStamp", "ecef_offset", "no_rot", "size"],
        "formats": [("<f4", 3), ("<f4", 3), ("<f4", 2), "<f4", ("<f4", 2), "<i8", ("<f4", 3), "u1", ("<u4", 2)],
        "offsets": [0, 12, 24, 32, 40, 48, 56, 68, 72],
        "itemsize": 80,
    }
)

KEYPOINT = np.dtype(
    {"names": ["parentId", "loc"], "formats": ["<i4", ("<f4", 2)], "offsets": [0, 8], "itemsize": 16}
)

MATCH = np.dtype(
    {"names": ["invalid", "keyPoints"], "formats": ["u1", (KEYPOINT, 2)], "offsets": [0, 8], "itemsize": 40}
)

DMATCH = np.dtype(
    {
        "names": ["invalid", "keyPoints", "distance"],
        "formats": ["u1", (KEYPOINT, 2), "<f4"],
        "offsets": [0, 8, 40],
        "itemsize": 48,
    }
)

SIFT_DESCRIPTOR = np.dtype(
    {"names": ["sigma", "theta", "values"], "formats": ["<f4", "<f4", ("u1", 128)], "offsets": [0, 4, 8], "itemsize": 136}
)

FEATUREMATCH = np.dtype(
    {
        "names": ["invalid", "keyPoints", "distance", "descriptors"],
        "formats": ["u1", (KEYPOINT, 2), "<f4", (SIFT_DESCRIPTOR, 2)],
        "offsets": [0, 8, 40, 44],
        "itemsize": 320,
    }
)

UINT2_PAIR = np.dtype({"names": ["a", "b"], "formats": [("<u4", 2), ("<u4", 2)], "offsets": [0, 8], "itemsize": 16})

MULTIMATCH = np.dtype({"names": ["numKeyPoints", "index"], "formats": ["<u4", "<i4"], "offsets": [0, 4], "itemsize": 8})

# C-ABI enums (include/ssrlcv_b200.h)
OUT_MATCH, OUT_DMATCH, OUT_PAIR, OUT_FEATUREMATCH = 0, 1, 2, 3
MASK_NONE, MASK_FUNDAMENTAL, MASK_DOUBLE = 0, 1, 2

OUT_DTYPES = {OUT_MATCH: MATCH, OUT_DMATCH: DMATCH, OUT_PAIR: UINT2_PAIR, OUT_FEATUREMATCH: FEATUREMATCH}


def make_features(values: np.ndarray, loc: np.ndarray | None = None, parent: int = -1) -> np.ndarray:
    """Build a Feature<SIFT_Descriptor>[n] array from uint8[n,128] descriptors (+ float32[n,2] locations)."""
    values = np.ascontiguousarray(values, dtype=np.uint8)
    n = values.shape[0]
    out = np.zeros(n, dtype=FEATURE)
    out["values"] = values
    out["parent"] = parent
    if loc is None:
        out["loc"] = -1.0
    else:
        out["loc"] = np.asarray(loc, dtype=np.float32)
    return out


def fields_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Field-wise equality of two record arrays of the same dtype (ignores `_pad*` bytes, bit-exact floats)."""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False

    def eq(x: np.ndarray, y: np.ndarray) -> bool:
        if x.dtype.names:
            return all(eq(x[n], y[n]) for n in x.dtype.names if not n.startswith("_pad"))
        if x.dtype.kind == "f":
            return np.array_equal(x.view(f"u{x.dtype.itemsize}"), y.view(f"u{y.dtype.itemsize}"))
        return np.array_equal(x, y)

    return eq(a, b)
