"""Loader for tests/golden/everest.npz (made by tools/make_golden.py from the reference's fixtures)."""
import os
from dataclasses import dataclass, field

import numpy as np

from ssrlcv_b200 import types as T

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "everest.npz")


@dataclass
class Everest:
    seed3: np.ndarray
    views3: list          # 3 FEATURE arrays (3-view fixture)
    seed2: np.ndarray
    views2: list          # 2 FEATURE arrays (2-view fixture)
    cameras: list         # 3 CAMERA scalars
    img_ids: list
    g2_keypoints: np.ndarray
    g2_multimatch: np.ndarray
    g3_keypoints: np.ndarray
    g3_multimatch: np.ndarray
    # the reference's parameters for these tests: test/Pipeline.cu:198,237 and src/Pipeline.cu:175
    epsilon: float = 25.0
    delta: float = 5.0
    abs_threshold: float = 200.0 * 200.0
    rel_threshold: float = 0.6


def _features(z, i, patch=None):
    f = np.zeros(len(z[f"v{i}_values"]), T.FEATURE)
    for n in ("values", "loc", "sigma", "theta", "parent"):
        f[n] = z[f"v{i}_{n}"]
    if patch is not None:
        flat = f["values"].reshape(-1).copy()
        flat[z[f"{patch}_v{i}_idx"]] = z[f"{patch}_v{i}_val"]
        f["values"] = flat.reshape(-1, 128)
    return f


def _records(dt, **cols):
    n = len(next(iter(cols.values())))
    r = np.zeros(n, dt)
    for k, v in cols.items():
        r[k] = v
    return r


def load_everest() -> Everest:
    z = np.load(_PATH)
    cams = [np.frombuffer(z[f"cam{i}"].tobytes(), T.CAMERA)[0] for i in range(3)]
    return Everest(
        seed3=_features(z, -1),
        views3=[_features(z, i) for i in range(3)],
        seed2=_features(z, -1, "p2"),
        views2=[_features(z, i, "p2") for i in range(2)],
        cameras=cams,
        img_ids=[int(x) for x in z["img_ids"]],
        g2_keypoints=_records(T.KEYPOINT, parentId=z["g2_kp_parent"], loc=z["g2_kp_loc"]),
        g2_multimatch=_records(T.MULTIMATCH, numKeyPoints=z["g2_mm_n"], index=z["g2_mm_index"]),
        g3_keypoints=_records(T.KEYPOINT, parentId=z["g3_kp_parent"], loc=z["g3_kp_loc"]),
        g3_multimatch=_records(T.MULTIMATCH, numKeyPoints=z["g3_mm_n"], index=z["g3_mm_index"]),
    )
