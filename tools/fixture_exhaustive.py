#!/usr/bin/env python
"""generateMatchesExhaustive on the reference's 3-view fixture (everest, test/checkpoints/Pipeline3View; GEO_ORBIT=1 pair loop +
adjacency fusion, ref: src/MatchFactory.cu:907-1028): drop-in vs the reference's own MatchFactory.cu, same harness, state `both`."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssrlcv_b200 import types as T  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

ev = load_everest()
os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
os.chdir("/tmp/ssrlcv_ref_run")
n = 3
feats = (C.c_void_p * n)(*[v.ctypes.data for v in ev.views3])
nf = (C.c_ulong * n)(*[len(v) for v in ev.views3])
cam_bytes = [np.asarray(c).tobytes() for c in ev.cameras]
cams = (C.c_char_p * n)(*cam_bytes)
ids = (C.c_int * n)(*ev.img_ids)
print("views", [len(v) for v in ev.views3], "seed", len(ev.seed3))
for so in ("libssrlcv_dropin.so", "libssrlcv_ref.so"):
    lib = C.CDLL(os.path.join(ROOT, "oracle/_ref", so), mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0))
    lib.ref_exhaustive.restype = C.c_long
    times = []
    for it in range(3):
        mm = np.zeros(40000, T.MULTIMATCH)
        kp = np.zeros(120000, T.KEYPOINT)
        nkp, ms = C.c_ulong(0), C.c_double(0)
        nmm = lib.ref_exhaustive(n, feats, nf, C.cast(cams, C.POINTER(C.c_void_p)), ids, ev.seed3.ctypes.data_as(C.c_void_p), C.c_ulong(len(ev.seed3)),
                                 C.c_float(ev.rel_threshold), C.c_float(ev.abs_threshold), C.c_float(ev.epsilon), C.c_float(ev.delta), 3,
                                 mm.ctypes.data_as(C.c_void_p), C.c_ulong(len(mm)), kp.ctypes.data_as(C.c_void_p), C.c_ulong(len(kp)), C.byref(nkp), C.byref(ms))
        times.append(ms.value)
    print(f"{so:24s} generateMatchesExhaustive {np.median(times[1:]):9.3f} ms   multi-matches {nmm}, keypoints {nkp.value}")
