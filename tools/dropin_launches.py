#!/usr/bin/env python
"""The reference's 2-view fixture through the drop-in MatchFactory (oracle/_ref/libssrlcv_dropin.so) -- for an ncu launch list."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssrlcv_b200 import types as T  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

ev = load_everest()
q, t, s = ev.views2[0], ev.views2[1], ev.seed2
os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
os.chdir("/tmp/ssrlcv_ref_run")
lib = C.CDLL(os.path.join(ROOT, "oracle/_ref", sys.argv[1] if len(sys.argv) > 1 else "libssrlcv_dropin.so"),
             mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0))
lib.ref_match.restype = C.c_long
out = np.zeros(len(q), T.DMATCH)
sd = np.zeros(len(q), np.float32)
for it in range(2):
    ms, ms2 = C.c_double(0), C.c_double(0)
    lib.ref_seed_distances(q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), s.ctypes.data_as(C.c_void_p), C.c_ulong(len(s)), 2,
                           sd.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ms2))
    n = lib.ref_match(1, 2, C.c_float(ev.abs_threshold), C.c_float(ev.rel_threshold), ev.img_ids[0], ev.img_ids[1],
                      q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                      sd.ctypes.data_as(C.POINTER(C.c_float)), 2, C.c_float(ev.epsilon), C.c_float(ev.delta), None,
                      np.asarray(ev.cameras[0]).tobytes(), np.asarray(ev.cameras[1]).tobytes(), out.ctypes.data_as(C.c_void_p), C.byref(ms))
    n2 = lib.ref_match(1, 0, C.c_float(ev.abs_threshold), C.c_float(ev.rel_threshold), 0, 1,
                       q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                       None, 2, C.c_float(0), C.c_float(0), None,
                       np.asarray(ev.cameras[0]).tobytes(), np.asarray(ev.cameras[1]).tobytes(), out.ctypes.data_as(C.c_void_p), C.byref(ms2))
    print("masked+seeded", n, round(ms.value, 3), "ms; unmasked unseeded", n2, round(ms2.value, 3), "ms")
