#!/usr/bin/env python
"""One call type of the everest fixture through the drop-in, repeated -- for an ncu launch list.  argv: out_kind mask seeded"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssrlcv_b200 import types as T  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

out_kind, mask, seeded = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ev = load_everest()
q, t, s = ev.views2[0], ev.views2[1], ev.seed2
os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
os.chdir("/tmp/ssrlcv_ref_run")
lib = C.CDLL(os.path.join(ROOT, "oracle/_ref", "libssrlcv_dropin.so"), mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0))
lib.ref_match.restype = C.c_long
out = np.zeros(len(q), T.OUT_DTYPES[out_kind])
sd = np.zeros(len(q), np.float32)
F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
lib.ref_seed_distances(q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), s.ctypes.data_as(C.c_void_p), C.c_ulong(len(s)), 2, sd.ctypes.data_as(C.POINTER(C.c_float)), None)
for it in range(3):
    ms = C.c_double(0)
    n = lib.ref_match(out_kind, mask, C.c_float(ev.abs_threshold), C.c_float(ev.rel_threshold), ev.img_ids[0], ev.img_ids[1],
                      q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                      sd.ctypes.data_as(C.POINTER(C.c_float)) if seeded else None, 2, C.c_float(ev.epsilon), C.c_float(ev.delta),
                      F.ctypes.data_as(C.POINTER(C.c_float)), np.asarray(ev.cameras[0]).tobytes(), np.asarray(ev.cameras[1]).tobytes(),
                      out.ctypes.data_as(C.c_void_p), C.byref(ms))
    print("call", it, "records", n, round(ms.value, 3), "ms", flush=True)
