#!/usr/bin/env python
"""Every matching call type on the reference's real 2-view fixture (everest, 32 425 x 36 251 features + 8 932 seed features),
device-resident inputs: the drop-in MatchFactory (oracle/_ref/libssrlcv_dropin.so) next to the reference's own CUDA MatchFactory
(libssrlcv_ref.so), both through the reference's public API in the same harness.  Time = wall time of the API call (ms),
median of 3 after 1 warm-up; records compared field by field."""
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
libs = {}
for name in ("libssrlcv_dropin.so", "libssrlcv_ref.so"):
    lib = C.CDLL(os.path.join(ROOT, "oracle/_ref", name), mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0))
    lib.ref_match.restype = C.c_long
    libs[name] = lib
sd = np.zeros(len(q), np.float32)
libs["libssrlcv_ref.so"].ref_seed_distances(q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), s.ctypes.data_as(C.c_void_p), C.c_ulong(len(s)), 2,
                                            sd.ctypes.data_as(C.POINTER(C.c_float)), None)
F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)


def call(lib, out_kind, mask, seeded, reps=4):
    out = np.zeros(len(q), T.OUT_DTYPES[out_kind])
    times = []
    for _ in range(reps):
        ms = C.c_double(0)
        n = lib.ref_match(out_kind, mask, C.c_float(ev.abs_threshold), C.c_float(ev.rel_threshold), ev.img_ids[0], ev.img_ids[1],
                          q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                          sd.ctypes.data_as(C.POINTER(C.c_float)) if seeded else None, 2, C.c_float(ev.epsilon), C.c_float(ev.delta),
                          F.ctypes.data_as(C.POINTER(C.c_float)), np.asarray(ev.cameras[0]).tobytes(), np.asarray(ev.cameras[1]).tobytes(),
                          out.ctypes.data_as(C.c_void_p), C.byref(ms))
        times.append(ms.value)
    return float(np.median(times[1:])), out[:n].copy()


names = {0: "generateMatches", 1: "generateDistanceMatches", 2: "generateMatchesIndexOnly"}
masks = {0: "", 1: "Constrained(F)", 2: "DoubleConstrained"}
print(f"{'call':58s} {'drop-in ms':>10s} {'reference ms':>12s} {'x':>7s} {'records':>8s}  identical")
for mask in (0, 1, 2):
    for seeded in (False, True):
        for out_kind in ((1, 2) if mask != 2 else (0, 1, 2)):
            a_ms, a = call(libs["libssrlcv_dropin.so"], out_kind, mask, seeded)
            b_ms, b = call(libs["libssrlcv_ref.so"], out_kind, mask, seeded, reps=3)
            label = f"{names[out_kind]} {masks[mask]}{' + seed distances' if seeded else ''}"
            print(f"{label:58s} {a_ms:10.3f} {b_ms:12.3f} {b_ms / a_ms:7.1f} {len(b):8d}  {T.fields_equal(a, b)}")
