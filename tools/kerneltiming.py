#!/usr/bin/env python
"""Per-role cycle breakdown of the matcher kernel (ssrlcv_b200_tune(SSRLCV_B200_TUNE_TIMING_BUFFER): the instrumented build of the kernel)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, synth_inputs  # noqa: E402
from ssrlcv_b200 import _lib, types as T  # noqa: E402
from ssrlcv_b200.match_factory import Context  # noqa: E402

# usage: kerneltiming.py [C1|C2|C3]   or   kerneltiming.py NQ NT [ABS_THRESHOLD [MASK_KIND [EPSILON]]]   (fold variant)
abs_thr, eps = 4e4, 25.0
if len(sys.argv) > 2:
    w = dict(nq=int(sys.argv[1]), nt=int(sys.argv[2]), ns=0, mask=0)
    if len(sys.argv) > 3:
        abs_thr = float(sys.argv[3])
    if len(sys.argv) > 4:
        w["mask"] = int(sys.argv[4])
    if len(sys.argv) > 5:
        eps = float(sys.argv[5])
else:
    w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
print("workload", w, "abs", abs_thr, "eps", eps)
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
# KT_FOLD / KT_SPLIT: force a matcher variant (ssrlcv_b200_tune)
if os.environ.get("KT_FOLD"):
    lib.ssrlcv_b200_tune(h, _lib.TUNE_FOLD, int(os.environ["KT_FOLD"]))
if os.environ.get("KT_COLMASK"):
    lib.ssrlcv_b200_tune(h, _lib.TUNE_COLUMN_MASKS, int(os.environ["KT_COLMASK"]))
print("fold", os.environ.get("KT_FOLD", "auto"), "colmask", os.environ.get("KT_COLMASK", "auto"))
q, t, s = synth_inputs(w, 0)
qd = torch.from_numpy(q.view(np.uint8).reshape(-1)).cuda()
td = torch.from_numpy(t.view(np.uint8).reshape(-1)).cuda()
out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
tim = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
p = _lib.MatchParams()
p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, w["mask"], abs_thr, 0.6, 1
if w["mask"]:
    from bench import cameras
    cam0, cam1 = cameras()
    p.epsilon, p.delta = eps, 5.0
    p.query_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(cam0).tobytes())
    p.target_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(cam1).tobytes())
n_out = C.c_uint64(0)
hq, ht = C.c_void_p(), C.c_void_p()
lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq))
lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib.ssrlcv_b200_set_profiling(h, 1)
for it in range(6):
    if it == 5:
        lib.ssrlcv_b200_tune(h, 6, tim.data_ptr())
    lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
    if it == 4:
        k_ms, k_n = C.c_double(0), C.c_uint64(0)
        lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
        print(f"production kernel (no counters): {1e3 * k_ms.value / k_n.value:.2f} us per launch over {k_n.value} launches; matches {n_out.value}")
torch.cuda.synchronize()
k_ms, k_n = C.c_double(0), C.c_uint64(0)
lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
print(f"instrumented kernel: {1e3 * k_ms.value / k_n.value:.2f} us")
x = tim.cpu().numpy().reshape(148, 16)
names = ["mma:wait_t_empty", "mma:wait_full", "mma:total", "tma:wait_empty", "epi0:wait_t_full", "epi0:work", "epi0:resolve", "epi0:total",
         "epi5:wait_t_full", "epi5:work", "epi5:resolve", "epi5:total", "res0:wait", "res0:work", "epi0:live_windows (pair walk: rare-path entries)", "epi0:wait_c_full (pair walk, fold: cycles in the rare path)"]
for i, n in enumerate(names):
    col = x[:, i]
    print(f"{n:62s} mean {col.mean():12.0f}  min {col.min():12d}  max {col.max():12d}")
