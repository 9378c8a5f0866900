#!/usr/bin/env python
"""Per-role cycle breakdown of the matcher on the reference's everest fixture (32 425 x 36 251 real SIFT features).
argv: mask kind (0 = none, 1 = F matrix, 2 = DoubleConstrained) [seeded: 0 / 1 -- with the seed distances of the fixture's seed image]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ssrlcv_b200 import _lib, types as T  # noqa: E402
from ssrlcv_b200.match_factory import Context  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

mask = int(sys.argv[1]) if len(sys.argv) > 1 else 2
seeded = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ev = load_everest()
q, t = ev.views2[0], ev.views2[1]
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
qd = torch.from_numpy(q.view(np.uint8).reshape(-1)).cuda()
td = torch.from_numpy(t.view(np.uint8).reshape(-1)).cuda()
out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
tim = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
p = _lib.MatchParams()
p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, mask, ev.abs_threshold, ev.rel_threshold, 1
p.epsilon, p.delta = ev.epsilon, ev.delta
p.query_image_id, p.target_image_id = ev.img_ids[0], ev.img_ids[1]
p.query_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(ev.cameras[0]).tobytes())
p.target_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(ev.cameras[1]).tobytes())
F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
for i in range(9):
    p.fundamental[i] = float(F.reshape(-1)[i])
n_out = C.c_uint64(0)
hq, ht = C.c_void_p(), C.c_void_p()
lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq))
lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht))
sd_ptr, sd_n = None, 0
if seeded:
    sf = torch.from_numpy(ev.seed2.view(np.uint8).reshape(-1)).cuda()
    hs = C.c_void_p()
    lib.ssrlcv_b200_pack(h, C.c_void_p(sf.data_ptr()), len(ev.seed2), C.byref(hs))
    sd = torch.empty(len(q), dtype=torch.float32, device="cuda")
    lib.ssrlcv_b200_seed_distances(h, hq, hs, C.c_void_p(sd.data_ptr()))
    lib.ssrlcv_b200_synchronize(h)
    sd_ptr, sd_n = C.c_void_p(sd.data_ptr()), len(q)
lib.ssrlcv_b200_set_profiling(h, 1)
k_ms, k_n = C.c_double(0), C.c_uint64(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("KT_LEAN"):
    lib.ssrlcv_b200_tune(h, _lib.TUNE_LEAN_MASKED, int(os.environ["KT_LEAN"]))
if os.environ.get("KT_PAIR"):
    lib.ssrlcv_b200_tune(h, _lib.TUNE_PAIR_UNITS, int(os.environ["KT_PAIR"]))
cm_env = os.environ.get("KT_COLMASK")
if cm_env:
    lib.ssrlcv_b200_tune(h, _lib.TUNE_COLUMN_MASKS, int(cm_env))
print("colmask", cm_env or "auto")
for sorted_ in ((1, 0) if (mask and not seeded) else (-1,)):
    lib.ssrlcv_b200_tune(h, _lib.TUNE_SORTED_TARGETS, sorted_)
    lib.ssrlcv_b200_tune(h, _lib.TUNE_TIMING_BUFFER, 0)
    for it in range(3):
        lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, sd_ptr, sd_n, C.c_void_p(out.data_ptr()), C.byref(n_out))
    lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for it in range(10):
        lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, sd_ptr, sd_n, C.c_void_p(out.data_ptr()), C.byref(n_out))
    wall = (time.perf_counter() - t0) / 10
    lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
    print(f"mask {mask} sorted {sorted_}: match call {1e3 * wall:.3f} ms (sorted copies cached on the sets), matcher kernel {1e3 * k_ms.value / k_n.value:.1f} us, records {n_out.value}")
    lib.ssrlcv_b200_tune(h, _lib.TUNE_TIMING_BUFFER, tim.data_ptr())
    for it in range(2):
        lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, sd_ptr, sd_n, C.c_void_p(out.data_ptr()), C.byref(n_out))
    lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
    x = tim.cpu().numpy().reshape(148, 16)
    names = ["mma:wait_t_empty", "mma:wait_full", "mma:total", "tma:wait_empty", "epi0:wait_t_full", "epi0:work", "epi0:resolve", "epi0:total",
             "epi5:wait_t_full", "epi5:work", "epi5:resolve", "epi5:total", "res0:wait", "res0:work", "epi0:live_windows (pair walk: rare-path entries)", "epi0:wait_c_full (pair walk, fold: cycles in the rare path)"]
    print(f"  instrumented kernel {1e3 * k_ms.value / k_n.value:.1f} us")
    for i, n in enumerate(names):
        col = x[:, i]
        print(f"  {n:62s} mean {col.mean():12.0f}  min {col.min():12d}  max {col.max():12d}")
