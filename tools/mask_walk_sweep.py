#!/usr/bin/env python
"""Masked searches WITHOUT seed distances on the reference's everest fixture (32 425 x 36 251): matcher launch time and record hash for a
sweep of the rare path's two knobs (ssrlcv_b200_tune MASK_COOPERATIVE / MASK_WALK; walk = 0 with cooperative = 10 is the path as it was
before the candidate walk).  Records must not depend on the knobs."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ssrlcv_b200 import _lib, types as T  # noqa: E402
from ssrlcv_b200.match_factory import Context  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

ev = load_everest()
q, t = ev.views2[0], ev.views2[1]
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
qd = torch.from_numpy(q.view(np.uint8).reshape(-1)).cuda()
td = torch.from_numpy(t.view(np.uint8).reshape(-1)).cuda()
out = torch.zeros(len(q) * 48, dtype=torch.uint8, device="cuda")
F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
hq, ht = C.c_void_p(), C.c_void_p()
lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq))
lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht))
lib.ssrlcv_b200_set_profiling(h, 1)
k_ms, k_n, n_out = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
settings = [(12, 32)]
for mask in (2, 1):
    p = _lib.MatchParams()
    p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, mask, ev.abs_threshold, ev.rel_threshold, 1
    p.epsilon, p.delta = ev.epsilon, ev.delta
    p.query_image_id, p.target_image_id = ev.img_ids[0], ev.img_ids[1]
    p.query_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(ev.cameras[0]).tobytes())
    p.target_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(ev.cameras[1]).tobytes())
    for i in range(9):
        p.fundamental[i] = float(F.reshape(-1)[i])
    for colmask in (-1,):
        lib.ssrlcv_b200_tune(h, _lib.TUNE_COLUMN_MASKS, colmask)
        for coop, walk in settings:
            lib.ssrlcv_b200_tune(h, _lib.TUNE_MASK_COOPERATIVE, coop)
            lib.ssrlcv_b200_tune(h, _lib.TUNE_MASK_WALK, walk)
            for it in range(3):
                lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
            lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
            for it in range(10):
                lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
            lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
            rec = out[: n_out.value * 48].cpu().numpy().view(T.DMATCH)
            digest = hashlib.sha1(np.ascontiguousarray(rec["distance"]).tobytes() + np.ascontiguousarray(rec["keyPoints"]["loc"]).tobytes()).hexdigest()[:12]
            print(f"mask {mask} colmask {colmask:2d} cooperative<={coop:2d} walk<={walk:2d}: matcher kernel {1e3 * k_ms.value / k_n.value:7.1f} us, records {n_out.value} {digest}", flush=True)
