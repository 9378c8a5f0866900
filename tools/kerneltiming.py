#!/usr/bin/env python
"""Per-role cycle breakdown of the matcher kernel (uses the SSRLCV_B200_TIMING diagnostics hook)."""
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

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
q, t, s = synth_inputs(w, 0)
qd = torch.from_numpy(q.view(np.uint8).reshape(-1)).cuda()
td = torch.from_numpy(t.view(np.uint8).reshape(-1)).cuda()
out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
tim = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
p = _lib.MatchParams()
p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, 0, 4e4, 0.6, 1
n_out = C.c_uint64(0)
hq, ht = C.c_void_p(), C.c_void_p()
lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq))
lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht))
for it in range(3):
    if it == 2:
        os.environ["SSRLCV_B200_TIMING"] = hex(tim.data_ptr())
    lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
torch.cuda.synchronize()
x = tim.cpu().numpy().reshape(148, 16)
names = ["mma:wait_t_empty", "mma:wait_full", "mma:total", "tma:wait_empty", "epi0:wait_t_full", "epi0:work", "epi0:resolve", "epi0:total",
         "epi5:wait_t_full", "epi5:work", "epi5:resolve", "epi5:total", "res0:wait", "res0:work"]
for i, n in enumerate(names):
    col = x[:, i]
    print(f"{n:18s} mean {col.mean():12.0f}  min {col.min():12d}  max {col.max():12d}")
