#!/usr/bin/env python
"""Host-side latency of each C-ABI call for one workload (perf_counter around each call, stream synchronised after)."""
import ctypes as C
import os
import sys
import time

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
p = _lib.MatchParams()
p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, 0, 4e4, 0.6, 1
n_out = C.c_uint64(0)


def timed(name, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    acc.setdefault(name, []).append(((t1 - t0) * 1e6, (t2 - t0) * 1e6))
    return r


acc = {}
for it in range(8):
    hq, ht = C.c_void_p(), C.c_void_p()
    timed("pack(q)", lambda: lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq)), acc)
    timed("pack(t)", lambda: lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht)), acc)
    timed("match", lambda: lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out)), acc)
    timed("free(q)", lambda: lib.ssrlcv_b200_set_free(hq), acc)
    timed("free(t)", lambda: lib.ssrlcv_b200_set_free(ht), acc)
for k, v in acc.items():
    v = v[2:]
    print(f"{k:10s} call returns after {np.median([a for a, _ in v]):9.1f} us ; done+synced after {np.median([b for _, b in v]):9.1f} us")
# back-to-back steps, wall clock
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(20):
    hq, ht = C.c_void_p(), C.c_void_p()
    lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq))
    lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht))
    lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
    lib.ssrlcv_b200_set_free(hq); lib.ssrlcv_b200_set_free(ht)
torch.cuda.synchronize()
print(f"20 back-to-back steps: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per step (wall), matches {n_out.value}")
