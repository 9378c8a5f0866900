#!/usr/bin/env python
"""Where the end-to-end time of ssrlcv_b200_match_host goes: each phase of the host-buffer path timed on its own
(perf_counter, stream synchronised), the plain pinned H2D / D2H bandwidth of this box, and the whole call."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ABS_THR, REL_THR, WORKLOADS, bind_to_gpu_numa, synth_inputs  # noqa: E402
from ssrlcv_b200 import _lib, types as T  # noqa: E402
from ssrlcv_b200.match_factory import Context  # noqa: E402

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
if os.environ.get("E2EPROF_NO_BIND") is None:
    print("affinity before", len(os.sched_getaffinity(0)), "cpus; bound:", bind_to_gpu_numa(0) is not None, "->", len(os.sched_getaffinity(0)), "cpus")
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
q, t, s = synth_inputs(w, 0)
pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
qh, th = pin(q), pin(t)
out_h = torch.empty(len(q) * 48, dtype=torch.uint8).pin_memory()
out_d = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
p = _lib.MatchParams()
p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, 0, ABS_THR, REL_THR, 1
n_out = C.c_uint64(0)


def med(fn, n=12):
    xs = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        xs.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(xs[2:]))


# plain copies
d = torch.empty_like(qh, device="cuda")
d2 = torch.empty_like(th, device="cuda")
us2 = med(lambda: d2.copy_(th, non_blocking=True))
print(f"H2D pinned (target buffer) {th.numel() / 1e6:.1f} MB: {us2:.1f} us = {th.numel() / us2 / 1e3:.1f} GB/s")
us = med(lambda: d.copy_(qh, non_blocking=True))
print(f"H2D pinned {qh.numel() / 1e6:.1f} MB: {us:.1f} us = {qh.numel() / us / 1e3:.1f} GB/s")
us = med(lambda: out_h.copy_(out_d, non_blocking=True))
print(f"D2H pinned {out_h.numel() / 1e6:.1f} MB: {us:.1f} us = {out_h.numel() / us / 1e3:.1f} GB/s")

hq, ht = C.c_void_p(), C.c_void_p()


def pack_host(buf, n, hs):
    if hs:
        lib.ssrlcv_b200_set_free(hs)
    ctx.check(lib.ssrlcv_b200_pack_host(h, C.c_void_p(buf.data_ptr()), n, C.byref(hs)))


print(f"pack_host(q)   {med(lambda: pack_host(qh, len(q), hq)):9.1f} us")
print(f"pack_host(t)   {med(lambda: pack_host(th, len(t), ht)):9.1f} us")
print(f"match          {med(lambda: ctx.check(lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out_d.data_ptr()), C.byref(n_out)))):9.1f} us  ({n_out.value} matches)")
nb = int(n_out.value) * 48
print(f"D2H records    {med(lambda: out_h[:nb].copy_(out_d[:nb], non_blocking=True)):9.1f} us")
print(f"match_host     {med(lambda: ctx.check(lib.ssrlcv_b200_match_host(h, C.byref(p), C.c_void_p(qh.data_ptr()), len(q), C.c_void_p(th.data_ptr()), len(t), None, 0, C.c_void_p(out_h.data_ptr()), C.byref(n_out)))):9.1f} us")
