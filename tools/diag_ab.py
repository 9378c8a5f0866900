#!/usr/bin/env python
"""What-if timing of the matcher kernel at C3 (ssrlcv_b200_tune(SSRLCV_B200_TUNE_DIAG)): how fast would the launch be with
half the L2->SM target-tile traffic / without the fold MMA / without the epilogue's max trees?  Results are wrong for every
non-zero setting by construction; only the kernel time and the in-kernel clock are read.  Needs a diagnostics build of the library
(SSRLCV_B200_DIAG_BUILD=1 python -m ssrlcv_b200.build --force): the product build compiles the hooks out."""
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
settings = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 7, 0]
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
q, t, s = synth_inputs(w, 0)
if os.environ.get("DIAG_ZERO"):      # operand toggling: all-zero descriptors
    q["values"][:] = 0
    t["values"][:] = 0
qd = torch.from_numpy(q.view(np.uint8).reshape(-1)).cuda()
td = torch.from_numpy(t.view(np.uint8).reshape(-1)).cuda()
out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tim = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
p = _lib.MatchParams()
p.out_kind, p.mask_kind, p.absolute_threshold, p.relative_threshold, p.compact = T.OUT_DMATCH, 0, 4e4, 0.6, 1
n_out = C.c_uint64(0)
hq, ht = C.c_void_p(), C.c_void_p()
lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), len(q), C.byref(hq))
lib.ssrlcv_b200_pack(h, C.c_void_p(td.data_ptr()), len(t), C.byref(ht))
lib.ssrlcv_b200_set_profiling(h, 1)
k_ms, k_n = C.c_double(0), C.c_uint64(0)
pairs = [int(x) for x in os.environ.get("DIAG_PAIR", "0,1").split(",")]
if os.environ.get("DIAG_COLMASK"):
    lib.ssrlcv_b200_tune(h, _lib.TUNE_COLUMN_MASKS, int(os.environ["DIAG_COLMASK"]))
for fold, pair in [(f, pr) for f in (1, 0) for pr in pairs]:
    lib.ssrlcv_b200_tune(h, _lib.TUNE_FOLD, fold)
    lib.ssrlcv_b200_tune(h, _lib.TUNE_PAIR_UNITS, pair)
    for d in settings:
        if d and lib.ssrlcv_b200_tune(h, _lib.TUNE_DIAG, d):
            print(f"diag={d}: skipped (product build of the library)")
            continue
        if not d:
            lib.ssrlcv_b200_tune(h, _lib.TUNE_DIAG, 0)
        lib.ssrlcv_b200_tune(h, _lib.TUNE_TIMING_BUFFER, 0)
        for it in range(3):
            lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
        lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
        for it in range(20):
            flush.fill_(it)
            torch.cuda.synchronize()
            lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
        lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
        ms = k_ms.value / k_n.value
        # in-kernel clock: the instrumented build's MMA-warp cycle count over its event time
        tim.zero_()
        lib.ssrlcv_b200_tune(h, _lib.TUNE_TIMING_BUFFER, tim.data_ptr())
        for it in range(3):
            lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out))
        lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
        x = tim.cpu().numpy().reshape(148, 16)
        if pair == 2:
            x = x[0::2]                # the leader CTA of each pair
        ims = k_ms.value / k_n.value
        print(f"fold={fold} pair={pair} diag={d}: kernel {1e3 * ms:8.2f} us  matches {n_out.value:6d} | instrumented {1e3 * ims:8.2f} us, mma cycles {x[:, 2].mean():9.0f} "
              f"-> {x[:, 2].mean() / (1e3 * ims):6.1f} MHz, mma wait_t_empty {x[:, 0].mean():8.0f} wait_full {x[:, 1].mean():7.0f}, epi0 wait {x[:, 4].mean():8.0f} work {x[:, 5].mean():8.0f} cwait|rare {x[:, 15].mean():7.0f} events {x[:, 14].mean():5.1f}" + (f" | ld {x[:, 8].mean():8.0f} hand-back {x[:, 9].mean():7.0f} " if pair == 2 else ""), flush=True)
