#!/usr/bin/env python
"""Seed pass (getSeedDistances) matcher launch time against the seed-set size, distances-only variant on / off (synthetic descriptors:
an uncapped search is bookkeeping-bound whatever the data)."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ssrlcv_b200 import _lib, types as T
from ssrlcv_b200.match_factory import MatchFactory
from ssrlcv_b200.unity import Unity, cpu, gpu
from tests import synth
nq = 32768
for ns in (8192, 16384, 32768, 65536):
    q, s = synth.make_pair(nq, ns, seed=ns)
    qu = Unity(q, T.FEATURE, cpu); qu.setMemoryState(gpu)
    for so in (1, 0):
        mf = MatchFactory(0.6, 4e4)
        mf.ctx.tune(_lib.TUNE_SEED_ONLY, so)
        mf.setSeedFeatures(Unity(s, T.FEATURE, cpu))
        lib, h = mf.ctx.lib, mf.ctx.handle
        for _ in range(3): mf.getSeedDistances(qu)
        lib.ssrlcv_b200_set_profiling(h, 1)
        for _ in range(10): r = mf.getSeedDistances(qu)
        torch.cuda.synchronize()
        k_ms, k_n = C.c_double(0), C.c_uint64(0); lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n)); lib.ssrlcv_b200_set_profiling(h, 0)
        print(f"{nq} x {ns} distances-only {so}: matcher {k_ms.value / max(k_n.value, 1) * 1e3:8.1f} us  ({nq * ns / (k_ms.value / max(k_n.value, 1) * 1e-3) / 1e12:.2f}e12 evals/s)", flush=True)
mf.ctx.tune(_lib.TUNE_SEED_ONLY, -1)
