#!/usr/bin/env python
"""Unmasked searches on the reference's real fixture (32 425 x 36 251, 8 932 seed features): matcher launch time by bookkeeping regime,
walk and threshold.  On real SIFT data the 'rare' path is not rare (84 % of the windows hold something below 200^2)."""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from ssrlcv_b200 import _lib, types as T
from ssrlcv_b200.match_factory import Image, MatchFactory
from ssrlcv_b200.unity import Unity, cpu, gpu
from tests.golden_data import load_everest
ev = load_everest(); q, t, s = ev.views2[0], ev.views2[1], ev.seed2
qu, tu, su = Unity(q, T.FEATURE, cpu), Unity(t, T.FEATURE, cpu), Unity(s, T.FEATURE, cpu)
for u in (qu, tu, su): u.setMemoryState(gpu)
qi, ti = Image(ev.img_ids[0], ev.cameras[0]), Image(ev.img_ids[1], ev.cameras[1])
def timed(name, mf, fn):
    lib, h = mf.ctx.lib, mf.ctx.handle
    for _ in range(2): fn()
    lib.ssrlcv_b200_set_profiling(h, 1)
    for _ in range(5): r = fn()
    torch.cuda.synchronize()
    k_ms, k_n = C.c_double(0), C.c_uint64(0); lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n)); lib.ssrlcv_b200_set_profiling(h, 0)
    print(f"{name:70s} matcher {k_ms.value / max(k_n.value, 1) * 1e3:8.1f} us x{k_n.value // 5}  records {r.size()}", flush=True)
for thr in (ev.abs_threshold, 3.4e38):
    mf = MatchFactory(ev.rel_threshold, thr)
    mf.setSeedFeatures(su)
    for cm in (1, 0):
        for pair in (1, 0):
            mf.ctx.tune(_lib.TUNE_COLUMN_MASKS, cm); mf.ctx.tune(_lib.TUNE_PAIR_UNITS, pair)
            timed(f"thr {thr:g} colmask {cm} pair {pair}: 32425 x 36251 unseeded", mf, lambda: mf.generateDistanceMatches(qi, qu, ti, tu))
            timed(f"thr {thr:g} colmask {cm} pair {pair}: seed pass 32425 x 8932", mf, lambda: mf.getSeedDistances(qu))
mf.ctx.tune(_lib.TUNE_COLUMN_MASKS, -1); mf.ctx.tune(_lib.TUNE_PAIR_UNITS, -1)
