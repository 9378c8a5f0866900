#!/usr/bin/env python
"""One ssrlcv_b200_match_host call sequence on the reference's 2-view fixture (for `ncu --metrics gpu__time_duration.sum`)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ssrlcv_b200 import types as T  # noqa: E402
from ssrlcv_b200.match_factory import Image, MatchFactory  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

ev = load_everest()
q, t, s = ev.views2[0], ev.views2[1], ev.seed2
mf = MatchFactory(ev.rel_threshold, ev.abs_threshold)
lib, h = mf.ctx.lib, mf.ctx.handle
p = mf._params(T.OUT_DMATCH, T.MASK_DOUBLE, Image(ev.img_ids[0], ev.cameras[0]), Image(ev.img_ids[1], ev.cameras[1]), epsilon=ev.epsilon, delta=ev.delta)
as_t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
qh, th, sh, oh = as_t(q), as_t(t), as_t(s), torch.empty(len(q) * 48, dtype=torch.uint8).pin_memory()
n = C.c_uint64(0)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    mf.ctx.check(lib.ssrlcv_b200_match_host(h, C.byref(p), C.c_void_p(qh.data_ptr()), len(q), C.c_void_p(th.data_ptr()), len(t),
                                            C.c_void_p(sh.data_ptr()), len(s), C.c_void_p(oh.data_ptr()), C.byref(n)))
torch.cuda.synchronize()
print("matches", n.value)
