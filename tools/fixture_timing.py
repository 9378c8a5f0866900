#!/usr/bin/env python
"""The reference's own 2-view fixture (everest, test/checkpoints/Pipeline2View: 32 425 x 36 251 features, 8 932 seed features,
DoubleConstrained eps = 25 px, delta = 5 km, MatchFactory(0.6, 200^2)) through the matching stage of doFeatureMatching
(src/Pipeline.cu:173-222), host-resident inputs to host-resident MatchSet: this repo's path next to the reference's own
CUDA MatchFactory (oracle/_ref, rebuilt for sm_100a) on the same box.  Wall clock, median of 7 after 2 warm-ups."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ssrlcv_b200 import types as T  # noqa: E402
from ssrlcv_b200.match_factory import Image  # noqa: E402
from ssrlcv_b200.pipeline import FeatureMatchingInput, doFeatureMatching  # noqa: E402
from ssrlcv_b200.unity import Unity, cpu  # noqa: E402
from tests.golden_data import load_everest  # noqa: E402

ev = load_everest()
q, t, s = ev.views2[0], ev.views2[1], ev.seed2
print(f"fixture: {len(q)} x {len(t)} features, {len(s)} seed features")


def ours():
    inp = FeatureMatchingInput(images=[Image(ev.img_ids[0], ev.cameras[0]), Image(ev.img_ids[1], ev.cameras[1])],
                               allFeatures=[Unity(q, T.FEATURE, cpu), Unity(t, T.FEATURE, cpu)], seedFeatures=Unity(s, T.FEATURE, cpu),
                               epsilon=ev.epsilon, delta=ev.delta)
    t0 = time.perf_counter()
    out = doFeatureMatching(inp)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, out.matchSet.matches.size()


xs = [ours() for _ in range(9)][2:]
print(f"this repo  : doFeatureMatching (seed distances + DoubleConstrained match + MatchSet), host in -> host out: "
      f"{np.median([x[0] for x in xs]):8.3f} ms   ({xs[0][1]} matches)")

path = os.path.join(ROOT, "oracle", "_ref", "libssrlcv_ref.so")
if os.path.exists(path):
    os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
    os.chdir("/tmp/ssrlcv_ref_run")
    ref = C.CDLL(path)
    ref.ref_match.restype = C.c_long
    sd = np.zeros(len(q), np.float32)
    out = np.zeros(len(q), T.DMATCH)

    def theirs():
        ms_seed, ms = C.c_double(0), C.c_double(0)
        t0 = time.perf_counter()
        ref.ref_seed_distances(q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), s.ctypes.data_as(C.c_void_p), C.c_ulong(len(s)), 2,
                               sd.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ms_seed))
        n = ref.ref_match(1, 2, C.c_float(ev.abs_threshold), C.c_float(ev.rel_threshold), ev.img_ids[0], ev.img_ids[1],
                          q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                          sd.ctypes.data_as(C.POINTER(C.c_float)), 2, C.c_float(ev.epsilon), C.c_float(ev.delta), None,
                          np.asarray(ev.cameras[0]).tobytes(), np.asarray(ev.cameras[1]).tobytes(), out.ctypes.data_as(C.c_void_p), C.byref(ms))
        return (time.perf_counter() - t0) * 1e3, ms_seed.value + ms.value, n

    ys = [theirs() for _ in range(5)][2:]
    print(f"reference  : getSeedDistances + generateDistanceMatchesDoubleConstrained, its own CUDA kernels on this GPU: "
          f"{np.median([y[1] for y in ys]):8.3f} ms in the API calls ({np.median([y[0] for y in ys]):.3f} ms wall incl. harness copies; {ys[0][2]} matches)")
    os.chdir(ROOT)
