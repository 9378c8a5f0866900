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

# the same through the C ABI with host buffers (pageable numpy memory, then pinned): ssrlcv_b200_match_host + ssrlcv_b200_match_set_2view
from ssrlcv_b200.match_factory import MatchFactory  # noqa: E402

mf = MatchFactory(ev.rel_threshold, ev.abs_threshold)
lib, h = mf.ctx.lib, mf.ctx.handle
p = mf._params(T.OUT_DMATCH, T.MASK_DOUBLE, Image(ev.img_ids[0], ev.cameras[0]), Image(ev.img_ids[1], ev.cameras[1]), epsilon=ev.epsilon, delta=ev.delta)
out_d = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
kp_d = torch.empty(len(q) * 32, dtype=torch.uint8, device="cuda")
mm_d = torch.empty(len(q) * 8, dtype=torch.uint8, device="cuda")


def cabi(qh, th, sh, out_h):
    n, mx = C.c_uint64(0), C.c_float(0)
    t0 = time.perf_counter()
    mf.ctx.check(lib.ssrlcv_b200_match_host(h, C.byref(p), C.c_void_p(qh.data_ptr()), len(q), C.c_void_p(th.data_ptr()), len(t),
                                            C.c_void_p(sh.data_ptr()), len(s), C.c_void_p(out_h.data_ptr()), C.byref(n)))
    return (time.perf_counter() - t0) * 1e3, n.value


as_t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1))
for name, prep in (("pageable", lambda x: x), ("pinned  ", lambda x: x.pin_memory())):
    qh, th, sh, oh = prep(as_t(q)), prep(as_t(t)), prep(as_t(s)), prep(torch.empty(len(q) * 48, dtype=torch.uint8))
    zs = [cabi(qh, th, sh, oh) for _ in range(9)][2:]
    print(f"this repo  : ssrlcv_b200_match_host ({name} host buffers in, DMatch records out): {np.median([z[0] for z in zs]):8.3f} ms   ({zs[0][1]} matches)")

for so, label in (("libssrlcv_dropin.so", "drop-in    : integration/MatchFactory_b200.cu behind the reference's API (same harness)"),
                  ("libssrlcv_ref.so", "reference  : its own CUDA MatchFactory.cu on this GPU                            ")):
    path = os.path.join(ROOT, "oracle", "_ref", so)
    if not os.path.exists(path):
        continue
    os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
    os.chdir("/tmp/ssrlcv_ref_run")
    ref = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0))
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
        return (time.perf_counter() - t0) * 1e3, ms_seed.value + ms.value, n, ms_seed.value, ms.value

    ys = [theirs() for _ in range(5)][2:]
    print(f"{label}: getSeedDistances + generateDistanceMatchesDoubleConstrained "
          f"{np.median([y[1] for y in ys]):8.3f} ms in the API calls = {np.median([y[3] for y in ys]):.3f} + {np.median([y[4] for y in ys]):.3f} ({np.median([y[0] for y in ys]):.3f} ms wall incl. harness copies; {ys[0][2]} matches)")
    os.chdir(ROOT)
