#!/usr/bin/env python
"""Where a C4 step goes on one rank of an 8-rank run (15 of the 120 pairs of 16 x 32 768 views): per-phase wall time
with the device synchronised between phases, next to the un-synchronised step."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ssrlcv_b200 import _lib, types as T  # noqa: E402
from ssrlcv_b200.match_factory import Context  # noqa: E402
from ssrlcv_b200.pair_scheduler import assign_pairs, pair_order  # noqa: E402
from tests import synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
V, n = 16, 32768
ctx = Context.get(0)
lib, h = ctx.lib, ctx.handle
views = []
for k in range(V // 2):
    a, b = synth.make_pair(n, n, seed=0x55524C43 + k)
    views += [a, b]
pairs = pair_order(V)
mine = assign_pairs(pairs, [n] * V, world)[0]
need = sorted({v for k in mine for v in pairs[k]})
devbuf = {v: torch.from_numpy(views[v].view(np.uint8).reshape(-1)).cuda() for v in need}
out = torch.zeros((len(mine), n, 4), dtype=torch.int32, device="cuda")
p = _lib.MatchParams()
p.absolute_threshold, p.relative_threshold, p.compact, p.reserved = 4e4, 0.6, 1, 0
p.out_kind, p.mask_kind = T.OUT_PAIR, 0
print(f"world {world}: {len(mine)} pairs, {len(need)} views on this rank")


def step(sync):
    t = [time.perf_counter()]
    sets = {}
    for v in need:
        hs = C.c_void_p()
        ctx.check(lib.ssrlcv_b200_pack(h, C.c_void_p(devbuf[v].data_ptr()), n, C.byref(hs)))
        sets[v] = hs
    t.append(time.perf_counter())
    if sync:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    m = len(mine)
    P = (_lib.MatchParams * m)()
    Q, Tt, O = (C.c_void_p * m)(), (C.c_void_p * m)(), (C.c_void_p * m)()
    for s_i, k in enumerate(mine):
        qv, tv = pairs[k]
        p.query_image_id, p.target_image_id = qv, tv
        C.memmove(C.byref(P, s_i * C.sizeof(_lib.MatchParams)), C.byref(p), C.sizeof(_lib.MatchParams))
        Q[s_i], Tt[s_i], O[s_i] = sets[qv], sets[tv], out[s_i].data_ptr()
    n_outs = (C.c_uint64 * m)()
    t.append(time.perf_counter())
    ctx.check(lib.ssrlcv_b200_match_batch(h, m, P, Q, Tt, None, O, n_outs))
    t.append(time.perf_counter())
    for hs in sets.values():
        lib.ssrlcv_b200_set_free(hs)
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return [1e6 * (b - a) for a, b in zip(t, t[1:])], sum(int(x) for x in n_outs)


for sync in (True, False):
    acc = []
    for it in range(8):
        d, tot = step(sync)
        if it >= 3:
            acc.append(d)
    med = np.median(np.array(acc), axis=0)
    names = ["pack calls (host)", "pack drain (sync)", "python arg arrays", "match_batch (incl. its sync)", "free + sync"]
    print("synchronised between phases" if sync else "back to back")
    for nm, v in zip(names, med):
        print(f"  {nm:30s} {v:9.1f} us")
    print(f"  {'step total':30s} {med.sum():9.1f} us   ({tot} records)")
lib.ssrlcv_b200_set_profiling(h, 1)
step(False)
k_ms, k_n = C.c_double(0), C.c_uint64(0)
lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
print(f"smax kernel: {1e3 * k_ms.value:.1f} us in {k_n.value} launch(es)")
