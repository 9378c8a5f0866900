#!/usr/bin/env python
"""Bring-up of the CTA-pair walk (smax_duo.cu, ssrlcv_b200_tune(PAIR_UNITS, 2)): DMatch records of a few shapes under the pair walk
against the one-tile walk of the same library, every (fold, column-mask) variant.  Prints one line per case; exit code 1 on a mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import synth  # noqa: E402
from ssrlcv_b200 import _lib, types as T  # noqa: E402
from ssrlcv_b200.match_factory import Context, Image, MatchFactory  # noqa: E402
from ssrlcv_b200.unity import Unity, cpu  # noqa: E402

ctx = Context.get(0)
bad = 0
shapes = [(300, 700), (513, 4097), (1000, 256), (4096 + 77, 8192 + 5), (16384, 16384)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for nq, nt in shapes:
    q, t = synth.make_pair(nq, nt, seed=nq + nt, dup=0.02)
    for thr in (3.4e38, 200.0 * 200.0):
        for fold in (1, 0):
            for cm in (1, 0):
                res = []
                for walk in (0, 1, 2):
                    ctx.tune(_lib.TUNE_FOLD, fold); ctx.tune(_lib.TUNE_COLUMN_MASKS, cm); ctx.tune(_lib.TUNE_PAIR_UNITS, walk)
                    mf = MatchFactory(0.0, thr)
                    m = mf.generateDistanceMatches(Image(0), Unity(q, T.FEATURE, cpu), Image(1), Unity(t, T.FEATURE, cpu))
                    res.append(m.to_numpy())
                ok = len(res[0]) == len(res[2]) and T.fields_equal(res[0], res[2])
                ok1 = len(res[0]) == len(res[1]) and T.fields_equal(res[0], res[1])
                bad += not ok
                print(f"{nq} x {nt} thr {thr:g} fold {fold} colmask {cm}: {len(res[0])} records, duo identical {ok}, pair identical {ok1}", flush=True)
                if not ok and len(res[0]) == len(res[2]):
                    a, b = res[0], res[2]
                    qa, qb = a["keyPoints"]["loc"][:, 0], b["keyPoints"]["loc"][:, 0]
                    diff = np.nonzero((a["distance"] != b["distance"]) | (a["keyPoints"]["loc"][:, 1] != b["keyPoints"]["loc"][:, 1]).any(-1))[0]
                    print(f"   {len(diff)} differing records; first rows {diff[:12].tolist()} last {diff[-4:].tolist()}")
                    for i in diff[:4]:
                        print(f"   row {i}: walk0 d {a['distance'][i]} tloc {a['keyPoints']['loc'][i, 1]}  duo d {b['distance'][i]} tloc {b['keyPoints']['loc'][i, 1]}")
sys.exit(1 if bad else 0)
