#!/usr/bin/env python
"""GPU bring-up ladder: each rung runs in its own process under a timeout so a hang or a wrong
descriptor encoding is localised.  Usage: python tools/bringup.py [rung ...]   (no args = all)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RUNGS = ["simt", "tc_nofold", "tc_fold", "tc_big", "masked", "trig", "golden2", "ref"]


def data(nq, nt, seed=0, sift_like=True):
    import numpy as np
    from ssrlcv_b200 import types as T
    rng = np.random.default_rng(seed)
    if sift_like:
        def gen(n):
            x = np.maximum(rng.normal(size=(n, 128)), 0) ** 2 + rng.exponential(0.05, size=(n, 128))
            x /= np.linalg.norm(x, axis=1, keepdims=True); x = np.minimum(x, 0.2); x /= np.linalg.norm(x, axis=1, keepdims=True)
            return np.round(255 * x).astype(np.uint8)
    else:
        def gen(n):
            return rng.integers(0, 256, (n, 128), dtype=np.uint8)
    qv, tv = gen(nq), gen(nt)
    k = min(nq, nt) // 3
    idx = rng.choice(nt, k, replace=False)
    tv[idx] = np.clip(qv[:k].astype(np.int32) + rng.integers(-3, 4, (k, 128)), 0, 255).astype(np.uint8)
    # exact duplicates (tie stress): lane rule != lowest index
    if nt > 200:
        tv[37] = tv[66] = qv[0]
    q = T.make_features(qv, rng.uniform(6, 1018, (nq, 2)).astype(np.float32), parent=0)
    t = T.make_features(tv, rng.uniform(6, 1018, (nt, 2)).astype(np.float32), parent=1)
    return q, t


def compare(name, got, want):
    import numpy as np
    from ssrlcv_b200 import types as T
    ok = T.fields_equal(got, want)
    print(f"[{name}] n_got={len(got)} n_want={len(want)} equal={ok}", flush=True)
    if not ok and len(got) == len(want) and len(got):
        for f in got.dtype.names:
            if f.startswith("_pad"):
                continue
            bad = np.flatnonzero([not np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(got[f], want[f])])
            if len(bad):
                print(f"   field {f}: {len(bad)} mismatches, first at {bad[:5]}: got {got[f][bad[0]]} want {want[f][bad[0]]}")
    return ok


def _fold(v):
    from ssrlcv_b200 import _lib
    from ssrlcv_b200.match_factory import Context
    Context.get(0).tune(_lib.TUNE_FOLD, v)


def run_engine_case(engine, nq, nt, tag, sift_like=True, seed=0):
    import numpy as np
    from oracle import oracle as O
    from ssrlcv_b200 import types as T
    from ssrlcv_b200.match_factory import Image, MatchFactory
    from ssrlcv_b200.unity import Unity, cpu
    q, t = data(nq, nt, seed, sift_like)
    mf = MatchFactory(1.0, 3.4e38, engine=engine)
    qi, ti = Image(0), Image(1)
    out = mf.generateDistanceMatches(qi, Unity(q, T.FEATURE, cpu), ti, Unity(t, T.FEATURE, cpu)).to_numpy()
    want = O.generate_matches(O.make_params(out_kind=T.OUT_DMATCH), q, t)
    return compare(f"{tag} {nq}x{nt}", out, want)


def rung(name):
    import numpy as np
    ok = True
    if name == "simt":
        for nq, nt in ((1, 1), (7, 33), (128, 256), (300, 1000)):
            ok &= run_engine_case(1, nq, nt, "simt")
    elif name == "tc_nofold":
        _fold(0)
        for nq, nt in ((128, 256), (1, 1), (300, 1000), (1000, 5000)):
            ok &= run_engine_case(0, nq, nt, "tc_nofold")
        _fold(-1)
    elif name == "tc_fold":
        for nq, nt in ((128, 256), (1, 1), (300, 1000), (1000, 5000)):
            ok &= run_engine_case(0, nq, nt, "tc_fold")
        ok &= run_engine_case(0, 500, 3000, "tc_fold_uniform_bytes", sift_like=False)
    elif name == "tc_big":
        ok &= run_engine_case(0, 16384, 16384, "tc_fold", seed=3)
        _fold(0)
        ok &= run_engine_case(0, 8192, 20000, "tc_nofold", seed=4)
        _fold(-1)
    elif name == "masked":
        from oracle import oracle as O
        from ssrlcv_b200 import types as T
        from ssrlcv_b200.match_factory import Image, MatchFactory
        from ssrlcv_b200.unity import Unity, cpu
        from tests.golden_data import load_everest
        ev = load_everest()
        q, t = data(3000, 5000, 5)
        sd_want = O.seed_distances(q, ev.seed2[:2000])
        for engine in (1, 0):
            mf = MatchFactory(0.8, 9.0e4, engine=engine)
            mf.setSeedFeatures(Unity(ev.seed2[:2000], T.FEATURE, cpu))
            qi, ti = Image(0, ev.cameras[0]), Image(1, ev.cameras[1])
            qu, tu = Unity(q, T.FEATURE, cpu), Unity(t, T.FEATURE, cpu)
            sd = mf.getSeedDistances(qu)
            print(f"[masked e{engine}] seed distances equal:", np.array_equal(sd.to_numpy(), sd_want), flush=True)
            got = mf.generateDistanceMatchesDoubleConstrained(qi, qu, ti, tu, 25.0, 5.0, sd).to_numpy()
            p = O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=9.0e4, rel_threshold=0.8,
                              query_id=0, target_id=1, epsilon=25, delta=5, query_camera=ev.cameras[0], target_camera=ev.cameras[1])
            ok &= compare(f"masked double e{engine}", got, O.generate_matches(p, q, t, sd_want))
            F = np.array([[0, 0, 0.002], [0, 0, -1.0], [-0.002, 1.0, 3.0]], np.float32)
            got = mf.generateMatchesConstrainedIndexOnly(qi, qu, ti, tu, 40.0, F, sd).to_numpy()
            p = O.make_params(out_kind=T.OUT_PAIR, mask_kind=T.MASK_FUNDAMENTAL, has_seed=True, abs_threshold=9.0e4, rel_threshold=0.8,
                              query_id=0, target_id=1, epsilon=40, fundamental=F)
            ok &= compare(f"masked F e{engine}", got, O.generate_matches(p, q, t, sd_want))
    elif name == "trig":
        import ctypes as C
        import torch
        from oracle import oracle as O
        from ssrlcv_b200.match_factory import Context
        ctx = Context.get(0)
        rng = np.random.default_rng(1)
        x = np.concatenate([rng.uniform(-7, 7, 200000), rng.uniform(-1000, 1000, 50000), [0.0, 2.0567966, 0.022177858, -0.041954666]]).astype(np.float32)
        xd = torch.from_numpy(x).cuda(); s = torch.empty_like(xd); c = torch.empty_like(xd)
        ctx.check(ctx.lib.ssrlcv_b200_device_sincos(ctx.handle, C.c_void_p(xd.data_ptr()), len(x), C.c_void_p(s.data_ptr()), C.c_void_p(c.data_ptr())))
        sw = np.array([O.lib().orc_cuda_sinf(float(v)) for v in x], np.float32)
        cw = np.array([O.lib().orc_cuda_cosf(float(v)) for v in x], np.float32)
        sg, cg = s.cpu().numpy(), c.cpu().numpy()
        print("[trig] sin mismatches", int((sg.view(np.uint32) != sw.view(np.uint32)).sum()), "cos mismatches", int((cg.view(np.uint32) != cw.view(np.uint32)).sum()), "of", len(x), flush=True)
        ok &= bool((sg.view(np.uint32) == sw.view(np.uint32)).all() and (cg.view(np.uint32) == cw.view(np.uint32)).all())
    elif name == "golden2":
        from ssrlcv_b200 import types as T
        from ssrlcv_b200.match_factory import Image, MatchFactory
        from ssrlcv_b200.unity import Unity, cpu, gpu, both
        from tests.golden_data import load_everest
        import time
        ev = load_everest()
        for engine in (0, 1):
            mf = MatchFactory(ev.rel_threshold, ev.abs_threshold, engine=engine)
            mf.setSeedFeatures(Unity(ev.seed2, T.FEATURE, cpu))
            qu, tu = Unity(ev.views2[0], T.FEATURE, both), Unity(ev.views2[1], T.FEATURE, both)
            t0 = time.time()
            sd = mf.getSeedDistances(qu)
            m = mf.generateDistanceMatchesDoubleConstrained(Image(ev.img_ids[0], ev.cameras[0]), qu, Image(ev.img_ids[1], ev.cameras[1]), tu, ev.epsilon, ev.delta, sd)
            raw = mf.getRawMatches(m).to_numpy()
            print(f"[golden2 e{engine}] {time.time() - t0:.3f}s matches {m.size()} (want 13534)", flush=True)
            ok &= compare(f"golden2 e{engine}", raw["keyPoints"].reshape(-1).copy(), ev.g2_keypoints)
    elif name == "ref":
        import ctypes as C
        from ssrlcv_b200 import types as T
        from oracle import oracle as O
        path = os.path.join(ROOT, "oracle", "_ref", "libssrlcv_ref.so")
        if not os.path.exists(path):
            print("[ref] oracle/_ref not built; skipped"); return True
        os.makedirs("/tmp/refrun", exist_ok=True); os.chdir("/tmp/refrun")
        ref = C.CDLL(path)
        ref.ref_match.restype = C.c_long
        q, t = data(3000, 5000, 5)
        out = np.zeros(len(q), T.DMATCH); ms = C.c_double(0)
        n = ref.ref_match(1, 0, C.c_float(3.4e38), C.c_float(1.0), 0, 1, q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                          None, 2, C.c_float(0), C.c_float(0), None, None, None, out.ctypes.data_as(C.c_void_p), C.byref(ms))
        want = O.generate_matches(O.make_params(out_kind=T.OUT_DMATCH), q, t)
        print(f"[ref] reference CUDA returned {n} matches in {ms.value:.2f} ms", flush=True)
        ok &= compare("reference-vs-oracle DMatch", out[:n], want)
    print(f"RUNG {name}: {'PASS' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--rung":
        sys.exit(0 if rung(sys.argv[2]) else 1)
    todo = sys.argv[1:] or RUNGS
    summary = {}
    for r in todo:
        try:
            p = subprocess.run(["timeout", "-k", "5", "240", sys.executable, os.path.abspath(__file__), "--rung", r], cwd=ROOT)
            summary[r] = p.returncode
        except Exception as e:  # noqa: BLE001
            summary[r] = repr(e)
    print("SUMMARY", summary, flush=True)
