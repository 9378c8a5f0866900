"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs,
against the reference's committed golden fixtures, and -- when oracle/_ref was built -- against the
reference's own CUDA kernels running on the same device.  Integer work is bit-exact; float fields are
compared by bit pattern."""
import ctypes as C
import os

import numpy as np
import pytest

from ssrlcv_b200 import types as T
from tests import synth

pytestmark = pytest.mark.gpu

FLT_MAX = float(np.finfo(np.float32).max)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from oracle import oracle as O
    from ssrlcv_b200.match_factory import Image, MatchFactory
    from ssrlcv_b200.unity import Unity, both, cpu, gpu

    class Api:
        pass

    a = Api()
    a.O, a.Image, a.MatchFactory, a.Unity, a.cpu, a.gpu, a.both = O, Image, MatchFactory, Unity, cpu, gpu, both
    _suite_default_knobs()
    return a


def _suite_default_knobs():
    """SSRLCV_TEST_FOLD=0|1 / SSRLCV_TEST_COLMASK=0|1 / SSRLCV_TEST_PAIR=0|1 run the whole suite under one forced matcher variant (tools/ab_kernel.sh)."""
    from ssrlcv_b200 import _lib
    from ssrlcv_b200.match_factory import Context
    f = os.environ.get("SSRLCV_TEST_FOLD")
    Context.get(0).tune(_lib.TUNE_FOLD, int(f) if f not in (None, "") else -1)
    cm = os.environ.get("SSRLCV_TEST_COLMASK")
    Context.get(0).tune(_lib.TUNE_COLUMN_MASKS, int(cm) if cm not in (None, "") else -1)
    pu = os.environ.get("SSRLCV_TEST_PAIR")
    Context.get(0).tune(_lib.TUNE_PAIR_UNITS, int(pu) if pu not in (None, "") else -1)


@pytest.fixture
def knobs(api):
    """ssrlcv_b200_tune for the duration of one test: knobs(fold=0, chunk_rows=512, ...); everything back to "library decides" afterwards."""
    from ssrlcv_b200 import _lib
    from ssrlcv_b200.match_factory import Context
    names = {"fold": [_lib.TUNE_FOLD], "colmask": [_lib.TUNE_COLUMN_MASKS], "sorted": [_lib.TUNE_SORTED_TARGETS], "pair": [_lib.TUNE_PAIR_UNITS],
             "cells": [_lib.TUNE_CELL_SORT], "seedonly": [_lib.TUNE_SEED_ONLY], "lean": [_lib.TUNE_LEAN_MASKED], "mask_coop": [_lib.TUNE_MASK_COOPERATIVE], "mask_walk": [_lib.TUNE_MASK_WALK],
             "chunk_rows": [_lib.TUNE_HOST_CHUNK_QUERY_ROWS, _lib.TUNE_HOST_CHUNK_TARGET_ROWS]}
    ctx = Context.get(0)

    def set_(**kw):
        for k, v in kw.items():
            for knob in names[k]:
                ctx.tune(knob, -1 if v is None else int(v))

    yield set_
    for ks in names.values():
        for knob in ks:
            ctx.tune(knob, -1)
    _suite_default_knobs()


def _unity(api, arr, state=None):
    return api.Unity(arr, T.FEATURE, api.cpu if state is None else state)


# ------------------------------------------------------------------------------------ unconstrained
@pytest.mark.parametrize("nq,nt", [(1, 1), (5, 31), (128, 256), (129, 257), (1000, 4097), (4096, 4096)])
@pytest.mark.parametrize("engine", [0, 1])
def test_bruteforce_dmatch_matches_oracle(api, nq, nt, engine):
    q, t = synth.make_pair(nq, nt, seed=nq + nt)
    mf = api.MatchFactory(1.0, FLT_MAX, engine=engine)
    got = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH), q, t)
    assert T.fields_equal(got, want)


@pytest.mark.parametrize("out_kind", [T.OUT_MATCH, T.OUT_DMATCH, T.OUT_PAIR])
@pytest.mark.parametrize("seeded", [False, True])
def test_output_kinds_and_ratio_rule(api, out_kind, seeded):
    """rel vs rel^2 depends on the output type."""
    q, t = synth.make_pair(3000, 3500, seed=11)
    s = synth.make_seed(800)
    mf = api.MatchFactory(0.75, 1.2e5)
    mf.setSeedFeatures(_unity(api, s))
    qu, tu = _unity(api, q), _unity(api, t)
    sd = mf.getSeedDistances(qu) if seeded else None
    sd_want = api.O.seed_distances(q, s) if seeded else None
    if seeded:
        assert np.array_equal(sd.to_numpy(), sd_want)
    fn = {T.OUT_MATCH: mf.generateMatches, T.OUT_DMATCH: mf.generateDistanceMatches, T.OUT_PAIR: mf.generateMatchesIndexOnly}[out_kind]
    got = fn(api.Image(4), qu, api.Image(9), tu, sd).to_numpy()
    p = api.O.make_params(out_kind=out_kind, has_seed=seeded, abs_threshold=1.2e5, rel_threshold=0.75, query_id=4, target_id=9)
    want = api.O.generate_matches(p, q, t, sd_want)
    assert len(want) > 100 and T.fields_equal(got, want)


def test_feature_match_keeps_all_entries(api):
    """generateFeatureMatches neither compacts nor drops invalid entries (ref: src/MatchFactory.cu:674-696)."""
    q, t = synth.make_pair(700, 900, seed=5)
    mf = api.MatchFactory(1.0, 3.0e4)
    got = mf.generateFeatureMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    assert len(got) == 700
    bd, bi, ok = api.O.match_winners(api.O.make_params(out_kind=T.OUT_FEATUREMATCH, abs_threshold=3.0e4), q, t)
    assert np.array_equal(got["invalid"] == 0, ok.astype(bool)) and 0 < ok.sum() < 700
    v = np.flatnonzero(ok)
    assert np.array_equal(got["distance"][v], bd[v])
    assert np.array_equal(got["descriptors"]["values"][v, 0], q["values"][v])
    assert np.array_equal(got["descriptors"]["values"][v, 1], t["values"][bi[v]])
    assert np.array_equal(got["keyPoints"]["loc"][v, 1], t["loc"][bi[v]])
    # and the constrained form does compact
    F = np.array([[0, 0, 0.001], [0, 0, -1.0], [-0.001, 1.0, 2.0]], np.float32)
    got = mf.generateFeatureMatchesConstrained(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t), 60.0, F).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_FEATUREMATCH, mask_kind=T.MASK_FUNDAMENTAL, abs_threshold=3.0e4,
                                                    epsilon=60.0, fundamental=F), q, t)
    assert T.fields_equal(got, want)


# ------------------------------------------------------------------------------------ ties / edge cases
@pytest.mark.parametrize("engine", [0, 1])
def test_tie_break_lane_then_index(api, engine):
    """ref: src/MatchFactory.cu:1476-1492 -- winner among equal distances is min (f mod 32, f)."""
    rng = np.random.default_rng(2)
    qv = synth.sift_descriptors(rng, 64)
    tv = synth.sift_descriptors(rng, 2048)
    for k in range(64):                       # duplicates in the same window, in different windows, in different tiles
        near = qv[k].copy(); near[k % 128] ^= 1
        for f in (37 + 64 * (k % 3), 66 + 64 * (k % 3), 300 + k, 1500 + (k * 7) % 500):
            tv[f] = near
    q, t = T.make_features(qv), T.make_features(tv)
    mf = api.MatchFactory(1.0, FLT_MAX, engine=engine)
    got = mf.generateMatchesIndexOnly(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_PAIR), q, t)
    assert T.fields_equal(got, want)
    assert (want["b"][:, 1] != np.array([min(f for f in range(2048) if (tv[f] == tv[w]).all()) for w in want["b"][:, 1]])).any()


@pytest.mark.parametrize("thr", [FLT_MAX, 30000.0])
def test_tie_break_both_bookkeeping_regimes(api, thr):
    """The tensor-core kernel records the exact columns that attained a window maximum for thresholded searches and short
    target sets, and only the window (re-evaluated afterwards) for long unthresholded ones (smax_tc.cu, kColMask).  20 480
    targets put FLT_MAX into the second regime; ties within a window, across windows, across tiles and across the CTAs'
    units, two and three ways (three tied windows go to the exact kernel)."""
    rng = np.random.default_rng(5)
    qv = synth.sift_descriptors(rng, 160)
    tv = synth.sift_descriptors(rng, 20480)
    for k in range(160):
        near = qv[k].copy(); near[(3 * k) % 128] ^= 1
        spots = [(40 + 97 * k) % 20480, (41 + 97 * k + 32 * (k % 5)) % 20480, (9000 + 131 * k) % 20480]
        if k % 4 == 0:
            spots.append((17000 + 61 * k) % 20480)
        for f in spots[: 2 + k % 3]:
            tv[f] = near
    q, t = T.make_features(qv), T.make_features(tv)
    mf = api.MatchFactory(1.0, thr)
    got = mf.generateMatchesIndexOnly(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_PAIR, abs_threshold=thr), q, t)
    assert len(want) >= 100 and T.fields_equal(got, want)
    simt = api.MatchFactory(1.0, thr, engine=1).generateMatchesIndexOnly(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    assert T.fields_equal(simt, want)


@pytest.mark.parametrize("kind", ["bright", "mixed", "uniform"])
@pytest.mark.parametrize("thr", [FLT_MAX, 3.0e5])
def test_large_norm_sets_take_the_epilogue_add_variant(api, kind, thr):
    """Targets whose |b|^2 exceeds what the fold digits can carry (> 4 032 057, pack.cu / matcher.cuh kFoldC) are matched by
    the kFold = false variant (W added in the epilogue); uniform random bytes (|b|^2 ~ 2.8 M) still fold.  Both are exact."""
    rng = np.random.default_rng(21)
    nq, nt = 700, 5000
    if kind == "bright":
        qv = rng.integers(180, 256, (nq, 128), dtype=np.uint8); tv = rng.integers(180, 256, (nt, 128), dtype=np.uint8)
    elif kind == "mixed":        # SIFT-like rows next to saturated ones, all-zero and all-255 rows included
        qv = synth.sift_descriptors(rng, nq); tv = synth.sift_descriptors(rng, nt)
        tv[::3] = rng.integers(200, 256, (len(tv[::3]), 128), dtype=np.uint8)
        tv[7] = 255; tv[8] = 0; qv[5] = 255; qv[6] = 0
    else:
        qv = rng.integers(0, 256, (nq, 128), dtype=np.uint8); tv = rng.integers(0, 256, (nt, 128), dtype=np.uint8)
    tv[rng.choice(nt, 300, replace=False)] = np.clip(qv[:300].astype(int) + rng.integers(-3, 4, (300, 128)), 0, 255)
    q, t = T.make_features(qv), T.make_features(tv)
    mf = api.MatchFactory(1.0, thr)
    got = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=thr), q, t)
    assert len(want) >= 300 and T.fields_equal(got, want)


@pytest.mark.parametrize("engine", [0, 1])
def test_degenerate_inputs(api, engine):
    """all-equal descriptors (every score ties), all-zero, all-255 (norm range beyond the fold), empty target set."""
    mf = api.MatchFactory(1.0, FLT_MAX, engine=engine)
    for qv, tv in [
        (np.full((40, 128), 7, np.uint8), np.full((700, 128), 7, np.uint8)),
        (np.zeros((40, 128), np.uint8), np.zeros((300, 128), np.uint8)),
        (np.concatenate([np.zeros((20, 128), np.uint8), np.full((20, 128), 255, np.uint8)]),
         np.concatenate([np.full((300, 128), 255, np.uint8), np.zeros((300, 128), np.uint8), np.full((100, 128), 128, np.uint8)])),
    ]:
        q, t = T.make_features(qv), T.make_features(tv)
        got = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
        want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH), q, t)
        assert T.fields_equal(got, want)
    q = T.make_features(np.ones((10, 128), np.uint8))
    empty = np.zeros(0, T.FEATURE)
    got = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, empty))
    # reference quirk: zero valid matches -> caller keeps the uncompacted, all-invalid array
    assert got.size() == 10 and np.all(got.to_numpy()["invalid"] == 1)
    mf.setSeedFeatures(_unity(api, empty))
    assert np.all(mf.getSeedDistances(_unity(api, q)).to_numpy() == np.float32(FLT_MAX))


def test_absolute_threshold_is_strict(api):
    q, t = synth.make_pair(500, 800, seed=21)
    bd, _, _ = api.O.match_winners(api.O.make_params(), q, t)
    thr = float(np.sort(bd)[250])             # some winner sits exactly at the threshold
    mf = api.MatchFactory(1.0, thr)
    got = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=thr), q, t)
    assert T.fields_equal(got, want) and got["distance"].max() < thr


@pytest.mark.parametrize("thr_kind", ["fraction", "below_all", "median", "median_plus_half", "tiny", "zero", "huge"])
def test_threshold_floor_never_prunes_a_valid_winner(api, thr_kind):
    """The tensor-core kernel starts every query at the lowest score that can still pass `d2 < abs` (smax_tc.cu); the
    winners must stay those of the reference for thresholds on, between and outside the distances that occur
    (ref: src/MatchFactory.cu:1472, :1494)."""
    q, t = synth.make_pair(700, 2600, seed=77)
    bd, _, _ = api.O.match_winners(api.O.make_params(), q, t)
    s = np.sort(bd)
    thr = {"fraction": float(s[300]) + 0.25, "below_all": float(s[0]), "median": float(s[350]), "median_plus_half": float(s[350]) + 0.5,
           "tiny": 1.0e-3, "zero": 0.0, "huge": 8323200.0}[thr_kind]
    mf = api.MatchFactory(1.0, thr)
    for kind, fn in ((T.OUT_DMATCH, mf.generateDistanceMatches), (T.OUT_PAIR, mf.generateMatchesIndexOnly)):
        out = fn(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t))
        want = api.O.generate_matches(api.O.make_params(out_kind=kind, abs_threshold=thr), q, t)
        if len(want):                           # (all invalid: the reference keeps the uncompacted array, checked elsewhere)
            assert T.fields_equal(out.to_numpy(), want)
    if thr_kind in ("median", "fraction"):
        assert 0 < len(want) < len(q)


@pytest.mark.parametrize("nq,nt", [(300, 20000), (130, 70000), (2000, 300), (640, 1024 + 7)])
def test_query_tiles_shared_by_many_ctas(api, nq, nt):
    """Shapes where the per-CTA tile ranges cut query tiles into many units (several key slots per query), with
    duplicates and planted matches spread over the whole target range."""
    q, t = synth.make_pair(nq, nt, seed=nq * 7 + nt, dup=0.01)
    for thr in (FLT_MAX, 45000.0):
        mf = api.MatchFactory(1.0, thr)
        got = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
        want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=thr), q, t)
        assert len(want) and T.fields_equal(got, want)


def test_memory_states_restored(api):
    q, t = synth.make_pair(300, 400, seed=1)
    qu, tu = _unity(api, q, api.both), _unity(api, t, api.cpu)
    mf = api.MatchFactory()
    m = mf.generateMatches(api.Image(0), qu, api.Image(1), tu)
    assert qu.getMemoryState() == api.both and tu.getMemoryState() == api.cpu and m.getMemoryState() == api.gpu
    from ssrlcv_b200.match_factory import MatchFactoryError
    with pytest.raises(MatchFactoryError):
        mf.generateMatches(api.Image(0), qu, api.Image(1), tu, api.Unity(np.zeros(5, np.float32), np.dtype("<f4"), api.cpu))


@pytest.mark.parametrize("fold", ["0", "1"])
@pytest.mark.parametrize("thr", [FLT_MAX, 9.0e4])
def test_forced_matcher_variants(api, everest, knobs, fold, thr):
    """SSRLCV_B200_TUNE_FOLD forces the tensor-core kernel variant: five MMAs per tile with W folded into the contraction ("1"), or
    four MMAs with the window-bound test and W added in the epilogue ("0", otherwise only taken by sets whose norms do not
    fit the fold digits).  Both are exact, thresholded (column regime) and not (window regime), masked and unmasked."""
    knobs(fold=fold)
    q, t = synth.make_pair(3000, 9000, seed=77)
    s = synth.make_seed(700)
    mf = api.MatchFactory(0.8, thr)
    mf.setSeedFeatures(_unity(api, s))
    qu, tu = _unity(api, q), _unity(api, t)
    sd = mf.getSeedDistances(qu)
    sd_want = api.O.seed_distances(q, s)
    assert np.array_equal(sd.to_numpy(), sd_want)
    got = mf.generateDistanceMatches(api.Image(0), qu, api.Image(1), tu, sd).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, has_seed=True, abs_threshold=thr, rel_threshold=0.8), q, t, sd_want)
    assert len(want) > 100 and T.fields_equal(got, want)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    got = mf.generateDistanceMatchesDoubleConstrained(qi, qu, ti, tu, 25.0, 5.0).to_numpy()
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, abs_threshold=thr, query_id=0, target_id=1, epsilon=25, delta=5,
                          query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    assert T.fields_equal(got, api.O.generate_matches(p, q, t))


# ------------------------------------------------------------------------------------ host-buffer entry point
def _match_host(api, mf, p, q, t, s, elt_dtype):
    """ssrlcv_b200_match_host on plain (pageable) numpy buffers -> record array"""
    lib, h = mf.ctx.lib, mf.ctx.handle
    out = np.zeros(max(len(q), 1), elt_dtype)
    n_out = C.c_uint64(0)
    ptr = lambda a: C.c_void_p(a.ctypes.data) if a is not None and len(a) else (C.c_void_p(a.ctypes.data) if a is not None else None)
    mf.ctx.check(lib.ssrlcv_b200_match_host(h, C.byref(p), ptr(q), len(q), ptr(t), len(t), ptr(s), len(s) if s is not None else 0,
                                            C.c_void_p(out.ctypes.data), C.byref(n_out)))
    return out[: n_out.value]


@pytest.mark.parametrize("chunk_rows", ["0", "256", "512", "1280", None])
@pytest.mark.parametrize("nq,nt", [(3000, 3500), (700, 2049), (2049, 300), (1024, 1024)])
def test_match_host_streamed_upload_equals_oracle(api, knobs, chunk_rows, nq, nt):
    """match_host uploads large sets chunk by chunk and matches every chunk against what has already arrived (api.cu
    match_host_streamed); the chunk length is forced down here so that small cases take that path: every cut of the query
    and target sets gives the oracle's records.  "0" = never cut (the one-piece path), None = the library's own rule."""
    knobs(chunk_rows=chunk_rows)
    q, t = synth.make_pair(nq, nt, seed=nq * 7 + nt)
    s = synth.make_seed(600)
    mf = api.MatchFactory(0.8, 1.1e5)
    sd_want = api.O.seed_distances(q, s)
    for out_kind, dt in ((T.OUT_DMATCH, T.DMATCH), (T.OUT_PAIR, T.UINT2_PAIR), (T.OUT_MATCH, T.MATCH)):
        for seeded in (False, True):
            p = mf._params(out_kind, T.MASK_NONE, api.Image(3), api.Image(8))
            got = _match_host(api, mf, p, q, t, s if seeded else None, dt)
            po = api.O.make_params(out_kind=out_kind, has_seed=seeded, abs_threshold=1.1e5, rel_threshold=0.8, query_id=3, target_id=8)
            want = api.O.generate_matches(po, q, t, sd_want if seeded else None)
            assert len(want) > 50 and T.fields_equal(got, want), (out_kind, seeded)


def test_match_host_streamed_masked_and_unthresholded(api, everest, knobs):
    knobs(chunk_rows=768)
    q, t = synth.make_pair(4000, 6000, seed=31)
    s = synth.make_seed(1000)
    sd_want = api.O.seed_distances(q, s)
    mf = api.MatchFactory(0.8, 1.0e5)
    p = mf._params(T.OUT_DMATCH, T.MASK_DOUBLE, api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1]), epsilon=25.0, delta=5.0)
    got = _match_host(api, mf, p, q, t, s, T.DMATCH)
    po = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=1.0e5, rel_threshold=0.8,
                           query_id=0, target_id=1, epsilon=25, delta=5, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    want = api.O.generate_matches(po, q, t, sd_want)
    assert len(want) > 20 and T.fields_equal(got, want)
    # no absolute threshold: the window regime of the matcher (long target sets) and the column regime (short chunks)
    mf2 = api.MatchFactory(1.0, FLT_MAX)
    p2 = mf2._params(T.OUT_DMATCH, T.MASK_NONE, api.Image(0), api.Image(1))
    got2 = _match_host(api, mf2, p2, q, t, None, T.DMATCH)
    assert T.fields_equal(got2, api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH), q, t))
    # FeatureMatch records need the target descriptors of the whole set at finalize time
    p3 = mf._params(T.OUT_FEATUREMATCH, T.MASK_NONE, api.Image(0), api.Image(1), compact=0)
    got3 = _match_host(api, mf, p3, q, t, None, T.FEATUREMATCH)
    bd, bi, ok = api.O.match_winners(api.O.make_params(out_kind=T.OUT_FEATUREMATCH, abs_threshold=1.0e5), q, t)
    v = np.flatnonzero(ok)
    assert len(got3) == len(q) and np.array_equal(got3["invalid"] == 0, ok.astype(bool)) and len(v) > 100
    assert np.array_equal(got3["distance"][v], bd[v]) and np.array_equal(got3["descriptors"]["values"][v, 1], t["values"][bi[v]])
    assert np.array_equal(got3["keyPoints"]["loc"][v, 1], t["loc"][bi[v]])


def test_match_host_streamed_large_norms_fall_back(api, knobs):
    """a target chunk whose norms do not fit the fold digits: the optimistic chunk launches are discarded and the pair is
    redone in one piece by the epilogue-add variant."""
    knobs(chunk_rows=1024)
    rng = np.random.default_rng(5)
    nq, nt = 1500, 5000
    qv = synth.sift_descriptors(rng, nq); tv = synth.sift_descriptors(rng, nt)
    tv[3000:3400] = rng.integers(200, 256, (400, 128), dtype=np.uint8)          # only the third chunk is "bright"
    tv[rng.choice(nt, 300, replace=False)] = np.clip(qv[:300].astype(int) + rng.integers(-3, 4, (300, 128)), 0, 255)
    q, t = T.make_features(qv), T.make_features(tv)
    mf = api.MatchFactory(1.0, 3.0e5)
    got = _match_host(api, mf, mf._params(T.OUT_DMATCH, T.MASK_NONE, api.Image(0), api.Image(1)), q, t, None, T.DMATCH)
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=3.0e5), q, t)
    assert len(want) >= 250 and T.fields_equal(got, want)


# ------------------------------------------------------------------------------------ masks
@pytest.mark.parametrize("engine", [0, 1])
def test_double_constrained_matches_oracle(api, everest, engine):
    q, t = synth.make_pair(4000, 6000, seed=31)
    s = synth.make_seed(1000)
    mf = api.MatchFactory(0.8, 1.0e5, engine=engine)
    mf.setSeedFeatures(_unity(api, s))
    qu, tu = _unity(api, q), _unity(api, t)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    sd = mf.getSeedDistances(qu)
    sd_want = api.O.seed_distances(q, s)
    for out_kind, fn in ((T.OUT_DMATCH, mf.generateDistanceMatchesDoubleConstrained), (T.OUT_MATCH, mf.generateMatchesDoubleConstrained),
                         (T.OUT_PAIR, mf.generateMatchesDoubleConstrainedIndexOnly)):
        got = fn(qi, qu, ti, tu, 25.0, 5.0, sd).to_numpy()
        p = api.O.make_params(out_kind=out_kind, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=1.0e5, rel_threshold=0.8,
                              query_id=0, target_id=1, epsilon=25, delta=5, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
        want = api.O.generate_matches(p, q, t, sd_want)
        assert len(want) > 20 and T.fields_equal(got, want)
    got = mf.generateDistanceMatchesDoubleConstrained(qi, qu, ti, tu, 25.0, 5.0).to_numpy()      # non-seed variant
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, abs_threshold=1.0e5, query_id=0, target_id=1, epsilon=25, delta=5,
                          query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    assert T.fields_equal(got, api.O.generate_matches(p, q, t))


def test_epipolar_endpoints_bitwise(api, everest):
    """the DoubleConstrained prologue (device float geometry) equals the oracle's restatement bit for bit."""
    from ssrlcv_b200.match_factory import Context, pack
    rng = np.random.default_rng(0)
    q = T.make_features(np.zeros((5000, 128), np.uint8), rng.uniform(0, 1024, (5000, 2)).astype(np.float32))
    mf = api.MatchFactory()
    qu = _unity(api, q, api.gpu) if False else _unity(api, q)
    qu.setMemoryState(api.gpu)
    import torch
    out = torch.empty(5000 * 4, dtype=torch.float32, device="cuda")
    p = mf._params(T.OUT_DMATCH, T.MASK_DOUBLE, api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1]), 25.0, 5.0)
    ctx = mf.ctx
    ctx.check(ctx.lib.ssrlcv_b200_epipolar_endpoints(ctx.handle, C.byref(p), pack(ctx, qu).handle, C.c_void_p(out.data_ptr())))
    got = out.cpu().numpy().reshape(-1, 4)
    P = api.O.projection_matrix(everest.cameras[1])
    for i in range(0, 5000, 7):
        p1, p2 = api.O.epipolar_endpoints(everest.cameras[0], P, q["loc"][i], 5.0)
        assert np.array_equal(got[i].view(np.uint32), np.concatenate([p1, p2]).view(np.uint32)), i


def test_device_trig_equals_oracle_port(api):
    import torch
    mf = api.MatchFactory()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-7, 7, 100000), rng.uniform(-2000, 2000, 20000)]).astype(np.float32)
    xd = torch.from_numpy(x).cuda(); s = torch.empty_like(xd); c = torch.empty_like(xd)
    mf.ctx.check(mf.ctx.lib.ssrlcv_b200_device_sincos(mf.ctx.handle, C.c_void_p(xd.data_ptr()), len(x), C.c_void_p(s.data_ptr()), C.c_void_p(c.data_ptr())))
    lib = api.O.lib()
    sw = np.array([lib.orc_cuda_sinf(float(v)) for v in x], np.float32)
    cw = np.array([lib.orc_cuda_cosf(float(v)) for v in x], np.float32)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), sw.view(np.uint32))
    assert np.array_equal(c.cpu().numpy().view(np.uint32), cw.view(np.uint32))


def test_fundamental_constrained_matches_oracle(api):
    q, t = synth.make_pair(2500, 3000, seed=41)
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    mf = api.MatchFactory(1.0, 2.0e5)
    for out_kind, fn in ((T.OUT_MATCH, mf.generateMatchesConstrained), (T.OUT_DMATCH, mf.generateDistanceMatchesConstrained),
                         (T.OUT_PAIR, mf.generateMatchesConstrainedIndexOnly)):
        got = fn(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t), 30.0, F).to_numpy()
        want = api.O.generate_matches(api.O.make_params(out_kind=out_kind, mask_kind=T.MASK_FUNDAMENTAL, abs_threshold=2.0e5,
                                                        epsilon=30.0, fundamental=F), q, t)
        assert len(want) > 20 and T.fields_equal(got, want)


# ------------------------------------------------------------------------------------ golden fixtures of the reference
def test_two_view_golden(api, everest):
    """ref: test/Pipeline.cu:183-216 through doFeatureMatching (src/Pipeline.cu:173-241)."""
    ev = everest
    mf = api.MatchFactory(ev.rel_threshold, ev.abs_threshold)
    mf.setSeedFeatures(_unity(api, ev.seed2))
    qu, tu = _unity(api, ev.views2[0], api.both), _unity(api, ev.views2[1], api.both)
    sd = mf.getSeedDistances(qu)
    m = mf.generateDistanceMatchesDoubleConstrained(api.Image(ev.img_ids[0], ev.cameras[0]), qu, api.Image(ev.img_ids[1], ev.cameras[1]), tu,
                                                    ev.epsilon, ev.delta, sd)
    assert m.size() == 13534
    kp = mf.getRawMatches(m).to_numpy()["keyPoints"].reshape(-1).copy()
    assert T.fields_equal(kp, ev.g2_keypoints)


def test_three_view_golden(api, everest):
    """ref: test/Pipeline.cu:218-255 -- generateMatchesExhaustive with GEO_ORBIT=1."""
    ev = everest
    mf = api.MatchFactory(ev.rel_threshold, ev.abs_threshold)
    mf.setSeedFeatures(_unity(api, ev.seed3))
    images = [api.Image(ev.img_ids[i], ev.cameras[i]) for i in range(3)]
    feats = [_unity(api, v, api.both) for v in ev.views3]
    ms = mf.generateMatchesExhaustive(images, feats, ev.epsilon, ev.delta)
    assert T.fields_equal(ms.matches.to_numpy(), ev.g3_multimatch)
    assert T.fields_equal(ms.keyPoints.to_numpy(), ev.g3_keypoints)


# ------------------------------------------------------------------------------------ post-processing
def test_refine_sort_raw_validate(api):
    q, t = synth.make_pair(5000, 5000, seed=51)
    mf = api.MatchFactory(1.0, FLT_MAX)
    m = mf.generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t))
    base = m.to_numpy().copy()
    mf.sortMatches(m)
    assert T.fields_equal(m.to_numpy(), api.O.sort(base))
    thr = float(np.median(base["distance"]))
    mf.refineMatches(m, thr)
    assert T.fields_equal(m.to_numpy(), api.O.refine(api.O.sort(base), thr))
    assert T.fields_equal(mf.getRawMatches(m).to_numpy(), api.O.raw_matches(m.to_numpy()))
    from ssrlcv_b200.match_factory import MatchFactoryError
    with pytest.raises(MatchFactoryError):
        mf.refineMatches(m, 0.0)
    dirty = base.copy(); dirty["invalid"][::3] = 1
    u = api.Unity(dirty, T.DMATCH, api.cpu)
    mf.validateMatches(u)
    assert T.fields_equal(u.to_numpy(), dirty[dirty["invalid"] == 0])


# ------------------------------------------------------------------------------------ pair scheduler: one batch, one synchronisation
def test_match_batch_equals_individual_calls(api, everest):
    """ssrlcv_b200_match_batch (the pair loop of generateMatchesExhaustive, ref: src/MatchFactory.cu:907-936) must return, pair by
    pair, exactly what the single calls return: mixed sizes, masks, output kinds, seed distances, and one pair with no survivor."""
    sets = [synth.make_pair(n, 16, seed=300 + i)[0] for i, n in enumerate((900, 2100, 513, 1280))]
    us = [_unity(api, f) for f in sets]
    imgs = [api.Image(i, everest.cameras[i % 2]) for i in range(4)]
    mf = api.MatchFactory(0.8, 70000.0)
    mf.setSeedFeatures(_unity(api, synth.make_seed(300)))
    sd0 = mf.getSeedDistances(us[0])
    jobs = [
        (mf._params(T.OUT_PAIR, T.MASK_NONE, imgs[0], imgs[1]), us[0], us[1], None),
        (mf._params(T.OUT_DMATCH, T.MASK_NONE, imgs[0], imgs[2]), us[0], us[2], sd0),
        (mf._params(T.OUT_PAIR, T.MASK_DOUBLE, imgs[0], imgs[1], 2000.0, 5.0), us[0], us[1], sd0),
        (mf._params(T.OUT_MATCH, T.MASK_NONE, imgs[1], imgs[3]), us[1], us[3], None),
        (mf._params(T.OUT_PAIR, T.MASK_NONE, imgs[2], imgs[3]), us[2], us[3], None),
    ]
    strict = api.MatchFactory(0.8, 1.0)                      # nothing survives d2 < 1: the all-invalid array is kept
    jobs.append((strict._params(T.OUT_PAIR, T.MASK_NONE, imgs[2], imgs[3]), us[2], us[3], None))
    got = mf.matchBatch(jobs)
    assert len(got) == len(jobs)
    for (p, qf, tf, sd), g in zip(jobs, got):
        want = mf._run(p, qf, tf, sd)
        assert g.size() == want.size() and T.fields_equal(g.to_numpy(), want.to_numpy())
    assert got[-1].size() == us[2].size() and got[0].size() > 0


def test_match_batch_one_launch_many_shapes_matches_oracle(api):
    """One persistent launch over pairs of very different shapes (CTA ranges cross job boundaries, query tiles shared by
    several CTAs, a one-tile job, a job with an empty target set in the middle) against the CPU oracle, pair by pair."""
    shapes = [(3000, 20000), (130, 70000), (1, 1), (4096, 4096), (700, 300), (50, 0), (5000, 9000), (129, 257)]
    mf = api.MatchFactory(1.0, 60000.0)
    jobs, feats = [], []
    for i, (nq, nt) in enumerate(shapes):
        q, t = synth.make_pair(nq, max(nt, 1), seed=900 + i)
        t = t[:nt]
        feats.append((q, t))
        jobs.append((mf._params(T.OUT_DMATCH, T.MASK_NONE, api.Image(2 * i), api.Image(2 * i + 1)), _unity(api, q), _unity(api, t), None))
    got = mf.matchBatch(jobs)
    for (q, t), g, (p, _, _, _) in zip(feats, got, jobs):
        want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=60000.0, query_id=p.query_image_id,
                                                        target_id=p.target_image_id), q, t)
        if len(want) == 0:      # reference quirk: nothing valid -> the uncompacted all-invalid array is kept
            assert g.size() == len(q) and np.all(g.to_numpy()["invalid"] == 1)
        else:
            assert T.fields_equal(g.to_numpy(), want)


def test_full_size_c3_matches_oracle_and_simt_engine(api):
    """BASELINE config C3 at full size (65 536 x 65 536, the bench workload): tensor-core engine == CPU oracle == SIMT engine,
    field by field; plus the size-independent property that matching a set against itself returns every query at distance 0
    with the reference's lane-rule winner among its duplicates."""
    q, t = synth.make_pair(65536, 65536, seed=0x55524C43)
    mf = api.MatchFactory(0.6, 200.0 * 200.0)
    qu, tu = _unity(api, q), _unity(api, t)
    got = mf.generateDistanceMatches(api.Image(0), qu, api.Image(1), tu).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=200.0 * 200.0, rel_threshold=0.6), q, t)
    assert len(got) == len(want) and T.fields_equal(got, want)
    simt = api.MatchFactory(0.6, 200.0 * 200.0, engine=1).generateDistanceMatches(api.Image(0), qu, api.Image(1), tu).to_numpy()
    assert T.fields_equal(simt, want)
    me = mf.generateDistanceMatches(api.Image(0), tu, api.Image(0), tu).to_numpy()
    assert len(me) == len(t) and np.all(me["distance"] == 0.0)
    assert np.array_equal(me["keyPoints"]["loc"][:, 0], t["loc"])


# ------------------------------------------------------------------------------------ sharded targets (multi-GPU data path on one device)
def test_sharded_targets_merge_equals_single_call(api):
    import torch
    from ssrlcv_b200.match_factory import pack
    q, t = synth.make_pair(3000, 8192 + 512, seed=61)
    mf = api.MatchFactory(1.0, 1.0e5)
    qu = _unity(api, q); qu.setMemoryState(api.gpu)
    want = mf.generateDistanceMatches(api.Image(0), qu, api.Image(1), _unity(api, t)).to_numpy()
    ctx = mf.ctx
    bounds = [0, 2048, 4096, 6400, len(t)]
    keys = torch.empty((4, len(q)), dtype=torch.int64, device="cuda")
    p = mf._params(T.OUT_DMATCH, T.MASK_NONE, api.Image(0), api.Image(1))
    shards = []
    for r in range(4):
        su = _unity(api, t[bounds[r]:bounds[r + 1]]); su.setMemoryState(api.gpu); shards.append(su)
        ctx.check(ctx.lib.ssrlcv_b200_match_candidates(ctx.handle, C.byref(p), pack(ctx, qu).handle, pack(ctx, su).handle, bounds[r],
                                                       C.c_void_p(keys[r].data_ptr())))
    tloc = torch.from_numpy(np.ascontiguousarray(t["loc"])).cuda()
    out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
    n = C.c_uint64(0)
    ctx.check(ctx.lib.ssrlcv_b200_finalize_candidates(ctx.handle, C.byref(p), pack(ctx, qu).handle, C.c_void_p(keys.data_ptr()), 4,
                                                      C.c_void_p(tloc.data_ptr()), None, C.c_void_p(out.data_ptr()), C.byref(n)))
    got = out[: n.value * 48].cpu().numpy().view(T.DMATCH)
    assert T.fields_equal(got, want)


def test_sharded_targets_with_seed_distances(api, everest):
    """target shards + seed distances (masked and not): each shard tracks only candidates that can pass the seed-ratio test
    (ssrlcv_b200_match_candidates_seeded); the merged result equals the single call and the oracle."""
    import torch
    from ssrlcv_b200.match_factory import pack
    q, t = synth.make_pair(3000, 8192 + 512, seed=62)
    s = synth.make_seed(900)
    sd_want = api.O.seed_distances(q, s)
    mf = api.MatchFactory(0.8, 1.0e5)
    mf.setSeedFeatures(_unity(api, s))
    qu = _unity(api, q); qu.setMemoryState(api.gpu)
    sd = mf.getSeedDistances(qu); sd.setMemoryState(api.gpu)
    ctx = mf.ctx
    bounds = [0, 2048, 4096, 6400, len(t)]
    tloc = torch.from_numpy(np.ascontiguousarray(t["loc"])).cuda()
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    for mask_kind, kw in ((T.MASK_NONE, {}), (T.MASK_DOUBLE, dict(epsilon=25.0, delta=5.0))):
        p = mf._params(T.OUT_DMATCH, mask_kind, qi, ti, **kw)
        keys = torch.empty((4, len(q)), dtype=torch.int64, device="cuda")
        for r in range(4):
            su = _unity(api, t[bounds[r]:bounds[r + 1]]); su.setMemoryState(api.gpu)
            ctx.check(ctx.lib.ssrlcv_b200_match_candidates_seeded(ctx.handle, C.byref(p), pack(ctx, qu).handle, pack(ctx, su).handle, bounds[r],
                                                                  C.c_void_p(sd.device_ptr()), C.c_void_p(keys[r].data_ptr())))
        out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
        n = C.c_uint64(0)
        ctx.check(ctx.lib.ssrlcv_b200_finalize_candidates(ctx.handle, C.byref(p), pack(ctx, qu).handle, C.c_void_p(keys.data_ptr()), 4,
                                                          C.c_void_p(tloc.data_ptr()), C.c_void_p(sd.device_ptr()), C.c_void_p(out.data_ptr()), C.byref(n)))
        got = out[: n.value * 48].cpu().numpy().view(T.DMATCH)
        po = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=mask_kind, has_seed=True, abs_threshold=1.0e5, rel_threshold=0.8, query_id=0, target_id=1,
                               epsilon=25, delta=5, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
        want = api.O.generate_matches(po, q, t, sd_want)
        assert len(want) > 10 and T.fields_equal(got, want), mask_kind


# ------------------------------------------------------------------------------------ the reference's own CUDA kernels
def _ref_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "libssrlcv_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    cwd = os.getcwd()
    os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
    os.chdir("/tmp/ssrlcv_ref_run")          # the reference's global logger writes ./out/ssrlcv.log
    lib = C.CDLL(path)
    os.chdir(cwd)
    lib.ref_match.restype = C.c_long
    return lib


def _ref_match(lib, out_kind, mask_kind, abs_thr, rel, q, t, seed, eps=0.0, delta=0.0, F=None, qcam=None, tcam=None, qid=0, tid=1):
    out = np.zeros(len(q), T.OUT_DTYPES[out_kind])
    ms = C.c_double(0)
    cwd = os.getcwd(); os.chdir("/tmp/ssrlcv_ref_run")
    try:
        n = lib.ref_match(out_kind, mask_kind, C.c_float(abs_thr), C.c_float(rel), qid, tid,
                          q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)),
                          None if seed is None else seed.ctypes.data_as(C.POINTER(C.c_float)), 2, C.c_float(eps), C.c_float(delta),
                          None if F is None else np.ascontiguousarray(F, np.float32).ctypes.data_as(C.POINTER(C.c_float)),
                          None if qcam is None else np.asarray(qcam).tobytes(), None if tcam is None else np.asarray(tcam).tobytes(),
                          out.ctypes.data_as(C.c_void_p), C.byref(ms))
    finally:
        os.chdir(cwd)
    assert n >= 0
    return out[:n]


def test_against_reference_cuda_unconstrained(api):
    lib = _ref_lib()
    q, t = synth.make_pair(3000, 5000, seed=71)
    s = synth.make_seed(700)
    sd = api.O.seed_distances(q, s)
    mf = api.MatchFactory(0.8, 1.5e5)
    for out_kind, fn in ((T.OUT_MATCH, mf.generateMatches), (T.OUT_DMATCH, mf.generateDistanceMatches), (T.OUT_PAIR, mf.generateMatchesIndexOnly)):
        for seed in (None, sd):
            ref = _ref_match(lib, out_kind, 0, 1.5e5, 0.8, q, t, seed)
            su = None if seed is None else api.Unity(seed, np.dtype("<f4"), api.cpu)
            got = fn(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t), su).to_numpy()
            assert len(ref) > 50 and T.fields_equal(got, ref), (out_kind, seed is not None)


def test_against_reference_cuda_masked(api, everest):
    lib = _ref_lib()
    q, t = synth.make_pair(3000, 5000, seed=81)
    sd = api.O.seed_distances(q, synth.make_seed(700))
    su = api.Unity(sd, np.dtype("<f4"), api.cpu)
    mf = api.MatchFactory(0.8, 1.5e5)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    for out_kind, fn in ((T.OUT_MATCH, mf.generateMatchesDoubleConstrained), (T.OUT_DMATCH, mf.generateDistanceMatchesDoubleConstrained),
                         (T.OUT_PAIR, mf.generateMatchesDoubleConstrainedIndexOnly)):
        ref = _ref_match(lib, out_kind, 2, 1.5e5, 0.8, q, t, sd, eps=25.0, delta=5.0, qcam=everest.cameras[0], tcam=everest.cameras[1])
        got = fn(qi, _unity(api, q), ti, _unity(api, t), 25.0, 5.0, su).to_numpy()
        assert len(ref) > 10 and T.fields_equal(got, ref), out_kind
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    for out_kind, fn in ((T.OUT_MATCH, mf.generateMatchesConstrained), (T.OUT_DMATCH, mf.generateDistanceMatchesConstrained),
                         (T.OUT_PAIR, mf.generateMatchesConstrainedIndexOnly)):
        ref = _ref_match(lib, out_kind, 1, 1.5e5, 0.8, q, t, sd, eps=30.0, F=F)
        got = fn(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t), 30.0, F, su).to_numpy()
        assert len(ref) > 10 and T.fields_equal(got, ref), out_kind


def test_reference_cuda_reproduces_golden(api, everest):
    """sanity of the rebuilt reference on B200: it reproduces its own 2-view fixture."""
    lib = _ref_lib()
    ev = everest
    sd = np.zeros(len(ev.views2[0]), np.float32)
    cwd = os.getcwd(); os.chdir("/tmp/ssrlcv_ref_run")
    try:
        lib.ref_seed_distances(ev.views2[0].ctypes.data_as(C.c_void_p), C.c_ulong(len(sd)), ev.seed2.ctypes.data_as(C.c_void_p),
                               C.c_ulong(len(ev.seed2)), 2, sd.ctypes.data_as(C.POINTER(C.c_float)), None)
    finally:
        os.chdir(cwd)
    assert np.array_equal(sd, api.O.seed_distances(ev.views2[0], ev.seed2))
    ref = _ref_match(lib, T.OUT_DMATCH, 2, ev.abs_threshold, ev.rel_threshold, ev.views2[0], ev.views2[1], sd, eps=ev.epsilon, delta=ev.delta,
                     qcam=ev.cameras[0], tcam=ev.cameras[1], qid=ev.img_ids[0], tid=ev.img_ids[1])
    assert len(ref) == 13534
    assert T.fields_equal(ref["keyPoints"].reshape(-1).copy(), ev.g2_keypoints)


# ------------------------------------------------------------------------------------ the drop-in MatchFactory.cu
def _load_harness(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference at build time)")
    cwd = os.getcwd()
    os.makedirs("/tmp/ssrlcv_ref_run", exist_ok=True)
    os.chdir("/tmp/ssrlcv_ref_run")
    lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW | getattr(os, "RTLD_DEEPBIND", 0))
    os.chdir(cwd)
    lib.ref_match.restype = C.c_long
    lib.ref_exhaustive.restype = C.c_long
    lib.ref_refine_sort.restype = C.c_long
    return lib


def test_dropin_matchfactory_equals_reference(api, everest):
    """integration/MatchFactory_b200.cu compiled against the reference's own headers and linked with the reference's other
    translation units: the reference's public API, our kernels.  Same inputs through both libraries -> same records."""
    ref, drop = _load_harness("libssrlcv_ref.so"), _load_harness("libssrlcv_dropin.so")
    q, t = synth.make_pair(2500, 4000, seed=91)
    sd = api.O.seed_distances(q, synth.make_seed(600))
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    for out_kind in (T.OUT_MATCH, T.OUT_DMATCH, T.OUT_PAIR):
        for mask_kind, kw in ((0, {}), (1, dict(eps=30.0, F=F)), (2, dict(eps=25.0, delta=5.0, qcam=everest.cameras[0], tcam=everest.cameras[1]))):
            for seed in (None, sd):
                a = _ref_match(ref, out_kind, mask_kind, 1.5e5, 0.8, q, t, seed, **kw)
                b = _ref_match(drop, out_kind, mask_kind, 1.5e5, 0.8, q, t, seed, **kw)
                assert len(a) > 5 and T.fields_equal(a, b), (out_kind, mask_kind, seed is not None)


def _exhaustive(lib, ev):
    n = 3
    feats = (C.c_void_p * n)(*[v.ctypes.data for v in ev.views3])
    nf = (C.c_ulong * n)(*[len(v) for v in ev.views3])
    cam_bytes = [np.asarray(c).tobytes() for c in ev.cameras]
    cams = (C.c_char_p * n)(*cam_bytes)
    ids = (C.c_int * n)(*ev.img_ids)
    mm = np.zeros(40000, T.MULTIMATCH); kp = np.zeros(120000, T.KEYPOINT)
    nkp = C.c_ulong(0); ms = C.c_double(0)
    cwd = os.getcwd(); os.chdir("/tmp/ssrlcv_ref_run")
    try:
        nmm = lib.ref_exhaustive(n, feats, nf, C.cast(cams, C.POINTER(C.c_void_p)), ids, ev.seed3.ctypes.data_as(C.c_void_p), C.c_ulong(len(ev.seed3)),
                                 C.c_float(ev.rel_threshold), C.c_float(ev.abs_threshold), C.c_float(ev.epsilon), C.c_float(ev.delta), 3,
                                 mm.ctypes.data_as(C.c_void_p), C.c_ulong(len(mm)), kp.ctypes.data_as(C.c_void_p), C.c_ulong(len(kp)), C.byref(nkp), C.byref(ms))
    finally:
        os.chdir(cwd)
    assert nmm >= 0
    return mm[:nmm], kp[: nkp.value], ms.value


def test_dropin_three_view_golden(api, everest):
    """PipelineTest.FeatureMatching3View (ref: test/Pipeline.cu:218-255) through the drop-in's generateMatchesExhaustive."""
    drop = _load_harness("libssrlcv_dropin.so")
    mm, kp, _ = _exhaustive(drop, everest)
    assert T.fields_equal(mm, everest.g3_multimatch) and T.fields_equal(kp, everest.g3_keypoints)


def test_dropin_refine_sort_raw(api):
    ref, drop = _load_harness("libssrlcv_ref.so"), _load_harness("libssrlcv_dropin.so")
    q, t = synth.make_pair(3000, 3000, seed=93)
    base = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH), q, t)
    thr = float(np.median(base["distance"]))
    outs = []
    for lib in (ref, drop):
        m = base.copy(); raw = np.zeros(len(m), T.MATCH)
        cwd = os.getcwd(); os.chdir("/tmp/ssrlcv_ref_run")
        try:
            n = lib.ref_refine_sort(m.ctypes.data_as(C.c_void_p), C.c_ulong(len(m)), C.c_float(thr), 1, raw.ctypes.data_as(C.c_void_p))
        finally:
            os.chdir(cwd)
        outs.append((m[:n], raw[:n]))
    (ma, ra), (mb, rb) = outs
    assert len(ma) == len(mb) and np.array_equal(ma["distance"], mb["distance"])          # same multiset, same order of distances
    want = api.O.sort(api.O.refine(base, thr))
    assert T.fields_equal(mb, want) and T.fields_equal(rb, api.O.raw_matches(want))
    # the reference's thrust::sort is unstable: inside runs of equal distance only the set of records is defined
    key = lambda r: (float(r["distance"]), r["keyPoints"]["loc"].tobytes())
    assert sorted(map(key, ma)) == sorted(map(key, mb))


def test_do_feature_matching_two_view_golden(api, everest):
    """ref: src/Pipeline.cu:173-241 through the mirror in ssrlcv_b200/pipeline.py: the 2-view tail (D2H, max-distance scan,
    getRawMatches, host loop) is one device pass (ssrlcv_b200_match_set_2view); output = the reference's golden MatchSet
    (test/checkpoints/Pipeline2View, test/Pipeline.cu:183-216)."""
    from ssrlcv_b200.pipeline import FeatureMatchingInput, doFeatureMatching
    ev = everest
    inp = FeatureMatchingInput(images=[api.Image(ev.img_ids[0], ev.cameras[0]), api.Image(ev.img_ids[1], ev.cameras[1])],
                               allFeatures=[_unity(api, ev.views2[0]), _unity(api, ev.views2[1])], seedFeatures=_unity(api, ev.seed2),
                               epsilon=ev.epsilon, delta=ev.delta)
    out = doFeatureMatching(inp)
    kp, mm = out.matchSet.keyPoints.to_numpy(), out.matchSet.matches.to_numpy()
    assert len(mm) == 13534 and len(kp) == 27068
    assert T.fields_equal(kp, ev.g2_keypoints) and T.fields_equal(mm, ev.g2_multimatch)
    # the largest distance the reference logs: against the oracle's DMatch list
    sd = api.O.seed_distances(ev.views2[0], ev.seed2)
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=ev.abs_threshold, rel_threshold=ev.rel_threshold,
                          query_id=ev.img_ids[0], target_id=ev.img_ids[1], epsilon=ev.epsilon, delta=ev.delta,
                          query_camera=ev.cameras[0], target_camera=ev.cameras[1])
    want = api.O.generate_matches(p, ev.views2[0], ev.views2[1], sd)
    assert out.maxDistance == float(want["distance"].max())
    # empty list
    from ssrlcv_b200.pipeline import match_set_2view
    mf = api.MatchFactory(0.6, 4e4)
    ms, mx = match_set_2view(mf, api.Unity(np.zeros(0, T.DMATCH), T.DMATCH, api.cpu))
    assert ms.matches.size() == 0 and ms.keyPoints.size() == 0 and mx == 0.0


def test_against_reference_cuda_featurematch(api):
    """generateFeatureMatches (ref: src/MatchFactory.cu:654-697) against the reference's own kernels, with and without seed
    distances.  It returns all numQuery entries uncompacted: the reference writes only `invalid` and `distance` of an invalid
    entry (the rest of its local struct is uninitialised, :2325-2339), so those are compared on `invalid` + `distance` and the
    valid ones field by field.
    generateFeatureMatchesConstrained (:699-750) cannot be driven through the reference: its validateMatches overload for the raw
    Unity<FeatureMatch>* hands RAW device pointers to thrust::remove_if (:110-114; the other three overloads wrap them in
    thrust::device_ptr), which dispatches to the host backend and dereferences device memory on the CPU -- a segmentation fault
    on every call (observed on the B200 box).  Its kernel (:2342-2404, :2460-2530) shares the winner / threshold code with the
    constrained DMatch kernels that ARE compared with the reference (test_against_reference_cuda_masked); the compacted
    FeatureMatch records are checked against the oracle here."""
    lib = _ref_lib()
    q, t = synth.make_pair(2000, 3000, seed=72)
    s = synth.make_seed(500)
    sd = api.O.seed_distances(q, s)
    su = api.Unity(sd, np.dtype("<f4"), api.cpu)
    mf = api.MatchFactory(0.8, 6.0e4)
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    for seed, seed_u in ((None, None), (sd, su)):
        ref = _ref_match(lib, T.OUT_FEATUREMATCH, 0, 6.0e4, 0.8, q, t, seed)
        got = mf.generateFeatureMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t), seed_u).to_numpy()
        assert len(ref) == len(got) == len(q)
        assert np.array_equal(ref["invalid"] != 0, got["invalid"] != 0)
        v = ref["invalid"] == 0
        assert 50 < v.sum() < len(q)
        assert T.fields_equal(got[v], ref[v]), seed is not None
        # no candidate / d2 >= abs: distance is the threshold itself; an entry that failed the ratio test keeps its d2
        assert np.array_equal(got["distance"].view(np.uint32), ref["distance"].view(np.uint32))
        got = mf.generateFeatureMatchesConstrained(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t), 30.0, F, seed_u).to_numpy()
        want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_FEATUREMATCH, mask_kind=T.MASK_FUNDAMENTAL, has_seed=seed is not None,
                                                        abs_threshold=6.0e4, rel_threshold=0.8, epsilon=30.0, fundamental=F), q, t, seed)
        assert len(want) > 10 and T.fields_equal(got, want), seed is not None


def test_dropin_featurematch_equals_reference(api):
    """the same calls through the drop-in MatchFactory.cu (raw Unity<FeatureMatch>* API): unconstrained against the reference,
    constrained (which the reference cannot run, see above) against the oracle"""
    ref, drop = _load_harness("libssrlcv_ref.so"), _load_harness("libssrlcv_dropin.so")
    q, t = synth.make_pair(1500, 2500, seed=73)
    sd = api.O.seed_distances(q, synth.make_seed(400))
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    for seed in (None, sd):
        a = _ref_match(ref, T.OUT_FEATUREMATCH, 0, 6.0e4, 0.8, q, t, seed)
        b = _ref_match(drop, T.OUT_FEATUREMATCH, 0, 6.0e4, 0.8, q, t, seed)
        assert len(a) == len(b) == len(q) and np.array_equal(a["invalid"] != 0, b["invalid"] != 0)
        v = a["invalid"] == 0
        assert v.sum() > 50 and T.fields_equal(a[v], b[v])
        assert np.array_equal(a["distance"].view(np.uint32), b["distance"].view(np.uint32))
        b = _ref_match(drop, T.OUT_FEATUREMATCH, 1, 6.0e4, 0.8, q, t, seed, eps=30.0, F=F)
        want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_FEATUREMATCH, mask_kind=T.MASK_FUNDAMENTAL, has_seed=seed is not None,
                                                        abs_threshold=6.0e4, rel_threshold=0.8, query_id=0, target_id=1, epsilon=30.0, fundamental=F), q, t, seed)
        assert len(want) > 10 and T.fields_equal(b, want)


@pytest.mark.parametrize("out_kind", [T.OUT_MATCH, T.OUT_DMATCH, T.OUT_PAIR, T.OUT_FEATUREMATCH])
def test_validate_every_record_kind(api, out_kind):
    """validateMatches x4 (ref: src/MatchFactory.cu:31-135): stable removal of invalid entries for Match / DMatch / uint2_pair (self
    pair = invalid) / FeatureMatch, against numpy, against the reference's thrust::remove_if and through the drop-in."""
    q, t = synth.make_pair(3000, 3000, seed=52 + out_kind)
    p = api.O.make_params(out_kind=out_kind, query_id=3, target_id=8)
    base = api.O.generate_matches(p, q, t)
    assert len(base) == 3000
    dirty = base.copy()
    rng = np.random.default_rng(out_kind)
    bad = rng.random(len(dirty)) < 0.37
    if out_kind == T.OUT_PAIR:
        dirty["b"][bad] = dirty["a"][bad]
    else:
        dirty["invalid"][bad] = 1
    want = dirty[~bad]
    mf = api.MatchFactory(1.0, FLT_MAX)
    u = api.Unity(dirty.copy(), T.OUT_DTYPES[out_kind], api.cpu)
    mf.validateMatches(u)
    assert u.size() == len(want) and T.fields_equal(u.to_numpy(), want)
    # all invalid: the reference's quirk leaves the caller's array as it was
    allbad = base.copy()
    if out_kind == T.OUT_PAIR:
        allbad["b"] = allbad["a"]
    else:
        allbad["invalid"] = 1
    u = api.Unity(allbad.copy(), T.OUT_DTYPES[out_kind], api.cpu)
    mf.validateMatches(u)
    assert u.size() == len(allbad)
    for name in ("libssrlcv_ref.so", "libssrlcv_dropin.so"):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(path):
            continue
        if out_kind == T.OUT_FEATUREMATCH and name == "libssrlcv_ref.so":
            continue      # the reference's FeatureMatch overload crashes (raw pointers to thrust::remove_if, src/MatchFactory.cu:110-114)
        lib = _load_harness(name)
        lib.ref_validate.restype = C.c_long
        rec = dirty.copy()
        cwd = os.getcwd(); os.chdir("/tmp/ssrlcv_ref_run")
        try:
            n = lib.ref_validate(out_kind, rec.ctypes.data_as(C.c_void_p), C.c_ulong(len(rec)))
        finally:
            os.chdir(cwd)
        assert n == len(want) and T.fields_equal(rec[:n], want), name


def _exhaustive2(lib, ev, ordered, overlap):
    n = 3
    feats = (C.c_void_p * n)(*[v.ctypes.data for v in ev.views3])
    nf = (C.c_ulong * n)(*[len(v) for v in ev.views3])
    cam_bytes = [np.asarray(c).tobytes() for c in ev.cameras]
    cams = (C.c_char_p * n)(*cam_bytes)
    ids = (C.c_int * n)(*ev.img_ids)
    mm = np.zeros(40000, T.MULTIMATCH); kp = np.zeros(120000, T.KEYPOINT)
    nkp = C.c_ulong(0); ms = C.c_double(0)
    lib.ref_exhaustive2.restype = C.c_long
    cwd = os.getcwd(); os.chdir("/tmp/ssrlcv_ref_run")
    try:
        nmm = lib.ref_exhaustive2(n, feats, nf, C.cast(cams, C.POINTER(C.c_void_p)), ids, ev.seed3.ctypes.data_as(C.c_void_p), C.c_ulong(len(ev.seed3)),
                                  C.c_float(ev.rel_threshold), C.c_float(ev.abs_threshold), C.c_float(ev.epsilon), C.c_float(ev.delta), 3,
                                  int(ordered), C.c_float(overlap),
                                  mm.ctypes.data_as(C.c_void_p), C.c_ulong(len(mm)), kp.ctypes.data_as(C.c_void_p), C.c_ulong(len(kp)), C.byref(nkp), C.byref(ms))
    finally:
        os.chdir(cwd)
    assert nmm >= 0
    return mm[:nmm], kp[: nkp.value]


@pytest.mark.parametrize("ordered,overlap", [(True, 0.7), (True, 0.5), (False, 0.7)])
def test_exhaustive_estimated_overlap_skip_rule(api, everest, ordered, overlap):
    """The pair loop's skip rule for ordered image sets (ref: src/MatchFactory.cu:925-934; the counter is shared between the skip test
    and the push): 3 views at overlap 0.7 run pairs 0:1 and 0:2, at 0.5 only 0:1, unordered all three.
    The reference itself cannot run the ordered case: with `ordered && estimatedOverlap > 0` its `matchIndices[i++]` (:934) indexes
    one past the vector it has just pushed to (i was already incremented by the skip test) -- undefined behaviour, a crash in
    practice -- so the reference harness is driven only on the unordered case (no skip, same code path otherwise); the ordered
    cases check the drop-in and the Python mirror against each other and against the oracle's fusion of exactly the expected pairs."""
    from ssrlcv_b200.pair_scheduler import pair_order
    want_pairs = {(True, 0.7): [(0, 1), (0, 2)], (True, 0.5): [(0, 1)], (False, 0.7): [(0, 1), (0, 2), (1, 2)]}[(ordered, overlap)]
    assert pair_order(3, ordered, overlap) == want_pairs
    ev = everest
    drop = _load_harness("libssrlcv_dropin.so")
    mm_d, kp_d = _exhaustive2(drop, ev, ordered, overlap)
    assert len(mm_d) > 1000
    if not ordered:
        mm_r, kp_r = _exhaustive2(_load_harness("libssrlcv_ref.so"), ev, ordered, overlap)
        assert T.fields_equal(mm_d, mm_r) and T.fields_equal(kp_d, kp_r)
        assert T.fields_equal(mm_d, ev.g3_multimatch) and T.fields_equal(kp_d, ev.g3_keypoints)
    mf = api.MatchFactory(ev.rel_threshold, ev.abs_threshold)
    mf.setSeedFeatures(_unity(api, ev.seed3))
    imgs = [api.Image(ev.img_ids[i], ev.cameras[i]) for i in range(3)]
    ms = mf.generateMatchesExhaustive(imgs, [_unity(api, v) for v in ev.views3], ev.epsilon, ev.delta, ordered, overlap)
    assert T.fields_equal(ms.matches.to_numpy(), mm_d) and T.fields_equal(ms.keyPoints.to_numpy(), kp_d)
    # oracle: pair lists of exactly the expected pairs, fused
    lists = []
    for (a, b) in want_pairs:
        sd = api.O.seed_distances(ev.views3[a], ev.seed3)
        p = api.O.make_params(out_kind=T.OUT_PAIR, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=ev.abs_threshold, rel_threshold=ev.rel_threshold,
                              query_id=ev.img_ids[a], target_id=ev.img_ids[b], epsilon=ev.epsilon, delta=ev.delta,
                              query_camera=ev.cameras[a], target_camera=ev.cameras[b])
        lists.append(api.O.generate_matches(p, ev.views3[a], ev.views3[b], sd))
    mm_o, kp_o = api.O.fuse_nview(ev.views3, lists)
    assert T.fields_equal(mm_d, mm_o) and T.fields_equal(kp_d, kp_o)


def test_exact_baseline_shapes_c2_and_c4_pair(api, everest):
    """BASELINE.json's configs at their exact shapes against the CPU oracle: C2 = 16 384 x 16 384 + 8 192 seed features, seed-ratio test +
    DoubleConstrained (the bench's parameters), and one pair of C4 = 32 768 x 32 768 generateMatchesIndexOnly."""
    q, t = synth.make_pair(16384, 16384, seed=0x55524C43)
    s = synth.make_seed(8192, seed=99)
    mf = api.MatchFactory(0.6, 200.0 * 200.0)
    mf.setSeedFeatures(_unity(api, s))
    qu, tu = _unity(api, q), _unity(api, t)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    sd = mf.getSeedDistances(qu)
    sd_want = api.O.seed_distances(q, s)
    assert np.array_equal(sd.to_numpy(), sd_want)
    got = mf.generateDistanceMatchesDoubleConstrained(qi, qu, ti, tu, 25.0, 5.0, sd).to_numpy()
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=200.0 * 200.0, rel_threshold=0.6,
                          query_id=0, target_id=1, epsilon=25.0, delta=5.0, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    want = api.O.generate_matches(p, q, t, sd_want)
    assert T.fields_equal(got, want)
    # the same shape without the mask must find the planted correspondences (a non-empty answer at this exact shape)
    got = mf.generateDistanceMatches(qi, qu, ti, tu, sd).to_numpy()
    p.mask_kind = T.MASK_NONE
    want = api.O.generate_matches(p, q, t, sd_want)
    assert len(want) > 1000 and T.fields_equal(got, want)
    q, t = synth.make_pair(32768, 32768, seed=0x55524C43 + 3)
    got = mf.generateMatchesIndexOnly(api.Image(6), _unity(api, q), api.Image(7), _unity(api, t)).to_numpy()
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_PAIR, abs_threshold=200.0 * 200.0, rel_threshold=0.6, query_id=6, target_id=7), q, t)
    assert len(want) > 5000 and T.fields_equal(got, want)


def test_match_features_one_call(api, everest):
    """ssrlcv_b200_match_features: pack + seed distances + match + release on device-resident AoS arrays in one entry (the bench's
    step): seeded DoubleConstrained DMatch and unseeded Match, records and seed distances equal the oracle's; a target set whose
    norms do not fit the fold digits (the seed pass and the match are first launched on the optimistic guess, then redone)."""
    import torch
    from ssrlcv_b200.match_factory import Context
    ctx = Context.get(0)
    lib, h = ctx.lib, ctx.handle
    q, t = synth.make_pair(3000, 4100, seed=4242)
    s = synth.make_seed(640)
    dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()
    mf = api.MatchFactory(0.8, 1.0e5)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    for bright_seed in (False, True):
        s2 = s.copy()
        if bright_seed:
            s2["values"][5] = 255
        qd, td, sd_ = dev(q), dev(t), dev(s2)
        sd_want = api.O.seed_distances(q, s2)
        for out_kind, mask_kind, dt in ((T.OUT_DMATCH, T.MASK_DOUBLE, T.DMATCH), (T.OUT_MATCH, T.MASK_NONE, T.MATCH)):
            p = mf._params(out_kind, mask_kind, qi, ti, 25.0, 5.0)
            out = torch.empty(len(q) * dt.itemsize, dtype=torch.uint8, device="cuda")
            seed_out = torch.zeros(len(q), dtype=torch.float32, device="cuda")
            n_out = C.c_uint64(0)
            ctx.check(lib.ssrlcv_b200_match_features(h, C.byref(p), C.c_void_p(qd.data_ptr()), len(q), C.c_void_p(td.data_ptr()), len(t),
                                                     C.c_void_p(sd_.data_ptr()), len(s2), C.c_void_p(seed_out.data_ptr()), C.c_void_p(out.data_ptr()), C.byref(n_out)))
            assert np.array_equal(seed_out.cpu().numpy(), sd_want)
            po = api.O.make_params(out_kind=out_kind, mask_kind=mask_kind, has_seed=True, abs_threshold=1.0e5, rel_threshold=0.8, query_id=0, target_id=1,
                                   epsilon=25.0, delta=5.0, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
            want = api.O.generate_matches(po, q, t, sd_want)
            got = out[: n_out.value * dt.itemsize].cpu().numpy().view(dt)
            assert len(want) >= 10 and T.fields_equal(got, want), (bright_seed, out_kind)
    p = mf._params(T.OUT_PAIR, T.MASK_NONE, qi, ti)
    out = torch.empty(len(q) * 16, dtype=torch.uint8, device="cuda")
    n_out = C.c_uint64(0)
    ctx.check(lib.ssrlcv_b200_match_features(h, C.byref(p), C.c_void_p(qd.data_ptr()), len(q), C.c_void_p(td.data_ptr()), len(t), None, 0, None,
                                             C.c_void_p(out.data_ptr()), C.byref(n_out)))
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_PAIR, abs_threshold=1.0e5, rel_threshold=0.8), q, t)
    assert T.fields_equal(out[: n_out.value * 16].cpu().numpy().view(T.UINT2_PAIR), want)


# ------------------------------------------------------------------------------------ the stage, stand-alone, from checkpoints
@pytest.mark.parametrize("n_views", [2, 3])
def test_stage_runner_from_checkpoint_tree(api, everest, tmp_path, n_views):
    """ref: src/Pipeline.cu:144-163 + src/SFM.cu:247-259 -- the matching stage from a feature-generation checkpoint directory (the
    layout of the reference's own fixtures, test/checkpoints/Pipeline{2,3}View) to the stage's output checkpoints:
    0_...KeyPointE.uty / 0_...MultiMatchE.uty hold the reference's golden MatchSet."""
    from ssrlcv_b200 import checkpoint as ck
    from ssrlcv_b200.pipeline import main as stage_main
    from tests.test_abi import _write_stage_inputs
    fg, out = str(tmp_path / "fg"), str(tmp_path / "fm")
    os.makedirs(fg)
    _write_stage_inputs(fg, everest, n_views)
    assert stage_main(["--fg", fg, "--out", out, "-n", str(n_views), "--epsilon", "25", "--delta", "5"]) == 0
    kp = ck.read_uty(ck.checkpoint_path(out, "keypoint", 0), "keypoint")
    mm = ck.read_uty(ck.checkpoint_path(out, "multimatch", 0), "multimatch")
    want_kp, want_mm = (everest.g2_keypoints, everest.g2_multimatch) if n_views == 2 else (everest.g3_keypoints, everest.g3_multimatch)
    assert T.fields_equal(kp.data, want_kp) and T.fields_equal(mm.data, want_mm)
    assert kp.state == ck.STATE_CPU and os.path.exists(os.path.join(out, "done"))


def test_match_set_stays_on_the_device_for_the_consumer(api, everest):
    """SURVEY 8f-4, consumer side: generateBundles moves matches and keyPoints to the GPU first thing (ref:
    src/PointCloudFactory.cu:837-838); doFeatureMatching(device_resident=True) hands them over in state gpu (no D2H + H2D pair)."""
    from ssrlcv_b200.pipeline import FeatureMatchingInput, doFeatureMatching
    ev = everest
    inp = FeatureMatchingInput(images=[api.Image(ev.img_ids[0], ev.cameras[0]), api.Image(ev.img_ids[1], ev.cameras[1])],
                               allFeatures=[_unity(api, ev.views2[0]), _unity(api, ev.views2[1])], seedFeatures=_unity(api, ev.seed2),
                               epsilon=ev.epsilon, delta=ev.delta)
    out = doFeatureMatching(inp, device_resident=True)
    assert out.matchSet.keyPoints.getMemoryState() == api.gpu and out.matchSet.matches.getMemoryState() == api.gpu
    assert out.matchSet.keyPoints.host is None
    assert T.fields_equal(out.matchSet.keyPoints.to_numpy(), ev.g2_keypoints) and T.fields_equal(out.matchSet.matches.to_numpy(), ev.g2_multimatch)


# ------------------------------------------------------------------------------------ producer hand-off
def test_producer_handoff_packed_writer(api):
    """A feature factory that writes the packed layout itself (include/ssrlcv_b200_packed.cuh: the patched tail of fillDescriptors,
    src/SIFT_FeatureFactory.cu:533-548) + ssrlcv_b200_set_commit gives sets that are byte-identical to what the pack kernel makes of
    the 152-byte AoS records -- descriptors, fold digits, norms, W, window bounds, locations, sigma/theta, padding included -- and
    the matcher returns the oracle's records on them.  No pack launch on that route."""
    import torch
    from ssrlcv_b200 import _lib
    from ssrlcv_b200.match_factory import Context
    path = os.path.join(ROOT, "tests", "native", "libpacked_writer_test.so")
    if not os.path.exists(path):
        pytest.skip("tests/native/libpacked_writer_test.so not built (__graft_entry__.build)")
    helper = C.CDLL(path)
    ctx = Context.get(0)
    lib, h = ctx.lib, ctx.handle
    stream = lib.ssrlcv_b200_stream(h)
    q, t = synth.make_pair(3000 + 77, 5000 + 13, seed=1234)
    rng = np.random.default_rng(3)
    q["sigma"], q["theta"] = rng.uniform(1, 4, len(q)), rng.uniform(0, 6.28, len(q))
    t["sigma"], t["theta"] = rng.uniform(1, 4, len(t)), rng.uniform(0, 6.28, len(t))
    t["values"][17] = 255                      # one row whose norm does not fit the fold digits: the verdict must travel as well
    handles = []

    def download(hs, which):
        nbytes = C.c_uint64(0)
        ctx.check(lib.ssrlcv_b200_set_download(h, hs, which, None, 0, C.byref(nbytes)))
        buf = np.zeros(nbytes.value, np.uint8)
        ctx.check(lib.ssrlcv_b200_set_download(h, hs, which, C.c_void_p(buf.ctypes.data), nbytes.value, C.byref(nbytes)))
        return buf

    for feats in (q, t):
        dev = torch.from_numpy(feats.view(np.uint8).reshape(-1)).cuda()
        packed = C.c_void_p()
        ctx.check(lib.ssrlcv_b200_pack(h, C.c_void_p(dev.data_ptr()), len(feats), C.byref(packed)))
        direct = C.c_void_p()
        launches0 = ctx.launch_count()
        ctx.check(lib.ssrlcv_b200_set_alloc(h, len(feats), C.byref(direct)))
        arrays = _lib.SetArrays()
        ctx.check(lib.ssrlcv_b200_set_arrays_of(direct, C.byref(arrays)))
        assert arrays.n == len(feats) and arrays.n_pad % 256 == 0
        assert helper.packed_writer_fill(C.byref(arrays), C.c_void_p(dev.data_ptr()), C.c_ulong(len(feats)), C.c_void_p(stream)) == 0
        ctx.check(lib.ssrlcv_b200_set_commit(h, direct))
        assert ctx.launch_count() - launches0 == 1          # the commit launch only
        for which in range(8):
            assert np.array_equal(download(packed, which), download(direct, which)), which
        lib.ssrlcv_b200_set_free(packed)
        handles.append(direct)
    mf = api.MatchFactory(1.0, 1.2e5)
    p = mf._params(T.OUT_DMATCH, T.MASK_NONE, api.Image(0), api.Image(1))
    out = torch.empty(len(q) * 48, dtype=torch.uint8, device="cuda")
    n_out = C.c_uint64(0)
    ctx.check(lib.ssrlcv_b200_match(h, C.byref(p), handles[0], handles[1], None, 0, C.c_void_p(out.data_ptr()), C.byref(n_out)))
    got = out[: n_out.value * 48].cpu().numpy().view(T.DMATCH)
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=1.2e5), q, t)
    assert len(want) > 500 and T.fields_equal(got, want)
    for hs in handles:
        lib.ssrlcv_b200_set_free(hs)


# ------------------------------------------------------------------------------------ the pair scheduler across devices (C ABI)
def _multi_gpu_case(api, everest, n_devices):
    """ssrlcv_b200_match_pairs_multi_gpu: 5 views of different sizes, all 10 pairs, seed-ratio test + DoubleConstrained (what
    generateMatchesExhaustive runs under GEO_ORBIT=1), views handed over as device arrays, pinned-free host arrays and a mix;
    every pair list equals the single-GPU batch and the oracle."""
    from ssrlcv_b200.match_factory import MultiGpu
    from ssrlcv_b200.pair_scheduler import pair_order
    mg = MultiGpu.get(n_devices)
    sizes = (1500, 2300, 900, 3100, 1280)
    views = [synth.make_pair(n, 16, seed=400 + i)[0] for i, n in enumerate(sizes)]
    seed = synth.make_seed(500)
    mf = api.MatchFactory(0.8, 9.0e4)
    imgs = [api.Image(i, everest.cameras[i % 3]) for i in range(5)]
    pairs = pair_order(5)
    params = [mf._params(T.OUT_PAIR, T.MASK_DOUBLE, imgs[a], imgs[b], 2000.0, 5.0) for a, b in pairs]
    dev_of = mg.assignment(list(sizes), pairs)
    assert len(set(dev_of)) == mg.device_count() and len(dev_of) == 10
    want = []
    for (a, b), p in zip(pairs, params):
        sd = api.O.seed_distances(views[a], seed)
        po = api.O.make_params(out_kind=T.OUT_PAIR, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=9.0e4, rel_threshold=0.8, query_id=a, target_id=b,
                               epsilon=2000.0, delta=5.0, query_camera=everest.cameras[a % 3], target_camera=everest.cameras[b % 3])
        want.append(api.O.generate_matches(po, views[a], views[b], sd))
    assert sum(len(x) for x in want) > 200
    on_dev = [_unity(api, v) for v in views]
    for u in on_dev:
        u.setMemoryState(api.gpu)
    for feats, sd_in in ((on_dev, _unity(api, seed)), (views, seed), ([on_dev[0], views[1], on_dev[2], views[3], on_dev[4]], seed)):
        got = mg.match_pairs(feats, pairs, params, seed=sd_in)
        for g, w_ in zip(got, want):
            if len(w_) == 0:
                assert g.size() == 0
            else:
                assert T.fields_equal(g.to_numpy(), w_)
    # unseeded, DMatch, unmasked: other record kind through the same entry
    params2 = [mf._params(T.OUT_DMATCH, T.MASK_NONE, imgs[a], imgs[b]) for a, b in pairs]
    got = mg.match_pairs(on_dev, pairs, params2)
    for (a, b), g in zip(pairs, got):
        w_ = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=9.0e4, query_id=a, target_id=b), views[a], views[b])
        assert T.fields_equal(g.to_numpy(), w_)


def test_multi_gpu_entry_on_one_device(api, everest):
    _multi_gpu_case(api, everest, 1)


def test_multi_gpu_entry_on_every_device(api, everest):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    _multi_gpu_case(api, everest, 0)


# ------------------------------------------------------------------------------------ size-independent properties at full size
def test_c5_size_self_match_and_sharded_merge(api):
    """BASELINE C5 size (1 048 576 x 1 048 576), where the CPU oracle would take minutes: a set matched against ITSELF has a known
    answer -- every query finds distance 0 at the member of its group of identical descriptors with the smallest key
    (f mod 32, f) (the reference's lane rule, src/MatchFactory.cu:1476-1492) -- and the answer must not depend on how the
    target set is sharded (4 shards merged through the candidate keys == one call)."""
    import torch
    from ssrlcv_b200.match_factory import pack
    n = 1 << 20
    rng = np.random.default_rng(2026)
    base = synth.sift_descriptors(rng, 1 << 16)
    v = base[rng.integers(0, len(base), n)].copy()                                  # many exact duplicates, all over the index range
    v[:, :4] = (v[:, :4].astype(np.int32) + rng.integers(0, 4, (n, 4))).astype(np.uint8)   # ... thinned out to groups of a few
    f = T.make_features(v, rng.uniform(0, 4096, (n, 2)).astype(np.float32))
    # expected winner per query: smallest (f % 32, f) within its group of identical descriptors
    _, inv = np.unique(v.view(np.dtype((np.void, 128))).reshape(-1), return_inverse=True)
    idx = np.arange(n, dtype=np.int64)
    key = ((idx & 31) << 32) | idx
    best = np.full(inv.max() + 1, np.iinfo(np.int64).max)
    np.minimum.at(best, inv, key)
    want_t = (best[inv] & 0xFFFFFFFF).astype(np.uint32)
    assert (want_t != idx).sum() > 1000                                              # the tie rule is really exercised
    mf = api.MatchFactory(1.0, FLT_MAX)
    u = _unity(api, f); u.setMemoryState(api.gpu)
    got = mf.generateMatchesIndexOnly(api.Image(0), u, api.Image(1), u).to_numpy()
    assert len(got) == n
    assert np.array_equal(got["a"][:, 1], idx.astype(np.uint32)) and np.array_equal(got["b"][:, 1], want_t)
    dm = mf.generateDistanceMatches(api.Image(0), u, api.Image(1), u).to_numpy()
    assert len(dm) == n and not dm["distance"].any()
    # sharded: 4 target shards (boundaries multiples of 32), merged candidate keys
    ctx = mf.ctx
    p = mf._params(T.OUT_PAIR, T.MASK_NONE, api.Image(0), api.Image(1))
    bounds = [0, 262144 + 256, 524288, 786432 - 512, n]
    keys = torch.empty((4, n), dtype=torch.int64, device="cuda")
    hq = pack(ctx, u)
    for r in range(4):
        su = _unity(api, f[bounds[r]:bounds[r + 1]]); su.setMemoryState(api.gpu)
        ctx.check(ctx.lib.ssrlcv_b200_match_candidates(ctx.handle, C.byref(p), hq.handle, pack(ctx, su).handle, bounds[r], C.c_void_p(keys[r].data_ptr())))
    out = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
    cnt = C.c_uint64(0)
    ctx.check(ctx.lib.ssrlcv_b200_finalize_candidates(ctx.handle, C.byref(p), hq.handle, C.c_void_p(keys.data_ptr()), 4, None, None,
                                                      C.c_void_p(out.data_ptr()), C.byref(cnt)))
    merged = out[: cnt.value * 16].cpu().numpy().view(T.UINT2_PAIR)
    assert T.fields_equal(merged, got)


def test_random_sweep_against_oracle(api, everest):
    """seeded sweep over shapes (ragged, tiny, one-sided), thresholds, seeds, mask and output kinds: every case equals the oracle"""
    rng = np.random.default_rng(7)
    F = np.array([[0, 0, 0.001], [0, 0, -1.0], [-0.001, 1.0, 2.0]], np.float32)
    qi, ti = api.Image(2, everest.cameras[0]), api.Image(5, everest.cameras[1])
    n_empty = 0
    for case in range(36):
        nq = int(rng.choice([1, 31, 127, 128, 129, 257, 700, 1500, 4097]))
        nt = int(rng.choice([1, 32, 255, 256, 257, 1000, 2047, 5000, 12000]))
        thr = float(rng.choice([FLT_MAX, 2.0e5, 6.0e4, 3.0e4]))
        rel = float(rng.choice([1.0, 0.9, 0.7]))
        out_kind = int(rng.choice([T.OUT_MATCH, T.OUT_DMATCH, T.OUT_PAIR]))
        mask_kind = int(rng.choice([T.MASK_NONE, T.MASK_NONE, T.MASK_FUNDAMENTAL, T.MASK_DOUBLE]))
        seeded = bool(rng.integers(0, 2))
        q, t = synth.make_pair(nq, nt, seed=1000 + case, uniform_bytes=bool(case % 7 == 3))
        s = synth.make_seed(int(rng.choice([1, 200, 900])), seed=case)
        mf = api.MatchFactory(rel, thr)
        mf.setSeedFeatures(_unity(api, s))
        qu, tu = _unity(api, q), _unity(api, t)
        sd = mf.getSeedDistances(qu) if seeded else None
        sd_want = api.O.seed_distances(q, s) if seeded else None
        p = mf._params(out_kind, mask_kind, qi, ti, epsilon=30.0, delta=5.0, fundamental=F)
        got = mf._run(p, qu, tu, sd).to_numpy()
        po = api.O.make_params(out_kind=out_kind, mask_kind=mask_kind, has_seed=seeded, abs_threshold=thr, rel_threshold=rel, query_id=2, target_id=5,
                               epsilon=30.0, delta=5.0, fundamental=F, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
        want = api.O.generate_matches(po, q, t, sd_want)
        if len(want) == 0:      # reference quirk (src/MatchFactory.cu:40-45): nothing valid -> the caller keeps the uncompacted array
            assert len(got) == nq, (case, nq, nt)
            if out_kind == T.OUT_PAIR:
                assert np.array_equal(got["a"], got["b"])            # uint2_pair marks "invalid" as a self pair
            else:
                assert np.all(got["invalid"] == 1)
            n_empty += 1
        else:
            assert T.fields_equal(got, want), (case, nq, nt, thr, rel, out_kind, mask_kind, seeded)
    assert n_empty < 18


@pytest.mark.parametrize("thr", [FLT_MAX, 1.2e5])
def test_masked_without_seeds_on_spatially_sorted_sets(api, everest, knobs, thr):
    """Masked searches without seed distances run on Z-curve-sorted copies of BOTH sets (api.cu sorted_of): 32-target windows and
    (query tile, target tile) pairs whose boxes an epipolar band misses are skipped (geometry.cuh box_skips, exact.cu cull_kernel).
    NaN / negative / repeated coordinates on both sides (a NaN coordinate passes the reference's mask and must never be culled), a
    ragged last window, both bookkeeping regimes, both mask kinds -- same records as the oracle and as the unsorted path, in the
    caller's query order."""
    q, t = synth.make_pair(2500, 8192 + 77, seed=123)
    rng = np.random.default_rng(9)
    t["loc"][rng.choice(len(t), 40, replace=False), 0] = np.nan          # NaN coordinates pass the reference's mask
    t["loc"][rng.choice(len(t), 40, replace=False), 0] = -5.0
    t["loc"][rng.choice(len(t), 300, replace=False), 0] = 512.0           # many equal keys
    t["loc"][rng.choice(len(t), 10, replace=False), 1] = np.nan
    q["loc"][rng.choice(len(q), 12, replace=False), 0] = np.nan          # a query without a location: its mask state is NaN, nothing is masked
    q["loc"][rng.choice(len(q), 12, replace=False), 1] = -40.0
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, abs_threshold=thr, query_id=0, target_id=1, epsilon=40, delta=5,
                          query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    want = api.O.generate_matches(p, q, t)
    assert len(want) > 50
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    pf = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_FUNDAMENTAL, abs_threshold=thr, query_id=0, target_id=1, epsilon=30.0, fundamental=F)
    want_f = api.O.generate_matches(pf, q, t)
    assert len(want_f) > 50
    # sorted by the one-block counting sort over image cells / by the radix sort of Morton keys / not at all; and every way the rare
    # path evaluates the mask (warp-cooperative, candidate walk, all 32 columns)
    for so, cells, coop, walk in (("1", None, None, None), ("1", "0", None, None), ("0", None, None, None), ("1", None, "0", "32"),
                                  ("1", None, "0", "0"), ("1", None, "32", "0"), ("0", None, "2", "3")):
        knobs(sorted=so, cells=cells, mask_coop=coop, mask_walk=walk)
        mf = api.MatchFactory(1.0, thr)
        got = mf.generateDistanceMatchesDoubleConstrained(qi, _unity(api, q), ti, _unity(api, t), 40.0, 5.0).to_numpy()
        assert T.fields_equal(got, want), (so, cells, coop, walk)
        got = mf.generateDistanceMatchesConstrained(qi, _unity(api, q), ti, _unity(api, t), 30.0, F).to_numpy()
        assert T.fields_equal(got, want_f), (so, cells, coop, walk)
    knobs(sorted=None, cells=None, mask_coop=None, mask_walk=None)
    pr = api.MatchFactory(1.0, thr).generateMatchesDoubleConstrainedIndexOnly(qi, _unity(api, q), ti, _unity(api, t), 40.0, 5.0).to_numpy()
    p.out_kind = T.OUT_PAIR
    assert T.fields_equal(pr, api.O.generate_matches(p, q, t))
    # a batch of sorted and unsorted pairs in one launch, sets shared between pairs (the sorted copy is cached on the set)
    knobs(sorted="1")
    mf = api.MatchFactory(1.0, thr)
    qu, tu = _unity(api, q), _unity(api, t)
    jobs = [(mf._params(T.OUT_DMATCH, T.MASK_DOUBLE, qi, ti, 40.0, 5.0), qu, tu, None),
            (mf._params(T.OUT_DMATCH, T.MASK_NONE, qi, ti), qu, tu, None),
            (mf._params(T.OUT_DMATCH, T.MASK_FUNDAMENTAL, qi, ti, 30.0, 0.0, F), qu, tu, None)]
    got = mf.matchBatch(jobs)
    assert T.fields_equal(got[0].to_numpy(), want) and T.fields_equal(got[2].to_numpy(), want_f)
    pn = api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=thr, query_id=0, target_id=1)
    assert T.fields_equal(got[1].to_numpy(), api.O.generate_matches(pn, q, t))


@pytest.mark.parametrize("thr", [FLT_MAX, 3.0e5])
def test_sorted_masked_search_with_ties_everywhere(api, everest, knobs, thr):
    """Masked searches on spatially sorted sets when ties are the rule: few distinct descriptors repeated all over the target set, so
    nearly every query ties in three or more windows of several units -- the tie pass (exact.cu, rows mapped back through q_orig /
    t_orig) and the lane rule on the CALLER's indices decide the records.  Same as the oracle, under both bookkeeping regimes."""
    rng = np.random.default_rng(77)
    base = synth.sift_descriptors(rng, 6)
    q, t = synth.make_pair(700, 4096 + 300, seed=5)
    t["values"] = base[rng.integers(0, 6, len(t))]
    q["values"][:500] = base[rng.integers(0, 6, 500)]
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    pf = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_FUNDAMENTAL, abs_threshold=thr, query_id=0, target_id=1, epsilon=60.0, fundamental=F)
    want_f = api.O.generate_matches(pf, q, t)
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, abs_threshold=thr, query_id=0, target_id=1, epsilon=80, delta=5,
                          query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    want = api.O.generate_matches(p, q, t)
    assert len(want_f) > 100 and len(want) > 100
    for cm in ("1", "0"):
        knobs(sorted="1", colmask=cm)
        mf = api.MatchFactory(1.0, thr)
        got = mf.generateDistanceMatchesConstrained(qi, _unity(api, q), ti, _unity(api, t), 60.0, F).to_numpy()
        assert T.fields_equal(got, want_f), cm
        got = mf.generateDistanceMatchesDoubleConstrained(qi, _unity(api, q), ti, _unity(api, t), 80.0, 5.0).to_numpy()
        assert T.fields_equal(got, want), cm


@pytest.mark.parametrize("thr", [FLT_MAX, 2.5e5])
def test_runner_up_search_is_the_second_smallest_key(api, everest, knobs, thr):
    """ssrlcv_b200_match_runner_up (an extension; the reference has no second-best output): a second search of the pair with every
    query's winner left out returns the second-smallest key (d2, f mod 32, f) of the brute-force distance matrix -- duplicates,
    ragged last window and all -- under every matcher variant; the winner keys themselves are what the plain search returns, so the
    top-1 path is untouched.  Masked searches: the tensor-core path equals the independent SIMT engine."""
    import torch
    from ssrlcv_b200.match_factory import pack
    q, t = synth.make_pair(1500, 5000 + 13, seed=31, dup=0.02)
    qv, tv = q["values"].astype(np.int64), t["values"].astype(np.int64)
    d2 = (qv * qv).sum(1)[:, None] + (tv * tv).sum(1)[None, :] - 2 * qv @ tv.T
    f = np.arange(len(t), dtype=np.int64)[None, :]
    key = (d2 << 37) | ((f & 31) << 32) | f
    key = np.where(d2.astype(np.float32) < np.float32(thr), key, -1).astype(np.uint64)       # all-ones: not below the absolute threshold
    two = np.sort(key, axis=1)[:, :2]
    mf = api.MatchFactory(1.0, thr)
    ctx = mf.ctx
    qu, tu = _unity(api, q), _unity(api, t)
    qu.setMemoryState(api.gpu); tu.setMemoryState(api.gpu)
    p = mf._params(T.OUT_DMATCH, T.MASK_NONE, api.Image(0), api.Image(1))
    k1 = torch.empty(len(q), dtype=torch.int64, device="cuda")
    k2 = torch.empty(len(q), dtype=torch.int64, device="cuda")

    def top2(pp):
        ctx.check(ctx.lib.ssrlcv_b200_match_candidates(ctx.handle, C.byref(pp), pack(ctx, qu).handle, pack(ctx, tu).handle, 0, C.c_void_p(k1.data_ptr())))
        ctx.check(ctx.lib.ssrlcv_b200_match_runner_up(ctx.handle, C.byref(pp), pack(ctx, qu).handle, pack(ctx, tu).handle,
                                                      C.c_void_p(k1.data_ptr()), C.c_void_p(k2.data_ptr())))
        ctx.check(ctx.lib.ssrlcv_b200_synchronize(ctx.handle))
        return k1.cpu().numpy().view(np.uint64).copy(), k2.cpu().numpy().view(np.uint64).copy()

    for fold, cm, pair in [(None, None, None), ("0", None, None), (None, "0", None), (None, "1", "1"), ("0", "0", "1"), (None, None, "0")]:
        knobs(fold=fold, colmask=cm, pair=pair)
        a, b = top2(p)
        assert np.array_equal(a, two[:, 0]), (fold, cm, pair)
        assert np.array_equal(b, two[:, 1]), (fold, cm, pair)
    knobs(fold=None, colmask=None, pair=None)
    # masked: tensor-core path == SIMT cross-check engine (two independent implementations of the exclusion)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    for pm in (mf._params(T.OUT_DMATCH, T.MASK_DOUBLE, qi, ti, 60.0, 5.0), mf._params(T.OUT_DMATCH, T.MASK_FUNDAMENTAL, qi, ti, 40.0, 0.0, F)):
        a, b = top2(pm)
        ctx.tune(_lib_mod().TUNE_CROSSCHECK_SIMT, 1)
        try:
            a2, b2 = top2(pm)
        finally:
            ctx.tune(_lib_mod().TUNE_CROSSCHECK_SIMT, -1)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
        assert (b >= a).all() and (b[a != np.uint64(0xFFFFFFFFFFFFFFFF)] != a[a != np.uint64(0xFFFFFFFFFFFFFFFF)]).all()


@pytest.mark.parametrize("nq,nt", [(300, 700), (513, 4097), (4096 + 77, 8192 + 5)])
@pytest.mark.parametrize("thr", [FLT_MAX, 200.0 * 200.0])
def test_cta_pair_walk_matches_oracle(api, knobs, nq, nt, thr):
    """smax_duo.cu (tcgen05 cta_group::2 on a cluster of two SMs, ssrlcv_b200_tune(PAIR_UNITS, 2); measured slower than smax_pair.cu,
    never selected automatically -- profiles/r2_kernel_experiments.txt section 8): every (fold, column-mask) variant gives the oracle's
    records, including sets whose last 512-row unit leaves the peer CTA nothing but padding."""
    q, t = synth.make_pair(nq, nt, seed=nq + nt, dup=0.02)
    want = api.O.generate_matches(api.O.make_params(out_kind=T.OUT_DMATCH, abs_threshold=thr), q, t)
    for fold, cm in [(None, None), ("0", None), (None, "0"), ("0", "0"), (None, "1")]:
        knobs(fold=fold, colmask=cm, pair="2")
        got = api.MatchFactory(1.0, thr).generateDistanceMatches(api.Image(0), _unity(api, q), api.Image(1), _unity(api, t)).to_numpy()
        assert len(got) == len(want) and T.fields_equal(got, want), (fold, cm)


@pytest.mark.parametrize("nq,ns", [(8192 + 77, 9600 + 5), (16384, 8192)])
def test_seed_pass_distances_only_variant(api, knobs, nq, ns):
    """Large seed passes take smax_pair.cu's distances-only variant (the epilogue maximises 2S - parity(|b|^2) branch-free: no window
    records, no resolvers): the same float distances as the ordinary search of the same library and as the oracle (ref:
    src/MatchFactory.cu:349-394), duplicates of a query among the seed features (distance 0) and odd / even norms included."""
    q, s = synth.make_pair(nq, ns, seed=nq ^ ns, dup=0.01)
    s["values"][::7] ^= 1                                              # flip the parity of every seventh seed feature's norm
    want = api.O.seed_distances(q, s)
    qu = _unity(api, q); qu.setMemoryState(api.gpu)
    got = {}
    for so in ("1", "0"):
        knobs(seedonly=so)
        mf = api.MatchFactory(0.6, 200.0 * 200.0)
        mf.setSeedFeatures(_unity(api, s))
        got[so] = mf.getSeedDistances(qu).to_numpy()
        assert np.array_equal(got[so].view(np.uint32), want.view(np.uint32)), so
    knobs(seedonly=None)


@pytest.mark.parametrize("ratio,thr", [(0.8, 1.5e5), (1.0, FLT_MAX)])
def test_seeded_masked_searches_lean_instantiation(api, everest, knobs, ratio, thr):
    """Seeded masked searches run smax_tc.cu's lean instantiation (the unmasked hot loop; mask state and target locations read from
    global memory in the bookkeeping path): same records as the ordinary masked instantiation and as the oracle -- tight seed
    distances (bookkeeping rare) and a loose ratio (bookkeeping frequent), duplicates and a ragged last tile included."""
    q, t = synth.make_pair(3000 + 7, 5000 + 13, seed=83, dup=0.02)
    sd = api.O.seed_distances(q, synth.make_seed(700))
    su = api.Unity(sd, np.dtype("<f4"), api.cpu)
    qi, ti = api.Image(0, everest.cameras[0]), api.Image(1, everest.cameras[1])
    F = np.array([[1e-6, 2e-6, 0.0015], [-3e-6, 1e-7, -1.0], [-0.0013, 1.0, 4.0]], np.float32)
    p = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, has_seed=True, abs_threshold=thr, rel_threshold=ratio, query_id=0, target_id=1,
                          epsilon=25, delta=5, query_camera=everest.cameras[0], target_camera=everest.cameras[1])
    want = api.O.generate_matches(p, q, t, sd)
    pf = api.O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_FUNDAMENTAL, has_seed=True, abs_threshold=thr, rel_threshold=ratio, query_id=0,
                           target_id=1, epsilon=30.0, fundamental=F)
    want_f = api.O.generate_matches(pf, q, t, sd)
    assert len(want) > 10 and len(want_f) > 10
    for lean in ("1", "0"):
        knobs(lean=lean)
        mf = api.MatchFactory(ratio, thr)
        got = mf.generateDistanceMatchesDoubleConstrained(qi, _unity(api, q), ti, _unity(api, t), 25.0, 5.0, su).to_numpy()
        assert T.fields_equal(got, want), lean
        got = mf.generateDistanceMatchesConstrained(qi, _unity(api, q), ti, _unity(api, t), 30.0, F, su).to_numpy()
        assert T.fields_equal(got, want_f), lean
    knobs(lean=None)


def _lib_mod():
    from ssrlcv_b200 import _lib
    return _lib
