"""Pins the CPU oracle to the reference's own golden vectors.

ref: test/Pipeline.cu:183-216 (FeatureMatching2View) and :218-255 (FeatureMatching3View):
doFeatureMatching (src/Pipeline.cu:173-241) with MatchFactory(0.6, 200*200), epsilon 25, delta 5.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from ssrlcv_b200 import checkpoint as ck
from ssrlcv_b200 import types as T


def _two_view(ev, trig_mode):
    q, t = ev.views2
    sd = O.seed_distances(q, ev.seed2)
    p = O.make_params(out_kind=T.OUT_DMATCH, mask_kind=T.MASK_DOUBLE, has_seed=True,
                      abs_threshold=ev.abs_threshold, rel_threshold=ev.rel_threshold,
                      query_id=ev.img_ids[0], target_id=ev.img_ids[1], epsilon=ev.epsilon, delta=ev.delta,
                      query_camera=ev.cameras[0], target_camera=ev.cameras[1], trig_mode=trig_mode)
    return O.generate_matches(p, q, t, sd)


@pytest.mark.parametrize("trig_mode", [0, 1])
def test_two_view_golden(everest, trig_mode):
    m = _two_view(everest, trig_mode)
    assert len(m) == 13534
    # src/Pipeline.cu:206-222: getRawMatches, then KeyPoint[2i], KeyPoint[2i+1], MultiMatch{2, 2i}
    kp = O.raw_matches(m)["keyPoints"].reshape(-1).copy()
    assert T.fields_equal(kp, everest.g2_keypoints)
    assert np.array_equal(everest.g2_multimatch["numKeyPoints"], np.full(len(m), 2))
    assert np.array_equal(everest.g2_multimatch["index"], 2 * np.arange(len(m)))
    # distances are exact integers below the absolute threshold
    assert np.all(m["distance"] == np.floor(m["distance"])) and m["distance"].max() < everest.abs_threshold


def test_three_view_golden(everest):
    ev = everest
    pairs = []
    for qi in range(2):
        sd = O.seed_distances(ev.views3[qi], ev.seed3)
        for ti in range(qi + 1, 3):
            p = O.make_params(out_kind=T.OUT_PAIR, mask_kind=T.MASK_DOUBLE, has_seed=True,
                              abs_threshold=ev.abs_threshold, rel_threshold=ev.rel_threshold,
                              query_id=ev.img_ids[qi], target_id=ev.img_ids[ti], epsilon=ev.epsilon, delta=ev.delta,
                              query_camera=ev.cameras[qi], target_camera=ev.cameras[ti])
            pairs.append(O.generate_matches(p, ev.views3[qi], ev.views3[ti], sd))
    assert [len(p) for p in pairs] == [15627, 15439, 12795]
    mm, kp = O.fuse_nview(ev.views3, pairs)
    assert len(mm) == 21177 and len(kp) == 51442
    assert T.fields_equal(mm, ev.g3_multimatch)
    assert T.fields_equal(kp, ev.g3_keypoints)


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/checkpoints"), reason="reference checkout not present")
def test_committed_golden_matches_reference_files(everest):
    """tests/golden/everest.npz is a faithful conversion of the reference's fixture files."""
    d = "/root/reference/test/checkpoints/Pipeline2View/"
    name = ck.TYPE_NAMES["feature"]
    for i, mine in ((-1, everest.seed2), (0, everest.views2[0]), (1, everest.views2[1])):
        ref = ck.read_uty(f"{d}{i}_{name}.uty", "feature").data
        assert T.fields_equal(ref, mine)
    ref_kp = ck.read_uty(d + "0_N6ssrlcv8KeyPointE.uty", "keypoint").data
    assert T.fields_equal(ref_kp, everest.g2_keypoints)


def test_tie_break_is_lane_then_index():
    """ref: src/MatchFactory.cu:1476-1492 -- winner among equal distances = min (f % 32, f), not lowest f."""
    rng = np.random.default_rng(7)
    q = T.make_features(rng.integers(0, 60, (1, 128), dtype=np.uint8))
    tv = rng.integers(100, 160, (96, 128), dtype=np.uint8)
    near = q["values"][0].copy()
    near[0] += 3
    tv[37] = near          # lane 5
    tv[66] = near          # lane 2  -> wins although 37 < 66
    tv[34 + 32] = near     # == 66
    t = T.make_features(tv)
    bd, bi, ok = O.match_winners(O.make_params(), q, t)
    assert bi[0] == 66 and bd[0] == 9.0 and ok[0] == 1
    tv[2] = near           # lane 2, lower f -> wins
    bd, bi, ok = O.match_winners(O.make_params(), q, T.make_features(tv))
    assert bi[0] == 2


def test_ratio_threshold_table():
    """rel vs rel^2 per kernel (ref: src/MatchFactory.cu:1695,1859,2112,2278,2751,2907)."""
    rel = np.float32(0.6)
    r2 = float(rel * rel)
    assert O.lib().orc_ratio_threshold(T.OUT_MATCH, T.MASK_NONE, rel) == float(rel)
    assert O.lib().orc_ratio_threshold(T.OUT_MATCH, T.MASK_FUNDAMENTAL, rel) == float(rel)
    assert O.lib().orc_ratio_threshold(T.OUT_MATCH, T.MASK_DOUBLE, rel) == r2
    for mk in (T.MASK_NONE, T.MASK_FUNDAMENTAL, T.MASK_DOUBLE):
        assert O.lib().orc_ratio_threshold(T.OUT_DMATCH, mk, rel) == r2
        assert O.lib().orc_ratio_threshold(T.OUT_PAIR, mk, rel) == float(rel)


def test_empty_and_threshold_edges():
    rng = np.random.default_rng(3)
    q = T.make_features(rng.integers(0, 255, (5, 128), dtype=np.uint8))
    empty = np.zeros(0, T.FEATURE)
    assert np.all(O.seed_distances(q, empty) == np.finfo(np.float32).max)     # ref :1439 init FLT_MAX
    bd, bi, ok = O.match_winners(O.make_params(), q, empty)
    assert np.all(bi == -1) and not ok.any()
    # a candidate exactly at the absolute threshold never wins (running best starts at abs, strict <) ref :1472,1478
    t = q.copy()
    t["values"][:, 0] = t["values"][:, 0] ^ 4                                   # d2 == 16 to its own query
    bd, bi, ok = O.match_winners(O.make_params(abs_threshold=16.0), q[:1], t[:1])
    assert bi[0] == -1 and not ok[0] and bd[0] == 16.0
    bd, bi, ok = O.match_winners(O.make_params(abs_threshold=17.0), q[:1], t[:1])
    assert bi[0] == 0 and ok[0] and bd[0] == 16.0
