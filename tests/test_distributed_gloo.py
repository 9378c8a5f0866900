"""world_size-2 gloo tests of the multi-GPU host logic, run on CPU.

The orchestration in ssrlcv_b200/distributed.py takes the per-rank compute as callables; here the callables are the CPU
oracle (this is a test, the only place allowed to do that), so what is checked is the sharding itself:
  * pair sharding: the pair lists come back in the reference's pair order (ref: src/MatchFactory.cu:925-936, :1025)
    and equal the single-process result;
  * target sharding: the minimum over ranks of the 64-bit candidate keys d2<<37 | (f%32)<<32 | f (f global) is the
    reference's winner (ref: src/MatchFactory.cu:1476-1492), including ties that straddle a shard boundary.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _keys_from_oracle(q, t, offset):
    """CPU stand-in for ssrlcv_b200_match_candidates: per query the key of the reference's winner in the shard."""
    from oracle import oracle as O
    from ssrlcv_b200 import types as T

    keys = np.full(len(q), -1, np.int64)            # all-ones = no candidate
    if len(t) == 0:
        return keys
    # the lane rule depends on the GLOBAL index: shard offsets are multiples of 32, so f % 32 is shard-invariant
    assert offset % 32 == 0
    bd, bi, _ = O.match_winners(O.make_params(out_kind=T.OUT_PAIR), q, t, None)
    for i in range(len(q)):
        if bi[i] < 0:
            continue
        g = int(bi[i]) + offset
        keys[i] = (int(bd[i]) << 37) | ((g & 31) << 32) | g
    return keys


def _worker_target_sharded(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ssrlcv_b200 import distributed as D
        from tests import synth

        q, t = synth.make_pair(96, 700, seed=5)
        # a tie that straddles the shard boundary: the same descriptor at a low-lane/high-index slot in shard 1
        # and a high-lane/low-index slot in shard 0 -> the lane rule must pick the shard-1 copy
        t["values"][31] = q["values"][7]
        t["values"][512 + 1] = q["values"][7]

        def candidates(b, e, off):
            return torch.from_numpy(_keys_from_oracle(q, t[b:e], off))

        def finalize(allk, n):
            return allk.numpy().view(np.uint64).min(axis=0)

        got = D.match_target_sharded(candidates, finalize, len(q), len(t), rank, world, dist)
        want = _keys_from_oracle(q, t, 0).view(np.uint64)
        ok = np.array_equal(got, want) and int(want[7] & 0xFFFFFFFF) == 513
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _worker_pair_sharded(rank, world, port, ret):
    import torch  # noqa: F401
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from ssrlcv_b200 import distributed as D
        from ssrlcv_b200 import types as T
        from tests import synth

        sizes = [150, 90, 200, 60]
        feats = [synth.make_pair(n, 8, seed=40 + v)[0] for v, n in enumerate(sizes)]
        calls = []

        def match_pair(qv, tv):
            calls.append((qv, tv))
            p = O.make_params(out_kind=T.OUT_PAIR, abs_threshold=60000.0, query_id=qv, target_id=tv)
            return O.generate_matches(p, feats[qv], feats[tv], None)

        got = D.exhaustive_pairs_sharded(match_pair, sizes, rank, world, dist)
        from ssrlcv_b200.pair_scheduler import assign_pairs, pair_order
        pairs = pair_order(len(sizes))
        mine = assign_pairs(pairs, sizes, world)[rank]
        ok = sorted(calls) == sorted(pairs[k] for k in mine)           # every rank ran only its own pairs
        for k, (qv, tv) in enumerate(pairs):
            p = O.make_params(out_kind=T.OUT_PAIR, abs_threshold=60000.0, query_id=qv, target_id=tv)
            want = O.generate_matches(p, feats[qv], feats[tv], None)
            ok = ok and T.fields_equal(got[k], want)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    import torch.multiprocessing as mp

    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(fn, args=(world, port, ret), nprocs=world, join=True)
        return dict(ret)


@pytest.mark.timeout(300)
def test_target_sharding_merges_to_reference_winner_gloo():
    ret = _spawn(_worker_target_sharded)
    assert ret == {0: True, 1: True}


@pytest.mark.timeout(300)
def test_pair_sharding_keeps_reference_pair_order_gloo():
    ret = _spawn(_worker_pair_sharded)
    assert ret == {0: True, 1: True}


def test_shard_bounds_are_lane_aligned():
    from ssrlcv_b200.distributed import shard_bounds

    for n, w in [(1 << 20, 8), (1000, 2), (31, 4), (65536 + 17, 8), (0, 2)]:
        b = shard_bounds(n, w)
        assert len(b) == w + 1 and b[0] == 0 and b[-1] == n
        assert all(b[i] <= b[i + 1] for i in range(w))
        # every non-empty shard starts on a lane boundary (f % 32 must be shard-invariant)
        assert all(b[i] % 32 == 0 for i in range(w) if b[i + 1] > b[i])


def test_assign_pairs_is_balanced_and_keeps_views_local():
    """pair_scheduler.assign_pairs: every pair exactly once, loads balanced by cost, and a rank touches few views
    (a rank uploads and packs every view it touches: 16 equal views over 8 ranks must not need all 16 per rank)."""
    from ssrlcv_b200.pair_scheduler import assign_pairs, pair_order
    pairs = pair_order(16)
    for world in (1, 2, 3, 4, 8):
        a = assign_pairs(pairs, [32768] * 16, world)
        assert sorted(k for r in a for k in r) == list(range(len(pairs)))
        sizes = [len(r) for r in a]
        assert max(sizes) - min(sizes) <= 1
        assert a == assign_pairs(pairs, [32768] * 16, world)            # deterministic
    views8 = [len({v for k in r for v in pairs[k]}) for r in assign_pairs(pairs, [32768] * 16, 8)]
    assert max(views8) <= 10, views8
    # unequal sizes: longest-processing-time balance within a few percent
    rng = np.random.default_rng(3)
    sizes = [int(x) for x in rng.integers(2000, 60000, 11)]
    pairs = pair_order(11)
    a = assign_pairs(pairs, sizes, 4)
    load = [sum(sizes[pairs[k][0]] * sizes[pairs[k][1]] for k in r) for r in a]
    assert max(load) / min(load) < 1.10, load
    # the reference's skip rule for ordered image sets survives the assignment (pairs are indices into pair_order)
    skipped = pair_order(6, ordered=True, estimatedOverlap=0.5)
    a = assign_pairs(skipped, [1000] * 6, 2)
    assert sorted(k for r in a for k in r) == list(range(len(skipped)))
