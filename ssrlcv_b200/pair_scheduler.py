"""Pair scheduler for N-view all-pairs matching.

Replaces the serial pair loop of generateMatchesExhaustive (ref: src/MatchFactory.cu:907-936):
directed pairs (i < j) in the reference's order 0:1, 0:2, ..., 1:2, ...; seed distances are computed once
per query view; packed descriptor sets are cached per view so every view is repacked once, not once per pair.
Across ranks the pairs are independent work units: `assign_pairs` deals them out by cost
Nq * Nt (longest-processing-time first), and results are re-ordered into the reference's pair order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import types as T


def pair_order(n_views: int, ordered: bool = True, estimatedOverlap: float = 0.0) -> List[Tuple[int, int]]:
    """The (query, target) pairs the reference visits, including its skip rule for ordered image sets
    (ref: src/MatchFactory.cu:925-934; the counter `i` there is shared between the skip test and the push)."""
    pairs = []
    i = 0
    for q in range(n_views - 1):
        for t in range(q + 1, n_views):
            if ordered and estimatedOverlap > 0.0:
                i += 1
                if i * (1 - estimatedOverlap) > 1.0:
                    continue
            pairs.append((q, t))
            i += 1
    return pairs


def assign_pairs(pairs: Sequence[Tuple[int, int]], sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """LPT assignment of pair indices to ranks by cost Nq*Nt; deterministic."""
    cost = [sizes[q] * sizes[t] for q, t in pairs]
    order = sorted(range(len(pairs)), key=lambda k: (-cost[k], k))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for k in order:
        r = min(range(world_size), key=lambda x: (load[x], x))
        out[r].append(k)
        load[r] += cost[k]
    for r in range(world_size):
        out[r].sort()
    return out


def exhaustive_pairs(mf, images, features, epsilon, delta, ordered=True, estimatedOverlap=0.0, geo_orbit=True,
                     only: Sequence[int] | None = None):
    """Run the pair loop on this process' GPU. `only`: indices into pair_order() to run (multi-GPU sharding)."""
    pairs = pair_order(len(images), ordered, estimatedOverlap)
    todo = list(range(len(pairs)) if only is None else only)
    seed_cache = {}
    jobs = []
    for k in todo:
        q, t = pairs[k]
        seed = None
        if mf.seedFeatures is not None:
            if q not in seed_cache:
                seed_cache[q] = mf.getSeedDistances(features[q])      # once per query view, not once per pair
            seed = seed_cache[q]
        if geo_orbit:
            p = mf._params(T.OUT_PAIR, T.MASK_DOUBLE, images[q], images[t], epsilon, delta)
        else:
            p = mf._params(T.OUT_PAIR, T.MASK_NONE, images[q], images[t])
        jobs.append((p, features[q], features[t], seed))
    # every pair is enqueued back to back on the device and synchronised once (ssrlcv_b200_match_batch)
    return mf.matchBatch(jobs)
