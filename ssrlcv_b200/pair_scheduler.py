"""Pair scheduler for N-view all-pairs matching.

Replaces the serial pair loop of generateMatchesExhaustive (ref: src/MatchFactory.cu:907-936):
directed pairs (i < j) in the reference's order 0:1, 0:2, ..., 1:2, ...; seed distances are computed once
per query view; packed descriptor sets are cached per view so every view is repacked once, not once per pair.
Across ranks the pairs are independent work units: `assign_pairs` deals them out balanced by cost
Nq * Nt with view affinity (a rank packs only the views it touches), and results are re-ordered into the reference's
pair order.
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
    """Deterministic assignment of pair indices to ranks: balanced by cost Nq*Nt, with view affinity.

    The least-loaded rank picks next (longest-processing-time order: only pairs within a factor two of the largest
    remaining cost are eligible), and among the eligible pairs it takes the one that brings the fewest views it does not
    hold yet, then the one whose new views the fewest other ranks hold.  A rank uploads and packs every view it touches,
    so cliques beat stars: 16 equal views over 8 ranks need 7-9 views per rank instead of all 16 (plain LPT), with the
    same 15 pairs each."""
    cost = [sizes[q] * sizes[t] for q, t in pairs]
    left = set(range(len(pairs)))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    held: List[set] = [set() for _ in range(world_size)]
    holders = [0] * len(sizes)                      # ranks holding each view
    while left:
        r = min(range(world_size), key=lambda x: (load[x], x))
        cmax = max(cost[k] for k in left)
        best_key, best = None, -1
        for k in left:
            if cost[k] * 2 < cmax:
                continue
            q, t = pairs[k]
            nq, nt = q not in held[r], t not in held[r]
            key = (int(nq) + int(nt and t != q), (holders[q] if nq else 0) + (holders[t] if nt else 0), -cost[k], k)
            if best_key is None or key < best_key:
                best_key, best = key, k
        left.discard(best)
        out[r].append(best)
        load[r] += cost[best]
        for v in set(pairs[best]):
            if v not in held[r]:
                held[r].add(v)
                holders[v] += 1
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
