"""Multi-GPU forms of the matching path (one process per GPU, torch.distributed for the plumbing).

Two shardings, as the path offers them:

* pair sharding (N-view all-pairs, `generateMatchesExhaustive`): image pairs are independent work units.  Pairs are
  dealt to ranks by cost (pair_scheduler.assign_pairs); every rank matches its pairs on its own GPU with no data-path
  collective; the variable-length index-only results are exchanged once at the end (all_gather of counts, then of
  padded pair arrays) and re-ordered into the reference's pair order (ref: src/MatchFactory.cu:925-936, :1025).

* target sharding (one huge pair): rank r owns targets [bounds[r], bounds[r+1]) (bounds multiples of 32, so f % 32 is
  shard-invariant); queries are replicated.  Each rank emits one 64-bit candidate per query
  (d2 << 37 | (f % 32) << 32 | f, f global); an all_gather over NVLink brings the G candidates per query together and the
  finalize kernel takes their minimum -- exactly the reference's winner rule -- before thresholds and compaction.

The compute callables are injected so the orchestration can be tested on CPU with the gloo backend.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np

from . import types as T
from .pair_scheduler import assign_pairs, pair_order


def shard_bounds(n_targets: int, world_size: int, align: int = 256) -> List[int]:
    """Contiguous target shards, every boundary a multiple of `align` (a multiple of 32, the reference's lane count)."""
    assert align % 32 == 0
    per = -(-n_targets // world_size)
    per = -(-per // align) * align
    b = [min(r * per, n_targets) for r in range(world_size)] + [n_targets]
    return b


def gather_pair_lists(local: dict, n_pairs: int, rank: int, world_size: int, dist, device="cpu") -> List[np.ndarray]:
    """local: {pair_index: UINT2_PAIR array} computed by this rank.  Returns all pair lists in pair order on every rank."""
    import torch

    counts = torch.zeros(n_pairs, dtype=torch.int64, device=device)
    for k, arr in local.items():
        counts[k] = len(arr)
    if world_size > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)        # every pair is owned by exactly one rank
    counts_h = counts.cpu().numpy()
    cap = int(counts_h.max()) if n_pairs else 0
    mine = sorted(local)
    per_rank = [0] * world_size
    if world_size > 1:
        n_mine = torch.tensor([len(mine)], dtype=torch.int64, device=device)
        all_n = [torch.zeros_like(n_mine) for _ in range(world_size)]
        dist.all_gather(all_n, n_mine)
        per_rank = [int(x.item()) for x in all_n]
    else:
        per_rank = [len(mine)]
    slots = max(per_rank) if per_rank else 0
    # payload per rank: [slots] pair indices + [slots, cap, 4] uint32 records
    idx = torch.full((max(slots, 1),), -1, dtype=torch.int64, device=device)
    buf = torch.zeros((max(slots, 1), max(cap, 1), 4), dtype=torch.int32, device=device)
    for s, k in enumerate(mine):
        idx[s] = k
        arr = np.ascontiguousarray(local[k], T.UINT2_PAIR)
        if len(arr):
            buf[s, : len(arr)] = torch.from_numpy(arr.view(np.int32).reshape(-1, 4)).to(device)
    if world_size > 1:
        all_idx = [torch.empty_like(idx) for _ in range(world_size)]
        all_buf = [torch.empty_like(buf) for _ in range(world_size)]
        dist.all_gather(all_idx, idx)
        dist.all_gather(all_buf, buf)
    else:
        all_idx, all_buf = [idx], [buf]
    out: List[np.ndarray] = [np.zeros(0, T.UINT2_PAIR) for _ in range(n_pairs)]
    for r in range(world_size):
        ih = all_idx[r].cpu().numpy()
        bh = all_buf[r].cpu().numpy()
        for s, k in enumerate(ih):
            if k < 0:
                continue
            n = int(counts_h[k])
            out[int(k)] = np.ascontiguousarray(bh[s, :n]).view(np.uint32).reshape(-1).view(T.UINT2_PAIR).copy()
    return out


def exhaustive_pairs_sharded(match_pair: Callable[[int, int], np.ndarray], sizes: Sequence[int], rank: int, world_size: int, dist,
                             ordered: bool = True, estimatedOverlap: float = 0.0, device="cpu") -> List[np.ndarray]:
    """Pair-sharded all-pairs matching.  match_pair(q, t) -> UINT2_PAIR array for views (q, t), run on this rank's GPU."""
    pairs = pair_order(len(sizes), ordered, estimatedOverlap)
    mine = assign_pairs(pairs, sizes, world_size)[rank]
    local = {k: match_pair(*pairs[k]) for k in mine}
    return gather_pair_lists(local, len(pairs), rank, world_size, dist, device)


def match_target_sharded(candidates: Callable[[int, int, int], "object"], finalize: Callable[["object", int], np.ndarray],
                         n_queries: int, n_targets: int, rank: int, world_size: int, dist, device="cpu"):
    """Target-sharded single pair.
    candidates(begin, end, offset) -> int64 tensor[n_queries] of candidate keys for targets [begin, end) (global index = offset + local)
    finalize(keys[world_size, n_queries], world_size) -> compacted records."""
    import torch

    b = shard_bounds(n_targets, world_size)
    # an empty trailing shard starts at n_targets, which need not be lane-aligned: its offset is irrelevant, pass 0
    mine = candidates(b[rank], b[rank + 1], b[rank] if b[rank + 1] > b[rank] else 0)
    if world_size > 1:
        allk = torch.empty((world_size, n_queries), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(allk, mine.contiguous()) if hasattr(dist, "all_gather_into_tensor") and mine.device.type == "cuda" else \
            _all_gather_rows(allk, mine, dist, world_size)
    else:
        allk = mine.reshape(1, -1)
    return finalize(allk, world_size)


def _all_gather_rows(out, mine, dist, world_size):
    import torch

    parts = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(parts, mine.contiguous())
    for r in range(world_size):
        out[r] = parts[r]
