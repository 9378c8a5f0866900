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


class PairListExchange:
    """The one exchange of pair-sharded N-view matching: every rank ends up with every pair's uint2_pair list, addressable
    in the reference's pair order (ref: src/MatchFactory.cu:925-936, :1025).

    Everything stays on the device (NCCL over NVLink; gloo on CPU for the tests): the ranks' record counts travel first
    (one small all-gather, the only host synchronisation of the exchange), then the lists, padded to the LONGEST list rather
    than to the number of queries.  Buffers are allocated once per job shape."""

    def __init__(self, n_pairs: int, assign: Sequence[Sequence[int]], capacity: int, rank: int, world_size: int, dist, device="cpu"):
        import torch

        self.torch, self.dist, self.rank, self.world, self.device = torch, dist, rank, world_size, device
        self.n_pairs, self.capacity = n_pairs, capacity
        self.slots = max(1, max((len(a) for a in assign), default=1))
        self.where = [(-1, -1)] * n_pairs                       # pair -> (rank, slot)
        for r, a in enumerate(assign):
            for s_i, k in enumerate(a):
                self.where[k] = (r, s_i)
        assert all(r >= 0 for r, _ in self.where), "every pair must be assigned to a rank"
        self._rank_of = np.array([r for r, _ in self.where], np.int64)
        self._slot_of = np.array([s_i for _, s_i in self.where], np.int64)
        pinned = str(device).startswith("cuda")
        self.counts_h = torch.zeros(self.slots, dtype=torch.int64, pin_memory=pinned)
        self.counts_d = torch.zeros(self.slots, dtype=torch.int64, device=device)
        self.all_counts_d = torch.zeros((world_size, self.slots), dtype=torch.int64, device=device)
        self.all_counts_h = torch.zeros((world_size, self.slots), dtype=torch.int64, pin_memory=pinned)
        self.gathered = torch.empty(world_size * self.slots * max(capacity, 1) * 4, dtype=torch.int32, device=device) if world_size > 1 else None
        self.last_bytes = 0

    def run(self, out, counts):
        """out: int32[slots, capacity, 4] records of this rank's pairs (slot order = its assignment order);
        counts: record count per owned pair (host).  Returns (counts per pair [n_pairs] on the host, PairLists)."""
        torch = self.torch
        m = len(counts)
        self.counts_h.zero_()
        if m:
            self.counts_h[:m] = torch.as_tensor(np.asarray(counts, dtype=np.int64))
        if self.world == 1:
            allc = self.counts_h.reshape(1, -1)
            view = out.reshape(1, self.slots, -1, 4) if out.numel() else out
            cap = out.shape[1] if out.dim() == 3 else 0
            self.last_bytes = 0
        else:
            self.counts_d.copy_(self.counts_h, non_blocking=True)
            self.dist.all_gather_into_tensor(self.all_counts_d.reshape(-1), self.counts_d)
            self.all_counts_h.copy_(self.all_counts_d, non_blocking=True)
            if str(self.device).startswith("cuda"):
                torch.cuda.current_stream().synchronize()
            allc = self.all_counts_h
            cap = max(int(allc.max()), 1)                        # longest list of the job: the lists travel padded to it
            send = out[:, :cap].contiguous()
            recv = self.gathered[: self.world * self.slots * cap * 4]
            self.dist.all_gather_into_tensor(recv, send.reshape(-1))
            view = recv.view(self.world, self.slots, cap, 4)
            self.last_bytes = recv.numel() * 4 + self.all_counts_d.numel() * 8
        ac = allc.numpy()
        per_pair = ac[self._rank_of, self._slot_of]               # counts in pair order (one gather, no per-pair Python work)
        return per_pair, PairLists(view, per_pair, self.where)


class PairLists:
    """every pair's records after the exchange, addressable in pair order: lists[k] is a device tensor int32[count_k, 4]"""

    def __init__(self, view, counts, where):
        self.view, self.counts, self.where = view, counts, where

    def __len__(self):
        return len(self.where)

    def __getitem__(self, k):
        r, s_i = self.where[k]
        return self.view[r, s_i, : int(self.counts[k])]


def exchange_pair_lists(n_pairs, assign, capacity, rank, world_size, dist, device="cpu") -> PairListExchange:
    return PairListExchange(n_pairs, assign, capacity, rank, world_size, dist, device)


def gather_pair_lists(local: dict, n_pairs: int, rank: int, world_size: int, dist, device="cpu", assign=None) -> List[np.ndarray]:
    """local: {pair_index: UINT2_PAIR array} computed by this rank.  Returns all pair lists in pair order on every rank
    (host arrays; the exchange itself runs on `device` through PairListExchange)."""
    import torch

    if assign is None:                                          # who owns what: one small all-gather of the owners' index lists
        mine = sorted(local)
        if world_size > 1:
            idx = torch.full((n_pairs,), -1, dtype=torch.int64, device=device)
            idx[: len(mine)] = torch.as_tensor(mine, dtype=torch.int64)
            alli = torch.empty((world_size, n_pairs), dtype=torch.int64, device=device)
            parts = [alli[r] for r in range(world_size)]
            dist.all_gather(parts, idx)
            assign = [[int(k) for k in alli[r].cpu().tolist() if k >= 0] for r in range(world_size)]
        else:
            assign = [mine]
    mine = list(assign[rank])
    cap = max([len(local[k]) for k in mine], default=0)
    if world_size > 1:
        capt = torch.tensor([cap], dtype=torch.int64, device=device)
        dist.all_reduce(capt, op=dist.ReduceOp.MAX)
        cap = int(capt.item())
    xch = PairListExchange(n_pairs, assign, max(cap, 1), rank, world_size, dist, device)
    out = torch.zeros((xch.slots, max(cap, 1), 4), dtype=torch.int32, device=device)
    for s_i, k in enumerate(mine):
        arr = np.ascontiguousarray(local[k], T.UINT2_PAIR)
        if len(arr):
            out[s_i, : len(arr)] = torch.from_numpy(arr.view(np.int32).reshape(-1, 4)).to(device)
    per_pair, lists = xch.run(out, [len(local[k]) for k in mine])
    res: List[np.ndarray] = []
    for k in range(n_pairs):
        if per_pair[k] == 0:
            res.append(np.zeros(0, T.UINT2_PAIR))
        else:
            res.append(np.ascontiguousarray(lists[k].cpu().numpy()).view(np.uint32).reshape(-1).view(T.UINT2_PAIR).copy())
    return res


def exhaustive_pairs_sharded(match_pair: Callable[[int, int], np.ndarray], sizes: Sequence[int], rank: int, world_size: int, dist,
                             ordered: bool = True, estimatedOverlap: float = 0.0, device="cpu") -> List[np.ndarray]:
    """Pair-sharded all-pairs matching.  match_pair(q, t) -> UINT2_PAIR array for views (q, t), run on this rank's GPU."""
    pairs = pair_order(len(sizes), ordered, estimatedOverlap)
    mine = assign_pairs(pairs, sizes, world_size)[rank]
    assign = assign_pairs(pairs, sizes, world_size)
    local = {k: match_pair(*pairs[k]) for k in mine}
    return gather_pair_lists(local, len(pairs), rank, world_size, dist, device, assign=assign)


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
