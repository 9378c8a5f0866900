"""N-view match fusion: index-only pair lists -> MultiMatch / KeyPoint sets.

Host-side step that follows the pair loop in generateMatchesExhaustive
(ref: src/MatchFactory.cu:942-1020): build per-feature adjacency lists, keep a track only if the
adjacency of each hop is consistent with the next one (std::set_intersection test), then flatten.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import types as T


def fuse_pairs(features: List[np.ndarray], pair_lists: List[np.ndarray]):
    n = len(features)
    adj = [dict() for _ in range(max(n - 1, 0))]          # adj[i][f] -> list of (view, index), insertion order
    for pl in pair_lists:
        if len(pl) == 0:
            continue
        a, b = pl["a"], pl["b"]
        for k in range(len(pl)):
            adj[int(a[k, 0])].setdefault(int(a[k, 1]), []).append((int(b[k, 0]), int(b[k, 1])))
    tracks = []
    for i in range(n - 1):
        if not i < n - 2:
            break
        ai = adj[i]
        for f in range(len(features[i])):
            a = ai.get(f)
            if not a:
                continue
            bad = False
            prev = a
            while True:
                if prev[0][0] == n - 1:
                    break
                nxt = adj[prev[0][0]].get(prev[0][1])
                if not nxt:
                    break
                # both lists are ascending by construction (pairs are visited in ascending target view)
                if len(set(prev) & set(nxt)) != len(nxt):
                    bad = True
                    break
                if len(nxt) == 1:
                    break
                prev = nxt
            if bad:
                ai[f] = []
                continue
            tracks.append([(i, f)] + list(a))
            for m in a[:-1]:
                if m[0] == n - 1:
                    break
                adj[m[0]][m[1]] = []
    mm = np.zeros(len(tracks), T.MULTIMATCH)
    nk = sum(len(t) for t in tracks)
    kp = np.zeros(nk, T.KEYPOINT)
    idx = 0
    for j, t in enumerate(tracks):
        mm[j] = (len(t), idx)
        for v, f in t:
            kp[idx]["parentId"] = v
            kp[idx]["loc"] = features[v][f]["loc"]
            idx += 1
    return mm, kp
