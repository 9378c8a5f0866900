"""Deterministic synthetic SIFT-shaped inputs: descriptors post-processed like the reference's
fillDescriptors (L2-normalise, clip 0.2, renormalise, round(255 x)), planted correspondences, exact duplicates
placed so that the lane rule differs from lowest-index, locations on the query's epipolar segment for masked runs."""
import numpy as np

from ssrlcv_b200 import types as T


def sift_descriptors(rng, n):
    x = np.maximum(rng.normal(size=(n, 128)), 0) ** 2 + rng.exponential(0.05, size=(n, 128))
    x[rng.random((n, 128)) < 0.1] = 0
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-9)
    x = np.minimum(x, 0.2)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-9)
    return np.round(255 * x).astype(np.uint8)


def make_pair(nq, nt, seed=0, planted=0.4, dup=0.002, noise=2, uniform_bytes=False):
    rng = np.random.default_rng(seed)
    gen = (lambda n: rng.integers(0, 256, (n, 128), dtype=np.uint8)) if uniform_bytes else (lambda n: sift_descriptors(rng, n))
    qv, tv = gen(nq), gen(nt)
    k = int(min(nq, nt) * planted)
    if k:
        slots = rng.choice(nt, k, replace=False)
        tv[slots] = np.clip(qv[:k].astype(np.int32) + np.round(rng.normal(0, noise, (k, 128))).astype(np.int32), 0, 255).astype(np.uint8)
    nd = int(nt * dup)
    for _ in range(nd):                      # exact duplicate targets: same descriptor at a higher-lane/lower-index and lower-lane/higher-index slot
        a = int(rng.integers(0, max(nt - 64, 1)))
        b = min(nt - 1, a + int(rng.integers(1, 64)))
        tv[b] = tv[a]
    ql = rng.uniform(6, 1018, (nq, 2)).astype(np.float32)
    tl = rng.uniform(6, 1018, (nt, 2)).astype(np.float32)
    return T.make_features(qv, ql, parent=0), T.make_features(tv, tl, parent=1)


def make_seed(ns, seed=99):
    rng = np.random.default_rng(seed)
    return T.make_features(sift_descriptors(rng, ns))
