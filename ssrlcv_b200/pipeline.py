"""The matching stage of the reference's pipeline, `ssrlcv::doFeatureMatching` (ref: src/Pipeline.cu:173-241), over the
B200 matcher.  Same inputs and outputs (FeatureMatchingInput / FeatureMatchingOutput, include/Pipeline.cuh); the 2-view
tail (:198-222 -- D2H of the DMatch array, host scan for the largest distance, getRawMatches, D2H, host loop writing
KeyPoint[2n] / MultiMatch[n]) is one device pass, `ssrlcv_b200_match_set_2view`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

from . import types as T
from .match_factory import Image, MatchFactory, MatchSet
from .unity import Unity, cpu, gpu


@dataclass
class FeatureMatchingInput:
    """ref: include/Pipeline.cuh (FeatureMatchingInput): images, allFeatures, seedFeatures, seedDistances, epsilon, delta"""
    images: List[Image]
    allFeatures: List[Unity]
    seedFeatures: Optional[Unity] = None
    seedDistances: Optional[Unity] = None
    epsilon: float = 0.0
    delta: float = 0.0


@dataclass
class FeatureMatchingOutput:
    matchSet: Optional[MatchSet] = None
    maxDistance: float = 0.0          # the reference only logs it ("max euclidean distance between features")
    log: List[str] = field(default_factory=list)


def match_set_2view(mf: MatchFactory, distanceMatches: Unity) -> tuple:
    """(MatchSet, max distance) of a 2-view DMatch list: KeyPoint[2n], MultiMatch[n] = {2, 2i} (ref: src/Pipeline.cu:211-222)."""
    import torch

    origin = distanceMatches.getMemoryState()
    if origin == cpu:
        distanceMatches.setMemoryState(gpu)
    n = distanceMatches.size()
    dev = f"cuda:{mf.device}"
    kp = torch.empty(max(2 * n, 1) * T.KEYPOINT.itemsize, dtype=torch.uint8, device=dev)
    mm = torch.empty(max(n, 1) * T.MULTIMATCH.itemsize, dtype=torch.uint8, device=dev)
    mx = C.c_float(0.0)
    mf.ctx.check(mf.ctx.lib.ssrlcv_b200_match_set_2view(mf.ctx.handle, C.c_void_p(distanceMatches.device_ptr() if n else 0), n,
                                                        C.c_void_p(kp.data_ptr()), C.c_void_p(mm.data_ptr()), C.byref(mx)))
    ms = MatchSet(keyPoints=Unity(kp[: 2 * n * T.KEYPOINT.itemsize], T.KEYPOINT, gpu, mf.device, 2 * n),
                  matches=Unity(mm[: n * T.MULTIMATCH.itemsize], T.MULTIMATCH, gpu, mf.device, n))
    ms.keyPoints.setMemoryState(cpu)          # the reference hands the MatchSet on in state cpu (:212-216, :229-230)
    ms.matches.setMemoryState(cpu)
    if origin == cpu:
        distanceMatches.setMemoryState(cpu)
    return ms, float(mx.value)


def doFeatureMatching(inp: FeatureMatchingInput, geo_orbit: bool = True, device: int = 0) -> FeatureMatchingOutput:
    """ref: src/Pipeline.cu:173-241.  MatchFactory(0.6, 200^2) as there (:175)."""
    out = FeatureMatchingOutput()
    mf = MatchFactory(0.6, 200.0 * 200.0, device=device)
    if inp.seedFeatures is not None:
        mf.setSeedFeatures(inp.seedFeatures)
    if inp.seedDistances is None:
        inp.seedDistances = mf.getSeedDistances(inp.allFeatures[0]) if inp.seedFeatures is not None else None
    if len(inp.images) == 2:
        if geo_orbit:
            dm = mf.generateDistanceMatchesDoubleConstrained(inp.images[0], inp.allFeatures[0], inp.images[1], inp.allFeatures[1],
                                                             inp.epsilon, inp.delta, inp.seedDistances)
        else:
            dm = mf.generateDistanceMatches(inp.images[0], inp.allFeatures[0], inp.images[1], inp.allFeatures[1], inp.seedDistances)
        out.matchSet, out.maxDistance = match_set_2view(mf, dm)
        out.log.append("max euclidean distance between features = %f" % out.maxDistance)
        out.log.append("Total Matches: %d" % out.matchSet.matches.size())
    else:
        out.matchSet = mf.generateMatchesExhaustive(inp.images, inp.allFeatures, inp.epsilon, inp.delta, geo_orbit=geo_orbit)
        out.matchSet.matches.setMemoryState(cpu)
        out.matchSet.keyPoints.setMemoryState(cpu)
    return out
