"""The matching stage of the reference's pipeline, `ssrlcv::doFeatureMatching` (ref: src/Pipeline.cu:173-241), over the
B200 matcher.  Same inputs and outputs (FeatureMatchingInput / FeatureMatchingOutput, include/Pipeline.cuh); the 2-view
tail (:198-222 -- D2H of the DMatch array, host scan for the largest distance, getRawMatches, D2H, host loop writing
KeyPoint[2n] / MultiMatch[n]) is one device pass, `ssrlcv_b200_match_set_2view`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import os

import numpy as np

from . import checkpoint as ck
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

    @classmethod
    def fromCheckpoint(cls, featureGenDirectory: str, poseDirectory: Optional[str], numImages: int, epsilon: float, delta: float,
                       device: int = 0) -> "FeatureMatchingInput":
        """ref: src/Pipeline.cu:144-163 -- the stage's inputs from the checkpoints of the stages before it: `<dir>/-1_<Feature>.uty`
        (seed image, optional), `<dir>/<i>_<Image>.cpimg` + `<dir>/<i>_<Feature>.uty` per image, and (2 views only) the pose
        stage's `<poseDirectory>/0_f.uty` seed distances if present."""
        inp = cls(images=[], allFeatures=[], epsilon=epsilon, delta=delta)
        seed_path = ck.checkpoint_path(featureGenDirectory, "feature", -1)
        if os.path.exists(seed_path):
            inp.seedFeatures = Unity.from_checkpoint(seed_path, device)
        for i in range(numImages):
            meta = ck.read_cpimg(ck.image_checkpoint_path(featureGenDirectory, i))
            inp.images.append(Image(i, meta.camera))           # the reference passes id = i to the Image constructor (:151) ...
            inp.images[-1].id = meta.id                        # ... which then takes the id stored in the file (src/Image.cu:52)
            inp.allFeatures.append(Unity.from_checkpoint(ck.checkpoint_path(featureGenDirectory, "feature", i), device))
        if numImages == 2 and poseDirectory:
            sd_path = ck.checkpoint_path(poseDirectory, "float", 0)
            if os.path.exists(sd_path):
                inp.seedDistances = Unity.from_checkpoint(sd_path, device)
        return inp


@dataclass
class FeatureMatchingOutput:
    matchSet: Optional[MatchSet] = None
    maxDistance: float = 0.0          # the reference only logs it ("max euclidean distance between features")
    log: List[str] = field(default_factory=list)

    def checkpoint(self, cpdir: str) -> List[str]:
        """what SFM does with the stage's output when checkpointing is on (ref: src/SFM.cu:253-259): keyPoints and matches as
        Unity checkpoints with id 0, then an empty `done` marker.  `cpdir` is used as the reference uses it -- as a plain
        string prefix (Unity<T>::checkpoint concatenates, include/Unity.cuh:933) -- so pass it with its trailing slash."""
        d = os.path.dirname(cpdir)
        if d:
            os.makedirs(d, exist_ok=True)
        paths = [self.matchSet.keyPoints.checkpoint(0, cpdir), self.matchSet.matches.checkpoint(0, cpdir)]
        open(os.path.join(d if d else ".", "done"), "w").close()
        return paths


def match_set_2view(mf: MatchFactory, distanceMatches: Unity, device_resident: bool = False) -> tuple:
    """(MatchSet, max distance) of a 2-view DMatch list: KeyPoint[2n], MultiMatch[n] = {2, 2i} (ref: src/Pipeline.cu:211-222).
    device_resident: leave both arrays in state gpu -- the consumer, PointCloudFactory::generateBundles, moves them to the GPU
    first thing (ref: src/PointCloudFactory.cu:837-838), so handing them over in state cpu (as the reference does) costs a
    D2H here and an H2D there."""
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
    if not device_resident:
        ms.keyPoints.setMemoryState(cpu)      # the reference hands the MatchSet on in state cpu (:212-216, :229-230)
        ms.matches.setMemoryState(cpu)
    if origin == cpu:
        distanceMatches.setMemoryState(cpu)
    return ms, float(mx.value)


def doFeatureMatching(inp: FeatureMatchingInput, geo_orbit: bool = True, device: int = 0, device_resident: bool = False) -> FeatureMatchingOutput:
    """ref: src/Pipeline.cu:173-241.  MatchFactory(0.6, 200^2) as there (:175).  device_resident: see match_set_2view."""
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
        out.matchSet, out.maxDistance = match_set_2view(mf, dm, device_resident)
        out.log.append("max euclidean distance between features = %f" % out.maxDistance)
        out.log.append("Total Matches: %d" % out.matchSet.matches.size())
    else:
        out.matchSet = mf.generateMatchesExhaustive(inp.images, inp.allFeatures, inp.epsilon, inp.delta, geo_orbit=geo_orbit)
        if device_resident:
            out.matchSet.matches.setMemoryState(gpu)
            out.matchSet.keyPoints.setMemoryState(gpu)
        else:
            out.matchSet.matches.setMemoryState(cpu)
            out.matchSet.keyPoints.setMemoryState(cpu)
    return out


def run_stage(featureGenDirectory: str, poseDirectory: Optional[str], outDirectory: str, numImages: int, epsilon: float, delta: float,
              geo_orbit: bool = True, device: int = 0) -> FeatureMatchingOutput:
    """The matching stage stand-alone, from a `--cpdir` tree to a `--cpdir` tree: what `SFM --cpdir ... ` does between its
    FEATURE_MATCHING label and the next stage (ref: src/SFM.cu:247-259) when it resumes from the pose-estimation checkpoints."""
    inp = FeatureMatchingInput.fromCheckpoint(featureGenDirectory, poseDirectory, numImages, epsilon, delta, device)
    out = doFeatureMatching(inp, geo_orbit=geo_orbit, device=device)
    out.checkpoint(outDirectory if outDirectory.endswith("/") else outDirectory + "/")
    return out


def main(argv=None) -> int:
    import argparse

    ap = argparse.ArgumentParser(description="SSRLCV feature-matching stage on B200, from and to Unity checkpoints")
    ap.add_argument("--fg", required=True, help="feature-generation checkpoint directory (<i>_...Feature....uty, <i>_...Image.cpimg, -1_... = seed image)")
    ap.add_argument("--pose", default=None, help="pose-estimation checkpoint directory (0_f.uty seed distances; 2 views only)")
    ap.add_argument("--out", required=True, help="directory the stage's checkpoints go to (0_...KeyPointE.uty, 0_...MultiMatchE.uty, done)")
    ap.add_argument("-n", "--num-images", type=int, required=True)
    ap.add_argument("--epsilon", type=float, default=25.0)
    ap.add_argument("--delta", type=float, default=5.0)
    ap.add_argument("--no-geo-orbit", action="store_true", help="GEO_ORBIT=0 build of the reference: unconstrained matching")
    a = ap.parse_args(argv)
    out = run_stage(a.fg, a.pose, a.out, a.num_images, a.epsilon, a.delta, not a.no_geo_orbit)
    for line in out.log:
        print(line)
    print(f"keyPoints {out.matchSet.keyPoints.size()}  matches {out.matchSet.matches.size()}  -> {a.out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
