#!/usr/bin/env python
"""Convert the reference's own matching fixtures into one committed golden file.

Source (read-only, only present in the build container):
  /root/reference/test/checkpoints/Pipeline2View  and  Pipeline3View
  (the inputs and expected outputs of PipelineTest.FeatureMatching2View/3View,
   ref: test/Pipeline.cu:183-255).

Output: tests/golden/everest.npz
  v{-1,0,1,2}_{values,loc,sigma,theta,parent}  the 3-view feature sets (-1 = seed image)
  p2_v{-1,0,1}_{idx,val}                       sparse byte overrides turning the 3-view sets into the
                                               2-view fixture's sets (the two fixture directories differ in
                                               a handful of descriptor bytes)
  cam{0,1,2}                                   Image::Camera records (80 B) ; img_ids
  g2_keypoints / g2_multimatch                 golden output of the 2-view test
  g3_keypoints / g3_multimatch                 golden output of the 3-view test

Nothing under tests/ or bench.py reads /root/reference at run time; they read this file.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssrlcv_b200 import checkpoint as ck  # noqa: E402

REF = "/root/reference/test/checkpoints"
FEAT = "N6ssrlcv7FeatureINS_15SIFT_DescriptorEEE.uty"


def main():
    d = {}
    d3, d2 = f"{REF}/Pipeline3View/", f"{REF}/Pipeline2View/"
    for i in (-1, 0, 1, 2):
        f = ck.read_uty(f"{d3}{i}_{FEAT}", "feature").data
        for n in ("values", "loc", "sigma", "theta", "parent"):
            d[f"v{i}_{n}"] = f[n]
        if i != 2:
            g = ck.read_uty(f"{d2}{i}_{FEAT}", "feature").data
            for n in ("loc", "sigma", "theta", "parent"):
                assert np.array_equal(f[n].view(np.uint32), g[n].view(np.uint32)), (i, n)
            idx = np.flatnonzero(f["values"].reshape(-1) != g["values"].reshape(-1))
            d[f"p2_v{i}_idx"] = idx.astype(np.int64)
            d[f"p2_v{i}_val"] = g["values"].reshape(-1)[idx]
    ids = []
    for i in (0, 1, 2):
        m = ck.read_cpimg(f"{d3}{i}_N6ssrlcv5ImageE.cpimg")
        d[f"cam{i}"] = np.frombuffer(np.asarray(m.camera).tobytes(), np.uint8)
        ids.append(m.id)
        if i != 2:
            m2 = ck.read_cpimg(f"{d2}{i}_N6ssrlcv5ImageE.cpimg")
            assert np.asarray(m2.camera).tobytes() == np.asarray(m.camera).tobytes() and m2.id == m.id
    d["img_ids"] = np.array(ids, np.int32)
    for tag, dd in (("g2", d2), ("g3", d3)):
        kp = ck.read_uty(f"{dd}0_N6ssrlcv8KeyPointE.uty", "keypoint").data
        mm = ck.read_uty(f"{dd}0_N6ssrlcv10MultiMatchE.uty", "multimatch").data
        d[f"{tag}_kp_parent"] = kp["parentId"]
        d[f"{tag}_kp_loc"] = kp["loc"]
        d[f"{tag}_mm_n"] = mm["numKeyPoints"]
        d[f"{tag}_mm_index"] = mm["index"]
    out = os.path.join(ROOT, "tests", "golden", "everest.npz")
    np.savez_compressed(out, **d)
    print(out, os.path.getsize(out) / 1e6, "MB")


if __name__ == "__main__":
    main()
