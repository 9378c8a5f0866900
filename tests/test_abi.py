"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
record layouts match the reference's, and the library refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ssrlcv_b200 import _lib, checkpoint as ck, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "ssrlcv_b200.h")).read()
    declared = set(re.findall(r"\b(ssrlcv_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ssrlcv_b200_abi_version() == 2


def test_params_struct_matches_header():
    # 2 i32 + 2 f32 + 2 u32 + 2 f32 + 9 f32 + 80 + 80 + 2 i32
    assert C.sizeof(_lib.MatchParams) == 8 * 4 + 36 + 160 + 8


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ssrlcv_b200_create(C.byref(h), 0, None) == -2        # SSRLCV_B200_ENODEVICE
    from ssrlcv_b200.match_factory import MatchFactory
    with pytest.raises(Exception):
        MatchFactory()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ssrlcv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), os.path.join(dirpath, f)


def test_record_layouts():
    assert T.FEATURE.itemsize == 152 and T.FEATURE.fields["values"][1] == 24
    assert T.CAMERA.itemsize == 80 and T.CAMERA.fields["size"][1] == 72
    assert (T.KEYPOINT.itemsize, T.MATCH.itemsize, T.DMATCH.itemsize, T.UINT2_PAIR.itemsize, T.MULTIMATCH.itemsize) == (16, 40, 48, 16, 8)
    assert T.FEATUREMATCH.itemsize == 320 and T.FEATUREMATCH.fields["descriptors"][1] == 44
    assert T.DMATCH.fields["distance"][1] == 40


def test_uty_roundtrip_and_hash(tmp_path, everest):
    # the hash is libstdc++'s typeid(T).hash_code(); checked against the value stored in the reference's own fixtures
    assert ck.libstdcxx_hash_bytes(b"f") != 0
    p = ck.write_uty(str(tmp_path) + "/", "keypoint", everest.g2_keypoints[:100], ck.STATE_CPU, id=7)
    assert os.path.basename(p) == "7_N6ssrlcv8KeyPointE.uty"
    # Unity<T>::checkpoint concatenates dirPath and the file name (ref: include/Unity.cuh:933): no slash, no directory
    p2 = ck.write_uty(str(tmp_path) + "/kp", "keypoint", everest.g2_keypoints[:3], ck.STATE_CPU, id=0)
    assert os.path.basename(p2) == "kp0_N6ssrlcv8KeyPointE.uty"
    back = ck.read_uty(p, "keypoint")
    assert back.state == ck.STATE_CPU and T.fields_equal(back.data, everest.g2_keypoints[:100])
    with pytest.raises(ValueError):
        ck.read_uty(p, "feature")


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/checkpoints"), reason="reference checkout not present")
def test_hash_matches_reference_fixture_header():
    d = "/root/reference/test/checkpoints/Pipeline2View/"
    for kind in ("feature", "keypoint", "multimatch"):
        name = ck.TYPE_NAMES[kind]
        fn = [f for f in os.listdir(d) if f.endswith(name + ".uty")][0]
        raw = open(d + fn, "rb").read(len(name) + 1 + 8)
        stored = int.from_bytes(raw[len(name) + 1:], "little")
        assert stored == ck.libstdcxx_hash_bytes(name.encode()), kind


def test_cpimg_roundtrip(tmp_path, everest):
    m = ck.ImageMeta(3, (1024, 1024), 1, everest.cameras[1], bytes(range(72)), 1)
    ck.write_cpimg(str(tmp_path / "3.cpimg"), m)
    b = ck.read_cpimg(str(tmp_path / "3.cpimg"))
    assert b.id == 3 and b.size == (1024, 1024) and b.camera.tobytes() == everest.cameras[1].tobytes()
    assert b.pushbroom == bytes(range(72)) and b.isPushbroom == 1          # read-then-write is lossless


def _write_stage_inputs(d, ev, n_views):
    """a feature-generation checkpoint directory like the reference's fixtures (test/checkpoints/Pipeline{2,3}View)"""
    views, seed = (ev.views2, ev.seed2) if n_views == 2 else (ev.views3, ev.seed3)
    ck.write_uty(d + "/", "feature", seed, ck.STATE_CPU, id=-1)
    for i in range(n_views):
        ck.write_uty(d + "/", "feature", views[i], ck.STATE_CPU, id=i)
        ck.write_cpimg(ck.image_checkpoint_path(d, i), ck.ImageMeta(ev.img_ids[i], (1024, 1024), 1, ev.cameras[i]))


def test_stage_inputs_from_checkpoint(tmp_path, everest):
    """FeatureMatchingInput::fromCheckpoint (ref: src/Pipeline.cu:144-163): seed image optional, one .cpimg + feature file per
    image, the pose stage's seed distances only for 2 views and only if present."""
    from ssrlcv_b200.pipeline import FeatureMatchingInput
    d = str(tmp_path / "fg"); os.makedirs(d)
    _write_stage_inputs(d, everest, 2)
    inp = FeatureMatchingInput.fromCheckpoint(d, None, 2, 25.0, 5.0)
    assert inp.seedDistances is None and len(inp.images) == 2 and [im.id for im in inp.images] == everest.img_ids[:2]
    assert T.fields_equal(inp.seedFeatures.to_numpy(), everest.seed2)
    for i in range(2):
        assert T.fields_equal(inp.allFeatures[i].to_numpy(), everest.views2[i])
        assert T.fields_equal(np.asarray(inp.images[i].camera).reshape(1), np.asarray(everest.cameras[i]).reshape(1))
    pose = str(tmp_path / "pose"); os.makedirs(pose)
    sd = np.arange(len(everest.views2[0]), dtype=np.float32)
    ck.write_uty(pose + "/", "float", sd, ck.STATE_CPU, id=0)
    inp = FeatureMatchingInput.fromCheckpoint(d, pose, 2, 25.0, 5.0)
    assert np.array_equal(inp.seedDistances.to_numpy(), sd)
    os.remove(ck.checkpoint_path(d, "feature", -1))
    assert FeatureMatchingInput.fromCheckpoint(d, None, 2, 25.0, 5.0).seedFeatures is None


@pytest.mark.skipif(not os.path.isdir("/root/reference/test/checkpoints"), reason="reference checkout not present")
def test_checkpoint_writer_reproduces_reference_files(tmp_path, everest):
    """the stage's output files as the reference wrote them (test/checkpoints/Pipeline2View/0_...KeyPointE.uty, ...MultiMatchE.uty):
    same header bytes, same records (KeyPoint has 4 padding bytes the reference leaves uninitialised: compared field-wise)."""
    d = "/root/reference/test/checkpoints/Pipeline2View/"
    for kind, data in (("keypoint", everest.g2_keypoints), ("multimatch", everest.g2_multimatch)):
        ours = ck.write_uty(str(tmp_path) + "/", kind, data, ck.STATE_CPU, id=0)
        ref_raw, our_raw = open(d + os.path.basename(ours), "rb").read(), open(ours, "rb").read()
        hdr = len(ck.TYPE_NAMES[kind]) + 1 + 9 + 13
        assert len(ref_raw) == len(our_raw) and ref_raw[:hdr] == our_raw[:hdr]
        assert T.fields_equal(ck.read_uty(d + os.path.basename(ours), kind).data, ck.read_uty(ours, kind).data)
        if kind == "multimatch":
            assert ref_raw == our_raw                       # no padding in MultiMatch: byte for byte


def test_pair_order_and_assignment():
    from ssrlcv_b200.pair_scheduler import assign_pairs, pair_order

    assert pair_order(4) == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    pairs = pair_order(16)
    assert len(pairs) == 120
    parts = assign_pairs(pairs, [32768] * 16, 8)
    assert sorted(sum(parts, [])) == list(range(120)) and all(len(p) == 15 for p in parts)
    parts = assign_pairs(pair_order(3), [10, 1000, 10], 2)
    assert sorted(sum(parts, [])) == [0, 1, 2]


def test_fusion_matches_oracle_on_synthetic():
    from oracle import oracle as O
    from ssrlcv_b200 import fusion
    from tests import synth

    feats = [synth.make_pair(300, 300, seed=s)[0] for s in range(4)]
    rng = np.random.default_rng(0)
    pairs = []
    for q in range(3):
        for t in range(q + 1, 4):
            n = 120
            qi = np.sort(rng.choice(300, n, replace=False))
            p = np.zeros(n, T.UINT2_PAIR)
            p["a"][:, 0], p["a"][:, 1], p["b"][:, 0], p["b"][:, 1] = q, qi, t, rng.integers(0, 300, n)
            pairs.append(p)
    mm, kp = fusion.fuse_pairs(feats, pairs)
    mm2, kp2 = O.fuse_nview(feats, pairs)
    assert T.fields_equal(mm, mm2) and T.fields_equal(kp, kp2)


def test_header_is_plain_c_and_links(tmp_path):
    """the boundary is a C ABI: include/ssrlcv_b200.h compiles as C99 (pedantic) and as C++11, and a C program links against
    the library and calls it (no CUDA or torch types in any signature)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "ssrlcv_b200.h"\n'
                   "int main(void) {\n"
                   "  ssrlcv_b200_match_params p; ssrlcv_b200_ctx* ctx = 0; (void)p;\n"
                   "  if (ssrlcv_b200_abi_version() != SSRLCV_B200_ABI_VERSION) return 1;\n"
                   "  if (ssrlcv_b200_create(0, 0, 0) == 0) return 2;          /* NULL out pointer: rejected */\n"
                   "  ssrlcv_b200_destroy(ctx);                                 /* NULL: no-op */\n"
                   "  return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "ssrlcv_b200")
    _lib.load()                                                       # builds the library if it is stale
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-c", str(src), "-o", str(tmp_path / "abi.o")])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-c", str(src), "-o", str(tmp_path / "abi_cxx.o")])
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", str(tmp_path / "abi.o"), "-L", libdir, "-lssrlcv_match_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_bench_gives_every_local_rank_its_own_cpus():
    """bench.py: the ranks of one host split the GPU-local CPU set (this pool's eight GPUs all report CPUs 0-31) into disjoint slices."""
    import bench
    cpus = set(range(32))
    slices = [bench.cpu_slice(cpus, r, 8) for r in range(8)]
    assert all(len(s) == 4 for s in slices) and set().union(*slices) == cpus
    assert all(slices[a].isdisjoint(slices[b]) for a in range(8) for b in range(a))
    assert bench.cpu_slice(cpus, 0, 1) == cpus
    assert bench.cpu_slice(set(range(6)), 3, 8) == set(range(6))          # fewer than two CPUs per rank: no slicing
    assert bench.cpu_slice({3, 9, 4, 8}, 1, 2) == {8, 9}
