"""Reader/writer for the reference's binary checkpoints on the matching path.

`.uty`  = Unity<T>::checkpoint  (ref: include/Unity.cuh:923-971, reader :466-531)
          header: mangled type name '\\n', hash_code (8 B) '\\n', MemoryState (4 B),
          numElements (8 B) '\\n'; body: raw T[numElements].
`.cpimg` = raw dump of the 240-byte ssrlcv::Image object (ref: src/Image.cu:274-303,
          reader :48-64): id@32, size@40, colorDepth@48, camera@56 (80 B),
          pushbroom@136, isPushbroom@208.

The hash_code is libstdc++'s typeid(T).hash_code() = std::_Hash_bytes(name, len,
0xc70f6907) (murmur2-64A); it is recomputed here so files written by this
module are accepted by the reference's reader.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

from . import types as T

# MemoryState, ref: include/Unity.cuh (null=0, cpu=1, gpu=2, both=3)
STATE_NULL, STATE_CPU, STATE_GPU, STATE_BOTH = 0, 1, 2, 3

TYPE_NAMES = {
    "feature": "N6ssrlcv7FeatureINS_15SIFT_DescriptorEEE",
    "keypoint": "N6ssrlcv8KeyPointE",
    "multimatch": "N6ssrlcv10MultiMatchE",
    "match": "N6ssrlcv5MatchE",
    "dmatch": "N6ssrlcv6DMatchE",
    "pair": "N6ssrlcv10uint2_pairE",
    "float": "f",
    "uchar": "h",
}
TYPE_DTYPES = {
    "feature": T.FEATURE,
    "keypoint": T.KEYPOINT,
    "multimatch": T.MULTIMATCH,
    "match": T.MATCH,
    "dmatch": T.DMATCH,
    "pair": T.UINT2_PAIR,
    "float": np.dtype("<f4"),
    "uchar": np.dtype("u1"),
}
_NAME_TO_KIND = {v: k for k, v in TYPE_NAMES.items()}


def libstdcxx_hash_bytes(data: bytes, seed: int = 0xC70F6907) -> int:
    """std::_Hash_bytes for 64-bit size_t (MurmurHash2, 64-bit version by Austin Appleby)."""
    mask = (1 << 64) - 1
    mul = (0xC6A4A793 << 32) + 0x5BD1E995
    n = len(data)
    h = (seed ^ (n * mul)) & mask
    nblocks = n // 8
    for i in range(nblocks):
        k = struct.unpack_from("<Q", data, i * 8)[0]
        k = (k * mul) & mask
        k ^= k >> 47
        k = (k * mul) & mask
        h ^= k
        h = (h * mul) & mask
    tail = data[nblocks * 8 :]
    if tail:
        k = int.from_bytes(tail, "little")
        h ^= k
        h = (h * mul) & mask
    h ^= h >> 47
    h = (h * mul) & mask
    h ^= h >> 47
    return h


@dataclass
class UnityFile:
    kind: str
    state: int
    data: np.ndarray


def read_uty(path: str, expect: str | None = None) -> UnityFile:
    with open(path, "rb") as f:
        raw = f.read()
    nl = raw.index(b"\n")
    name = raw[:nl].decode()
    if name not in _NAME_TO_KIND:
        raise ValueError(f"{path}: unknown Unity element type {name!r}")
    kind = _NAME_TO_KIND[name]
    if expect is not None and kind != expect:
        raise ValueError(f"{path}: names of type T do not match up ({kind} vs {expect})")
    off = nl + 1
    (hash_code,) = struct.unpack_from("<Q", raw, off)
    off += 9
    if hash_code != libstdcxx_hash_bytes(name.encode()):
        raise ValueError(f"{path}: hash_codes of type T do not match up")
    state, n = struct.unpack_from("<iQ", raw, off)
    off += 13
    if state == STATE_NULL:
        raise ValueError(f"{path}: read origin in Unity checkpoint header shows null")
    dt = TYPE_DTYPES[kind]
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).copy()
    return UnityFile(kind, state, data)


def checkpoint_path(directory: str, kind: str, id: int) -> str:
    """getCheckpoint<T>(directory, id) of the stage loaders (ref: src/Pipeline.cu:4-6): directory + "/" + id + "_" + name + ".uty" """
    return f"{directory}/{id}_{TYPE_NAMES[kind]}.uty"


def image_checkpoint_path(directory: str, id: int) -> str:
    """getImgCheckpoint (ref: src/Pipeline.cu:8-10)"""
    return f"{directory}/{id}_N6ssrlcv5ImageE.cpimg"


def write_uty(path_or_dir: str, kind: str, data: np.ndarray, state: int = STATE_CPU, id: int | None = None) -> str:
    """Write `data` as a Unity<T> checkpoint.  If `id` is given the file name is `path_or_dir + "<id>_<mangled name>.uty"` --
    plain string concatenation, exactly like Unity<T>::checkpoint(id, dirPath) (ref: include/Unity.cuh:933-935): a dirPath
    without a trailing slash becomes a file-name prefix there, and does here."""
    name = TYPE_NAMES[kind]
    dt = TYPE_DTYPES[kind]
    data = np.ascontiguousarray(data, dtype=dt)
    path = f"{path_or_dir}{id}_{name}.uty" if id is not None else path_or_dir
    with open(path, "wb") as f:
        f.write(name.encode() + b"\n")
        f.write(struct.pack("<Q", libstdcxx_hash_bytes(name.encode())) + b"\n")
        f.write(struct.pack("<iQ", state, data.shape[0]) + b"\n")
        f.write(data.tobytes())
    return path


@dataclass
class ImageMeta:
    id: int
    size: tuple
    colorDepth: int
    camera: np.ndarray  # shape (), dtype types.CAMERA
    pushbroom: bytes = bytes(72)   # Image::PushbroomCamera @136 (the reference's reader copies it, src/Image.cu:55): carried through
    isPushbroom: int = 0           # bool @208


def read_cpimg(path: str) -> ImageMeta:
    with open(path, "rb") as f:
        raw = f.read(240)
    if len(raw) < 240:
        raise ValueError(f"{path}: short .cpimg ({len(raw)} bytes)")
    (img_id,) = struct.unpack_from("<i", raw, 32)
    size = struct.unpack_from("<II", raw, 40)
    (depth,) = struct.unpack_from("<I", raw, 48)
    cam = np.frombuffer(raw, dtype=T.CAMERA, count=1, offset=56)[0].copy()
    return ImageMeta(img_id, size, depth, cam, bytes(raw[136:208]), raw[208])


def write_cpimg(path: str, meta: ImageMeta) -> None:
    buf = bytearray(240)
    struct.pack_into("<i", buf, 32, meta.id)
    struct.pack_into("<II", buf, 40, *meta.size)
    struct.pack_into("<I", buf, 48, meta.colorDepth)
    buf[56:136] = np.asarray(meta.camera, dtype=T.CAMERA).tobytes()
    buf[136:208] = bytes(meta.pushbroom)[:72].ljust(72, b"\0")
    buf[208] = 1 if meta.isPushbroom else 0
    with open(path, "wb") as f:
        f.write(bytes(buf))
