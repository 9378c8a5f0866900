"""Host-side mirror of ssrlcv::MatchFactory<SIFT_Descriptor> over the C ABI.

Same method names, argument meaning, post-conditions and error behaviour as the reference class
(ref: include/MatchFactory.cuh:131-311, src/MatchFactory.cu:14-1028):
  * inputs are borrowed Unity arrays; their MemoryState is forced to gpu for the call and restored on exit
    (ref: src/MatchFactory.cu:350-353, :390-391)
  * outputs are new Unity arrays in state gpu, sized to the number of valid matches, ascending query index
  * argument errors raise MatchFactoryError where the reference logs and exit(-1)s
    (seed size mismatch :370-373, threshold == 0 :138-141)
All arithmetic happens in libssrlcv_match_b200.so (hand-written sm_100a CUDA); this module only marshals.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from . import types as T
from .unity import MemoryState, Unity, both, cpu, gpu

FLT_MAX = float(np.finfo(np.float32).max)
ENGINE_TCGEN05, ENGINE_SIMT = 0, 1


class MatchFactoryError(RuntimeError):
    """Raised where the reference prints to logger.err and calls exit(-1)."""


@dataclass
class Image:
    """The fields of ssrlcv::Image the matching path reads (ref: include/Image.cuh:79-85)."""
    id: int = -1
    camera: np.ndarray = field(default_factory=lambda: np.zeros((), T.CAMERA))

    @property
    def size(self):
        return tuple(int(x) for x in self.camera["size"])


@dataclass
class MatchSet:
    keyPoints: Unity
    matches: Unity


class Context:
    """One per (process, device): owns the library context bound to torch's current stream."""
    _instances: dict = {}

    def __init__(self, device: int = 0):
        import torch

        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.B200Error(-2, "no CUDA device: the matching path has no CPU fallback")
        torch.cuda.set_device(device)
        self.device = device
        self.handle = C.c_void_p()
        # torch's default stream is the legacy default stream (handle 0): name it explicitly (cudaStreamLegacy = 1) so the
        # library's work is ordered with the caller's torch ops instead of running on a private non-blocking stream
        stream = torch.cuda.current_stream(device).cuda_stream or 1
        rc = self.lib.ssrlcv_b200_create(C.byref(self.handle), device, C.c_void_p(stream))
        if rc != 0:
            raise _lib.B200Error(rc, "ssrlcv_b200_create failed (needs an sm_100 GPU)")

    @classmethod
    def get(cls, device: int = 0) -> "Context":
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    def check(self, rc: int) -> None:
        _lib.check(self.handle, rc)

    def launch_count(self) -> int:
        return int(self.lib.ssrlcv_b200_launch_count(self.handle))

    def tune(self, knob: int, value: int) -> None:
        """ssrlcv_b200_tune: -1 = the library decides (tests / A-B timing only, see include/ssrlcv_b200.h)."""
        self.check(self.lib.ssrlcv_b200_tune(self.handle, knob, int(value)))


class PackedSet:
    def __init__(self, ctx: Context, handle: C.c_void_p, n: int):
        self.ctx, self.handle, self.n = ctx, handle, n

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.ssrlcv_b200_set_free(self.handle)
                self.handle = None
        except Exception:
            pass


def _pack_key(features: Unity):
    d = features.device
    return (d.data_ptr(), d._version, features.size()) if d is not None else (0, 0, features.size())


def pack(ctx: Context, features: Unity) -> PackedSet:
    """Repack (and cache on the Unity) the device copy of a Feature<SIFT_Descriptor> array.  The cache is keyed on the
    device buffer's address, torch version counter and length, so an in-place edit of the buffer repacks."""
    key = _pack_key(features)
    if features._packed is not None and features._packed.ctx is ctx and getattr(features._packed, "key", None) == key:
        return features._packed
    h = C.c_void_p()
    ctx.check(ctx.lib.ssrlcv_b200_pack(ctx.handle, C.c_void_p(features.device_ptr() if features.size() else 0),
                                       features.size(), C.byref(h)))
    features._packed = PackedSet(ctx, h, features.size())
    features._packed.key = key
    return features._packed


def pack_many(ctx: Context, sets) -> list:
    """pack() for several Unity arrays with ONE launch for all that are not cached yet (ssrlcv_b200_pack_batch)."""
    todo = []
    for u in sets:
        if not any(u is v for v in todo) and not (u._packed is not None and u._packed.ctx is ctx and getattr(u._packed, "key", None) == _pack_key(u)):
            todo.append(u)
    if len(todo) > 1:
        n = len(todo)
        ptrs = (C.c_void_p * n)(*[u.device_ptr() if u.size() else 0 for u in todo])
        sizes = (C.c_uint64 * n)(*[u.size() for u in todo])
        outs = (C.c_void_p * n)()
        ctx.check(ctx.lib.ssrlcv_b200_pack_batch(ctx.handle, n, ptrs, sizes, outs))
        for u, h in zip(todo, outs):
            u._packed = PackedSet(ctx, C.c_void_p(h), u.size())
            u._packed.key = _pack_key(u)
    return [pack(ctx, u) for u in sets]


class MultiGpu:
    """The pair scheduler across the sm_100 devices of the box (ssrlcv_b200_match_pairs_multi_gpu): one process, one library
    context + host thread per device.  replaces: the pair loop of generateMatchesExhaustive (ref: src/MatchFactory.cu:907-936)."""
    _instances: dict = {}

    def __init__(self, n_devices: int = 0):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        rc = self.lib.ssrlcv_b200_multi_create(C.byref(self.handle), int(n_devices), None)
        if rc != 0:
            raise _lib.B200Error(rc, "ssrlcv_b200_multi_create failed (needs sm_100 GPUs)")

    @classmethod
    def get(cls, n_devices: int = 0) -> "MultiGpu":
        if n_devices not in cls._instances:
            cls._instances[n_devices] = cls(n_devices)
        return cls._instances[n_devices]

    def device_count(self) -> int:
        return int(self.lib.ssrlcv_b200_multi_device_count(self.handle))

    @staticmethod
    def _pointer(u):
        """(address, rows) of a Feature array: Unity (device copy if it has one, else host), numpy array or torch tensor"""
        if isinstance(u, Unity):
            if u.getMemoryState() in (gpu, both) and u.device is not None:
                return u.device_ptr() if u.size() else 0, u.size()
            return (u.host.ctypes.data if u.size() else 0), u.size()
        if isinstance(u, np.ndarray):
            return (u.ctypes.data if len(u) else 0), len(u)
        return (u.data_ptr() if u.numel() else 0), u.numel() // T.FEATURE.itemsize

    def assignment(self, sizes, pairs):
        n = len(pairs)
        dev = (C.c_int32 * max(n, 1))()
        pq = (C.c_uint32 * max(n, 1))(*[p[0] for p in pairs])
        pt = (C.c_uint32 * max(n, 1))(*[p[1] for p in pairs])
        rc = self.lib.ssrlcv_b200_multi_assignment(self.handle, len(sizes), (C.c_uint64 * len(sizes))(*sizes), n, pq, pt, dev)
        if rc != 0:
            raise _lib.B200Error(rc, "ssrlcv_b200_multi_assignment")
        return [int(x) for x in dev[:n]]

    def match_pairs(self, features, pairs, params, seed=None, out_device: int = 0):
        """features: one Feature array per view; pairs: [(query view, target view)]; params: one MatchParams per pair.
        Returns one Unity per pair (state gpu on `out_device`), exactly what the single-GPU generate* call returns."""
        import torch

        n_views, n = len(features), len(pairs)
        ptrs, sizes = zip(*[self._pointer(f) for f in features]) if n_views else ((), ())
        F = (C.c_void_p * max(n_views, 1))(*ptrs)
        N = (C.c_uint64 * max(n_views, 1))(*sizes)
        sp, sn = self._pointer(seed) if seed is not None else (None, 0)
        P = (_lib.MatchParams * max(n, 1))()
        for i, p in enumerate(params):
            C.memmove(C.byref(P, i * C.sizeof(_lib.MatchParams)), C.byref(p), C.sizeof(_lib.MatchParams))
        pq = (C.c_uint32 * max(n, 1))(*[p[0] for p in pairs])
        pt = (C.c_uint32 * max(n, 1))(*[p[1] for p in pairs])
        outs, O = [], (C.c_void_p * max(n, 1))()
        for i, (q, _) in enumerate(pairs):
            dt = T.OUT_DTYPES[params[i].out_kind]
            o = torch.empty(max(sizes[q], 1) * dt.itemsize, dtype=torch.uint8, device=f"cuda:{out_device}")
            outs.append(o)
            O[i] = o.data_ptr()
        torch.cuda.synchronize()                          # the workers' streams are not ordered after torch's
        n_out = (C.c_uint64 * max(n, 1))()
        rc = self.lib.ssrlcv_b200_match_pairs_multi_gpu(self.handle, n_views, F, N, C.c_void_p(sp) if sp else None, sn, n, pq, pt, P, O, n_out)
        if rc != 0:
            msg = self.lib.ssrlcv_b200_multi_last_error(self.handle)
            raise _lib.B200Error(rc, msg.decode() if msg else "")
        res = []
        for i in range(n):
            dt = T.OUT_DTYPES[params[i].out_kind]
            k = int(n_out[i])
            res.append(Unity(outs[i][: k * dt.itemsize], dt, gpu, out_device, k))
        return res


class MatchFactory:
    def __init__(self, relativeThreshold: float = 1.0, absoluteThreshold: float = FLT_MAX, device: int = 0,
                 engine: int = ENGINE_TCGEN05):
        self.relativeThreshold = float(np.float32(relativeThreshold))
        self.absoluteThreshold = float(np.float32(absoluteThreshold))
        self.seedFeatures: Unity | None = None
        self.device = device
        self.engine = engine
        self.ctx = Context.get(device)

    # ------------------------------------------------------------------ helpers
    def setSeedFeatures(self, seedFeatures: Unity) -> None:
        self.seedFeatures = seedFeatures

    def _torch(self):
        import torch

        return torch

    def _params(self, out_kind, mask_kind, query: Image, target: Image, epsilon=0.0, delta=0.0, fundamental=None,
                compact=1) -> _lib.MatchParams:
        p = _lib.MatchParams()
        p.out_kind, p.mask_kind = out_kind, mask_kind
        p.absolute_threshold, p.relative_threshold = self.absoluteThreshold, self.relativeThreshold
        p.query_image_id, p.target_image_id = query.id & 0xFFFFFFFF, target.id & 0xFFFFFFFF
        p.epsilon, p.delta = float(np.float32(epsilon)), float(np.float32(delta))
        if fundamental is not None:
            p.fundamental = (C.c_float * 9)(*np.asarray(fundamental, np.float32).reshape(9))
        p.query_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(query.camera, T.CAMERA).tobytes())
        p.target_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(target.camera, T.CAMERA).tobytes())
        p.compact, p.reserved = compact, 0
        return p

    def _select_engine(self) -> None:
        # engine 1 = the SIMT cross-check kernel (tests only): a knob of the context, not a parameter of the call
        self.ctx.tune(_lib.TUNE_CROSSCHECK_SIMT, 1 if self.engine == ENGINE_SIMT else -1)

    def _run(self, p: _lib.MatchParams, queryFeatures: Unity, targetFeatures: Unity, seedDistances: Unity | None) -> Unity:
        torch = self._torch()
        origin = (queryFeatures.getMemoryState(), targetFeatures.getMemoryState())
        if origin[0] != gpu:
            queryFeatures.setMemoryState(gpu)
        if origin[1] != gpu:
            targetFeatures.setMemoryState(gpu)
        try:
            nq = queryFeatures.size()
            seed_ptr, n_seed = None, 0
            seed_origin = None
            if seedDistances is not None:
                if seedDistances.size() != nq:
                    raise MatchFactoryError("ERROR: seedDistances should have come from matching a seed image to queryFeatures")
                seed_origin = seedDistances.getMemoryState()
                if seed_origin != gpu:
                    seedDistances.setMemoryState(gpu)
                seed_ptr, n_seed = C.c_void_p(seedDistances.device_ptr()), nq
            q, t = pack_many(self.ctx, [queryFeatures, targetFeatures])
            self._select_engine()
            dt = T.OUT_DTYPES[p.out_kind]
            out = torch.empty(max(nq, 1) * dt.itemsize, dtype=torch.uint8, device=f"cuda:{self.device}")
            n_out = C.c_uint64(0)
            self.ctx.check(self.ctx.lib.ssrlcv_b200_match(self.ctx.handle, C.byref(p), q.handle, t.handle, seed_ptr, n_seed,
                                                          C.c_void_p(out.data_ptr()), C.byref(n_out)))
            if seedDistances is not None and seed_origin != gpu:
                seedDistances.setMemoryState(seed_origin)
            n = int(n_out.value)
            if p.compact and n == 0 and nq > 0:
                # reference quirk (src/MatchFactory.cu:40-45): validateMatches clears only its own copy of the
                # smart pointer, so the caller keeps the uncompacted array of invalid entries
                p2 = _lib.MatchParams.from_buffer_copy(bytes(p))
                p2.compact = 0
                self.ctx.check(self.ctx.lib.ssrlcv_b200_match(self.ctx.handle, C.byref(p2), q.handle, t.handle, seed_ptr, n_seed,
                                                              C.c_void_p(out.data_ptr()), C.byref(n_out)))
                n = int(n_out.value)
            return Unity(out[: n * dt.itemsize].clone() if n != nq else out[: n * dt.itemsize], dt, gpu, self.device, n)
        finally:
            if origin[0] != gpu:
                queryFeatures.setMemoryState(origin[0])
            if origin[1] != gpu:
                targetFeatures.setMemoryState(origin[1])

    # ------------------------------------------------------------------ seed distances
    def getSeedDistances(self, features: Unity) -> Unity:
        """ref: src/MatchFactory.cu:315-346"""
        torch = self._torch()
        if self.seedFeatures is None:
            raise MatchFactoryError("getSeedDistances called without seed features")
        origin = features.getMemoryState()
        if self.seedFeatures.getMemoryState() != gpu:
            self.seedFeatures.setMemoryState(gpu)
        if origin != gpu:
            features.setMemoryState(gpu)
        try:
            q, s = pack_many(self.ctx, [features, self.seedFeatures])
            self._select_engine()
            out = torch.empty(max(features.size(), 1), dtype=torch.float32, device=f"cuda:{self.device}")
            self.ctx.check(self.ctx.lib.ssrlcv_b200_seed_distances(self.ctx.handle, q.handle, s.handle,
                                                                   C.c_void_p(out.data_ptr())))
            self.ctx.check(self.ctx.lib.ssrlcv_b200_synchronize(self.ctx.handle))
            return Unity(out[: features.size()], np.dtype("<f4"), gpu, self.device, features.size())
        finally:
            if origin != gpu:
                features.setMemoryState(origin)

    # ------------------------------------------------------------------ Match
    def generateMatches(self, query, queryFeatures, target, targetFeatures, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:349-394"""
        return self._run(self._params(T.OUT_MATCH, T.MASK_NONE, query, target), queryFeatures, targetFeatures, seedDistances)

    def generateMatchesConstrained(self, query, queryFeatures, target, targetFeatures, epsilon, fundamental, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:395-445"""
        return self._run(self._params(T.OUT_MATCH, T.MASK_FUNDAMENTAL, query, target, epsilon, 0.0, fundamental),
                         queryFeatures, targetFeatures, seedDistances)

    def generateMatchesDoubleConstrained(self, query, queryFeatures, target, targetFeatures, epsilon, delta, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:447-501"""
        return self._run(self._params(T.OUT_MATCH, T.MASK_DOUBLE, query, target, epsilon, delta), queryFeatures, targetFeatures, seedDistances)

    # ------------------------------------------------------------------ DMatch
    def generateDistanceMatches(self, query, queryFeatures, target, targetFeatures, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:504-547"""
        return self._run(self._params(T.OUT_DMATCH, T.MASK_NONE, query, target), queryFeatures, targetFeatures, seedDistances)

    def generateDistanceMatchesConstrained(self, query, queryFeatures, target, targetFeatures, epsilon, fundamental, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:549-597"""
        return self._run(self._params(T.OUT_DMATCH, T.MASK_FUNDAMENTAL, query, target, epsilon, 0.0, fundamental),
                         queryFeatures, targetFeatures, seedDistances)

    def generateDistanceMatchesDoubleConstrained(self, query, queryFeatures, target, targetFeatures, epsilon, delta, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:600-651"""
        return self._run(self._params(T.OUT_DMATCH, T.MASK_DOUBLE, query, target, epsilon, delta), queryFeatures, targetFeatures, seedDistances)

    # ------------------------------------------------------------------ FeatureMatch
    def generateFeatureMatches(self, query, queryFeatures, target, targetFeatures, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:654-697 -- quirk kept: neither synchronises nor compacts, returns numQuery raw entries."""
        return self._run(self._params(T.OUT_FEATUREMATCH, T.MASK_NONE, query, target, compact=0), queryFeatures, targetFeatures, seedDistances)

    def generateFeatureMatchesConstrained(self, query, queryFeatures, target, targetFeatures, epsilon, fundamental, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:699-750"""
        return self._run(self._params(T.OUT_FEATUREMATCH, T.MASK_FUNDAMENTAL, query, target, epsilon, 0.0, fundamental),
                         queryFeatures, targetFeatures, seedDistances)

    # ------------------------------------------------------------------ uint2_pair
    def generateMatchesIndexOnly(self, query, queryFeatures, target, targetFeatures, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:755-800"""
        return self._run(self._params(T.OUT_PAIR, T.MASK_NONE, query, target), queryFeatures, targetFeatures, seedDistances)

    def generateMatchesConstrainedIndexOnly(self, query, queryFeatures, target, targetFeatures, epsilon, fundamental, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:802-851"""
        return self._run(self._params(T.OUT_PAIR, T.MASK_FUNDAMENTAL, query, target, epsilon, 0.0, fundamental),
                         queryFeatures, targetFeatures, seedDistances)

    def generateMatchesDoubleConstrainedIndexOnly(self, query, queryFeatures, target, targetFeatures, epsilon, delta, seedDistances=None) -> Unity:
        """ref: src/MatchFactory.cu:853-905"""
        return self._run(self._params(T.OUT_PAIR, T.MASK_DOUBLE, query, target, epsilon, delta), queryFeatures, targetFeatures, seedDistances)

    # ------------------------------------------------------------------ post-processing
    def _kind_of(self, matches: Unity) -> int:
        for k, dt in T.OUT_DTYPES.items():
            if dt == matches.dtype:
                return k
        raise MatchFactoryError("not a match array")

    def _inplace(self, matches: Unity, call) -> None:
        origin = matches.getMemoryState()
        if origin != gpu:
            matches.setMemoryState(gpu)
        n_out = C.c_uint64(0)
        self.ctx.check(call(C.c_void_p(matches.device_ptr() if matches.size() else 0), matches.size(), n_out))
        n = int(n_out.value)
        if n == 0:
            # reference quirk: "matches.clear(); return;" on a by-value smart pointer leaves the caller's array untouched
            if origin != gpu:
                matches.setMemoryState(origin)
            return
        kept = matches.device[: n * matches.dtype.itemsize].clone()
        matches.setData(kept, n, gpu)
        if origin != gpu:
            matches.setMemoryState(origin)

    def validateMatches(self, matches: Unity) -> None:
        """ref: src/MatchFactory.cu:31-135"""
        kind = self._kind_of(matches)
        self._inplace(matches, lambda p, n, out: self.ctx.lib.ssrlcv_b200_validate(self.ctx.handle, kind, p, n, C.byref(out)))

    def refineMatches(self, matches: Unity, threshold: float) -> None:
        """ref: src/MatchFactory.cu:137-160"""
        if float(threshold) == 0.0:
            raise MatchFactoryError("ERROR illegal value used for threshold: 0.0")
        self._inplace(matches, lambda p, n, out: self.ctx.lib.ssrlcv_b200_refine(self.ctx.handle, p, n, float(np.float32(threshold)), C.byref(out)))

    def sortMatches(self, matches: Unity) -> None:
        """ref: src/MatchFactory.cu:191-222 (device branch; equal distances keep input order)"""
        if matches.getFore() not in (cpu, gpu, both):
            raise MatchFactoryError(f"ERROR cannot perform sortMatches with matches->getMemoryState() = {int(matches.getMemoryState())}")
        state = matches.getMemoryState()
        if matches.getFore() == cpu:
            matches.transferMemoryTo(gpu) if state == both else matches.setMemoryState(gpu)
        self.ctx.check(self.ctx.lib.ssrlcv_b200_sort(self.ctx.handle, C.c_void_p(matches.device_ptr() if matches.size() else 0), matches.size()))
        matches.setFore(gpu)
        if state == both:
            matches.transferMemoryTo(cpu)
        elif state == cpu:
            matches.setMemoryState(cpu)

    def getRawMatches(self, matches: Unity) -> Unity:
        """ref: src/MatchFactory.cu:257-280"""
        torch = self._torch()
        origin = matches.getMemoryState()
        if origin == cpu:
            matches.setMemoryState(gpu)
        n = matches.size()
        out = torch.empty(max(n, 1) * T.MATCH.itemsize, dtype=torch.uint8, device=f"cuda:{self.device}")
        self.ctx.check(self.ctx.lib.ssrlcv_b200_raw_matches(self.ctx.handle, C.c_void_p(matches.device_ptr() if n else 0), n,
                                                            C.c_void_p(out.data_ptr())))
        res = Unity(out[: n * T.MATCH.itemsize], T.MATCH, gpu, self.device, n)
        if origin == cpu:
            matches.setMemoryState(cpu)
            res.setMemoryState(cpu)
        return res

    # ------------------------------------------------------------------ pair scheduler (device side)
    def matchBatch(self, jobs) -> list:
        """Run many independent generate* calls with one synchronisation (ssrlcv_b200_match_batch).
        jobs: list of (params, queryFeatures, targetFeatures, seedDistances-or-None); returns one Unity per job, exactly what
        the corresponding generate* call returns (ref: the pair loop of src/MatchFactory.cu:907-936)."""
        torch = self._torch()
        n = len(jobs)
        if n == 0:
            return []
        touched = {}
        for _, qf, tf, sd in jobs:
            for u in (qf, tf, sd):
                if u is not None and id(u) not in touched:
                    touched[id(u)] = (u, u.getMemoryState())
                    if u.getMemoryState() != gpu:
                        u.setMemoryState(gpu)
        try:
            P = (_lib.MatchParams * n)()
            Q, Tt, S, O = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
            outs, keep = [], []
            pack_many(self.ctx, [u for _, qf, tf, _ in jobs for u in (qf, tf)])      # one launch for every view of the batch
            self._select_engine()
            for i, (p, qf, tf, sd) in enumerate(jobs):
                if sd is not None and sd.size() != qf.size():
                    raise MatchFactoryError("ERROR: seedDistances should have come from matching a seed image to queryFeatures")
                C.memmove(C.byref(P, i * C.sizeof(_lib.MatchParams)), C.byref(p), C.sizeof(_lib.MatchParams))
                q, t = pack(self.ctx, qf), pack(self.ctx, tf)
                keep += [q, t]
                Q[i], Tt[i] = q.handle, t.handle
                S[i] = sd.device_ptr() if sd is not None else None
                dt = T.OUT_DTYPES[p.out_kind]
                out = torch.empty(max(qf.size(), 1) * dt.itemsize, dtype=torch.uint8, device=f"cuda:{self.device}")
                outs.append(out)
                O[i] = out.data_ptr()
            n_out = (C.c_uint64 * n)()
            self.ctx.check(self.ctx.lib.ssrlcv_b200_match_batch(self.ctx.handle, n, P, Q, Tt, S, O, n_out))
            res = []
            for i, (p, qf, tf, sd) in enumerate(jobs):
                k = int(n_out[i])
                dt = T.OUT_DTYPES[p.out_kind]
                if p.compact and k == 0 and qf.size() > 0:
                    res.append(self._run(p, qf, tf, sd))          # rare reference quirk path (all-invalid array is kept), see _run
                else:
                    res.append(Unity(outs[i][: k * dt.itemsize], dt, gpu, self.device, k))
            return res
        finally:
            for u, st in touched.values():
                if st != gpu:
                    u.setMemoryState(st)

    # ------------------------------------------------------------------ N-view
    def generateMatchesExhaustive(self, images, features, epsilon, delta, ordered=True, estimatedOverlap=0.0,
                                  geo_orbit: bool = True) -> MatchSet:
        """ref: src/MatchFactory.cu:907-1028.  Pair loop on the GPU; the adjacency fusion follows in pair_fusion."""
        from .pair_scheduler import exhaustive_pairs
        from .fusion import fuse_pairs

        pair_lists = exhaustive_pairs(self, images, features, epsilon, delta, ordered, estimatedOverlap, geo_orbit)
        if sum(p.size() for p in pair_lists) == 0:
            raise MatchFactoryError("There were no matches found in the set of images, likely due to unreasonable threshold")
        mm, kp = fuse_pairs([f.to_numpy() for f in features], [p.to_numpy() for p in pair_lists])
        return MatchSet(keyPoints=Unity(kp, T.KEYPOINT, cpu, self.device), matches=Unity(mm, T.MULTIMATCH, cpu, self.device))
