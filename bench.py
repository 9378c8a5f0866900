#!/usr/bin/env python
"""bench.py -- headline benchmark of the SIFT matching hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3|C2|C1|C4|C5]
  (N > 1: launched by torchrun, one rank per GPU.  C1-C3: each rank matches its own image pair of the same
   shape -- pair-sharded N-view matching, no data-path collective -- so scaling is "weak".
   C4: 16 views x 32 768 features, the 120 pairs dealt to the ranks (strong scaling, one result exchange at the end).
   C5: one 1 M x 1 M pair, target set sharded over the ranks, candidates merged through an NCCL all-gather.)

A "step" is one pass of the hot path over one synthetic 2-view batch:
  C3 (default): generateDistanceMatches on 65 536 x 65 536 SIFT-shaped uint8[128] descriptors, unconstrained
  C2          : getSeedDistances (8 192 seed) + generateDistanceMatchesDoubleConstrained on 16 384 x 16 384 (GEO_ORBIT=1)
  C1          : getSeedDistances (1 024 seed) + generateMatches on 4 096 x 4 096
`value`  = distance evaluations / s, whole job, with the Feature<SIFT_Descriptor> arrays already resident in HBM
           (the step still repacks them: the reference API hands over 152-byte AoS records).
`e2e`    = same metric through ssrlcv_b200_match_host: HOST buffers in, H2D + pack + match + D2H inside the timed region.
`roofline` = the tensor-core matcher kernel alone (CUDA events on its stream): 256 int8 ops per distance evaluation.
`cpu_baseline` = the CPU oracle (OpenMP port of the same arithmetic) on a bounded sample of the same workload.
--impl reference times the reference's own CUDA MatchFactory kernels (oracle/_ref, rebuilt for sm_100a) on the same
inputs through the reference's public API; if that library is absent, the CPU oracle port on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C1": dict(nq=4096, nt=4096, ns=1024, mask=0, desc="2-view 4096x4096 + 1024 seed, unconstrained generateMatches"),
    "C2": dict(nq=16384, nt=16384, ns=8192, mask=2, desc="2-view 16384x16384 + 8192 seed, seed ratio test + DoubleConstrained (GEO_ORBIT=1)"),
    "C3": dict(nq=65536, nt=65536, ns=0, mask=0, desc="2-view 65536x65536 unconstrained generateDistanceMatches (tensor-pipe roofline case)"),
}
WORKLOADS["C4"] = dict(views=16, n=32768, nq=32768, nt=32768, ns=0, mask=0,
                       desc="N-view all-pairs: 16 views x 32768 features, 120 pairs, generateMatchesIndexOnly, pair-sharded over the ranks")
WORKLOADS["C5"] = dict(nq=1 << 20, nt=1 << 20, ns=0, mask=0,
                       desc="one 1048576 x 1048576 pair, generateDistanceMatches, target set sharded over the ranks + NCCL all-gather merge")
ABS_THR, REL_THR = float(os.environ.get("SSRLCV_BENCH_ABS", 200.0 * 200.0)), 0.6          # the reference's MatchFactory(0.6f, 200*200), src/Pipeline.cu:175
EPS, DELTA = 25.0, 5.0                         # test/Pipeline.cu:198


def synth_inputs(w, rank):
    from tests import synth
    q, t = synth.make_pair(w["nq"], w["nt"], seed=0x55524C43 + 17 * rank)
    s = synth.make_seed(w["ns"], seed=99 + rank) if w["ns"] else None
    return q, t, s


def cameras():
    from tests.golden_data import load_everest
    ev = load_everest()
    return ev.cameras[0], ev.cameras[1]


def config_of(name, w):
    """the `config` object both arms print (identical keys and values: the driver compares them)"""
    return {"workload": f"{name}: {w['desc']}", "nq": w["nq"], "nt": w["nt"], "ns": w["ns"],
            "l2": "256 MiB flush between timed iterations", "inputs": "Feature<SIFT_Descriptor> AoS arrays resident in HBM"}


def evals_per_step(w):
    if "views" in w:
        return (w["views"] * (w["views"] - 1) // 2) * w["n"] * w["n"]
    return w["nq"] * w["nt"] + w["nq"] * w["ns"]


def big_pair(nq, nt, seed):
    """make_pair in 65 536-row chunks (bounded host memory); planted correspondences stay inside a chunk pair."""
    from tests import synth
    qs, ts = [], []
    for c in range(max(nq, nt) // 65536):
        q, t = synth.make_pair(65536, 65536, seed=seed + 1000 * c)
        qs.append(q); ts.append(t)
    return np.concatenate(qs)[:nq], np.concatenate(ts)[:nt]


def cpu_slice(cpus, index, local_world):
    """The part of the CPU set `cpus` that local rank `index` of `local_world` ranks keeps for itself: equal contiguous slices in
    ascending CPU order; the whole set when the ranks would get fewer than two CPUs each."""
    order = sorted(cpus)
    per = len(order) // max(local_world, 1)
    if local_world <= 1 or per < 2:
        return set(order)
    r = index % local_world
    return set(order[r * per:(r + 1) * per])


def bind_to_gpu_numa(index, local_world=1):
    """Run this process on the CPUs local to GPU `index` (NVML's ideal CPU affinity) so that the pinned host buffers of the
    e2e leg are allocated, and their copies issued, on the NUMA node the GPU hangs off; a pinned buffer on the far socket
    costs up to 3x in H2D time (profiles/r1f_e2eprof_c3.txt).  With several ranks on one host whose GPUs share a CPU set (all
    eight GPUs of this pool's boxes report CPUs 0-31), every rank takes its own slice of it: the launch chain of a step is a handful
    of host calls a few microseconds apart, and eight ranks' threads migrating over the same cores show up as launch gaps.
    Returns the previous mask (restore before the CPU baseline, which wants every core), or None when NVML is not available."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        prev = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(prev) + 64) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1} & prev
        if cpus:
            cpus = cpu_slice(cpus, index, local_world)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def ensure(self, step_fn, sync_fn, min_rows=3, max_seconds=1.5):
        """A timed region shorter than nvidia-smi's first answer yields no sample: keep the same steps running (untimed)
        until a few samples under load exist."""
        if self.proc is None:
            return
        t0 = time.time()
        while len(self.rows) < min_rows and time.time() - t0 < max_seconds:
            step_fn()
            sync_fn()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(w, q, t, s, seconds=15.0):
    """CPU oracle on a bounded query sample of the same workload."""
    from oracle import oracle as O
    from ssrlcv_b200 import types as T
    cam0, cam1 = cameras()
    n = 256
    def run(nq_s):
        t0 = time.perf_counter()
        sd = O.seed_distances(q[:nq_s], s) if s is not None else None
        p = O.make_params(out_kind=T.OUT_DMATCH, mask_kind=w["mask"], has_seed=s is not None, abs_threshold=ABS_THR, rel_threshold=REL_THR,
                          epsilon=EPS, delta=DELTA, query_camera=cam0, target_camera=cam1)
        O.generate_matches(p, q[:nq_s], t, sd)
        return time.perf_counter() - t0
    dt = run(n)
    n2 = int(min(len(q), max(n, n * seconds / max(dt, 1e-6))))
    n2 = max(256, (n2 // 256) * 256)
    dt = run(n2)
    ev = n2 * (len(t) + (len(s) if s is not None else 0))
    return {"value": ev / dt, "unit": "dist-evals/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{n2} of {len(q)} queries against all {len(t)} targets" + (f" + {len(s)} seed" if s is not None else "") + f", {dt:.1f} s of CPU work"}


def run_reference(args, w, rank, world):
    """--impl reference: the reference's own CUDA kernels (oracle/_ref) through its public API."""
    if rank != 0:
        return None
    from ssrlcv_b200 import types as T
    scale, sample_note = 1.0, "full workload"
    if "views" in w:            # C4: every pair has the same shape -> time one pair, 120 pairs extrapolated
        from tests import synth
        q, t = synth.make_pair(w["n"], w["n"], seed=0x55524C43)
        s = None
        scale = w["views"] * (w["views"] - 1) / 2
        sample_note = f"one {w['n']} x {w['n']} pair of the {int(scale)}; ms_per_step extrapolated x{int(scale)}"
    elif args.workload == "C5":  # C5: 16 384-query subsample against the full target set, linear in Nq
        q, t = big_pair(w["nq"], w["nt"], 0x55524C43)
        q, s = q[:16384], None
        scale = w["nq"] / len(q)
        sample_note = f"{len(q)} of {w['nq']} queries against all {w['nt']} targets; ms_per_step extrapolated x{int(scale)}"
    else:
        q, t, s = synth_inputs(w, 0)
    cam0, cam1 = cameras()
    path = os.path.join(ROOT, "oracle", "_ref", "libssrlcv_ref.so")
    w = dict(w, mask=w.get("mask", 0))
    line = {"impl": "reference", "metric": "sift_dist_evals_per_s", "unit": "dist-evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if args.workload in ("C4", "C5") else "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": config_of(args.workload, w)}
    import torch
    if os.path.exists(path) and torch.cuda.is_available():
        run_dir = "/tmp/ssrlcv_ref_run"; os.makedirs(run_dir, exist_ok=True); os.chdir(run_dir)
        # the reference's global logger announces its log file on stdout when the library is loaded and may print later: this
        # process' stdout carries exactly one JSON line, so the C-level stdout goes to stderr while the reference runs
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        ref = C.CDLL(path); ref.ref_match.restype = C.c_long
        out = np.zeros(len(q), T.DMATCH)
        sd = np.zeros(len(q), np.float32)
        times = []
        for it in range(args.warmup + args.steps):
            ms_seed, ms = C.c_double(0), C.c_double(0)
            if s is not None:
                ref.ref_seed_distances(q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)), s.ctypes.data_as(C.c_void_p), C.c_ulong(len(s)), 2,
                                       sd.ctypes.data_as(C.POINTER(C.c_float)), C.byref(ms_seed))
            n = ref.ref_match(1, w["mask"], C.c_float(ABS_THR), C.c_float(REL_THR), 0, 1, q.ctypes.data_as(C.c_void_p), C.c_ulong(len(q)),
                              t.ctypes.data_as(C.c_void_p), C.c_ulong(len(t)), sd.ctypes.data_as(C.POINTER(C.c_float)) if s is not None else None, 2,
                              C.c_float(EPS), C.c_float(DELTA), None, np.asarray(cam0).tobytes(), np.asarray(cam1).tobytes(),
                              out.ctypes.data_as(C.c_void_p), C.byref(ms))
            if it >= args.warmup:
                times.append(ms.value + ms_seed.value)
        os.chdir(ROOT)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
        ms_step = sum(times) / len(times)
        v = (len(q) * (len(t) + (len(s) if s is not None else 0))) / (ms_step * 1e-3)
        line.update({"value": v, "ms_per_step": ms_step * scale, "matches_per_step": int(n),
                     "cpu_baseline": {"value": v, "unit": "dist-evals/s", "cores": 0, "kind": "reference",
                                      "sample": sample_note + "; the reference has no CPU path -- this is its own CUDA MatchFactory rebuilt for sm_100a, "
                                                "run on the GPU with device-resident inputs (API wall time incl. its cudaMalloc/compaction)"},
                     "e2e": {"value": v, "unit": "dist-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:
        cb = cpu_baseline(w, q, t, s, seconds=20.0)
        line.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "dist-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return line


class Env:
    """one per process: device, library context, (for N > 1) the NCCL process group"""

    def __init__(self, rank, world, local_rank):
        import torch
        from ssrlcv_b200.match_factory import Context
        self.torch, self.rank, self.world, self.local_rank = torch, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.prev_affinity = bind_to_gpu_numa(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            self.dist = dist
        self.ctx = Context.get(local_rank)
        # A/B timing of kernel variants: SSRLCV_BENCH_TUNE="fold=0,colmask=1,chunk_q=16384,chunk_t=16384" (tools/ab_fold.sh, e2e_chunks.sh)
        from ssrlcv_b200 import _lib
        knob_ids = {"fold": _lib.TUNE_FOLD, "colmask": _lib.TUNE_COLUMN_MASKS, "sorted": _lib.TUNE_SORTED_TARGETS, "pair": _lib.TUNE_PAIR_UNITS, "seedonly": _lib.TUNE_SEED_ONLY,
                    "chunk_q": _lib.TUNE_HOST_CHUNK_QUERY_ROWS, "chunk_t": _lib.TUNE_HOST_CHUNK_TARGET_ROWS}
        for kv in filter(None, os.environ.get("SSRLCV_BENCH_TUNE", "").split(",")):
            k, v = kv.split("=")
            self.ctx.tune(knob_ids[k.strip()], int(v))
        self.dev = f"cuda:{local_rank}"
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)       # > 126 MB L2

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """CUDA events around every step on the library's stream, an L2 flush (not timed) in front of each"""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        wall0 = time.perf_counter()
        n = 0
        for i in range(steps):
            self.flush.fill_(i & 0xFF)
            ev[i][0].record()
            n = fn()
            ev[i][1].record()
        self.barrier()
        wall = time.perf_counter() - wall0
        return sum(a.elapsed_time(b) for a, b in ev), n, wall

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def roofline_of(world, evals_per_launch_set, kern_ms, k_n, steps, kernel_note=""):
    bf16_burst, bf16_sust, hbm, src = measured_peaks()
    peak_tops = 2.0 * bf16_burst                     # dense int8 tensor rate = 2 x dense bf16
    ops_per_launch = 256.0 * evals_per_launch_set / max(k_n / steps, 1)
    achieved = ops_per_launch / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    mma_meas = None
    mp = os.path.join(ROOT, "profiles", "int8_mma_peak.json")       # tools/microbench.cu on this pool (MMA-only issue rate)
    if os.path.exists(mp):
        mma_meas = json.load(open(mp)).get("int8_tops")
    r = {"bound": "tensor", "achieved": achieved, "peak": peak_tops, "unit": "TFLOP/s", "frac": achieved / peak_tops,
         "frac_vs_nominal": achieved / 4500.0, "frac_vs_int8_mma_measured": (achieved / mma_meas) if mma_meas else None,
         "int8_mma_measured": mma_meas,
         "kernel": "smax_pair_kernel / smax_kernel (tcgen05.mma kind::i8 + fused argmax epilogue; large unmasked searches take two query tiles per unit)" + kernel_note, "kernel_ms": kern_ms, "launches_timed": int(k_n),
         "ops": "int8 dense: 256 ops per distance evaluation (128 u8 x u8 MACs); the fold's extra K block is not counted",
         "peak_source": f"2 x bf16_tflops burst, {src}; nominal dense int8 is 4500; int8_mma_measured = MMA-only microbenchmark (tools/microbench.cu)"}
    return r, peak_tops


def run_ours(args, w, env):
    torch = env.torch
    from ssrlcv_b200 import _lib, types as T
    rank, world, dev = env.rank, env.world, env.dev
    ctx = env.ctx
    lib, h = ctx.lib, ctx.handle
    q, t, s = synth_inputs(w, rank)
    cam0, cam1 = cameras()
    to_dev = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    qd, td = to_dev(q), to_dev(t)
    sd_feat = to_dev(s) if s is not None else None
    out = torch.empty(len(q) * 48, dtype=torch.uint8, device=dev)
    seed_dist = torch.empty(len(q), dtype=torch.float32, device=dev)

    p = _lib.MatchParams()
    p.out_kind, p.mask_kind = T.OUT_DMATCH, w["mask"]
    p.absolute_threshold, p.relative_threshold = ABS_THR, REL_THR
    p.query_image_id, p.target_image_id, p.epsilon, p.delta = 0, 1, EPS, DELTA
    p.query_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(cam0).tobytes())
    p.target_camera = (C.c_uint8 * 80).from_buffer_copy(np.asarray(cam1).tobytes())
    p.compact, p.reserved = 1, 0
    n_out = C.c_uint64(0)
    seed_ptr = C.c_void_p(sd_feat.data_ptr()) if s is not None else None
    seed_out = C.c_void_p(seed_dist.data_ptr()) if s is not None else None

    def step():
        """pack every set (one launch), seed distances, match, release; inputs are device-resident AoS arrays."""
        ctx.check(lib.ssrlcv_b200_match_features(h, C.byref(p), C.c_void_p(qd.data_ptr()), len(q), C.c_void_p(td.data_ptr()), len(t),
                                                 seed_ptr, len(s) if s is not None else 0, seed_out, C.c_void_p(out.data_ptr()), C.byref(n_out)))
        return int(n_out.value)

    # ---------------- device-resident timing ----------------
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(env.local_rank) if rank == 0 else None      # one sampler per job, not per rank
    launches0 = ctx.launch_count()
    total_ms, n_matches, wall = env.timed(step, args.steps)
    launches = ctx.launch_count() - launches0
    # The matcher's own duration (roofline): CUDA events around every matcher launch on the library's stream, over the same K
    # steps run once more.  Not in the pass above: an event record between two kernels serialises them, and the chain's
    # kernels are launched with programmatic stream serialisation (each sets itself up while its predecessor drains).
    k_ms, k_n = C.c_double(0), C.c_uint64(0)
    lib.ssrlcv_b200_set_profiling(h, 1)
    env.timed(step, args.steps)
    lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
    lib.ssrlcv_b200_set_profiling(h, 0)
    clocks = None
    if sampler is not None:
        sampler.ensure(step, torch.cuda.synchronize)    # short timed regions: same steps, untimed, until nvidia-smi has answered
        clocks = sampler.stop()

    # ---------------- end-to-end through the C ABI with host buffers ----------------
    pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
    qh, th = pin(q), pin(t)
    sh = pin(s) if s is not None else None
    out_h = torch.empty(len(q) * 48, dtype=torch.uint8).pin_memory()

    def step_e2e():
        ctx.check(lib.ssrlcv_b200_match_host(h, C.byref(p), C.c_void_p(qh.data_ptr()), len(q), C.c_void_p(th.data_ptr()), len(t),
                                             C.c_void_p(sh.data_ptr()) if sh is not None else None, len(s) if s is not None else 0,
                                             C.c_void_p(out_h.data_ptr()), C.byref(n_out)))
        return int(n_out.value)

    for _ in range(2):
        step_e2e()
    e2e_ms, n_e2e, _ = env.timed(step_e2e, args.steps)

    # ---------------- max over ranks ----------------
    tt = torch.tensor([total_ms, e2e_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if env.dist is not None:
        env.dist.all_reduce(tt, op=env.dist.ReduceOp.MAX)
    total_ms, e2e_ms, wall_ms = (float(x) for x in tt.cpu())
    if rank != 0:
        return None

    evals = evals_per_step(w)
    value = world * evals * args.steps / (total_ms * 1e-3)
    kern_ms = k_ms.value / max(k_n.value, 1)
    # per launch of the matcher kernel: the main (query x target) contraction dominates; seed launches are averaged in
    roof, peak_tops = roofline_of(world, evals, kern_ms, k_n.value, args.steps)
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    roof["traffic"] = json.load(open(tp)).get(args.workload) if os.path.exists(tp) else None
    line = {
        "metric": "sift_dist_evals_per_s", "value": value, "unit": "dist-evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(args.workload, w),
        "parallelism": f"pair-sharded x{world}: every rank matches its own image pair of this shape, no data-path collective",
        "step": "ssrlcv_b200_match_features: pack(AoS->SoA, one launch) + [seed distances] + match + compaction, device-resident inputs",
        "matches_per_s": world * n_matches * args.steps / (total_ms * 1e-3), "queries_per_s": world * w["nq"] * args.steps / (total_ms * 1e-3),
        "matches_per_step": n_matches, "wall_ms_per_step": wall_ms / args.steps,
        "int8_tops": 256.0 * value / 1e12, "frac_of_int8_peak_step": 256.0 * value / 1e12 / (world * peak_tops),
        "frac_of_nominal_int8_step": 256.0 * value / 1e12 / (world * 4500.0),
        "clocks": clocks,
        "e2e": {"value": world * evals * args.steps / (e2e_ms * 1e-3), "unit": "dist-evals/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": (len(q) + len(t) + (len(s) if s is not None else 0)) * 152, "d2h_bytes_per_step": n_e2e * 48 + 8,
                "api": "ssrlcv_b200_match_host (pinned host buffers in, records out)",
                "host": "process bound to the GPU-local CPUs (NVML affinity) while the pinned buffers are allocated and used" if env.prev_affinity else "no NUMA binding"},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    line["cpu_baseline"] = None
    if world == 1 and not args.no_cpu_baseline:
        if env.prev_affinity:
            os.sched_setaffinity(0, env.prev_affinity)       # the CPU baseline uses every core the box gives us
        try:
            line["cpu_baseline"] = cpu_baseline(w, q, t, s)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "dist-evals/s", "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}
    return line


def run_sharded(args, name, w, env, with_clocks=True, with_e2e=True):
    """C4 (pairs dealt to ranks) and C5 (target set sharded over ranks, NCCL all-gather of candidate keys).  After the timed
    regions rank 0 recomputes the whole job alone on its GPU and compares the sharded records with it byte for byte
    (`parity`, outside the timed regions)."""
    torch = env.torch
    from ssrlcv_b200 import _lib, types as T
    from ssrlcv_b200.distributed import exchange_pair_lists, shard_bounds
    from ssrlcv_b200.pair_scheduler import assign_pairs, pair_order
    rank, world, dev, dist = env.rank, env.world, env.dev, env.dist
    ctx = env.ctx
    lib, h = ctx.lib, ctx.handle
    pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
    p = _lib.MatchParams()
    p.absolute_threshold, p.relative_threshold, p.compact, p.reserved = ABS_THR, REL_THR, 1, 0
    n_out = C.c_uint64(0)
    evals = evals_per_step(w)
    is_c4 = "views" in w
    state = {"exchange_bytes": 0}

    if is_c4:
        from tests import synth
        V, n = w["views"], w["n"]
        views = []
        for k in range(V // 2):
            a, b = synth.make_pair(n, n, seed=0x55524C43 + k)
            views += [a, b]
        pairs = pair_order(V)
        assign = assign_pairs(pairs, [n] * V, world)
        mine = assign[rank]
        need = sorted({v for k in mine for v in pairs[k]})
        host = {v: pin(views[v]) for v in need}
        devbuf = {v: host[v].to(dev) for v in need}
        slots = max(len(a) for a in assign)
        out = torch.zeros((max(slots, 1), n, 4), dtype=torch.int32, device=dev)      # uint2_pair records, one slot per owned pair
        p.out_kind, p.mask_kind = T.OUT_PAIR, 0
        h2d = sum(len(views[v]) * 152 for v in need)
        # per-pair parameters and output slots do not change from step to step; only the packed-set handles do
        m = len(mine)
        P = (_lib.MatchParams * max(m, 1))()
        Q, Tt, O = (C.c_void_p * max(m, 1))(), (C.c_void_p * max(m, 1))(), (C.c_void_p * max(m, 1))()
        n_outs = (C.c_uint64 * max(m, 1))()
        n_outs_np = np.ctypeslib.as_array(n_outs)
        for s_i, k in enumerate(mine):
            p.query_image_id, p.target_image_id = pairs[k]
            C.memmove(C.byref(P, s_i * C.sizeof(_lib.MatchParams)), C.byref(p), C.sizeof(_lib.MatchParams))
            O[s_i] = out[s_i].data_ptr()
        nv = len(need)
        vptr = (C.c_void_p * nv)(*[devbuf[v].data_ptr() for v in need])
        vsize = (C.c_uint64 * nv)(*([n] * nv))
        vsets = (C.c_void_p * nv)()
        vpos = {v: i for i, v in enumerate(need)}
        xch = exchange_pair_lists(len(pairs), assign, n, rank, world, dist, dev)

        def step(from_host=False):
            if from_host:
                for v in need:
                    devbuf[v].copy_(host[v], non_blocking=True)
            ctx.check(lib.ssrlcv_b200_pack_batch(h, nv, vptr, vsize, vsets))      # every view this rank touches: ONE pack launch per step
            # the pair scheduler's device side: all of this rank's pairs in one batch, one synchronisation
            for s_i, k in enumerate(mine):
                Q[s_i], Tt[s_i] = vsets[vpos[pairs[k][0]]], vsets[vpos[pairs[k][1]]]
            ctx.check(lib.ssrlcv_b200_match_batch(h, m, P, Q, Tt, None, O, n_outs))
            for hs in vsets:
                lib.ssrlcv_b200_set_free(hs)
            # the one exchange of the path: every rank ends up with every pair list, in the reference's pair order
            state["lists"] = xch.run(out, n_outs_np[:m])
            state["exchange_bytes"] = xch.last_bytes
            return int(n_outs_np[:m].sum())

        d2h = lambda: len(pairs) * 8
        par = f"pair-sharded x{world} ({len(mine)} of {len(pairs)} pairs on rank 0)"

        def parity():
            """rank 0: all 120 pairs alone, pair by pair == what the exchange delivered"""
            allv = {v: torch.from_numpy(views[v].view(np.uint8).reshape(-1)).to(dev) for v in range(V)}
            ok, total = True, 0
            got_counts, got_buf = state["lists"]
            solo = torch.zeros((n, 4), dtype=torch.int32, device=dev)
            hq, ht = C.c_void_p(), C.c_void_p()
            for k, (a, b) in enumerate(pairs):
                ctx.check(lib.ssrlcv_b200_pack(h, C.c_void_p(allv[a].data_ptr()), n, C.byref(hq)))
                ctx.check(lib.ssrlcv_b200_pack(h, C.c_void_p(allv[b].data_ptr()), n, C.byref(ht)))
                p.query_image_id, p.target_image_id = a, b
                ctx.check(lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(solo.data_ptr()), C.byref(n_out)))
                lib.ssrlcv_b200_set_free(hq); lib.ssrlcv_b200_set_free(ht)
                c = int(n_out.value)
                total += c
                ok = ok and c == int(got_counts[k]) and bool(torch.equal(solo[:c], got_buf[k][:c]))
            return ok, total
    else:
        nq, nt = w["nq"], w["nt"]
        q, t = big_pair(nq, nt, 0x55524C43)
        b = shard_bounds(nt, world)
        lo, hi = b[rank], b[rank + 1]
        qh, th = pin(q), pin(t[lo:hi])
        qd, td = qh.to(dev), th.to(dev)
        tloc = torch.from_numpy(np.ascontiguousarray(t["loc"])).to(dev)            # locations of the whole target set (8 B each)
        keys = torch.empty(nq, dtype=torch.int64, device=dev)
        allk = torch.empty((world, nq), dtype=torch.int64, device=dev)
        out = torch.empty(nq * 48, dtype=torch.uint8, device=dev)
        out_h = torch.empty(nq * 48, dtype=torch.uint8).pin_memory()
        p.out_kind, p.mask_kind, p.query_image_id, p.target_image_id = T.OUT_DMATCH, 0, 0, 1
        h2d = (nq + hi - lo) * 152
        ptrs = (C.c_void_p * 2)(qd.data_ptr(), td.data_ptr())
        sizes = (C.c_uint64 * 2)(nq, hi - lo)
        sets = (C.c_void_p * 2)()

        def step(from_host=False):
            if from_host:
                qd.copy_(qh, non_blocking=True); td.copy_(th, non_blocking=True)
            ctx.check(lib.ssrlcv_b200_pack_batch(h, 2, ptrs, sizes, sets))
            ctx.check(lib.ssrlcv_b200_match_candidates(h, C.byref(p), sets[0], sets[1], lo if hi > lo else 0, C.c_void_p(keys.data_ptr())))
            if dist is not None:
                dist.all_gather_into_tensor(allk, keys)                            # 8 B per query per rank over NVLink
                src = allk
                state["exchange_bytes"] = allk.numel() * 8
            else:
                src = keys
            ctx.check(lib.ssrlcv_b200_finalize_candidates(h, C.byref(p), sets[0], C.c_void_p(src.data_ptr()), world, C.c_void_p(tloc.data_ptr()),
                                                          None, C.c_void_p(out.data_ptr()), C.byref(n_out)))
            lib.ssrlcv_b200_set_free(sets[0]); lib.ssrlcv_b200_set_free(sets[1])
            if from_host:
                out_h[: n_out.value * 48].copy_(out[: n_out.value * 48]); torch.cuda.synchronize()
            return int(n_out.value)

        d2h = lambda: int(n_out.value) * 48 + 8
        par = f"target-sharded x{world} (rank 0 owns targets [{lo}, {hi}))"

        def parity():
            """rank 0: the whole 1 M x 1 M pair alone on its GPU == the merged result of the shards"""
            n_sh = step()
            got = out[: n_sh * 48].clone()
            tfull = torch.from_numpy(t.view(np.uint8).reshape(-1)).to(dev)
            hq, ht = C.c_void_p(), C.c_void_p()
            ctx.check(lib.ssrlcv_b200_pack(h, C.c_void_p(qd.data_ptr()), nq, C.byref(hq)))
            ctx.check(lib.ssrlcv_b200_pack(h, C.c_void_p(tfull.data_ptr()), nt, C.byref(ht)))
            solo = torch.empty(nq * 48, dtype=torch.uint8, device=dev)
            ctx.check(lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(solo.data_ptr()), C.byref(n_out)))
            lib.ssrlcv_b200_set_free(hq); lib.ssrlcv_b200_set_free(ht)
            c = int(n_out.value)
            return c == n_sh and bool(torch.equal(solo[: c * 48], got)), c

    for _ in range(max(args.warmup, 3)):
        step()
    lib.ssrlcv_b200_set_profiling(h, 1)
    sampler = ClockSampler(env.local_rank) if (with_clocks and rank == 0) else None
    launches0 = ctx.launch_count()
    total_ms, n_matches, _ = env.timed(step, args.steps)
    launches = ctx.launch_count() - launches0
    k_ms, k_n = C.c_double(0), C.c_uint64(0)
    lib.ssrlcv_b200_kernel_time(h, C.byref(k_ms), C.byref(k_n))
    lib.ssrlcv_b200_set_profiling(h, 0)
    e2e_ms = 0.0
    if with_e2e:
        step(True)
        e2e_ms, _, _ = env.timed(lambda: step(True), args.steps)
    tt = torch.tensor([total_ms, e2e_ms, float(n_matches), k_ms.value / max(k_n.value, 1)], dtype=torch.float64, device=dev)
    tm = tt.clone()
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, kern_ms_max = float(tm[0]), float(tm[1]), float(tm[3])
    n_total = int(tt[2]) if is_c4 else n_matches
    clocks = None
    if with_clocks:
        # short timed regions: the same steps, untimed, until nvidia-smi has had time to answer.  The steps contain collectives,
        # so every rank runs the same number of them (derived from the all-reduced time).
        if total_ms < 400.0:
            for _ in range(min(400, int(400.0 * args.steps / max(total_ms, 1e-3)) + 1)):
                step()
            torch.cuda.synchronize()
        if sampler is not None:
            clocks = sampler.stop()
    # ---- parity against the single-GPU result (rank 0 computes it; the others wait at the barrier)
    par_ok, par_n = None, None
    if is_c4:
        step()
    if rank == 0:
        par_ok, par_n = parity() if is_c4 else (None, None)
    if not is_c4:                      # C5's check runs one more sharded step: every rank takes part in its all-gather
        if rank == 0:
            par_ok, par_n = parity()
        else:
            step()
    env.barrier()
    if rank != 0:
        return None
    value = evals * args.steps / (total_ms * 1e-3)
    kern_ms = k_ms.value / max(k_n.value, 1)
    my_evals = evals / world if not is_c4 else len(mine) * w["n"] * w["n"]
    roof, peak_tops = roofline_of(world, my_evals, kern_ms, k_n.value, args.steps, ", rank 0")
    roof["traffic"] = None
    roof["kernel_ms_max_over_ranks"] = kern_ms_max
    line = {
        "metric": "sift_dist_evals_per_s", "value": value, "unit": "dist-evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(name, w), "parallelism": par,
        "step": "pack + match of this rank's share + result exchange + finalize, device-resident inputs",
        "matches_per_s": n_total * args.steps / (total_ms * 1e-3), "matches_per_step": n_total,
        "int8_tops": 256.0 * value / 1e12, "frac_of_int8_peak_step": 256.0 * value / 1e12 / (world * peak_tops),
        "frac_of_nominal_int8_step": 256.0 * value / 1e12 / (world * 4500.0), "clocks": clocks,
        "exchange_bytes_per_step": int(state["exchange_bytes"]),
        "parity": par_ok, "parity_note": f"sharded records == the single-GPU result recomputed on rank 0 ({par_n} records), byte for byte",
        "e2e": ({"value": evals * args.steps / (e2e_ms * 1e-3), "unit": "dist-evals/s", "ms_per_step": e2e_ms / args.steps,
                 "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h(), "api": "C ABI with pinned host inputs copied in every step, counts / records read back"}
                if with_e2e else None),
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": None,
    }
    if world == 1 and not args.no_cpu_baseline and with_e2e:
        # the CPU oracle on a bounded sample of the same workload: C4 = a query sample of ONE of its 120 equal pairs,
        # C5 = a query sample against the whole target set (dist-evals/s does not depend on which pair / which queries)
        try:
            if env.prev_affinity:
                os.sched_setaffinity(0, env.prev_affinity)
            cq, ct = (views[0], views[1]) if is_c4 else (q, t)
            cb = cpu_baseline(dict(w, mask=0), cq, ct, None)
            if is_c4:
                cb["sample"] += f" (one pair of the {len(pairs)})"
            line["cpu_baseline"] = cb
        except Exception as e:  # the oracle library is test infrastructure; its absence must not hide the GPU numbers
            line["cpu_baseline"] = {"unavailable": str(e)}
    return line


def run_c4_one_process(args, w, env):
    """C4 through ssrlcv_b200_match_pairs_multi_gpu: ONE process (rank 0) drives every GPU of the job through the C ABI's own
    pair scheduler (one context + host thread per device) -- the route the C++ drop-in's generateMatchesExhaustive takes.
    View v lives on GPU v mod N (a multi-GPU feature-extraction stage leaves them spread like that); every pair list is
    delivered to GPU 0.  The other ranks wait on the process group's store (host side: no NCCL kernel occupies their GPUs).
    Timed per device with CUDA events inside the library (max over devices) and by the host clock around the call."""
    torch = env.torch
    from ssrlcv_b200 import _lib, types as T
    from ssrlcv_b200.match_factory import MultiGpu
    from ssrlcv_b200.pair_scheduler import pair_order
    from tests import synth
    rank, world = env.rank, env.world
    store = None
    if env.dist is not None:
        from torch.distributed import distributed_c10d
        store = distributed_c10d._get_default_store()
    key = f"c4_one_process_done_{args.steps}"
    if rank != 0:
        env.torch.cuda.synchronize()
        store.wait([key])
        return None
    line = None
    mask_here = os.sched_getaffinity(0)
    if env.prev_affinity:
        os.sched_setaffinity(0, env.prev_affinity)   # this ONE process drives every GPU with a host thread per device: all the CPUs, not rank 0's slice
    try:
        V, n = w["views"], w["n"]
        views = []
        for k in range(V // 2):
            a, b = synth.make_pair(n, n, seed=0x55524C43 + k)
            views += [a, b]
        pairs = pair_order(V)
        mg = MultiGpu.get(world)
        lib = mg.lib
        G = mg.device_count()
        dev_views = [torch.from_numpy(views[v].view(np.uint8).reshape(-1)).to(f"cuda:{v % G}") for v in range(V)]
        p = _lib.MatchParams()
        p.absolute_threshold, p.relative_threshold, p.compact, p.reserved = ABS_THR, REL_THR, 1, 0
        p.out_kind, p.mask_kind = T.OUT_PAIR, 0
        m = len(pairs)
        P = (_lib.MatchParams * m)()
        for i, (a, b) in enumerate(pairs):
            p.query_image_id, p.target_image_id = a, b
            C.memmove(C.byref(P, i * C.sizeof(_lib.MatchParams)), C.byref(p), C.sizeof(_lib.MatchParams))
        F = (C.c_void_p * V)(*[t.data_ptr() for t in dev_views])
        N = (C.c_uint64 * V)(*([n] * V))
        pq = (C.c_uint32 * m)(*[a for a, _ in pairs])
        pt = (C.c_uint32 * m)(*[b for _, b in pairs])
        out = torch.zeros((m, n, 4), dtype=torch.int32, device="cuda:0")
        O = (C.c_void_p * m)(*[out[i].data_ptr() for i in range(m)])
        n_out = (C.c_uint64 * m)()
        for d in range(G):
            torch.cuda.synchronize(d)

        def step():
            rc = lib.ssrlcv_b200_match_pairs_multi_gpu(mg.handle, V, F, N, None, 0, m, pq, pt, P, O, n_out)
            if rc != 0:
                raise RuntimeError(lib.ssrlcv_b200_multi_last_error(mg.handle))

        for _ in range(max(args.warmup, 3)):
            step()
        flushes = [torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{d}") for d in range(G)]
        dev_ms, wall_ms = [], []
        for i in range(args.steps):
            for d in range(G):
                flushes[d].fill_(i & 0xFF)
            for d in range(G):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            step()
            wall_ms.append((time.perf_counter() - t0) * 1e3)
            dev_ms.append(max(lib.ssrlcv_b200_multi_device_ms(mg.handle, d) for d in range(G)))
        counts = np.ctypeslib.as_array(n_out).copy()
        # parity: every pair list == the single-GPU call on GPU 0
        ctx = env.ctx
        h = ctx.handle
        allv = {v: (dev_views[v] if v % G == 0 else dev_views[v].to("cuda:0")) for v in range(V)}
        solo = torch.zeros((n, 4), dtype=torch.int32, device="cuda:0")
        ok, total, one = True, 0, C.c_uint64(0)
        hq, ht = C.c_void_p(), C.c_void_p()
        for k, (a, b) in enumerate(pairs):
            ctx.check(ctx.lib.ssrlcv_b200_pack(h, C.c_void_p(allv[a].data_ptr()), n, C.byref(hq)))
            ctx.check(ctx.lib.ssrlcv_b200_pack(h, C.c_void_p(allv[b].data_ptr()), n, C.byref(ht)))
            p.query_image_id, p.target_image_id = a, b
            ctx.check(ctx.lib.ssrlcv_b200_match(h, C.byref(p), hq, ht, None, 0, C.c_void_p(solo.data_ptr()), C.byref(one)))
            ctx.lib.ssrlcv_b200_set_free(hq); ctx.lib.ssrlcv_b200_set_free(ht)
            c = int(one.value)
            total += c
            ok = ok and c == int(counts[k]) and bool(torch.equal(solo[:c], out[k][:c]))
        evals = evals_per_step(w)
        t_dev, t_wall = sum(dev_ms) / len(dev_ms), sum(wall_ms) / len(wall_ms)
        line = {"api": "ssrlcv_b200_match_pairs_multi_gpu (one process, one context + host thread per device)", "n_gpus": G,
                "ms_per_step": t_dev, "wall_ms_per_step": t_wall, "scaling": "strong",
                "value": evals / (t_dev * 1e-3), "value_wall": evals / (t_wall * 1e-3), "unit": "dist-evals/s",
                "timing": "ms_per_step = max over devices of the CUDA-event time of each device's share (fetch ... delivery), mean over steps; "
                          "wall_ms_per_step = host clock around the call (includes thread wake-up and the final joins)",
                "inputs": f"view v resident on GPU v mod {G}; pair lists delivered to GPU 0", "matches_per_step": int(counts.sum()),
                "parity": ok, "parity_note": f"every pair list == the single-GPU call on GPU 0 ({total} records)"}
    except Exception as e:  # noqa: BLE001
        line = {"error": repr(e)}
    finally:
        os.sched_setaffinity(0, mask_here)
        if store is not None:
            store.set(key, "1")
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the C4 / C5 sub-records of the default line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        line = run_reference(args, w, rank, world)
    else:
        env = Env(rank, world, local_rank)
        if args.workload in ("C4", "C5"):
            line = run_sharded(args, args.workload, w, env)
        else:
            line = run_ours(args, w, env)
            if world > 1 and args.workload == "C3" and not args.no_sharded:
                # The multi-GPU configurations of BASELINE.json, measured by the same launch: C4 (120 pairs dealt to the ranks)
                # and C5 (one 1 M x 1 M pair, target set sharded, NCCL all-gather of the candidate keys), strong scaling,
                # each with its record-equality check against the single-GPU result.
                sub = {}
                sub_args = argparse.Namespace(**vars(args))
                sub_args.steps = min(args.steps, 10)
                for name in ("C4", "C5"):
                    env.torch.cuda.empty_cache()
                    try:
                        r = run_sharded(sub_args, name, WORKLOADS[name], env, with_clocks=False, with_e2e=False)
                    except Exception as e:  # noqa: BLE001 -- the headline line must survive a failure of a sub-record
                        r = {"error": repr(e)}
                    if rank == 0:
                        sub[name] = r
                env.torch.cuda.empty_cache()
                r = run_c4_one_process(sub_args, WORKLOADS["C4"], env)
                if rank == 0:
                    sub["C4_one_process"] = r
                if line is not None:
                    line["sharded"] = sub
        env.close()
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
