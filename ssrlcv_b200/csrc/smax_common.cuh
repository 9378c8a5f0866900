// smax_common.cuh -- device helpers shared by the two tensor-core matcher kernels (smax_tc.cu: one query tile per unit, every
// mask kind; smax_pair.cu: two query tiles per unit, unmasked searches): the max trees, the per-query window state and its
// bookkeeping, the resolvers that turn recorded windows into the reference's winner keys, and the unit iterator.
#pragma once
#include "geometry.cuh"
#include "matcher.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr uint32_t kNoWin = 0xFFFFFFFFu;

__device__ __forceinline__ int32_t max3(int32_t a, int32_t b, int32_t c) { return max(max(a, b), c); }

// maximum of 32 values as a depth-4 tree of 3-input max (VIMNMX3): 17 instructions, no long dependency chain
__device__ __forceinline__ int32_t max32(const uint32_t (&x)[32]) {
  int32_t t[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) t[i] = max3(static_cast<int32_t>(x[3 * i]), static_cast<int32_t>(x[3 * i + 1]), static_cast<int32_t>(x[3 * i + 2]));
  t[10] = max(static_cast<int32_t>(x[30]), static_cast<int32_t>(x[31]));
  const int32_t u0 = max3(t[0], t[1], t[2]), u1 = max3(t[3], t[4], t[5]), u2 = max3(t[6], t[7], t[8]), u3 = max(t[9], t[10]);
  return max(max3(u0, u1, u2), u3);
}

// per-query running state of the window search: best score and up to two (window, column mask) records that reached it.
// `best` starts at (lowest score that can still pass the absolute threshold) - 1, so with a finite threshold
// (the reference's MatchFactory(0.6, 200*200), src/Pipeline.cu:175) almost no tile ever touches the bookkeeping;
// w1 == kNoWin means "no acceptable window yet".
struct WinState { int32_t best; uint32_t w1, m1, w2, m2, ovf; };
// branch-free update with one window maximum and the columns of the window that attained it
__device__ __forceinline__ void win_update(WinState& s, int32_t m, uint32_t win, uint32_t mask) {
  const bool gt = m > s.best;
  const bool eq = (m == s.best) & (s.w1 != kNoWin);
  const bool second = eq & (s.w2 == kNoWin);
  s.ovf = gt ? 0u : ((eq & !second) ? 1u : s.ovf);
  s.w2 = gt ? kNoWin : (second ? win : s.w2);
  s.m2 = second ? mask : s.m2;
  s.w1 = gt ? win : s.w1;
  s.m1 = gt ? mask : s.m1;
  s.best = max(s.best, m);
}
__device__ __forceinline__ WinState win_merge(const WinState& a, const WinState& b) {
  if (a.w1 == kNoWin) return b;
  if (b.w1 == kNoWin) return a;
  if (a.best > b.best) return a;
  if (b.best > a.best) return b;
  WinState r = a;                       // equal scores: union of the windows, at most two survive
  r.ovf = a.ovf | b.ovf;
  if (a.w2 == kNoWin) r.w2 = b.w1; else r.ovf = 1;
  if (b.w2 != kNoWin) r.ovf = 1;
  return r;
}
// Which columns of a window attained its maximum is recorded (kColMask) only where that is cheap: thresholded searches
// (almost no window reaches the floor) and short target sets (where re-reading 4 KB of descriptors per query and unit
// would cost more than the unit's tensor work).  A long unthresholded search updates some lane's running best in ~14 %
// of its windows; there only the window is recorded and the one or two surviving windows are re-evaluated afterwards.
// bit e set <=> x[e] == m
__device__ __forceinline__ uint32_t eq_mask32(const uint32_t (&x)[32], int32_t m) {
  uint32_t mask = 0;
#pragma unroll
  for (int e = 0; e < 32; ++e) mask |= (static_cast<int32_t>(x[e]) == m) ? (1u << e) : 0u;
  return mask;
}

// The rare path's view of two job fields it alone needs: volatile loads, so that the compiler cannot hoist them into registers
// that would live across the matcher's hot loop (the job lives in device memory or in the kernel's parameter block).
__device__ __forceinline__ const uint64_t* load_exclude_ptr(const SmaxArgs& job) {
  unsigned long long p;
  asm volatile("ld.u64 %0, [%1];" : "=l"(p) : "l"(&job.q_exclude) : "memory");
  return reinterpret_cast<const uint64_t*>(p);
}
__device__ __forceinline__ uint32_t excluded_row(const uint64_t* q_exclude, uint32_t row) { return static_cast<uint32_t>(__ldg(q_exclude + row)); }
__device__ __forceinline__ uint32_t load_nq(const SmaxArgs& job) {
  uint32_t v;
  asm volatile("ld.u32 %0, [%1];" : "=r"(v) : "l"(&job.nq) : "memory");
  return v;
}
// Runner-up searches: the excluded target of this lane's query (row `ex` of the target set) scores 0 in the windows that could
// matter.  A window holding a query's excluded WINNER always reaches that query's running best, so the rare path -- the only
// caller -- sees every case; the common path is untouched.  x holds exact scores here (fold, or W already added).
template <int kWin>
__device__ __forceinline__ void exclude_columns(uint32_t (&v)[kWin][32], int32_t (&wm)[kWin], int32_t best, uint32_t win0, uint32_t ex) {
#pragma unroll
  for (int k = 0; k < kWin; ++k) {
    const bool hit = ((ex >> 5) == win0 + k) && (wm[k] >= best);
    if (!__any_sync(0xffffffffu, hit)) continue;
    uint32_t(&x)[32] = v[k];
    const uint32_t col = ex & 31u;
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = (hit && col == static_cast<uint32_t>(e)) ? 0u : x[e];
    const int32_t m = max32(x);
    wm[k] = hit ? m : wm[k];
  }
}

// kWin 32-target windows of one query row, in registers (unmasked searches), in two steps: the window maxima (the common path:
// max trees and nothing else) and -- only when some lane's maximum reaches its running best -- the bookkeeping.
// Fold: the maxima ARE the windows' best scores.  No-fold: the accumulator holds a.b only; max(a.b) + wb[k] (the largest W of
// window k, pack.cu wmax) bounds every score of the window.
template <int kWin, bool kFold>
__device__ __forceinline__ int32_t window_maxima(const uint32_t (&v)[kWin][32], const int32_t (&wb)[kWin], int32_t (&wm)[kWin], uint32_t diag) {
  if (kDiagBuild && (diag & 4u)) {
#pragma unroll
    for (int k = 0; k < kWin; ++k) wm[k] = static_cast<int32_t>((v[k][0] | v[k][31]) >> 24);
  } else {
#pragma unroll
    for (int k = 0; k < kWin; ++k) wm[k] = max32(v[k]) + (kFold ? 0 : wb[k]);
  }
  int32_t mt = wm[0];
#pragma unroll
  for (int k = 1; k < kWin; ++k) mt = max(mt, wm[k]);
  return mt;
}
// The rare path (all lanes of the warp take it together).  No-fold: only windows whose bound reaches some lane's best get W
// added (s_w: shared-window address of the first window's 32 W values) and an exact maximum; the others stay bounds below every
// lane's best.  win0: global index of the first window.  kSplit (smax_duo.cu, 128-column accumulators fed by both CTAs of a pair):
// the four windows are two runs of two, four windows apart -- k -> win0 + (k >> 1) * 4 + (k & 1).
template <int kWin, bool kFold, bool kColMask, bool kSplit = false>
__device__ __forceinline__ void window_bookkeeping(uint32_t (&v)[kWin][32], int32_t (&wm)[kWin], WinState& st, uint32_t win0, uint32_t s_w,
                                                   bool has_exclude, const SmaxArgs& job, uint32_t row) {
  auto wofs = [](int k) { return kSplit ? static_cast<uint32_t>((k >> 1) * 4 + (k & 1)) : static_cast<uint32_t>(k); };
  if (!kFold) {
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      if (!__any_sync(0xffffffffu, wm[k] >= st.best)) continue;
      uint32_t(&x)[32] = v[k];
      const uint32_t wp = s_w + wofs(k) * (kChunk * 4);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const int4 w4 = ld_shared_v4(wp + 4 * e);
        x[e] += static_cast<uint32_t>(w4.x); x[e + 1] += static_cast<uint32_t>(w4.y);
        x[e + 2] += static_cast<uint32_t>(w4.z); x[e + 3] += static_cast<uint32_t>(w4.w);
      }
      wm[k] = max32(x);
    }
  }
  if (has_exclude) {                                         // runner-up search (launch-uniform)
    const uint64_t* q_exclude = load_exclude_ptr(job);
    if (q_exclude != nullptr) exclude_columns<kWin>(v, wm, st.best, win0, row < load_nq(job) ? excluded_row(q_exclude, row) : 0xFFFFFFFFu);
  }
  if (kColMask) {
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      if (__any_sync(0xffffffffu, wm[k] >= st.best))         // which columns attained it: only for windows that matter
        win_update(st, wm[k], win0 + wofs(k), eq_mask32(v[k], wm[k]));
    }
  } else {
#pragma unroll
    for (int k = 0; k < kWin; ++k) win_update(st, wm[k], win0 + wofs(k), 0u);
  }
}

// Where query row `row` of the (possibly spatially sorted) query arrays keeps its key: the caller's row.  Padding rows of the
// last tile keep their own (unused) row.
__device__ __forceinline__ uint32_t key_row(const uint32_t* q_orig, uint32_t nq, uint32_t row) {
  return (q_orig != nullptr && row < nq) ? __ldg(q_orig + row) : row;
}
// The live target tiles of a unit, in ascending order.  Without a dead-tile bitmap: every tile of [j0, j1).  With one (masked
// searches on spatially sorted sets, cull_kernel): the set bits of row qt, one 32-tile word at a time, the next word requested
// before the current one is consumed -- one load latency per 32 tiles and nothing per dead tile (on the reference's fixture only
// 5-30 % of a CTA's tiles are live).  Every role walks the same sequence.
struct LiveTiles {
  const uint32_t* row;     // nullptr: all live
  uint32_t j, j1;          // all live: next tile / end
  uint32_t base, bits, ahead;
  __device__ __forceinline__ uint32_t word(uint32_t b) const {          // tiles [b, b + 32) of the row, limited to [.., j1)
    if (b >= j1) return 0u;
    const uint32_t w = __ldg(row + (b >> 5));
    return (j1 - b >= 32u) ? w : (w & ((1u << (j1 - b)) - 1u));
  }
  __device__ __forceinline__ LiveTiles(const uint32_t* live_bits, uint32_t stride, uint32_t qt, uint32_t j0, uint32_t j1_) {
    row = live_bits ? live_bits + static_cast<uint64_t>(qt) * stride : nullptr;
    j = j0; j1 = j1_;
    base = j0 & ~31u; bits = 0; ahead = 0;
    if (row != nullptr) {
      bits = word(base) & (~0u << (j0 & 31u));
      ahead = word(base + 32u);
    }
  }
  __device__ __forceinline__ bool next(uint32_t& out) {
    if (row == nullptr) {
      if (j >= j1) return false;
      out = j++;
      return true;
    }
    while (bits == 0u) {
      base += 32u;
      if (base >= j1) return false;
      bits = ahead;
      ahead = word(base + 32u);
    }
    out = base + (__ffs(bits) - 1);
    bits &= bits - 1u;
    return true;
  }
};

// Turns the recorded states of query row `row_base + r` (a, b: the states of the warps that shared the row's columns; an empty
// state -- w1 == kNoWin -- where only one warp saw them) into the reference's winner key.  Every recorded column holds exactly the score `top`, so its distance follows from
// the score and the parity of the target's norm; the minimum key over the recorded columns is the winner.
__device__ __forceinline__ void resolve_row(const SmaxArgs& job, const WinState& a, const WinState& b, uint32_t row_base, uint32_t slot, uint32_t r,
                                            int32_t w_offset) {
  const uint32_t row = row_base + r;                    // < nq_pad: rows beyond nq land in the padding of the key array
  const bool va = a.w1 != kNoWin, vb = b.w1 != kNoWin;
  uint64_t key = kKeyNone;
  if (va | vb) {
    const int32_t top = va ? (vb ? max(a.best, b.best) : a.best) : b.best;
    const int32_t base = __ldg(job.q_norm + row) - 2 * top + 2 * w_offset;
    const int32_t* t_norm = job.t_norm;
    const uint32_t* t_orig = job.t_orig;
    const uint64_t off = job.target_offset;
    bool ovf = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const WinState& s = h ? b : a;
      if (s.w1 == kNoWin || s.best != top) continue;
      ovf |= s.ovf != 0;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t w = i ? s.w2 : s.w1;
        uint32_t m = i ? s.m2 : s.m1;
        if (w == kNoWin) continue;
        while (m) {
          const uint32_t e = __ffs(m) - 1;
          m &= m - 1;
          const uint64_t f = static_cast<uint64_t>(w) * kChunk + e;
          const uint32_t d2 = static_cast<uint32_t>(base + (__ldg(t_norm + f) & 1));
          const uint64_t k = make_key(d2, (t_orig ? static_cast<uint64_t>(__ldg(t_orig + f)) : f) + off);
          key = k < key ? k : key;
        }
      }
    }
    if (ovf) {                                          // three or more windows tied: exact SIMT kernel
      key = kKeyNeedsExact;
      if (row < job.nq) {
        const uint32_t at = atomicAdd(job.n_flagged, 1u);
        job.flagged[at] = slot * job.nq_pad + row;
      }
    }
  }
  job.keys[static_cast<uint64_t>(slot) * job.nq_pad + key_row(job.q_orig, job.nq, row)] = key;
}

// Key slots [slot + 1, fill_end) of a query tile whose last unit this is have no unit of their own (the tile is crossed by
// fewer CTAs than the widest one): mark them "no candidate" for rows [r_begin, r_begin + r_count).
__device__ __forceinline__ void fill_unused_slots(const SmaxArgs& job, uint32_t row_base, uint32_t slot, uint32_t fill_end, uint32_t r_begin,
                                                  uint32_t r_count, int lane) {
  if (slot + 1 >= fill_end) return;
  uint64_t* keys = job.keys;
  const uint64_t stride = job.nq_pad;
  const uint32_t* q_orig = job.q_orig;
  const uint32_t nq = job.nq;
  for (uint32_t s = slot + 1; s < fill_end; ++s)
    for (uint32_t r = lane; r < r_count; r += 32) keys[s * stride + key_row(q_orig, nq, row_base + r_begin + r)] = kKeyNone;
}

__device__ __forceinline__ uint32_t word_of(const uint4 (&v)[8], int w) {
  const uint4 q = v[w >> 2];
  return (w & 3) == 0 ? q.x : (w & 3) == 1 ? q.y : (w & 3) == 2 ? q.z : q.w;
}

// Window regime (no column masks were recorded, see smax_kernel): exact re-evaluation of the recorded windows.  One warp,
// kFly queries in flight: lane l owns target l of each query's window (32 lanes = the 32 targets of the window = the 32 lanes of the
// reference's block, ref: src/MatchFactory.cu:1476-1492) and holds its 128 descriptor bytes in registers; the query
// descriptor is spread one 32-bit word per lane and broadcast by shuffle.  Exact integer d2 by dp4a, key
// (d2, f mod 32, f), warp-wide minimum.  Rows row_base + [r_begin, r_begin + r_count) of the unit, r_count % kFly == 0;
// xb == nullptr: one state per row (smax_pair.cu).
template <int kMask, int kFly>
__device__ __forceinline__ void resolve_rows_windows(const SmaxArgs& job, const WinState* xa, const WinState* xb,
                                             uint32_t row_base, uint32_t slot, uint32_t r_begin, uint32_t r_count, int lane) {
  constexpr uint32_t kFull = 0xffffffffu;
  // `job` lives in device memory: take the fields once (the stores below could alias it as far as the compiler knows)
  struct {
    uint32_t nq, nt, nq_pad; float eps; uint64_t target_offset;
    const uint8_t* q_desc; const int32_t* q_norm; const float4* q_mask;
    const uint8_t* t_desc; const int32_t* t_norm; const float2* t_loc; const uint32_t* t_orig; const uint32_t* q_orig;
    uint64_t* keys; uint32_t* flagged; uint32_t* n_flagged; const uint64_t* q_exclude;
  } args = {job.nq, job.nt, job.nq_pad, job.eps, job.target_offset, job.q_desc, job.q_norm, job.q_mask,
            job.t_desc, job.t_norm, job.t_loc, job.t_orig, job.q_orig, job.keys, job.flagged, job.n_flagged, job.q_exclude};
  for (uint32_t g = 0; g < r_count; g += kFly) {
    WinState ms[kFly];
    uint32_t qrow[kFly];
    bool any = false;
#pragma unroll
    for (int k = 0; k < kFly; ++k) {
      const uint32_t r = r_begin + g + k;
      ms[k] = xb != nullptr ? win_merge(xa[r], xb[r]) : xa[r];
      qrow[k] = min(row_base + r, args.nq_pad - 1);
      any |= ms[k].w1 != kNoWin;
    }
    uint64_t res[kFly];
#pragma unroll
    for (int k = 0; k < kFly; ++k) res[k] = kKeyNone;
    if (any) {                                                   // warp-uniform: every lane read the same states
      uint4 tv[kFly][8];
      uint32_t qw[kFly];
      int32_t tn[kFly], qn[kFly];
      uint64_t fs[kFly];
      bool ok[kFly];
#pragma unroll
      for (int k = 0; k < kFly; ++k) {                              // all loads first
        const uint32_t w = ms[k].w1 == kNoWin ? 0u : ms[k].w1;
        const uint64_t f = static_cast<uint64_t>(w) * kChunk + lane;
        ok[k] = (f < args.nt) & (ms[k].w1 != kNoWin);
        if (args.q_exclude != nullptr && qrow[k] < args.nq && f == excluded_row(args.q_exclude, qrow[k])) ok[k] = false;   // runner-up search
        fs[k] = ok[k] ? f : 0;
        const uint4* td = reinterpret_cast<const uint4*>(args.t_desc + fs[k] * kDescBytes);
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tv[k][w8] = __ldg(td + w8);
        qw[k] = __ldg(reinterpret_cast<const uint32_t*>(args.q_desc + static_cast<uint64_t>(qrow[k]) * kDescBytes) + lane);
        tn[k] = __ldg(args.t_norm + fs[k]);
        qn[k] = __ldg(args.q_norm + qrow[k]);
      }
#pragma unroll
      for (int k = 0; k < kFly; ++k) {
        if (kMask != SSRLCV_B200_MASK_NONE) {
          const QueryMask qm = load_query_mask<kMask>(args.q_mask + 2 * static_cast<uint64_t>(qrow[k]));
          const float2 tl = args.t_loc[fs[k]];
          if (mask_skips<kMask>(qm, tl.x, tl.y, args.eps)) ok[k] = false;
        }
        uint32_t dot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) dot = __dp4a(__shfl_sync(kFull, qw[k], w), word_of(tv[k], w), dot);
        const uint32_t d2 = static_cast<uint32_t>(qn[k] + tn[k] - 2 * static_cast<int32_t>(dot));
        res[k] = ok[k] ? make_key(d2, (args.t_orig ? static_cast<uint64_t>(__ldg(args.t_orig + fs[k])) : fs[k]) + args.target_offset) : kKeyNone;
      }
      // ---- rare: a second window reached the same score ----
#pragma unroll
      for (int k = 0; k < kFly; ++k) {
        if (ms[k].w1 != kNoWin && ms[k].w2 != kNoWin) {          // warp-uniform
          const uint64_t f = static_cast<uint64_t>(ms[k].w2) * kChunk + lane;
          bool ok2 = f < args.nt;
          if (args.q_exclude != nullptr && qrow[k] < args.nq && f == excluded_row(args.q_exclude, qrow[k])) ok2 = false;
          const uint64_t f2 = ok2 ? f : 0;
          if (kMask != SSRLCV_B200_MASK_NONE) {
            const QueryMask qm = load_query_mask<kMask>(args.q_mask + 2 * static_cast<uint64_t>(qrow[k]));
            const float2 tl = args.t_loc[f2];
            if (mask_skips<kMask>(qm, tl.x, tl.y, args.eps)) ok2 = false;
          }
          const uint4* td = reinterpret_cast<const uint4*>(args.t_desc + f2 * kDescBytes);
          uint32_t dot = 0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) {
            const uint4 b4 = __ldg(td + w8);
            dot = __dp4a(__shfl_sync(kFull, qw[k], 4 * w8 + 0), b4.x, dot);
            dot = __dp4a(__shfl_sync(kFull, qw[k], 4 * w8 + 1), b4.y, dot);
            dot = __dp4a(__shfl_sync(kFull, qw[k], 4 * w8 + 2), b4.z, dot);
            dot = __dp4a(__shfl_sync(kFull, qw[k], 4 * w8 + 3), b4.w, dot);
          }
          const uint32_t d2 = static_cast<uint32_t>(qn[k] + __ldg(args.t_norm + f2) - 2 * static_cast<int32_t>(dot));
          const uint64_t k2 = ok2 ? make_key(d2, (args.t_orig ? static_cast<uint64_t>(__ldg(args.t_orig + f2)) : f2) + args.target_offset) : kKeyNone;
          res[k] = k2 < res[k] ? k2 : res[k];
        }
      }
      // ---- warp-wide minima of the queries in flight, interleaved ----
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
        for (int k = 0; k < kFly; ++k) {
          const uint32_t lo = __shfl_xor_sync(kFull, static_cast<uint32_t>(res[k]), o);
          const uint32_t hi = __shfl_xor_sync(kFull, static_cast<uint32_t>(res[k] >> 32), o);
          const uint64_t other = (static_cast<uint64_t>(hi) << 32) | lo;
          res[k] = other < res[k] ? other : res[k];
        }
      }
#pragma unroll
      for (int k = 0; k < kFly; ++k) {
        if (ms[k].w1 == kNoWin) res[k] = kKeyNone;
        else if (ms[k].ovf) {                                    // three or more windows tied: exact SIMT kernel
          res[k] = kKeyNeedsExact;
          const uint32_t row = row_base + r_begin + g + k;
          if (lane == 0 && row < args.nq) {
            const uint32_t at = atomicAdd(args.n_flagged, 1u);
            args.flagged[at] = slot * args.nq_pad + row;
          }
        }
      }
    }
    // rows beyond nq land in the padding of the key array
    if (lane < kFly) {
      uint64_t mine = res[0];
#pragma unroll
      for (int k = 1; k < kFly; ++k) mine = lane == k ? res[k] : mine;
      args.keys[static_cast<uint64_t>(slot) * args.nq_pad + key_row(args.q_orig, args.nq, row_base + r_begin + g + lane)] = mine;
    }
  }
}

// Iterates the units of one CTA: the parts of its contiguous range of the launch's linear tile space that fall inside
// one query tile of one job.  The jobs' tile spaces are laid end to end (SmaxArgs::tile_begin / tile_end).
struct UnitIter {
  uint64_t L, L_end;                           // linear tile indices of the launch
  const SmaxJob* J;                            // job of the current unit
  uint32_t job;
  uint32_t qt, j0, j1, slot;
  uint32_t fill_end;                           // key slots [slot + 1, fill_end) of this query tile have no unit: the unit that ends the
                                               // tile's target range marks them "no candidate" (nothing is memset beforehand)
  bool last;                                   // the CTA's final unit
  const SmaxJob* jobs;                         // device memory, or the launch's parameter block (single pair)
  __device__ __forceinline__ UnitIter(const SmaxBatch& b, uint32_t cta) {
    jobs = b.jobs ? b.jobs : &b.inline_job;
    L = cta_tile_begin(b.total_tiles, b.n_ctas, cta);
    L_end = cta_tile_begin(b.total_tiles, b.n_ctas, cta + 1);
    uint32_t lo = 0, hi = b.n_jobs - 1;        // the job whose range contains L (every job of a launch has tiles)
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (L >= jobs[mid].a.tile_end) lo = mid + 1; else hi = mid;
    }
    job = lo;
    J = jobs + job;
  }
  // check_fold: the launch is the fold variant; a job whose target set turned out not to fit the fold digits (the host
  // launched optimistically, before the set's pack had reported back) is skipped by every role alike and redone by the
  // host with the other variant.
  __device__ __forceinline__ bool next(const SmaxBatch& b, uint32_t cta, bool check_fold) {
    for (;;) {
      if (L >= L_end) return false;
      while (L >= jobs[job].a.tile_end) ++job;
      J = jobs + job;
      if (!check_fold || J->a.t_meta[0] == 0) break;
      L = J->a.tile_end;
    }
    const uint32_t n_tt = J->a.n_ttiles;
    const uint64_t begin = J->a.tile_begin;
    const uint32_t rel = static_cast<uint32_t>(L - begin);         // a job has fewer than 2^32 tiles
    qt = rel / n_tt;
    j0 = rel - qt * n_tt;
    const uint64_t room = L_end - L;
    j1 = (n_tt - j0 <= room) ? n_tt : j0 + static_cast<uint32_t>(room);
    slot = cta - cta_of_tile(b.total_tiles, b.n_ctas, begin + static_cast<uint64_t>(qt) * n_tt);
    fill_end = j1 == n_tt ? J->a.n_splits : slot + 1;
    L += j1 - j0;
    last = L >= L_end;
    return true;
  }
};

}  // namespace b200
