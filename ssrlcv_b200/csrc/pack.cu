// pack.cu -- repack Feature<SIFT_Descriptor>[n] (152-byte AoS, descriptor at byte 24, ref:
// include/Feature.cuh:30-38,68-94) into the layout the tensor pipeline wants:
//   desc  uint8[n_pad][128]   one 128-byte row per feature (TMA box rows, SWIZZLE_128B)
//   ext   uint8[n_pad][32]    base-255 digits of W = kFoldC - (|b|^2 >> 1) (extra K block, SWIZZLE_32B)
//   norm  int32[n_pad]        |b|^2            wcol int32[n_pad]   kNoFoldC - (|b|^2 >> 1)
//   wmax  int32[n_pad / 32]   largest wcol of each 32-row window (0 for windows of padding only): the bound the matcher's
//                             no-fold variant tests before it adds W to a window's scores
//   loc   float2[n_pad]       sigtheta float2[n_pad]
// HBM-bound byte shuffling: 152 B read, 188 B written per feature (+ 4 B per 32 features).
#include <atomic>

#include <cub/cub.cuh>

#include "matcher.cuh"
#include "ptx.cuh"

namespace b200 {

// Per-set device scalars ("meta", 4 x int32 from the context's zero-initialised slot pool, api.cu):
//   meta[0] != 0  <=>  some target's W does not fit the fold digits (the matcher then runs its epilogue-add variant)
// W is measured against a CONSTANT offset (kFoldC / kNoFoldC, matcher.cuh), not against the set's own maximum norm, so
// the digits of a row depend on that row alone: one kernel, no grid-wide dependency, no host round trip.  The (rare) bad
// verdict is also stored straight into the set's pinned, device-mapped host word, so no copy node follows the kernel.
//
// One launch packs up to kPackInline sets (the job table travels as a kernel parameter: no staging copy).  The unit of work
// is a 32-row window -- the matcher's candidate window -- handled by ONE warp in 16 steps of two rows (16 lanes x 8 bytes
// cover one descriptor; the record base is only 8-byte aligned), so the window's largest W is a register maximum: no atomics
// and nothing to zero beforehand.  Rows [n, n_rows) are the zero padding of the set and are written here too.
__global__ void __launch_bounds__(256) pack_kernel(const __grid_constant__ PackBatch b) {
  pdl_wait();                  // the rows may be recycled memory a previous kernel of the stream still reads
  pdl_launch_dependents();     // the matcher behind this launch sets itself up while the windows are packed
  const int lane = threadIdx.x & 31, half = lane >> 4, sub = lane & 15;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t w = warp_global; w < b.total_windows; w += n_warps) {
    uint32_t j = 0;
    while (j + 1 < b.n_jobs && w >= b.jobs[j + 1].win_begin) ++j;
    const RefFeature* __restrict__ in = b.jobs[j].in;
    const uint64_t n = b.jobs[j].n;
    const uint32_t win = w - b.jobs[j].win_begin;
    uint8_t* __restrict__ desc = b.jobs[j].desc;
    uint8_t* __restrict__ ext = b.jobs[j].ext;
    const uint64_t row0 = static_cast<uint64_t>(win) * kChunk;
    // every load of the window first: 16 descriptor slices per lane (rows row0 + 2k + half) and the scalars of "its own" row
    // row0 + 2 * sub + half, the row whose norm this lane will keep
    const bool commit = in == nullptr;                                // rows [0, n) of desc / loc / sigtheta are already in place
    uint2 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint64_t i = row0 + 2 * k + half;
      v[k] = make_uint2(0u, 0u);
      if (i < n)
        v[k] = commit ? *reinterpret_cast<const uint2*>(desc + i * kDescBytes + sub * 8)
                      : *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(in + i) + 24 + sub * 8);
    }
    const uint64_t own = row0 + 2 * sub + half;
    uint2 own_loc = make_uint2(0u, 0u), own_st = make_uint2(0u, 0u);
    if (own < n && !commit) {
      own_loc = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(in + own) + 8);
      own_st = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(in + own) + 16);
    }
    uint32_t own_s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint64_t i = row0 + 2 * k + half;                         // < n_rows: windows are whole
      if (!commit || i >= n) *reinterpret_cast<uint2*>(desc + i * kDescBytes + sub * 8) = v[k];
      uint32_t s = __dp4a(v[k].x, v[k].x, 0u);
      s = __dp4a(v[k].y, v[k].y, s);
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // every lane of the half-warp: |b|^2 of row i
      own_s = k == sub ? s : own_s;
      // W = kFoldC - h as  digits[0] + 255 * sum(digits[1..31]),  every digit <= 255, spread greedily over digits 1..31: the first
      // nfull of them are 255, the next one holds the remainder.  Zero on padding rows and on rows whose W does not fit (the
      // whole set then takes the epilogue-add variant, which reads wcol = kNoFoldC - h).  Lane `sub` writes digits 2 sub, 2 sub + 1.
      const int32_t W = kFoldC - static_cast<int32_t>(s >> 1);
      uint32_t pair16 = 0;
      if (i < n && W >= 2) {
        const uint32_t rest = static_cast<uint32_t>(W) / 255u, d0 = static_cast<uint32_t>(W) % 255u;
        const uint32_t nfull = rest / 255u, rem = rest % 255u;       // rest <= 31 * 255
        const uint32_t j0 = 2u * sub, j1 = j0 + 1u;
        const uint32_t b0 = j0 == 0 ? d0 : (j0 <= nfull ? 255u : (j0 == nfull + 1 ? rem : 0u));
        const uint32_t b1 = j1 <= nfull ? 255u : (j1 == nfull + 1 ? rem : 0u);
        pair16 = b0 | (b1 << 8);
      }
      *reinterpret_cast<uint16_t*>(ext + i * kExtBytes + 2 * sub) = static_cast<uint16_t>(pair16);
    }
    // the window's scalars, one row per lane
    int32_t wn = 0;
    bool bad = false;
    if (own < n) {
      const int32_t h = static_cast<int32_t>(own_s >> 1);
      wn = kNoFoldC - h;                                              // >= 2
      bad = kFoldC - h < 2;
    }
    b.jobs[j].norm[own] = own < n ? static_cast<int32_t>(own_s) : 0;
    b.jobs[j].wcol[own] = wn;
    if (!commit || own >= n) {
      b.jobs[j].loc[own] = make_float2(__uint_as_float(own_loc.x), __uint_as_float(own_loc.y));
      b.jobs[j].sigtheta[own] = make_float2(__uint_as_float(own_st.x), __uint_as_float(own_st.y));
    }
    const int32_t wtop = __reduce_max_sync(0xffffffffu, wn);
    if (lane == 0) b.jobs[j].wmax[win] = wtop;                        // 0 for windows of padding only
    if (__any_sync(0xffffffffu, bad) && lane == 0) {
      atomicOr(b.jobs[j].meta, 1);
      *reinterpret_cast<volatile int32_t*>(b.jobs[j].h_meta) = 1;     // pinned host word, read after the set's `ready` event
    }
  }
}

// jobs: host array; each launch takes up to kPackInline of them.  Every job's n_rows is a multiple of 32.
void launch_pack(cudaStream_t st, const PackJob* jobs, uint32_t n_jobs, int sm_count, uint64_t* launches) {
  for (uint32_t at = 0; at < n_jobs;) {
    PackBatch b;
    b.n_jobs = 0; b.total_windows = 0;
    while (at < n_jobs && b.n_jobs < static_cast<uint32_t>(kPackInline)) {
      const uint64_t windows = jobs[at].n_rows / kChunk;
      if (windows == 0) { ++at; continue; }
      if (static_cast<uint64_t>(b.total_windows) + windows > 0xFFFFFFFFull) break;
      b.jobs[b.n_jobs] = jobs[at];
      b.jobs[b.n_jobs].win_begin = b.total_windows;
      b.total_windows += static_cast<uint32_t>(windows);
      ++b.n_jobs; ++at;
    }
    if (b.n_jobs == 0) continue;
    uint64_t blocks = (static_cast<uint64_t>(b.total_windows) + 7) / 8;
    const uint64_t cap = static_cast<uint64_t>(sm_count) * 8;
    if (blocks > cap) blocks = cap;
    launch_pdl(pack_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, b);
    if (launches) *launches += 1;
  }
}

// ---------------------------------------------------------------- spatially sorted copy of a packed set (masked searches, api.cu sorted_of)
// Rows are ordered along a Z-curve of their image location, so that the 32 rows of a candidate window and the 256 / 128 rows of a
// target / query tile cover a compact patch of the image: an epipolar band then misses most windows and most tiles, and the
// matcher skips them from their bounding boxes (smax_tc.cu, cull_kernel in exact.cu).  The key is only a locality heuristic --
// correctness rests on the boxes, which are taken from the rows themselves: Morton code of the top 16 bits of the two
// coordinates' order-preserving integer images (sign + exponent + 7 mantissa bits: cells of 4 px between 512 and 1024, finer
// below).  A NaN coordinate passes the reference's mask (src/MatchFactory.cu:2241-2254, :1627): such rows sort last and make the
// boxes of their window and tile NaN, which no test can cull.
__device__ __forceinline__ uint32_t spread16(uint32_t v) {      // abcd -> 0a0b0c0d
  v &= 0xFFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
__device__ __forceinline__ uint32_t orderable_f32(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void sort_keys_zcurve_kernel(const float2* __restrict__ loc, uint32_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 l = loc[i];
  keys[i] = (l.x != l.x || l.y != l.y) ? 0xFFFFFFFFu : ((spread16(orderable_f32(l.y) >> 16) << 1) | spread16(orderable_f32(l.x) >> 16));
  vals[i] = i;
}
// One block per 256-row tile, one warp per 32-row window: bounding box (x0, x1, y0, y1) of the window's real rows -- all NaN if
// any coordinate is NaN, (+inf, -inf, +inf, -inf) if the window holds padding only -- then the tile's box from its 8 windows.
__global__ void __launch_bounds__(256) boxes_kernel(const float2* __restrict__ loc, uint32_t n, float4* __restrict__ wbox, float4* __restrict__ tbox) {
  __shared__ float4 s_box[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 256u + threadIdx.x;
  const float inf = __int_as_float(0x7f800000);
  float x0 = inf, x1 = -inf, y0 = inf, y1 = -inf;
  bool bad = false;
  if (r < n) {
    const float2 l = loc[r];
    bad = (l.x != l.x) || (l.y != l.y);
    if (!bad) { x0 = x1 = l.x; y0 = y1 = l.y; }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (__any_sync(0xffffffffu, bad)) x0 = x1 = y0 = y1 = __uint_as_float(0x7fc00000u);
  if (lane == 0) {
    const float4 b4 = make_float4(x0, x1, y0, y1);
    wbox[blockIdx.x * 8u + warp] = b4;
    s_box[warp] = b4;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 t = s_box[0];
    bool nan = t.x != t.x;
    for (int w = 1; w < 8; ++w) {
      const float4 b4 = s_box[w];
      nan |= b4.x != b4.x;
      t.x = fminf(t.x, b4.x); t.y = fmaxf(t.y, b4.y); t.z = fminf(t.z, b4.z); t.w = fmaxf(t.w, b4.w);
    }
    if (nan) t = make_float4(__uint_as_float(0x7fc00000u), __uint_as_float(0x7fc00000u), __uint_as_float(0x7fc00000u), __uint_as_float(0x7fc00000u));
    tbox[blockIdx.x] = t;
  }
}
// one warp per output row: row r of the sorted set = row perm[r] of the source; rows [n, n_rows) are padding (zero descriptor
// and digits, W = 0)
__global__ void __launch_bounds__(256) gather_sorted_kernel(const uint32_t* __restrict__ perm, uint32_t n, uint32_t n_rows,
                                                            const uint8_t* __restrict__ s_desc, const uint8_t* __restrict__ s_ext,
                                                            const int32_t* __restrict__ s_norm, const int32_t* __restrict__ s_wcol,
                                                            const float2* __restrict__ s_loc, const float2* __restrict__ s_sigtheta,
                                                            uint8_t* __restrict__ desc, uint8_t* __restrict__ ext, int32_t* __restrict__ norm,
                                                            int32_t* __restrict__ wcol, int32_t* __restrict__ wmax, float2* __restrict__ loc,
                                                            float2* __restrict__ sigtheta, uint32_t* __restrict__ orig) {
  const int lane = threadIdx.x & 31;
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n_rows) return;
  uint32_t* d = reinterpret_cast<uint32_t*>(desc + static_cast<uint64_t>(r) * kDescBytes);
  uint32_t* e = reinterpret_cast<uint32_t*>(ext + static_cast<uint64_t>(r) * kExtBytes);
  if (r < n) {
    const uint32_t src = perm[r];
    d[lane] = reinterpret_cast<const uint32_t*>(s_desc + static_cast<uint64_t>(src) * kDescBytes)[lane];
    if (lane < 8) e[lane] = reinterpret_cast<const uint32_t*>(s_ext + static_cast<uint64_t>(src) * kExtBytes)[lane];
    if (lane == 0) {
      norm[r] = s_norm[src];
      const int32_t w = s_wcol[src];
      wcol[r] = w;
      atomicMax(&wmax[r >> 5], w);
      loc[r] = s_loc[src];
      sigtheta[r] = s_sigtheta[src];
      orig[r] = src;
    }
  } else {
    d[lane] = 0u;
    if (lane < 8) e[lane] = 0u;
    if (lane == 0) {
      norm[r] = 0; wcol[r] = 0;
      loc[r] = make_float2(0.f, 0.f);
      sigtheta[r] = make_float2(0.f, 0.f);
      orig[r] = 0u;
    }
  }
}
// ---- small sets (n < 65 536: every set of the reference's fixtures): the whole ordering in ONE block, the copy in one more launch.
// A radix sort of 32 k keys is ten launches of a few microseconds each (histogram, scan, four onesweep passes, ...), ~70 us per set
// against a matcher launch of ~230; a counting sort over image cells needs no passes at all.  Cells: 128 x 128 over the bounding box
// of the finite locations, in Z order (14-bit Morton code; one more bin for rows with a NaN coordinate, which sort last); rows of a
// cell keep their original order, so the result is a stable sort by cell and does not depend on the order the atomics land in.
constexpr int kCellBits = 7;                                   // per axis
constexpr uint32_t kCellBins = (1u << (2 * kCellBits)) + 1;    // + the NaN bin
constexpr int kCellThreads = 1024;
constexpr int kCellBatch = 8;
constexpr uint32_t kCellSortMaxRows = 65535;                   // row indices travel as u16 in shared memory
struct CellSortSmem {
  uint32_t cursor[kCellBins + 3];        // histogram -> exclusive scan -> running cursor (ends at each bin's end)
  uint32_t part[kCellThreads / 32];
  float red[4][kCellThreads / 32];
  float lim[4];
  uint16_t perm[kCellSortMaxRows + 1];
};
__device__ __forceinline__ uint32_t spread7(uint32_t v) {     // 0000000 gfedcba -> 0g0f0e0d0c0b0a
  v &= 0x7Fu;
  v = (v | (v << 4)) & 0x070Fu;
  v = (v | (v << 2)) & 0x1333u;
  v = (v | (v << 1)) & 0x1555u;
  return v;
}
// position of cell (x, y) of the 128 x 128 grid along the Hilbert curve: consecutive positions are always neighbouring cells, so a
// window of 32 consecutive rows never straddles one of the Z curve's jumps and its bounding box stays tight
constexpr bool kCellHilbert = true;
__device__ __forceinline__ uint32_t hilbert7(uint32_t x, uint32_t y) {
  uint32_t d = 0;
#pragma unroll
  for (uint32_t s = 1u << (kCellBits - 1); s > 0; s >>= 1) {
    const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
    d += s * s * ((3u * rx) ^ ry);
    if (ry == 0) {
      if (rx == 1) { x = (1u << kCellBits) - 1 - x; y = (1u << kCellBits) - 1 - y; }
      const uint32_t t = x; x = y; y = t;
    }
  }
  return d;
}
__global__ void __launch_bounds__(kCellThreads, 1) cell_sort_kernel(const float2* __restrict__ loc, uint32_t n, uint16_t* __restrict__ cell_of,
                                                                    uint32_t* __restrict__ perm_out) {
  extern __shared__ __align__(16) uint8_t cs_raw[];
  CellSortSmem& sm = *reinterpret_cast<CellSortSmem*>(cs_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t b = tid; b < kCellBins + 3; b += kCellThreads) sm.cursor[b] = 0;
  pdl_wait();                                                   // loc is the pack launch's output
  pdl_launch_dependents();
  // ---- bounding box of the finite locations
  const float inf = __int_as_float(0x7f800000);
  // (one block has to cover the memory latency by itself: every pass keeps kCellBatch independent loads per thread in flight)
  float x0 = inf, x1 = -inf, y0 = inf, y1 = -inf;
  for (uint32_t base = tid; base < n; base += kCellBatch * kCellThreads) {
    float2 l[kCellBatch];
#pragma unroll
    for (int k = 0; k < kCellBatch; ++k) { const uint32_t i = base + k * kCellThreads; l[k] = i < n ? loc[i] : make_float2(inf, inf); }
#pragma unroll
    for (int k = 0; k < kCellBatch; ++k)
      if (fabsf(l[k].x) < inf && fabsf(l[k].y) < inf) { x0 = fminf(x0, l[k].x); x1 = fmaxf(x1, l[k].x); y0 = fminf(y0, l[k].y); y1 = fmaxf(y1, l[k].y); }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0) { sm.red[0][warp] = x0; sm.red[1][warp] = x1; sm.red[2][warp] = y0; sm.red[3][warp] = y1; }
  __syncthreads();
  if (warp == 0) {
    x0 = sm.red[0][lane]; x1 = sm.red[1][lane]; y0 = sm.red[2][lane]; y1 = sm.red[3][lane];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
      y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if (lane == 0) {
      const float cells = static_cast<float>(1 << kCellBits);
      sm.lim[0] = x0; sm.lim[1] = (x1 > x0) ? cells / (x1 - x0) : 0.0f;
      sm.lim[2] = y0; sm.lim[3] = (y1 > y0) ? cells / (y1 - y0) : 0.0f;
    }
  }
  __syncthreads();
  const float ox = sm.lim[0], sx = sm.lim[1], oy = sm.lim[2], sy = sm.lim[3];
  // ---- cell of every row, histogram
  for (uint32_t base = tid; base < n; base += kCellBatch * kCellThreads) {
    float2 l[kCellBatch];
#pragma unroll
    for (int k = 0; k < kCellBatch; ++k) { const uint32_t i = base + k * kCellThreads; l[k] = i < n ? loc[i] : make_float2(inf, inf); }
#pragma unroll
    for (int k = 0; k < kCellBatch; ++k) {
      const uint32_t i = base + k * kCellThreads;
      if (i >= n) break;
      uint32_t c = kCellBins - 1;                               // a NaN / infinite coordinate: last
      if (fabsf(l[k].x) < inf && fabsf(l[k].y) < inf) {
        const int cx = min((1 << kCellBits) - 1, max(0, static_cast<int>((l[k].x - ox) * sx)));
        const int cy = min((1 << kCellBits) - 1, max(0, static_cast<int>((l[k].y - oy) * sy)));
        c = kCellHilbert ? hilbert7(static_cast<uint32_t>(cx), static_cast<uint32_t>(cy))
                         : ((spread7(static_cast<uint32_t>(cy)) << 1) | spread7(static_cast<uint32_t>(cx)));
      }
      cell_of[i] = static_cast<uint16_t>(c);
      atomicAdd(&sm.cursor[c], 1u);
    }
  }
  __syncthreads();
  // ---- exclusive scan of the bins: each thread a run of consecutive bins, then the block
  constexpr uint32_t kRun = (kCellBins + kCellThreads - 1) / kCellThreads;   // 17
  const uint32_t b0 = tid * kRun;
  uint32_t sum = 0;
  for (uint32_t k = 0; k < kRun; ++k) if (b0 + k < kCellBins) sum += sm.cursor[b0 + k];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) sm.part[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = sm.part[lane], w = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
    sm.part[lane] = w - v;
  }
  __syncthreads();
  uint32_t run = sm.part[warp] + incl - sum;
  for (uint32_t k = 0; k < kRun; ++k)
    if (b0 + k < kCellBins) { const uint32_t c = sm.cursor[b0 + k]; sm.cursor[b0 + k] = run; run += c; }
  __syncthreads();
  // ---- scatter (order inside a cell: whatever the atomics give), then every cell's rows back into ascending order
  for (uint32_t base = tid; base < n; base += kCellBatch * kCellThreads) {
    uint32_t c[kCellBatch];
#pragma unroll
    for (int k = 0; k < kCellBatch; ++k) { const uint32_t i = base + k * kCellThreads; c[k] = i < n ? cell_of[i] : 0u; }
#pragma unroll
    for (int k = 0; k < kCellBatch; ++k) {
      const uint32_t i = base + k * kCellThreads;
      if (i >= n) break;
      const uint32_t at = atomicAdd(&sm.cursor[c[k]], 1u);
      sm.perm[at] = static_cast<uint16_t>(i);
    }
  }
  __syncthreads();
  for (uint32_t k = 0; k < kRun; ++k) {
    const uint32_t b = b0 + k;
    if (b >= kCellBins) break;
    const uint32_t end = sm.cursor[b], begin = b ? sm.cursor[b - 1] : 0u;   // cursor[b] now ends bin b = begins bin b + 1
    for (uint32_t i = begin + 1; i < end; ++i) {                             // insertion sort: a cell holds a handful of rows
      const uint16_t v = sm.perm[i];
      uint32_t j = i;
      while (j > begin && sm.perm[j - 1] > v) { sm.perm[j] = sm.perm[j - 1]; --j; }
      sm.perm[j] = v;
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += kCellThreads) perm_out[i] = sm.perm[i];
}

// One block per 256-row tile of the sorted set, one warp per 32-row window: lane l takes the scalars of row l, the descriptor and
// digit rows are copied one row per step by the whole warp; the window's largest W and its bounding box are warp reductions, the
// tile's box comes from its 8 windows -- no atomics, nothing to zero beforehand, no separate box pass.
__global__ void __launch_bounds__(256) gather_tiles_kernel(const uint32_t* __restrict__ perm, uint32_t n, uint32_t n_rows,
                                                           const uint8_t* __restrict__ s_desc, const uint8_t* __restrict__ s_ext,
                                                           const int32_t* __restrict__ s_norm, const int32_t* __restrict__ s_wcol,
                                                           const float2* __restrict__ s_loc, const float2* __restrict__ s_sigtheta,
                                                           uint8_t* __restrict__ desc, uint8_t* __restrict__ ext, int32_t* __restrict__ norm,
                                                           int32_t* __restrict__ wcol, int32_t* __restrict__ wmax, float2* __restrict__ loc,
                                                           float2* __restrict__ sigtheta, uint32_t* __restrict__ orig,
                                                           float4* __restrict__ wbox, float4* __restrict__ tbox) {
  __shared__ float4 s_box[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 256u + threadIdx.x;          // this lane's row (n_rows % 256 == 0)
  pdl_wait();
  pdl_launch_dependents();
  const bool real = r < n;
  const uint32_t src = real ? perm[r] : 0u;
  const float inf = __int_as_float(0x7f800000);
  float x0 = inf, x1 = -inf, y0 = inf, y1 = -inf;
  bool bad = false;
  int32_t w = 0;
  if (real) {
    const float2 l = s_loc[src];
    w = s_wcol[src];
    norm[r] = s_norm[src]; wcol[r] = w; loc[r] = l; sigtheta[r] = s_sigtheta[src]; orig[r] = src;
    bad = (l.x != l.x) || (l.y != l.y);
    if (!bad) { x0 = x1 = l.x; y0 = y1 = l.y; }
  } else {
    norm[r] = 0; wcol[r] = 0; loc[r] = make_float2(0.f, 0.f); sigtheta[r] = make_float2(0.f, 0.f); orig[r] = 0u;
  }
  // descriptor + digit rows of the window, one row per step (128 + 32 contiguous bytes each)
  const uint32_t row0 = blockIdx.x * 256u + warp * 32u;
#pragma unroll 16
  for (int k = 0; k < 32; ++k) {
    const uint32_t sk = __shfl_sync(0xffffffffu, src, k);
    const bool rk = row0 + k < n;
    uint32_t d = 0u, e = 0u;
    if (rk) {
      d = reinterpret_cast<const uint32_t*>(s_desc + static_cast<uint64_t>(sk) * kDescBytes)[lane];
      if (lane < 8) e = reinterpret_cast<const uint32_t*>(s_ext + static_cast<uint64_t>(sk) * kExtBytes)[lane];
    }
    reinterpret_cast<uint32_t*>(desc + static_cast<uint64_t>(row0 + k) * kDescBytes)[lane] = d;
    if (lane < 8) reinterpret_cast<uint32_t*>(ext + static_cast<uint64_t>(row0 + k) * kExtBytes)[lane] = e;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (__any_sync(0xffffffffu, bad)) x0 = x1 = y0 = y1 = __uint_as_float(0x7fc00000u);
  if (lane == 0) {
    const float4 b4 = make_float4(x0, x1, y0, y1);
    wmax[row0 >> 5] = w;
    wbox[row0 >> 5] = b4;
    s_box[warp] = b4;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 t = s_box[0];
    bool nan = t.x != t.x;
    for (int k = 1; k < 8; ++k) {
      const float4 b4 = s_box[k];
      nan |= b4.x != b4.x;
      t.x = fminf(t.x, b4.x); t.y = fmaxf(t.y, b4.y); t.z = fminf(t.z, b4.z); t.w = fmaxf(t.w, b4.w);
    }
    if (nan) t = make_float4(__uint_as_float(0x7fc00000u), __uint_as_float(0x7fc00000u), __uint_as_float(0x7fc00000u), __uint_as_float(0x7fc00000u));
    tbox[blockIdx.x] = t;
  }
}
bool cell_sort_serves(uint64_t n) { return n >= 1 && n <= kCellSortMaxRows; }
// scratch: perm (n u32) then cell ids (n u16); two launches in the stream's dependent-launch chain
cudaError_t launch_cell_sort_gather(cudaStream_t st, const ssrlcv_b200_set_impl& src, ssrlcv_b200_set_impl& dst, uint32_t* perm, uint16_t* cell_of) {
  const uint32_t n = static_cast<uint32_t>(src.n), n_rows = static_cast<uint32_t>(dst.n_pad);
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(cell_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(CellSortSmem)));
    if (e != cudaSuccess) return e;
    if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
  }
  cudaError_t e = launch_pdl(cell_sort_kernel, dim3(1), dim3(kCellThreads), sizeof(CellSortSmem), st, src.loc, n, cell_of, perm);
  if (e != cudaSuccess) return e;
  return launch_pdl(gather_tiles_kernel, dim3(n_rows / 256), dim3(256), 0, st, static_cast<const uint32_t*>(perm), n, n_rows,
                    static_cast<const uint8_t*>(src.desc), static_cast<const uint8_t*>(src.ext), static_cast<const int32_t*>(src.norm),
                    static_cast<const int32_t*>(src.wcol), static_cast<const float2*>(src.loc), static_cast<const float2*>(src.sigtheta),
                    dst.desc, dst.ext, dst.norm, dst.wcol, dst.wmax, dst.loc, dst.sigtheta, dst.orig, dst.wbox, dst.tbox);
}

size_t sort_x_temp_bytes(uint32_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                  static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), static_cast<int>(n));
  return bytes;
}
// scratch: [keys_in | keys_out | vals_in | vals_out] (4 x n u32, each 256-byte aligned by the caller) followed by cub's temp storage
cudaError_t launch_sort_gather(cudaStream_t st, const ssrlcv_b200_set_impl& src, ssrlcv_b200_set_impl& dst, uint32_t* keys_in,
                               uint32_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, void* temp, size_t temp_bytes) {
  const uint32_t n = static_cast<uint32_t>(src.n), n_rows = static_cast<uint32_t>(dst.n_pad);
  if (n) {
    sort_keys_zcurve_kernel<<<(n + 255) / 256, 256, 0, st>>>(src.loc, n, keys_in, vals_in);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, static_cast<int>(n), 0, 32, st);
    if (e != cudaSuccess) return e;
  }
  gather_sorted_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(vals_out, n, n_rows, src.desc, src.ext, src.norm, src.wcol, src.loc, src.sigtheta,
                                                         dst.desc, dst.ext, dst.norm, dst.wcol, dst.wmax, dst.loc, dst.sigtheta, dst.orig);
  boxes_kernel<<<n_rows / 256, 256, 0, st>>>(dst.loc, n, dst.wbox, dst.tbox);
  return cudaGetLastError();
}

}  // namespace b200
