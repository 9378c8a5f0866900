// exact.cu -- exact SIMT evaluation of one query against a target range.
// Two uses:
//   * tie resolution: queries whose maximum score occurred in more than one 32-target window of the
//     tensor-core kernel (rare: needs equal squared distances up to parity) are re-evaluated here;
//   * SSRLCV_B200_ENGINE_SIMT: the whole problem, to cross-check the tcgen05 path on the device.
// One CTA per (split, query): thread t owns targets t, t+256, ... so f % 32 is the thread's lane, as in the reference's
// 32-thread blocks (ref: src/MatchFactory.cu:1466-1492); the winner is the minimum of the key (d2, f%32, f).
#include "geometry.cuh"
#include "matcher.cuh"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v), o);
    const uint32_t hi = __shfl_xor_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), o);
    const uint64_t w = (static_cast<uint64_t>(hi) << 32) | lo;
    v = w < v ? w : v;
  }
  return v;
}

// One CTA (8 warps) per (split, query) item; lane/warp-strided over the target range, then a block-wide min.
template <int kMask>
__device__ __forceinline__ void exact_items(const SmaxArgs& a, int use_list, uint64_t* s_min) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float eps = a.eps;
  const uint32_t total = use_list ? *a.n_flagged : a.n_splits * a.nq;
  for (uint32_t item = blockIdx.x; item < total; item += gridDim.x) {
    uint32_t split, row;
    uint64_t f_begin, f_end;
    if (use_list) {
      // entry = key slot * nq_pad + row; the slot's target range is the tensor-core kernel's unit (matcher.cuh)
      const uint32_t e = a.flagged[item];
      split = e / a.nq_pad; row = e % a.nq_pad;
      const uint32_t qt = row / a.q_tile_rows;
      const uint32_t cta = cta_of_tile(a.total_tiles, a.n_ctas, a.tile_begin + static_cast<uint64_t>(qt) * a.n_ttiles) + split;
      uint32_t j0, j1;
      unit_range(a, cta, qt, j0, j1);
      f_begin = static_cast<uint64_t>(j0) * kBN;
      f_end = static_cast<uint64_t>(j1) * kBN;
    } else {
      split = item / a.nq; row = item % a.nq;
      f_begin = static_cast<uint64_t>(split) * a.tiles_per_split * kBN;
      f_end = f_begin + static_cast<uint64_t>(a.tiles_per_split) * kBN;
    }
    if (f_end > a.nt) f_end = a.nt;
    const uint4* qd = reinterpret_cast<const uint4*>(a.q_desc + static_cast<uint64_t>(row) * kDescBytes);
    uint4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __ldg(qd + i);
    const int32_t qn = a.q_norm[row];
    QueryMask qm;
    if (kMask != SSRLCV_B200_MASK_NONE) qm = load_query_mask<kMask>(a.q_mask + 2 * static_cast<uint64_t>(row));
    uint64_t best = kKeyNone;
    const uint64_t excluded = a.q_exclude ? (a.q_exclude[row] & 0xFFFFFFFFull) : ~0ull;   // runner-up search: one target row this query must not match
    for (uint64_t f = f_begin + threadIdx.x; f < f_end; f += 256) {     // f % 32 == lane for every f of this thread
      if (f == excluded) continue;
      if (kMask != SSRLCV_B200_MASK_NONE) {
        const float2 tl = a.t_loc[f];
        if (mask_skips<kMask>(qm, tl.x, tl.y, eps)) continue;
      }
      const uint4* td = reinterpret_cast<const uint4*>(a.t_desc + f * kDescBytes);
      uint32_t dot = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 t = __ldg(td + i);
        dot = __dp4a(q[i].x, t.x, dot); dot = __dp4a(q[i].y, t.y, dot);
        dot = __dp4a(q[i].z, t.z, dot); dot = __dp4a(q[i].w, t.w, dot);
      }
      const uint32_t d2 = static_cast<uint32_t>(qn + a.t_norm[f] - 2 * static_cast<int32_t>(dot));
      const uint64_t key = make_key(d2, (a.t_orig ? static_cast<uint64_t>(a.t_orig[f]) : f) + a.target_offset);   // x-sorted rows: the caller's index
      best = key < best ? key : best;
    }
    best = warp_min_u64(best);
    if (lane == 0) s_min[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t m = s_min[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = s_min[w] < m ? s_min[w] : m;
      a.keys[static_cast<uint64_t>(split) * a.nq_pad + (a.q_orig ? a.q_orig[row] : row)] = m;   // row < nq here
    }
    __syncthreads();
  }
}

// SSRLCV_B200_ENGINE_SIMT: the whole problem of one pair
template <int kMask>
__global__ void __launch_bounds__(256) exact_kernel(SmaxArgs a) {
  __shared__ uint64_t s_min[8];
  exact_items<kMask>(a, 0, s_min);
}
// tie fallback of the tensor-core engine: the flagged (slot, query) entries of every job of a launch.  The tie counters are
// persistent device words in two banks: a matcher launch counts into one bank and clears the other for the call after it
// (smax_tc.cu), so nothing is zeroed or staged in front of the matcher and nothing is reset here.
template <int kMask>
__global__ void __launch_bounds__(256) exact_list_kernel(const __grid_constant__ SmaxBatch b) {
  __shared__ uint64_t s_min[8];
  __shared__ int s_any;
  pdl_launch_dependents();
  pdl_wait();                                         // the matcher launch whose tie counters are read next
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < b.n_jobs; j += blockDim.x)
    if (b.flags[j] != 0) s_any = 1;                   // the counters of the launch's jobs are contiguous: one parallel look
  __syncthreads();
  if (!s_any) return;                                 // the common case: no tie anywhere in the launch
  const SmaxJob* jobs = b.jobs ? b.jobs : &b.inline_job;
  for (uint32_t j = 0; j < b.n_jobs; ++j) {
    if (b.flags[j] == 0) continue;
    const SmaxArgs a = jobs[j].a;
    exact_items<kMask>(a, 1, s_min);
  }
}

void launch_exact(cudaStream_t st, const SmaxArgs& a, int mask_kind, int sm_count) {
  const int blocks = sm_count * 8;
  switch (mask_kind) {
    case SSRLCV_B200_MASK_NONE: exact_kernel<SSRLCV_B200_MASK_NONE><<<blocks, 256, 0, st>>>(a); break;
    case SSRLCV_B200_MASK_FUNDAMENTAL: exact_kernel<SSRLCV_B200_MASK_FUNDAMENTAL><<<blocks, 256, 0, st>>>(a); break;
    default: exact_kernel<SSRLCV_B200_MASK_DOUBLE><<<blocks, 256, 0, st>>>(a); break;
  }
}
void launch_exact_list(cudaStream_t st, const SmaxBatch& b, int mask_kind, int sm_count) {
  const int blocks = sm_count * 2;
  switch (mask_kind) {
    case SSRLCV_B200_MASK_NONE: launch_pdl(exact_list_kernel<SSRLCV_B200_MASK_NONE>, dim3(blocks), dim3(256), 0, st, b); break;
    case SSRLCV_B200_MASK_FUNDAMENTAL: launch_pdl(exact_list_kernel<SSRLCV_B200_MASK_FUNDAMENTAL>, dim3(blocks), dim3(256), 0, st, b); break;
    default: launch_pdl(exact_list_kernel<SSRLCV_B200_MASK_DOUBLE>, dim3(blocks), dim3(256), 0, st, b); break;
  }
}

// ---- per-query mask prologue ---------------------------------------------------------------------
template <int kMask>
__global__ void __launch_bounds__(128) query_mask_kernel(MaskParams mp, const float2* __restrict__ loc, uint64_t n, float4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();                                         // loc comes from the pack launch in front of this one
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  store_query_mask<kMask>(out + 2 * i, make_query_mask<kMask>(mp, loc[i]));
}
void launch_query_masks(cudaStream_t st, const MaskParams& mp, const float2* loc, uint64_t n, float4* out) {
  if (n == 0 || mp.kind == SSRLCV_B200_MASK_NONE) return;
  const unsigned blocks = static_cast<unsigned>((n + 127) / 128);
  if (mp.kind == SSRLCV_B200_MASK_DOUBLE) launch_pdl(query_mask_kernel<SSRLCV_B200_MASK_DOUBLE>, dim3(blocks), dim3(128), 0, st, mp, loc, n, out);
  else launch_pdl(query_mask_kernel<SSRLCV_B200_MASK_FUNDAMENTAL>, dim3(blocks), dim3(128), 0, st, mp, loc, n, out);
}

// ---- dead tile pairs of a masked search on spatially sorted sets --------------------------------------------
// One block per (query tile, word of 32 target tiles), one thread per query: bit j of row qt of `bits` is set iff some query of the
// tile can match some target whose location lies in the box of target tile j (geometry.cuh box_skips: exact and conservative).
// The matcher's roles skip the dead pairs before the MMA.
template <int kMask>
__global__ void __launch_bounds__(kBM) cull_kernel(const float4* __restrict__ q_mask, uint32_t nq, const float4* __restrict__ tbox,
                                                   uint32_t n_ttiles, float eps, uint32_t* __restrict__ bits, uint32_t stride) {
  __shared__ uint32_t s_word;
  pdl_launch_dependents();
  pdl_wait();                                         // q_mask comes from the launch in front of this one
  const uint32_t row = blockIdx.x * kBM + threadIdx.x, w = blockIdx.y;
  const bool real = row < nq;
  if (threadIdx.x == 0) s_word = 0;
  __syncthreads();
  const QueryMask qm = load_query_mask<kMask>(q_mask + 2 * static_cast<uint64_t>(real ? row : 0));
  uint32_t mine = 0;
  const uint32_t j_end = min(n_ttiles - w * 32, 32u);
  if (real)
    for (uint32_t b = 0; b < j_end; ++b) mine |= box_skips<kMask>(qm, __ldg(tbox + w * 32 + b), eps) ? 0u : (1u << b);
  mine = __reduce_or_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicOr(&s_word, mine);
  __syncthreads();
  if (threadIdx.x == 0) bits[blockIdx.x * stride + w] = s_word;
}
void launch_cull(cudaStream_t st, int mask_kind, const float4* q_mask, uint32_t nq, uint32_t n_qtiles, const float4* tbox, uint32_t n_ttiles,
                 float eps, uint32_t* bits, uint32_t stride) {
  if (n_qtiles == 0 || n_ttiles == 0) return;
  const dim3 grid(n_qtiles, stride);                  // stride <= 8192 (capped by the caller)
  if (mask_kind == SSRLCV_B200_MASK_DOUBLE) launch_pdl(cull_kernel<SSRLCV_B200_MASK_DOUBLE>, grid, dim3(kBM), 0, st, q_mask, nq, tbox, n_ttiles, eps, bits, stride);
  else launch_pdl(cull_kernel<SSRLCV_B200_MASK_FUNDAMENTAL>, grid, dim3(kBM), 0, st, q_mask, nq, tbox, n_ttiles, eps, bits, stride);
}

// ---- small diagnostics kernels -----------------------------------------------------------------
__global__ void endpoints_kernel(MaskParams mp, const float2* loc, uint64_t n, float4* out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = epipolar_endpoints_ref(mp, loc[i]);
}
void launch_endpoints(cudaStream_t st, const MaskParams& mp, const float2* loc, uint64_t n, float4* out) {
  if (n) endpoints_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(mp, loc, n, out);
}
__global__ void sincos_kernel(const float* x, uint64_t n, float* s, float* c) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) { s[i] = sinf(x[i]); c[i] = cosf(x[i]); }
}
void launch_sincos(cudaStream_t st, const float* x, uint64_t n, float* s, float* c) {
  if (n) sincos_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(x, n, s, c);
}

}  // namespace b200
