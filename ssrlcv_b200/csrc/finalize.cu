// finalize.cu -- everything after the per-query winner is known:
//   * merge candidate keys over target splits / shards (min of d2<<37 | lane<<32 | f)
//   * absolute threshold, seed-distance ratio test with the reference's float compares
//     (ref: src/MatchFactory.cu:1494, :1691-1697, :2270-2281, :2899-2910)
//   * writers for Match / DMatch / uint2_pair / FeatureMatch
//   * stable compaction by query index == thrust::remove_if in validateMatches (ref: :31-135)
//   * refineMatches / sortMatches / getRawMatches equivalents (ref: :137-160, :191-222, :257-280)
// HBM/latency-bound record shuffling: 8 B x slots read and <= 48 B written per query.
#include <cfloat>

#include <cub/cub.cuh>

#include "matcher.cuh"
#include "ptx.cuh"

namespace b200 {

struct Decision { bool valid; float cur; uint32_t f; bool has; };

__device__ __forceinline__ Decision decide(const FinalizeArgs& a, uint32_t i) {
  uint64_t key = kKeyNone;
  for (uint32_t s = 0; s < a.n_splits; ++s) {
    const uint64_t k = a.keys[static_cast<uint64_t>(s) * a.nq_pad + i];
    key = k < key ? k : key;
  }
  Decision d;
  d.has = false; d.f = 0; d.cur = a.abs_thr;
  if (key < kKeyNeedsExact) {                                                  // neither "no candidate" nor an unresolved tie marker
    const float dist = static_cast<float>(static_cast<uint32_t>(key >> 37));   // exact: d2 < 2^24
    if (a.abs_thr > dist) { d.has = true; d.cur = dist; d.f = static_cast<uint32_t>(key); }
  }
  if (a.seed) {
    d.valid = !(d.cur >= a.abs_thr || !d.has);
    if (d.valid && (__fdiv_rn(d.cur, a.seed[i]) > a.ratio)) d.valid = false;
  } else {
    d.valid = !(d.cur >= a.abs_thr);
  }
  // uint2_pair encodes "invalid" as a self pair and validateMatches removes by that predicate
  if (a.out_kind == SSRLCV_B200_OUT_PAIR && d.valid && a.qid == a.tid && d.f == i) d.valid = false;
  return d;
}

__device__ __forceinline__ void store_keypoints(RefKeyPoint* kp, uint32_t qid, uint32_t tid, float2 ql, float2 tl) {
  kp[0].parentId = static_cast<int32_t>(qid); kp[0].pad_ = 0; kp[0].loc = ql;
  kp[1].parentId = static_cast<int32_t>(tid); kp[1].pad_ = 0; kp[1].loc = tl;
}

__device__ __forceinline__ uint64_t ld_acquire_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Thresholds, writers and the stable compaction in ONE pass.  grid = (blocks of 256 queries, jobs).  A block draws a ticket
// (its position in scheduling order, so every block it will wait for is already running), decides its 256 queries,
// publishes its count stamped with this call's epoch and sums the counts of the tickets before it -- they are published
// without waiting for anything, so the wait is bounded by one block's decide().  Output order = ascending query index,
// exactly thrust::remove_if's (ref: src/MatchFactory.cu:31-135).
__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeBatch fb) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_pre[8];
  __shared__ uint32_t s_ticket;
  pdl_launch_dependents();
  pdl_wait();                                         // keys (and a staged job table) come from the launches in front of this one
  const FinalizeArgs a = fb.jobs ? fb.jobs[blockIdx.y] : fb.inline_job;
  const uint32_t nq = a.nq, n_blocks = (nq + 255) / 256;
  if (blockIdx.x >= n_blocks) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(a.ticket, 1u);
    if (t == n_blocks - 1) *a.ticket = 0;             // the last ticket of this job: ready for the next call
    s_ticket = t;
  }
  __syncthreads();
  const uint32_t vb = s_ticket;
  const uint32_t i = vb * 256 + threadIdx.x;

  Decision d; d.valid = false; d.has = false; d.cur = 0.f; d.f = 0;
  if (i < nq) d = decide(a, i);
  const bool emit = (i < nq) && (a.compact ? d.valid : true);
  const uint32_t bal = __ballot_sync(0xffffffffu, emit);
  if (lane == 0) s_warp[warp] = __popc(bal);
  // the record's gathers are issued before the wait on the preceding blocks' counts
  const float2 zero2 = make_float2(0.f, 0.f);
  const float2 ql = d.valid ? a.q_loc[i] : zero2;
  const float2 tl = (d.valid && a.t_loc) ? a.t_loc[d.f] : zero2;
  __syncthreads();
  uint32_t mine = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) mine += s_warp[w];
  const uint64_t stamp = static_cast<uint64_t>(a.epoch) << 32;
  if (threadIdx.x == 0) st_release_u64(a.state + vb, stamp | mine);
  // exclusive prefix over the tickets before this one
  uint32_t part = 0;
  for (uint32_t b = threadIdx.x; b < vb; b += 256) {
    uint64_t v;
    do { v = ld_acquire_u64(a.state + b); } while ((v >> 32) != a.epoch);
    part += static_cast<uint32_t>(v);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_pre[warp] = part;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) base += s_pre[w];
  if (vb == n_blocks - 1 && threadIdx.x == 0) {
    if (a.total) *a.total = base + mine;
    if (a.total_host) *reinterpret_cast<volatile uint32_t*>(a.total_host) = base + mine;
  }
  if (!emit) return;
  uint32_t off = base;
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  off += __popc(bal & ((1u << lane) - 1u));

  switch (a.out_kind) {
    case SSRLCV_B200_OUT_PAIR: {
      RefPair p;
      p.a[0] = a.qid; p.a[1] = i;
      if (d.valid) { p.b[0] = a.tid; p.b[1] = d.f; } else { p.b[0] = a.qid; p.b[1] = i; }
      static_cast<RefPair*>(a.out)[off] = p;
    } break;
    case SSRLCV_B200_OUT_MATCH: {
      RefMatch m;
      m.invalid = d.valid ? 0 : 1;
      for (int k = 0; k < 7; ++k) m.pad_[k] = 0;
      store_keypoints(m.kp, d.valid ? a.qid : 0, d.valid ? a.tid : 0, ql, tl);
      static_cast<RefMatch*>(a.out)[off] = m;
    } break;
    case SSRLCV_B200_OUT_DMATCH: {
      RefDMatch m;
      m.invalid = d.valid ? 0 : 1;
      for (int k = 0; k < 7; ++k) m.pad_[k] = 0;
      store_keypoints(m.kp, d.valid ? a.qid : 0, d.valid ? a.tid : 0, ql, tl);
      m.distance = d.cur; m.pad2_ = 0.f;
      static_cast<RefDMatch*>(a.out)[off] = m;
    } break;
    default: {
      RefFeatureMatch* o = static_cast<RefFeatureMatch*>(a.out) + off;
      o->invalid = d.valid ? 0 : 1;
      for (int k = 0; k < 7; ++k) o->pad_[k] = 0;
      store_keypoints(o->kp, d.valid ? a.qid : 0, d.valid ? a.tid : 0, ql, tl);
      o->distance = d.cur; o->pad2_ = 0.f;
      // descriptors start at byte 44 of a 320-byte record: 4-byte aligned words only
      uint32_t* w0 = reinterpret_cast<uint32_t*>(&o->desc[0]);
      uint32_t* w1 = reinterpret_cast<uint32_t*>(&o->desc[1]);
      if (d.valid) {
        const float2 qs = a.q_sigtheta[i], ts = a.t_sigtheta[d.f];
        w0[0] = __float_as_uint(qs.x); w0[1] = __float_as_uint(qs.y);
        w1[0] = __float_as_uint(ts.x); w1[1] = __float_as_uint(ts.y);
        const uint32_t* qd = reinterpret_cast<const uint32_t*>(a.q_desc + static_cast<uint64_t>(i) * kDescBytes);
        const uint32_t* td = reinterpret_cast<const uint32_t*>(a.t_desc + static_cast<uint64_t>(d.f) * kDescBytes);
        for (int k = 0; k < 32; ++k) { w0[2 + k] = qd[k]; w1[2 + k] = td[k]; }
      } else {
        for (int k = 0; k < 34; ++k) { w0[k] = 0; w1[k] = 0; }
      }
    } break;
  }
}

// fb.jobs: n_jobs FinalizeArgs in device memory, or nullptr with the single job inline; max_nq: the largest nq among them
// (> 0).  Jobs with nq == 0 must have had their totals zeroed by the caller (no block of theirs runs).
void launch_finalize(cudaStream_t st, const FinalizeBatch& fb, uint32_t n_jobs, uint32_t max_nq) {
  if (n_jobs == 0 || max_nq == 0) return;
  const dim3 grid((max_nq + 255) / 256, n_jobs);
  launch_pdl(finalize_kernel, grid, dim3(256), 0, st, fb);
}

// keys -> float seed distances (min over splits; FLT_MAX when there is no seed feature)
__global__ void seed_from_keys_kernel(const uint64_t* keys, uint32_t nq, uint32_t nq_pad, uint32_t n_splits, float* out) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  uint64_t key = kKeyNone;
  for (uint32_t s = 0; s < n_splits; ++s) {
    const uint64_t k = keys[static_cast<uint64_t>(s) * nq_pad + i];
    key = k < key ? k : key;
  }
  out[i] = key >= kKeyNeedsExact ? FLT_MAX : static_cast<float>(static_cast<uint32_t>(key >> 37));
}
void launch_seed_from_keys(cudaStream_t st, const uint64_t* keys, uint32_t nq, uint32_t nq_pad, uint32_t n_splits, float* out) {
  if (nq) launch_pdl(seed_from_keys_kernel, dim3((nq + 255) / 256), dim3(256), 0, st, keys, nq, nq_pad, n_splits, out);
}

__global__ void min_keys_kernel(const uint64_t* keys, uint32_t nq, uint32_t nq_pad, uint32_t n_splits, uint64_t* out) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  uint64_t key = kKeyNone;
  for (uint32_t s = 0; s < n_splits; ++s) {
    const uint64_t k = keys[static_cast<uint64_t>(s) * nq_pad + i];
    key = k < key ? k : key;
  }
  out[i] = key;
}
void launch_min_keys(cudaStream_t st, const uint64_t* keys, uint32_t nq, uint32_t nq_pad, uint32_t n_splits, uint64_t* out) {
  if (nq) launch_pdl(min_keys_kernel, dim3((nq + 255) / 256), dim3(256), 0, st, keys, nq, nq_pad, n_splits, out);
}

__global__ void fill_u64_kernel(uint64_t* p, uint64_t n, uint64_t v) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
void launch_fill_u64(cudaStream_t st, uint64_t* p, uint64_t n, uint64_t v) {
  if (n) fill_u64_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p, n, v);
}

// ---------------------------------------------------------------- generic stable compaction of records
// pred: 0 = Match-family `invalid` byte at offset 0 ; 1 = uint2_pair self pair ; 2 = DMatch.distance > threshold
struct CompactArgs {
  const uint8_t* in; uint8_t* out; uint32_t n; uint32_t elt; int32_t pred; float threshold;
  uint32_t* block_counts; uint32_t* total;
};
__device__ __forceinline__ bool compact_keep(const CompactArgs& a, uint32_t i) {
  const uint8_t* r = a.in + static_cast<uint64_t>(i) * a.elt;
  if (a.pred == 0) return r[0] == 0;
  if (a.pred == 1) { const uint32_t* w = reinterpret_cast<const uint32_t*>(r); return !(w[0] == w[2] && w[1] == w[3]); }
  return !(*reinterpret_cast<const float*>(r + 40) > a.threshold);
}
__global__ void __launch_bounds__(256) compact_count_kernel(CompactArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = __syncthreads_count(i < a.n && compact_keep(a, i));
  if (threadIdx.x == 0) a.block_counts[blockIdx.x] = static_cast<uint32_t>(c);
}
__global__ void __launch_bounds__(256) compact_write_kernel(CompactArgs a) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_base;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t part = 0;
  for (uint32_t b = threadIdx.x; b < blockIdx.x; b += blockDim.x) part += a.block_counts[b];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_warp[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += s_warp[w]; s_base = t; }
  __syncthreads();
  const uint32_t base = s_base;
  __syncthreads();
  const bool keep = i < a.n && compact_keep(a, i);
  const uint32_t bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  uint32_t off = base;
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  off += __popc(bal & ((1u << lane) - 1u));
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { uint32_t t = base; for (int w = 0; w < 8; ++w) t += s_warp[w]; *a.total = t; }
  if (!keep) return;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(a.in + static_cast<uint64_t>(i) * a.elt);
  uint32_t* dst = reinterpret_cast<uint32_t*>(a.out + static_cast<uint64_t>(off) * a.elt);
  for (uint32_t k = 0; k < a.elt / 4; ++k) dst[k] = src[k];
}
void launch_compact(cudaStream_t st, CompactArgs a) {
  if (a.n == 0) { cudaMemsetAsync(a.total, 0, sizeof(uint32_t), st); return; }
  const unsigned blocks = (a.n + 255) / 256;
  compact_count_kernel<<<blocks, 256, 0, st>>>(a);
  compact_write_kernel<<<blocks, 256, 0, st>>>(a);
}

// ---------------------------------------------------------------- getRawMatches
__global__ void raw_matches_kernel(const RefDMatch* in, uint32_t n, RefMatch* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t* s = reinterpret_cast<const uint64_t*>(in + i);
  uint64_t* d = reinterpret_cast<uint64_t*>(out + i);
#pragma unroll
  for (int k = 0; k < 5; ++k) d[k] = s[k];
}
void launch_raw_matches(cudaStream_t st, const void* in, uint32_t n, void* out) {
  if (n) raw_matches_kernel<<<(n + 255) / 256, 256, 0, st>>>(static_cast<const RefDMatch*>(in), n, static_cast<RefMatch*>(out));
}

// ---------------------------------------------------------------- 2-view MatchSet (tail of doFeatureMatching)
// ref: src/Pipeline.cu:198-222 -- D2H of the DMatch array, host scan for the largest distance, getRawMatches, D2H of the
// Match array and a host loop writing KeyPoint[2n] / MultiMatch[n] = {2, 2i}.  One pass on the device instead: the MatchSet
// arrays are written straight from the DMatch records and the largest distance is reduced on the way (distances are >= 0,
// so their bit patterns order like the floats; the reference's scan starts at 0.0f and uses `<`, NaN never wins).
struct RefMultiMatch { uint32_t numKeyPoints; int32_t index; };
__global__ void match_set_2view_kernel(const RefDMatch* in, uint32_t n, RefKeyPoint* kp, RefMultiMatch* mm, uint32_t* max_bits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t bits = 0;
  if (i < n) {
    const uint64_t* s = reinterpret_cast<const uint64_t*>(in + i);
    uint64_t* d = reinterpret_cast<uint64_t*>(kp + 2 * static_cast<uint64_t>(i));
    d[0] = s[1]; d[1] = s[2]; d[2] = s[3]; d[3] = s[4];            // keyPoints[0], keyPoints[1] (16 B each, at byte 8 of the record)
    mm[i].numKeyPoints = 2u;
    mm[i].index = static_cast<int32_t>(2 * i);
    const float dist = in[i].distance;
    if (dist > 0.0f) bits = __float_as_uint(dist);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  if ((threadIdx.x & 31) == 0 && bits) atomicMax(max_bits, bits);
}
void launch_match_set_2view(cudaStream_t st, const void* in, uint32_t n, void* kp, void* mm, uint32_t* max_bits) {
  if (n) match_set_2view_kernel<<<(n + 255) / 256, 256, 0, st>>>(static_cast<const RefDMatch*>(in), n, static_cast<RefKeyPoint*>(kp),
                                                                  static_cast<RefMultiMatch*>(mm), max_bits);
}

// ---------------------------------------------------------------- sortMatches: radix sort of (orderable distance bits, index)
__device__ __forceinline__ uint32_t orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void sort_keys_kernel(const RefDMatch* m, uint32_t n, uint64_t* keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (static_cast<uint64_t>(orderable(m[i].distance)) << 32) | i;
}
__global__ void sort_gather_kernel(const RefDMatch* in, const uint64_t* keys, uint32_t n, RefDMatch* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t* s = reinterpret_cast<const uint64_t*>(in + static_cast<uint32_t>(keys[i]));
  uint64_t* d = reinterpret_cast<uint64_t*>(out + i);
#pragma unroll
  for (int w = 0; w < 6; ++w) d[w] = s[w];
}
size_t sort_temp_bytes(uint32_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr), static_cast<int>(n));
  return bytes;
}
// keys_in / keys_out: n u64 each; tmp_records: n DMatch.  The key's low word is the record's position, so equal distances
// keep their input order: deterministic, and one valid outcome of the reference's unstable thrust::sort (ref: :191-222).
cudaError_t launch_sort(cudaStream_t st, void* dmatches, uint32_t n, uint64_t* keys_in, uint64_t* keys_out, void* temp, size_t temp_bytes,
                        void* tmp_records, uint64_t* launches) {
  if (n < 2) return cudaSuccess;
  const unsigned blocks = (n + 255) / 256;
  sort_keys_kernel<<<blocks, 256, 0, st>>>(static_cast<const RefDMatch*>(dmatches), n, keys_in);
  cudaError_t e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys_in, keys_out, static_cast<int>(n), 0, 64, st);
  if (e != cudaSuccess) return e;
  sort_gather_kernel<<<blocks, 256, 0, st>>>(static_cast<const RefDMatch*>(dmatches), keys_out, n, static_cast<RefDMatch*>(tmp_records));
  if (launches) *launches += 3;
  return cudaMemcpyAsync(dmatches, tmp_records, static_cast<size_t>(n) * sizeof(RefDMatch), cudaMemcpyDeviceToDevice, st);
}

}  // namespace b200
