// api.cu -- the C ABI (include/ssrlcv_b200.h): context, packed sets, and the orchestration of
// pack -> tensor-core matcher -> exact tie resolution -> finalize.  Host code only launches our own
// kernels; there is no CPU fallback: without an sm_100 device every entry point fails.
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "matcher.cuh"

namespace b200 {
// pack.cu
void launch_pack(cudaStream_t, const PackJob* jobs, uint32_t n_jobs, int sm_count, uint64_t* launches);
size_t sort_x_temp_bytes(uint32_t n);
bool cell_sort_serves(uint64_t n);
cudaError_t launch_cell_sort_gather(cudaStream_t, const ssrlcv_b200_set_impl& src, ssrlcv_b200_set_impl& dst, uint32_t* perm, uint16_t* cell_of);
cudaError_t launch_sort_gather(cudaStream_t, const ssrlcv_b200_set_impl& src, ssrlcv_b200_set_impl& dst, uint32_t* keys_in, uint32_t* keys_out,
                               uint32_t* vals_in, uint32_t* vals_out, void* temp, size_t temp_bytes);
// smax_tc.cu
cudaError_t launch_smax(cudaStream_t, const SmaxBatch&, int mask_kind, bool fold, bool colmask, bool lean);
cudaError_t launch_smax_pair(cudaStream_t, const SmaxBatch&, bool fold, bool colmask, bool seed_only);
cudaError_t launch_smax_duo(cudaStream_t, const SmaxBatch&, bool fold, bool colmask);
int smax_duo_max_pairs();
// exact.cu
void launch_exact(cudaStream_t, const SmaxArgs&, int mask_kind, int sm_count);
void launch_exact_list(cudaStream_t, const SmaxBatch&, int mask_kind, int sm_count);
void launch_endpoints(cudaStream_t, const MaskParams&, const float2*, uint64_t, float4*);
void launch_query_masks(cudaStream_t, const MaskParams&, const float2*, uint64_t, float4*);
void launch_cull(cudaStream_t, int mask_kind, const float4* q_mask, uint32_t nq, uint32_t n_qtiles, const float4* tbox, uint32_t n_ttiles,
                 float eps, uint32_t* bits, uint32_t stride);
void launch_sincos(cudaStream_t, const float*, uint64_t, float*, float*);
// finalize.cu
struct CompactArgs {
  const uint8_t* in; uint8_t* out; uint32_t n; uint32_t elt; int32_t pred; float threshold;
  uint32_t* block_counts; uint32_t* total;
};
void launch_finalize(cudaStream_t, const FinalizeBatch&, uint32_t, uint32_t);
void launch_seed_from_keys(cudaStream_t, const uint64_t*, uint32_t, uint32_t, uint32_t, float*);
void launch_fill_u64(cudaStream_t, uint64_t*, uint64_t, uint64_t);
void launch_min_keys(cudaStream_t, const uint64_t*, uint32_t, uint32_t, uint32_t, uint64_t*);
void launch_compact(cudaStream_t, CompactArgs);
void launch_raw_matches(cudaStream_t, const void*, uint32_t, void*);
void launch_match_set_2view(cudaStream_t, const void*, uint32_t, void*, void*, uint32_t*);
size_t sort_temp_bytes(uint32_t n);
cudaError_t launch_sort(cudaStream_t, void*, uint32_t, uint64_t*, uint64_t*, void*, size_t, void*, uint64_t*);
}  // namespace b200

using namespace b200;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Workspace {
  void* p = nullptr;
  size_t cap = 0;
};

struct ssrlcv_b200_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  EncodeTiledFn encode = nullptr;
  std::string err;
  uint64_t launches = 0;
  Workspace keys, flagged, counts, scratch, hostpin, qmask, sortws, seedws, cullws;
  // zero-initialised once, left zero by the kernels that use them: tie counters (two banks: a matcher launch counts into
  // one and clears the other for the next call), finalize tickets, finalize block states (stamped with `epoch`, so stale
  // entries read as "not yet")
  Workspace nflag, tickets, fstate;
  uint32_t epoch = 0;
  uint32_t flag_bank = 0;
  uint32_t* d_scalars = nullptr;   // small device scalars
  // job descriptors (SmaxJob / FinalizeArgs arrays) travel host -> device through a small ring of pinned buffers
  struct Staged { void* h = nullptr; void* d = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool pending = false; };
  Staged staged[8];
  unsigned staged_next = 0;
  uint32_t* h_scalars = nullptr;   // pinned mirror
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  // pack() does not synchronise and enqueues nothing but its kernel: a set's meta scalars live in a slot of a pool that is
  // zero between uses (4 x int32 on the device + a pinned, device-mapped host mirror the pack kernel writes its rare "does
  // not fit the fold digits" verdict to) and are read when a match first needs them (set_resolve).  Slots are recycled
  // oldest first, so a slot that comes round again has long been quiescent.
  struct PackSlot { int32_t* mm; cudaEvent_t ready; int32_t* d_meta; };
  std::deque<PackSlot> free_slots;
  std::vector<int32_t*> slot_blocks;
  std::vector<int32_t*> slot_dev_blocks;
  // tuning / diagnostics knobs (ssrlcv_b200_tune); -1 = the library decides
  int64_t knob[SSRLCV_B200_TUNE_COUNT_];
  // seed passes launched with the fold variant before the seed set's pack had reported back (seed_distances does not
  // wait): checked at the context's next synchronisation point, redone with the other variant if the guess was wrong
  struct PendingSeed { const ssrlcv_b200_set* q; const ssrlcv_b200_set* seed; float* out; };
  std::vector<PendingSeed> pending;
  // match_host with large sets: uploads run on their own stream, chunk by chunk, while the matcher already works on the
  // chunks that have arrived (match_host_streamed)
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  Workspace upload;
};

static int slot_acquire(ssrlcv_b200_ctx* ctx, ssrlcv_b200_ctx::PackSlot& out) {
  if (ctx->free_slots.empty()) {
    constexpr int kPerBlock = 256;
    int32_t* block = nullptr;
    int32_t* dblock = nullptr;
    cudaError_t e = cudaHostAlloc(&block, kPerBlock * 4 * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) return static_cast<int>(e);
    memset(block, 0, kPerBlock * 4 * sizeof(int32_t));
    ctx->slot_blocks.push_back(block);
    e = cudaMalloc(&dblock, kPerBlock * 4 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(dblock, 0, kPerBlock * 4 * sizeof(int32_t));
    if (e != cudaSuccess) return static_cast<int>(e);
    ctx->slot_dev_blocks.push_back(dblock);
    for (int i = 0; i < kPerBlock; ++i) {
      ssrlcv_b200_ctx::PackSlot s{block + 4 * i, nullptr, dblock + 4 * i};
      e = cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming);
      if (e != cudaSuccess) return static_cast<int>(e);
      ctx->free_slots.push_back(s);
    }
  }
  out = ctx->free_slots.front();
  ctx->free_slots.pop_front();
  // the slot's previous user: normally finished long ago (oldest-first recycling); its verdict decides whether the device
  // word has to be cleared (only sets that did not fit the fold digits ever dirty it)
  if (cudaEventQuery(out.ready) == cudaErrorNotReady) cudaEventSynchronize(out.ready);
  if (out.mm[0] != 0) {
    cudaMemsetAsync(out.d_meta, 0, 4 * sizeof(int32_t), ctx->stream);
    out.mm[0] = 0;
  }
  return 0;
}

static int fail(ssrlcv_b200_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    c->err = buf;
  }
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(ctx, static_cast<int>(e_), "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static int ws_reserve(ssrlcv_b200_ctx* ctx, Workspace& w, size_t bytes) {
  if (bytes <= w.cap) return 0;
  if (w.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFree(w.p)); w.p = nullptr; w.cap = 0; }
  size_t cap = bytes + bytes / 4 + 256;
  CU(cudaMalloc(&w.p, cap));
  w.cap = cap;
  return 0;
}

// same, and the buffer is zero whenever it is (re)allocated (workspaces whose users leave them zero)
static int ws_reserve_zero(ssrlcv_b200_ctx* ctx, Workspace& w, size_t bytes) {
  if (bytes <= w.cap) return 0;
  int rc = ws_reserve(ctx, w, bytes * 2);
  if (rc) return rc;
  CU(cudaMemsetAsync(w.p, 0, w.cap, ctx->stream));
  return 0;
}

static int verify_pending(ssrlcv_b200_ctx* ctx, bool* redone);

extern "C" {

int ssrlcv_b200_abi_version(void) { return SSRLCV_B200_ABI_VERSION; }

int ssrlcv_b200_create(ssrlcv_b200_ctx** out, int device, void* stream) {
  if (!out) return SSRLCV_B200_EINVAL;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) return SSRLCV_B200_ENODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SSRLCV_B200_ENODEVICE;
  if (prop.major != 10) return SSRLCV_B200_ENODEVICE;   // tcgen05 / TMEM: sm_100 family only
  ssrlcv_b200_ctx* ctx = new (std::nothrow) ssrlcv_b200_ctx();
  if (!ctx) return SSRLCV_B200_ENOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  for (int64_t& k : ctx->knob) k = -1;
  cudaSetDevice(device);
  if (stream) ctx->stream = static_cast<cudaStream_t>(stream);
  else { cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking); ctx->own_stream = true; }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    delete ctx;
    return SSRLCV_B200_ENODEVICE;
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  {  // packed sets come from the stream-ordered pool; keep freed blocks cached instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (cudaMalloc(&ctx->d_scalars, 16 * sizeof(uint32_t)) != cudaSuccess || cudaHostAlloc(&ctx->h_scalars, 16 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    delete ctx;
    return SSRLCV_B200_ENOMEM;
  }
  *out = ctx;
  return 0;
}

void ssrlcv_b200_destroy(ssrlcv_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (Workspace* w : {&ctx->keys, &ctx->flagged, &ctx->counts, &ctx->scratch, &ctx->qmask, &ctx->nflag, &ctx->tickets, &ctx->fstate, &ctx->sortws, &ctx->seedws, &ctx->cullws}) if (w->p) cudaFree(w->p);
  for (auto& st : ctx->staged) {
    if (st.h) cudaFreeHost(st.h);
    if (st.d) cudaFree(st.d);
    if (st.done) cudaEventDestroy(st.done);
  }
  if (ctx->hostpin.p) cudaFreeHost(ctx->hostpin.p);
  if (ctx->upload.p) cudaFree(ctx->upload.p);
  for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& sl : ctx->free_slots) cudaEventDestroy(sl.ready);
  for (int32_t* b : ctx->slot_blocks) cudaFreeHost(b);
  for (int32_t* b : ctx->slot_dev_blocks) cudaFree(b);
  if (ctx->d_scalars) cudaFree(ctx->d_scalars);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  for (auto& pe : ctx->prof_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ssrlcv_b200_last_error(const ssrlcv_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
int ssrlcv_b200_synchronize(ssrlcv_b200_ctx* ctx) {
  if (!ctx) return SSRLCV_B200_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  bool redone = false;
  int rc = verify_pending(ctx, &redone);
  if (rc) return rc;
  if (redone) CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int ssrlcv_b200_tune(ssrlcv_b200_ctx* ctx, int knob, int64_t value) {
  if (!ctx || knob < 0 || knob >= SSRLCV_B200_TUNE_COUNT_) return fail(ctx, SSRLCV_B200_EINVAL, "tune: unknown knob");
  if (knob == SSRLCV_B200_TUNE_DIAG && value > 0 && !kDiagBuild)
    return fail(ctx, SSRLCV_B200_EINVAL, "tune: SSRLCV_B200_TUNE_DIAG needs a diagnostics build of the library (SSRLCV_B200_DIAG_BUILD=1 python -m ssrlcv_b200.build)");
  ctx->knob[knob] = value;
  return 0;
}
uint64_t ssrlcv_b200_launch_count(const ssrlcv_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* ssrlcv_b200_stream(const ssrlcv_b200_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int ssrlcv_b200_set_profiling(ssrlcv_b200_ctx* ctx, int enabled) {
  if (!ctx) return SSRLCV_B200_EINVAL;
  ctx->profiling = enabled != 0;
  return 0;
}
int ssrlcv_b200_kernel_time(ssrlcv_b200_ctx* ctx, double* total_ms, uint64_t* n_launches) {
  if (!ctx || !total_ms || !n_launches) return SSRLCV_B200_EINVAL;
  CU(cudaStreamSynchronize(ctx->stream));
  double t = 0.0;
  for (size_t i = 0; i < ctx->prof_used; ++i) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->prof_events[i].first, ctx->prof_events[i].second));
    t += ms;
  }
  *total_ms = t;
  *n_launches = ctx->prof_used;
  ctx->prof_used = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------- sets
static int encode_map(ssrlcv_b200_ctx* ctx, CUtensorMap* m, void* base, uint64_t rows, uint32_t row_bytes, uint32_t box_rows,
                      CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {row_bytes, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {row_bytes, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, SSRLCV_B200_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// Allocates a set of n rows (one allocation, tensor maps, a slot of the meta pool); nothing is packed yet and nothing is
// enqueued: the pack kernel writes every row including the zero padding and every window bound.
static int set_alloc(ssrlcv_b200_ctx* ctx, uint64_t n, ssrlcv_b200_set** out, bool with_orig = false) {
  ssrlcv_b200_set* s = new (std::nothrow) ssrlcv_b200_set();
  if (!s) return SSRLCV_B200_ENOMEM;
  memset(static_cast<ssrlcv_b200_set_impl*>(s), 0, sizeof(ssrlcv_b200_set_impl));
  s->ctx = ctx;
  s->n = n;
  s->n_pad = ((n + kRowPad - 1) / kRowPad) * kRowPad;
  if (s->n_pad == 0) s->n_pad = kRowPad;
  ssrlcv_b200_ctx::PackSlot slot;
  int rc = slot_acquire(ctx, slot);
  if (rc) { delete s; return fail(ctx, rc, "pack: meta slot: %s", cudaGetErrorString(static_cast<cudaError_t>(rc))); }
  s->h_mm = slot.mm; s->ready = slot.ready; s->meta = slot.d_meta;
  s->resolved = false; s->is_view = false; s->sorted = nullptr;
  // one allocation: desc | ext | norm | wcol | loc | sigtheta | wmax | [orig]   (all segments 32-byte aligned since n_pad % 256 == 0)
  const size_t np = s->n_pad;
  const size_t wmax_bytes = ((np / kChunk) * 4 + 255) / 256 * 256;
  const size_t box_bytes = with_orig ? ((np / kChunk + np / kRowPad) * sizeof(float4) + 255) / 256 * 256 : 0;
  const size_t bytes = np * (kDescBytes + kExtBytes + 4 + 4 + 8 + 8 + (with_orig ? 4 : 0)) + wmax_bytes + box_bytes + 256;
  uint8_t* base = nullptr;
  cudaError_t e = cudaMallocAsync(&base, bytes, ctx->stream);
  if (e != cudaSuccess) {
    ctx->free_slots.push_back(slot); delete s;
    return fail(ctx, static_cast<int>(e), "pack: cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e));
  }
  s->desc = base;
  s->ext = s->desc + np * kDescBytes;
  s->norm = reinterpret_cast<int32_t*>(s->ext + np * kExtBytes);
  s->wcol = s->norm + np;
  s->loc = reinterpret_cast<float2*>(s->wcol + np);
  s->sigtheta = s->loc + np;
  s->wmax = reinterpret_cast<int32_t*>(s->sigtheta + np);
  s->orig = with_orig ? reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s->wmax) + wmax_bytes) : nullptr;
  s->wbox = with_orig ? reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(s->orig) + np * 4) : nullptr;   // np * 4 is a multiple of 1024
  s->tbox = with_orig ? s->wbox + np / kChunk : nullptr;
  s->has_map_a = s->has_map_bx = s->has_map_exth = false;
  *out = s;
  return 0;
}
// TMA descriptors of a set, encoded when the set is first used as a query (A tiles) / as a target (B + digit tiles)
static int ensure_maps(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* cs, bool as_query, bool as_target, bool as_half_target = false) {
  ssrlcv_b200_set* s = const_cast<ssrlcv_b200_set*>(cs);
  int rc = 0;
  if (as_query && !s->has_map_a) {
    rc = encode_map(ctx, &s->map_a, s->desc, s->n_pad, kDescBytes, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    s->has_map_a = true;
  }
  if (as_target && !s->has_map_bx) {
    rc = encode_map(ctx, &s->map_b, s->desc, s->n_pad, kDescBytes, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!rc) rc = encode_map(ctx, &s->map_ext, s->ext, s->n_pad, kExtBytes, kBN, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
    s->has_map_bx = true;
  }
  if (as_half_target && !s->has_map_exth) {              // (its descriptor half-tiles go through map_a: same box)
    rc = encode_map(ctx, &s->map_exth, s->ext, s->n_pad, kExtBytes, kBN / 2, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
    s->has_map_exth = true;
  }
  return 0;
}
// After the last pack launch of a set: the event set_resolve waits for before it reads the verdict from the pinned word.
static int set_finish(ssrlcv_b200_ctx* ctx, ssrlcv_b200_set* s) {
  cudaEventRecord(s->ready, ctx->stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, static_cast<int>(e), "pack: %s", cudaGetErrorString(e));
  return 0;
}
// rows [row0, row0 + n_rows) of `s` <- n records at feat_dev (+ zero padding); row0 % 32 == 0, n_rows % 32 == 0
static PackJob pack_job(const ssrlcv_b200_set* s, const void* feat_dev, uint64_t n, uint64_t n_rows, uint64_t row0) {
  PackJob j;
  j.in = static_cast<const RefFeature*>(feat_dev);
  j.n = n; j.n_rows = n_rows;
  j.desc = s->desc + row0 * kDescBytes; j.ext = s->ext + row0 * kExtBytes;
  j.norm = s->norm + row0; j.wcol = s->wcol + row0; j.wmax = s->wmax + row0 / kChunk;
  j.loc = s->loc + row0; j.sigtheta = s->sigtheta + row0;
  j.meta = s->meta; j.h_meta = s->h_mm;
  j.win_begin = 0;
  return j;
}

// Enqueue-only: nothing here waits for the device, and the pack kernel is the only node that enters the stream.  Whether
// every norm of a set fits the fold digits, which selects the kernel variant, is read by set_resolve() from the set's pinned
// word when a match first needs it.  All sets of the call share ONE launch (N-view jobs: one launch for all views).
int ssrlcv_b200_pack_batch(ssrlcv_b200_ctx* ctx, uint32_t n_sets, const void* const* features_dev, const uint64_t* n,
                           ssrlcv_b200_set** out) {
  if (!ctx || (n_sets && (!features_dev || !n || !out))) return fail(ctx, SSRLCV_B200_EINVAL, "pack_batch: bad arguments");
  for (uint32_t i = 0; i < n_sets; ++i) {
    out[i] = nullptr;
    if ((n[i] && !features_dev[i]) || n[i] > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "pack_batch: set %u: bad arguments", i);
  }
  CU(cudaSetDevice(ctx->device));
  std::vector<PackJob> jobs(n_sets);
  int rc = 0;
  for (uint32_t i = 0; i < n_sets && !rc; ++i) {
    rc = set_alloc(ctx, n[i], &out[i]);
    if (!rc) jobs[i] = pack_job(out[i], features_dev[i], n[i], out[i]->n_pad, 0);
  }
  if (!rc) {
    launch_pack(ctx->stream, jobs.data(), n_sets, ctx->sm_count, &ctx->launches);
    for (uint32_t i = 0; i < n_sets && !rc; ++i) rc = set_finish(ctx, out[i]);
  }
  if (rc) {
    for (uint32_t i = 0; i < n_sets; ++i) { ssrlcv_b200_set_free(out[i]); out[i] = nullptr; }
    return rc;
  }
  return 0;
}

// ---- producer hand-off: the feature factory writes descriptors / locations straight into a packed set -------------
int ssrlcv_b200_set_alloc(ssrlcv_b200_ctx* ctx, uint64_t n, ssrlcv_b200_set** out) {
  if (!ctx || !out || n > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "set_alloc: bad arguments");
  CU(cudaSetDevice(ctx->device));
  return set_alloc(ctx, n, out);
}
int ssrlcv_b200_set_arrays_of(const ssrlcv_b200_set* s, ssrlcv_b200_set_arrays* out) {
  if (!s || !out) return SSRLCV_B200_EINVAL;
  out->desc = s->desc; out->loc = reinterpret_cast<float*>(s->loc); out->sigma_theta = reinterpret_cast<float*>(s->sigtheta);
  out->n = s->n; out->n_pad = s->n_pad;
  return 0;
}
// Enqueue-only: one launch derives norms, fold digits, W and the window bounds from the rows the producer wrote and zero-fills
// the padding; the producer's kernel must be ordered before the context's stream at this point.
int ssrlcv_b200_set_commit(ssrlcv_b200_ctx* ctx, ssrlcv_b200_set* s) {
  if (!ctx || !s || s->ctx != ctx || s->is_view) return fail(ctx, SSRLCV_B200_EINVAL, "set_commit: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const PackJob j = pack_job(s, nullptr, s->n, s->n_pad, 0);
  launch_pack(ctx->stream, &j, 1, ctx->sm_count, &ctx->launches);
  s->resolved = false;
  return set_finish(ctx, s);
}

int ssrlcv_b200_pack(ssrlcv_b200_ctx* ctx, const void* features_dev, uint64_t n, ssrlcv_b200_set** out) {
  if (!ctx || !out) return fail(ctx, SSRLCV_B200_EINVAL, "pack: bad arguments");
  return ssrlcv_b200_pack_batch(ctx, 1, &features_dev, &n, out);
}

// first use of a set's norm range on the host: wait for its pack (normally long finished)
static int set_resolve(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* cs) {
  ssrlcv_b200_set* s = const_cast<ssrlcv_b200_set*>(cs);
  if (s->resolved) return 0;
  CU(cudaEventSynchronize(s->ready));
  s->fold_ok = s->h_mm[0] == 0;
  s->resolved = true;
  return 0;
}

// the same without waiting: true if the verdict is known now
static bool set_try_resolve(const ssrlcv_b200_set* cs) {
  ssrlcv_b200_set* s = const_cast<ssrlcv_b200_set*>(cs);
  if (s->resolved) return true;
  if (cudaEventQuery(s->ready) != cudaSuccess) { cudaGetLastError(); return false; }
  s->fold_ok = s->h_mm[0] == 0;
  s->resolved = true;
  return true;
}

static int pack_host_async(ssrlcv_b200_ctx* ctx, const void* features_host, uint64_t n, ssrlcv_b200_set** out) {
  CU(cudaSetDevice(ctx->device));
  const size_t bytes = static_cast<size_t>(n) * sizeof(RefFeature);
  int rc = ws_reserve(ctx, ctx->scratch, bytes ? bytes : 16);
  if (rc) return rc;
  if (bytes) CU(cudaMemcpyAsync(ctx->scratch.p, features_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return ssrlcv_b200_pack(ctx, ctx->scratch.p, n, out);   // stream order keeps the scratch buffer safe to reuse
}

int ssrlcv_b200_pack_host(ssrlcv_b200_ctx* ctx, const void* features_host, uint64_t n, ssrlcv_b200_set** out) {
  if (!ctx || !out || (n && !features_host)) return fail(ctx, SSRLCV_B200_EINVAL, "pack_host: bad arguments");
  int rc = pack_host_async(ctx, features_host, n, out);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));                  // the caller's host buffer is free again on return
  return 0;
}

uint64_t ssrlcv_b200_set_size(const ssrlcv_b200_set* s) { return s ? s->n : 0; }
const float* ssrlcv_b200_set_loc(const ssrlcv_b200_set* s) { return s ? reinterpret_cast<const float*>(s->loc) : nullptr; }
void ssrlcv_b200_set_free(ssrlcv_b200_set* s) {
  if (!s) return;
  ssrlcv_b200_ctx* ctx = s->ctx;
  cudaSetDevice(ctx->device);
  if (!s->is_view && !ctx->pending.empty()) {
    // a seed pass that still awaits its verdict reads this set: settle it while the set is alive
    bool touched = false;
    for (const auto& p : ctx->pending) touched |= p.q == s || p.seed == s;
    if (touched) { bool redone = false; verify_pending(ctx, &redone); }
  }
  if (s->is_view) return;                              // borrowed rows of another set (match_host_streamed): nothing to release
  if (s->sorted) { ssrlcv_b200_set_free(s->sorted); s->sorted = nullptr; }
  if (s->desc) cudaFreeAsync(s->desc, ctx->stream);   // stream-ordered: every kernel that reads the set was enqueued before this
  if (s->ready) ctx->free_slots.push_back(ssrlcv_b200_ctx::PackSlot{s->h_mm, s->ready, s->meta});
  delete s;
}

// ------------------------------------------------------------------------------------------- matching
// ref: src/Image.cu:498-539 -- host code, glibc trig, every operation separately rounded (no FMA):
// compiled here with -ffp-contract=off semantics by writing through volatile temporaries.
static void projection_matrix_host(const RefCamera& cam, float P[12]) {
  volatile float cx = cosf(cam.cam_rot[0]), sx = sinf(cam.cam_rot[0]);
  volatile float cy = cosf(cam.cam_rot[1]), sy = sinf(cam.cam_rot[1]);
  volatile float cz = cosf(cam.cam_rot[2]), sz = sinf(cam.cam_rot[2]);
  volatile float K[3][3] = {{cam.foc / cam.dpix[0], 0.0f, static_cast<float>(cam.size[0]) / 2.0f},
                            {0.0f, cam.foc / cam.dpix[1], static_cast<float>(cam.size[1]) / 2.0f},
                            {0.0f, 0.0f, 1.0f}};
  volatile float R[3][4];
  volatile float t1, t2, t3;
  R[0][0] = cz * cy; R[0][1] = sz * cy; R[0][2] = -1 * sy; R[0][3] = 0.0f;
  t1 = cz * sy; t1 = t1 * sx; t2 = sz * cx; R[1][0] = t1 - t2;
  t1 = sz * sy; t1 = t1 * sx; t2 = cz * cx; R[1][1] = t1 + t2;
  R[1][2] = cy * sx; R[1][3] = 0.0f;
  t1 = cz * sy; t1 = t1 * cx; t2 = sz * sx; R[2][0] = t1 + t2;
  t1 = sz * sy; t1 = t1 * cx; t2 = cz * sx; R[2][1] = t1 - t2;
  R[2][2] = cy * cx; R[2][3] = 0.0f;
  volatile float ex = cam.cam_pos[0] + cam.ecef_offset[0];
  volatile float ey = cam.cam_pos[1] + cam.ecef_offset[1];
  volatile float ez = cam.cam_pos[2] + cam.ecef_offset[2];
  for (int i = 0; i < 3; ++i) {
    t1 = R[i][0] * ex; t2 = R[i][1] * ey; t3 = R[i][2] * ez;
    t1 = t1 + t2; t1 = t1 + t3;
    R[i][3] = R[i][3] - t1;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      t1 = K[r][0] * R[0][c]; t2 = K[r][1] * R[1][c]; t3 = K[r][2] * R[2][c];
      t1 = t1 + t2; t1 = t1 + t3;
      P[r * 4 + c] = t1;
    }
}

// which ratio the reference kernel of (out_kind, mask_kind) applies
static float ratio_threshold(int out_kind, int mask_kind, float rel) {
  volatile float rel2 = rel * rel;
  switch (out_kind) {
    case SSRLCV_B200_OUT_MATCH: return mask_kind == SSRLCV_B200_MASK_DOUBLE ? rel2 : rel;
    case SSRLCV_B200_OUT_PAIR: return rel;
    default: return rel2;
  }
}

static void fill_mask(const ssrlcv_b200_match_params* p, MaskParams& mp) {
  memset(&mp, 0, sizeof mp);
  mp.kind = p->mask_kind;
  mp.eps = p->epsilon;
  mp.delta = p->delta;
  memcpy(mp.F, p->fundamental, sizeof mp.F);
  memcpy(&mp.qcam, p->query_camera, sizeof mp.qcam);
  if (p->mask_kind == SSRLCV_B200_MASK_DOUBLE) {
    RefCamera tcam;
    memcpy(&tcam, p->target_camera, sizeof tcam);
    projection_matrix_host(tcam, mp.P);
  }
}

// ---- job staging ------------------------------------------------------------------------------------------------
// Returns a pinned buffer of `bytes` (fill it, then stage_upload) whose device mirror the kernels of this call read.
static int stage_acquire(ssrlcv_b200_ctx* ctx, size_t bytes, ssrlcv_b200_ctx::Staged** out) {
  ssrlcv_b200_ctx::Staged& st = ctx->staged[ctx->staged_next++ % 8];
  if (st.pending) { CU(cudaEventSynchronize(st.done)); st.pending = false; }   // its previous upload has long finished
  if (!st.done) CU(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
  if (bytes > st.cap) {
    if (st.h) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFreeHost(st.h)); CU(cudaFree(st.d)); st.h = st.d = nullptr; st.cap = 0; }
    const size_t cap = bytes + bytes / 2 + 4096;
    CU(cudaMallocHost(&st.h, cap));
    CU(cudaMalloc(&st.d, cap));
    st.cap = cap;
  }
  *out = &st;
  return 0;
}
static int stage_upload(ssrlcv_b200_ctx* ctx, ssrlcv_b200_ctx::Staged* st, size_t bytes) {
  CU(cudaMemcpyAsync(st->d, st->h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(st->done, ctx->stream));
  st->pending = true;
  return 0;
}

// ---- planning ---------------------------------------------------------------------------------------------------
// SIMT engine: fixed-length target splits so that (queries x splits) fills the device
static void plan_simt(const ssrlcv_b200_ctx* ctx, SmaxArgs& a) {
  const uint32_t sms = static_cast<uint32_t>(ctx->sm_count);
  uint32_t s = 1;
  while (static_cast<uint64_t>(a.nq) * s < 8ull * sms && s * 4 <= a.n_ttiles) s *= 2;
  a.tiles_per_split = (a.n_ttiles + s - 1) / s;
  if (a.tiles_per_split == 0) a.tiles_per_split = 1;
  a.n_splits = a.n_ttiles ? (a.n_ttiles + a.tiles_per_split - 1) / a.tiles_per_split : 1;
  a.n_ctas = 1;
  a.total_tiles = static_cast<uint64_t>(a.n_qtiles) * a.n_ttiles;
  a.tile_begin = 0; a.tile_end = a.total_tiles;
}

// Tensor-core engine: the tile spaces (query tile x target tile) of all jobs of a launch are laid end to end and cut into
// one contiguous, equally long range per CTA (matcher.cuh).  Every SM gets the same number of tiles (+-1) whatever the
// shapes; a query tile is shared by the few CTAs whose ranges cross it, each writing its own key slot.
static uint32_t launch_ctas(const ssrlcv_b200_ctx* ctx, uint64_t total_tiles, bool cta_pairs = false) {
  const uint64_t min_tiles = 4;                         // per-unit fixed cost (A tile, pipeline fill, resolve) ~ a few tiles
  const uint64_t most = cta_pairs ? static_cast<uint64_t>(smax_duo_max_pairs()) : static_cast<uint64_t>(ctx->sm_count);
  uint64_t g = total_tiles / min_tiles;
  if (g > most) g = most;
  if (g < 1) g = 1;
  return static_cast<uint32_t>(g);
}
// key slots of a job = the widest crossing of one of its query tiles (total_tiles, n_ctas, tile_begin already set)
static uint32_t job_slots(const SmaxArgs& a) {
  const uint64_t per = a.total_tiles / a.n_ctas;        // shortest range
  uint32_t slots = static_cast<uint32_t>((a.n_ttiles + per - 1) / per) + 1;
  if (slots > a.n_ctas) slots = a.n_ctas;
  if (a.n_qtiles <= 8192) {                             // exact
    uint32_t widest = 1;
    for (uint32_t qt = 0; qt < a.n_qtiles; ++qt) {
      const uint64_t base = a.tile_begin + static_cast<uint64_t>(qt) * a.n_ttiles;
      const uint32_t w = cta_of_tile(a.total_tiles, a.n_ctas, base + a.n_ttiles - 1) - cta_of_tile(a.total_tiles, a.n_ctas, base) + 1;
      if (w > widest) widest = w;
    }
    slots = widest;
  }
  return slots;
}

// largest integer d2 with (float)d2 < abs_thr, i.e. that the reference's `d2 >= abs` test lets through
// (ref: src/MatchFactory.cu:1472, :1494); INT32_MAX when the threshold cannot bind (d2 <= 128 * 255^2)
static int32_t d2_cap_of(float abs_thr) {
  if (!(abs_thr > 0.0f)) return -1;                   // also NaN: nothing passes
  if (abs_thr > 8323200.0f) return 0x7fffffff;
  const float c = ceilf(abs_thr);
  return static_cast<int32_t>(c) - 1;
}

// The spatially sorted copy of a packed set (cached on the set, built on first use; enqueue-only).  Rows are ordered along a
// Z-curve of loc (NaN locations last); `orig` maps a row back to the caller's index; wbox / tbox hold the bounding boxes of its
// 32-row windows and 256-row tiles.  Used on BOTH sides of masked searches that have no seed distances to bound their
// candidates (matcher.cuh t_orig / q_orig / live_bits).
static int sorted_of(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* ct, const ssrlcv_b200_set** out) {
  ssrlcv_b200_set* t = const_cast<ssrlcv_b200_set*>(ct);
  if (t->sorted) { *out = t->sorted; return 0; }
  int rc = set_resolve(ctx, t);                            // same rows, same fold verdict: the copy starts out resolved
  if (rc) return rc;
  ssrlcv_b200_set* s2 = nullptr;
  rc = set_alloc(ctx, t->n, &s2, true);
  if (rc) return rc;
  const size_t n4 = ((static_cast<size_t>(t->n) * 4 + 255) / 256) * 256;
  const size_t temp = sort_x_temp_bytes(static_cast<uint32_t>(t->n));
  rc = ws_reserve(ctx, ctx->sortws, 4 * n4 + temp + 256);
  if (rc) { ssrlcv_b200_set_free(s2); return rc; }
  uint8_t* w = static_cast<uint8_t*>(ctx->sortws.p);
  cudaError_t e = cudaSuccess;
  bool cells = cell_sort_serves(t->n);                     // one-block counting sort over image cells + one copy launch (pack.cu)
  if (ctx->knob[SSRLCV_B200_TUNE_CELL_SORT] >= 0) cells = cells && ctx->knob[SSRLCV_B200_TUNE_CELL_SORT] != 0;
  if (cells) {
    e = launch_cell_sort_gather(ctx->stream, *t, *s2, reinterpret_cast<uint32_t*>(w), reinterpret_cast<uint16_t*>(w + n4));
    ctx->launches += 2;
  } else {
    e = cudaMemsetAsync(s2->wmax, 0, (s2->n_pad / kChunk) * 4, ctx->stream);   // the gather accumulates the window bounds with atomicMax
    if (e == cudaSuccess)
      e = launch_sort_gather(ctx->stream, *t, *s2, reinterpret_cast<uint32_t*>(w), reinterpret_cast<uint32_t*>(w + n4),
                             reinterpret_cast<uint32_t*>(w + 2 * n4), reinterpret_cast<uint32_t*>(w + 3 * n4), w + 4 * n4, temp);
    ctx->launches += 4;
  }
  if (e == cudaSuccess && !t->fold_ok) {
    e = cudaMemsetAsync(s2->meta, 1, sizeof(int32_t), ctx->stream);
    s2->h_mm[0] = 1;
  }
  if (e != cudaSuccess) { ssrlcv_b200_set_free(s2); return fail(ctx, static_cast<int>(e), "sorted copy: %s", cudaGetErrorString(e)); }
  rc = set_finish(ctx, s2);
  if (rc) { ssrlcv_b200_set_free(s2); return rc; }
  s2->fold_ok = t->fold_ok; s2->resolved = true;
  t->sorted = s2;
  *out = s2;
  return 0;
}

// One pair of a batch: inputs, and (after engine_enqueue) where its keys are.
static inline int engine_of(const ssrlcv_b200_ctx* ctx) {
  return ctx->knob[SSRLCV_B200_TUNE_CROSSCHECK_SIMT] == 1 ? SSRLCV_B200_ENGINE_SIMT : SSRLCV_B200_ENGINE_TCGEN05;
}

struct PairWork {
  const ssrlcv_b200_set* q = nullptr;
  const ssrlcv_b200_set* t = nullptr;
  MaskParams mp;
  int engine = 0;                 // filled by engine_enqueue (the context's cross-check knob)
  uint64_t target_offset = 0;
  float abs_thr = 0.f;
  const float* seed = nullptr;    // seed distances of the query set (device) when the ratio test will drop what exceeds it
  float seed_ratio = 0.f;
  const uint64_t* exclude = nullptr;   // runner-up search: winner keys of a previous search of the same pair (matcher.cuh q_exclude)
  bool seed_only = false;         // the caller wants each query's smallest d2 and nothing else (seed passes): keys may carry f = 0
  bool optimistic_fold = false;   // out: the fold variant was launched before the target set's pack had reported back
  const ssrlcv_b200_set* tt = nullptr;   // out: the set the kernel reads its targets from (t, or t's spatially sorted copy)
  const ssrlcv_b200_set* qq = nullptr;   // out: the same for the queries
  SmaxArgs a;                     // out: a.keys, a.n_splits, a.nq_pad describe the pair's candidate keys
};

// Runs the distance engine over n pairs: exact candidate keys land in w[i].a.keys ([n_splits][nq_pad]).  All tensor-core
// pairs with the same (mask kind, fold variant) share ONE persistent launch; nothing here waits for the device.
// optimistic != NULL: a target set whose pack has not reported back yet is not waited for; the fold variant is launched
// (the kernel skips the job if the set turns out not to fit the fold digits) and *optimistic tells the caller to check
// those sets once the stream has been synchronised.  optimistic == NULL: wait for the sets' packs first.
static int engine_enqueue(ssrlcv_b200_ctx* ctx, PairWork* w, uint32_t n, bool* optimistic) {
  struct Group { std::vector<uint32_t> idx; uint64_t total = 0; };
  // how a job walks the score matrix = which kernel family / instantiation serves it (one launch per non-empty group)
  enum Walk { kWalkTile = 0,       // smax_tc.cu: one query tile per unit (every mask kind, any size)
              kWalkPair = 1,       // smax_pair.cu: two query tiles per unit (large unmasked searches)
              kWalkDuo = 2,        // smax_duo.cu: four per CTA pair, tcgen05 cta_group::2 (knob only: measured slower)
              kWalkPairSeed = 3,   // smax_pair.cu kSeedOnly: seed passes, distances only
              kWalkTileLean = 4,   // smax_tc.cu kLean: seeded masked searches, unmasked hot loop
              kWalks = 5 };
  constexpr int kGroups = 3 * 4 * kWalks;                 // [mask kind][fold][column-mask regime][walk]
  Group groups[kGroups];
  size_t keys_total = 0, qmask_total = 0;
  if (optimistic) *optimistic = false;
  for (uint32_t i = 0; i < n; ++i) {
    PairWork& p = w[i];
    SmaxArgs& a = p.a;
    memset(&a, 0, sizeof a);
    p.optimistic_fold = false;
    p.engine = engine_of(ctx);
    a.nq = static_cast<uint32_t>(p.q->n); a.nt = static_cast<uint32_t>(p.t->n); a.nq_pad = static_cast<uint32_t>(p.q->n_pad);
    a.n_ttiles = (a.nt + kBN - 1) / kBN;
    // Large unmasked searches walk the score matrix two query tiles at a time (smax_pair.cu: every target tile is staged once
    // per 256 queries -- half the L2 -> SM traffic -- and each accumulator has its own epilogue warps); everything else, and
    // pairs too small to balance 256-row units over the SMs, one tile at a time (smax_tc.cu).
    bool pair_units = p.engine != SSRLCV_B200_ENGINE_SIMT && p.mp.kind == SSRLCV_B200_MASK_NONE && a.nq_pad % (2 * kBM) == 0 && p.exclude == nullptr &&
                      static_cast<uint64_t>((a.nq + 2 * kBM - 1) / (2 * kBM)) * a.n_ttiles >= 8ull * static_cast<uint64_t>(ctx->sm_count);
    if (ctx->knob[SSRLCV_B200_TUNE_PAIR_UNITS] >= 0)
      pair_units = ctx->knob[SSRLCV_B200_TUNE_PAIR_UNITS] != 0 && p.engine != SSRLCV_B200_ENGINE_SIMT && p.mp.kind == SSRLCV_B200_MASK_NONE &&
                   a.nq_pad % (2 * kBM) == 0 && p.exclude == nullptr;
    // ... and, on a CTA pair (smax_duo.cu, tcgen05 cta_group::2), four query tiles per unit with every target tile staged half in
    // each SM of the pair: bit-exact, measured slower (the accumulator hand-off crosses the pair twice), only ever run by the knob
    bool duo_units = false;
    if (ctx->knob[SSRLCV_B200_TUNE_PAIR_UNITS] == 2) duo_units = pair_units && smax_duo_max_pairs() > 0;
    int walk = duo_units ? kWalkDuo : (pair_units ? kWalkPair : kWalkTile);
    a.q_tile_rows = duo_units ? 4 * kBM : (pair_units ? 2 * kBM : kBM);
    a.n_qtiles = (a.nq + a.q_tile_rows - 1) / a.q_tile_rows;
    const uint64_t tiles = static_cast<uint64_t>(a.n_qtiles) * a.n_ttiles;
    if (tiles >= (1ull << 32))
      return fail(ctx, SSRLCV_B200_EINVAL, "%u x %u features is too large for one call: shard the target set (ssrlcv_b200_match_candidates)", a.nq, a.nt);
    a.target_offset = p.target_offset;
    a.eps = p.mp.eps;
    a.d2_cap = d2_cap_of(p.abs_thr);
    a.q_seed = p.seed; a.seed_ratio = p.seed_ratio;
    a.q_exclude = p.exclude;
    p.tt = p.t; p.qq = p.q;
    if (p.engine != SSRLCV_B200_ENGINE_SIMT && tiles != 0 && p.mp.kind != SSRLCV_B200_MASK_NONE && p.seed == nullptr && p.exclude == nullptr &&
        !p.t->is_view && !p.q->is_view) {
      // Masked search with nothing but the mask to prune it: both sets are matched in spatial (Z-curve) order, so that an
      // epipolar band misses most 32-target windows and most (query tile, target tile) pairs outright.  Worth its launches
      // (~4 per set, once per set) only for pairs whose matcher runs much longer than that.
      bool use_sorted = a.nt >= 16384 && static_cast<uint64_t>(a.nq) * a.nt >= (1ull << 28);
      if (ctx->knob[SSRLCV_B200_TUNE_SORTED_TARGETS] >= 0) use_sorted = ctx->knob[SSRLCV_B200_TUNE_SORTED_TARGETS] != 0;
      if (use_sorted) {
        int rc = sorted_of(ctx, p.t, &p.tt);
        if (!rc) rc = sorted_of(ctx, p.q, &p.qq);
        if (rc) return rc;
      }
    }
    a.t_meta = p.tt->meta;
    a.t_orig = p.tt->orig;
    a.t_wbox = p.tt->wbox;
    a.q_orig = p.qq->orig;
    a.q_desc = p.qq->desc; a.q_norm = p.qq->norm; a.q_loc = p.qq->loc;
    a.t_desc = p.tt->desc; a.t_norm = p.tt->norm; a.t_loc = p.tt->loc; a.t_wcol = p.tt->wcol; a.t_wmax = p.tt->wmax;
    if (p.engine == SSRLCV_B200_ENGINE_SIMT || tiles == 0) {
      plan_simt(ctx, a);
    } else {
      // Variant (smax_tc.cu): the fold kernel (five MMAs per tile, the accumulator IS the score) whenever the target set's
      // norms fit the fold digits; otherwise the no-fold kernel (four MMAs, W added in the epilogue to the few windows whose
      // bound reaches a running best), which serves any norm range.  Measured at 65 536 x 65 536 (profiles/r1f_fold_ab.txt):
      // fold 0.371 ms, no-fold 0.431 ms -- with four MMAs the tile period drops below the epilogue's own time per tile and
      // every rare bookkeeping event of any of the 8 epilogue warps stalls the two-accumulator hand-off; the fifth MMA is the
      // slack that hides them.  ssrlcv_b200_tune(SSRLCV_B200_TUNE_FOLD) forces a variant (tests, A/B timing).
      bool fold = ctx->knob[SSRLCV_B200_TUNE_FOLD] != 0;
      if (fold) {
        if (set_try_resolve(p.t) || !optimistic) {
          int rc = set_resolve(ctx, p.t);
          if (rc) return rc;
          fold = p.t->fold_ok;
        } else {
          *optimistic = true;
          p.optimistic_fold = true;
        }
      }
      // bookkeeping regime (smax_tc.cu): exact column masks for thresholded searches and short target sets
      bool colmask = a.d2_cap != 0x7fffffff || a.q_seed != nullptr || a.n_ttiles < 64;
      if (ctx->knob[SSRLCV_B200_TUNE_COLUMN_MASKS] >= 0) colmask = ctx->knob[SSRLCV_B200_TUNE_COLUMN_MASKS] != 0;   // A/B timing, tests
      // Seed passes on the pair walk (smax_pair.cu kSeedOnly): an uncapped search is bookkeeping-bound, a minimum distance needs no
      // bookkeeping -- the fold variant maximises 2S - (parity of |b|^2) branch-free and writes keys without an index
      // (up to 65 536 seed features: record updates thin out as 1 / position, and beyond that the ordinary search's common path beats an
      // epilogue that is bound by its shared-memory loads, profiles/r2_seed_pass_scaling.txt)
      const bool seed_variant = ctx->knob[SSRLCV_B200_TUNE_SEED_ONLY] >= 0 ? ctx->knob[SSRLCV_B200_TUNE_SEED_ONLY] != 0 : a.n_ttiles <= 256;
      if (walk == kWalkPair && fold && p.seed_only && a.q_seed == nullptr && a.d2_cap == 0x7fffffff && seed_variant) {
        walk = kWalkPairSeed;
        colmask = false;
      }
      // Seeded masked searches (the reference pipeline's call): the bookkeeping path is next to never taken, so they run the
      // instantiation with the unmasked hot loop (smax_tc.cu kLean: no column ring, no boxes, no dead-tile bitmap)
      if (walk == kWalkTile && p.mp.kind != SSRLCV_B200_MASK_NONE && fold && colmask && a.q_seed != nullptr && p.tt->orig == nullptr && p.qq->orig == nullptr &&
          p.exclude == nullptr && ctx->knob[SSRLCV_B200_TUNE_LEAN_MASKED] != 0)
        walk = kWalkTileLean;
      Group& g = groups[kWalks * (p.mp.kind * 4 + (fold ? 2 : 0) + (colmask ? 1 : 0)) + walk];
      a.tile_begin = g.total;
      g.total += tiles;
      a.tile_end = g.total;
      g.idx.push_back(i);
    }
  }
  uint32_t n_tc = 0;
  for (int gi2 = 0; gi2 < kGroups; ++gi2) {
    Group& g = groups[gi2];
    if (g.idx.empty()) continue;
    const uint32_t n_ctas = launch_ctas(ctx, g.total, gi2 % kWalks == kWalkDuo);
    for (uint32_t i : g.idx) {
      SmaxArgs& a = w[i].a;
      a.total_tiles = g.total; a.n_ctas = n_ctas; a.tiles_per_split = a.n_ttiles;
      a.n_splits = job_slots(a);
    }
    n_tc += static_cast<uint32_t>(g.idx.size());
  }
  // ---- workspaces: one region per pair in each
  std::vector<size_t> keys_off(n), qmask_off(n), cull_off(n);
  size_t cull_total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    keys_off[i] = keys_total;
    keys_total += static_cast<size_t>(w[i].a.n_splits) * w[i].a.nq_pad;
    qmask_off[i] = qmask_total;
    if (w[i].mp.kind != SSRLCV_B200_MASK_NONE && w[i].a.nq && w[i].a.nt) qmask_total += static_cast<size_t>(w[i].a.nq_pad) * 2;
    cull_off[i] = cull_total;
    if (w[i].tt->tbox != nullptr && w[i].qq->orig != nullptr && w[i].a.n_ttiles <= 32 * 8192) {   // both sides spatially sorted: find the dead tile pairs (bitmap row <= 32 KB of shared memory)
      w[i].a.live_stride = (w[i].a.n_ttiles + 31) / 32;
      cull_total += static_cast<size_t>(w[i].a.n_qtiles) * w[i].a.live_stride;
    }
  }
  int rc = ws_reserve(ctx, ctx->keys, keys_total * sizeof(uint64_t));
  if (!rc) rc = ws_reserve(ctx, ctx->flagged, keys_total * sizeof(uint32_t));
  if (!rc && qmask_total) rc = ws_reserve(ctx, ctx->qmask, qmask_total * sizeof(float4));
  if (!rc && n_tc) rc = ws_reserve_zero(ctx, ctx->nflag, 2 * static_cast<size_t>(n_tc) * sizeof(uint32_t));
  if (!rc && cull_total) rc = ws_reserve(ctx, ctx->cullws, cull_total * sizeof(uint32_t));
  if (rc) return rc;
  for (uint32_t i = 0; i < n; ++i) {
    SmaxArgs& a = w[i].a;
    a.keys = static_cast<uint64_t*>(ctx->keys.p) + keys_off[i];
    a.flagged = static_cast<uint32_t*>(ctx->flagged.p) + keys_off[i];
    // The tensor-core kernel writes every key slot of every query tile itself (units that end a tile also mark the slots
    // no unit owns); only pairs it never sees -- empty target sets, whose single slot must read "no candidate" -- are preset.
    const bool tc = w[i].engine != SSRLCV_B200_ENGINE_SIMT && a.n_qtiles != 0 && a.n_ttiles != 0;
    if (!tc && a.nt == 0 && a.nq_pad)
      CU(cudaMemsetAsync(a.keys, 0xFF, static_cast<size_t>(a.n_splits) * a.nq_pad * sizeof(uint64_t), ctx->stream));
  }
  // ---- per-query mask state, once per masked pair (all padded rows: loc = 0 there)
  for (uint32_t i = 0; i < n; ++i) {
    PairWork& p = w[i];
    if (p.mp.kind == SSRLCV_B200_MASK_NONE || p.a.nq == 0 || p.a.nt == 0) continue;
    float4* qm = static_cast<float4*>(ctx->qmask.p) + qmask_off[i];
    launch_query_masks(ctx->stream, p.mp, p.qq->loc, p.a.nq_pad, qm);
    p.a.q_mask = qm;
    ctx->launches += 1;
    if (p.a.live_stride) {                               // which (query tile, target tile) pairs can hold a match at all
      uint32_t* bits = static_cast<uint32_t*>(ctx->cullws.p) + cull_off[i];
      launch_cull(ctx->stream, p.mp.kind, qm, p.a.nq, p.a.n_qtiles, p.tt->tbox, p.a.n_ttiles, p.mp.eps, bits, p.a.live_stride);
      p.a.live_bits = bits;
      ctx->launches += 1;
    }
  }
  // ---- tensor-core launches: one per (mask kind, fold variant, bookkeeping regime) group
  if (n_tc) {
    // The job table of a single pair rides in the kernel's parameter block; several pairs are staged through a pinned ring
    // (one small copy for the whole batch).  Tie counters: persistent words in two banks of the zero-initialised workspace;
    // this call counts into one bank, its matcher launches clear the other for the call after it.
    const uint32_t bank_len = static_cast<uint32_t>(ctx->nflag.cap / (2 * sizeof(uint32_t)));
    ctx->flag_bank ^= 1u;
    uint32_t* d_nflag = static_cast<uint32_t*>(ctx->nflag.p) + ctx->flag_bank * bank_len;
    uint32_t* d_nflag_other = static_cast<uint32_t*>(ctx->nflag.p) + (ctx->flag_bank ^ 1u) * bank_len;
    ssrlcv_b200_ctx::Staged* st = nullptr;
    SmaxJob* hj = nullptr;
    SmaxJob single;
    if (n_tc > 1) {
      rc = stage_acquire(ctx, static_cast<size_t>(n_tc) * sizeof(SmaxJob), &st);
      if (rc) return rc;
      hj = static_cast<SmaxJob*>(st->h);
    }
    uint32_t at = 0;
    for (Group& g : groups)
      for (uint32_t i : g.idx) {
        w[i].a.n_flagged = d_nflag + at;
        const bool duo = w[i].a.q_tile_rows == 4 * kBM;
        rc = ensure_maps(ctx, w[i].qq, true, false);
        if (!rc) rc = duo ? ensure_maps(ctx, w[i].tt, true, false, true) : ensure_maps(ctx, w[i].tt, false, true);
        if (rc) return rc;
        SmaxJob& j = n_tc > 1 ? hj[at] : single;
        ++at;
        j.map_q = w[i].qq->map_a;
        j.map_t = duo ? w[i].tt->map_a : w[i].tt->map_b;
        j.map_text = duo ? w[i].tt->map_exth : w[i].tt->map_ext;
        j.a = w[i].a;
      }
    if (n_tc > 1) {
      rc = stage_upload(ctx, st, static_cast<size_t>(n_tc) * sizeof(SmaxJob));
      if (rc) return rc;
    }
    at = 0;
    for (int gi2 = 0; gi2 < kGroups; ++gi2) {
      Group& g = groups[gi2];
      const int gi = gi2 / kWalks, walk = gi2 % kWalks;     // gi = mask kind * 4 + fold * 2 + column masks
      if (g.idx.empty()) continue;
      SmaxBatch b;
      memset(&b, 0, sizeof b);
      if (n_tc > 1) b.jobs = static_cast<const SmaxJob*>(st->d) + at;
      else { b.jobs = nullptr; b.inline_job = single; }
      b.n_jobs = static_cast<uint32_t>(g.idx.size());
      b.n_ctas = w[g.idx[0]].a.n_ctas;
      b.total_tiles = g.total;
      b.flags = d_nflag + at;
      b.clear = d_nflag_other; b.clear_n = bank_len;
      for (uint32_t i : g.idx) if (w[i].exclude != nullptr) b.has_exclude = 1;
      b.mask_coop_max = ctx->knob[SSRLCV_B200_TUNE_MASK_COOPERATIVE] >= 0 ? static_cast<uint32_t>(ctx->knob[SSRLCV_B200_TUNE_MASK_COOPERATIVE]) : 12u;
      b.mask_walk_max = ctx->knob[SSRLCV_B200_TUNE_MASK_WALK] >= 0 ? static_cast<uint32_t>(ctx->knob[SSRLCV_B200_TUNE_MASK_WALK]) : 32u;
      b.diag = ctx->knob[SSRLCV_B200_TUNE_DIAG] > 0 ? static_cast<uint32_t>(ctx->knob[SSRLCV_B200_TUNE_DIAG]) : 0u;
      b.timing = ctx->knob[SSRLCV_B200_TUNE_TIMING_BUFFER] > 0 ? reinterpret_cast<unsigned long long*>(ctx->knob[SSRLCV_B200_TUNE_TIMING_BUFFER]) : nullptr;
      if (ctx->profiling) {
        if (ctx->prof_used == ctx->prof_events.size()) {
          cudaEvent_t e0, e1;
          CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
          ctx->prof_events.emplace_back(e0, e1);
        }
        CU(cudaEventRecord(ctx->prof_events[ctx->prof_used].first, ctx->stream));
      }
      cudaError_t e = walk == kWalkDuo      ? launch_smax_duo(ctx->stream, b, (gi & 2) != 0, (gi & 1) != 0)
                    : walk == kWalkPairSeed ? launch_smax_pair(ctx->stream, b, true, false, true)
                    : walk == kWalkPair     ? launch_smax_pair(ctx->stream, b, (gi & 2) != 0, (gi & 1) != 0, false)
                                            : launch_smax(ctx->stream, b, gi / 4, (gi & 2) != 0, (gi & 1) != 0, walk == kWalkTileLean);
      if (e != cudaSuccess) return fail(ctx, static_cast<int>(e), "smax launch: %s", cudaGetErrorString(e));
      if (ctx->profiling) { CU(cudaEventRecord(ctx->prof_events[ctx->prof_used].second, ctx->stream)); ctx->prof_used++; }
      launch_exact_list(ctx->stream, b, gi / 4, ctx->sm_count);
      ctx->launches += 2;
      at += b.n_jobs;
    }
  }
  // ---- SIMT engine pairs (cross-check engine), one launch each
  for (uint32_t i = 0; i < n; ++i) {
    PairWork& p = w[i];
    if (p.engine != SSRLCV_B200_ENGINE_SIMT || p.a.nq == 0 || p.a.nt == 0) continue;
    launch_exact(ctx->stream, p.a, p.mp.kind, ctx->sm_count);
    ctx->launches += 1;
  }
  CU(cudaGetLastError());
  return 0;
}

// enqueue one seed pass; `optimistic`: do not wait for the seed set's pack (see engine_enqueue)
static int seed_pass_enqueue(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* query, const ssrlcv_b200_set* seed, float* out_dev, bool* optimistic) {
  PairWork w;
  memset(&w.mp, 0, sizeof w.mp);
  w.q = query; w.t = seed; w.abs_thr = FLT_MAX;
  w.seed_only = true;
  int rc = engine_enqueue(ctx, &w, 1, optimistic);
  if (rc) return rc;
  if (w.a.nq) { launch_seed_from_keys(ctx->stream, w.a.keys, w.a.nq, w.a.nq_pad, w.a.n_splits, out_dev); ctx->launches += 1; }
  CU(cudaGetLastError());
  return 0;
}

// Enqueue-only.  If the seed set's pack has not reported back yet the fold variant is launched on the (almost always right)
// guess that its norms fit the fold digits, and the pass is remembered: the next synchronisation point of the context
// (ssrlcv_b200_match*, ssrlcv_b200_synchronize, or the release of either set) checks the verdict and redoes the pass with
// the other variant if the guess was wrong -- ssrlcv_b200_match* then also redo the matches that consumed the distances.
int ssrlcv_b200_seed_distances(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* query, const ssrlcv_b200_set* seed, float* out_dev) {
  if (!ctx || !query || !seed || (query->n && !out_dev)) return fail(ctx, SSRLCV_B200_EINVAL, "seed_distances: bad arguments");
  CU(cudaSetDevice(ctx->device));
  bool optimistic = false;
  int rc = seed_pass_enqueue(ctx, query, seed, out_dev, &optimistic);
  if (rc) return rc;
  if (optimistic) ctx->pending.push_back(ssrlcv_b200_ctx::PendingSeed{query, seed, out_dev});
  return 0;
}

}  // extern "C"

// Called right after a synchronisation of the context's stream: every pack has reported, so every optimistic seed pass can
// be checked.  *redone = true: some pass was enqueued again (the caller synchronises again; whoever consumed the
// distances in between must be redone as well).
static int verify_pending(ssrlcv_b200_ctx* ctx, bool* redone) {
  if (ctx->pending.empty()) return 0;
  std::vector<ssrlcv_b200_ctx::PendingSeed> todo;
  todo.swap(ctx->pending);
  for (const auto& p : todo) {
    int rc = set_resolve(ctx, p.seed);
    if (rc) return rc;
    if (p.seed->fold_ok || p.q->n == 0 || p.seed->n == 0) continue;
    rc = seed_pass_enqueue(ctx, p.q, p.seed, p.out, nullptr);
    if (rc) return rc;
    *redone = true;
  }
  return 0;
}

extern "C" {

static size_t out_elt(int kind) {
  switch (kind) {
    case SSRLCV_B200_OUT_MATCH: return sizeof(RefMatch);
    case SSRLCV_B200_OUT_DMATCH: return sizeof(RefDMatch);
    case SSRLCV_B200_OUT_PAIR: return sizeof(RefPair);
    case SSRLCV_B200_OUT_FEATUREMATCH: return sizeof(RefFeatureMatch);
  }
  return 0;
}

// One pair's thresholds / writers / compaction request.
struct FinalizeWork {
  const ssrlcv_b200_match_params* p;
  const ssrlcv_b200_set* q;
  const uint64_t* keys; uint32_t n_splits; uint32_t key_stride;   // keys: [n_splits][key_stride]
  const float2* t_loc; const ssrlcv_b200_set* t;                  // t may be NULL (sharded targets: locations only)
  const float* seed;
  void* out_dev;
};

// enqueue thresholds / writers / compaction of n pairs in ONE launch; the number of records of pair i lands in totals_host[i]
// (pinned, device-mapped host memory, written by the kernel itself: valid after the stream has been synchronised)
static int finalize_enqueue(ssrlcv_b200_ctx* ctx, const FinalizeWork* fw, uint32_t n, uint32_t* totals_host) {
  if (n == 0) return 0;
  size_t blocks_total = 0;
  uint32_t max_nq = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t nq = static_cast<uint32_t>(fw[i].q->n);
    blocks_total += (nq + 255) / 256;
    if (nq > max_nq) max_nq = nq;
    totals_host[i] = 0;                                    // pairs without queries launch no block
  }
  int rc = ws_reserve_zero(ctx, ctx->fstate, (blocks_total + 1) * sizeof(uint64_t));
  if (!rc) rc = ws_reserve_zero(ctx, ctx->tickets, static_cast<size_t>(n) * sizeof(uint32_t));
  if (rc) return rc;
  if (++ctx->epoch == 0) ctx->epoch = 1;
  // the job table of a single pair rides in the kernel's parameter block; several pairs are staged (one small copy)
  ssrlcv_b200_ctx::Staged* st = nullptr;
  FinalizeBatch fb;
  memset(&fb, 0, sizeof fb);
  FinalizeArgs* hf = &fb.inline_job;
  if (n > 1) {
    rc = stage_acquire(ctx, static_cast<size_t>(n) * sizeof(FinalizeArgs), &st);
    if (rc) return rc;
    hf = static_cast<FinalizeArgs*>(st->h);
    fb.jobs = static_cast<const FinalizeArgs*>(st->d);
  }
  size_t blocks_at = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const FinalizeWork& x = fw[i];
    FinalizeArgs& f = hf[i];
    memset(&f, 0, sizeof f);
    f.nq = static_cast<uint32_t>(x.q->n); f.nq_pad = x.key_stride; f.n_splits = x.n_splits;
    f.keys = x.keys; f.seed = x.seed;
    f.abs_thr = x.p->absolute_threshold;
    f.ratio = ratio_threshold(x.p->out_kind, x.p->mask_kind, x.p->relative_threshold);
    f.out_kind = x.p->out_kind; f.compact = x.p->compact;
    f.qid = x.p->query_image_id; f.tid = x.p->target_image_id;
    f.q_loc = x.q->loc; f.t_loc = x.t_loc;
    f.q_sigtheta = x.q->sigtheta; f.t_sigtheta = x.t ? x.t->sigtheta : nullptr;
    f.q_desc = x.q->desc; f.t_desc = x.t ? x.t->desc : nullptr;
    f.out = x.out_dev;
    f.state = static_cast<uint64_t*>(ctx->fstate.p) + blocks_at;
    f.ticket = static_cast<uint32_t*>(ctx->tickets.p) + i;
    f.epoch = ctx->epoch;
    f.total = nullptr;
    f.total_host = totals_host + i;
    blocks_at += (f.nq + 255) / 256;
  }
  if (n > 1) {
    rc = stage_upload(ctx, st, static_cast<size_t>(n) * sizeof(FinalizeArgs));
    if (rc) return rc;
  }
  if (max_nq) {
    for (uint32_t at = 0; at < n; at += 32768) {           // gridDim.y limit
      const uint32_t m = n - at < 32768 ? n - at : 32768;
      FinalizeBatch part = fb;
      if (n > 1) part.jobs = fb.jobs + at;
      launch_finalize(ctx->stream, part, m, max_nq);
      ctx->launches += 1;
    }
  }
  CU(cudaGetLastError());
  return 0;
}

static int check_params(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p) {
  if (!p) return fail(ctx, SSRLCV_B200_EINVAL, "match: params is NULL");
  if (p->out_kind < 0 || p->out_kind > 3 || p->mask_kind < 0 || p->mask_kind > 2) return fail(ctx, SSRLCV_B200_EINVAL, "match: unknown out_kind/mask_kind");
  if (p->out_kind == SSRLCV_B200_OUT_FEATUREMATCH && p->mask_kind == SSRLCV_B200_MASK_DOUBLE)
    return fail(ctx, SSRLCV_B200_EINVAL, "the reference has no DoubleConstrained FeatureMatch kernel");
  return 0;
}

// The pair scheduler's device side (replaces the pair loop of generateMatchesExhaustive, ref: src/MatchFactory.cu:907-936):
// all pairs of the batch go through ONE persistent tensor-core launch (per mask kind / fold variant), one tie-fallback
// launch and two finalize launches; the per-pair record counts are read back with ONE synchronisation.
int ssrlcv_b200_match_batch(ssrlcv_b200_ctx* ctx, uint32_t n_pairs, const ssrlcv_b200_match_params* params,
                            const ssrlcv_b200_set* const* queries, const ssrlcv_b200_set* const* targets,
                            const float* const* seed_dist_dev, void* const* out_dev, uint64_t* n_out) {
  if (!ctx || (n_pairs && (!params || !queries || !targets || !out_dev || !n_out))) return fail(ctx, SSRLCV_B200_EINVAL, "match_batch: bad arguments");
  if (n_pairs == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  int rc = 0;
  for (uint32_t i = 0; i < n_pairs; ++i) {
    rc = check_params(ctx, &params[i]);
    if (rc) return rc;
    if (!queries[i] || !targets[i] || (queries[i]->n && !out_dev[i])) return fail(ctx, SSRLCV_B200_EINVAL, "match_batch: pair %u: bad arguments", i);
  }
  if (ctx->hostpin.cap < n_pairs * sizeof(uint32_t)) {
    if (ctx->hostpin.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFreeHost(ctx->hostpin.p)); ctx->hostpin.p = nullptr; ctx->hostpin.cap = 0; }
    CU(cudaHostAlloc(&ctx->hostpin.p, n_pairs * sizeof(uint32_t) * 2, cudaHostAllocMapped | cudaHostAllocPortable));
    ctx->hostpin.cap = n_pairs * sizeof(uint32_t) * 2;
  }
  uint32_t* totals = static_cast<uint32_t*>(ctx->hostpin.p);
  std::vector<PairWork> w(n_pairs);
  std::vector<FinalizeWork> fw(n_pairs);
  for (int attempt = 0; attempt < 3; ++attempt) {
    for (uint32_t i = 0; i < n_pairs; ++i) {
      fill_mask(&params[i], w[i].mp);
      w[i].q = queries[i]; w[i].t = targets[i]; w[i].abs_thr = params[i].absolute_threshold;
      // invalid entries are compacted away: candidates the seed-ratio test would reject need not be tracked (matcher.cuh)
      const float* sd = seed_dist_dev ? seed_dist_dev[i] : nullptr;
      w[i].seed = (sd && params[i].compact) ? sd : nullptr;
      w[i].seed_ratio = ratio_threshold(params[i].out_kind, params[i].mask_kind, params[i].relative_threshold);
    }
    bool optimistic = false;
    rc = engine_enqueue(ctx, w.data(), n_pairs, &optimistic);
    if (rc) return rc;
    for (uint32_t i = 0; i < n_pairs; ++i)
      fw[i] = FinalizeWork{&params[i], queries[i], w[i].a.keys, w[i].a.n_splits, w[i].a.nq_pad, targets[i]->loc, targets[i],
                           seed_dist_dev ? seed_dist_dev[i] : nullptr, out_dev[i]};
    rc = finalize_enqueue(ctx, fw.data(), n_pairs, totals);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));               // the record counts were written to pinned memory by the kernel
    CU(cudaGetLastError());
    // every pack has reported by now: seed passes launched on a guess (they may have fed this batch), then this batch's own
    bool redo = false;
    rc = verify_pending(ctx, &redo);
    if (rc) return rc;
    if (optimistic) {
      for (uint32_t i = 0; i < n_pairs; ++i) {
        rc = set_resolve(ctx, targets[i]);
        if (rc) return rc;
        if (w[i].optimistic_fold && queries[i]->n && targets[i]->n && !targets[i]->fold_ok) redo = true;
      }
    }
    if (!redo) break;                                    // otherwise: once more, every set resolved, right variants
  }
  for (uint32_t i = 0; i < n_pairs; ++i) n_out[i] = totals[i];
  return 0;
}

int ssrlcv_b200_match(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query, const ssrlcv_b200_set* target,
                      const float* seed_dist_dev, uint64_t n_seed_dist, void* out_dev, uint64_t* n_out) {
  if (!ctx || !query || !target || !n_out || (query->n && !out_dev)) return fail(ctx, SSRLCV_B200_EINVAL, "match: bad arguments");
  if (seed_dist_dev && n_seed_dist != query->n)
    return fail(ctx, SSRLCV_B200_ESEEDSIZE, "seedDistances should have come from matching a seed image to queryFeatures");
  const float* seeds[1] = {seed_dist_dev};
  void* outs[1] = {out_dev};
  return ssrlcv_b200_match_batch(ctx, 1, p, &query, &target, seed_dist_dev ? seeds : nullptr, outs, n_out);
}

// The whole 2-view call sequence on device-resident AoS arrays in ONE entry: pack (one launch for all sets), seed distances
// (if a seed set is given), match, release of the packed sets.  What a caller of the reference does with
// getSeedDistances + generate* on Unity arrays in state gpu (ref: src/Pipeline.cu:176-197).
int ssrlcv_b200_match_features(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const void* query_dev, uint64_t n_query,
                               const void* target_dev, uint64_t n_target, const void* seed_dev, uint64_t n_seed, float* seed_dist_out_dev,
                               void* out_dev, uint64_t* n_out) {
  if (!ctx || !n_out) return fail(ctx, SSRLCV_B200_EINVAL, "match_features: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  const bool seeded = seed_dev != nullptr;
  const void* ptrs[3] = {query_dev, target_dev, seed_dev};
  const uint64_t sizes[3] = {n_query, n_target, n_seed};
  ssrlcv_b200_set* sets[3] = {nullptr, nullptr, nullptr};
  rc = ssrlcv_b200_pack_batch(ctx, seeded ? 3 : 2, ptrs, sizes, sets);
  if (rc) return rc;
  float* sd = nullptr;
  if (seeded) {
    sd = seed_dist_out_dev;
    if (!sd) {                                            // the caller does not want them back: context scratch
      rc = ws_reserve(ctx, ctx->seedws, (n_query ? n_query : 1) * sizeof(float));
      sd = static_cast<float*>(ctx->seedws.p);
    }
    if (!rc) rc = ssrlcv_b200_seed_distances(ctx, sets[0], sets[2], sd);
  }
  if (!rc) rc = ssrlcv_b200_match(ctx, p, sets[0], sets[1], sd, sd ? n_query : 0, out_dev, n_out);
  for (ssrlcv_b200_set* s : sets) ssrlcv_b200_set_free(s);
  return rc;
}

int ssrlcv_b200_match_candidates(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query,
                                 const ssrlcv_b200_set* shard, uint64_t target_offset, uint64_t* keys_dev) {
  return ssrlcv_b200_match_candidates_seeded(ctx, p, query, shard, target_offset, nullptr, keys_dev);
}

int ssrlcv_b200_match_candidates_seeded(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query,
                                        const ssrlcv_b200_set* shard, uint64_t target_offset, const float* seed_dist_dev,
                                        uint64_t* keys_dev) {
  if (!ctx || !query || !shard || (query->n && !keys_dev) || (target_offset & 31)) return fail(ctx, SSRLCV_B200_EINVAL, "match_candidates: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  PairWork w;
  fill_mask(p, w.mp);
  w.q = query; w.t = shard; w.target_offset = target_offset; w.abs_thr = p->absolute_threshold;
  // the merged winner is dropped when it fails the seed-ratio test, so no shard needs candidates beyond that bound (matcher.cuh)
  w.seed = (seed_dist_dev && p->compact) ? seed_dist_dev : nullptr;
  w.seed_ratio = ratio_threshold(p->out_kind, p->mask_kind, p->relative_threshold);
  rc = engine_enqueue(ctx, &w, 1, nullptr);
  if (rc) return rc;
  if (w.a.nq) {                                            // reduce the local key slots to one key per query
    launch_min_keys(ctx->stream, w.a.keys, w.a.nq, w.a.nq_pad, w.a.n_splits, keys_dev);
    ctx->launches += 1;
  }
  CU(cudaGetLastError());
  return 0;
}

int ssrlcv_b200_match_runner_up(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query,
                                const ssrlcv_b200_set* target, const uint64_t* winner_keys_dev, uint64_t* keys_dev) {
  if (!ctx || !query || !target || (query->n && (!keys_dev || !winner_keys_dev))) return fail(ctx, SSRLCV_B200_EINVAL, "match_runner_up: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  PairWork w;
  fill_mask(p, w.mp);
  w.q = query; w.t = target; w.abs_thr = p->absolute_threshold;
  w.exclude = winner_keys_dev;
  rc = engine_enqueue(ctx, &w, 1, nullptr);
  if (rc) return rc;
  if (w.a.nq) {
    launch_min_keys(ctx->stream, w.a.keys, w.a.nq, w.a.nq_pad, w.a.n_splits, keys_dev);
    ctx->launches += 1;
  }
  CU(cudaGetLastError());
  return 0;
}

int ssrlcv_b200_finalize_candidates(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query,
                                    const uint64_t* keys_dev, uint32_t n_shards, const float* target_loc_dev,
                                    const float* seed_dist_dev, void* out_dev, uint64_t* n_out) {
  if (!ctx || !query || !n_out || (query->n && (!keys_dev || !out_dev)) || n_shards == 0) return fail(ctx, SSRLCV_B200_EINVAL, "finalize_candidates: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  if (p->out_kind == SSRLCV_B200_OUT_FEATUREMATCH) return fail(ctx, SSRLCV_B200_EINVAL, "FeatureMatch output is not available for sharded targets");
  if (p->out_kind != SSRLCV_B200_OUT_PAIR && !target_loc_dev) return fail(ctx, SSRLCV_B200_EINVAL, "target locations required");
  CU(cudaSetDevice(ctx->device));
  // keys_dev is [n_shards][nq] (dense)
  FinalizeWork fw{p, query, keys_dev, n_shards, static_cast<uint32_t>(query->n), reinterpret_cast<const float2*>(target_loc_dev), nullptr,
                  seed_dist_dev, out_dev};
  rc = finalize_enqueue(ctx, &fw, 1, ctx->h_scalars + 1);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  *n_out = ctx->h_scalars[1];
  return 0;
}

// A window of rows [row0, row0 + rows_pad) of a packed set, usable wherever a set is (its own tensor maps; rows_real of
// them are features, the rest the parent's zero padding).  row0 and rows_pad are multiples of kRowPad.  Borrowed memory:
// never passed to ssrlcv_b200_set_free.
static int make_view(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* parent, uint64_t row0, uint64_t rows_real, uint64_t rows_pad,
                     ssrlcv_b200_set* v) {
  *v = *parent;
  v->n = rows_real; v->n_pad = rows_pad;
  v->desc += row0 * kDescBytes; v->ext += row0 * kExtBytes;
  v->norm += row0; v->wcol += row0; v->loc += row0; v->sigtheta += row0; v->wmax += row0 / kChunk;
  v->resolved = false;                                   // the parent's packs may still be in flight: optimistic fold launch
  v->is_view = true; v->sorted = nullptr;
  v->has_map_a = v->has_map_bx = v->has_map_exth = false;                  // its own tensor maps, encoded on first use
  (void)ctx;
  return 0;
}

// How match_host cuts a set of n rows for the streamed upload: the first row of every chunk (multiples of kRowPad) + n.
// Two equal chunks per set of at least 32 768 rows.  The available work grows with the SQUARE of the arrived fraction while
// the transfer is linear, so the pipeline is transfer-bound early and compute-bound late whatever the cut: at 65 536 x 65 536
// (0.40 ms of PCIe, 0.37 ms of matching) uneven schedules (small first / last chunks) and 4 + 4 chunks measure the same
// 0.72 ms as 2 + 2 (profiles/r2_e2e_chunks.txt), every extra chunk costing a pack + matcher + tie-list + key-reduction launch.
// ssrlcv_b200_tune(SSRLCV_B200_TUNE_HOST_CHUNK_{QUERY,TARGET}_ROWS) forces a chunk length (0 = never cut; tests use tiny ones).
static std::vector<uint64_t> chunk_bounds(const ssrlcv_b200_ctx* ctx, uint64_t n, int which) {
  std::vector<uint64_t> b{0};
  const int64_t knob = ctx->knob[which == 1 ? SSRLCV_B200_TUNE_HOST_CHUNK_TARGET_ROWS : SSRLCV_B200_TUNE_HOST_CHUNK_QUERY_ROWS];
  auto round_rows = [](uint64_t r) { return ((r + kRowPad - 1) / kRowPad) * kRowPad; };
  if (knob > 0) {
    const uint64_t rows = round_rows(static_cast<uint64_t>(knob));
    for (uint64_t r = rows; r < n; r += rows) b.push_back(r);
  } else if (knob < 0 && n >= 32768) {
    const uint64_t r = round_rows((n + 1) / 2);
    if (r < n) b.push_back(r);
  }
  b.push_back(n);
  return b;
}

// match_host for sets large enough that the PCIe transfer matters: both sets are uploaded chunk by chunk on a copy
// stream (query and target chunks interleaved) and every chunk is packed and matched against everything that has already
// arrived while the next one is still on the bus:
//   target chunk j arrives: (queries so far) x (chunk j)            -> key slot j of those queries
//   query chunk i arrives : (chunk i) x (targets so far, one view)  -> key slot n_tchunks of those queries
// Each (query, target) combination is evaluated exactly once; the minimum over a query's slots is the reference's winner
// (keys carry the global target index, chunk boundaries are multiples of 32 so f mod 32 is unchanged), then one finalize.
static int match_host_streamed(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const void* query_host, uint64_t nq,
                               const void* target_host, uint64_t nt, const void* seed_host, uint64_t ns, void* out_host,
                               uint64_t* n_out, const std::vector<uint64_t>& qb, const std::vector<uint64_t>& tb) {
  CU(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  // Every exit after this point goes through cleanup(): on failure both streams are drained first, because copies from the
  // CALLER's host buffers may still be in flight and the caller is free to release them once we return.
  ssrlcv_b200_set *q = nullptr, *t = nullptr, *s = nullptr;
  auto cleanup = [&](int code) {
    if (code != 0) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); }
    ssrlcv_b200_set_free(q); ssrlcv_b200_set_free(t); ssrlcv_b200_set_free(s);
    return code;
  };
#define CUS(call)                                                                                                   \
  do {                                                                                                              \
    cudaError_t e_ = (call);                                                                                        \
    if (e_ != cudaSuccess) return cleanup(fail(ctx, static_cast<int>(e_), "%s: %s", #call, cudaGetErrorString(e_))); \
  } while (0)
  const uint64_t n_qc = qb.size() - 1, n_tc = tb.size() - 1;      // chunk i of a set = rows [b[i], b[i + 1])
  while (ctx->copy_events.size() < n_qc + n_tc + 2) {
    cudaEvent_t e;
    CUS(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_events.push_back(e);
  }
  const size_t rec = sizeof(RefFeature), elt = out_elt(p->out_kind);
  int rc = ws_reserve(ctx, ctx->upload, (nq + nt + ns) * rec + 16);
  if (rc) return cleanup(rc);
  uint8_t* up_q = static_cast<uint8_t*>(ctx->upload.p);
  uint8_t* up_t = up_q + nq * rec;
  uint8_t* up_s = up_t + nt * rec;
  // scratch: [out records | seed distances | chunk keys (n_tc + 1) x nq]
  const size_t out_bytes = ((nq * elt + 255) / 256) * 256, seed_bytes = ((nq * sizeof(float) + 255) / 256) * 256;
  const size_t n_slots = n_tc + 1;
  rc = ws_reserve(ctx, ctx->scratch, out_bytes + seed_bytes + n_slots * nq * sizeof(uint64_t) + 256);
  if (rc) return cleanup(rc);
  void* out_dev = ctx->scratch.p;
  float* seed_dev = reinterpret_cast<float*>(static_cast<uint8_t*>(ctx->scratch.p) + out_bytes);
  uint64_t* ckeys = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(ctx->scratch.p) + out_bytes + seed_bytes);

  // ---- the copy stream starts after everything already enqueued on the main stream (earlier readers of the upload buffer)
  cudaEvent_t ev_fence = ctx->copy_events[n_qc + n_tc + 1], ev_seed = ctx->copy_events[n_qc + n_tc];
  CUS(cudaEventRecord(ev_fence, ctx->stream));
  CUS(cudaStreamWaitEvent(ctx->copy_stream, ev_fence, 0));
  // upload order: seed set, then query / target chunks interleaved in proportion (query chunk first)
  struct Up { bool is_q; uint64_t idx; };
  std::vector<Up> order;
  {
    // With a seed set every query chunk goes first: the seed distances (needed to bound the candidates, matcher.cuh
    // q_seed) exist once the last query has arrived, before the first (queries x target chunk) job is enqueued.
    uint64_t qi = 0, ti = 0;
    while (qi < n_qc || ti < n_tc) {
      // the set whose arrived FRACTION is smaller goes next (query first on a tie): qb[qi] / nq <= tb[ti] / nt
      const bool take_q = ti >= n_tc || (qi < n_qc && (seed_host != nullptr || qb[qi] * nt <= tb[ti] * nq));
      if (take_q) order.push_back(Up{true, qi++}); else order.push_back(Up{false, ti++});
    }
  }
  if (seed_host && ns) CUS(cudaMemcpyAsync(up_s, seed_host, ns * rec, cudaMemcpyHostToDevice, ctx->copy_stream));
  if (seed_host) CUS(cudaEventRecord(ev_seed, ctx->copy_stream));
  for (size_t k = 0; k < order.size(); ++k) {
    const Up& u = order[k];
    const std::vector<uint64_t>& bb = u.is_q ? qb : tb;
    const uint64_t r0 = bb[u.idx], cnt = bb[u.idx + 1] - r0;
    const uint8_t* src = static_cast<const uint8_t*>(u.is_q ? query_host : target_host) + r0 * rec;
    CUS(cudaMemcpyAsync((u.is_q ? up_q : up_t) + r0 * rec, src, cnt * rec, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUS(cudaEventRecord(ctx->copy_events[k], ctx->copy_stream));
  }

  // ---- main stream: pack and match as the chunks arrive
  rc = set_alloc(ctx, nq, &q);
  if (!rc) rc = set_alloc(ctx, nt, &t);
  if (rc) return cleanup(rc);
  if (seed_host) {
    CUS(cudaStreamWaitEvent(ctx->stream, ev_seed, 0));
    rc = ssrlcv_b200_pack(ctx, up_s, ns, &s);
    if (rc) return cleanup(rc);
  }
  launch_fill_u64(ctx->stream, ckeys, n_slots * nq, kKeyNone);
  ctx->launches += 1;
  PairWork w;
  fill_mask(p, w.mp);
  w.abs_thr = p->absolute_threshold;
  uint64_t q_arrived = 0, t_arrived = 0;                      // rows
  bool any_optimistic = false;
  for (size_t k = 0; k < order.size(); ++k) {
    const Up& u = order[k];
    ssrlcv_b200_set* set = u.is_q ? q : t;
    const std::vector<uint64_t>& bb = u.is_q ? qb : tb;
    const uint64_t r0 = bb[u.idx];
    const bool last = u.idx + 2 == bb.size();
    const uint64_t cnt = bb[u.idx + 1] - r0, pad = last ? set->n_pad - r0 : cnt;
    CUS(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[k], 0));
    const PackJob pj = pack_job(set, (u.is_q ? up_q : up_t) + r0 * rec, cnt, pad, r0);
    launch_pack(ctx->stream, &pj, 1, ctx->sm_count, &ctx->launches);
    ssrlcv_b200_set qv, tv;
    uint64_t q0 = 0, slot = 0;
    bool run = false;
    if (u.is_q) {
      if (last) { rc = set_finish(ctx, q); if (rc) return cleanup(rc); }
      if (t_arrived) {                                         // (chunk i) x (targets so far)
        rc = make_view(ctx, q, r0, cnt, pad, &qv);
        if (!rc) rc = make_view(ctx, t, 0, t_arrived, t_arrived == t->n ? t->n_pad : t_arrived, &tv);
        if (rc) return cleanup(rc);
        w.target_offset = 0; q0 = r0; slot = n_tc; run = cnt != 0;
        w.seed = nullptr;                                        // (a query chunk's seed distances do not exist yet)
      }
      q_arrived = r0 + cnt;
    } else {
      if (last) { rc = set_finish(ctx, t); if (rc) return cleanup(rc); }
      if (q_arrived) {                                         // (queries so far) x (chunk j)
        const uint64_t q_pad = q_arrived == q->n ? q->n_pad : q_arrived;
        rc = make_view(ctx, q, 0, q_arrived, q_pad, &qv);
        if (!rc) rc = make_view(ctx, t, r0, cnt, pad, &tv);
        if (rc) return cleanup(rc);
        w.target_offset = r0; q0 = 0; slot = u.idx; run = cnt != 0;
        // the whole query set has arrived and its seed distances are enqueued (see the upload order above)
        w.seed = (s && p->compact && q_arrived == q->n) ? seed_dev : nullptr;
        w.seed_ratio = ratio_threshold(p->out_kind, p->mask_kind, p->relative_threshold);
      }
      t_arrived = r0 + cnt;                                    // a multiple of kRowPad until the last chunk
    }
    if (run) {
      w.q = &qv; w.t = &tv;
      bool optimistic = false;
      rc = engine_enqueue(ctx, &w, 1, &optimistic);
      if (rc) return cleanup(rc);
      any_optimistic |= optimistic;
      launch_min_keys(ctx->stream, w.a.keys, w.a.nq, w.a.nq_pad, w.a.n_splits, ckeys + slot * nq + q0);
      ctx->launches += 1;
    }
    if (u.is_q && last && s) {                                 // every query is on the device: seed distances
      rc = ssrlcv_b200_seed_distances(ctx, q, s, seed_dev);
      if (rc) return cleanup(rc);
    }
  }
  FinalizeWork fw{p, q, ckeys, static_cast<uint32_t>(n_slots), static_cast<uint32_t>(nq), t->loc, t, s ? seed_dev : nullptr, out_dev};
  rc = finalize_enqueue(ctx, &fw, 1, ctx->h_scalars + 2);
  if (rc) return cleanup(rc);
  CUS(cudaStreamSynchronize(ctx->stream));
  *n_out = ctx->h_scalars[2];
  bool seeds_redone = false;
  rc = verify_pending(ctx, &seeds_redone);
  if (!rc) rc = set_resolve(ctx, t);
  if (rc) return cleanup(rc);
  if (seeds_redone || (any_optimistic && !t->fold_ok)) {
    // some norm does not fit the fold digits: the optimistic launches skipped their jobs (or ran on seed distances that
    // have only now been computed).  Both sets are packed by now; redo the pair in one piece.
    rc = ssrlcv_b200_match(ctx, p, q, t, s ? seed_dev : nullptr, s ? nq : 0, out_dev, n_out);
    if (rc) return cleanup(rc);
  }
  if (*n_out) {
    cudaError_t e = cudaMemcpyAsync(out_host, out_dev, *n_out * elt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cleanup(fail(ctx, static_cast<int>(e), "match_host: D2H: %s", cudaGetErrorString(e)));
  }
  return cleanup(0);
}
#undef CUS

int ssrlcv_b200_match_host(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const void* query_host, uint64_t n_query,
                           const void* target_host, uint64_t n_target, const void* seed_host, uint64_t n_seed, void* out_host,
                           uint64_t* n_out) {
  if (!ctx || !n_out) return fail(ctx, SSRLCV_B200_EINVAL, "match_host: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  ssrlcv_b200_set *q = nullptr, *t = nullptr, *s = nullptr;
  float* seed_dev = nullptr;
  void* out_dev = nullptr;
  const size_t elt = out_elt(p->out_kind);
  if ((n_query && !query_host) || (n_target && !target_host) || (seed_host == nullptr && n_seed)) return fail(ctx, SSRLCV_B200_EINVAL, "match_host: bad arguments");
  if (n_query > 0x7fffffffull || n_target > 0x7fffffffull || n_seed > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "match_host: too many features");
  {
    const std::vector<uint64_t> qb = chunk_bounds(ctx, n_query, 0), tb = chunk_bounds(ctx, n_target, 1);
    // also for smaller seeded calls: the target set travels on the copy stream while the seed pass (queries x seed set) runs
    const bool overlap_seed_pass = seed_host != nullptr && n_seed != 0 && n_target >= 8192 &&
                                   ctx->knob[SSRLCV_B200_TUNE_HOST_CHUNK_QUERY_ROWS] < 0 && ctx->knob[SSRLCV_B200_TUNE_HOST_CHUNK_TARGET_ROWS] < 0;
    if (n_query && n_target && engine_of(ctx) != SSRLCV_B200_ENGINE_SIMT && (qb.size() > 2 || tb.size() > 2 || overlap_seed_pass))
      return match_host_streamed(ctx, p, query_host, n_query, target_host, n_target, seed_host, n_seed, out_host, n_out, qb, tb);
  }
  rc = pack_host_async(ctx, query_host, n_query, &q);
  if (!rc) rc = pack_host_async(ctx, target_host, n_target, &t);
  if (!rc && seed_host) rc = pack_host_async(ctx, seed_host, n_seed, &s);
  if (!rc) {
    // scratch layout after the packs: [out records | seed distances]
    const size_t out_bytes = ((n_query * elt + 255) / 256) * 256;
    rc = ws_reserve(ctx, ctx->scratch, out_bytes + n_query * sizeof(float) + 256);
    if (!rc) {
      out_dev = ctx->scratch.p;
      seed_dev = reinterpret_cast<float*>(static_cast<uint8_t*>(ctx->scratch.p) + out_bytes);
    }
  }
  if (!rc && s) rc = ssrlcv_b200_seed_distances(ctx, q, s, seed_dev);
  if (!rc) rc = ssrlcv_b200_match(ctx, p, q, t, s ? seed_dev : nullptr, s ? n_query : 0, out_dev, n_out);
  if (!rc && *n_out) {
    cudaError_t e = cudaMemcpyAsync(out_host, out_dev, *n_out * elt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, static_cast<int>(e), "match_host: D2H: %s", cudaGetErrorString(e));
  }
  if (rc) cudaStreamSynchronize(ctx->stream);              // copies from the caller's host buffers may still be in flight
  ssrlcv_b200_set_free(q); ssrlcv_b200_set_free(t); ssrlcv_b200_set_free(s);
  return rc;
}

// ------------------------------------------------------------------------------------------- post-processing
static int compact_inplace(ssrlcv_b200_ctx* ctx, void* dev, uint64_t n, uint32_t elt, int pred, float thr, uint64_t* n_out) {
  if (n > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "too many records");
  CU(cudaSetDevice(ctx->device));
  const uint32_t blocks = static_cast<uint32_t>((n + 255) / 256);
  int rc = ws_reserve(ctx, ctx->counts, (blocks + 1) * sizeof(uint32_t));
  if (!rc) rc = ws_reserve(ctx, ctx->scratch, n * elt + 16);
  if (rc) return rc;
  CompactArgs c;
  c.in = static_cast<const uint8_t*>(dev); c.out = static_cast<uint8_t*>(ctx->scratch.p);
  c.n = static_cast<uint32_t>(n); c.elt = elt; c.pred = pred; c.threshold = thr;
  c.block_counts = static_cast<uint32_t*>(ctx->counts.p); c.total = ctx->d_scalars + 1;
  launch_compact(ctx->stream, c);
  if (n) ctx->launches += 2;
  CU(cudaMemcpyAsync(ctx->h_scalars + 1, c.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const uint64_t kept = ctx->h_scalars[1];
  if (kept) CU(cudaMemcpyAsync(dev, ctx->scratch.p, kept * elt, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (n_out) *n_out = kept;
  return 0;
}

int ssrlcv_b200_validate(ssrlcv_b200_ctx* ctx, int out_kind, void* matches_dev, uint64_t n, uint64_t* n_out) {
  if (!ctx || (n && !matches_dev) || out_kind < 0 || out_kind > 3) return fail(ctx, SSRLCV_B200_EINVAL, "validate: bad arguments");
  return compact_inplace(ctx, matches_dev, n, static_cast<uint32_t>(out_elt(out_kind)), out_kind == SSRLCV_B200_OUT_PAIR ? 1 : 0, 0.f, n_out);
}

int ssrlcv_b200_refine(ssrlcv_b200_ctx* ctx, void* dmatches_dev, uint64_t n, float threshold, uint64_t* n_out) {
  if (!ctx || (n && !dmatches_dev)) return fail(ctx, SSRLCV_B200_EINVAL, "refine: bad arguments");
  if (threshold == 0.0f) return fail(ctx, SSRLCV_B200_ETHRESHOLD, "ERROR illegal value used for threshold: 0.0");
  return compact_inplace(ctx, dmatches_dev, n, sizeof(RefDMatch), 2, threshold, n_out);
}

int ssrlcv_b200_sort(ssrlcv_b200_ctx* ctx, void* dmatches_dev, uint64_t n) {
  if (!ctx || (n && !dmatches_dev) || n > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "sort: bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (n < 2) return 0;
  const size_t kb = ((n * sizeof(uint64_t) + 255) / 256) * 256;
  const size_t temp = sort_temp_bytes(static_cast<uint32_t>(n));
  int rc = ws_reserve(ctx, ctx->sortws, 2 * kb + temp + 256);
  if (!rc) rc = ws_reserve(ctx, ctx->scratch, n * sizeof(RefDMatch));
  if (rc) return rc;
  uint8_t* w = static_cast<uint8_t*>(ctx->sortws.p);
  cudaError_t e = launch_sort(ctx->stream, dmatches_dev, static_cast<uint32_t>(n), reinterpret_cast<uint64_t*>(w), reinterpret_cast<uint64_t*>(w + kb),
                              w + 2 * kb, temp, ctx->scratch.p, &ctx->launches);
  if (e != cudaSuccess) return fail(ctx, static_cast<int>(e), "sort: %s", cudaGetErrorString(e));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return 0;
}

int ssrlcv_b200_raw_matches(ssrlcv_b200_ctx* ctx, const void* dmatches_dev, uint64_t n, void* matches_dev) {
  if (!ctx || (n && (!dmatches_dev || !matches_dev)) || n > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "raw_matches: bad arguments");
  CU(cudaSetDevice(ctx->device));
  launch_raw_matches(ctx->stream, dmatches_dev, static_cast<uint32_t>(n), matches_dev);
  if (n) ctx->launches += 1;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return 0;
}

int ssrlcv_b200_match_set_2view(ssrlcv_b200_ctx* ctx, const void* dmatches_dev, uint64_t n, void* keypoints_dev, void* multimatches_dev,
                                float* max_distance) {
  if (!ctx || (n && (!dmatches_dev || !keypoints_dev || !multimatches_dev)) || n > 0x3fffffffull)
    return fail(ctx, SSRLCV_B200_EINVAL, "match_set_2view: bad arguments");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemsetAsync(ctx->d_scalars + 3, 0, sizeof(uint32_t), ctx->stream));
  launch_match_set_2view(ctx->stream, dmatches_dev, static_cast<uint32_t>(n), keypoints_dev, multimatches_dev, ctx->d_scalars + 3);
  if (n) ctx->launches += 1;
  CU(cudaMemcpyAsync(ctx->h_scalars + 3, ctx->d_scalars + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  if (max_distance) memcpy(max_distance, ctx->h_scalars + 3, sizeof(float));
  return 0;
}

// ------------------------------------------------------------------------------------------- diagnostics
// one array of a packed set to the host (parity tests: the producer hand-off must give byte-identical sets)
int ssrlcv_b200_set_download(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* s, int which, void* host_out, uint64_t capacity, uint64_t* bytes) {
  if (!ctx || !s || !bytes) return fail(ctx, SSRLCV_B200_EINVAL, "set_download: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const void* src = nullptr;
  uint64_t n = 0;
  switch (which) {
    case 0: src = s->desc; n = s->n_pad * kDescBytes; break;
    case 1: src = s->ext; n = s->n_pad * kExtBytes; break;
    case 2: src = s->norm; n = s->n_pad * 4; break;
    case 3: src = s->wcol; n = s->n_pad * 4; break;
    case 4: src = s->wmax; n = (s->n_pad / kChunk) * 4; break;
    case 5: src = s->loc; n = s->n_pad * 8; break;
    case 6: src = s->sigtheta; n = s->n_pad * 8; break;
    case 7: src = s->meta; n = 4; break;
    default: return fail(ctx, SSRLCV_B200_EINVAL, "set_download: unknown array");
  }
  *bytes = n;
  if (!host_out) return 0;
  if (capacity < n) return fail(ctx, SSRLCV_B200_EINVAL, "set_download: buffer too small");
  CU(cudaMemcpyAsync(host_out, src, n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int ssrlcv_b200_epipolar_endpoints(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query, float* out_dev) {
  if (!ctx || !p || !query || (query->n && !out_dev)) return fail(ctx, SSRLCV_B200_EINVAL, "epipolar_endpoints: bad arguments");
  CU(cudaSetDevice(ctx->device));
  ssrlcv_b200_match_params pp = *p;
  pp.mask_kind = SSRLCV_B200_MASK_DOUBLE;
  MaskParams mp; fill_mask(&pp, mp);
  launch_endpoints(ctx->stream, mp, query->loc, query->n, reinterpret_cast<float4*>(out_dev));
  if (query->n) ctx->launches += 1;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return 0;
}

int ssrlcv_b200_device_sincos(ssrlcv_b200_ctx* ctx, const float* x_dev, uint64_t n, float* sin_dev, float* cos_dev) {
  if (!ctx || (n && (!x_dev || !sin_dev || !cos_dev))) return fail(ctx, SSRLCV_B200_EINVAL, "device_sincos: bad arguments");
  CU(cudaSetDevice(ctx->device));
  launch_sincos(ctx->stream, x_dev, n, sin_dev, cos_dev);
  if (n) ctx->launches += 1;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return 0;
}

}  // extern "C"
