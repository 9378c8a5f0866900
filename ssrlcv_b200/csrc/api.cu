// api.cu -- the C ABI (include/ssrlcv_b200.h): context, packed sets, and the orchestration of
// pack -> tensor-core matcher -> exact tie resolution -> finalize.  Host code only launches our own
// kernels; there is no CPU fallback: without an sm_100 device every entry point fails.
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "matcher.cuh"

namespace b200 {
// pack.cu
void launch_pack(cudaStream_t, const void*, uint64_t, ssrlcv_b200_set_impl&, int32_t*, int);
// smax_tc.cu
cudaError_t launch_smax(cudaStream_t, int, const ssrlcv_b200_set_impl&, const ssrlcv_b200_set_impl&, const SmaxArgs&, const MaskParams&);
// exact.cu
void launch_exact(cudaStream_t, const SmaxArgs&, const MaskParams&, int, int);
void launch_endpoints(cudaStream_t, const MaskParams&, const float2*, uint64_t, float4*);
void launch_query_masks(cudaStream_t, const MaskParams&, const float2*, uint64_t, float4*);
void launch_sincos(cudaStream_t, const float*, uint64_t, float*, float*);
// finalize.cu
struct FinalizeArgs {
  uint32_t nq, nq_pad, n_splits;
  const uint64_t* keys;
  const float* seed;
  float abs_thr, ratio;
  int32_t out_kind, compact;
  uint32_t qid, tid;
  const float2* q_loc; const float2* t_loc;
  const float2* q_sigtheta; const float2* t_sigtheta;
  const uint8_t* q_desc; const uint8_t* t_desc;
  void* out;
  uint32_t* block_counts;
  uint32_t* total;
};
struct CompactArgs {
  const uint8_t* in; uint8_t* out; uint32_t n; uint32_t elt; int32_t pred; float threshold;
  uint32_t* block_counts; uint32_t* total;
};
void launch_finalize(cudaStream_t, FinalizeArgs);
void launch_seed_from_keys(cudaStream_t, const uint64_t*, uint32_t, uint32_t, uint32_t, float*);
void launch_fill_u64(cudaStream_t, uint64_t*, uint64_t, uint64_t);
void launch_min_keys(cudaStream_t, const uint64_t*, uint32_t, uint32_t, uint32_t, uint64_t*);
void launch_compact(cudaStream_t, CompactArgs);
void launch_raw_matches(cudaStream_t, const void*, uint32_t, void*);
uint64_t launch_sort(cudaStream_t, void*, uint32_t, uint64_t*, void*);
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
  Workspace keys, flagged, counts, scratch, hostpin, qmask, totals;
  uint32_t* d_scalars = nullptr;   // [0] n_flagged, [1] total, [2..3] hmin/hmax staging
  uint32_t* h_scalars = nullptr;   // pinned mirror
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  // pack() does not synchronise: the (hmin, hmax) of a set lands in a pinned slot and is read when a match first
  // needs it (set_resolve).  Slots and their events are recycled through this free list.
  struct PackSlot { int32_t* mm; cudaEvent_t ready; };
  std::vector<PackSlot> free_slots;
  std::vector<int32_t*> slot_blocks;
};

static int slot_acquire(ssrlcv_b200_ctx* ctx, ssrlcv_b200_ctx::PackSlot& out) {
  if (ctx->free_slots.empty()) {
    constexpr int kPerBlock = 256;
    int32_t* block = nullptr;
    cudaError_t e = cudaMallocHost(&block, kPerBlock * 2 * sizeof(int32_t));
    if (e != cudaSuccess) return static_cast<int>(e);
    ctx->slot_blocks.push_back(block);
    for (int i = 0; i < kPerBlock; ++i) {
      ssrlcv_b200_ctx::PackSlot s{block + 2 * i, nullptr};
      e = cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming);
      if (e != cudaSuccess) return static_cast<int>(e);
      ctx->free_slots.push_back(s);
    }
  }
  out = ctx->free_slots.back();
  ctx->free_slots.pop_back();
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
  if (cudaMalloc(&ctx->d_scalars, 16 * sizeof(uint32_t)) != cudaSuccess || cudaMallocHost(&ctx->h_scalars, 16 * sizeof(uint32_t)) != cudaSuccess) {
    delete ctx;
    return SSRLCV_B200_ENOMEM;
  }
  ctx->h_scalars[2] = 0x7fffffff; ctx->h_scalars[3] = 0xffffffffu;   // constant: initial (hmin, hmax) = (INT_MAX, -1)
  *out = ctx;
  return 0;
}

void ssrlcv_b200_destroy(ssrlcv_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (Workspace* w : {&ctx->keys, &ctx->flagged, &ctx->counts, &ctx->scratch, &ctx->qmask, &ctx->totals}) if (w->p) cudaFree(w->p);
  if (ctx->hostpin.p) cudaFreeHost(ctx->hostpin.p);
  for (auto& sl : ctx->free_slots) cudaEventDestroy(sl.ready);
  for (int32_t* b : ctx->slot_blocks) cudaFreeHost(b);
  if (ctx->d_scalars) cudaFree(ctx->d_scalars);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  for (auto& pe : ctx->prof_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* ssrlcv_b200_last_error(const ssrlcv_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
int ssrlcv_b200_synchronize(ssrlcv_b200_ctx* ctx) {
  if (!ctx) return SSRLCV_B200_EINVAL;
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}
uint64_t ssrlcv_b200_launch_count(const ssrlcv_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

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

// Enqueue-only: nothing here waits for the device.  The (hmin, hmax) range of the set's norms, which selects the kernel
// variant and offsets the scores, is copied to a pinned slot and read by set_resolve() when a match first needs it.
int ssrlcv_b200_pack(ssrlcv_b200_ctx* ctx, const void* features_dev, uint64_t n, ssrlcv_b200_set** out) {
  if (!ctx || !out || (n && !features_dev) || n > 0x7fffffffull) return fail(ctx, SSRLCV_B200_EINVAL, "pack: bad arguments");
  CU(cudaSetDevice(ctx->device));
  ssrlcv_b200_set* s = new (std::nothrow) ssrlcv_b200_set();
  if (!s) return SSRLCV_B200_ENOMEM;
  memset(static_cast<ssrlcv_b200_set_impl*>(s), 0, sizeof(ssrlcv_b200_set_impl));
  s->ctx = ctx;
  s->n = n;
  s->n_pad = ((n + kRowPad - 1) / kRowPad) * kRowPad;
  if (s->n_pad == 0) s->n_pad = kRowPad;
  ssrlcv_b200_ctx::PackSlot slot;
  int rc = slot_acquire(ctx, slot);
  if (rc) { delete s; return fail(ctx, rc, "pack: pinned slot: %s", cudaGetErrorString(static_cast<cudaError_t>(rc))); }
  s->h_mm = slot.mm; s->ready = slot.ready; s->resolved = false;
  // one allocation: desc | ext | norm | wcol | loc | sigtheta   (all segments 256-byte aligned since n_pad % 256 == 0)
  const size_t np = s->n_pad;
  const size_t bytes = np * (kDescBytes + kExtBytes + 4 + 4 + 8 + 8);
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
  // only the padding rows need zeroing (the kernels write every real row); everything when the set is tiny
  if (np - n >= n) {
    cudaMemsetAsync(base, 0, bytes, ctx->stream);
  } else {
    const size_t pad = np - n;
    if (pad) {
      cudaMemsetAsync(s->desc + n * kDescBytes, 0, pad * kDescBytes, ctx->stream);
      cudaMemsetAsync(s->ext + n * kExtBytes, 0, pad * kExtBytes, ctx->stream);
      cudaMemsetAsync(s->norm + n, 0, pad * 4, ctx->stream);
      cudaMemsetAsync(s->wcol + n, 0, pad * 4, ctx->stream);
      cudaMemsetAsync(s->loc + n, 0, pad * 8, ctx->stream);
      cudaMemsetAsync(s->sigtheta + n, 0, pad * 8, ctx->stream);
    }
  }
  int32_t* hmm = reinterpret_cast<int32_t*>(ctx->d_scalars + 2);   // shared staging: every use is ordered by the stream
  cudaMemcpyAsync(hmm, ctx->h_scalars + 2, 8, cudaMemcpyHostToDevice, ctx->stream);
  launch_pack(ctx->stream, features_dev, n, *s, hmm, ctx->sm_count);
  if (n) ctx->launches += 2;
  cudaMemcpyAsync(s->h_mm, hmm, 8, cudaMemcpyDeviceToHost, ctx->stream);
  cudaEventRecord(s->ready, ctx->stream);
  e = cudaGetLastError();
  if (e != cudaSuccess) { ssrlcv_b200_set_free(s); return fail(ctx, static_cast<int>(e), "pack: %s", cudaGetErrorString(e)); }
  rc = encode_map(ctx, &s->map_a, s->desc, np, kDescBytes, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!rc) rc = encode_map(ctx, &s->map_b, s->desc, np, kDescBytes, kBN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!rc) rc = encode_map(ctx, &s->map_ext, s->ext, np, kExtBytes, kBN, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) { ssrlcv_b200_set_free(s); return rc; }
  *out = s;
  return 0;
}

// first use of a set's norm range on the host: wait for its pack (normally long finished)
static int set_resolve(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* cs) {
  ssrlcv_b200_set* s = const_cast<ssrlcv_b200_set*>(cs);
  if (s->resolved) return 0;
  CU(cudaEventSynchronize(s->ready));
  s->hmin = s->h_mm[0];
  s->hmax = s->h_mm[1];
  s->fold_ok = (s->n == 0) || (static_cast<int64_t>(s->hmax) - s->hmin + 2 <= kFoldMaxW);
  if (getenv("SSRLCV_B200_NOFOLD")) s->fold_ok = false;   // test hook: force the epilogue-add variant
  s->resolved = true;
  return 0;
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
  cudaSetDevice(s->ctx->device);
  if (s->desc) cudaFreeAsync(s->desc, s->ctx->stream);   // stream-ordered: every kernel that reads the set was enqueued before this
  if (s->ready) {
    // the slot's D2H was enqueued before this point; a later pack that reuses it is ordered after it on the same stream
    s->ctx->free_slots.push_back(ssrlcv_b200_ctx::PackSlot{s->h_mm, s->ready});
  }
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
}

// Tensor-core engine: cut the linearised (query tile, target tile) space into one contiguous, equally long range per
// CTA (matcher.cuh).  Every SM gets the same number of tiles (+-1) whatever the shape; a query tile is shared by the
// few CTAs whose ranges cross it, each writing its own key slot.
static void plan_tc(const ssrlcv_b200_ctx* ctx, SmaxArgs& a, int& grid) {
  const uint32_t sms = static_cast<uint32_t>(ctx->sm_count);
  a.total_tiles = static_cast<uint64_t>(a.n_qtiles) * a.n_ttiles;
  const uint64_t min_tiles = 4;                       // per-unit fixed cost (A tile, pipeline fill, resolve) ~ a few tiles
  uint64_t g = a.total_tiles / min_tiles;
  if (g < 1) g = 1;
  if (g > sms) g = sms;
  a.n_ctas = static_cast<uint32_t>(g);
  a.tiles_per_split = a.n_ttiles;
  uint32_t slots = 1;
  if (a.total_tiles) {
    const uint64_t per = a.total_tiles / a.n_ctas;    // shortest range
    slots = static_cast<uint32_t>((a.n_ttiles + per - 1) / per) + 1;
    if (slots > a.n_ctas) slots = a.n_ctas;
    if (a.n_qtiles <= 8192) {                         // exact: the widest query tile
      uint32_t widest = 1;
      for (uint32_t qt = 0; qt < a.n_qtiles; ++qt) {
        const uint64_t base = static_cast<uint64_t>(qt) * a.n_ttiles;
        const uint32_t w = cta_of_tile(a.total_tiles, a.n_ctas, base + a.n_ttiles - 1) - cta_of_tile(a.total_tiles, a.n_ctas, base) + 1;
        if (w > widest) widest = w;
      }
      slots = widest;
    }
  }
  a.n_splits = slots;
  grid = static_cast<int>(a.n_ctas);
}

// largest integer d2 with (float)d2 < abs_thr, i.e. that the reference's `d2 >= abs` test lets through
// (ref: src/MatchFactory.cu:1472, :1494); INT32_MAX when the threshold cannot bind (d2 <= 128 * 255^2)
static int32_t d2_cap_of(float abs_thr) {
  if (!(abs_thr > 0.0f)) return -1;                   // also NaN: nothing passes
  if (abs_thr > 8323200.0f) return 0x7fffffff;
  const float c = ceilf(abs_thr);
  return static_cast<int32_t>(c) - 1;
}

// shape-only part of the plan: tile counts, CTA ranges, key slots
static int make_plan(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* q, const ssrlcv_b200_set* t, int engine, SmaxArgs& a, int& grid) {
  memset(&a, 0, sizeof a);
  a.nq = static_cast<uint32_t>(q->n); a.nt = static_cast<uint32_t>(t->n); a.nq_pad = static_cast<uint32_t>(q->n_pad);
  a.n_qtiles = (a.nq + kBM - 1) / kBM;
  a.n_ttiles = (a.nt + kBN - 1) / kBN;
  grid = 1;
  if (static_cast<uint64_t>(a.n_qtiles) * a.n_ttiles >= (1ull << 32))
    return fail(ctx, SSRLCV_B200_EINVAL, "%u x %u features is too large for one call: shard the target set (ssrlcv_b200_match_candidates)", a.nq, a.nt);
  if (engine == SSRLCV_B200_ENGINE_SIMT || a.nq == 0 || a.nt == 0) plan_simt(ctx, a);
  else plan_tc(ctx, a, grid);
  return 0;
}

// runs the distance engine; leaves exact keys in ctx->keys ([n_splits][nq_pad])
static int run_engine(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* q, const ssrlcv_b200_set* t, const MaskParams& mp, int engine,
                      uint64_t target_offset, float abs_thr, SmaxArgs& a) {
  int grid = 1;
  int rc = make_plan(ctx, q, t, engine, a, grid);
  if (!rc) rc = set_resolve(ctx, t);
  if (rc) return rc;
  a.target_offset = target_offset;
  a.t_hmax = t->hmax;
  a.d2_cap = d2_cap_of(abs_thr);
  a.q_desc = q->desc; a.q_norm = q->norm; a.q_loc = q->loc;
  a.t_desc = t->desc; a.t_norm = t->norm; a.t_loc = t->loc; a.t_wcol = t->wcol;
  const size_t n_keys = static_cast<size_t>(a.n_splits) * a.nq_pad;
  rc = ws_reserve(ctx, ctx->keys, n_keys * sizeof(uint64_t));
  if (!rc) rc = ws_reserve(ctx, ctx->flagged, n_keys * sizeof(uint32_t));
  if (rc) return rc;
  a.keys = static_cast<uint64_t*>(ctx->keys.p);
  a.flagged = static_cast<uint32_t*>(ctx->flagged.p);
  a.n_flagged = ctx->d_scalars;
  a.timing = nullptr;
  if (const char* tp = getenv("SSRLCV_B200_TIMING")) {       // diagnostics: device buffer address (hex) for cycle counters
    a.timing = reinterpret_cast<unsigned long long*>(strtoull(tp, nullptr, 16));
  }
  if (a.nq == 0) return 0;
  if (mp.kind != SSRLCV_B200_MASK_NONE && a.nt != 0) {      // per-query mask state, once per call (all padded rows: loc = 0 there)
    rc = ws_reserve(ctx, ctx->qmask, static_cast<size_t>(a.nq_pad) * 2 * sizeof(float4));
    if (rc) return rc;
    launch_query_masks(ctx->stream, mp, q->loc, a.nq_pad, static_cast<float4*>(ctx->qmask.p));
    a.q_mask = static_cast<const float4*>(ctx->qmask.p);
    ctx->launches += 1;
  }
  if (a.nt == 0) { launch_fill_u64(ctx->stream, a.keys, n_keys, kKeyNone); ctx->launches += 1; return 0; }
  if (engine == SSRLCV_B200_ENGINE_SIMT) {
    launch_exact(ctx->stream, a, mp, 0, ctx->sm_count);
    ctx->launches += 1;
  } else {
    // key slots a query tile does not use stay "no candidate" (kKeyNone is all ones)
    if (a.n_splits > 1) CU(cudaMemsetAsync(a.keys, 0xFF, n_keys * sizeof(uint64_t), ctx->stream));
    CU(cudaMemsetAsync(ctx->d_scalars, 0, sizeof(uint32_t), ctx->stream));
    if (ctx->profiling) {
      if (ctx->prof_used == ctx->prof_events.size()) {
        cudaEvent_t e0, e1;
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        ctx->prof_events.emplace_back(e0, e1);
      }
      CU(cudaEventRecord(ctx->prof_events[ctx->prof_used].first, ctx->stream));
    }
    cudaError_t e = launch_smax(ctx->stream, grid, *q, *t, a, mp);
    if (e != cudaSuccess) return fail(ctx, static_cast<int>(e), "smax launch: %s", cudaGetErrorString(e));
    if (ctx->profiling) { CU(cudaEventRecord(ctx->prof_events[ctx->prof_used].second, ctx->stream)); ctx->prof_used++; }
    launch_exact(ctx->stream, a, mp, 1, ctx->sm_count);
    ctx->launches += 2;
  }
  CU(cudaGetLastError());
  return 0;
}

int ssrlcv_b200_seed_distances(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* query, const ssrlcv_b200_set* seed, float* out_dev, int engine) {
  if (!ctx || !query || !seed || (query->n && !out_dev)) return fail(ctx, SSRLCV_B200_EINVAL, "seed_distances: bad arguments");
  CU(cudaSetDevice(ctx->device));
  MaskParams mp; memset(&mp, 0, sizeof mp);
  SmaxArgs a;
  int rc = run_engine(ctx, query, seed, mp, engine, 0, FLT_MAX, a);
  if (rc) return rc;
  if (a.nq) { launch_seed_from_keys(ctx->stream, a.keys, a.nq, a.nq_pad, a.n_splits, out_dev); ctx->launches += 1; }
  CU(cudaGetLastError());
  return 0;
}

static size_t out_elt(int kind) {
  switch (kind) {
    case SSRLCV_B200_OUT_MATCH: return sizeof(RefMatch);
    case SSRLCV_B200_OUT_DMATCH: return sizeof(RefDMatch);
    case SSRLCV_B200_OUT_PAIR: return sizeof(RefPair);
    case SSRLCV_B200_OUT_FEATUREMATCH: return sizeof(RefFeatureMatch);
  }
  return 0;
}

// enqueue thresholds / writers / compaction; the number of records written lands in *total_dev (device)
static int finalize_enqueue(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* q, const uint64_t* keys,
                            uint32_t n_splits, const float2* t_loc, const ssrlcv_b200_set* t, const float* seed, void* out_dev,
                            uint32_t* total_dev) {
  FinalizeArgs f; memset(&f, 0, sizeof f);
  f.nq = static_cast<uint32_t>(q->n); f.nq_pad = static_cast<uint32_t>(q->n_pad); f.n_splits = n_splits;
  f.keys = keys; f.seed = seed;
  f.abs_thr = p->absolute_threshold;
  f.ratio = ratio_threshold(p->out_kind, p->mask_kind, p->relative_threshold);
  f.out_kind = p->out_kind; f.compact = p->compact;
  f.qid = p->query_image_id; f.tid = p->target_image_id;
  f.q_loc = q->loc; f.t_loc = t_loc;
  f.q_sigtheta = q->sigtheta; f.t_sigtheta = t ? t->sigtheta : nullptr;
  f.q_desc = q->desc; f.t_desc = t ? t->desc : nullptr;
  f.out = out_dev;
  const uint32_t blocks = (f.nq + 255) / 256;
  int rc = ws_reserve(ctx, ctx->counts, (blocks + 1) * sizeof(uint32_t));
  if (rc) return rc;
  f.block_counts = static_cast<uint32_t*>(ctx->counts.p);
  f.total = total_dev;
  launch_finalize(ctx->stream, f);
  if (f.nq) ctx->launches += 2;
  return 0;
}

static int finalize_common(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* q, const uint64_t* keys,
                           uint32_t n_splits, const float2* t_loc, const ssrlcv_b200_set* t, const float* seed, void* out_dev,
                           uint64_t* n_out) {
  int rc = finalize_enqueue(ctx, p, q, keys, n_splits, t_loc, t, seed, out_dev, ctx->d_scalars + 1);
  if (rc) return rc;
  CU(cudaMemcpyAsync(ctx->h_scalars + 1, ctx->d_scalars + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  if (n_out) *n_out = ctx->h_scalars[1];
  return 0;
}

static int check_params(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p) {
  if (!p) return fail(ctx, SSRLCV_B200_EINVAL, "match: params is NULL");
  if (p->out_kind < 0 || p->out_kind > 3 || p->mask_kind < 0 || p->mask_kind > 2) return fail(ctx, SSRLCV_B200_EINVAL, "match: unknown out_kind/mask_kind");
  return 0;
}

int ssrlcv_b200_match(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query, const ssrlcv_b200_set* target,
                      const float* seed_dist_dev, uint64_t n_seed_dist, void* out_dev, uint64_t* n_out) {
  if (!ctx || !query || !target || !n_out || (query->n && !out_dev)) return fail(ctx, SSRLCV_B200_EINVAL, "match: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  if (seed_dist_dev && n_seed_dist != query->n)
    return fail(ctx, SSRLCV_B200_ESEEDSIZE, "seedDistances should have come from matching a seed image to queryFeatures");
  CU(cudaSetDevice(ctx->device));
  if (p->out_kind == SSRLCV_B200_OUT_FEATUREMATCH && p->mask_kind == SSRLCV_B200_MASK_DOUBLE)
    return fail(ctx, SSRLCV_B200_EINVAL, "the reference has no DoubleConstrained FeatureMatch kernel");
  MaskParams mp; fill_mask(p, mp);
  SmaxArgs a;
  rc = run_engine(ctx, query, target, mp, p->engine, 0, p->absolute_threshold, a);
  if (rc) return rc;
  return finalize_common(ctx, p, query, a.keys, a.n_splits, target->loc, target, seed_dist_dev, out_dev, n_out);
}

// The pair scheduler's device side: every pair of the batch is enqueued back to back on the context's stream (workspaces are
// reused in stream order), the per-pair record counts are collected on the device and read back with ONE synchronisation.
int ssrlcv_b200_match_batch(ssrlcv_b200_ctx* ctx, uint32_t n_pairs, const ssrlcv_b200_match_params* params,
                            const ssrlcv_b200_set* const* queries, const ssrlcv_b200_set* const* targets,
                            const float* const* seed_dist_dev, void* const* out_dev, uint64_t* n_out) {
  if (!ctx || (n_pairs && (!params || !queries || !targets || !out_dev || !n_out))) return fail(ctx, SSRLCV_B200_EINVAL, "match_batch: bad arguments");
  if (n_pairs == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  int rc = 0;
  size_t max_keys = 0, max_rows = 0;
  for (uint32_t i = 0; i < n_pairs; ++i) {
    rc = check_params(ctx, &params[i]);
    if (rc) return rc;
    if (!queries[i] || !targets[i] || (queries[i]->n && !out_dev[i])) return fail(ctx, SSRLCV_B200_EINVAL, "match_batch: pair %u: bad arguments", i);
    if (params[i].out_kind == SSRLCV_B200_OUT_FEATUREMATCH && params[i].mask_kind == SSRLCV_B200_MASK_DOUBLE)
      return fail(ctx, SSRLCV_B200_EINVAL, "the reference has no DoubleConstrained FeatureMatch kernel");
    SmaxArgs a; int grid;
    rc = make_plan(ctx, queries[i], targets[i], params[i].engine, a, grid);
    if (rc) return rc;
    const size_t nk = static_cast<size_t>(a.n_splits) * a.nq_pad;
    if (nk > max_keys) max_keys = nk;
    if (queries[i]->n_pad > max_rows) max_rows = queries[i]->n_pad;
  }
  // grow every workspace once, up front: a reallocation in the middle of the batch would have to drain the stream
  rc = ws_reserve(ctx, ctx->keys, max_keys * sizeof(uint64_t));
  if (!rc) rc = ws_reserve(ctx, ctx->flagged, max_keys * sizeof(uint32_t));
  if (!rc) rc = ws_reserve(ctx, ctx->qmask, max_rows * 2 * sizeof(float4));
  if (!rc) rc = ws_reserve(ctx, ctx->counts, (max_rows / 256 + 2) * sizeof(uint32_t));
  if (!rc) rc = ws_reserve(ctx, ctx->totals, static_cast<size_t>(n_pairs) * sizeof(uint32_t));
  if (rc) return rc;
  if (ctx->hostpin.cap < n_pairs * sizeof(uint32_t)) {
    if (ctx->hostpin.p) { CU(cudaStreamSynchronize(ctx->stream)); CU(cudaFreeHost(ctx->hostpin.p)); ctx->hostpin.p = nullptr; ctx->hostpin.cap = 0; }
    CU(cudaMallocHost(&ctx->hostpin.p, n_pairs * sizeof(uint32_t) * 2));
    ctx->hostpin.cap = n_pairs * sizeof(uint32_t) * 2;
  }
  uint32_t* totals = static_cast<uint32_t*>(ctx->totals.p);
  for (uint32_t i = 0; i < n_pairs; ++i) {
    const ssrlcv_b200_match_params* p = &params[i];
    MaskParams mp; fill_mask(p, mp);
    SmaxArgs a;
    rc = run_engine(ctx, queries[i], targets[i], mp, p->engine, 0, p->absolute_threshold, a);
    if (!rc) rc = finalize_enqueue(ctx, p, queries[i], a.keys, a.n_splits, targets[i]->loc, targets[i],
                                   seed_dist_dev ? seed_dist_dev[i] : nullptr, out_dev[i], totals + i);
    if (rc) return rc;
  }
  CU(cudaMemcpyAsync(ctx->hostpin.p, totals, n_pairs * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  for (uint32_t i = 0; i < n_pairs; ++i) n_out[i] = static_cast<const uint32_t*>(ctx->hostpin.p)[i];
  return 0;
}

int ssrlcv_b200_match_candidates(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* p, const ssrlcv_b200_set* query,
                                 const ssrlcv_b200_set* shard, uint64_t target_offset, uint64_t* keys_dev) {
  if (!ctx || !query || !shard || (query->n && !keys_dev) || (target_offset & 31)) return fail(ctx, SSRLCV_B200_EINVAL, "match_candidates: bad arguments");
  int rc = check_params(ctx, p);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  MaskParams mp; fill_mask(p, mp);
  SmaxArgs a;
  rc = run_engine(ctx, query, shard, mp, p->engine, target_offset, p->absolute_threshold, a);
  if (rc) return rc;
  if (a.nq) {
    // reduce the local splits to one key per query, reusing the seed kernel's min loop via finalize-free path
    if (a.n_splits == 1) {
      CU(cudaMemcpyAsync(keys_dev, a.keys, static_cast<size_t>(a.nq) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      launch_min_keys(ctx->stream, a.keys, a.nq, a.nq_pad, a.n_splits, keys_dev);
      ctx->launches += 1;
    }
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
  // keys_dev is [n_shards][nq] (dense): finalize reads with stride nq_pad, so pass nq as the stride
  ssrlcv_b200_set qview = *query;
  qview.n_pad = query->n;
  return finalize_common(ctx, p, &qview, keys_dev, n_shards, reinterpret_cast<const float2*>(target_loc_dev), nullptr, seed_dist_dev, out_dev, n_out);
}

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
  if (!rc && s) rc = ssrlcv_b200_seed_distances(ctx, q, s, seed_dev, p->engine);
  if (!rc) rc = ssrlcv_b200_match(ctx, p, q, t, s ? seed_dev : nullptr, s ? n_query : 0, out_dev, n_out);
  if (!rc && *n_out) {
    cudaError_t e = cudaMemcpyAsync(out_host, out_dev, *n_out * elt, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, static_cast<int>(e), "match_host: D2H: %s", cudaGetErrorString(e));
  }
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
  uint64_t p2 = 1;
  while (p2 < n) p2 <<= 1;
  int rc = ws_reserve(ctx, ctx->keys, p2 * sizeof(uint64_t));
  if (!rc) rc = ws_reserve(ctx, ctx->scratch, n * sizeof(RefDMatch));
  if (rc) return rc;
  ctx->launches += launch_sort(ctx->stream, dmatches_dev, static_cast<uint32_t>(n), static_cast<uint64_t*>(ctx->keys.p), ctx->scratch.p);
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

// ------------------------------------------------------------------------------------------- diagnostics
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
