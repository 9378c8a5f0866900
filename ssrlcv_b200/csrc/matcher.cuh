// matcher.cuh -- internal types shared by the kernels and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/ssrlcv_b200.h"

namespace b200 {

// ----- reference record layouts (byte-compatible, see include/ssrlcv_b200.h) -----
struct alignas(8) RefFeature {   // Feature<SIFT_Descriptor>, include/Feature.cuh:30-38,68-94
  int32_t parent;
  int32_t pad_;
  float2  loc;
  float   sigma, theta;
  uint8_t values[128];
};
static_assert(sizeof(RefFeature) == 152, "Feature<SIFT_Descriptor> is 152 bytes");

struct RefCamera {               // Image::Camera, include/Image.cuh:41-57
  float    cam_pos[3];
  float    cam_rot[3];
  float    fov[2];
  float    foc;
  float    pad0_;
  float    dpix[2];
  long long timeStamp;
  float    ecef_offset[3];
  uint8_t  no_rot;
  uint8_t  pad1_[3];
  uint32_t size[2];
};
static_assert(sizeof(RefCamera) == 80, "Image::Camera is 80 bytes");

struct alignas(8) RefKeyPoint { int32_t parentId; int32_t pad_; float2 loc; };                         // 16
struct alignas(8) RefMatch { uint8_t invalid; uint8_t pad_[7]; RefKeyPoint kp[2]; };                  // 40
struct alignas(8) RefDMatch { uint8_t invalid; uint8_t pad_[7]; RefKeyPoint kp[2]; float distance; float pad2_; };  // 48
struct RefSift { float sigma, theta; uint8_t values[128]; };                                          // 136
struct alignas(8) RefFeatureMatch {                                                                   // 320
  uint8_t invalid; uint8_t pad_[7]; RefKeyPoint kp[2]; float distance; RefSift desc[2]; float pad2_;
};
struct alignas(8) RefPair { uint32_t a[2]; uint32_t b[2]; };                                          // 16
static_assert(sizeof(RefKeyPoint) == 16 && sizeof(RefMatch) == 40 && sizeof(RefDMatch) == 48, "match layouts");
static_assert(sizeof(RefFeatureMatch) == 320 && offsetof(RefFeatureMatch, desc) == 44, "FeatureMatch layout");
static_assert(sizeof(RefPair) == 16, "uint2_pair layout");

// What-if timing hooks of the matcher (SSRLCV_B200_TUNE_DIAG, tools/diag_ab.py) exist only in a diagnostics build of the library
// (SSRLCV_B200_DIAG_BUILD=1 python -m ssrlcv_b200.build): even one extra test in the epilogue's hot loop costs the 65 536 x 65 536
// launch 2-3 %, so the product build compiles them out.
#ifdef SSRLCV_B200_DIAG_BUILD
constexpr bool kDiagBuild = true;
#else
constexpr bool kDiagBuild = false;
#endif

// ----- tiling -----
constexpr int kBM = 128;          // query rows per CTA tile  (UMMA M, cta_group::1)
constexpr int kBN = 256;          // target rows per pipeline stage (UMMA N)
constexpr int kDescBytes = 128;   // SIFT descriptor length = K of the main contraction
constexpr int kExtBytes = 32;     // extra K block carrying the folded -|b|^2/2 term
constexpr int kRowPad = 256;      // packed sets are padded to a multiple of this many rows
constexpr int kChunk = 32;        // candidate window (targets) resolved exactly per query

// Fold constants: the extra K block multiplies the constant vector c = [1, 255 x 31] (u8) with
// per-target digits so that c . digits = W = kFoldC - (|b|^2 >> 1)  (>= 2 for every real target,
// 0 for padding rows).  Representable range of W: [0, 255 + 255*31*255], so the fold serves every set whose
// largest |b|^2 is at most 2 * (kFoldC - 2) + 1 = 4 032 057 (SIFT descriptors: ~65 000; uniform random bytes: ~2.8 M).
// Sets beyond that take the epilogue-add variant, whose W = kNoFoldC - (|b|^2 >> 1) >= 2 for ANY uint8[128] row.
// Constant offsets (rather than the set's own maximum norm) make a row's W a function of that row alone.
constexpr int32_t kFoldMaxW = 255 + 255 * 31 * 255;   // 2 016 030
constexpr int32_t kFoldC = kFoldMaxW;
constexpr int32_t kNoFoldC = 128 * 255 * 255 / 2 + 2;  // 4 161 602

// 64-bit candidate key: d2 << 37 | (f & 31) << 32 | f   (min == the reference's winner rule)
constexpr uint64_t kKeyNone = ~0ull;                 // no candidate
constexpr uint64_t kKeyNeedsExact = ~0ull - 1;       // tie between windows: resolve by the exact kernel
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t d2, uint64_t f) {
  return (static_cast<uint64_t>(d2) << 37) | ((f & 31ull) << 32) | (f & 0xFFFFFFFFull);
}

// ----- per-query epipolar mask state -----
struct MaskParams {
  int32_t kind;            // SSRLCV_B200_MASK_*
  float   eps, delta;
  float   F[9];
  RefCamera qcam;
  float   P[12];           // target projection (host, getProjectionMatrix)
};

struct ssrlcv_b200_set_impl {
  uint64_t n, n_pad;
  uint8_t* desc;      // [n_pad][128]
  uint8_t* ext;       // [n_pad][32]   fold digits
  int32_t* norm;      // [n_pad]       |b|^2
  int32_t* wcol;      // [n_pad]       kNoFoldC - (|b|^2 >> 1) (epilogue-add variant), 0 on padding rows
  float2*  loc;       // [n_pad]
  float2*  sigtheta;  // [n_pad]
  int32_t* wmax;      // [n_pad / 32]  largest wcol of each 32-row window, 0 for windows of padding only
  uint32_t* orig;     // [n_pad]       spatially sorted copies of a set only (api.cu sorted_of): row -> index in the original set
  float4*  wbox;      // [n_pad / 32]  sorted copies only: bounding box (x0, x1, y0, y1) of each 32-row window's locations
  float4*  tbox;      // [n_pad / 256] sorted copies only: the same per 256-row tile
  int32_t* meta;      // [4] device scalars written by the pack kernel: [0] != 0 <=> some W does not fit the fold digits
  bool     fold_ok;      // every W fits the digit range -> fused extra-K-block path (host copy, valid after set_resolve)
  CUtensorMap map_a;     // desc, box 128 x kBM rows, SWIZZLE_128B
  CUtensorMap map_b;     // desc, box 128 x kBN rows, SWIZZLE_128B
  CUtensorMap map_ext;   // ext,  box  32 x kBN rows, SWIZZLE_32B
  CUtensorMap map_exth;  // ext,  box  32 x kBN / 2 rows, SWIZZLE_32B (smax_duo.cu: each CTA of a pair stages half a target tile; its
                         // descriptor half is map_a's box)
};

// arguments of the fused tcgen05 kernel
struct SmaxArgs {
  uint32_t nq, nt;               // real sizes
  uint32_t nq_pad;
  uint32_t q_tile_rows;          // query rows per work unit: kBM (smax_tc.cu), 2 * kBM (smax_pair.cu: two A tiles per unit) or
                                 // 4 * kBM (smax_duo.cu: two A tiles in each CTA of a pair)
  uint32_t n_qtiles;             // ceil(nq / q_tile_rows)
  uint32_t n_ttiles;             // ceil(nt / kBN)
  uint32_t n_splits;             // key slots per query: keys is [n_splits][nq_pad]
  uint32_t tiles_per_split;      // SIMT engine only: fixed-length target splits, slot = split
  // tensor-core engine: the linearised tile space L = qtile * n_ttiles + ttile is cut into n_ctas equal contiguous
  // ranges (one per CTA); a "unit" is the part of a CTA's range inside one query tile, and writes key slot
  // (cta - first cta touching that query tile).  Balanced to one tile, whatever the shape.
  uint32_t n_ctas;
  uint64_t total_tiles;
  // a launch may cover several independent pairs (SmaxJob): the tile spaces of the jobs are laid end to end, this job's
  // tiles are [tile_begin, tile_end) of the launch's linear space; total_tiles / n_ctas describe the whole launch
  uint64_t tile_begin, tile_end;
  float    eps;                  // epipolar mask tolerance (masked modes)
  const int32_t* t_meta;         // device scalars of the target set (pack.cu): [0] == 0 <=> fold digits valid
  int32_t  d2_cap;               // largest integer d2 that still passes the absolute threshold; INT32_MAX = no threshold
  // Seeded searches whose invalid entries are compacted away: the seed-ratio test (ref: src/MatchFactory.cu:1691-1697) applies to
  // the WINNER, so a query whose best d2 exceeds ratio * seed[q] is dropped whatever else it found -- candidates beyond that
  // bound need no bookkeeping.  On real SIFT data this, not the absolute threshold, is what makes windows rare
  // (everest fixture: 84 % of the (query, window) pairs reach d2 < 200^2, 0.04 % reach 0.36 * seed[q]).
  const float* q_seed;           // [nq] seed distances (exact integers as float, FLT_MAX = none) or nullptr
  float    seed_ratio;           // rel or rel^2 as the output kind prescribes (api.cu ratio_threshold)
  // Runner-up search (ssrlcv_b200_match_runner_up): the low 32 bits of q_exclude[q] -- a winner key of a previous search of the
  // same pair, all-ones = none -- name the row of the target set that query q must not match.  It scores 0, as a masked target
  // would, wherever it could matter: in the epilogue's rare path, in the resolvers and in the exact kernels.  nullptr in every
  // reference-API call.
  const uint64_t* q_exclude;
  uint64_t target_offset;        // global index of target 0 of this set (sharded mode)
  const uint8_t* q_desc; const int32_t* q_norm; const float2* q_loc;
  const float4* q_mask;          // [nq_pad][2] per-query epipolar mask state (geometry.cuh), masked modes only
  const uint8_t* t_desc; const int32_t* t_norm; const float2* t_loc; const int32_t* t_wcol;
  const int32_t* t_wmax;         // [n_ttiles * 8] per-window maximum of t_wcol (no-fold variant)
  // Spatially sorted sets (masked searches without seed distances, api.cu sorted_of): row r of the target arrays is feature
  // t_orig[r] of the caller's set and keys carry the caller's index, so the winner rule is unchanged; row r of the query arrays
  // is query q_orig[r] and its key is stored at the caller's row.  A 32-target window then covers a small patch of the image
  // (t_wbox) and a query whose epipolar band misses the patch skips the window; whole (query tile, target tile) pairs no query
  // of which can match are marked dead beforehand (live_bits, cull_kernel) and skipped by every role before the MMA.
  const uint32_t* t_orig;        // nullptr: rows are in the caller's order
  const uint32_t* q_orig;        // nullptr: rows are in the caller's order
  const float4* t_wbox;          // [n_ttiles * 8] window boxes of the sorted target set, or nullptr
  const uint32_t* live_bits;     // [n_qtiles][live_stride] bit j of row qt: tile pair (qt, j) may hold a match; nullptr: all live
  uint32_t live_stride;
  uint64_t* keys;                // [n_splits][nq_pad]
  uint32_t* flagged;             // list of (slot * nq_pad + row) needing the exact kernel
  uint32_t* n_flagged;
};

// One (query set, target set) pair of a launch; an array of these lives in device memory.
struct alignas(128) SmaxJob {
  CUtensorMap map_q;             // query descriptors,  box 128 B x kBM rows
  CUtensorMap map_t;             // target descriptors, box 128 B x kBN rows (smax_duo.cu jobs: kBN / 2 rows)
  CUtensorMap map_text;          // target fold digits, box  32 B x kBN rows (smax_duo.cu jobs: kBN / 2 rows)
  SmaxArgs a;
};
// The launch's job table: device memory (many pairs, staged by the host) or, for a single pair, carried in the kernel's
// parameter block itself (no staging copy in front of the launch; tensor maps are read straight from the parameter space).
struct SmaxBatch {
  const SmaxJob* jobs;           // nullptr: the single job is `inline_job`
  uint32_t n_jobs;
  uint32_t n_ctas;               // ranges the unit space is cut into: CTAs, or CTA pairs (smax_duo.cu)
  uint64_t total_tiles;
  unsigned long long* timing;    // optional [gridDim.x][16] cycle counters (diagnostics build), or nullptr
  const uint32_t* flags;         // this launch's tie counters, one per job (job i counts into flags[i])
  uint32_t* clear;               // the tie counters of the OTHER bank: zeroed by CTA 0 for the call after this one
  uint32_t clear_n;
  uint32_t diag;                 // what-if timing bits (SSRLCV_B200_TUNE_DIAG); 0 in every product launch
  // masked searches, rare path (smax_tc.cu): a window that matters to at most mask_coop_max queries of a warp is tested by the warp
  // together (one query at a time, lane l = target l); beyond that every lane walks the candidate columns of its own row unless
  // some row has more than mask_walk_max of them (then all 32 columns are tested)
  uint32_t mask_coop_max, mask_walk_max;
  uint32_t has_exclude;          // some job of the launch is a runner-up search (SmaxArgs::q_exclude): read from the parameter bank,
                                 // so the rare path of every other launch pays nothing for the feature
  SmaxJob inline_job;
};

// pack.cu: one set (or one row range of a set) of a pack launch
struct PackJob {
  const RefFeature* in;          // n records; nullptr = "commit": the producer already wrote desc / loc / sigtheta of rows [0, n) in
                                 // place (include/ssrlcv_b200_packed.cuh) and only the derived arrays and the padding are written
  uint64_t n, n_rows;            // rows [0, n) are features, [n, n_rows) zero padding; n_rows % 32 == 0
  uint8_t* desc; uint8_t* ext; int32_t* norm; int32_t* wcol; int32_t* wmax; float2* loc; float2* sigtheta;   // at the range's first row
  int32_t* meta;                 // device scalars of the set (zero unless some W does not fit the fold digits)
  int32_t* h_meta;               // the same verdict in pinned, device-mapped host memory
  uint32_t win_begin;            // first 32-row window of this job in the launch's linear window space
};
constexpr int kPackInline = 8;
struct PackBatch {
  PackJob jobs[kPackInline];
  uint32_t n_jobs, total_windows;
};

// thresholds / writers / compaction of one pair (finalize.cu); an array of these lives in device memory
struct FinalizeArgs {
  uint32_t nq, nq_pad, n_splits;
  const uint64_t* keys;       // [n_splits][nq_pad]
  const float* seed;          // nullable
  float abs_thr, ratio;
  int32_t out_kind, compact;
  uint32_t qid, tid;
  const float2* q_loc; const float2* t_loc;            // t_loc indexed by global f
  const float2* q_sigtheta; const float2* t_sigtheta;  // FeatureMatch only
  const uint8_t* q_desc; const uint8_t* t_desc;
  void* out;
  // single-pass stable compaction: block b (in ticket order) publishes (epoch << 32 | count) in state[b] and sums its predecessors
  uint64_t* state;            // [ceil(nq / 256)], zero-initialised once; entries of other epochs are "not yet"
  uint32_t* ticket;           // [1] zero before the launch, reset by the block that draws the last ticket
  uint32_t epoch;             // this call's stamp (never 0)
  uint32_t* total;            // [1] device
  uint32_t* total_host;       // [1] pinned, device-mapped host word (or nullptr)
};
struct FinalizeBatch {
  const FinalizeArgs* jobs;   // device memory, or nullptr: the single job is `inline_job`
  FinalizeArgs inline_job;
};

__host__ __device__ __forceinline__ uint64_t cta_tile_begin(uint64_t total, uint32_t n_ctas, uint32_t c) {
  return (total * c) / n_ctas;
}
// the CTA whose range contains linear tile X
__host__ __device__ __forceinline__ uint32_t cta_of_tile(uint64_t total, uint32_t n_ctas, uint64_t X) {
  return static_cast<uint32_t>(((X + 1) * n_ctas - 1) / total);
}
// target tile range [j0, j1) of the unit (cta, qt); empty if the CTA does not touch that query tile
__host__ __device__ __forceinline__ void unit_range(const SmaxArgs& a, uint32_t cta, uint32_t qt, uint32_t& j0, uint32_t& j1) {
  const uint64_t lo = cta_tile_begin(a.total_tiles, a.n_ctas, cta), hi = cta_tile_begin(a.total_tiles, a.n_ctas, cta + 1);
  const uint64_t base = a.tile_begin + static_cast<uint64_t>(qt) * a.n_ttiles;
  const uint64_t b = lo > base ? lo : base, e = hi < base + a.n_ttiles ? hi : base + a.n_ttiles;
  if (b >= e) { j0 = j1 = 0; return; }
  j0 = static_cast<uint32_t>(b - base); j1 = static_cast<uint32_t>(e - base);
}

// Launch with programmatic stream serialisation: the kernel may be scheduled while the previous kernel of the stream drains
// (ptx.cuh pdl_wait / pdl_launch_dependents).  The call chain of one match is 4-8 short kernels around the matcher; the
// launch gaps between them were a fifth of the small configurations' step.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace b200

struct ssrlcv_b200_set : b200::ssrlcv_b200_set_impl {
  ssrlcv_b200_ctx* ctx;
  // pack() is enqueue-only: the set's meta scalars arrive in this pinned slot once `ready` has fired; fold_ok of the
  // impl is valid on the host only after set_resolve() (api.cu)
  int32_t* h_mm;
  cudaEvent_t ready;
  bool resolved;
  bool has_map_a, has_map_bx, has_map_exth;    // tensor maps are encoded on first use as a query / as a target (api.cu ensure_maps)
  bool is_view;                  // rows of another set (match_host_streamed): borrowed memory
  ssrlcv_b200_set* sorted;       // cached x-sorted copy (owned), or nullptr
};
