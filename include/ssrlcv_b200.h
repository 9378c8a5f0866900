/*
 * ssrlcv_b200.h -- C ABI of libssrlcv_match_b200.so
 *
 * B200-native (sm_100a) replacement for the SIFT descriptor matching path of
 * uga-ssrl/SSRLCV: the kernels behind ssrlcv::MatchFactory<SIFT_Descriptor>
 * (reference: include/MatchFactory.cuh:131-311, src/MatchFactory.cu).
 * The reference has no FFI layer; its "operator API" is that C++ class.  The
 * drop-in is therefore a replacement MatchFactory.cu
 * whose methods marshal Unity<> device pointers into the entry points below.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary
 *   - every function returns 0 on success, a negative SSRLCV_B200_E* code on
 *     error, or a positive cudaError_t; nothing in the library calls exit()
 *     (the reference's CudaSafeCall does, include/Memory.cuh:33-40 -- the shim
 *     maps a non-zero return to that behaviour)
 *   - "dev" pointers are device memory of the context's device; "host" are host
 *   - record layouts are the reference's (sizes in bytes):
 *       Feature<SIFT_Descriptor> 152 (include/Feature.cuh:30-38,68-94)
 *       Image::Camera 80            (include/Image.cuh:41-57)
 *       Match 40, DMatch 48, FeatureMatch<SIFT> 320, uint2_pair 16,
 *       KeyPoint 16, MultiMatch 8   (include/MatchFactory.cuh:23-72)
 *   - all work is enqueued on the context's stream; functions that return a
 *     count to the host synchronise that stream before returning, mirroring the
 *     reference's cudaDeviceSynchronize() after each kernel
 *     (src/MatchFactory.cu:383-384)
 */
#ifndef SSRLCV_B200_H
#define SSRLCV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSRLCV_B200_ABI_VERSION 2

enum {
  SSRLCV_B200_EINVAL = -1,      /* bad argument (NULL, size mismatch, unknown kind) */
  SSRLCV_B200_ENODEVICE = -2,   /* no sm_100 device / driver entry point missing    */
  SSRLCV_B200_ESEEDSIZE = -3,   /* seed distances do not belong to the query set (src/MatchFactory.cu:370-373) */
  SSRLCV_B200_ETHRESHOLD = -4,  /* refine threshold == 0 (src/MatchFactory.cu:138-141) */
  SSRLCV_B200_ENOMEM = -5
};

/* output record selected by the reference method family */
enum {
  SSRLCV_B200_OUT_MATCH = 0,        /* generateMatches*            -> Match        */
  SSRLCV_B200_OUT_DMATCH = 1,       /* generateDistanceMatches*    -> DMatch       */
  SSRLCV_B200_OUT_PAIR = 2,         /* generateMatches*IndexOnly   -> uint2_pair   */
  SSRLCV_B200_OUT_FEATUREMATCH = 3  /* generateFeatureMatches*     -> FeatureMatch */
};
/* epipolar mask applied to the candidate targets of each query */
enum {
  SSRLCV_B200_MASK_NONE = 0,        /* matchFeaturesBruteForce        */
  SSRLCV_B200_MASK_FUNDAMENTAL = 1, /* matchFeaturesConstrained       (F matrix, src/MatchFactory.cu:1598-1657) */
  SSRLCV_B200_MASK_DOUBLE = 2       /* matchFeaturesDoubleConstrained (GEO_ORBIT=1, src/MatchFactory.cu:1507-1597) */
};
/* Tuning / diagnostics knobs of a context (ssrlcv_b200_tune).  Value -1 (the default) lets the library decide; the product
 * path never needs them.  They exist for A/B timing and so that the parity tests can force every kernel variant and the
 * second, independent device implementation (CROSSCHECK_SIMT) the tensor-core path is checked against on the device. */
enum {
  SSRLCV_B200_TUNE_FOLD = 0,                  /* 0 / 1: matcher variant without / with the fold K block (smax_tc.cu)            */
  SSRLCV_B200_TUNE_COLUMN_MASKS = 1,          /* 0 / 1: bookkeeping regime of the matcher's epilogue                            */
  SSRLCV_B200_TUNE_SORTED_TARGETS = 2,        /* 0 / 1: x-sorted target copy for unseeded DoubleConstrained searches           */
  SSRLCV_B200_TUNE_HOST_CHUNK_QUERY_ROWS = 3, /* match_host: rows per upload chunk of the query set (0 = never cut)            */
  SSRLCV_B200_TUNE_HOST_CHUNK_TARGET_ROWS = 4,/* match_host: rows per upload chunk of the target set (0 = never cut)           */
  SSRLCV_B200_TUNE_CROSSCHECK_SIMT = 5,       /* 1: evaluate distances with the SIMT cross-check kernel instead of tcgen05     */
  SSRLCV_B200_TUNE_TIMING_BUFFER = 6,         /* device address of u64[148][16] cycle counters: runs the instrumented build of the
                                                 matcher kernel (tools/kerneltiming.py); 0 / -1 = off                           */
  SSRLCV_B200_TUNE_DIAG = 7,                  /* diagnostics builds of the library only (SSRLCV_B200_DIAG_BUILD; the product build rejects it):
                                                 what-if timing of the matcher (tools/diag_ab.py), RESULTS ARE WRONG when non-zero:
                                                 bit 0 = load only every other target tile, bit 1 = skip the fold MMA,
                                                 bit 2 = skip the epilogue's max trees, bit 5 = no W ring (no-fold, smax_pair.cu);
                                                 0 / -1 = off                                                                    */
  SSRLCV_B200_TUNE_PAIR_UNITS = 8,            /* 0 / 1: unmasked searches one / two query tiles per work unit (smax_tc.cu / smax_pair.cu);
                                                 2: four per CTA PAIR (smax_duo.cu, tcgen05 cta_group::2 on a cluster of two SMs) -- bit-exact,
                                                 measured slower than 1, never chosen by the library (profiles/r2_kernel_experiments.txt 8) */
  SSRLCV_B200_TUNE_MASK_COOPERATIVE = 9,      /* masked searches: most live queries of a warp for which a window's mask is evaluated by the
                                                 warp together (default 12); beyond, per row over its candidate columns (smax_tc.cu)   */
  SSRLCV_B200_TUNE_MASK_WALK = 10,            /* ... unless some row has more candidate columns than this (default 32 = never): all 32 are tested */
  SSRLCV_B200_TUNE_CELL_SORT = 11,            /* 0: spatially sorted copies always by the radix sort of Morton keys; default: sets below 65 536
                                                 rows by a one-block counting sort over image cells (pack.cu)                            */
  SSRLCV_B200_TUNE_SEED_ONLY = 12,            /* 0: seed passes through the ordinary search (winner keys); default: large seed passes take the
                                                 distances-only variant of smax_pair.cu (no bookkeeping, no resolvers)                    */
  SSRLCV_B200_TUNE_LEAN_MASKED = 13,          /* 0: seeded masked searches through the ordinary masked instantiation; default: the one with the
                                                 unmasked hot loop (smax_tc.cu kLean)                                                     */
  SSRLCV_B200_TUNE_COUNT_ = 14
};
/* internal engine ids (kept for the library's own use) */
enum { SSRLCV_B200_ENGINE_TCGEN05 = 0, SSRLCV_B200_ENGINE_SIMT = 1 };

typedef struct ssrlcv_b200_ctx ssrlcv_b200_ctx;   /* device, stream, workspaces */
typedef struct ssrlcv_b200_set ssrlcv_b200_set;   /* a feature set repacked for TMA / tcgen05 */

/* replaces: the arguments every matchFeatures* kernel receives (src/MatchFactory.cu:1463-1466,
 * 1659-1662, 1599-1601, 2195-2198) plus MatchFactory::{absolute,relative}Threshold */
typedef struct {
  int32_t  out_kind;            /* SSRLCV_B200_OUT_*  */
  int32_t  mask_kind;           /* SSRLCV_B200_MASK_* */
  float    absolute_threshold;  /* MatchFactory::absoluteThreshold (FLT_MAX by default) */
  float    relative_threshold;  /* MatchFactory::relativeThreshold; the library applies rel or rel*rel
                                   exactly as the reference kernel of (out_kind, mask_kind) does */
  uint32_t query_image_id;      /* query->id  */
  uint32_t target_image_id;     /* target->id */
  float    epsilon;             /* px; Constrained / DoubleConstrained */
  float    delta;               /* km; DoubleConstrained */
  float    fundamental[9];      /* row-major float[3][3]; Constrained */
  uint8_t  query_camera[80];    /* Image::Camera of the query image; DoubleConstrained */
  uint8_t  target_camera[80];   /* Image::Camera of the target image; the library derives the 3x4
                                   projection on the host like getProjectionMatrix (src/Image.cu:498-539) */
  int32_t  compact;             /* 1: stable-compact valid entries (validateMatches, src/MatchFactory.cu:31-135);
                                   0: keep all numQuery entries, invalid ones flagged (generateFeatureMatches
                                   quirk, src/MatchFactory.cu:674-696) */
  int32_t  reserved;            /* must be 0 */
} ssrlcv_b200_match_params;

/* ---- context ---------------------------------------------------------------------------- */
int  ssrlcv_b200_abi_version(void);
/* stream: a cudaStream_t to enqueue on, or NULL to let the library create its own. */
int  ssrlcv_b200_create(ssrlcv_b200_ctx** ctx, int device, void* stream);
void ssrlcv_b200_destroy(ssrlcv_b200_ctx* ctx);
const char* ssrlcv_b200_last_error(const ssrlcv_b200_ctx* ctx);
int  ssrlcv_b200_synchronize(ssrlcv_b200_ctx* ctx);
/* the cudaStream_t the context enqueues on (the one given to create(), or the library's own) */
void* ssrlcv_b200_stream(const ssrlcv_b200_ctx* ctx);
/* number of kernels this library has launched on the context so far (bench bookkeeping) */
uint64_t ssrlcv_b200_launch_count(const ssrlcv_b200_ctx* ctx);
/* bench instrumentation: when enabled, CUDA events bracket every launch of the tensor-core matcher kernel on the
 * context's stream; kernel_time() synchronises and returns the accumulated device time (ms) and launch count
 * since the last call, then resets both. */
int  ssrlcv_b200_tune(ssrlcv_b200_ctx* ctx, int knob, int64_t value);
int  ssrlcv_b200_set_profiling(ssrlcv_b200_ctx* ctx, int enabled);
int  ssrlcv_b200_kernel_time(ssrlcv_b200_ctx* ctx, double* total_ms, uint64_t* n_launches);

/* ---- feature sets -------------------------------------------------------------------------
 * replaces: Unity<Feature<SIFT_Descriptor>>::device as consumed by the matching kernels.  The
 * 152-byte AoS records cannot be read by TMA (stride 152 = 8 mod 16); pack() repacks them once
 * into dense uint8[n][128] rows (+ norms, locations, sigma/theta). */
int  ssrlcv_b200_pack(ssrlcv_b200_ctx* ctx, const void* features_dev, uint64_t n, ssrlcv_b200_set** out);
/* n_sets feature arrays in ONE launch (the views of an N-view job, or the query / target / seed sets of a 2-view call).
 * Enqueue-only, like pack(). */
int  ssrlcv_b200_pack_batch(ssrlcv_b200_ctx* ctx, uint32_t n_sets, const void* const* features_dev, const uint64_t* n,
                            ssrlcv_b200_set** out);
int  ssrlcv_b200_pack_host(ssrlcv_b200_ctx* ctx, const void* features_host, uint64_t n, ssrlcv_b200_set** out);
/* Producer hand-off (replaces: the 152-byte AoS stores at the end of fillDescriptors, src/SIFT_FeatureFactory.cu:533-548, and
 * with them the whole pack pass): set_alloc reserves an EMPTY packed set of n rows; set_arrays_of returns the device arrays the
 * producer's kernel writes in place -- desc[row][128] descriptor bytes, loc[row] = (x, y), sigma_theta[row] = (sigma, theta),
 * rows [0, n) (device-side helpers: include/ssrlcv_b200_packed.cuh); set_commit (enqueue-only, one launch) then derives the
 * norms, fold digits and window bounds and zero-fills the padding.  The producer's work must be ordered before the context's
 * stream when set_commit is called (same stream, or an event wait). */
typedef struct {
  uint8_t* desc;         /* uint8[n_pad][128] */
  float*   loc;          /* float2[n_pad]     */
  float*   sigma_theta;  /* float2[n_pad]     */
  uint64_t n, n_pad;
} ssrlcv_b200_set_arrays;
int  ssrlcv_b200_set_alloc(ssrlcv_b200_ctx* ctx, uint64_t n, ssrlcv_b200_set** out);
int  ssrlcv_b200_set_arrays_of(const ssrlcv_b200_set* s, ssrlcv_b200_set_arrays* out);
int  ssrlcv_b200_set_commit(ssrlcv_b200_ctx* ctx, ssrlcv_b200_set* s);
uint64_t ssrlcv_b200_set_size(const ssrlcv_b200_set* s);
void ssrlcv_b200_set_free(ssrlcv_b200_set* s);

/* ---- matching -----------------------------------------------------------------------------
 * replaces: MatchFactory<T>::getSeedDistances (src/MatchFactory.cu:315-346) + getSeedMatchDistances
 * (:1433-1460).  out_dev: float[size(query)]; FLT_MAX when the seed set is empty.  Enqueue-only: the distances are
 * complete for later calls on the same context (stream order); call ssrlcv_b200_synchronize before reading them on
 * the host.  Both sets must stay alive until then (releasing one settles the pass first). */
int  ssrlcv_b200_seed_distances(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* query, const ssrlcv_b200_set* seed,
                                float* out_dev);

/* replaces: generateMatches / generateDistanceMatches / generateMatchesIndexOnly / generateFeatureMatches
 * and their Constrained / DoubleConstrained forms (src/MatchFactory.cu:349-905) including the
 * validateMatches compaction.  seed_dist_dev: float[size(query)] from seed_distances, or NULL.
 * out_dev: capacity size(query) records of params->out_kind.  *n_out: number of records written
 * (ascending query index).  Synchronises. */
int  ssrlcv_b200_match(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params,
                       const ssrlcv_b200_set* query, const ssrlcv_b200_set* target,
                       const float* seed_dist_dev, uint64_t n_seed_dist, void* out_dev, uint64_t* n_out);

/* The same on DEVICE-resident Feature arrays in one call (pack of every set in one launch + [seed distances] + match +
 * release): what getSeedDistances + generate* do on Unity arrays in state gpu (src/Pipeline.cu:176-197).  seed_dev == NULL:
 * no seed-ratio test.  seed_dist_out_dev: float[n_query] that receives the seed distances, or NULL. */
int  ssrlcv_b200_match_features(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params, const void* query_dev,
                                uint64_t n_query, const void* target_dev, uint64_t n_target, const void* seed_dev,
                                uint64_t n_seed, float* seed_dist_out_dev, void* out_dev, uint64_t* n_out);

/* replaces: the pair loop of generateMatchesExhaustive (src/MatchFactory.cu:907-936) -- the pair scheduler's device side.
 * n_pairs independent ssrlcv_b200_match calls (params[i], queries[i], targets[i], seed_dist_dev[i] or NULL, out_dev[i]) are
 * enqueued back to back and synchronised once; n_out[i] receives the record count of pair i.  seed_dist_dev may be NULL
 * (no pair uses seed distances); seed arrays must hold size(queries[i]) floats. */
int  ssrlcv_b200_match_batch(ssrlcv_b200_ctx* ctx, uint32_t n_pairs, const ssrlcv_b200_match_params* params,
                             const ssrlcv_b200_set* const* queries, const ssrlcv_b200_set* const* targets,
                             const float* const* seed_dist_dev, void* const* out_dev, uint64_t* n_out);

/* ---- the pair scheduler across the GPUs of one box ----------------------------------------------
 * replaces: the pair loop of generateMatchesExhaustive (src/MatchFactory.cu:907-936) when more than one sm_100 device is
 * visible.  One process, one context + persistent host thread per device.  multi_create: devices == NULL takes the first
 * n_devices sm_100 devices (n_devices <= 0: all of them). */
typedef struct ssrlcv_b200_multi ssrlcv_b200_multi;
int  ssrlcv_b200_multi_create(ssrlcv_b200_multi** out, int n_devices, const int* devices);
void ssrlcv_b200_multi_destroy(ssrlcv_b200_multi* m);
int  ssrlcv_b200_multi_device_count(const ssrlcv_b200_multi* m);
const char* ssrlcv_b200_multi_last_error(const ssrlcv_b200_multi* m);
/* device time (ms, CUDA events on that device's stream) of the share device `index` ran in the last call: fetch of its views ...
 * delivery of its last pair list (bench bookkeeping; the call's time is the maximum over the devices) */
double ssrlcv_b200_multi_device_ms(const ssrlcv_b200_multi* m, int index);
/* n_pairs independent matches (pair i: query view pair_query[i], target view pair_target[i], params[i]) dealt to the devices
 * by cost with view affinity; every device fetches the views it needs (features[v]: n_features[v] Feature records behind ANY
 * unified-address pointer -- host memory or any device's), packs them once, computes the seed distances of its query views
 * (seed_features != NULL: every pair applies the seed-ratio test, like the reference's loop :915-917) and runs its pairs as
 * one batch.  out[i]: any unified-address buffer with room for n_features[pair_query[i]] records of params[i].out_kind;
 * n_out[i]: records written (ascending query index).  Returns when every device is done. */
int  ssrlcv_b200_match_pairs_multi_gpu(ssrlcv_b200_multi* m, uint32_t n_views, const void* const* features,
                                       const uint64_t* n_features, const void* seed_features, uint64_t n_seed,
                                       uint32_t n_pairs, const uint32_t* pair_query, const uint32_t* pair_target,
                                       const ssrlcv_b200_match_params* params, void* const* out, uint64_t* n_out);
/* which device the scheduler gives each pair to (device_of_pair[n_pairs]); deterministic */
int  ssrlcv_b200_multi_assignment(const ssrlcv_b200_multi* m, uint32_t n_views, const uint64_t* n_features, uint32_t n_pairs,
                                  const uint32_t* pair_query, const uint32_t* pair_target, int32_t* device_of_pair);

/* Whole reference call sequence with HOST buffers (what MatchFactory does when the Unity arrays are in
 * state `cpu`): H2D of both feature sets (+ seed set if seed_host != NULL), pack, seed distances, match,
 * D2H of the result.  out_host: capacity n_query records. */
int  ssrlcv_b200_match_host(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params,
                            const void* query_host, uint64_t n_query, const void* target_host, uint64_t n_target,
                            const void* seed_host, uint64_t n_seed, void* out_host, uint64_t* n_out);

/* Sharded reference set (multi-GPU): each rank matches the queries against its shard of the
 * target set and emits one 64-bit candidate per query, key = d2 << 37 | (f mod 32) << 32 | f with f the
 * GLOBAL target index (all-ones = no candidate); the minimum over shards is the reference's winner.
 * target_offset must be a multiple of 32. */
int  ssrlcv_b200_match_candidates(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params,
                                  const ssrlcv_b200_set* query, const ssrlcv_b200_set* target_shard,
                                  uint64_t target_offset, uint64_t* keys_dev);
/* Same, for searches that will be finalized with seed distances (ssrlcv_b200_finalize_candidates, params->compact != 0): the
 * seed-ratio test (src/MatchFactory.cu:1691-1697) applies to the merged winner, so a shard only has to track candidates with
 * d2 <= ratio * seed[q].  seed_dist_dev: float[size(query)] on the device, may be NULL (then identical to the call above). */
int  ssrlcv_b200_match_candidates_seeded(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params,
                                         const ssrlcv_b200_set* query, const ssrlcv_b200_set* target_shard,
                                         uint64_t target_offset, const float* seed_dist_dev, uint64_t* keys_dev);
/* Runner-up ("top-2") -- an extension: the reference has NO second-best output (its ratio test divides by the SEED distance,
 * src/MatchFactory.cu:1691-1697), so nothing in the MatchFactory API maps to this entry and the top-1 path does not
 * depend on it.  For hosts that want a Lowe-style ratio: a second search of the same pair in which every query's winner is left out.
 * winner_keys_dev: uint64[size(query)] keys of a previous ssrlcv_b200_match_candidates(..., target_offset = 0, ...) of this pair
 * (all-ones = no winner: nothing is excluded); keys_dev: uint64[size(query)], the best REMAINING target of every query under the
 * same winner rule (min of (d2, f mod 32, f) over the mask), all-ones = none below the absolute threshold.  Enqueue-only. */
int  ssrlcv_b200_match_runner_up(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params, const ssrlcv_b200_set* query,
                                 const ssrlcv_b200_set* target, const uint64_t* winner_keys_dev, uint64_t* keys_dev);
/* Merge n_shards candidate arrays (keys_dev: [n_shards][size(query)], e.g. the output of an all-gather),
 * then thresholds / ratio test / writers / compaction exactly as ssrlcv_b200_match.
 * target_loc_dev: float2[n_target_total] locations of the whole target set (index = global f);
 * needed for Match/DMatch outputs, may be NULL for SSRLCV_B200_OUT_PAIR. */
int  ssrlcv_b200_finalize_candidates(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params,
                                     const ssrlcv_b200_set* query, const uint64_t* keys_dev, uint32_t n_shards,
                                     const float* target_loc_dev, const float* seed_dist_dev,
                                     void* out_dev, uint64_t* n_out);
/* device pointer to the float2[n] locations of a packed set (borrowed) */
const float* ssrlcv_b200_set_loc(const ssrlcv_b200_set* s);

/* ---- post-processing ------------------------------------------------------------------------
 * replaces: validateMatches (src/MatchFactory.cu:31-135): stable removal of invalid entries, in place. */
int  ssrlcv_b200_validate(ssrlcv_b200_ctx* ctx, int out_kind, void* matches_dev, uint64_t n, uint64_t* n_out);
/* replaces: refineMatches (src/MatchFactory.cu:137-160): keep distance <= threshold, stable, in place. */
int  ssrlcv_b200_refine(ssrlcv_b200_ctx* ctx, void* dmatches_dev, uint64_t n, float threshold, uint64_t* n_out);
/* replaces: sortMatches (src/MatchFactory.cu:191-222): ascending distance; equal distances keep their
 * input order (deterministic; one valid outcome of the reference's unstable thrust::sort). */
int  ssrlcv_b200_sort(ssrlcv_b200_ctx* ctx, void* dmatches_dev, uint64_t n);
/* replaces: getRawMatches + convertMatchToRaw (src/MatchFactory.cu:257-280, :2921-2926). */
int  ssrlcv_b200_raw_matches(ssrlcv_b200_ctx* ctx, const void* dmatches_dev, uint64_t n, void* matches_dev);
/* replaces: the 2-view tail of doFeatureMatching (src/Pipeline.cu:198-222): D2H of the DMatch array, host scan for the
 * largest distance, getRawMatches, D2H, host loop writing the MatchSet.  One device pass: keypoints_dev receives
 * KeyPoint[2n] (16 B each: match i -> entries 2i, 2i+1), multimatches_dev MultiMatch[n] = {numKeyPoints 2, index 2i}
 * (8 B each), *max_distance the largest DMatch.distance (0 if n == 0; may be NULL). */
int  ssrlcv_b200_match_set_2view(ssrlcv_b200_ctx* ctx, const void* dmatches_dev, uint64_t n, void* keypoints_dev,
                                 void* multimatches_dev, float* max_distance);

/* ---- diagnostics used by the parity tests ----------------------------------------------------
 * per-query epipolar segment end points (p1.x,p1.y,p2.x,p2.y) as the DoubleConstrained prologue
 * computes them (replaces getEpipolarEndpoints, src/MatchFactory.cu:1241-1277). out_dev: float4[n]. */
int  ssrlcv_b200_epipolar_endpoints(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_match_params* params,
                                    const ssrlcv_b200_set* query, float* out_dev);
/* one array of a packed set copied to the host: which = 0 desc, 1 fold digits, 2 norms, 3 W, 4 window bounds, 5 loc,
 * 6 sigma/theta, 7 meta; *bytes = its size (host_out == NULL: size query only) */
int  ssrlcv_b200_set_download(ssrlcv_b200_ctx* ctx, const ssrlcv_b200_set* s, int which, void* host_out, uint64_t capacity,
                              uint64_t* bytes);
/* sinf/cosf of x[n] computed on the device (validates the oracle's libdevice restatement) */
int  ssrlcv_b200_device_sincos(ssrlcv_b200_ctx* ctx, const float* x_dev, uint64_t n, float* sin_dev, float* cos_dev);

#ifdef __cplusplus
}
#endif
#endif /* SSRLCV_B200_H */
