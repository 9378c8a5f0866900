/*
 * MatchFactory_b200.cu -- drop-in replacement for the reference's src/MatchFactory.cu.
 *
 * Compile this file INSTEAD of src/MatchFactory.cu (same header include/MatchFactory.cuh, same mangled
 * symbols for ssrlcv::MatchFactory<ssrlcv::SIFT_Descriptor>) and link libssrlcv_match_b200.so: the rest of
 * the SfM pipeline (SIFT_FeatureFactory -> MatchFactory -> PointCloudFactory, src/Pipeline.cu:173-241) calls
 * it unchanged.  Every method keeps the reference's observable contract:
 *   - borrowed Unity inputs are forced to MemoryState gpu for the call and restored afterwards
 *     (ref: src/MatchFactory.cu:350-353, :390-391)
 *   - outputs are new Unity arrays in state gpu whose device memory comes from ptr::device<T> (cudaMalloc,
 *     freed by the reference's own shared_ptr deleter, include/Memory.cuh:183-216)
 *   - argument errors log through `logger.err` and exit(-1) exactly where the reference does
 *   - a non-zero return of the C ABI is treated like a failed CudaSafeCall (include/Memory.cuh:33-40)
 * All of the arithmetic lives behind include/ssrlcv_b200.h; this file only marshals.
 *
 * Only the SIFT_Descriptor instantiation is provided (the Window_* SAD descriptors have no caller and are
 * out of scope); the pixel-disparity matcher and the match-file IO of the original TU
 * are likewise not part of the hot path and are not provided here.
 */
#include "MatchFactory.cuh"

#include <algorithm>
#include <cstring>
#include <map>
#include <set>

#include "ssrlcv_b200.h"

namespace {

ssrlcv_b200_ctx* b200_context() {
  static ssrlcv_b200_ctx* ctx = nullptr;
  if (!ctx) {
    int dev = 0;
    cudaGetDevice(&dev);
    // the reference runs everything on the legacy default stream and synchronises after each kernel
    int rc = ssrlcv_b200_create(&ctx, dev, (void*)cudaStreamLegacy);
    if (rc != 0) {
      logger.err.printf("ERROR: ssrlcv_b200_create failed with code %d (an sm_100 GPU is required)", rc);
      exit(-1);
    }
  }
  return ctx;
}

void b200_check(int rc, const char* what) {
  if (rc == 0) return;
  logger.err.printf("ERROR: %s failed with code %d: %s", what, rc, ssrlcv_b200_last_error(b200_context()));
  exit(-1);
}

struct PackedSet {   // RAII over ssrlcv_b200_set
  ssrlcv_b200_set* s = nullptr;
  PackedSet(void* features_dev, unsigned long n) { b200_check(ssrlcv_b200_pack(b200_context(), features_dev, n, &s), "ssrlcv_b200_pack"); }
  ~PackedSet() { ssrlcv_b200_set_free(s); }
  PackedSet(const PackedSet&) = delete;
  PackedSet& operator=(const PackedSet&) = delete;
};

template <typename Out, typename T>
ssrlcv::ptr::value<ssrlcv::Unity<Out>> b200_generate(
    float relativeThreshold, float absoluteThreshold, int out_kind, int mask_kind, bool compact,
    ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Feature<T>>> queryFeatures,
    ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Feature<T>>> targetFeatures,
    float epsilon, float delta, float (*fundamental)[3], ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances,
    const PackedSet* queryPacked = nullptr, const PackedSet* targetPacked = nullptr) {
  using namespace ssrlcv;
  static_assert(sizeof(Feature<T>) == 152, "only Feature<SIFT_Descriptor> is supported");
  MemoryState origin[2] = {queryFeatures->getMemoryState(), targetFeatures->getMemoryState()};
  if (origin[0] != gpu) queryFeatures->setMemoryState(gpu);
  if (origin[1] != gpu) targetFeatures->setMemoryState(gpu);

  const unsigned long numPossibleMatches = queryFeatures->size();
  // The reference allocates numPossibleMatches records, compacts into a second allocation and frees the first
  // (src/MatchFactory.cu:31-60, :355-357): two cudaMalloc and a cudaFree of megabyte blocks cost milliseconds, more than the
  // whole match.  Here the records land in a grow-only scratch buffer first and the Unity the caller gets is allocated once,
  // at its final size.
  static void* scratch = nullptr;
  static size_t scratch_bytes = 0;
  if (scratch_bytes < numPossibleMatches * sizeof(Out)) {
    if (scratch) CudaSafeCall(cudaFree(scratch));
    scratch_bytes = numPossibleMatches * sizeof(Out) + (numPossibleMatches * sizeof(Out)) / 2 + 256;
    CudaSafeCall(cudaMalloc(&scratch, scratch_bytes));
  }
  ptr::value<Unity<Out>> matches = nullptr;

  ssrlcv_b200_match_params p;
  memset(&p, 0, sizeof p);
  p.out_kind = out_kind;
  p.mask_kind = mask_kind;
  p.absolute_threshold = absoluteThreshold;
  p.relative_threshold = relativeThreshold;
  p.query_image_id = (unsigned int)query->id;
  p.target_image_id = (unsigned int)target->id;
  p.epsilon = epsilon;
  p.delta = delta;
  if (fundamental) memcpy(p.fundamental, fundamental, 9 * sizeof(float));
  static_assert(sizeof(Image::Camera) == 80, "Image::Camera layout");
  memcpy(p.query_camera, &query->camera, 80);
  memcpy(p.target_camera, &target->camera, 80);
  p.compact = compact ? 1 : 0;
  p.engine = SSRLCV_B200_ENGINE_TCGEN05;

  clock_t timer = clock();
  MemoryState seedOrigin = gpu;
  const float* seed_dev = nullptr;
  if (seedDistances != nullptr) {
    if (seedDistances->size() != queryFeatures->size()) {
      logger.err << "ERROR: seedDistances should have come from matching a seed image to queryFeatures";
      exit(-1);
    }
    seedOrigin = seedDistances->getMemoryState();
    if (seedOrigin != gpu) seedDistances->setMemoryState(gpu);
    seed_dev = seedDistances->device.get();
  }
  {
    PackedSet* own_q = queryPacked ? nullptr : new PackedSet(queryFeatures->device.get(), queryFeatures->size());
    PackedSet* own_t = targetPacked ? nullptr : new PackedSet(targetFeatures->device.get(), targetFeatures->size());
    const ssrlcv_b200_set* qs = queryPacked ? queryPacked->s : own_q->s;
    const ssrlcv_b200_set* ts = targetPacked ? targetPacked->s : own_t->s;
    uint64_t n_out = 0;
    b200_check(ssrlcv_b200_match(b200_context(), &p, qs, ts, seed_dev, seed_dev ? numPossibleMatches : 0, scratch, &n_out), "ssrlcv_b200_match");
    if (compact) {
      // validateMatches (ref: src/MatchFactory.cu:31-135), already fused into the call above
      if (n_out == 0) {
        logger.info << "No valid matches found";
        // reference quirk: validateMatches clears only its own copy of the smart pointer, so the caller keeps
        // the uncompacted array -> materialise it
        p.compact = 0;
        b200_check(ssrlcv_b200_match(b200_context(), &p, qs, ts, seed_dev, seed_dev ? numPossibleMatches : 0, scratch, &n_out), "ssrlcv_b200_match");
      } else {
        logger.info.printf("%d valid matches found out of %lu original matches", (int)n_out, numPossibleMatches);
      }
    }
    ptr::device<Out> matches_device(n_out);
    if (n_out) CudaSafeCall(cudaMemcpy(matches_device.get(), scratch, n_out * sizeof(Out), cudaMemcpyDeviceToDevice));
    matches = ptr::value<Unity<Out>>(matches_device, n_out, gpu);
    delete own_q;
    delete own_t;
  }
  if (seedDistances != nullptr && seedOrigin != gpu) seedDistances->setMemoryState(seedOrigin);
  logger.info.printf("done in %f seconds.", ((float)clock() - timer) / CLOCKS_PER_SEC);

  if (origin[0] != gpu) queryFeatures->setMemoryState(origin[0]);
  if (origin[1] != gpu) targetFeatures->setMemoryState(origin[1]);
  return matches;
}

template <typename U>
void b200_compact_inplace(U matches, int out_kind, size_t elt, bool refine, float threshold) {
  using namespace ssrlcv;
  MemoryState origin = matches->getMemoryState();
  if (origin != gpu) matches->setMemoryState(gpu);
  uint64_t n_out = 0;
  if (refine) b200_check(ssrlcv_b200_refine(b200_context(), matches->device.get(), matches->size(), threshold, &n_out), "ssrlcv_b200_refine");
  else b200_check(ssrlcv_b200_validate(b200_context(), out_kind, matches->device.get(), matches->size(), &n_out), "ssrlcv_b200_validate");
  if (n_out == 0) {
    if (!refine) logger.info << "No valid matches found";
    return;   // the reference resets/deletes only its local handle here; the caller's array is left as is
  }
  if (refine) logger.info.printf("%lu matches have been refined to %u matches using a cutoff of %f", matches->size(), (unsigned)n_out, threshold);
  else logger.info.printf("%d valid matches found out of %lu original matches", (int)n_out, matches->size());
  typedef typename std::remove_reference<decltype(*matches->device.get())>::type Elt;
  ptr::device<Elt> kept(n_out);
  CudaSafeCall(cudaMemcpy(kept.get(), matches->device.get(), n_out * elt, cudaMemcpyDeviceToDevice));
  matches->setData(kept, n_out, gpu);
  if (origin != gpu) matches->setMemoryState(origin);
}

}  // namespace

template class ssrlcv::MatchFactory<ssrlcv::SIFT_Descriptor>;

// ------------------------------------------------------------------------------------------------ construction
template <typename T>
ssrlcv::MatchFactory<T>::MatchFactory() {
  this->relativeThreshold = 1.0f;
  this->absoluteThreshold = FLT_MAX;
  this->seedFeatures = nullptr;
}
template <typename T>
ssrlcv::MatchFactory<T>::MatchFactory(float relativeThreshold, float absoluteThreshold)
    : relativeThreshold(relativeThreshold), absoluteThreshold(absoluteThreshold) {
  this->seedFeatures = nullptr;
}
template <typename T>
void ssrlcv::MatchFactory<T>::setSeedFeatures(ssrlcv::ptr::value<Unity<Feature<T>>> seedFeatures) {
  this->seedFeatures = seedFeatures;
}

// ------------------------------------------------------------------------------------------------ validate / refine / sort / raw
template <typename T>
void ssrlcv::MatchFactory<T>::validateMatches(ssrlcv::ptr::value<ssrlcv::Unity<uint2_pair>> matches) {
  b200_compact_inplace(matches, SSRLCV_B200_OUT_PAIR, sizeof(uint2_pair), false, 0.0f);
}
template <typename T>
void ssrlcv::MatchFactory<T>::validateMatches(ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Match>> matches) {
  b200_compact_inplace(matches, SSRLCV_B200_OUT_MATCH, sizeof(Match), false, 0.0f);
}
template <typename T>
void ssrlcv::MatchFactory<T>::validateMatches(ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::DMatch>> matches) {
  b200_compact_inplace(matches, SSRLCV_B200_OUT_DMATCH, sizeof(DMatch), false, 0.0f);
}
template <typename T>
void ssrlcv::MatchFactory<T>::validateMatches(ssrlcv::Unity<ssrlcv::FeatureMatch<T>>* matches) {
  b200_compact_inplace(matches, SSRLCV_B200_OUT_FEATUREMATCH, sizeof(FeatureMatch<T>), false, 0.0f);
}
template <typename T>
void ssrlcv::MatchFactory<T>::refineMatches(ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::DMatch>> matches, float threshold) {
  if (threshold == 0.0f) {
    logger.err << "ERROR illegal value used for threshold: 0.0";
    exit(-1);
  }
  b200_compact_inplace(matches, SSRLCV_B200_OUT_DMATCH, sizeof(DMatch), true, threshold);
}
template <typename T>
void ssrlcv::MatchFactory<T>::refineMatches(ssrlcv::Unity<ssrlcv::FeatureMatch<T>>* matches, float threshold) {
  if (threshold == 0.0f) {
    logger.err << "ERROR illegal value used for cutoff ratio: 0.0";
    exit(-1);
  }
  // FeatureMatch records carry `distance` at the same offset as DMatch but are 320 bytes: compact through validate's
  // generic path after flagging (no caller in the reference; kept for API completeness)
  MemoryState origin = matches->getMemoryState();
  if (origin != cpu) matches->setMemoryState(cpu);
  FeatureMatch<T>* h = matches->host.get();
  unsigned long kept = 0;
  for (unsigned long i = 0; i < matches->size(); ++i)
    if (!(h[i].distance > threshold)) h[kept++] = h[i];
  if (kept == 0) return;
  ssrlcv::ptr::host<FeatureMatch<T>> compacted(kept);
  memcpy((void*)compacted.get(), (void*)h, kept * sizeof(FeatureMatch<T>));
  matches->setData(compacted, kept, cpu);
  if (origin != cpu) matches->setMemoryState(origin);
}
template <typename T>
void ssrlcv::MatchFactory<T>::sortMatches(ssrlcv::ptr::value<ssrlcv::Unity<DMatch>> matches) {
  if (matches->getFore() == gpu || matches->getFore() == both) {
    b200_check(ssrlcv_b200_sort(b200_context(), matches->device.get(), matches->size()), "ssrlcv_b200_sort");
    matches->setFore(gpu);
    if (matches->getMemoryState() == both) matches->transferMemoryTo(cpu);
  } else if (matches->getFore() == cpu) {
    DMatch* mhost = matches->host.get();
    std::stable_sort(mhost, mhost + matches->size(), [](const DMatch& a, const DMatch& b) { return a.distance < b.distance; });
    if (matches->getMemoryState() == both) matches->transferMemoryTo(gpu);
  } else {
    logger.err << "ERROR cannot perform sortMatches with matches->getMemoryState() = " << std::to_string(matches->getMemoryState());
    exit(-1);
  }
}
template <typename T>
void ssrlcv::MatchFactory<T>::sortMatches(Unity<FeatureMatch<T>>* matches) {
  MemoryState origin = matches->getMemoryState();
  if (origin == null) {
    logger.err << "ERROR cannot perform sortMatches with matches->getMemoryState() = " << std::to_string(origin);
    exit(-1);
  }
  if (matches->getFore() == gpu) matches->transferMemoryTo(cpu);
  FeatureMatch<T>* mhost = matches->host.get();
  std::stable_sort(mhost, mhost + matches->size(), [](const FeatureMatch<T>& a, const FeatureMatch<T>& b) { return a.distance < b.distance; });
  matches->setFore(cpu);
  if (origin == gpu || origin == both) matches->transferMemoryTo(gpu);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Match>> ssrlcv::MatchFactory<T>::getRawMatches(ssrlcv::ptr::value<ssrlcv::Unity<DMatch>> matches) {
  if (matches->getMemoryState() == gpu || matches->getFore() == gpu) {
    ssrlcv::ptr::device<Match> rawMatches_device(matches->size());
    b200_check(ssrlcv_b200_raw_matches(b200_context(), matches->device.get(), matches->size(), rawMatches_device.get()), "ssrlcv_b200_raw_matches");
    return ssrlcv::ptr::value<Unity<Match>>(rawMatches_device, matches->size(), gpu);
  }
  ssrlcv::ptr::host<Match> rawMatches_host(matches->size());
  for (unsigned long i = 0; i < matches->size(); ++i) rawMatches_host.get()[i] = Match(matches->host.get()[i]);
  return ssrlcv::ptr::value<Unity<Match>>(rawMatches_host, matches->size(), cpu);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Match>> ssrlcv::MatchFactory<T>::getRawMatches(Unity<FeatureMatch<T>>* matches) {
  MemoryState origin = matches->getMemoryState();
  if (origin == gpu) matches->transferMemoryTo(cpu);
  ssrlcv::ptr::host<Match> rawMatches_host(matches->size());
  for (unsigned long i = 0; i < matches->size(); ++i) rawMatches_host.get()[i] = Match(matches->host.get()[i]);
  ssrlcv::ptr::value<Unity<Match>> raw(rawMatches_host, matches->size(), cpu);
  if (origin == gpu) raw->setMemoryState(gpu);
  return raw;
}

// ------------------------------------------------------------------------------------------------ seed distances
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<float>> ssrlcv::MatchFactory<T>::getSeedDistances(ssrlcv::ptr::value<Unity<Feature<T>>> features) {
  MemoryState origin = features->getMemoryState();
  if (this->seedFeatures->getMemoryState() != gpu) this->seedFeatures->setMemoryState(gpu);
  if (origin != gpu) features->setMemoryState(gpu);
  unsigned int numPossibleMatches = features->size();
  ssrlcv::ptr::value<ssrlcv::Unity<float>> matchDistances = ssrlcv::ptr::value<ssrlcv::Unity<float>>(nullptr, numPossibleMatches, gpu);
  clock_t timer = clock();
  {
    PackedSet q(features->device.get(), features->size());
    PackedSet s(this->seedFeatures->device.get(), this->seedFeatures->size());
    b200_check(ssrlcv_b200_seed_distances(b200_context(), q.s, s.s, matchDistances->device.get(), SSRLCV_B200_ENGINE_TCGEN05),
               "ssrlcv_b200_seed_distances");
    b200_check(ssrlcv_b200_synchronize(b200_context()), "ssrlcv_b200_synchronize");
  }
  logger.info.printf("seed match distances computed in %f seconds.", ((float)clock() - timer) / CLOCKS_PER_SEC);
  if (origin != gpu) features->setMemoryState(origin);
  return matchDistances;
}

// ------------------------------------------------------------------------------------------------ generate*
#define B200_ARGS query, queryFeatures, target, targetFeatures
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Match>> ssrlcv::MatchFactory<T>::generateMatches(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<Match, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_MATCH, SSRLCV_B200_MASK_NONE, true, B200_ARGS, 0.f, 0.f, nullptr, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Match>> ssrlcv::MatchFactory<T>::generateMatchesConstrained(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float fundamental[3][3], ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<Match, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_MATCH, SSRLCV_B200_MASK_FUNDAMENTAL, true, B200_ARGS, epsilon, 0.f, fundamental, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Match>> ssrlcv::MatchFactory<T>::generateMatchesDoubleConstrained(ssrlcv::ptr::value<Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float delta, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<Match, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_MATCH, SSRLCV_B200_MASK_DOUBLE, true, B200_ARGS, epsilon, delta, nullptr, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::DMatch>> ssrlcv::MatchFactory<T>::generateDistanceMatches(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<DMatch, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_DMATCH, SSRLCV_B200_MASK_NONE, true, B200_ARGS, 0.f, 0.f, nullptr, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::DMatch>> ssrlcv::MatchFactory<T>::generateDistanceMatchesConstrained(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float fundamental[3][3], ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<DMatch, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_DMATCH, SSRLCV_B200_MASK_FUNDAMENTAL, true, B200_ARGS, epsilon, 0.f, fundamental, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::DMatch>> ssrlcv::MatchFactory<T>::generateDistanceMatchesDoubleConstrained(ssrlcv::ptr::value<Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float delta, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<DMatch, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_DMATCH, SSRLCV_B200_MASK_DOUBLE, true, B200_ARGS, epsilon, delta, nullptr, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::uint2_pair>> ssrlcv::MatchFactory<T>::generateMatchesIndexOnly(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<uint2_pair, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_PAIR, SSRLCV_B200_MASK_NONE, true, B200_ARGS, 0.f, 0.f, nullptr, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::uint2_pair>> ssrlcv::MatchFactory<T>::generateMatchesConstrainedIndexOnly(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float fundamental[3][3], ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<uint2_pair, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_PAIR, SSRLCV_B200_MASK_FUNDAMENTAL, true, B200_ARGS, epsilon, 0.f, fundamental, seedDistances);
}
template <typename T>
ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::uint2_pair>> ssrlcv::MatchFactory<T>::generateMatchesDoubleConstrainedIndexOnly(ssrlcv::ptr::value<Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float delta, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  return b200_generate<uint2_pair, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_PAIR, SSRLCV_B200_MASK_DOUBLE, true, B200_ARGS, epsilon, delta, nullptr, seedDistances);
}
// FeatureMatch: raw Unity pointers; the unconstrained form returns all numQuery entries uncompacted (ref quirk, :674-696)
template <typename T>
ssrlcv::Unity<ssrlcv::FeatureMatch<T>>* ssrlcv::MatchFactory<T>::generateFeatureMatches(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  ssrlcv::ptr::value<Unity<FeatureMatch<T>>> m = b200_generate<FeatureMatch<T>, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_FEATUREMATCH, SSRLCV_B200_MASK_NONE, false, B200_ARGS, 0.f, 0.f, nullptr, seedDistances);
  return new Unity<FeatureMatch<T>>(m->device, m->size(), gpu);
}
template <typename T>
ssrlcv::Unity<ssrlcv::FeatureMatch<T>>* ssrlcv::MatchFactory<T>::generateFeatureMatchesConstrained(ssrlcv::ptr::value<ssrlcv::Image> query, ssrlcv::ptr::value<Unity<Feature<T>>> queryFeatures, ssrlcv::ptr::value<ssrlcv::Image> target, ssrlcv::ptr::value<Unity<Feature<T>>> targetFeatures, float epsilon, float fundamental[3][3], ssrlcv::ptr::value<ssrlcv::Unity<float>> seedDistances) {
  ssrlcv::ptr::value<Unity<FeatureMatch<T>>> m = b200_generate<FeatureMatch<T>, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_FEATUREMATCH, SSRLCV_B200_MASK_FUNDAMENTAL, true, B200_ARGS, epsilon, 0.f, fundamental, seedDistances);
  return new Unity<FeatureMatch<T>>(m->device, m->size(), gpu);
}
#undef B200_ARGS

// ------------------------------------------------------------------------------------------------ N-view
// Pair scheduler + fusion (ref: src/MatchFactory.cu:907-1028).  Every view is packed once (the reference re-uploads
// and re-reads the AoS records for every pair); the adjacency derivation is the reference's algorithm.
template <typename T>
ssrlcv::MatchSet ssrlcv::MatchFactory<T>::generateMatchesExhaustive(std::vector<ssrlcv::ptr::value<ssrlcv::Image>> images, std::vector<ssrlcv::ptr::value<ssrlcv::Unity<ssrlcv::Feature<T>>>> features, float epsilon, float delta, bool ordered, float estimatedOverlap) {
  MatchSet matchSet;
  matchSet.keyPoints = nullptr;
  matchSet.matches = nullptr;
  if (estimatedOverlap == 0) logger.warn << "WARNING: estimated overlap fraction of 0.0f requires unordered match interpolation";
  const size_t n = images.size();
  std::vector<MemoryState> origin(n);
  std::vector<PackedSet*> packed(n, nullptr);
  for (size_t v = 0; v < n; ++v) {
    origin[v] = features[v]->getMemoryState();
    if (origin[v] != gpu) features[v]->setMemoryState(gpu);
    packed[v] = new PackedSet(features[v]->device.get(), features[v]->size());
  }
  std::vector<ssrlcv::ptr::value<ssrlcv::Unity<uint2_pair>>> matchIndices;
  unsigned long long totalMatches = 0;
  logger.info << "matching images";
  // The pair scheduler: seed distances once per query view, then every pair of the reference's loop (same order,
  // same skip rule, ref: src/MatchFactory.cu:925-936) is handed to the library in ONE batch and synchronised once.
#if GEO_ORBIT == 1
  const int mask = SSRLCV_B200_MASK_DOUBLE;
#else
  const int mask = SSRLCV_B200_MASK_NONE;
#endif
  std::vector<ssrlcv::ptr::value<ssrlcv::Unity<float>>> seedOfView(n, nullptr);
  std::vector<ssrlcv_b200_match_params> params;
  std::vector<const ssrlcv_b200_set*> qsets, tsets;
  std::vector<const float*> seeds;
  std::vector<void*> outs;
  std::vector<std::pair<size_t, size_t>> pairOf;
  int i = 0;
  for (size_t q = 0; q + 1 < n; ++q) {
    if (this->seedFeatures != nullptr) seedOfView[q] = this->getSeedDistances(features[q]);
    for (size_t t = q + 1; t < n; ++t) {
      if (ordered && estimatedOverlap > 0.0f && ++i * (1 - estimatedOverlap) > 1.0f) continue;
      ++i;
      ssrlcv_b200_match_params p;
      memset(&p, 0, sizeof p);
      p.out_kind = SSRLCV_B200_OUT_PAIR;
      p.mask_kind = mask;
      p.absolute_threshold = absoluteThreshold;
      p.relative_threshold = relativeThreshold;
      p.query_image_id = (unsigned int)images[q]->id;
      p.target_image_id = (unsigned int)images[t]->id;
      p.epsilon = epsilon;
      p.delta = delta;
      memcpy(p.query_camera, &images[q]->camera, 80);
      memcpy(p.target_camera, &images[t]->camera, 80);
      p.compact = 1;
      p.engine = SSRLCV_B200_ENGINE_TCGEN05;
      ptr::device<uint2_pair> dev(features[q]->size());
      matchIndices.push_back(ptr::value<Unity<uint2_pair>>(dev, features[q]->size(), gpu));
      params.push_back(p);
      qsets.push_back(packed[q]->s);
      tsets.push_back(packed[t]->s);
      seeds.push_back(seedOfView[q] != nullptr ? seedOfView[q]->device.get() : nullptr);
      outs.push_back(matchIndices.back()->device.get());
      pairOf.push_back(std::make_pair(q, t));
    }
  }
  std::vector<uint64_t> counts(params.size(), 0);
  b200_check(ssrlcv_b200_match_batch(b200_context(), (uint32_t)params.size(), params.data(), qsets.data(), tsets.data(), seeds.data(),
                                     outs.data(), counts.data()),
             "ssrlcv_b200_match_batch");
  for (size_t k = 0; k < params.size(); ++k) {
    if (counts[k] == 0) {
      // reference quirk (validateMatches, src/MatchFactory.cu:118-123): the caller keeps the uncompacted all-invalid array
      matchIndices[k] = b200_generate<uint2_pair, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_PAIR, mask, true, images[pairOf[k].first],
                                                     features[pairOf[k].first], images[pairOf[k].second], features[pairOf[k].second], epsilon, delta,
                                                     nullptr, seedOfView[pairOf[k].first], packed[pairOf[k].first], packed[pairOf[k].second]);
    } else {
      ptr::device<uint2_pair> validated(counts[k]);
      CudaSafeCall(cudaMemcpy(validated.get(), matchIndices[k]->device.get(), counts[k] * sizeof(uint2_pair), cudaMemcpyDeviceToDevice));
      matchIndices[k]->setData(validated, counts[k], gpu);
    }
    totalMatches += matchIndices[k]->size();
  }
  for (size_t v = 0; v < n; ++v) delete packed[v];
  if (totalMatches == 0) {
    logger.err << "There were no matches found in the set of images, likely due to unreasonable threshold";
    logger.err << "exiting...";
    exit(0);
  }
  logger.info << "building adjacency list";
  typedef std::pair<unsigned int, unsigned int> Node;   // (view, feature); std::pair orders like uint2's operator<
  std::vector<std::map<unsigned int, std::vector<Node>>> adjacency(n ? n - 1 : 0);
  for (auto& m : matchIndices) {
    if (m->getMemoryState() != cpu) m->setMemoryState(cpu);
    const uint2_pair* h = m->host.get();
    for (unsigned long p = 0; p < m->size(); ++p) adjacency[h[p].a.x][h[p].a.y].push_back(Node(h[p].b.x, h[p].b.y));
  }
  logger.info << "deriving matches from adjacency";
  std::vector<std::vector<Node>> multiMatch_vec;
  for (size_t v = 0; v + 1 < n; ++v) {
    if (features[v]->getMemoryState() != cpu && features[v]->getMemoryState() != both) features[v]->setMemoryState(both);
    for (unsigned long f = 0; v + 2 < n && f < features[v]->size(); ++f) {
      auto it = adjacency[v].find((unsigned int)f);
      if (it == adjacency[v].end() || it->second.empty()) continue;
      std::vector<Node>* adj = &it->second;
      bool badMatch = false;
      std::vector<Node>* prev_adj = adj;
      while (true) {
        if (prev_adj->begin()->first == n - 1) break;
        std::vector<Node>* next_adj = &adjacency[prev_adj->begin()->first][prev_adj->begin()->second];
        if (next_adj->empty()) break;
        std::vector<Node> intersection;
        std::set_intersection(prev_adj->begin(), prev_adj->end(), next_adj->begin(), next_adj->end(), std::back_inserter(intersection));
        if (intersection.size() != next_adj->size()) { badMatch = true; break; }
        else if (next_adj->size() == 1) break;
        else prev_adj = next_adj;
      }
      if (badMatch) { adj->clear(); continue; }
      std::vector<Node> match;
      match.push_back(Node((unsigned int)v, (unsigned int)f));
      match.insert(match.end(), adj->begin(), adj->end());
      multiMatch_vec.push_back(match);
      for (auto m = adj->begin(); m != adj->end() - 1; ++m) {
        if (m->first == n - 1) break;
        adjacency[m->first][m->second].clear();
      }
    }
  }
  logger.info.printf("total matches found in set = %d", (int)multiMatch_vec.size());
  for (size_t v = 0; v < n; ++v)
    if (features[v]->getMemoryState() == gpu) features[v]->setMemoryState(both);
  matchSet.matches = ssrlcv::ptr::value<ssrlcv::Unity<MultiMatch>>(nullptr, multiMatch_vec.size(), cpu);
  std::vector<KeyPoint> kp_vec;
  int index = 0;
  MultiMatch* mhost = matchSet.matches->host.get();
  size_t mi = 0;
  for (auto& m : multiMatch_vec) {
    mhost[mi++] = {(unsigned int)m.size(), index};
    index += (int)m.size();
    for (auto& kp : m) kp_vec.push_back({(int)kp.first, features[kp.first]->host.get()[kp.second].loc});
  }
  matchSet.keyPoints = ssrlcv::ptr::value<ssrlcv::Unity<KeyPoint>>(nullptr, kp_vec.size(), gpu);
  CudaSafeCall(cudaMemcpy(matchSet.keyPoints->device.get(), &kp_vec[0], kp_vec.size() * sizeof(KeyPoint), cudaMemcpyHostToDevice));
  for (size_t v = 0; v < n; ++v)
    if (features[v]->getMemoryState() != origin[v]) features[v]->setMemoryState(origin[v]);
  return matchSet;
}
