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
#include <chrono>
#include <map>
#include <set>
#include <vector>

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

// More than one sm_100 device visible: generateMatchesExhaustive deals its pairs to all of them (one context + host thread per
// device inside the library).  SSRLCV_B200_MAX_DEVICES (read once) caps the number of devices; 1 keeps everything on the
// current device.
ssrlcv_b200_multi* b200_multi() {
  static ssrlcv_b200_multi* multi = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    int cap = 0;
    if (const char* e = getenv("SSRLCV_B200_MAX_DEVICES")) cap = atoi(e);
    int visible = 0;
    cudaGetDeviceCount(&visible);
    if (visible > 1 && cap != 1) {
      ssrlcv_b200_multi* m = nullptr;
      if (ssrlcv_b200_multi_create(&m, cap, nullptr) == 0) {
        if (ssrlcv_b200_multi_device_count(m) > 1) multi = m;
        else ssrlcv_b200_multi_destroy(m);
      }
    }
  }
  return multi;
}

void b200_check(int rc, const char* what) {
  if (rc == 0) return;
  logger.err.printf("ERROR: %s failed with code %d: %s", what, rc, ssrlcv_b200_last_error(b200_context()));
  exit(-1);
}

// Read-only inputs: a Unity in state `both` whose device side is current needs no transition (the reference drops the host copy
// and re-downloads it afterwards, src/MatchFactory.cu:350-353 / :390-391 -- two transfers of the whole array per call).
template <typename U>
inline bool b200_device_current(const U& u) {
  const ssrlcv::MemoryState st = u->getMemoryState();
  return st == ssrlcv::gpu || (st == ssrlcv::both && u->getFore() != ssrlcv::cpu);
}

// Device memory for the Unity arrays this file returns.  ptr::device<T>(n) is a cudaMalloc whose shared_ptr deleter cudaFrees
// (include/Memory.cuh:183-216): milliseconds per call for megabyte blocks, more than the whole match.  The Unity only needs a
// ptr::device<T> that owns its memory, so the block comes from the device's stream-ordered pool instead and the deleter hands it
// back to the pool (both on the legacy default stream the reference runs on): microseconds, same ownership semantics.
template <typename E>
inline ssrlcv::ptr::device<E> b200_pool_device(size_t n) {
  ssrlcv::ptr::device<E> d;
  E* raw = nullptr;
  CudaSafeCall(cudaMallocAsync((void**)&raw, (n ? n : 1) * sizeof(E), cudaStreamLegacy));
  d.ptr.reset(raw, [](E* p) { cudaFreeAsync(p, cudaStreamLegacy); });
  return d;
}

struct PackedSet {   // RAII over ssrlcv_b200_set
  ssrlcv_b200_set* s = nullptr;
  PackedSet(void* features_dev, unsigned long n) { b200_check(ssrlcv_b200_pack(b200_context(), features_dev, n, &s), "ssrlcv_b200_pack"); }
  explicit PackedSet(ssrlcv_b200_set* adopted) : s(adopted) {}
  ~PackedSet() { ssrlcv_b200_set_free(s); }
  PackedSet(const PackedSet&) = delete;
  PackedSet& operator=(const PackedSet&) = delete;
};
// Packs several device arrays with ONE launch (ssrlcv_b200_pack_batch).  Packed sets are never cached across calls: a
// (device pointer, size) key cannot tell a buffer from a later allocation at the same address, and validating the contents
// costs as much as repacking them (one HBM pass); within a call every array is packed exactly once.
inline void b200_pack_all(const std::vector<const void*>& dev, const std::vector<uint64_t>& n, std::vector<PackedSet*>& out) {
  std::vector<ssrlcv_b200_set*> sets(dev.size(), nullptr);
  b200_check(ssrlcv_b200_pack_batch(b200_context(), (uint32_t)dev.size(), dev.data(), n.data(), sets.data()), "ssrlcv_b200_pack_batch");
  out.resize(dev.size());
  for (size_t i = 0; i < dev.size(); ++i) out[i] = new PackedSet(sets[i]);
}

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
  if (!b200_device_current(queryFeatures)) queryFeatures->setMemoryState(gpu);
  if (!b200_device_current(targetFeatures)) targetFeatures->setMemoryState(gpu);

  const unsigned long numPossibleMatches = queryFeatures->size();
  // The reference allocates numPossibleMatches records, compacts into a second allocation and frees the first
  // (src/MatchFactory.cu:31-60, :355-357): two cudaMalloc and a cudaFree of megabyte blocks cost milliseconds, more than the
  // whole match.  Here the records are written once, compacted, into a pool block of that capacity, which the returned Unity
  // then owns (its size is the number of valid matches).
  ptr::device<Out> matches_device = b200_pool_device<Out>(numPossibleMatches);
  void* scratch = (void*)matches_device.get();
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

  clock_t timer = clock();
  MemoryState seedOrigin = gpu;
  const float* seed_dev = nullptr;
  if (seedDistances != nullptr) {
    if (seedDistances->size() != queryFeatures->size()) {
      logger.err << "ERROR: seedDistances should have come from matching a seed image to queryFeatures";
      exit(-1);
    }
    seedOrigin = seedDistances->getMemoryState();
    if (!b200_device_current(seedDistances)) seedDistances->setMemoryState(gpu);
    seed_dev = seedDistances->device.get();
  }
  {
    PackedSet* own_q = nullptr;
    PackedSet* own_t = nullptr;
    if (!queryPacked && !targetPacked) {                   // the usual 2-view call: both sets in one pack launch
      std::vector<PackedSet*> both;
      b200_pack_all({queryFeatures->device.get(), targetFeatures->device.get()}, {queryFeatures->size(), targetFeatures->size()}, both);
      own_q = both[0]; own_t = both[1];
    } else {
      if (!queryPacked) own_q = new PackedSet(queryFeatures->device.get(), queryFeatures->size());
      if (!targetPacked) own_t = new PackedSet(targetFeatures->device.get(), targetFeatures->size());
    }
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
    matches = ptr::value<Unity<Out>>(matches_device, n_out, gpu);
    delete own_q;
    delete own_t;
  }
  if (seedDistances != nullptr && seedDistances->getMemoryState() != seedOrigin) seedDistances->setMemoryState(seedOrigin);
  logger.info.printf("done in %f seconds.", ((float)clock() - timer) / CLOCKS_PER_SEC);

  if (queryFeatures->getMemoryState() != origin[0]) queryFeatures->setMemoryState(origin[0]);
  if (targetFeatures->getMemoryState() != origin[1]) targetFeatures->setMemoryState(origin[1]);
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
  ptr::device<Elt> kept = b200_pool_device<Elt>(n_out);
  CudaSafeCall(cudaMemcpyAsync(kept.get(), matches->device.get(), n_out * elt, cudaMemcpyDeviceToDevice, cudaStreamLegacy));
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
    ssrlcv::ptr::device<Match> rawMatches_device = b200_pool_device<Match>(matches->size());
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
  if (!b200_device_current(this->seedFeatures)) this->seedFeatures->setMemoryState(gpu);
  if (!b200_device_current(features)) features->setMemoryState(gpu);
  unsigned int numPossibleMatches = features->size();
  ssrlcv::ptr::value<ssrlcv::Unity<float>> matchDistances = ssrlcv::ptr::value<ssrlcv::Unity<float>>(b200_pool_device<float>(numPossibleMatches), numPossibleMatches, gpu);
  clock_t timer = clock();
  {
    std::vector<PackedSet*> qs;
    b200_pack_all({features->device.get(), this->seedFeatures->device.get()}, {features->size(), this->seedFeatures->size()}, qs);
    b200_check(ssrlcv_b200_seed_distances(b200_context(), qs[0]->s, qs[1]->s, matchDistances->device.get()), "ssrlcv_b200_seed_distances");
    b200_check(ssrlcv_b200_synchronize(b200_context()), "ssrlcv_b200_synchronize");
    delete qs[0];
    delete qs[1];
  }
  logger.info.printf("seed match distances computed in %f seconds.", ((float)clock() - timer) / CLOCKS_PER_SEC);
  if (features->getMemoryState() != origin) features->setMemoryState(origin);
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
#ifdef SSRLCV_B200_DEBUG_PHASES
  auto phase0 = std::chrono::steady_clock::now();
  auto phase = [&](const char* what) {
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "generateMatchesExhaustive: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - phase0).count());
    phase0 = now;
  };
#else
  auto phase = [](const char*) {};
#endif
  std::vector<MemoryState> origin(n);
  std::vector<PackedSet*> packed(n, nullptr);
  PackedSet* seedPacked = nullptr;
  ssrlcv_b200_multi* multi = n > 2 ? b200_multi() : nullptr;
  for (size_t v = 0; v < n; ++v) {
    origin[v] = features[v]->getMemoryState();
    if (!b200_device_current(features[v])) features[v]->setMemoryState(gpu);
  }
  if (this->seedFeatures != nullptr && !b200_device_current(this->seedFeatures)) this->seedFeatures->setMemoryState(gpu);
  {
    // every view (and the seed image's features) is packed once on this device, all of them in ONE launch (with several
    // devices the pairs are matched elsewhere, but the key-point locations below still come from these packed sets)
    std::vector<const void*> dev;
    std::vector<uint64_t> cnt;
    for (size_t v = 0; v < n; ++v) { dev.push_back(features[v]->device.get()); cnt.push_back(features[v]->size()); }
    if (this->seedFeatures != nullptr && !multi) { dev.push_back(this->seedFeatures->device.get()); cnt.push_back(this->seedFeatures->size()); }
    b200_pack_all(dev, cnt, packed);
    if (this->seedFeatures != nullptr && !multi) { seedPacked = packed.back(); packed.pop_back(); }
  }
  unsigned long long totalMatches = 0;
  phase("states + pack");
  logger.info << "matching images";
  // The pair scheduler: seed distances once per query view, then every pair of the reference's loop (same order,
  // same skip rule, ref: src/MatchFactory.cu:925-936) is handed to the library in ONE batch and synchronised once --
  // on this device, or dealt to every sm_100 device of the box (ssrlcv_b200_match_pairs_multi_gpu).
  // The index pairs land in one grow-only device buffer (the reference cudaMallocs, compacts into a second allocation and
  // frees per pair, :909-936 via :853-905) and only the valid ones travel to the host.
#if GEO_ORBIT == 1
  const int mask = SSRLCV_B200_MASK_DOUBLE;
#else
  const int mask = SSRLCV_B200_MASK_NONE;
#endif
  std::vector<ssrlcv::ptr::value<ssrlcv::Unity<float>>> seedOfView(n, nullptr);
  std::vector<ssrlcv_b200_match_params> params;
  std::vector<const ssrlcv_b200_set*> qsets, tsets;
  std::vector<const float*> seeds;
  std::vector<size_t> outOffset;
  std::vector<std::pair<size_t, size_t>> pairOf;
  size_t outRecords = 0;
  int i = 0;
  for (size_t q = 0; q + 1 < n; ++q) {
    if (seedPacked != nullptr) {
      // getSeedDistances (ref: :915-917) on the sets that are packed already; stream-ordered before the batch below
      clock_t timer = clock();
      seedOfView[q] = ssrlcv::ptr::value<ssrlcv::Unity<float>>(b200_pool_device<float>(features[q]->size()), features[q]->size(), gpu);
      b200_check(ssrlcv_b200_seed_distances(b200_context(), packed[q]->s, seedPacked->s, seedOfView[q]->device.get()), "ssrlcv_b200_seed_distances");
      logger.info.printf("seed match distances computed in %f seconds.", ((float)clock() - timer) / CLOCKS_PER_SEC);
    }
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
      params.push_back(p);
      if (!multi) {
        qsets.push_back(packed[q]->s);
        tsets.push_back(packed[t]->s);
        seeds.push_back(seedOfView[q] != nullptr ? seedOfView[q]->device.get() : nullptr);
      }
      outOffset.push_back(outRecords);
      outRecords += features[q]->size();
      pairOf.push_back(std::make_pair(q, t));
    }
  }
  static uint2_pair* pairScratch = nullptr;
  static size_t pairScratchRecords = 0;
  if (pairScratchRecords < outRecords) {
    if (pairScratch) CudaSafeCall(cudaFree(pairScratch));
    pairScratchRecords = outRecords + outRecords / 2 + 64;
    CudaSafeCall(cudaMalloc(&pairScratch, pairScratchRecords * sizeof(uint2_pair)));
  }
  std::vector<void*> outs(params.size());
  for (size_t k = 0; k < params.size(); ++k) outs[k] = pairScratch + outOffset[k];
  std::vector<uint64_t> counts(params.size(), 0);
  if (multi) {
    std::vector<const void*> dev(n);
    std::vector<uint64_t> cnt(n);
    std::vector<uint32_t> pq(params.size()), pt(params.size());
    for (size_t v = 0; v < n; ++v) { dev[v] = features[v]->device.get(); cnt[v] = features[v]->size(); }
    for (size_t k = 0; k < params.size(); ++k) { pq[k] = (uint32_t)pairOf[k].first; pt[k] = (uint32_t)pairOf[k].second; }
    CudaSafeCall(cudaDeviceSynchronize());               // the workers' streams do not order themselves after the legacy stream
    int rc = ssrlcv_b200_match_pairs_multi_gpu(multi, (uint32_t)n, dev.data(), cnt.data(),
                                               this->seedFeatures != nullptr ? this->seedFeatures->device.get() : nullptr,
                                               this->seedFeatures != nullptr ? this->seedFeatures->size() : 0, (uint32_t)params.size(), pq.data(),
                                               pt.data(), params.data(), outs.data(), counts.data());
    if (rc != 0) {
      logger.err.printf("ERROR: ssrlcv_b200_match_pairs_multi_gpu failed with code %d: %s", rc, ssrlcv_b200_multi_last_error(multi));
      exit(-1);
    }
  } else {
    b200_check(ssrlcv_b200_match_batch(b200_context(), (uint32_t)params.size(), params.data(), qsets.data(), tsets.data(), seeds.data(),
                                       outs.data(), counts.data()),
               "ssrlcv_b200_match_batch");
  }
  phase("seed distances + batch");
  std::vector<std::vector<uint2_pair>> matchIndices(params.size());
  for (size_t k = 0; k < params.size(); ++k) {
    if (counts[k] == 0) {
      // reference quirk (validateMatches, src/MatchFactory.cu:118-123): the caller keeps the uncompacted all-invalid array
      if (this->seedFeatures != nullptr && seedOfView[pairOf[k].first] == nullptr) seedOfView[pairOf[k].first] = this->getSeedDistances(features[pairOf[k].first]);
      ptr::value<Unity<uint2_pair>> all = b200_generate<uint2_pair, T>(relativeThreshold, absoluteThreshold, SSRLCV_B200_OUT_PAIR, mask, true,
                                                                      images[pairOf[k].first], features[pairOf[k].first], images[pairOf[k].second],
                                                                      features[pairOf[k].second], epsilon, delta, nullptr, seedOfView[pairOf[k].first],
                                                                      packed[pairOf[k].first], packed[pairOf[k].second]);
      all->setMemoryState(cpu);
      matchIndices[k].assign(all->host.get(), all->host.get() + all->size());
    } else {
      matchIndices[k].resize(counts[k]);
      CudaSafeCall(cudaMemcpy(matchIndices[k].data(), outs[k], counts[k] * sizeof(uint2_pair), cudaMemcpyDeviceToHost));
    }
    totalMatches += matchIndices[k].size();
  }
  if (totalMatches == 0) {
    logger.err << "There were no matches found in the set of images, likely due to unreasonable threshold";
    logger.err << "exiting...";
    exit(0);
  }
  phase("records to host");
  // Key-point locations: 8 bytes per feature from the packed sets instead of the whole 152-byte records the reference moves to
  // the host (:966-967); the Unity arrays were forced to state gpu above, so none of them has a host copy right now.
  std::vector<std::vector<float2>> locOf(n);
  for (size_t v = 0; v < n; ++v) {
    locOf[v].resize(features[v]->size());
    if (!locOf[v].empty())
      CudaSafeCall(cudaMemcpy(locOf[v].data(), ssrlcv_b200_set_loc(packed[v]->s), locOf[v].size() * sizeof(float2), cudaMemcpyDeviceToHost));
  }
  for (size_t v = 0; v < n; ++v) delete packed[v];
  delete seedPacked;

  // Match interpolation (ref: src/MatchFactory.cu:943-1020), same visiting order and the same side effects, on flat arrays:
  // a feature of view v has at most one partner per later view, so its adjacency list is a fixed-stride slot of (n - 1 - v)
  // nodes, filled in ascending partner view (the pair order) -- already sorted the way std::set_intersection needs it.
  phase("locations to host");
  logger.info << "building adjacency list";
  typedef std::pair<unsigned int, unsigned int> Node;   // (view, feature); std::pair orders like uint2's operator<
  std::vector<std::vector<Node>> adjNodes(n ? n - 1 : 0);
  std::vector<std::vector<unsigned short>> adjCount(n ? n - 1 : 0);
  std::vector<size_t> stride(n ? n - 1 : 0);
  for (size_t v = 0; v + 1 < n; ++v) {
    stride[v] = n - 1 - v;
    adjNodes[v].resize(features[v]->size() * stride[v]);
    adjCount[v].assign(features[v]->size(), 0);
  }
  for (auto& m : matchIndices)
    for (const uint2_pair& pr : m) {
      const size_t v = pr.a.x, f = pr.a.y;
      if (v + 1 >= n || f >= adjCount[v].size() || adjCount[v][f] >= stride[v]) {
        logger.err << "ERROR: image ids must be 0..n-1 in view order for match interpolation";
        exit(-1);
      }
      adjNodes[v][f * stride[v] + adjCount[v][f]++] = Node(pr.b.x, pr.b.y);
    }
  logger.info << "deriving matches from adjacency";
  std::vector<Node> flatMatches;            // all multi-matches back to back
  std::vector<MultiMatch> mm_vec;
  for (size_t v = 0; v + 2 < n; ++v) {
    const unsigned long nf = features[v]->size();
    for (unsigned long f = 0; f < nf; ++f) {
      const unsigned short cnt = adjCount[v][f];
      if (!cnt) continue;
      const Node* adj = &adjNodes[v][f * stride[v]];
      bool badMatch = false;
      const Node* prev = adj;
      unsigned short prevCnt = cnt;
      while (true) {
        if (prev[0].first == n - 1) break;
        const size_t nv = prev[0].first, nfeat = prev[0].second;
        const unsigned short nextCnt = adjCount[nv][nfeat];
        if (!nextCnt) break;
        const Node* next = &adjNodes[nv][nfeat * stride[nv]];
        unsigned short a = 0, b = 0, common = 0;      // |prev ∩ next| by merge (both ascending)
        while (a < prevCnt && b < nextCnt) {
          if (prev[a] < next[b]) ++a;
          else if (next[b] < prev[a]) ++b;
          else { ++common; ++a; ++b; }
        }
        if (common != nextCnt) { badMatch = true; break; }
        else if (nextCnt == 1) break;
        prev = next; prevCnt = nextCnt;
      }
      if (badMatch) { adjCount[v][f] = 0; continue; }
      mm_vec.push_back({(unsigned int)(cnt + 1), (int)flatMatches.size()});
      flatMatches.push_back(Node((unsigned int)v, (unsigned int)f));
      flatMatches.insert(flatMatches.end(), adj, adj + cnt);
      for (unsigned short m = 0; m + 1 < cnt; ++m) {
        if (adj[m].first == n - 1) break;
        adjCount[adj[m].first][adj[m].second] = 0;
      }
    }
  }
  phase("match interpolation");
  logger.info.printf("total matches found in set = %d", (int)mm_vec.size());
  matchSet.matches = ssrlcv::ptr::value<ssrlcv::Unity<MultiMatch>>(nullptr, mm_vec.size(), cpu);
  if (!mm_vec.empty()) memcpy((void*)matchSet.matches->host.get(), mm_vec.data(), mm_vec.size() * sizeof(MultiMatch));
  std::vector<KeyPoint> kp_vec(flatMatches.size());
  for (size_t k = 0; k < flatMatches.size(); ++k) {
    const Node& nd = flatMatches[k];
    const float2 loc = locOf[nd.first][nd.second];
    kp_vec[k] = {(int)nd.first, loc};
  }
  matchSet.keyPoints = ssrlcv::ptr::value<ssrlcv::Unity<KeyPoint>>(nullptr, kp_vec.size(), gpu);
  CudaSafeCall(cudaMemcpy(matchSet.keyPoints->device.get(), &kp_vec[0], kp_vec.size() * sizeof(KeyPoint), cudaMemcpyHostToDevice));
  phase("MatchSet");
  for (size_t v = 0; v < n; ++v)
    if (features[v]->getMemoryState() != origin[v]) features[v]->setMemoryState(origin[v]);
  phase("restore states");
  return matchSet;
}
