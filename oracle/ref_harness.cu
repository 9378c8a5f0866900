/*
 * ref_harness.cu -- TEST/BASELINE INFRASTRUCTURE ONLY.
 *
 * A thin extern "C" face over the UNMODIFIED reference MatchFactory<SIFT_Descriptor>
 * (compiled from the sources where they lie under /root/reference by build_ref.sh;
 * no reference source is copied into this repository).  It lets the parity tests
 * and bench.py call the reference's own CUDA kernels on the B200 box with plain
 * host buffers, through the reference's public API (ref: include/MatchFactory.cuh:131-311).
 *
 * Everything in this file is our code; it only *uses* the reference's classes.
 */
#include "MatchFactory.cuh"

#include <chrono>
#include <cstring>
#include <sys/stat.h>
#include <vector>

using namespace ssrlcv;
typedef Feature<SIFT_Descriptor> Feat;
typedef ptr::value<Unity<Feat>> FeatSet;

/* ---- the five io_util symbols Image.cu references (ref: src/io_util.cpp:27-70), so that
 *      libpng/libjpeg/libtiff are not needed ---- */
namespace ssrlcv {
bool fileExists(std::string fileName) { std::ifstream f(fileName.c_str()); return f.good(); }
std::string getFileExtension(std::string path) {
  size_t p = path.find_last_of('.');
  return p == std::string::npos ? std::string() : path.substr(p + 1);
}
std::string getFolderFromFilePath(std::string path) {
  size_t p = path.find_last_of('/');
  return p == std::string::npos ? std::string(".") : path.substr(0, p);
}
std::string getFileFromFilePath(std::string path) {
  size_t p = path.find_last_of('/');
  return p == std::string::npos ? path : path.substr(p + 1);
}
ptr::host<unsigned char> readImage(const char*, unsigned int&, unsigned int&, unsigned int&) {
  fprintf(stderr, "ref_harness: readImage is not available (no image codecs linked)\n");
  abort();
}
}  // namespace ssrlcv

static FeatSet make_set(const void* host, unsigned long n, int state) {
  FeatSet u(nullptr, n, cpu);
  memcpy((void*)u->host.get(), host, n * sizeof(Feat));
  if (state == 2) u->setMemoryState(gpu);
  else if (state == 3) u->setMemoryState(both);
  return u;
}

static ptr::value<Image> make_image(int id, const void* camera80) {
  ptr::value<Image> img;
  img.construct();
  img->id = id;
  if (camera80) memcpy((void*)&img->camera, camera80, sizeof(Image::Camera));
  img->size = img->camera.size;
  return img;
}

template <typename U>
static unsigned long copy_out(U m, void* out_host, size_t elt) {
  if (m == nullptr) return 0;
  unsigned long n = m->size();
  if (m->getMemoryState() == gpu) m->transferMemoryTo(cpu);
  memcpy(out_host, (void*)m->host.get(), n * elt);
  return n;
}

extern "C" {

int ref_sizes(int* feat, int* camera, int* match, int* dmatch, int* pair, int* fmatch) {
  *feat = sizeof(Feat); *camera = sizeof(Image::Camera); *match = sizeof(Match); *dmatch = sizeof(DMatch);
  *pair = sizeof(uint2_pair); *fmatch = sizeof(FeatureMatch<SIFT_Descriptor>);
  return 0;
}

/* getSeedDistances (ref: src/MatchFactory.cu:315-346).  state: 1 cpu, 2 gpu, 3 both. */
int ref_seed_distances(const void* q, unsigned long nq, const void* seed, unsigned long ns, int state,
                       float* out, double* ms) {
  mkdir("out", 0755);
  FeatSet qs = make_set(q, nq, state), ss = make_set(seed, ns, state);
  MatchFactory<SIFT_Descriptor> mf(1.0f, FLT_MAX);
  mf.setSeedFeatures(ss);
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  ptr::value<Unity<float>> d = mf.getSeedDistances(qs);
  cudaDeviceSynchronize();
  if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  d->transferMemoryTo(cpu);
  memcpy(out, d->host.get(), nq * sizeof(float));
  return 0;
}

/* out_kind: 0 Match, 1 DMatch, 2 uint2_pair, 3 FeatureMatch (mask_kind 0 or 1 only: the reference has no DoubleConstrained
 * FeatureMatch kernel; the unconstrained form returns all nq entries uncompacted, ref: src/MatchFactory.cu:654-697).
 * mask_kind: 0 none, 1 fundamental, 2 double constrained.
 * seed_dist may be NULL.  Returns the number of entries the reference returned (after its own
 * validateMatches), copied to out_host.  ms = wall time of the API call (inputs already in `state`). */
long ref_match(int out_kind, int mask_kind, float abs_thr, float rel_thr, int qid, int tid,
               const void* q, unsigned long nq, const void* t, unsigned long nt, const float* seed_dist, int state,
               float eps, float delta, const float* F9, const void* qcam80, const void* tcam80,
               void* out_host, double* ms) {
  mkdir("out", 0755);
  FeatSet qs = make_set(q, nq, state), ts = make_set(t, nt, state);
  ptr::value<Image> qi = make_image(qid, qcam80), ti = make_image(tid, tcam80);
  MatchFactory<SIFT_Descriptor> mf(rel_thr, abs_thr);
  ptr::value<Unity<float>> sd = nullptr;
  if (seed_dist) {
    sd = ptr::value<Unity<float>>(nullptr, nq, cpu);
    memcpy(sd->host.get(), seed_dist, nq * sizeof(float));
    if (state != 1) sd->setMemoryState(gpu);
  }
  float F[3][3];
  if (F9) memcpy(F, F9, sizeof F);
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  long n = -1;
  ptr::value<Unity<Match>> m0 = nullptr;
  ptr::value<Unity<DMatch>> m1 = nullptr;
  ptr::value<Unity<uint2_pair>> m2 = nullptr;
  Unity<FeatureMatch<SIFT_Descriptor>>* m3 = nullptr;     // raw pointer API (ref: include/MatchFactory.cuh:288,293)
  if (out_kind == 3 && mask_kind == 2) return -2;
  if (out_kind == 0) {
    if (mask_kind == 0) m0 = mf.generateMatches(qi, qs, ti, ts, sd);
    else if (mask_kind == 1) m0 = mf.generateMatchesConstrained(qi, qs, ti, ts, eps, F, sd);
    else m0 = mf.generateMatchesDoubleConstrained(qi, qs, ti, ts, eps, delta, sd);
  } else if (out_kind == 1) {
    if (mask_kind == 0) m1 = mf.generateDistanceMatches(qi, qs, ti, ts, sd);
    else if (mask_kind == 1) m1 = mf.generateDistanceMatchesConstrained(qi, qs, ti, ts, eps, F, sd);
    else m1 = mf.generateDistanceMatchesDoubleConstrained(qi, qs, ti, ts, eps, delta, sd);
  } else if (out_kind == 2) {
    if (mask_kind == 0) m2 = mf.generateMatchesIndexOnly(qi, qs, ti, ts, sd);
    else if (mask_kind == 1) m2 = mf.generateMatchesConstrainedIndexOnly(qi, qs, ti, ts, eps, F, sd);
    else m2 = mf.generateMatchesDoubleConstrainedIndexOnly(qi, qs, ti, ts, eps, delta, sd);
  } else if (out_kind == 3) {
    if (mask_kind == 0) m3 = mf.generateFeatureMatches(qi, qs, ti, ts, sd);
    else m3 = mf.generateFeatureMatchesConstrained(qi, qs, ti, ts, eps, F, sd);
  } else {
    return -2;
  }
  cudaDeviceSynchronize();
  if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (out_kind == 0) n = (long)copy_out(m0, out_host, sizeof(Match));
  else if (out_kind == 1) n = (long)copy_out(m1, out_host, sizeof(DMatch));
  else if (out_kind == 2) n = (long)copy_out(m2, out_host, sizeof(uint2_pair));
  else { n = (long)copy_out(m3, out_host, sizeof(FeatureMatch<SIFT_Descriptor>)); delete m3; }
  return n;
}

/* generateMatchesExhaustive (ref: src/MatchFactory.cu:907-1028) over n_views host feature arrays.
 * Returns the number of MultiMatch entries (copied to out_mm, capacity cap_mm); *n_kp keypoints to out_kp. */
long ref_exhaustive2(int n_views, const void* const* feats, const unsigned long* n_feats, const void* const* cams80, const int* ids,
                     const void* seed, unsigned long n_seed, float rel_thr, float abs_thr, float eps, float delta, int state,
                     int ordered, float estimated_overlap,
                     void* out_mm, unsigned long cap_mm, void* out_kp, unsigned long cap_kp, unsigned long* n_kp, double* ms) {
  mkdir("out", 0755);
  std::vector<ptr::value<Image>> images;
  std::vector<FeatSet> features;
  for (int v = 0; v < n_views; ++v) {
    images.push_back(make_image(ids[v], cams80[v]));
    features.push_back(make_set(feats[v], n_feats[v], state));
  }
  MatchFactory<SIFT_Descriptor> mf(rel_thr, abs_thr);
  if (seed) mf.setSeedFeatures(make_set(seed, n_seed, 2));
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  MatchSet set = mf.generateMatchesExhaustive(images, features, eps, delta, ordered != 0, estimated_overlap);
  cudaDeviceSynchronize();
  if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (set.matches->size() > cap_mm || set.keyPoints->size() > cap_kp) return -3;
  if (set.matches->getMemoryState() == gpu) set.matches->transferMemoryTo(cpu);
  if (set.keyPoints->getMemoryState() == gpu) set.keyPoints->transferMemoryTo(cpu);
  memcpy(out_mm, (void*)set.matches->host.get(), set.matches->size() * sizeof(MultiMatch));
  memcpy(out_kp, (void*)set.keyPoints->host.get(), set.keyPoints->size() * sizeof(KeyPoint));
  *n_kp = set.keyPoints->size();
  return (long)set.matches->size();
}

long ref_exhaustive(int n_views, const void* const* feats, const unsigned long* n_feats, const void* const* cams80, const int* ids,
                    const void* seed, unsigned long n_seed, float rel_thr, float abs_thr, float eps, float delta, int state,
                    void* out_mm, unsigned long cap_mm, void* out_kp, unsigned long cap_kp, unsigned long* n_kp, double* ms) {
  return ref_exhaustive2(n_views, feats, n_feats, cams80, ids, seed, n_seed, rel_thr, abs_thr, eps, delta, state, 1, 0.0f,
                         out_mm, cap_mm, out_kp, cap_kp, n_kp, ms);
}

/* validateMatches (ref: src/MatchFactory.cu:31-135) through the class API on a host record array of out_kind 0..3 (in place);
 * returns the size of the caller's array afterwards (unchanged when nothing was valid -- the reference's quirk, except for the raw
 * FeatureMatch pointer, which the reference deletes: not driven with an all-invalid array). */
long ref_validate(int out_kind, void* records, unsigned long n) {
  mkdir("out", 0755);
  MatchFactory<SIFT_Descriptor> mf;
  if (out_kind == 0) {
    ptr::value<Unity<Match>> m(nullptr, n, cpu);
    memcpy((void*)m->host.get(), records, n * sizeof(Match));
    m->setMemoryState(gpu); mf.validateMatches(m); m->transferMemoryTo(cpu);
    memcpy(records, (void*)m->host.get(), m->size() * sizeof(Match));
    return (long)m->size();
  } else if (out_kind == 1) {
    ptr::value<Unity<DMatch>> m(nullptr, n, cpu);
    memcpy((void*)m->host.get(), records, n * sizeof(DMatch));
    m->setMemoryState(gpu); mf.validateMatches(m); m->transferMemoryTo(cpu);
    memcpy(records, (void*)m->host.get(), m->size() * sizeof(DMatch));
    return (long)m->size();
  } else if (out_kind == 2) {
    ptr::value<Unity<uint2_pair>> m(nullptr, n, cpu);
    memcpy((void*)m->host.get(), records, n * sizeof(uint2_pair));
    m->setMemoryState(gpu); mf.validateMatches(m); m->transferMemoryTo(cpu);
    memcpy(records, (void*)m->host.get(), m->size() * sizeof(uint2_pair));
    return (long)m->size();
  } else if (out_kind == 3) {
    typedef FeatureMatch<SIFT_Descriptor> FM;
    Unity<FM>* m = new Unity<FM>(nullptr, n, cpu);
    memcpy((void*)m->host.get(), records, n * sizeof(FM));
    m->setMemoryState(gpu); mf.validateMatches(m); m->transferMemoryTo(cpu);
    memcpy(records, (void*)m->host.get(), m->size() * sizeof(FM));
    long k = (long)m->size();
    delete m;
    return k;
  }
  return -2;
}

/* refineMatches + sortMatches + getRawMatches through the class API on a host DMatch array (in place); returns kept count */
long ref_refine_sort(void* dmatches, unsigned long n, float threshold, int do_sort, void* raw_out) {
  mkdir("out", 0755);
  ptr::value<Unity<DMatch>> m(nullptr, n, cpu);
  memcpy((void*)m->host.get(), dmatches, n * sizeof(DMatch));
  m->setMemoryState(gpu);
  MatchFactory<SIFT_Descriptor> mf;
  if (threshold != 0.0f) mf.refineMatches(m, threshold);
  if (do_sort) mf.sortMatches(m);
  ptr::value<Unity<Match>> raw = mf.getRawMatches(m);
  m->transferMemoryTo(cpu);
  raw->transferMemoryTo(cpu);
  memcpy(dmatches, (void*)m->host.get(), m->size() * sizeof(DMatch));
  if (raw_out) memcpy(raw_out, (void*)raw->host.get(), raw->size() * sizeof(Match));
  return (long)m->size();
}

}  // extern "C"
