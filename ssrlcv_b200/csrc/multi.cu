// multi.cu -- the pair scheduler across the sm_100 devices of one box, behind the C ABI (include/ssrlcv_b200.h).
//
// replaces: the serial pair loop of generateMatchesExhaustive (ref: src/MatchFactory.cu:907-936), which re-uploads and
// re-reads both feature arrays for every pair on one GPU.  Image pairs are independent work units: ONE process
// drives every device through one context and one persistent host thread per device.  A call
//   1. deals the pairs to the devices by cost Nq x Nt with view affinity (a device fetches and packs only the views it
//      touches -- the same rule as ssrlcv_b200/pair_scheduler.py::assign_pairs),
//   2. on every device, concurrently: fetches the views it needs from wherever they live (any unified-address pointer: host
//      memory or any device; peer copies go over NVLink), packs them in one launch, computes the seed distances of its
//      query views, runs all of its pairs as ONE batch (ssrlcv_b200_match_batch: one persistent tensor-core launch, one
//      synchronisation) and sends every pair's records to the caller's buffer for that pair,
//   3. returns when every device is done: records and counts are indexed by pair, i.e. in the reference's pair order.
// There is no collective: the exchange is the direct copy of each pair list to its destination.
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ssrlcv_b200.h"

namespace {

constexpr size_t kFeatureBytes = 152;   // Feature<SIFT_Descriptor>, include/Feature.cuh:30-38,68-94

size_t record_bytes(int out_kind) {
  switch (out_kind) {
    case SSRLCV_B200_OUT_MATCH: return 40;
    case SSRLCV_B200_OUT_DMATCH: return 48;
    case SSRLCV_B200_OUT_PAIR: return 16;
    case SSRLCV_B200_OUT_FEATUREMATCH: return 320;
  }
  return 0;
}

// one call's arguments, shared read-only by the workers
struct Call {
  uint32_t n_views = 0;
  const void* const* features = nullptr;
  const uint64_t* n_features = nullptr;
  const void* seed = nullptr;
  uint64_t n_seed = 0;
  uint32_t n_pairs = 0;
  const uint32_t* pair_query = nullptr;
  const uint32_t* pair_target = nullptr;
  const ssrlcv_b200_match_params* params = nullptr;
  void* const* out = nullptr;
  uint64_t* n_out = nullptr;
};

struct Buffer {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return static_cast<int>(e);
    cap = want;
    return 0;
  }
};

struct Worker {
  int device = 0;
  ssrlcv_b200_ctx* ctx = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  bool has_job = false, stop = false, idle = true;
  const Call* call = nullptr;
  std::vector<uint32_t> pairs;     // this call's share (indices into the call's pair arrays)
  int rc = 0;
  std::string err;
  Buffer stage, scratch;           // fetched views (+ seed set); seed distances + records of this device's pairs
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device time of the last share (fetch ... last delivery)
  float last_ms = 0.f;
};

}  // namespace

struct ssrlcv_b200_multi {
  std::vector<Worker*> workers;
  std::string err;
};

namespace {

int fail(Worker& w, int code, const char* what) {
  char buf[256];
  snprintf(buf, sizeof buf, "device %d: %s (%d%s%s)", w.device, what, code, code > 0 ? ": " : "",
           code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "");
  w.err = buf;
  return code;
}

// true if `p` is device memory of `device` (usable in place)
bool resident_on(const void* p, int device) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice && a.device == device;
}

int run_share(Worker& w) {
  const Call& c = *w.call;
  w.last_ms = 0.f;
  if (w.pairs.empty()) return 0;
  cudaError_t e = cudaSetDevice(w.device);
  if (e != cudaSuccess) return fail(w, static_cast<int>(e), "cudaSetDevice");
  cudaStream_t stream = static_cast<cudaStream_t>(ssrlcv_b200_stream(w.ctx));
  if (!w.ev0) { cudaEventCreate(&w.ev0); cudaEventCreate(&w.ev1); }
  cudaEventRecord(w.ev0, stream);
  // ---- the views this device touches (ascending), and which of them are query views
  std::vector<uint32_t> need;
  {
    std::set<uint32_t> s;
    for (uint32_t k : w.pairs) { s.insert(c.pair_query[k]); s.insert(c.pair_target[k]); }
    need.assign(s.begin(), s.end());
  }
  std::vector<int> pos(c.n_views, -1);
  for (size_t i = 0; i < need.size(); ++i) pos[need[i]] = static_cast<int>(i);
  const bool seeded = c.seed != nullptr;
  // ---- fetch: views that do not already live on this device are copied into the staging buffer (peer copy / H2D)
  size_t stage_bytes = 0;
  std::vector<size_t> stage_off(need.size(), 0);
  std::vector<bool> in_place(need.size(), false);
  for (size_t i = 0; i < need.size(); ++i) {
    const uint64_t n = c.n_features[need[i]];
    in_place[i] = n == 0 || resident_on(c.features[need[i]], w.device);
    if (!in_place[i]) { stage_off[i] = stage_bytes; stage_bytes += ((n * kFeatureBytes + 255) / 256) * 256; }
  }
  size_t seed_off = 0;
  const bool seed_in_place = !seeded || c.n_seed == 0 || resident_on(c.seed, w.device);
  if (!seed_in_place) { seed_off = stage_bytes; stage_bytes += ((c.n_seed * kFeatureBytes + 255) / 256) * 256; }
  int rc = w.stage.reserve(stage_bytes ? stage_bytes : 256);
  if (rc) return fail(w, rc, "staging buffer");
  std::vector<const void*> dev(need.size() + (seeded ? 1 : 0));
  std::vector<uint64_t> cnt(dev.size());
  for (size_t i = 0; i < need.size(); ++i) {
    const uint64_t n = c.n_features[need[i]];
    cnt[i] = n;
    if (in_place[i]) { dev[i] = c.features[need[i]]; continue; }
    uint8_t* dst = static_cast<uint8_t*>(w.stage.p) + stage_off[i];
    e = cudaMemcpyAsync(dst, c.features[need[i]], n * kFeatureBytes, cudaMemcpyDefault, stream);
    if (e != cudaSuccess) return fail(w, static_cast<int>(e), "fetching a view");
    dev[i] = dst;
  }
  if (seeded) {
    cnt[need.size()] = c.n_seed;
    if (seed_in_place) dev[need.size()] = c.seed;
    else {
      uint8_t* dst = static_cast<uint8_t*>(w.stage.p) + seed_off;
      e = cudaMemcpyAsync(dst, c.seed, c.n_seed * kFeatureBytes, cudaMemcpyDefault, stream);
      if (e != cudaSuccess) return fail(w, static_cast<int>(e), "fetching the seed set");
      dev[need.size()] = dst;
    }
  }
  // ---- pack everything in one launch
  std::vector<ssrlcv_b200_set*> sets(dev.size(), nullptr);
  rc = ssrlcv_b200_pack_batch(w.ctx, static_cast<uint32_t>(dev.size()), dev.data(), cnt.data(), sets.data());
  if (rc) { w.err = ssrlcv_b200_last_error(w.ctx); return rc; }
  auto release = [&]() { for (ssrlcv_b200_set* s : sets) ssrlcv_b200_set_free(s); };
  // ---- scratch: seed distances of the query views, then the records of every pair (capacity = queries of the pair)
  size_t scratch_bytes = 0;
  std::vector<size_t> seed_dist_off(need.size(), 0);
  std::vector<bool> is_query(need.size(), false);
  if (seeded)
    for (uint32_t k : w.pairs) is_query[pos[c.pair_query[k]]] = true;
  for (size_t i = 0; i < need.size(); ++i)
    if (is_query[i]) { seed_dist_off[i] = scratch_bytes; scratch_bytes += ((c.n_features[need[i]] * sizeof(float) + 255) / 256) * 256; }
  std::vector<size_t> out_off(w.pairs.size());
  for (size_t j = 0; j < w.pairs.size(); ++j) {
    const uint32_t k = w.pairs[j];
    out_off[j] = scratch_bytes;
    scratch_bytes += ((c.n_features[c.pair_query[k]] * record_bytes(c.params[k].out_kind) + 255) / 256) * 256;
  }
  rc = w.scratch.reserve(scratch_bytes ? scratch_bytes : 256);
  if (rc) { release(); return fail(w, rc, "scratch buffer"); }
  uint8_t* scr = static_cast<uint8_t*>(w.scratch.p);
  if (seeded)
    for (size_t i = 0; i < need.size(); ++i) {
      if (!is_query[i]) continue;
      // getSeedDistances once per query view (ref: src/MatchFactory.cu:915-917), enqueue-only
      rc = ssrlcv_b200_seed_distances(w.ctx, sets[i], sets[need.size()], reinterpret_cast<float*>(scr + seed_dist_off[i]));
      if (rc) { w.err = ssrlcv_b200_last_error(w.ctx); release(); return rc; }
    }
  // ---- all of this device's pairs: one batch, one synchronisation
  const size_t m = w.pairs.size();
  std::vector<ssrlcv_b200_match_params> params(m);
  std::vector<const ssrlcv_b200_set*> qs(m), ts(m);
  std::vector<const float*> sd(m, nullptr);
  std::vector<void*> outs(m);
  std::vector<uint64_t> counts(m, 0);
  for (size_t j = 0; j < m; ++j) {
    const uint32_t k = w.pairs[j];
    params[j] = c.params[k];
    qs[j] = sets[pos[c.pair_query[k]]];
    ts[j] = sets[pos[c.pair_target[k]]];
    if (seeded) sd[j] = reinterpret_cast<const float*>(scr + seed_dist_off[pos[c.pair_query[k]]]);
    outs[j] = scr + out_off[j];
  }
  rc = ssrlcv_b200_match_batch(w.ctx, static_cast<uint32_t>(m), params.data(), qs.data(), ts.data(), seeded ? sd.data() : nullptr, outs.data(), counts.data());
  if (rc) { w.err = ssrlcv_b200_last_error(w.ctx); release(); return rc; }
  // ---- the exchange: every pair list straight to the caller's buffer for that pair
  bool any = false;
  for (size_t j = 0; j < m; ++j) {
    const uint32_t k = w.pairs[j];
    c.n_out[k] = counts[j];
    if (counts[j] == 0) continue;
    e = cudaMemcpyAsync(c.out[k], outs[j], counts[j] * record_bytes(c.params[k].out_kind), cudaMemcpyDefault, stream);
    if (e != cudaSuccess) { release(); return fail(w, static_cast<int>(e), "delivering a pair list"); }
    any = true;
  }
  release();                                           // stream-ordered: after everything that reads the sets
  cudaEventRecord(w.ev1, stream);
  (void)any;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return fail(w, static_cast<int>(e), "cudaStreamSynchronize");
  cudaEventElapsedTime(&w.last_ms, w.ev0, w.ev1);
  return 0;
}

void worker_main(Worker* w) {
  cudaSetDevice(w->device);
  std::unique_lock<std::mutex> lk(w->mu);
  for (;;) {
    w->cv.wait(lk, [&] { return w->has_job || w->stop; });
    if (w->stop) return;
    lk.unlock();
    const int rc = run_share(*w);
    lk.lock();
    w->rc = rc;
    w->has_job = false;
    w->idle = true;
    w->cv.notify_all();
  }
}

// Deterministic assignment of pairs to devices: balanced by cost Nq * Nt with view affinity.  The least-loaded device picks
// next (only pairs within a factor two of the largest remaining cost are eligible) and takes the pair that brings the fewest
// views it does not hold yet, then the one whose new views the fewest other devices hold (cliques beat stars: 16 equal
// views over 8 devices need 7-9 views per device instead of all 16).  Same rule as pair_scheduler.py::assign_pairs.
void assign_pairs(const Call& c, size_t n_dev, std::vector<std::vector<uint32_t>>& out) {
  const uint32_t P = c.n_pairs;
  std::vector<unsigned long long> cost(P);
  for (uint32_t k = 0; k < P; ++k) cost[k] = static_cast<unsigned long long>(c.n_features[c.pair_query[k]]) * c.n_features[c.pair_target[k]];
  std::vector<bool> left(P, true);
  uint32_t n_left = P;
  std::vector<unsigned long long> load(n_dev, 0);
  std::vector<std::vector<bool>> held(n_dev, std::vector<bool>(c.n_views, false));
  std::vector<uint32_t> holders(c.n_views, 0);
  out.assign(n_dev, {});
  while (n_left) {
    size_t r = 0;
    for (size_t d = 1; d < n_dev; ++d) if (load[d] < load[r]) r = d;
    unsigned long long cmax = 0;
    for (uint32_t k = 0; k < P; ++k) if (left[k] && cost[k] > cmax) cmax = cost[k];
    bool have = false;
    unsigned long long best_new = 0, best_hold = 0, best_cost = 0;
    uint32_t best = 0;
    for (uint32_t k = 0; k < P; ++k) {
      if (!left[k] || cost[k] * 2 < cmax) continue;
      const uint32_t q = c.pair_query[k], t = c.pair_target[k];
      const bool nq = !held[r][q], nt = !held[r][t];
      const unsigned long long n_new = (nq ? 1 : 0) + ((nt && t != q) ? 1 : 0);
      const unsigned long long n_hold = (nq ? holders[q] : 0) + (nt ? holders[t] : 0);
      // key (new views, holders of the new views, -cost, index), smallest wins
      const bool better = !have || n_new < best_new || (n_new == best_new && (n_hold < best_hold || (n_hold == best_hold && cost[k] > best_cost)));
      if (better) { have = true; best = k; best_new = n_new; best_hold = n_hold; best_cost = cost[k]; }
    }
    left[best] = false; --n_left;
    out[r].push_back(best);
    load[r] += cost[best];
    for (uint32_t v : {c.pair_query[best], c.pair_target[best]})
      if (!held[r][v]) { held[r][v] = true; holders[v] += 1; }
  }
  for (auto& v : out) std::sort(v.begin(), v.end());
}

}  // namespace

extern "C" {

int ssrlcv_b200_multi_create(ssrlcv_b200_multi** out, int n_devices, const int* devices) {
  if (!out) return SSRLCV_B200_EINVAL;
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) return SSRLCV_B200_ENODEVICE;
  std::vector<int> ids;
  if (devices) ids.assign(devices, devices + (n_devices > 0 ? n_devices : 0));
  else
    for (int d = 0; d < visible && (n_devices <= 0 || static_cast<int>(ids.size()) < n_devices); ++d) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ids.push_back(d);
    }
  if (ids.empty()) return SSRLCV_B200_ENODEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  ssrlcv_b200_multi* m = new (std::nothrow) ssrlcv_b200_multi();
  if (!m) return SSRLCV_B200_ENOMEM;
  for (int d : ids) {
    Worker* w = new Worker();
    w->device = d;
    int rc = ssrlcv_b200_create(&w->ctx, d, nullptr);
    if (rc) { delete w; ssrlcv_b200_multi_destroy(m); cudaSetDevice(prev); return rc; }
    m->workers.push_back(w);
  }
  // peer access both ways, so that views and pair lists travel over NVLink directly (already enabled / unsupported: ignored)
  for (int a : ids)
    for (int b : ids) {
      if (a == b) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
        cudaSetDevice(a);
        cudaDeviceEnablePeerAccess(b, 0);
        cudaGetLastError();
      }
    }
  cudaSetDevice(prev);
  for (Worker* w : m->workers) w->th = std::thread(worker_main, w);
  *out = m;
  return 0;
}

void ssrlcv_b200_multi_destroy(ssrlcv_b200_multi* m) {
  if (!m) return;
  int prev = 0;
  cudaGetDevice(&prev);
  for (Worker* w : m->workers) {
    if (w->th.joinable()) {
      { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; }
      w->cv.notify_all();
      w->th.join();
    }
    cudaSetDevice(w->device);
    if (w->ev0) { cudaEventDestroy(w->ev0); cudaEventDestroy(w->ev1); }
    if (w->stage.p) cudaFree(w->stage.p);
    if (w->scratch.p) cudaFree(w->scratch.p);
    ssrlcv_b200_destroy(w->ctx);
    delete w;
  }
  cudaSetDevice(prev);
  delete m;
}

int ssrlcv_b200_multi_device_count(const ssrlcv_b200_multi* m) { return m ? static_cast<int>(m->workers.size()) : 0; }
const char* ssrlcv_b200_multi_last_error(const ssrlcv_b200_multi* m) { return m ? m->err.c_str() : "no multi-GPU context"; }
double ssrlcv_b200_multi_device_ms(const ssrlcv_b200_multi* m, int index) {
  return (m && index >= 0 && index < static_cast<int>(m->workers.size())) ? static_cast<double>(m->workers[index]->last_ms) : 0.0;
}

int ssrlcv_b200_multi_assignment(const ssrlcv_b200_multi* m, uint32_t n_views, const uint64_t* n_features, uint32_t n_pairs,
                                 const uint32_t* pair_query, const uint32_t* pair_target, int32_t* device_of_pair) {
  if (!m || (n_pairs && (!n_features || !pair_query || !pair_target || !device_of_pair))) return SSRLCV_B200_EINVAL;
  Call c;
  c.n_views = n_views; c.n_features = n_features; c.n_pairs = n_pairs; c.pair_query = pair_query; c.pair_target = pair_target;
  for (uint32_t k = 0; k < n_pairs; ++k)
    if (pair_query[k] >= n_views || pair_target[k] >= n_views) return SSRLCV_B200_EINVAL;
  std::vector<std::vector<uint32_t>> share;
  assign_pairs(c, m->workers.size(), share);
  for (size_t d = 0; d < share.size(); ++d)
    for (uint32_t k : share[d]) device_of_pair[k] = m->workers[d]->device;
  return 0;
}

int ssrlcv_b200_match_pairs_multi_gpu(ssrlcv_b200_multi* m, uint32_t n_views, const void* const* features, const uint64_t* n_features,
                                      const void* seed_features, uint64_t n_seed, uint32_t n_pairs, const uint32_t* pair_query,
                                      const uint32_t* pair_target, const ssrlcv_b200_match_params* params, void* const* out, uint64_t* n_out) {
  if (!m) return SSRLCV_B200_EINVAL;
  if (n_pairs == 0) return 0;
  if (!features || !n_features || !pair_query || !pair_target || !params || !out || !n_out) { m->err = "match_pairs_multi_gpu: bad arguments"; return SSRLCV_B200_EINVAL; }
  for (uint32_t k = 0; k < n_pairs; ++k) {
    if (pair_query[k] >= n_views || pair_target[k] >= n_views || record_bytes(params[k].out_kind) == 0 ||
        (n_features[pair_query[k]] && !out[k])) { m->err = "match_pairs_multi_gpu: bad pair"; return SSRLCV_B200_EINVAL; }
    n_out[k] = 0;
  }
  for (uint32_t v = 0; v < n_views; ++v)
    if ((n_features[v] && !features[v]) || n_features[v] > 0x7fffffffull) { m->err = "match_pairs_multi_gpu: bad view"; return SSRLCV_B200_EINVAL; }
  Call c;
  c.n_views = n_views; c.features = features; c.n_features = n_features; c.seed = seed_features; c.n_seed = n_seed;
  c.n_pairs = n_pairs; c.pair_query = pair_query; c.pair_target = pair_target; c.params = params; c.out = out; c.n_out = n_out;
  std::vector<std::vector<uint32_t>> share;
  assign_pairs(c, m->workers.size(), share);
  for (size_t d = 0; d < m->workers.size(); ++d) {
    Worker* w = m->workers[d];
    std::lock_guard<std::mutex> lk(w->mu);
    w->call = &c; w->pairs = share[d]; w->rc = 0; w->err.clear();
    w->has_job = true; w->idle = false;
    w->cv.notify_all();
  }
  int rc = 0;
  for (Worker* w : m->workers) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->idle; });
    if (w->rc && !rc) { rc = w->rc; m->err = w->err; }
  }
  return rc;
}

}  // extern "C"
