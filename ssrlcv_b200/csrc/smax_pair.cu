// smax_pair.cu -- the tensor-core matcher for large unmasked searches: TWO query tiles per work unit.
//
// Same arithmetic, same bookkeeping, same hand-off and same keys as smax_tc.cu (see its header); what changes is how a CTA walks
// the score matrix.  Measured on B200 (profiles/r2t_diag_ab.txt) the 65 536 x 65 536 launch of smax_tc.cu sits on the tensor
// pipe's issue rate in CYCLES but runs at ~1.6 GHz of 1.965: the SM clock is power-limited, so the launch pays in clock for
// every joule it spends besides the MMAs.  With one 128-row query tile per unit every target tile (32 KB + 8 KB of fold digits)
// crosses the crossbar and is written to shared memory once per 128 x 256 scores: 5.4 GB per launch, ~15 TB/s.  Loading only every
// other target tile (wrong results, timing only) made that launch 5.3 % faster at the same cycle count.  Here a unit is a PAIR of
// query tiles (256 rows): each staged target tile is contracted with both A tiles, into the two TMEM accumulators -- half the L2
// reads, half the TMA shared-memory writes, half the TMA issue rate, for the same MMAs.
//
// Pipeline per CTA (persistent over (query tile pair, target range) units; 384 threads):
//   warp 0  TMA producer : the unit's two A tiles (2 x 128 x 128 B); per stage a B tile (256 x 128 B) + its fold digits
//                          (256 x 32 B), 3-stage ring; the no-fold variant stages W and the window bounds instead of the digits
//   warp 1  MMA issuer   : per target tile, for p = 0, 1: 4 (+1 fold) tcgen05.mma M128 N256 K32 of A_p against the stage
//                          into accumulator p, commit -> t_full[p]; after p = 1 the stage goes back to the producer
//   warps 2-3            : TMEM allocation / the fold's constant operand, then resolvers (128 rows each) of the previous unit
//   warps 4-11 epilogue  : as in smax_tc.cu all eight drain EVERY accumulator (two warps per TMEM lane quarter, one per 128-column
//                          half: one tcgen05.ld 32x32b.x128, accumulator handed back at once, four max trees, one warp-uniform
//                          test), so a thread serves two query rows in turn, one per accumulator.  Only each row's running best
//                          lives in a register; the window records behind it stay in shared memory (where the resolvers read
//                          them when the unit ends) and are touched in the rare path only.
// Two other walks were built and measured slower (profiles/r2_kernel_experiments.txt): one epilogue warp set per accumulator
// (a single warp drains TMEM at half the sub-partition's rate: 256 columns take ~500 cycles, the accumulator comes back too late),
// and 64-column software pipelining of the epilogue (a tcgen05.ld costs ~250 cycles whatever its width under MMA load).
#include <atomic>

#include "smax_common.cuh"

namespace b200 {
namespace pairk {

constexpr int kStages = 3;                             // a stage serves two MMA groups: three of them buffer as long as smax_tc.cu's four (four stages + one A pair measured the same)
constexpr int kASlots = 2;                             // the next unit's A pair loads while this unit's last tiles are contracted
constexpr int kColStages = 6;
constexpr uint32_t kPairRows = 2 * kBM;               // 256 query rows per unit
constexpr uint32_t kABytes = kBM * kDescBytes;        // 16384
constexpr uint32_t kAPairBytes = 2 * kABytes;
constexpr uint32_t kBBytes = kBN * kDescBytes;        // 32768
constexpr uint32_t kBExtBytes = kBN * kExtBytes;      // 8192
constexpr uint32_t kAExtBytes = kBM * kExtBytes;      // 4096
constexpr uint32_t kColWBytes = kBN * 4;
constexpr uint32_t kColWmaxBytes = (kBN / kChunk) * 4;
constexpr uint32_t kTmemCols = 512;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;        // 384

struct Smem {
  static constexpr uint32_t a = 0;                                       // kASlots x 2 tiles x 16 KB
  static constexpr uint32_t b = a + kASlots * kAPairBytes;               // kStages x 32 KB
  static constexpr uint32_t bext = b + kStages * kBBytes;                // kStages x 8 KB (fold digits)
  static constexpr uint32_t col_w = bext;                                // no-fold: kColStages x 1 KB of W ...
  static constexpr uint32_t col_wmax = col_w + kColStages * kColWBytes;  // ... and kColStages x 32 B of window bounds, in the digit region
  static constexpr uint32_t aext = bext + kStages * kBExtBytes;          // 4 KB
  static constexpr uint32_t bars = aext + kAExtBytes;
  static constexpr uint32_t n_bars = 2 * kStages + 4 + 4 + 2 * kColStages + 4;
  static constexpr uint32_t tmem_ptr = bars + n_bars * 8;
  static constexpr uint32_t xch = tmem_ptr + 16;                         // 2 (unit parity) x 2 (column half) x 256 rows x 24 B
  static constexpr uint32_t total = xch + 2 * 2 * kPairRows * 24;
};
static_assert(kColStages * (kColWBytes + kColWmaxBytes) <= kStages * kBExtBytes, "the no-fold rings live in the digit region");
constexpr uint32_t kSmemBytes = Smem::total + 1024;
static_assert(kSmemBytes <= 232448, "227 KB of shared memory per CTA");

// a row's window state in shared memory, by shared-window address (24 bytes, 8-byte aligned)
__device__ __forceinline__ void ld_state(uint32_t addr, WinState& s) {
  uint32_t a, b, c, d, e, f;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(c), "=r"(d) : "r"(addr + 8) : "memory");
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e), "=r"(f) : "r"(addr + 16) : "memory");
  s.best = static_cast<int32_t>(a); s.w1 = b; s.m1 = c; s.w2 = d; s.m2 = e; s.ovf = f;
}
__device__ __forceinline__ void st_state(uint32_t addr, const WinState& s) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(static_cast<uint32_t>(s.best)), "r"(s.w1) : "memory");
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + 8), "r"(s.m1), "r"(s.w2) : "memory");
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + 16), "r"(s.m2), "r"(s.ovf) : "memory");
}

// kSeedOnly (fold variant only): the search behind getSeedDistances (ref: src/MatchFactory.cu:349-394) wants each query's smallest
// d2 and nothing else -- no index, no tie rule, no threshold.  An uncapped search is bookkeeping-bound (a running best keeps
// improving, profiles/r2_fixture_regimes.txt), but a minimum needs no bookkeeping: with |b|^2 = 2h + p the score S = a.b + C - h
// loses the parity bit p, and d2 = |a|^2 + 2C - (2S - p) exactly.  The epilogue therefore maximises T = 2S - p over ALL columns
// branch-free (one add per score in front of the max trees) and keeps one register per row: no window records, no rare path, no
// resolvers.  The resolver warps, idle in this mode, turn each target tile's norms into -p and stage them in shared memory a few
// tiles ahead of the epilogue (the c_full / c_empty ring; the buffers live in the window-state region this mode does not use).
constexpr uint32_t kNegParBytes = kBN * 4;              // one int32 (0 or -1) per target of a tile
static_assert(kColStages * kNegParBytes + 2 * 2 * kPairRows * 4 <= 2 * 2 * kPairRows * 24, "seed-only buffers fit the window-state region");

template <bool kFold, bool kColMask, bool kTiming, bool kSeedOnly = false>
__global__ void __launch_bounds__(kThreads, 1)
smax_pair_kernel(const __grid_constant__ SmaxBatch batch) {
  static_assert(!kSeedOnly || (kFold && !kColMask && !kTiming), "seed-only: the fold variant, no bookkeeping regime, no instrumentation");
  constexpr int32_t kWOffset = kFold ? kFoldC : kNoFoldC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;                         // [kStages]    TMA -> MMA
  uint64_t* empty = bars + kStages;              // [kStages]    MMA -> TMA
  uint64_t* a_full = bars + 2 * kStages;         // [2]
  uint64_t* a_empty = a_full + 2;                // [2]
  uint64_t* t_full = a_empty + 2;                // [2]          MMA -> epilogue (accumulator p = query tile p of the pair ready)
  uint64_t* t_empty = t_full + 2;                // [2]          epilogue -> MMA (accumulator p copied out of TMEM)
  uint64_t* c_full = t_empty + 2;                // [kColStages] TMA -> epilogue (W / window bounds, no-fold)
  uint64_t* c_empty = c_full + kColStages;       // [kColStages]
  uint64_t* x_full = c_empty + kColStages;       // [2] epilogue -> resolver warps (window states of a finished unit)
  uint64_t* x_empty = x_full + 2;                // [2] resolver -> epilogue
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kEpiWarps); }
    for (int i = 0; i < kColStages; ++i) { mbar_init(&c_full[i], kSeedOnly ? 2 : 1); mbar_init(&c_empty[i], kEpiWarps); }   // seed-only: filled by the two resolver warps
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], kEpiWarps); mbar_init(&x_empty[i], 2); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc<kTmemCols>(tmem_ptr); tmem_relinquish(); }
  if (warp == 3 && kFold) {
    // constant A-side operand of the fold MMA: every row = [1, 255 x 31] in the SWIZZLE_32B K-major layout (smax_tc.cu)
    for (int r = lane; r < kBM; r += 32) {
      const uint32_t sw = (r >> 2) & 1;
      uint4 c0 = make_uint4(0xFFFFFF01u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      uint4 c1 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      uint4* row = reinterpret_cast<uint4*>(smem + Smem::aext + r * kExtBytes);
      row[0 ^ sw] = c0;
      row[1 ^ sw] = c1;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // everything above overlapped the tail of the previous kernel of the stream (the pack launch); its output is read from here on
  pdl_launch_dependents();
  pdl_wait();
  if (blockIdx.x == 0)                                 // tie counters of the other bank (the next call's): see exact.cu.  After the
    for (uint32_t i = threadIdx.x; i < batch.clear_n; i += kThreads) batch.clear[i] = 0;   // wait: the previous call's tie pass reads that bank

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, aslot = 0, aphase = 0, cs = 0, cphase = 0;
      long long tp_wait = 0;
      uint32_t acquired = 0xFFFFFFFFu;
      UnitIter u(batch, blockIdx.x);
      while (u.next(batch, blockIdx.x, kFold)) {
        const CUtensorMap* map_q = &u.J->map_q;
        const CUtensorMap* map_t = &u.J->map_t;
        const CUtensorMap* map_text = &u.J->map_text;
        if (u.job != acquired && batch.jobs != nullptr) {          // first use of this job's tensor maps (global memory) by this CTA
          tma_acquire_desc(map_q); tma_acquire_desc(map_t);
          if (kFold) tma_acquire_desc(map_text);
          acquired = u.job;
        }
        const int32_t* t_wcol = u.J->a.t_wcol;
        const int32_t* t_wmax = u.J->a.t_wmax;
        mbar_wait(&a_empty[aslot], aphase ^ 1);
        mbar_arrive_expect_tx(&a_full[aslot], kAPairBytes);
        tma_load_2d(smem + Smem::a + aslot * kAPairBytes, map_q, &a_full[aslot], 0, static_cast<int32_t>(u.qt * kPairRows));
        tma_load_2d(smem + Smem::a + aslot * kAPairBytes + kABytes, map_q, &a_full[aslot], 0, static_cast<int32_t>(u.qt * kPairRows + kBM));
        for (uint32_t j = u.j0; j < u.j1; ++j) {
          long long p0 = 0;
          if (kTiming) p0 = clock64();
          mbar_wait(&empty[stage], phase ^ 1);
          if (kTiming) tp_wait += clock64() - p0;
          mbar_arrive_expect_tx(&full[stage], kBBytes + (kFold ? kBExtBytes : 0));
          tma_load_2d(smem + Smem::b + stage * kBBytes, map_t, &full[stage], 0, static_cast<int32_t>(j * kBN));
          if (kFold) tma_load_2d(smem + Smem::bext + stage * kBExtBytes, map_text, &full[stage], 0, static_cast<int32_t>(j * kBN));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (!kFold && !(kDiagBuild && (batch.diag & 32u))) {   // W of the tile's 256 targets and the largest W of each of its 8 windows (pack.cu wcol / wmax)
            mbar_wait(&c_empty[cs], cphase ^ 1);
            mbar_arrive_expect_tx(&c_full[cs], kColWBytes + kColWmaxBytes);
            bulk_load_1d(smem + Smem::col_w + cs * kColWBytes, t_wcol + static_cast<uint64_t>(j) * kBN, kColWBytes, &c_full[cs]);
            bulk_load_1d(smem + Smem::col_wmax + cs * kColWmaxBytes, t_wmax + static_cast<uint64_t>(j) * (kBN / kChunk), kColWmaxBytes, &c_full[cs]);
            if (++cs == kColStages) { cs = 0; cphase ^= 1; }
          }
        }
        if (++aslot == kASlots) { aslot = 0; aphase ^= 1; }
      }
      if (kTiming) batch.timing[blockIdx.x * 16 + 3] = tp_wait;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(kBM, kBN, 0, 0);
      const uint64_t aext_desc = make_smem_desc(smem_u32(smem + Smem::aext), 256, kSwizzle32B);
      uint32_t stage = 0, phase = 0, aslot = 0, aphase = 0, accphase = 0;
      long long tm_wait_acc = 0, tm_wait_full = 0, tm_total0 = 0;
      if (kTiming) tm_total0 = clock64();
      UnitIter u(batch, blockIdx.x);
      while (u.next(batch, blockIdx.x, kFold)) {
        mbar_wait(&a_full[aslot], aphase);
        const uint64_t a_desc0 = make_smem_desc(smem_u32(smem + Smem::a + aslot * kAPairBytes), 1024, kSwizzle128B);
        const uint64_t a_desc1 = make_smem_desc(smem_u32(smem + Smem::a + aslot * kAPairBytes + kABytes), 1024, kSwizzle128B);
        for (uint32_t j = u.j0; j < u.j1; ++j) {
          long long c0 = 0, c1 = 0;
          if (kTiming) c0 = clock64();
          mbar_wait(&full[stage], phase);
          if (kTiming) { c1 = clock64(); tm_wait_full += c1 - c0; }
          const uint64_t b_desc = make_smem_desc(smem_u32(smem + Smem::b + stage * kBBytes), 1024, kSwizzle128B);
          const uint64_t bext_desc = make_smem_desc(smem_u32(smem + Smem::bext + stage * kBExtBytes), 256, kSwizzle32B);
#pragma unroll
          for (uint32_t p = 0; p < 2; ++p) {
            const uint32_t d_tmem = tmem_base + p * kBN;
            const uint64_t a_desc = p ? a_desc1 : a_desc0;
            long long w0 = 0;
            if (kTiming) w0 = clock64();
            mbar_wait(&t_empty[p], accphase ^ 1);
            if (kTiming) tm_wait_acc += clock64() - w0;
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)   // +32 bytes along K inside the 128-byte swizzle atom = +2 in the address field
              mma_i8_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k);
            if (kFold && !(kDiagBuild && (batch.diag & 2u))) mma_i8_ss(d_tmem, aext_desc, bext_desc, idesc, 1);
            mma_commit(&t_full[p]);
          }
          mma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          accphase ^= 1;
        }
        mma_commit(&a_empty[aslot]);
        if (++aslot == kASlots) { aslot = 0; aphase ^= 1; }
      }
      if (kTiming) {
        batch.timing[blockIdx.x * 16 + 0] = tm_wait_acc;
        batch.timing[blockIdx.x * 16 + 1] = tm_wait_full;
        batch.timing[blockIdx.x * 16 + 2] = clock64() - tm_total0;
      }
    }
  } else if (warp == 2 || warp == 3) {
    // =============================== resolver ===============================
    // Turns the window states of the PREVIOUS unit into keys while the pipe and the epilogue warps work on the next one
    // (128 rows per warp); the CTA's last unit is resolved by the epilogue warps themselves.
    const WinState* xch = reinterpret_cast<const WinState*>(smem + Smem::xch);
    const uint32_t rw = warp - 2;
    uint32_t uparity = 0, uiter = 0;
    long long tr_wait = 0, tr_work = 0;
    UnitIter u(batch, blockIdx.x);
    if (kSeedOnly) {
      // -p of every target of every tile of this CTA's units, kColStages tiles ahead of the epilogue: warp rw takes columns
      // [128 rw, +128) of the tile, lane l four of them
      uint32_t cs = 0, cphase = 0;
      while (u.next(batch, blockIdx.x, kFold)) {
        const int32_t* t_norm = u.J->a.t_norm;
        for (uint32_t j = u.j0; j < u.j1; ++j) {
          const int4 n4 = __ldg(reinterpret_cast<const int4*>(t_norm + static_cast<uint64_t>(j) * kBN + rw * 128) + lane);
          mbar_wait(&c_empty[cs], cphase ^ 1);
          const uint32_t dst = smem_u32(smem + Smem::xch + cs * kNegParBytes) + (rw * 128 + lane * 4) * 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(-(n4.x & 1)), "r"(-(n4.y & 1)), "r"(-(n4.z & 1)), "r"(-(n4.w & 1)) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_full[cs]);
          if (++cs == kColStages) { cs = 0; cphase ^= 1; }
        }
      }
    } else
    while (u.next(batch, blockIdx.x, kFold)) {
      if (u.last) break;
      long long w0 = 0, w1 = 0;
      if (kTiming) w0 = clock64();
      mbar_wait(&x_full[uparity], (uiter >> 1) & 1);
      if (kTiming) w1 = clock64();
      const uint32_t row_base = u.qt * kPairRows;
      const WinState* xa = xch + (uparity * 2 + 0) * kPairRows;
      const WinState* xb = xch + (uparity * 2 + 1) * kPairRows;
      if (kColMask) {
#pragma unroll
        for (uint32_t g = 0; g < 4; ++g) {
          const uint32_t r = rw * 128 + g * 32 + lane;
          resolve_row(u.J->a, xa[r], xb[r], row_base, u.slot, r, kWOffset);
        }
      } else {
        resolve_rows_windows<SSRLCV_B200_MASK_NONE, 2>(u.J->a, xa, xb, row_base, u.slot, rw * 128, 128, lane);
      }
      fill_unused_slots(u.J->a, row_base, u.slot, u.fill_end, rw * 128, 128, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[uparity]);
      uparity ^= 1;
      ++uiter;
      if (kTiming) { const long long w2 = clock64(); tr_wait += w1 - w0; tr_work += w2 - w1; }
    }
    if (kTiming && lane == 0 && rw == 0) { batch.timing[blockIdx.x * 16 + 12] = tr_wait; batch.timing[blockIdx.x * 16 + 13] = tr_work; }
  } else {
    // =============================== epilogue ===============================
    const uint32_t quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const uint32_t half = (warp - 4) >> 2;             // which 128 of the 256 accumulator columns
    const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + half * 128;
    const uint32_t my_r = quarter * 32 + lane;         // this thread's row of either query tile
    WinState* xch = reinterpret_cast<WinState*>(smem + Smem::xch);
    uint32_t it = 0, uiter = 0;                       // it: target tiles consumed so far -> phases of the accumulators and the column ring
    long long te_wait = 0, te_work = 0, te_resolve = 0, te_total0 = 0, te_cwait = 0, te_rare = 0, te_rare_n = 0;
    if (kTiming) te_total0 = clock64();
    UnitIter u(batch, blockIdx.x);
    if (kSeedOnly) {
      int32_t* sbest = reinterpret_cast<int32_t*>(smem + Smem::xch + kColStages * kNegParBytes);   // [unit parity][column half][256 rows]
      while (u.next(batch, blockIdx.x, kFold)) {
        const SmaxArgs& args = u.J->a;
        const uint32_t row_base = u.qt * kPairRows;
        int32_t best[2] = {0, 0};                                  // T >= 3 for every real target; padding columns give 0
        for (uint32_t j = u.j0; j < u.j1; ++j) {
          const uint32_t accphase = it & 1, cs = it % kColStages, cphase = (it / kColStages) & 1;
          ++it;
          mbar_wait(&c_full[cs], cphase);
          const uint32_t s_np = smem_u32(smem + Smem::xch + cs * kNegParBytes) + half * (128 * 4);
#pragma unroll
          for (uint32_t p = 0; p < 2; ++p) {
            mbar_wait(&t_full[p], accphase);
            tc_fence_after();
            uint32_t v[4][32];
            tmem_ld_32x128(taddr0 + p * kBN, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[p]);
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
#pragma unroll
              for (uint32_t e = 0; e < 32; e += 4) {
                const int4 np4 = ld_shared_v4(s_np + (k * 32 + e) * 4);
                v[k][e] = v[k][e] + v[k][e] + static_cast<uint32_t>(np4.x); v[k][e + 1] = v[k][e + 1] + v[k][e + 1] + static_cast<uint32_t>(np4.y);
                v[k][e + 2] = v[k][e + 2] + v[k][e + 2] + static_cast<uint32_t>(np4.z); v[k][e + 3] = v[k][e + 3] + v[k][e + 3] + static_cast<uint32_t>(np4.w);
              }
            }
            const int32_t m = max(max(max32(v[0]), max32(v[1])), max(max32(v[2]), max32(v[3])));
            best[p] = max(best[p], m);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_empty[cs]);
        }
        // the two column halves of a row meet in shared memory; half 0 writes the unit's key: d2 = |a|^2 + 2C - T, no index
        const uint32_t uparity = uiter & 1;
        int32_t* mine = sbest + (uparity * 2 + half) * kPairRows;
        mine[my_r] = best[0]; mine[kBM + my_r] = best[1];
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (half == 0) {
          const int32_t* other = sbest + (uparity * 2 + 1) * kPairRows;
#pragma unroll
          for (uint32_t p = 0; p < 2; ++p) {
            const uint32_t row = row_base + p * kBM + my_r;                 // < nq_pad: rows beyond nq land in the padding of the key array
            const int32_t T = max(best[p], other[p * kBM + my_r]);
            const uint32_t d2 = static_cast<uint32_t>(__ldg(args.q_norm + row) + 2 * kFoldC - T);
            args.keys[static_cast<uint64_t>(u.slot) * args.nq_pad + key_row(args.q_orig, args.nq, row)] = T > 0 ? make_key(d2, 0) : kKeyNone;
          }
        }
        fill_unused_slots(args, row_base, u.slot, u.fill_end, (warp - 4) * 32, 32, lane);
        ++uiter;                                                   // (the next unit's halves meet in the other buffer)
      }
    } else
    while (u.next(batch, blockIdx.x, kFold)) {
      const SmaxArgs& args = u.J->a;                               // device memory: read what the unit needs up front
      const uint32_t row_base = u.qt * kPairRows;
      const uint32_t uparity = uiter & 1;
      // this unit's window states (one per row and column half) live in shared memory; the buffer was last read by the
      // resolvers of the unit before the previous one
      mbar_wait(&x_empty[uparity], ((uiter >> 1) & 1) ^ 1);
      WinState* mine = xch + (uparity * 2 + half) * kPairRows + my_r;   // row of query tile p: mine[p * kBM]
      const uint32_t mine_addr = smem_u32(mine);
      int32_t best[2];
#pragma unroll
      for (uint32_t p = 0; p < 2; ++p) {
        const uint32_t row = row_base + p * kBM + my_r;
        const uint32_t row_c = min(row, args.nq_pad - 1);
        // lowest score that can still pass the thresholds (smax_tc.cu)
        int32_t floor_s = 2;
        int32_t d2_cap = args.d2_cap;
        if (args.q_seed != nullptr && row < args.nq) {
          const float c = __fmul_rn(__fmul_rn(args.seed_ratio, args.q_seed[row]), 1.00001f);
          if (c < 8.4e6f) d2_cap = min(d2_cap, static_cast<int32_t>(c) + 1);
        }
        if (d2_cap != 0x7fffffff) {
          const int32_t qn = args.q_norm[row_c];
          floor_s = max(2, kWOffset + ((qn - d2_cap + 1) >> 1));
        }
        WinState st;
        st.best = floor_s - 1; st.w1 = kNoWin; st.w2 = kNoWin; st.m1 = 0; st.m2 = 0; st.ovf = 0;
        mine[p * kBM] = st;
        best[p] = st.best;
      }
      for (uint32_t j = u.j0; j < u.j1; ++j) {
        const uint32_t accphase = it & 1, cs = it % kColStages, cphase = (it / kColStages) & 1;
        ++it;
        int32_t wb[4] = {0, 0, 0, 0};
        if (!kFold && !(kDiagBuild && (batch.diag & 32u))) {
          long long c0 = 0;
          if (kTiming) c0 = clock64();
          mbar_wait(&c_full[cs], cphase);
          const int4 b4 = ld_shared_v4(smem_u32(smem + Smem::col_wmax + cs * kColWmaxBytes + half * 16));
          wb[0] = b4.x; wb[1] = b4.y; wb[2] = b4.z; wb[3] = b4.w;
          if (kTiming) te_cwait += clock64() - c0;
        }
        const uint32_t s_w = smem_u32(smem + Smem::col_w + cs * kColWBytes) + half * (128 * 4);
#pragma unroll
        for (uint32_t p = 0; p < 2; ++p) {
          long long e0 = 0, e1 = 0;
          if (kTiming) e0 = clock64();
          mbar_wait(&t_full[p], accphase);
          if (kTiming) e1 = clock64();
          tc_fence_after();
          // this warp's 32 x 128 slice of accumulator p into registers, and the accumulator goes straight back to the MMA warp
          uint32_t v[4][32];
          tmem_ld_32x128(taddr0 + p * kBN, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[p]);
          int32_t wm[4];
          const int32_t mt = window_maxima<4, kFold>(v, wb, wm, batch.diag);
          if (__any_sync(0xffffffffu, mt >= best[p])) {            // thresholded search: next to never
            // The warp that gets here reaches its next accumulator late, and every tile needs all eight warps: these ~180 entries
            // per CTA cost the 65 536 x 65 536 launch 8 % (profiles/r2_kernel_experiments.txt), so the path is kept short.
            long long r0 = 0;
            if (kTiming) r0 = clock64();
            const uint32_t st_addr = mine_addr + p * (kBM * 24);
            WinState st;
            if (kFold && kColMask) {
              // which windows matter is decided on the register copy of the running best (== st.best), in ONE warp reduction, so
              // the state's load overlaps the equality masks; a window that stops mattering once an earlier one has raised the
              // best is a no-op in win_update
              uint32_t live = 0;
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) live |= (wm[k] >= best[p]) ? (1u << k) : 0u;
              live = __reduce_or_sync(0xffffffffu, live);
              ld_state(st_addr, st);
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k)
                if ((live >> k) & 1u) win_update(st, wm[k], j * 8 + half * 4 + k, eq_mask32(v[k], wm[k]));
            } else {
              ld_state(st_addr, st);
              window_bookkeeping<4, kFold, kColMask>(v, wm, st, j * 8 + half * 4, s_w, /* runner-up searches take smax_tc.cu */ false, args, 0u);
            }
            st_state(st_addr, st);
            best[p] = st.best;
            if (kTiming) { te_rare += clock64() - r0; ++te_rare_n; }
          }
          if (kTiming) { const long long e2 = clock64(); te_wait += e1 - e0; te_work += e2 - e1; }
        }
        if (!kFold && !(kDiagBuild && (batch.diag & 32u))) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_empty[cs]);
        }
      }
      long long r0 = 0;
      if (kTiming) r0 = clock64();
      // ---------------- the unit's window states go to the resolver warps ----------------
      if (!u.last) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_full[uparity]);
      } else {
        // last unit of this CTA: nothing left to overlap with, the epilogue warps turn it into keys themselves (32 rows each)
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        const WinState* xa = xch + (uparity * 2 + 0) * kPairRows;
        const WinState* xb = xch + (uparity * 2 + 1) * kPairRows;
        const uint32_t r = (warp - 4) * 32 + lane;
        if (!kColMask) resolve_rows_windows<SSRLCV_B200_MASK_NONE, 2>(args, xa, xb, row_base, u.slot, (warp - 4) * 32, 32, lane);
        else resolve_row(args, xa[r], xb[r], row_base, u.slot, r, kWOffset);
        fill_unused_slots(args, row_base, u.slot, u.fill_end, (warp - 4) * 32, 32, lane);
      }
      ++uiter;
      if (kTiming) te_resolve += clock64() - r0;
    }
    if (kTiming && lane == 0) {
      const uint32_t e = warp - 4;
      if (e == 0 || e == 5) {
        const uint32_t o = blockIdx.x * 16 + (e == 0 ? 4 : 8);
        batch.timing[o + 0] = te_wait; batch.timing[o + 1] = te_work; batch.timing[o + 2] = te_resolve; batch.timing[o + 3] = clock64() - te_total0;
        if (e == 0) { batch.timing[blockIdx.x * 16 + 14] = te_rare_n; batch.timing[blockIdx.x * 16 + 15] = kFold ? te_rare : te_cwait; }   // fold: cycles inside the rare path
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

template <bool kFold, bool kColMask, bool kTiming, bool kSeedOnly = false>
static cudaError_t launch_one(cudaStream_t st, const SmaxBatch& b) {
  auto kern = smax_pair_kernel<kFold, kColMask, kTiming, kSeedOnly>;
  // the dynamic shared-memory opt-in is a per-device attribute of the function (one context per device in multi-GPU jobs)
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
  }
  return launch_pdl(kern, dim3(b.n_ctas), dim3(kThreads), kSmemBytes, st, b);
}

}  // namespace pairk

// one launch over all jobs of `b` (unmasked, same fold variant and bookkeeping regime, SmaxArgs::q_tile_rows == 256)
cudaError_t launch_smax_pair(cudaStream_t st, const SmaxBatch& b, bool fold, bool colmask, bool seed_only) {
  using namespace pairk;
  if (seed_only) return launch_one<true, false, false, true>(st, b);   // (the host asks for it with the fold variant only)
  if (b.timing) {   // diagnostics build (tools/kerneltiming.py)
    if (fold) return colmask ? launch_one<true, true, true>(st, b) : launch_one<true, false, true>(st, b);
    return colmask ? launch_one<false, true, true>(st, b) : launch_one<false, false, true>(st, b);
  }
  if (fold) return colmask ? launch_one<true, true, false>(st, b) : launch_one<true, false, false>(st, b);
  return colmask ? launch_one<false, true, false>(st, b) : launch_one<false, false, false>(st, b);
}

}  // namespace b200
