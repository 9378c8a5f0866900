// smax_duo.cu -- the tensor-core matcher for large unmasked searches on a CTA PAIR (thread-block cluster of two SMs of one TPC,
// tcgen05 cta_group::2): FOUR query tiles per work unit, two per CTA, every staged target tile split between the two SMs.
//
// Same arithmetic, same bookkeeping, same epilogue and same keys as smax_pair.cu (see its header and smax_tc.cu's); what changes is
// who stages and who reads the target tiles.  smax_pair.cu's launch is power-limited: it sits on the tensor pipe's issue rate in
// cycles and pays in clock for every byte it moves besides the MMAs (profiles/r2_kernel_experiments.txt).  Here one
// tcgen05.mma.cta_group::2 (M = 256: 128 rows of each CTA, N = 256, K = 32) contracts query tile p of BOTH CTAs with a target
// tile of which each CTA holds only half (128 x 128 B + 128 x 32 B of fold digits) in its shared memory: per 128 x 256 scores
// an SM stages a quarter of what smax_tc.cu stages (half of smax_pair.cu), and the tensor core reads half the B-operand bytes
// from its own shared memory.
//
// Pipeline per CTA pair (persistent over (query tile quad, target range) units; 384 threads per CTA, cluster rank r = 0, 1):
//   warp 0  TMA producer, both CTAs : the unit's two A tiles of this CTA (rows quad * 512 + r * 256 ...), per stage this CTA's half of
//                          the B tile (rows j * 256 + r * 128 ...) + its digits; every load signals the LEADER's (rank 0) full barrier
//                          (cp.async.bulk.tensor .cta_group::2), whose transaction count covers both halves
//   warp 1  MMA issuer, leader only : per target tile, for p = 0, 1: 4 (+1 fold) tcgen05.mma.cta_group::2 into accumulator p of both
//                          CTAs' TMEM; tcgen05.commit multicast to both CTAs' t_full[p] / empty[stage] / a_empty[slot]
//   warps 2-3            : TMEM allocation (cta_group::2, the same warp of both CTAs) / fold constant, then resolvers of this CTA's rows
//   warps 4-11 epilogue  : exactly smax_pair.cu's, on this CTA's 2 x 128 rows; an accumulator goes back to the leader's MMA warp once
//                          the 16 epilogue warps of the pair have copied it out (remote mbarrier arrive on the leader's t_empty[p])
// The launch's unit space is that of smax_pair.cu with q_tile_rows = 512 and "CTA" = cluster (SmaxBatch::n_ctas counts clusters).
//
// MEASURED SLOWER than smax_pair.cu (65 536 x 65 536: 366.6 us against 344.0; profiles/r2_kernel_experiments.txt section 8), so the library
// never picks this walk by itself -- ssrlcv_b200_tune(SSRLCV_B200_TUNE_PAIR_UNITS, 2) runs it (parity suite green with it forced).  The
// accumulator hand-off crosses the pair twice (commit multicast to the peer's t_full, remote arrive on the leader's t_empty: ~215 cycles
// each) and TMEM holds only two 128 x 256 accumulators: with N = 256 MMAs the pipe idles waiting for the peer (395.6 us); four 128-column
// accumulators (kSplitN, below) hide the latency but double the MMA instructions and the A-operand reads.
#include <atomic>

#include "smax_common.cuh"

namespace b200 {
namespace duok {

constexpr int kStages = 5;                             // a stage is HALF a target tile here (20 KB) and serves two MMA groups
constexpr int kASlots = 2;                             // the next unit's A pair loads while this unit's last tiles are contracted
constexpr int kColStages = 6;
constexpr uint32_t kPairRows = 2 * kBM;               // 256 query rows per unit and CTA
constexpr uint32_t kUnitRows = 2 * kPairRows;         // 512 query rows per unit
constexpr uint32_t kHalfN = kBN / 2;                  // target rows of a tile staged by each CTA
constexpr uint32_t kABytes = kBM * kDescBytes;        // 16384
constexpr uint32_t kAPairBytes = 2 * kABytes;
constexpr uint32_t kBBytes = kHalfN * kDescBytes;     // 16384: this CTA's half of a target tile
constexpr uint32_t kBExtBytes = kHalfN * kExtBytes;   // 4096: and of its fold digits
constexpr uint32_t kAExtBytes = kBM * kExtBytes;      // 4096
constexpr uint32_t kColWBytes = kBN * 4;
constexpr uint32_t kColWmaxBytes = (kBN / kChunk) * 4;
constexpr uint32_t kTmemCols = 512;
// Accumulators: kSplitN = false: two of 256 columns (query tile p), N = 256 MMAs.  kSplitN = true: four of 128 columns (query tile p,
// column half h), N = 128 MMAs of rows [h * 64, +64) of each CTA's half tile, each handed back by the four warps per CTA that drain it:
// the hand-off crosses the pair twice (commit multicast ~215 cycles, remote arrive ~215), which a two-deep accumulator ring cannot hide.
// Columns [0, 64) of accumulator (p, h) are targets h * 64 ... of the tile, columns [64, 128) targets 128 + h * 64 ...
constexpr bool kSplitN = true;
constexpr uint32_t kAccs = kSplitN ? 4 : 2;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;        // 384

struct Smem {
  static constexpr uint32_t a = 0;                                       // kASlots x 2 tiles x 16 KB
  static constexpr uint32_t b = a + kASlots * kAPairBytes;               // kStages x 16 KB
  static constexpr uint32_t bext = b + kStages * kBBytes;                // kStages x 4 KB (fold digits)
  static constexpr uint32_t col_w = bext;                                // no-fold: kColStages x 1 KB of W ...
  static constexpr uint32_t col_wmax = col_w + kColStages * kColWBytes;  // ... and kColStages x 32 B of window bounds, in the digit region
  static constexpr uint32_t aext = bext + kStages * kBExtBytes;          // 4 KB
  static constexpr uint32_t bars = aext + kAExtBytes;
  static constexpr uint32_t n_bars = 2 * kStages + 4 + 8 + 2 * kColStages + 4;
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

template <bool kFold, bool kColMask, bool kTiming>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
smax_duo_kernel(const __grid_constant__ SmaxBatch batch) {
  constexpr int32_t kWOffset = kFold ? kFoldC : kNoFoldC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;                         // [kStages]    TMA -> MMA
  uint64_t* empty = bars + kStages;              // [kStages]    MMA -> TMA
  uint64_t* a_full = bars + 2 * kStages;         // [2]
  uint64_t* a_empty = a_full + 2;                // [2]
  uint64_t* t_full = a_empty + 2;                // [kAccs <= 4] MMA -> epilogue (accumulator ready: query tile p, or (p, column half h) at p * 2 + h)
  uint64_t* t_empty = t_full + 4;                // [kAccs <= 4] epilogue -> MMA (accumulator copied out of TMEM)
  uint64_t* c_full = t_empty + 4;                // [kColStages] TMA -> epilogue (W / window bounds, no-fold)
  uint64_t* c_empty = c_full + kColStages;       // [kColStages]
  uint64_t* x_full = c_empty + kColStages;       // [2] epilogue -> resolver warps (window states of a finished unit)
  uint64_t* x_empty = x_full + 2;                // [2] resolver -> epilogue
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();             // 0: leader (issues the MMAs, owns the full / t_empty barriers that count)
  const uint32_t unit_cta = blockIdx.x >> 1;           // the pair's index in the launch's unit space

  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    // t_empty: the leader's counts the epilogue warps of BOTH CTAs (the peer's arrive remotely); the peer's own is never used
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kSplitN ? kEpiWarps : 2 * kEpiWarps); }
    for (int i = 0; i < kColStages; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], kEpiWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], kEpiWarps); mbar_init(&x_empty[i], 2); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc_pair<kTmemCols>(tmem_ptr); tmem_relinquish_pair(); }
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
  cluster_sync_all();                                  // both CTAs' barriers are initialised before either signals the other
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
      UnitIter u(batch, unit_cta);
      while (u.next(batch, unit_cta, kFold)) {
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
        // the leader's barrier expects the bytes of both CTAs; each CTA's loads complete on it
        if (rank == 0) mbar_arrive_expect_tx(&a_full[aslot], 2 * kAPairBytes);
        const uint32_t a_bar = mapa_shared(smem_u32(&a_full[aslot]), 0);
        const uint32_t my_row0 = u.qt * kUnitRows + rank * kPairRows;
        tma_load_2d_pair(smem + Smem::a + aslot * kAPairBytes, map_q, a_bar, 0, static_cast<int32_t>(my_row0));
        tma_load_2d_pair(smem + Smem::a + aslot * kAPairBytes + kABytes, map_q, a_bar, 0, static_cast<int32_t>(my_row0 + kBM));
        for (uint32_t j = u.j0; j < u.j1; ++j) {
          long long p0 = 0;
          if (kTiming) p0 = clock64();
          mbar_wait(&empty[stage], phase ^ 1);
          if (kTiming) tp_wait += clock64() - p0;
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (kBBytes + (kFold ? kBExtBytes : 0)));
          const uint32_t f_bar = mapa_shared(smem_u32(&full[stage]), 0);
          const int32_t my_t0 = static_cast<int32_t>(j * kBN + rank * kHalfN);      // accumulator columns [rank * 128, +128)
          tma_load_2d_pair(smem + Smem::b + stage * kBBytes, map_t, f_bar, 0, my_t0);
          if (kFold) tma_load_2d_pair(smem + Smem::bext + stage * kBExtBytes, map_text, f_bar, 0, my_t0);
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
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(2 * kBM, kSplitN ? kBN / 2 : kBN, 0, 0);       // M = 256: 128 rows of each CTA
      const uint64_t aext_desc = make_smem_desc(smem_u32(smem + Smem::aext), 256, kSwizzle32B);
      uint32_t stage = 0, phase = 0, aslot = 0, aphase = 0, accphase = 0;
      long long tm_wait_acc = 0, tm_wait_full = 0, tm_total0 = 0;
      if (kTiming) tm_total0 = clock64();
      UnitIter u(batch, unit_cta);
      while (u.next(batch, unit_cta, kFold)) {
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
          for (uint32_t acc = 0; acc < kAccs; ++acc) {
            const uint32_t p = kSplitN ? acc >> 1 : acc, h = kSplitN ? acc & 1u : 0u;
            const uint32_t d_tmem = tmem_base + p * kBN + h * (kBN / 2);
            const uint64_t a_desc = p ? a_desc1 : a_desc0;
            // split: rows [h * 64, +64) of each CTA's half tile (address field is in 16-byte units)
            const uint64_t b_desc_h = b_desc + h * ((kHalfN / 2) * kDescBytes >> 4), bext_desc_h = bext_desc + h * ((kHalfN / 2) * kExtBytes >> 4);
            long long w0 = 0;
            if (kTiming) w0 = clock64();
            mbar_wait(&t_empty[acc], accphase ^ 1);
            if (kTiming) tm_wait_acc += clock64() - w0;
            tc_fence_after();
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)   // +32 bytes along K inside the 128-byte swizzle atom = +2 in the address field
              mma_i8_ss_pair(d_tmem, a_desc + 2 * k, b_desc_h + 2 * k, idesc, k);
            if (kFold && !(kDiagBuild && (batch.diag & 2u))) mma_i8_ss_pair(d_tmem, aext_desc, bext_desc_h, idesc, 1);
            mma_commit_pair(&t_full[acc]);
          }
          mma_commit_pair(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          accphase ^= 1;
        }
        mma_commit_pair(&a_empty[aslot]);
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
    UnitIter u(batch, unit_cta);
    while (u.next(batch, unit_cta, kFold)) {
      if (u.last) break;
      long long w0 = 0, w1 = 0;
      if (kTiming) w0 = clock64();
      mbar_wait(&x_full[uparity], (uiter >> 1) & 1);
      if (kTiming) w1 = clock64();
      const uint32_t row_base = u.qt * kUnitRows + rank * kPairRows;
      const WinState* xa = xch + (uparity * 2 + 0) * kPairRows;
      const WinState* xb = xch + (uparity * 2 + 1) * kPairRows;
      if (row_base < u.J->a.nq_pad) {                    // the peer's half of a set's last unit may lie beyond the padded rows: no keys there
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
      }
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
    const uint32_t t_empty_leader = mapa_shared(smem_u32(&t_empty[0]), 0);
    // first window (of the tile's eight) and step of this warp's four: columns [32k, 32k + 32) of its accumulator slice
    const uint32_t win_first = kSplitN ? half * 2 : half * 4;
    auto win_of = [&](uint32_t k) { return kSplitN ? win_first + (k >> 1) * 4 + (k & 1u) : win_first + k; };
    uint32_t it = 0, uiter = 0;                       // it: target tiles consumed so far -> phases of the accumulators and the column ring
    long long te_wait = 0, te_work = 0, te_resolve = 0, te_total0 = 0, te_cwait = 0, te_rare = 0, te_rare_n = 0, te_ld = 0, te_arr = 0;
    if (kTiming) te_total0 = clock64();
    UnitIter u(batch, unit_cta);
    while (u.next(batch, unit_cta, kFold)) {
      const SmaxArgs& args = u.J->a;                               // device memory: read what the unit needs up front
      const uint32_t row_base = u.qt * kUnitRows + rank * kPairRows;
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
          const uint32_t wmax_addr = smem_u32(smem + Smem::col_wmax + cs * kColWmaxBytes);
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k) asm volatile("ld.shared.s32 %0, [%1];" : "=r"(wb[k]) : "r"(wmax_addr + 4 * win_of(k)) : "memory");
          if (kTiming) te_cwait += clock64() - c0;
        }
        const uint32_t s_w = smem_u32(smem + Smem::col_w + cs * kColWBytes) + win_first * (kChunk * 4);
#pragma unroll
        for (uint32_t p = 0; p < 2; ++p) {
          long long e0 = 0, e1 = 0;
          if (kTiming) e0 = clock64();
          const uint32_t acc = kSplitN ? p * 2 + half : p;
          mbar_wait(&t_full[acc], accphase);
          if (kTiming) e1 = clock64();
          tc_fence_after();
          // this warp's 32 x 128 slice of accumulator p into registers, and the accumulator goes straight back to the MMA warp
          uint32_t v[4][32];
          tmem_ld_32x128(taddr0 + p * kBN, v);
          tmem_ld_wait();
          long long l1 = 0;
          if (kTiming) { l1 = clock64(); te_ld += l1 - e1; }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&t_empty[acc]); else mbar_arrive_cluster(t_empty_leader + acc * 8);
          }
          if (kTiming) { __syncwarp(); te_arr += clock64() - l1; }
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
                if ((live >> k) & 1u) win_update(st, wm[k], j * 8 + win_of(k), eq_mask32(v[k], wm[k]));
            } else {
              ld_state(st_addr, st);
              window_bookkeeping<4, kFold, kColMask, kSplitN>(v, wm, st, j * 8 + win_first, s_w, /* runner-up searches take smax_tc.cu */ false, args, 0u);
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
        if (row_base < args.nq_pad) {
          if (!kColMask) resolve_rows_windows<SSRLCV_B200_MASK_NONE, 2>(args, xa, xb, row_base, u.slot, (warp - 4) * 32, 32, lane);
          else resolve_row(args, xa[r], xb[r], row_base, u.slot, r, kWOffset);
          fill_unused_slots(args, row_base, u.slot, u.fill_end, (warp - 4) * 32, 32, lane);
        }
      }
      ++uiter;
      if (kTiming) te_resolve += clock64() - r0;
    }
    if (kTiming && lane == 0) {
      const uint32_t e = warp - 4;
      if (e == 0) {
        const uint32_t o = blockIdx.x * 16 + 4;
        batch.timing[o + 0] = te_wait; batch.timing[o + 1] = te_work; batch.timing[o + 2] = te_resolve; batch.timing[o + 3] = clock64() - te_total0;
        batch.timing[o + 4] = te_ld; batch.timing[o + 5] = te_arr;        // of te_work: the TMEM load, the hand-back
        if (e == 0) { batch.timing[blockIdx.x * 16 + 14] = te_rare_n; batch.timing[blockIdx.x * 16 + 15] = kFold ? te_rare : te_cwait; }   // fold: cycles inside the rare path
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                                  // neither CTA's shared memory or TMEM goes away while the pair still uses it
  if (warp == 2) tmem_dealloc_pair<kTmemCols>(tmem_base);
}

template <bool kFold, bool kColMask, bool kTiming>
static cudaError_t launch_one(cudaStream_t st, const SmaxBatch& b) {
  auto kern = smax_duo_kernel<kFold, kColMask, kTiming>;
  // the dynamic shared-memory opt-in is a per-device attribute of the function (one context per device in multi-GPU jobs)
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64 || !((configured.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes));
    if (e != cudaSuccess) return e;
    if (dev < 64) configured.fetch_or(1ull << dev, std::memory_order_release);
  }
  return launch_pdl(kern, dim3(2 * b.n_ctas), dim3(kThreads), kSmemBytes, st, b);   // n_ctas counts CTA pairs
}

}  // namespace duok

// How many CTA pairs of the matcher the current device holds at once (one CTA per SM, both SMs of a TPC): the launch's unit space
// is cut into that many ranges.  0: the device cannot run the pair kernel.
int smax_duo_max_pairs() {
  using namespace duok;
  static std::atomic<int> cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64) { const int c = cached[dev].load(std::memory_order_acquire); if (c != 0) return c > 0 ? c : 0; }
  auto kern = smax_duo_kernel<true, true, false>;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBytes)) == cudaSuccess) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(sms & ~1)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
    if (n > sms / 2) n = sms / 2;
  } else {
    cudaGetLastError();
  }
  if (dev < 64) cached[dev].store(n > 0 ? n : -1, std::memory_order_release);
  return n;
}

// one launch over all jobs of `b` (unmasked, same fold variant and bookkeeping regime, SmaxArgs::q_tile_rows == 512, n_ctas = CTA pairs)
cudaError_t launch_smax_duo(cudaStream_t st, const SmaxBatch& b, bool fold, bool colmask) {
  using namespace duok;
  if (b.timing) {   // diagnostics build (tools/kerneltiming.py)
    if (fold) return colmask ? launch_one<true, true, true>(st, b) : launch_one<true, false, true>(st, b);
    return colmask ? launch_one<false, true, true>(st, b) : launch_one<false, false, true>(st, b);
  }
  if (fold) return colmask ? launch_one<true, true, false>(st, b) : launch_one<true, false, false>(st, b);
  return colmask ? launch_one<false, true, false>(st, b) : launch_one<false, false, false>(st, b);
}

}  // namespace b200
