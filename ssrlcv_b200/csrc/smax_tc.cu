// smax_tc.cu -- the fused tensor-core matcher for sm_100a.
//
// Brute-force 128-D nearest neighbour as the dense contraction it is.  For query a and target b
//     d2(a,b) = |a|^2 + |b|^2 - 2 a.b ,        |b|^2 = 2h + p  (p = parity)
// so minimising d2 over b is maximising the integer score
//     S(a,b) = a.b + W_b ,   W_b = C - h_b  (C constant)    (d2 = |a|^2 - 2 (S - C) + p)
// S is produced ENTIRELY by the tensor core: four tcgen05.mma kind::i8 (u8 x u8 -> s32, K = 32 each)
// contract the 128 descriptor bytes and a fifth contracts a constant vector [1,255,...,255] with 32
// base-255 digits of W_b ("fold"), accumulating in TMEM.  The epilogue therefore does no arithmetic on
// the accumulator except a running 3-input max: per 32-target window it keeps the window maximum and,
// only when a window reaches the query's running best, records the window and WHICH of its 32 columns
// attained the maximum (a 32-bit mask).  Nothing of the 128 x Nt score matrix ever reaches shared or
// global memory.
//
// S is an exact integer, so the recorded columns ARE the exact answer: for a column f with score S
//     d2 = |a|^2 - 2 S + 2 C + (|b_f|^2 & 1)
// and the reference's winner (ref: src/MatchFactory.cu:1476-1492, min of (d2, f mod 32, f)) is the
// minimum of that key over the recorded columns -- equal scores differ in d2 only by the parity bit.
// No descriptor is read a second time.  Three or more windows tying at the best score (rare) do not
// fit the per-query state and are flagged for the exact SIMT kernel (exact.cu) instead of being guessed.
//
// Pipeline per CTA (one CTA per SM, persistent over (query tile, target split) work units):
//   warp 0  TMA producer : A tile (128 x 128 B, SWIZZLE_128B) per unit; per stage a B tile
//                          (256 x 128 B, SWIZZLE_128B) + its digit tile (256 x 32 B, SWIZZLE_32B)
//   warp 1  MMA issuer   : 5 x tcgen05.mma (M128 N256 K32) per stage into one of two 256-column
//                          TMEM accumulators; tcgen05.commit releases the smem stage and publishes
//                          the accumulator
//   warp 2  TMEM allocator, warp 3 writes the constant fold operand
//   warps 4-11 epilogue  : two warps per TMEM lane quarter (one per half of the 256 columns) so every SM
//                          sub-partition has two epilogue warps to hide latency: tcgen05.ld 32x32b.x32,
//                          max3 tree, window bookkeeping
//   warps 2-3 also turn the recorded (score, window, column mask) states of a finished unit into keys while
//   the pipe works on the next unit (one lane per query: a norm parity load and a 64-bit key each)
// Masked variants (epipolar constraints) evaluate the reference's float mask per (query, target) in the
// epilogue on target locations staged through a small shared-memory ring.
#include <atomic>

#include "smax_common.cuh"

namespace b200 {

constexpr int kStages = 4;
constexpr int kColStages = 8;   // deeper than the B ring: the producer must never wait for the epilogue before it can fetch B tiles
constexpr uint32_t kABytes = kBM * kDescBytes;       // 16384
constexpr uint32_t kBBytes = kBN * kDescBytes;       // 32768
constexpr uint32_t kBExtBytes = kBN * kExtBytes;     // 8192
constexpr uint32_t kAExtBytes = kBM * kExtBytes;     // 4096
constexpr uint32_t kColLocBytes = kBN * 8;           // float2 per target
constexpr uint32_t kColWBytes = kBN * 4;             // int32 W per target
constexpr uint32_t kColWmaxBytes = (kBN / kChunk) * 4;  // int32 largest W per 32-target window
constexpr uint32_t kColBoxBytes = (kBN / kChunk) * 16;  // float4 bounding box per 32-target window (spatially sorted target sets)
constexpr uint32_t kTmemCols = 512;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (4 + kEpiWarps) * 32;      // 384

struct SmemLayout {
  static constexpr uint32_t a = 0;                                       // 2 x 16 KB
  static constexpr uint32_t b = a + 2 * kABytes;                         // kStages x 32 KB
  static constexpr uint32_t bext = b + kStages * kBBytes;                // kStages x 8 KB
  static constexpr uint32_t aext = bext + kStages * kBExtBytes;          // 4 KB
  static constexpr uint32_t col_loc = aext + kAExtBytes;                 // kColStages x 2 KB   (masked variants)
  // the no-fold variants have no digit tiles: their per-target W ring and window-bound ring live in the bext region
  static constexpr uint32_t col_w = bext;                                // kColStages x 1 KB   (no-fold variants)
  static constexpr uint32_t col_wmax = col_w + kColStages * kColWBytes;  // kColStages x 32 B   (8 window bounds per tile)
  static constexpr uint32_t col_box = col_loc + kColStages * kColLocBytes;   // kColStages x 128 B: the tile's 8 window boxes (sorted sets)
  static constexpr uint32_t bars = col_box + kColStages * kColBoxBytes;
  static constexpr uint32_t n_bars = 2 * kStages + 4 + 4 + 2 * kColStages + 4;
  static constexpr uint32_t tmem_ptr = bars + n_bars * 8;
  static constexpr uint32_t xch = tmem_ptr + 16;                         // 2 (unit parity) x 2 (half) x 128 rows x 24 B
  static constexpr uint32_t total = xch + 2 * 2 * kBM * 24;
};
constexpr uint32_t kSmemBytes = SmemLayout::total + 1024;   // + slack for manual 1024-byte alignment

// kMask: SSRLCV_B200_MASK_*.
// kFold: W_b carried by the fifth MMA (true: +25 % tensor-pipe time, the accumulator IS the score) or left out of the
//        contraction (false: four MMAs per tile, no digit tile; the epilogue takes the window maximum of a.b, adds the
//        window's largest W -- a bound on every score of the window, pack.cu wmax -- and only when that bound reaches some
//        lane's running best adds W to the 32 values and takes the exact maximum.  Any W, no fold-digit limit.)
//        (A third way -- the epilogue warps storing W_b into the accumulator with tcgen05.st so that four accumulating MMAs
//        suffice -- was built and measured: the tensor pipe is undisturbed, but the epilogue warps' serial ld -> st -> max
//        chain takes 1 200 cycles per tile; 0.76 ms against 0.37 ms at 65 536 x 65 536.  profiles/r1c_init_variant.txt)
// kColMask: record WHICH columns of a window attained its maximum (true) or only the window (false), see above
//        unit_records_columns(); chosen per job by the host (job_records_columns).
// kTiming: per-role cycle counters (diagnostics build of the same kernel; never used by the product path).
// kLean (masked, fold): for searches whose bookkeeping path is next to never taken -- SEEDED masked searches, the reference
//        pipeline's own call: the per-query seed cap leaves 0.04 % of the (query, window) pairs on real data.  The masked
//        instantiations carry a column-data ring (target locations / window boxes staged per tile, two more barrier operations per
//        tile in the epilogue), the dead-tile iterator and the query's mask state through the hot loop; ptxas, capped at 168
//        registers beside a 128-register accumulator slice, answers with local-memory traffic on the hand-off chain: 168 us for the
//        fixture's seeded DoubleConstrained search against 108 us unmasked (profiles/r2_kernel_experiments.txt section 9).  The lean
//        instantiation has the UNMASKED hot loop: no ring, no boxes, no dead-tile bitmap (sorted copies exist for unseeded searches
//        only); the bookkeeping path reads the mask state and the 32 target locations of a window from global memory.
template <int kMask, bool kFold, bool kColMask, bool kTiming, bool kLean = false>
__global__ void __launch_bounds__(kThreads, 1)
smax_kernel(const __grid_constant__ SmaxBatch batch) {
  static_assert(!kLean || (kMask != SSRLCV_B200_MASK_NONE && kFold && !kTiming), "lean: masked fold variants");
  constexpr bool kCols = ((kMask != SSRLCV_B200_MASK_NONE) && !kLean) || !kFold;   // epilogue needs staged per-target data (locations / W)
  constexpr int32_t kWOffset = kFold ? kFoldC : kNoFoldC;               // the C of W_b = C - h_b
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::bars);
  uint64_t* full = bars;                         // [kStages]   TMA -> MMA
  uint64_t* empty = bars + kStages;              // [kStages]   MMA -> TMA
  uint64_t* a_full = bars + 2 * kStages;         // [2]
  uint64_t* a_empty = a_full + 2;                // [2]
  uint64_t* t_full = a_empty + 2;                // [2]         MMA -> epilogue (accumulator ready)
  uint64_t* t_empty = t_full + 2;                // [2]         epilogue -> MMA (accumulator copied out of TMEM)
  uint64_t* c_full = t_empty + 2;                // [kColStages] TMA -> epilogue (target columns data)
  uint64_t* c_empty = c_full + kColStages;       // [kColStages]
  uint64_t* x_full = c_empty + kColStages;       // [2] epilogue -> resolver warps (window states of a finished unit)
  uint64_t* x_empty = x_full + 2;                // [2] resolver -> epilogue
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SmemLayout::tmem_ptr);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], kEpiWarps); }
    for (int i = 0; i < kColStages; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], kEpiWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], kEpiWarps); mbar_init(&x_empty[i], 2); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc<kTmemCols>(tmem_ptr); tmem_relinquish(); }
  if (warp == 3) {
    // constant A-side operand of the fold MMA: every row = [1, 255 x 31], stored in the SWIZZLE_32B
    // K-major layout (32-byte rows, 16-byte chunk index ^= bit 7 of the byte address).
    for (int r = lane; r < kBM; r += 32) {
      const uint32_t sw = (r >> 2) & 1;
      uint4 c0 = make_uint4(0xFFFFFF01u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      uint4 c1 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      uint4* row = reinterpret_cast<uint4*>(smem + SmemLayout::aext + r * kExtBytes);
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
        const float2* t_loc = u.J->a.t_loc;
        const int32_t* t_wcol = u.J->a.t_wcol;
        const int32_t* t_wmax = u.J->a.t_wmax;
        const float4* t_wbox = u.J->a.t_wbox;
        const uint32_t* live_bits = (kMask != SSRLCV_B200_MASK_NONE && !kLean) ? u.J->a.live_bits : nullptr;
        const uint32_t live_stride = u.J->a.live_stride;
        mbar_wait(&a_empty[aslot], aphase ^ 1);
        mbar_arrive_expect_tx(&a_full[aslot], kABytes);
        tma_load_2d(smem + SmemLayout::a + aslot * kABytes, map_q, &a_full[aslot], 0, static_cast<int32_t>(u.qt * kBM));
        LiveTiles tiles(live_bits, live_stride, u.qt, u.j0, u.j1);     // dead pairs: every role skips them
        uint32_t j;
        while (tiles.next(j)) {
          long long p0 = 0;
          if (kTiming) p0 = clock64();
          mbar_wait(&empty[stage], phase ^ 1);
          if (kTiming) tp_wait += clock64() - p0;
          if ((kDiagBuild && (batch.diag & 1u)) && (j & 1u)) {
            mbar_arrive(&full[stage]);
          } else {
          mbar_arrive_expect_tx(&full[stage], kBBytes + (kFold ? kBExtBytes : 0));
          tma_load_2d(smem + SmemLayout::b + stage * kBBytes, map_t, &full[stage], 0, static_cast<int32_t>(j * kBN));
          if (kFold) tma_load_2d(smem + SmemLayout::bext + stage * kBExtBytes, map_text, &full[stage], 0, static_cast<int32_t>(j * kBN));
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (kCols) {
            mbar_wait(&c_empty[cs], cphase ^ 1);
            const uint32_t bytes = (kMask != SSRLCV_B200_MASK_NONE ? kColLocBytes : 0) + (!kFold ? kColWBytes + kColWmaxBytes : 0) +
                                   (t_wbox != nullptr ? kColBoxBytes : 0);
            mbar_arrive_expect_tx(&c_full[cs], bytes);
            if (t_wbox != nullptr)
              bulk_load_1d(smem + SmemLayout::col_box + cs * kColBoxBytes, t_wbox + static_cast<uint64_t>(j) * (kBN / kChunk), kColBoxBytes, &c_full[cs]);
            if (kMask != SSRLCV_B200_MASK_NONE)
              bulk_load_1d(smem + SmemLayout::col_loc + cs * kColLocBytes, t_loc + static_cast<uint64_t>(j) * kBN, kColLocBytes, &c_full[cs]);
            if (!kFold) {   // W of the tile's 256 targets and the largest W of each of its 8 windows (pack.cu wcol / wmax)
              bulk_load_1d(smem + SmemLayout::col_w + cs * kColWBytes, t_wcol + static_cast<uint64_t>(j) * kBN, kColWBytes, &c_full[cs]);
              bulk_load_1d(smem + SmemLayout::col_wmax + cs * kColWmaxBytes, t_wmax + static_cast<uint64_t>(j) * (kBN / kChunk), kColWmaxBytes, &c_full[cs]);
            }
            if (++cs == kColStages) { cs = 0; cphase ^= 1; }
          }
        }
        if (++aslot == 2) { aslot = 0; aphase ^= 1; }
      }
      if (kTiming) batch.timing[blockIdx.x * 16 + 3] = tp_wait;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_i8(kBM, kBN, 0, 0);
      const uint64_t aext_desc = make_smem_desc(smem_u32(smem + SmemLayout::aext), 256, kSwizzle32B);
      uint32_t stage = 0, phase = 0, aslot = 0, aphase = 0, acc = 0, accphase = 0;
      long long tm_wait_acc = 0, tm_wait_full = 0, tm_total0 = 0;
      if (kTiming) tm_total0 = clock64();
      UnitIter u(batch, blockIdx.x);
      while (u.next(batch, blockIdx.x, kFold)) {
        mbar_wait(&a_full[aslot], aphase);
        const uint64_t a_desc = make_smem_desc(smem_u32(smem + SmemLayout::a + aslot * kABytes), 1024, kSwizzle128B);
        const uint32_t* live_bits = (kMask != SSRLCV_B200_MASK_NONE && !kLean) ? u.J->a.live_bits : nullptr;
        const uint32_t live_stride = u.J->a.live_stride;
        LiveTiles tiles(live_bits, live_stride, u.qt, u.j0, u.j1);
        uint32_t j;
        while (tiles.next(j)) {
          long long c0 = 0, c1 = 0, c2 = 0;
          if (kTiming) c0 = clock64();
          mbar_wait(&t_empty[acc], accphase ^ 1);
          if (kTiming) c1 = clock64();
          mbar_wait(&full[stage], phase);
          if (kTiming) { c2 = clock64(); tm_wait_acc += c1 - c0; tm_wait_full += c2 - c1; }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kBN;
          const uint64_t b_desc = make_smem_desc(smem_u32(smem + SmemLayout::b + stage * kBBytes), 1024, kSwizzle128B);
#pragma unroll
          for (uint32_t k = 0; k < 4; ++k)   // +32 bytes along K inside the 128-byte swizzle atom = +2 in the address field
            mma_i8_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k);
          if (kFold && !(kDiagBuild && (batch.diag & 2u))) {
            const uint64_t bext_desc = make_smem_desc(smem_u32(smem + SmemLayout::bext + stage * kBExtBytes), 256, kSwizzle32B);
            mma_i8_ss(d_tmem, aext_desc, bext_desc, idesc, 1);
          }
          mma_commit(&empty[stage]);
          mma_commit(&t_full[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (++acc == 2) { acc = 0; accphase ^= 1; }
        }
        mma_commit(&a_empty[aslot]);
        if (++aslot == 2) { aslot = 0; aphase ^= 1; }
      }
      if (kTiming) {
        batch.timing[blockIdx.x * 16 + 0] = tm_wait_acc;
        batch.timing[blockIdx.x * 16 + 1] = tm_wait_full;
        batch.timing[blockIdx.x * 16 + 2] = clock64() - tm_total0;
      }
    }
  } else if (warp == 2 || warp == 3) {
    // =============================== resolver ===============================
    // Exact evaluation of the recorded windows (resolve_rows), off the epilogue's critical path: while the tensor pipe
    // and the epilogue warps already work on the next unit, these two warps (64 queries each) finish the previous one.
    // The CTA's last unit has nothing to overlap with and is resolved by the eight epilogue warps themselves.
    const WinState* xch = reinterpret_cast<const WinState*>(smem + SmemLayout::xch);
    const uint32_t rw = warp - 2;
    uint32_t uparity = 0, uiter = 0;
    long long tr_wait = 0, tr_work = 0;
    UnitIter u(batch, blockIdx.x);
    while (u.next(batch, blockIdx.x, kFold)) {
      if (u.last) break;
      long long w0 = 0, w1 = 0;
      if (kTiming) w0 = clock64();
      mbar_wait(&x_full[uparity], (uiter >> 1) & 1);
      if (kTiming) w1 = clock64();
      if (kColMask) {
#pragma unroll
        for (uint32_t g = 0; g < 2; ++g) {
          const uint32_t r = rw * 64 + g * 32 + lane;
          resolve_row(u.J->a, xch[(uparity * 2 + 0) * kBM + r], xch[(uparity * 2 + 1) * kBM + r], u.qt * kBM, u.slot, r, kWOffset);
        }
      } else {
        resolve_rows_windows<kMask, 2>(u.J->a, xch + (uparity * 2 + 0) * kBM, xch + (uparity * 2 + 1) * kBM, u.qt * kBM, u.slot, rw * 64, 64, lane);
      }
      fill_unused_slots(u.J->a, u.qt * kBM, u.slot, u.fill_end, rw * 64, 64, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[uparity]);
      uparity ^= 1;
      ++uiter;
      if (kTiming) { const long long w2 = clock64(); tr_wait += w1 - w0; tr_work += w2 - w1; }
    }
    if (kTiming && lane == 0 && rw == 0) { batch.timing[blockIdx.x * 16 + 12] = tr_wait; batch.timing[blockIdx.x * 16 + 13] = tr_work; }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const uint32_t quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const uint32_t half = (warp - 4) >> 2;             // which 128 of the 256 accumulator columns
    const uint32_t lane_addr = (quarter * 32u) << 16;
    constexpr uint32_t kWinPerTile = kBN / kChunk;     // 8
    constexpr uint32_t kWinPerWarp = kWinPerTile / 2;  // 4
    WinState* xch = reinterpret_cast<WinState*>(smem + SmemLayout::xch);
    uint32_t it = 0, uiter = 0;                       // it: tiles consumed so far -> accumulator slot / column stage and their phases
    long long te_wait = 0, te_work = 0, te_resolve = 0, te_total0 = 0, te_live = 0, te_cwait = 0;
    if (kTiming) te_total0 = clock64();
    UnitIter u(batch, blockIdx.x);
    while (u.next(batch, blockIdx.x, kFold)) {
      const SmaxArgs& args = u.J->a;                               // device memory: read what the unit needs up front
      const uint32_t row = u.qt * kBM + quarter * 32 + lane;       // this thread's query
      const uint32_t row_c = min(row, args.nq_pad - 1);
      const float eps = args.eps;
      const bool boxed = !kLean && args.t_wbox != nullptr;
      const uint32_t* live_bits = (kMask != SSRLCV_B200_MASK_NONE && !kLean) ? args.live_bits : nullptr;
      const uint32_t live_stride = args.live_stride;
      QueryMask qm;
      if (kMask != SSRLCV_B200_MASK_NONE && !kLean) qm = load_query_mask<kMask>(args.q_mask + 2 * static_cast<uint64_t>(row_c));

      // Lowest score that can still pass the absolute threshold: d2 = |a|^2 - 2 (S - C) + p <= d2_cap
      // <=> S >= C + ceil((|a|^2 - d2_cap) / 2)  (p = 0 is the permissive case).  Real scores are >= 2,
      // padding / masked entries are 0, so the floor never drops below 2.
      WinState st;
      {
        int32_t floor_s = 2;
        int32_t d2_cap = args.d2_cap;
        if (args.q_seed != nullptr && row < args.nq) {
          // largest d2 whose ratio test (float)d2 / seed <= ratio can still pass, generously rounded up (finalize applies the
          // reference's exact float test; this only has to keep every candidate that could pass it)
          const float c = __fmul_rn(__fmul_rn(args.seed_ratio, args.q_seed[row]), 1.00001f);
          if (c < 8.4e6f) d2_cap = min(d2_cap, static_cast<int32_t>(c) + 1);
        }
        if (d2_cap != 0x7fffffff) {
          const int32_t qn = args.q_norm[row_c];
          floor_s = max(2, kWOffset + ((qn - d2_cap + 1) >> 1));
        }
        st.best = floor_s - 1; st.w1 = kNoWin; st.w2 = kNoWin; st.m1 = 0; st.m2 = 0; st.ovf = 0;
      }
      LiveTiles tiles(live_bits, live_stride, u.qt, u.j0, u.j1);
      uint32_t j;
      while (tiles.next(j)) {
        long long e0 = 0, e1 = 0;
        if (kTiming) e0 = clock64();
        const uint32_t acc = it & 1, accphase = (it >> 1) & 1, cs = it % kColStages, cphase = (it / kColStages) & 1;
        ++it;
        // the tile's column data (staged several tiles ahead) first: nothing then stands between the accumulator and the max trees
        long long cw0 = 0;
        if (kTiming) cw0 = clock64();
        if (kCols) mbar_wait(&c_full[cs], cphase);
        if (kTiming) te_cwait += clock64() - cw0;
        int32_t wbound[kWinPerWarp] = {0, 0, 0, 0};                // no-fold: largest W of each of this warp's windows
        if (!kFold) {
          const int4 wb = ld_shared_v4(smem_u32(smem + SmemLayout::col_wmax + cs * kColWmaxBytes + half * (kWinPerWarp * 4)));
          wbound[0] = wb.x; wbound[1] = wb.y; wbound[2] = wb.z; wbound[3] = wb.w;
        }
        mbar_wait(&t_full[acc], accphase);
        if (kTiming) e1 = clock64();
        tc_fence_after();
        const float2* s_loc = reinterpret_cast<const float2*>(smem + SmemLayout::col_loc + cs * kColLocBytes);
        const int32_t* s_w = reinterpret_cast<const int32_t*>(smem + SmemLayout::col_w + cs * kColWBytes);
        const uint32_t c_begin = half * kWinPerWarp;
        const uint32_t taddr = tmem_base + acc * kBN + lane_addr + c_begin * kChunk;
        // Copy this warp's 32 x 128 slice of the accumulator into registers and hand the TMEM buffer straight back to
        // the MMA warp: the tensor pipe never waits for the reduction below, only for these four loads.
        uint32_t v[kWinPerWarp][32];
        tmem_ld_32x128(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);

        // Window maxima of this warp's 4 windows, straight from the accumulator: one branch-free block (the four max trees
        // interleave).  Fold, unmasked: these ARE the windows' best scores.  Otherwise they are upper bounds of them:
        //   no-fold: the accumulator holds a.b only; max(a.b) + (largest W of the window) bounds every score of the window
        //   masked : the unmasked maximum bounds the masked one
        int32_t wm[kWinPerWarp];
        if (kDiagBuild && (batch.diag & 4u)) {
#pragma unroll
          for (uint32_t k = 0; k < kWinPerWarp; ++k) wm[k] = static_cast<int32_t>((v[k][0] | v[k][31]) >> 24);
        } else {
#pragma unroll
        for (uint32_t k = 0; k < kWinPerWarp; ++k) wm[k] = max32(v[k]) + wbound[k];
        }
        uint32_t culled = 0;                                       // bit k: this lane's query cannot match anything in window k
        if (kMask != SSRLCV_B200_MASK_NONE && boxed) {
          // spatially sorted targets: a window's rows lie in a small box of the image; a query whose epipolar band misses the box
          // (geometry.cuh box_skips: exact, NaN-safe) skips the window without looking at it
#pragma unroll
          for (uint32_t k = 0; k < kWinPerWarp; ++k) {
            const int4 bx = ld_shared_v4(smem_u32(smem + SmemLayout::col_box + cs * kColBoxBytes + (c_begin + k) * 16));
            const bool outside = box_skips<kMask>(qm, make_float4(__int_as_float(bx.x), __int_as_float(bx.y), __int_as_float(bx.z), __int_as_float(bx.w)), eps);
            wm[k] = outside ? 0 : wm[k];
            culled |= outside ? (1u << k) : 0u;
          }
        }
        // One warp-uniform test per tile: does any window reach some lane's running best?  (Thresholded search: next to never.)
        const int32_t mt = max(max(wm[0], wm[1]), max(wm[2], wm[3]));
        if (__any_sync(0xffffffffu, mt >= st.best)) {
          if (kMask != SSRLCV_B200_MASK_NONE || !kFold) {
            // turn the bounds of the windows that matter into exact maxima; the others stay bounds below every lane's best
#pragma unroll
            for (uint32_t k = 0; k < kWinPerWarp; ++k) {
              if (!__any_sync(0xffffffffu, wm[k] >= st.best)) continue;
              uint32_t(&x)[32] = v[k];
              if (kTiming) ++te_live;
              if (!kFold) {                                        // exact scores, as the fold variant gets them from the tensor pipe
                const uint32_t wp = smem_u32(s_w + (c_begin + k) * kChunk);
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                  const int4 w4 = ld_shared_v4(wp + 4 * e);
                  x[e] += static_cast<uint32_t>(w4.x); x[e + 1] += static_cast<uint32_t>(w4.y);
                  x[e + 2] += static_cast<uint32_t>(w4.z); x[e + 3] += static_cast<uint32_t>(w4.w);
                }
                wm[k] = ((culled >> k) & 1u) ? 0 : max32(x);
                if (kMask != SSRLCV_B200_MASK_NONE && !__any_sync(0xffffffffu, wm[k] >= st.best)) continue;
              }
              if (kMask != SSRLCV_B200_MASK_NONE) {
                if (kLean) qm = load_query_mask_rare<kMask>(args.q_mask + 2 * static_cast<uint64_t>(row_c));   // (rare: not kept across the hot loop)
                const float2* g_loc = args.t_loc + (static_cast<uint64_t>(j) * kBN + (c_begin + k) * kChunk);   // lean: the window's locations, global
                // The reference's float mask (ref: src/MatchFactory.cu:2241-2254, :1613-1627) for the queries (lanes) this window
                // matters to -- usually one.  The warp evaluates it together: the live query's mask state is broadcast, lane l
                // tests target l of the window, the ballot comes back to the query's lane as a column mask.  Masked entries
                // score 0, below every floor (as do the rows past the end of the target set: zero padding with W = 0).
                uint32_t todo = __ballot_sync(0xffffffffu, wm[k] >= st.best);
                const bool mine = (todo >> lane) & 1u;             // this lane's query is one of the live ones
                if (static_cast<uint32_t>(__popc(todo)) <= batch.mask_coop_max) {
                  uint32_t skip = 0;
                  const float2 tl = kLean ? __ldg(g_loc + lane) : ld_shared_f2(smem_u32(s_loc + (c_begin + k) * kChunk + lane));
                  while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const QueryMask qs = shfl_query_mask<kMask>(qm, src);
                    const uint32_t b = __ballot_sync(0xffffffffu, mask_skips<kMask>(qs, tl.x, tl.y, eps));
                    skip = lane == src ? b : skip;
                  }
#pragma unroll
                  for (int e = 0; e < 32; ++e) x[e] = (skip >> e) & 1u ? 0u : x[e];
                } else {
                  // Many live queries (loose thresholds on real data).  Only a column whose score reaches its row's running best can
                  // change anything, and once a query has seen a few masked-in targets that is one or two of a window's 32: every
                  // lane walks the candidate columns of its own row (location by shared-memory address) and zeroes the masked-out
                  // ones; the other columns stay unverified below the best.  A window in which some row has many candidates
                  // (a query's first windows) is tested column by column as before.
                  const uint32_t lp = smem_u32(s_loc + (c_begin + k) * kChunk);
                  uint32_t cand = 0;
                  if (mine) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) cand |= (static_cast<int32_t>(x[e]) >= st.best) ? (1u << e) : 0u;
                  }
                  if (__reduce_max_sync(0xffffffffu, static_cast<uint32_t>(__popc(cand))) <= batch.mask_walk_max) {
                    uint32_t dead = 0;
                    while (__any_sync(0xffffffffu, cand != 0u)) {
                      if (cand) {
                        const uint32_t e = __ffs(cand) - 1;
                        cand &= cand - 1u;
                        const float2 tl = kLean ? __ldg(g_loc + e) : ld_shared_f2(lp + 8 * e);
                        dead |= mask_skips<kMask>(qm, tl.x, tl.y, eps) ? (1u << e) : 0u;
                      }
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) x[e] = (dead >> e) & 1u ? 0u : x[e];
                  } else {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                      const int4 l2 = kLean ? __ldg(reinterpret_cast<const int4*>(g_loc + e)) : ld_shared_v4(lp + 8 * e);
                      if (mask_skips<kMask>(qm, __int_as_float(l2.x), __int_as_float(l2.y), eps)) x[e] = 0;
                      if (mask_skips<kMask>(qm, __int_as_float(l2.z), __int_as_float(l2.w), eps)) x[e + 1] = 0;
                    }
                  }
                }
                // a lane that was not live keeps its bound (below its best; its row may not have been masked)
                wm[k] = mine ? max32(x) : min(wm[k], st.best - 1);
              }
            }
          }
          if (batch.has_exclude) {                                 // runner-up search (launch-uniform): the excluded target scores 0
            const uint64_t* q_exclude = load_exclude_ptr(args);
            if (q_exclude != nullptr)
              exclude_columns<static_cast<int>(kWinPerWarp)>(v, wm, st.best, j * kWinPerTile + c_begin,
                                                             row < load_nq(args) ? excluded_row(q_exclude, row) : 0xFFFFFFFFu);
          }
          // bookkeeping (branch-free body)
          if (kColMask) {
#pragma unroll
            for (uint32_t k = 0; k < kWinPerWarp; ++k) {
              if (__any_sync(0xffffffffu, wm[k] >= st.best))     // which columns attained it: only for windows that matter
                win_update(st, wm[k], j * kWinPerTile + c_begin + k, eq_mask32(v[k], wm[k]));
            }
          } else {
#pragma unroll
            for (uint32_t k = 0; k < kWinPerWarp; ++k) win_update(st, wm[k], j * kWinPerTile + c_begin + k, 0u);
          }
        }
        if (kCols) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_empty[cs]);
        }
        if (kTiming) { const long long e2 = clock64(); te_wait += e1 - e0; te_work += e2 - e1; }
      }
      long long r0 = 0;
      if (kTiming) r0 = clock64();
      // ---------------- hand the window states of this unit to the resolver warps ----------------
      const uint32_t uparity = uiter & 1;
      mbar_wait(&x_empty[uparity], ((uiter >> 1) & 1) ^ 1);
      xch[(uparity * 2 + half) * kBM + quarter * 32 + lane] = st;
      if (!u.last) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_full[uparity]);
      } else {
        // last unit of this CTA: nothing left to overlap with, the epilogue warps turn it into keys themselves
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (!kColMask) {
          resolve_rows_windows<kMask, 2>(args, xch + (uparity * 2 + 0) * kBM, xch + (uparity * 2 + 1) * kBM, u.qt * kBM, u.slot, (warp - 4) * 16, 16, lane);
        } else if (half == 0) {
          const uint32_t r = quarter * 32 + lane;
          resolve_row(args, xch[(uparity * 2 + 0) * kBM + r], xch[(uparity * 2 + 1) * kBM + r], u.qt * kBM, u.slot, r, kWOffset);
        }
        fill_unused_slots(args, u.qt * kBM, u.slot, u.fill_end, (warp - 4) * 16, 16, lane);
      }
      ++uiter;
      if (kTiming) te_resolve += clock64() - r0;
    }
    if (kTiming && lane == 0) {
      const uint32_t e = warp - 4;
      if (e == 0 || e == 5) {
        const uint32_t o = blockIdx.x * 16 + (e == 0 ? 4 : 8);
        batch.timing[o + 0] = te_wait; batch.timing[o + 1] = te_work; batch.timing[o + 2] = te_resolve; batch.timing[o + 3] = clock64() - te_total0;
        // windows refined in the rare path (no-fold bounds / masks) and the time spent waiting for the column ring
        if (e == 0) { batch.timing[blockIdx.x * 16 + 14] = te_live; batch.timing[blockIdx.x * 16 + 15] = te_cwait; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

template <int kMask, bool kFold, bool kColMask, bool kTiming, bool kLean = false>
static cudaError_t launch_one(cudaStream_t st, const SmaxBatch& b) {
  auto kern = smax_kernel<kMask, kFold, kColMask, kTiming, kLean>;
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

template <int kMask>
static cudaError_t launch_mask(cudaStream_t st, const SmaxBatch& b, bool fold, bool colmask, bool lean) {
  if constexpr (kMask != SSRLCV_B200_MASK_NONE) {
    if (lean && fold && colmask && !b.timing) return launch_one<kMask, true, true, false, true>(st, b);
  }
  if (b.timing) {   // diagnostics build (tools/kerneltiming.py); fold variants of the unconstrained and DoubleConstrained kernels
    if (kMask != SSRLCV_B200_MASK_FUNDAMENTAL && fold)
      return colmask ? launch_one<kMask == SSRLCV_B200_MASK_FUNDAMENTAL ? 0 : kMask, true, true, true>(st, b)
                     : launch_one<kMask == SSRLCV_B200_MASK_FUNDAMENTAL ? 0 : kMask, true, false, true>(st, b);
    if (kMask == SSRLCV_B200_MASK_NONE && !fold)
      return colmask ? launch_one<SSRLCV_B200_MASK_NONE, false, true, true>(st, b) : launch_one<SSRLCV_B200_MASK_NONE, false, false, true>(st, b);
  }
  if (fold) return colmask ? launch_one<kMask, true, true, false>(st, b) : launch_one<kMask, true, false, false>(st, b);
  return colmask ? launch_one<kMask, false, true, false>(st, b) : launch_one<kMask, false, false, false>(st, b);
}

// one launch over all jobs of `b` (same mask kind, fold variant and bookkeeping regime); b.jobs is device memory
cudaError_t launch_smax(cudaStream_t st, const SmaxBatch& b, int mask_kind, bool fold, bool colmask, bool lean) {
  switch (mask_kind) {
    case SSRLCV_B200_MASK_NONE: return launch_mask<SSRLCV_B200_MASK_NONE>(st, b, fold, colmask, false);
    case SSRLCV_B200_MASK_FUNDAMENTAL: return launch_mask<SSRLCV_B200_MASK_FUNDAMENTAL>(st, b, fold, colmask, lean);
    default: return launch_mask<SSRLCV_B200_MASK_DOUBLE>(st, b, fold, colmask, lean);
  }
}

}  // namespace b200
