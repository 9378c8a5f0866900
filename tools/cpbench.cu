// cpbench.cu -- can the per-target constant W_b be put into the TMEM accumulator by tcgen05.cp (smem -> TMEM, issued
// by the MMA thread, pipelined with tcgen05.mma) instead of a fifth "fold" MMA?
//   1. layout probe: one tcgen05.cp.128x256b from a word-index pattern, dump of the TMEM rectangle for several (LBO, SBO)
//   2. rate: per tile 32 x tcgen05.cp.128x256b (256 columns) + 4 x tcgen05.mma (accumulate), against 4 and 5 MMAs
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I ssrlcv_b200/csrc tools/cpbench.cu -o tools/cpbench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t gtimer() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ uint64_t desc_raw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(swz) << 61;
  return d;
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void cp_128x128b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void cp_32x128b_x4(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

// ---- 1. layout probe ------------------------------------------------------------------------------
// shape: 0 = 128x256b, 1 = 128x128b, 2 = 32x128b.warpx4
__global__ void __launch_bounds__(128, 1) probe_kernel(int shape, uint32_t lbo, uint32_t sbo, uint32_t* out /*128 x 8*/) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;   // word index
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc<512>(&tptr); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 1 && elect_one()) {
    const uint64_t d = desc_raw(smem_u32(smem), lbo, sbo, 0);
    if (shape == 0) cp_128x256b(tbase, d);
    else if (shape == 1) cp_128x128b(tbase, d);
    else cp_32x128b_x4(tbase, d);
    mma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tbase + ((warp * 32u) << 16), v);
  tmem_ld_wait();
  for (int c = 0; c < 8; ++c) out[(warp * 32 + lane) * 8 + c] = v[c];
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ---- 2. rate ----------------------------------------------------------------------------------------
// MODE 0: 4 MMA; 1: 5 MMA (fold); 2: 32 x cp.128x256b + 4 MMA (all accumulating); 3: 64 x cp.32x128b.warpx4 + 4 MMA;
// 4: 64 x cp.128x128b + 4 MMA
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out_cycles, unsigned long long* out_ns, uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768 + 8192 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc<512>(&tptr); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_i8(128, 256, 0, 0);
    const uint64_t a = make_smem_desc(smem_u32(smem), 1024, kSwizzle128B);
    const uint64_t b = make_smem_desc(smem_u32(smem + 16384), 1024, kSwizzle128B);
    const uint64_t ae = make_smem_desc(smem_u32(smem + 16384 + 32768 + 8192), 256, kSwizzle32B);
    const uint64_t be = make_smem_desc(smem_u32(smem + 16384 + 32768), 256, kSwizzle32B);
    const uint32_t wbase = smem_u32(smem + 16384 + 32768);   // 8 KB: the replicated W tile
    const long long c0 = clock64(); const uint64_t t0 = gtimer();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tbase + (it & 1) * 256;
      if (MODE == 2) {
#pragma unroll
        for (int c = 0; c < 32; ++c) cp_128x256b(d + c * 8, desc_raw(wbase + c * 256, lbo, sbo, 0));
      } else if (MODE == 3) {
#pragma unroll
        for (int c = 0; c < 64; ++c) cp_32x128b_x4(d + c * 4, desc_raw(wbase + c * 128, lbo, sbo, 0));
      } else if (MODE == 4) {
#pragma unroll
        for (int c = 0; c < 64; ++c) cp_128x128b(d + c * 4, desc_raw(wbase + c * 128, lbo, sbo, 0));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_i8_ss(d, a + 2 * k, b + 2 * k, idesc, (MODE >= 2) ? 1 : k);
      if (MODE == 1) mma_i8_ss(d, ae, be, idesc, 1);
      if ((it & 15) == 15) { mma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1; }
    }
    mma_commit(&bar); mbar_wait(&bar, ph);
    const long long c1 = clock64(); const uint64_t t1 = gtimer();
    if (blockIdx.x == 0) { out_cycles[0] = c1 - c0; out_ns[0] = t1 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}


// ---- 3. TMEM store contention ---------------------------------------------------------------------------
// Warp 1 issues 4 MMAs per tile flat out; warps 4-11 (two per lane quarter, one per 128-column half) free-run
// "4 x tcgen05.ld.x32 [+ 4 x tcgen05.st.x32]" over their slice of the accumulators.  Reports MMA cycles per tile and
// how many slice passes every epilogue warp completed per MMA tile (>= 1 means the epilogue keeps up).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
        "r"(taddr)
      : "memory");
}

template <int MODE>   // 0: ld only, 1: ld + st, 2: st only, 3: no epilogue traffic
__global__ void __launch_bounds__(384, 1) st_rate_kernel(int iters, long long* out_cycles, unsigned long long* out_passes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile uint32_t done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768 + 8192 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc<512>(&tptr); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tptr;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_i8(128, 256, 0, 0);
      const uint64_t a = make_smem_desc(smem_u32(smem), 1024, kSwizzle128B);
      const uint64_t b = make_smem_desc(smem_u32(smem + 16384), 1024, kSwizzle128B);
      const long long c0 = clock64();
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tbase + (it & 1) * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_i8_ss(d, a + 2 * k, b + 2 * k, idesc, 1);
        if ((it & 15) == 15) { mma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1; }
      }
      mma_commit(&bar); mbar_wait(&bar, ph);
      const long long c1 = clock64();
      done = 1;
      if (blockIdx.x == 0) out_cycles[0] = c1 - c0;
    }
  } else if (warp >= 4) {
    const uint32_t quarter = warp & 3, half = (warp - 4) >> 2;
    const uint32_t base = tbase + ((quarter * 32u) << 16) + half * 128;
    unsigned long long passes = 0;
    uint32_t acc = 0;
    uint32_t w[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) w[e] = e + warp;
    if (MODE != 3) {
      while (!done) {
        const uint32_t t = base + (passes & 1) * 256;
        uint32_t v[4][32];
        if (MODE != 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_ld_32x32(t + k * 32, v[k]);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 32; e += 4) acc ^= v[k][e];
        }
        if (MODE != 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tmem_st_32x32(t + k * 32, w);
          tmem_st_wait();
        }
        ++passes;
      }
    }
    if (blockIdx.x == 0 && lane == 0) out_passes[warp - 4] = passes + (acc == 0x12345u);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s sms %d\n", prop.name, sms);
  long long* d_cyc; unsigned long long* d_ns; uint32_t* d_out;
  CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_ns, 64)); CK(cudaMalloc(&d_out, 128 * 8 * 4));
  const int smem = 16384 + 32768 + 8192 + 4096 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  static uint32_t h[128 * 8];
  const uint32_t cfgs[][3] = {{0, 128, 256}, {0, 128, 0}, {0, 256, 128}, {0, 0, 128}, {0, 16, 0}, {0, 0, 0},
                              {1, 128, 128}, {1, 128, 0}, {1, 0, 0}, {2, 128, 128}, {2, 128, 0}, {2, 0, 0}};
  for (auto& c : cfgs) {
    CK(cudaMemset(d_out, 0xff, sizeof(h)));
    probe_kernel<<<1, 128, smem>>>(c[0], c[1], c[2], d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe shape %u lbo %u sbo %u: %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    printf("probe shape %u lbo %u sbo %u (smem word index at TMEM lane r, columns 0..7)\n", c[0], c[1], c[2]);
    for (int r : {0, 1, 2, 7, 8, 9, 16, 31, 32, 33, 64, 127}) {
      printf("  lane %3d:", r);
      for (int k = 0; k < 8; ++k) printf(" %5d", (int)h[r * 8 + k]);
      printf("\n");
    }
  }
  auto run = [&](auto kern, const char* name, uint32_t lbo, uint32_t sbo) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 20000;
    long long cyc; unsigned long long ns;
    for (int rep = 0; rep < 2; ++rep) {
      kern<<<sms, 128, smem>>>(iters, d_cyc, d_ns, lbo, sbo);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&ns, d_ns, 8, cudaMemcpyDeviceToHost));
    printf("%-34s lbo %3u sbo %3u: %.1f cyc/tile (SM clock %.0f MHz while running)\n", name, lbo, sbo, (double)cyc / iters, (double)cyc / ns * 1e3);
  };
  run(rate_kernel<0>, "4 MMA", 0, 0);
  run(rate_kernel<1>, "5 MMA (fold)", 0, 0);
  run(rate_kernel<2>, "32 cp.128x256b + 4 MMA", 128, 0);
  run(rate_kernel<2>, "32 cp.128x256b + 4 MMA", 128, 256);
  run(rate_kernel<3>, "64 cp.32x128b.warpx4 + 4 MMA", 128, 0);
  run(rate_kernel<3>, "64 cp.32x128b.warpx4 + 4 MMA", 128, 128);
  run(rate_kernel<4>, "64 cp.128x128b + 4 MMA", 128, 0);
  run(rate_kernel<4>, "64 cp.128x128b + 4 MMA", 128, 128);
  {
    unsigned long long* d_p; CK(cudaMalloc(&d_p, 64));
    auto run2 = [&](auto kern, const char* name) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int iters = 20000;
      long long cyc; unsigned long long p[8];
      for (int rep = 0; rep < 2; ++rep) { kern<<<sms, 384, smem>>>(iters, d_cyc, d_p); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(p, d_p, 64, cudaMemcpyDeviceToHost));
      printf("%-30s: %.1f cyc/tile; slice passes per tile per warp:", name, (double)cyc / iters);
      for (int i = 0; i < 8; ++i) printf(" %.2f", (double)p[i] / iters);
      printf("\n");
    };
    run2(st_rate_kernel<3>, "4 MMA, idle epilogue");
    run2(st_rate_kernel<0>, "4 MMA + 8 warps ld");
    run2(st_rate_kernel<1>, "4 MMA + 8 warps ld+st");
    run2(st_rate_kernel<2>, "4 MMA + 8 warps st");
  }
  return 0;
}
