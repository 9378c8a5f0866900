// microbench.cu -- measures the three B200 rates the matcher's design rests on (run on the GPU box):
//   1. tcgen05.mma kind::i8 issue rate (M128 N256 K32 and N128), no epilogue
//   2. tcgen05.ld 32x32b.x32 bandwidth with 4 and 8 reader warps
//   3. integer max3 (VIMNMX3) and IMAD throughput per SM
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I ssrlcv_b200/csrc tools/microbench.cu -o tools/microbench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"
using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t gtimer() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---- 1. MMA issue rate -------------------------------------------------------------------------
template <int N, int KBLOCKS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* out_cycles, unsigned long long* out_ns) {
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
    const uint32_t idesc = make_idesc_i8(128, N, 0, 0);
    const uint64_t a = make_smem_desc(smem_u32(smem), 1024, kSwizzle128B);
    const uint64_t b = make_smem_desc(smem_u32(smem + 16384), 1024, kSwizzle128B);
    const uint64_t ae = make_smem_desc(smem_u32(smem + 16384 + 32768 + 8192), 256, kSwizzle32B);
    const uint64_t be = make_smem_desc(smem_u32(smem + 16384 + 32768), 256, kSwizzle32B);
    const long long c0 = clock64(); const uint64_t t0 = gtimer();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tbase + (it & 1) * 256;
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_i8_ss(d, a + 2 * k, b + 2 * k, idesc, k);
      if (KBLOCKS == 5) mma_i8_ss(d, ae, be, idesc, 1);
      if ((it & 15) == 15) { mma_commit(&bar); mbar_wait(&bar, ph); ph ^= 1; }
    }
    mma_commit(&bar); mbar_wait(&bar, ph);
    const long long c1 = clock64(); const uint64_t t1 = gtimer();
    if (blockIdx.x == 0) { out_cycles[0] = c1 - c0; out_ns[0] = t1 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ---- 2. tcgen05.ld bandwidth -------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) tmem_ld_kernel(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<512>(&tptr); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tptr;
  const uint32_t lane_addr = ((warp & 3) * 32u) << 16;
  const uint32_t col0 = (WARPS == 8 && warp >= 4) ? 256 : 0;
  const int ncol = (WARPS == 8) ? 256 : 512;
  uint32_t acc = 0;
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[2][32];
    tmem_ld_32x32(tbase + lane_addr + col0, v[0]);
    for (int c = 0; c < ncol / 32; ++c) {
      tmem_ld_wait();
      if (c + 1 < ncol / 32) tmem_ld_32x32(tbase + lane_addr + col0 + (c + 1) * 32, v[(c + 1) & 1]);
#pragma unroll
      for (int e = 0; e < 32; e += 8) acc ^= v[c & 1][e];
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out_cycles[0] = c1 - c0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

// ---- 3. ALU rates ------------------------------------------------------------------------------
template <int MODE>   // 0: max3 chain x8 independent, 1: IMAD x8 independent, 2: both interleaved
__global__ void __launch_bounds__(512, 1) alu_kernel(int iters, long long* out_cycles, int* sink, int seed) {
  int a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; b[i] = seed * 3 + i; }
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) a[i] = max(max(a[i], b[i] ^ it), b[(i + 1) & 7]);
      if (MODE == 1 || MODE == 2) b[i] = b[i] * seed + a[(i + 3) & 7];
    }
  }
  __syncthreads();
  const long long c1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out_cycles[0] = c1 - c0;
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + b[i];
  if (s == 0x7fffffff) sink[0] = s;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s sms %d cc %d.%d clock %d kHz\n", prop.name, sms, prop.major, prop.minor, prop.clockRate);
  long long* d_cyc; unsigned long long* d_ns; uint32_t* d_sink;
  CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_ns, 64)); CK(cudaMalloc(&d_sink, 64));
  long long cyc; unsigned long long ns;
  const int smem = 16384 + 32768 + 8192 + 4096 + 1024;

  {  // MMA rates
    auto run = [&](auto kern, const char* name, int n, int kb) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      const int iters = 20000;
      for (int rep = 0; rep < 2; ++rep) {
        kern<<<sms, 128, smem>>>(iters, d_cyc, d_ns);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&ns, d_ns, 8, cudaMemcpyDeviceToHost));
      const double mmas = (double)iters * kb;
      const double ops = mmas * 128.0 * n * 32 * 2;
      printf("MMA %-14s: %.1f cyc/MMA(K32), %.1f cyc/tile, per-SM %.0f ops/cyc, chip %.1f TOPS (SM clock %.0f MHz while running)\n", name,
             cyc / mmas, (double)cyc / iters, ops / cyc, ops * sms / ns / 1e3, (double)cyc / ns * 1e3);
    };
    run(mma_rate_kernel<256, 4>, "N256 k4", 256, 4);
    run(mma_rate_kernel<256, 5>, "N256 k4+fold", 256, 5);
    run(mma_rate_kernel<128, 4>, "N128 k4", 128, 4);
  }
  {  // TMEM load
    const int iters = 2000;
    tmem_ld_kernel<4><<<sms, 128>>>(iters, d_cyc, d_sink); CK(cudaDeviceSynchronize());
    tmem_ld_kernel<4><<<sms, 128>>>(iters, d_cyc, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("TMEM ld 4 warps: %.1f B/cyc/SM (%.1f cyc per 128x512 sweep)\n", (double)iters * 128 * 512 * 4 / cyc, (double)cyc / iters);
    tmem_ld_kernel<8><<<sms, 256>>>(iters, d_cyc, d_sink); CK(cudaDeviceSynchronize());
    tmem_ld_kernel<8><<<sms, 256>>>(iters, d_cyc, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("TMEM ld 8 warps: %.1f B/cyc/SM (%.1f cyc per 128x512 sweep)\n", (double)iters * 128 * 512 * 4 / cyc, (double)cyc / iters);
  }
  {  // ALU
    const int iters = 20000;
    alu_kernel<0><<<sms, 512>>>(iters, d_cyc, (int*)d_sink, 3); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("ALU max3 (2 min/max per op written as nested max): %.1f thread-ops/cyc/SM\n", (double)iters * 8 * 512 / cyc);
    alu_kernel<1><<<sms, 512>>>(iters, d_cyc, (int*)d_sink, 3); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("ALU imad: %.1f thread-ops/cyc/SM\n", (double)iters * 8 * 512 / cyc);
    alu_kernel<2><<<sms, 512>>>(iters, d_cyc, (int*)d_sink, 3); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("ALU max3+imad pairs: %.1f pair-ops/cyc/SM\n", (double)iters * 8 * 512 / cyc);
  }
  return 0;
}
