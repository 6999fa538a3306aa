// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M = 128, K = 16) for the operand modes the attention
// kernels use.  One CTA, one issuing lane, operands are whatever is in shared memory (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../uniception_b200/csrc -o mma_bench mma_bench.cu
#include <cstdio>
#include "common.cuh"
using namespace uc;

namespace uc {
void set_error(const char*, ...) {}
int check_launch(const char*) { return 0; }
}  // namespace uc


template <int MODE>
__device__ __forceinline__ void issue(int it, uint32_t tm, uint32_t sA, uint32_t sB, uint32_t sC, uint32_t sD) {
  const int k = it & 3;
  if (MODE == 0) umma_ss(tm, umma_desc_kmajor(sA) + k * 2, umma_desc_kmajor(sB) + k * 2, umma_idesc_bf16(128, 128, 0, 0), 1);
  if (MODE == 1) umma_ss(tm, umma_desc_kmajor(sA) + k * 2, umma_desc_kmajor(sB) + k * 2, umma_idesc_bf16(128, 64, 0, 0), 1);
  if (MODE == 2) umma_ss(tm, umma_desc_kmajor(sA) + k * 2, umma_desc_mnmajor(sB, 8192) + k * 128, umma_idesc_bf16(128, 64, 0, 1), 1);
  if (MODE == 3) umma_ss(tm, umma_desc_mnmajor(sA, 16384) + k * 128, umma_desc_mnmajor(sB, 8192) + k * 128, umma_idesc_bf16(128, 64, 1, 1), 1);
  if (MODE == 4) umma_ts(tm, tm + 256 + k * 8, umma_desc_mnmajor(sB, 8192) + k * 128, umma_idesc_bf16(128, 64, 0, 1), 1);
  if (MODE == 5) umma_ss(tm, umma_desc_kmajor(sA) + k * 2, umma_desc_kmajor(sC) + k * 2, umma_idesc_bf16(128, 256, 0, 0), 1);
  if (MODE == 6) umma_ss(tm + (it & 1) * 64, umma_desc_kmajor((it & 1) ? sD : sA) + k * 2, umma_desc_kmajor(sB) + k * 2, umma_idesc_bf16(128, 64, 0, 0), 1);
  if (MODE == 7) umma_ts(tm, tm + 256 + k * 8, umma_desc_kmajor(sB) + k * 2, umma_idesc_bf16(128, 128, 0, 0), 1);
}

template <int MODE>
__device__ __noinline__ void run(long long* out, int iters, uint32_t tm, uint32_t sA, uint32_t sB, uint32_t sC, uint32_t sD, uint32_t bar,
                                 uint32_t& ph) {
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        for (int o = 0; o < iters; o += 32) {
#pragma unroll
          for (int it = 0; it < 32; ++it) issue<MODE>(it, tm, sA, sB, sC, sD);
        }
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, ph);
      ph ^= 1u;
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[MODE] = t1 - t0;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 160 * 1024, slot = bar + 64;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tm) : "r"(slot));
  const uint32_t sA = base, sB = base + 32768, sC = base + 65536, sD = base + 98304;
  // mode: 0 SS N=128 K/K | 1 SS N=64 K/K | 2 SS N=64 A K-major, B MN | 3 SS N=64 A MN, B MN | 4 TS N=64 B MN | 5 SS N=256 K/K
  //       6 two interleaved accumulators of mode 1 | 7 TS N=128 B K-major
  uint32_t ph = 0;
  run<0>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<1>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<2>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<3>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<4>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<5>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<6>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  run<7>(out, iters, tm, sA, sB, sC, sD, bar, ph);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaMemset(d, 0, 64);
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 512;
  bench<<<1, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  const char* names[8] = {"SS 128x128x16 K/K", "SS 128x64x16 K/K", "SS 128x64x16 A K-major, B MN-major", "SS 128x64x16 A MN-major, B MN-major",
                          "TS 128x64x16 B MN-major", "SS 128x256x16 K/K", "SS 128x64x16 K/K, two accumulators interleaved", "TS 128x128x16 B K-major"};
  for (int m = 0; m < 8; ++m) printf("%-48s %7.1f clk / MMA  (%lld clk for %d)\n", names[m], double(h[m]) / iters, h[m], iters);
  return 0;
}
