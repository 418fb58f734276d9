// Microbenchmark: tcgen05.mma issue/execute cost vs N, operand source and accumulator dependence.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I facet_graph_convolution_b200/csrc -I include tests/micro/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace fgc;

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: TS (A tmem), B K-major smem; mode 1: SS A K-major, B K-major; mode 2: SS A K-major, B MN-major
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int nacc, int count, long long* out, int use_elect) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t idesc = (1u << 4) | (mode == 2 ? (1u << 16) : 0u) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (use_elect ? tc::elect_one() : (lane == 0)) {
        const uint32_t amask = nacc - 1, astride = (nacc > 1) ? N : 0;
        const uint64_t bd0 = (mode == 2) ? desc_mn(sb + 32768, 1024, 1024) : tc::smem_desc_k_sw128(sb + 32768);
        const uint64_t ad0 = tc::smem_desc_k_sw128(sb);
        const uint64_t bstep = (mode == 2) ? (2048 >> 4) : (32 >> 4);
#pragma unroll 4
        for (int i = 0; i < count; ++i) {
          const uint32_t d = tmem + 256 + (i & amask) * astride;
          const uint32_t ks = i & 3;
          if (mode == 0) tc::mma_f16_ts(d, tmem + ks * 8, bd0 + ks * bstep, idesc, i >= nacc ? 1u : 0u);
          else tc::mma_f16_ss(d, ad0 + ks * 2, bd0 + ks * bstep, idesc, i >= nacc ? 1u : 0u);
        }
        tc::tc_commit(&bar);
      }
      __syncwarp();
      long long t1 = clock64();
      tc::mbar_wait(&bar, rep & 1);
      long long t2 = clock64();
      if (lane == 0 && rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

__global__ void spin(long long cycles) { const long long t0 = clock64(); while (clock64() - t0 < cycles) {} }

int main() {
  long long* d; cudaMalloc(&d, 16);
  spin<<<148, 128>>>(400000000ll);   // ~0.2 s: let the clocks ramp before the one-CTA measurements
  cudaDeviceSynchronize();
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const char* names[] = {"TS  B K-major ", "SS  B K-major ", "SS  B MN-major"};
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {32, 64, 128, 256})
      for (int nacc : {1, 10, 4}) {
        const int use_elect = nacc != 10;
        if (nacc == 10) nacc = 1;
        if (nacc * N > 256) continue;
        bench<<<1, 128, 96 * 1024>>>(mode, N, nacc, 64, d, use_elect);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%s N=%3d accumulators=%d elect=%d: issue %6.1f cyc/mma, complete %6.1f cyc/mma  (%s)\n", names[mode], N, nacc, use_elect, h[0] / 64.0, h[1] / 64.0, cudaGetErrorString(e));
      }
  return 0;
}
