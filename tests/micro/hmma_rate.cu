// Micro-benchmark (B200): issue rate of the legacy warp-level mma.sync.m16n8k16 (f16 x f16 -> f32) per SM as a
// function of resident warps, next to packed fp32x2 FMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o hmma_rate.bin hmma_rate.cu ; prints cycles per instruction per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void hmma_kernel(int iters, float* out, long long* cyc) {
  float d[NACC][4];
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) hmma(d[i], a, b);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ffma2_kernel(int iters, float* out, long long* cyc) {
  unsigned long long d[16];
  unsigned long long a = 0x3f8000003f800000ull + threadIdx.x, b = 0x3f0000003f000000ull;
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(a), "l"(b));
  }
  __syncthreads();
  const long long t1 = clock64();
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {4, 8, 12, 16, 24, 32}) {
    hmma_kernel<8><<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    hmma_kernel<8><<<148, warps * 32>>>(iters, out, cyc);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double n = double(iters) * 8 * warps;
    printf("HMMA.16816.F32 warps/SM=%2d: %.3f cycles per warp-instruction per SM  (%.0f MAC/clk/SM)\n", warps, mx / n,
           2048.0 * n / mx);
  }
  for (int warps : {4, 8, 16, 32}) {
    ffma2_kernel<<<148, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    ffma2_kernel<<<148, warps * 32>>>(iters, out, cyc);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double n = double(iters) * 16 * warps;
    printf("FFMA2 warps/SM=%2d: %.3f cycles per warp-instruction per SM  (%.0f FMA/clk/SM)\n", warps, mx / n, 64.0 * n / mx);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
