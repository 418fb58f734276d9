// Per-facet linear layers (reference Code/model.py:763-769 custom_lin) and the fused regression
// head lrelu(x@W1+b1)@W2+b2 (reference Code/model.py:936-941) that never materialises the
// 1024-wide hidden activation.
#include "conv_common.cuh"
#include "conv_launch.cuh"

namespace fgc {

// ------------------------------------------------------------------ y = act(x @ W + b)
__global__ void __launch_bounds__(kThreads)
lin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
               float* __restrict__ y, int64_t rows, int Cin, int Cout, int act, float alpha) {
  extern __shared__ __align__(16) float sm[];
  const int lda = (Cin + 3) & ~3;
  float* A = sm;                       // [32][lda]
  float* Bs = A + kTileFacets * lda;   // [32][128]
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileFacets;
    __syncthreads();
    for (int e = threadIdx.x; e < kTileFacets * lda; e += kThreads) {
      const int f = e / lda, c = e % lda;
      A[e] = (r0 + f < rows && c < Cin) ? __ldg(x + (r0 + f) * Cin + c) : 0.f;
    }
    for (int o0 = 0; o0 < Cout; o0 += 128) {
      const int ncols = min(128, Cout - o0);
      const TileGemmMap mp(ncols);
      float acc[4][4];
      tile_gemm(A, lda, Cin, W, Cout, o0, ncols, Bs, mp, acc);
      if (mp.ty < mp.TY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = mp.ty + mp.TY * i;
          const int64_t r = r0 + f;
          if (i < mp.RF && f < kTileFacets && r < rows) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = o0 + 4 * mp.tx + j;
              if (o < Cout) {
                float v = acc[i][j] + __ldg(b + o);
                if (act == FGC_ACT_LRELU) v = lrelu_f(v, alpha);
                y[r * Cout + o] = v;
              }
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ y = act(x @ W + b), Cout <= 4 (the 1024 -> 3 output
// layer of the regression head in training): warp per row, the lanes split Cin in float4 steps, one shuffle tree per
// output in a fixed order; pure bandwidth (the row is read once, coalesced) where the tile GEMM left 31 of 32 columns
// of its thread map idle (0.59 -> 0.0x ms per 32 768 rows of 1024)
template <int NO>
__global__ void __launch_bounds__(256)
lin_fwd_narrow_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                      float* __restrict__ y, int64_t rows, int Cin, int act, float alpha) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = w0; r < rows; r += nw) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * Cin);
    float acc[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) acc[o] = 0.f;
    for (int k4 = lane; k4 < Cin / 4; k4 += 32) {
      const float4 v = __ldg(xr + k4);
      const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[o] = fmaf(xv[j], __ldg(W + static_cast<int64_t>(4 * k4 + j) * NO + o), acc[o]);
    }
#pragma unroll
    for (int o = 0; o < NO; ++o) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], sft);
    }
    if (lane < NO) {
      float v = acc[0];
#pragma unroll
      for (int o = 1; o < NO; ++o)
        if (lane == o) v = acc[o];
      v += __ldg(b + lane);
      if (act == FGC_ACT_LRELU) v = lrelu_f(v, alpha);
      y[r * NO + lane] = v;
    }
  }
}

// ------------------------------------------------------------------ fused head
constexpr int kHeadMaxOut = 4;
__global__ void __launch_bounds__(kThreads)
mlp_head_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                const float* __restrict__ b1, const float* __restrict__ W2,
                const float* __restrict__ b2, float* __restrict__ y, int64_t rows, int Cin, int H,
                int Cout, float alpha) {
  extern __shared__ __align__(16) float sm[];
  const int lda = (Cin + 3) & ~3;
  float* A = sm;                                 // [32][lda]
  float* Bs = A + kTileFacets * lda;             // [32][128]
  float* part = Bs + kChunkK * 128;              // [32 facets][32 tx][4]
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileFacets;
    __syncthreads();
    for (int e = threadIdx.x; e < kTileFacets * lda; e += kThreads) {
      const int f = e / lda, c = e % lda;
      A[e] = (r0 + f < rows && c < Cin) ? __ldg(x + (r0 + f) * Cin + c) : 0.f;
    }
    float yacc[4][kHeadMaxOut];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < kHeadMaxOut; ++j) yacc[i][j] = 0.f;
    // hidden columns are always processed in full 128-wide blocks => fixed thread map (TX=32,TY=8,RF=4)
    const TileGemmMap mp(128);
    for (int h0 = 0; h0 < H; h0 += 128) {
      const int ncols = min(128, H - h0);
      float acc[4][4];
      tile_gemm(A, lda, Cin, W1, H, h0, ncols, Bs, mp, acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int h = h0 + 4 * mp.tx + j;
        if (h < H) {
          const float bb = __ldg(b1 + h);
          float w2[kHeadMaxOut];
#pragma unroll
          for (int o = 0; o < kHeadMaxOut; ++o) w2[o] = (o < Cout) ? __ldg(W2 + h * Cout + o) : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float hv = lrelu_f(acc[i][j] + bb, alpha);
#pragma unroll
            for (int o = 0; o < kHeadMaxOut; ++o) yacc[i][o] = fmaf(hv, w2[o], yacc[i][o]);
          }
        }
      }
    }
    // fixed-order reduction over the 32 column groups of every facet
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = mp.ty + mp.TY * i;
#pragma unroll
      for (int o = 0; o < kHeadMaxOut; ++o) part[(f * 32 + mp.tx) * kHeadMaxOut + o] = yacc[i][o];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTileFacets * Cout; e += kThreads) {
      const int f = e / Cout, o = e % Cout;
      if (r0 + f < rows) {
        float a = 0.f;
        for (int t = 0; t < 32; ++t) a += part[(f * 32 + t) * kHeadMaxOut + o];
        y[(r0 + f) * Cout + o] = a + __ldg(b2 + o);
      }
    }
  }
}

// ------------------------------------------------------------------ backward
__global__ void transpose2d_kernel(const float* __restrict__ W, float* __restrict__ Wt, int R, int C) {
  const int total = R * C;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int r = e / C, c = e % C;
    Wt[static_cast<int64_t>(c) * R + r] = W[e];
  }
}

// gx = gy @ W^T for a WIDE layer with a narrow input (Cin = 32 outputs of this kernel, Cout = hundreds of columns to reduce
// over: the 32 -> 1024 layer of the regression head in training).  A warp takes eight rows, lane i owns output i of each:
// per four columns one coalesced 128-byte read of each of the four Wt rows and one broadcast 16-byte read of gy per row
// (a lane-per-column split made every Wt load touch 32 lines: 0.49 ms; the tile GEMM needed a 128 KB row tile per CTA for
// this shape: 0.73 ms).  Fixed summation order (k ascending).
__global__ void __launch_bounds__(256)
lin_bwd_x_wide_kernel(const float* __restrict__ gy, const float* __restrict__ Wt, float* __restrict__ gx, int64_t rows,
                      int Cout) {
  constexpr int R = 8;
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r0 = R * w0; r0 < rows; r0 += R * nw) {
    float acc[R];
    const float4* g[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      acc[i] = 0.f;
      g[i] = reinterpret_cast<const float4*>(gy + (r0 + i < rows ? r0 + i : r0) * Cout);
    }
#pragma unroll 2
    for (int k4 = 0; k4 < Cout / 4; ++k4) {
      const float* wr = Wt + static_cast<int64_t>(4 * k4) * 32 + lane;
      const float w0v = __ldg(wr), w1v = __ldg(wr + 32), w2v = __ldg(wr + 64), w3v = __ldg(wr + 96);
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const float4 v = __ldg(g[i] + k4);
        acc[i] = fmaf(v.x, w0v, acc[i]);
        acc[i] = fmaf(v.y, w1v, acc[i]);
        acc[i] = fmaf(v.z, w2v, acc[i]);
        acc[i] = fmaf(v.w, w3v, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
      if (r0 + i < rows) gx[(r0 + i) * 32 + lane] = acc[i];
  }
}

// gx = gy @ W^T   (A = gy tile, B = Wt[Cout][Cin])
__global__ void __launch_bounds__(kThreads)
lin_bwd_x_kernel(const float* __restrict__ gy, const float* __restrict__ Wt, float* __restrict__ gx,
                 int64_t rows, int Cin, int Cout) {
  extern __shared__ __align__(16) float sm[];
  const int lda = (Cout + 3) & ~3;
  float* A = sm;
  float* Bs = A + kTileFacets * lda;
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileFacets;
    __syncthreads();
    for (int e = threadIdx.x; e < kTileFacets * lda; e += kThreads) {
      const int f = e / lda, c = e % lda;
      A[e] = (r0 + f < rows && c < Cout) ? __ldg(gy + (r0 + f) * Cout + c) : 0.f;
    }
    for (int c0 = 0; c0 < Cin; c0 += 128) {
      const int ncols = min(128, Cin - c0);
      const TileGemmMap mp(ncols);
      float acc[4][4];
      tile_gemm(A, lda, Cout, Wt, Cin, c0, ncols, Bs, mp, acc);
      if (mp.ty < mp.TY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = mp.ty + mp.TY * i;
          const int64_t r = r0 + f;
          if (i < mp.RF && f < kTileFacets && r < rows) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + 4 * mp.tx + j;
              if (c < Cin) gx[r * Cin + c] = acc[i][j];
            }
          }
        }
      }
    }
  }
}

// gW[c][o] = sum_r x[r][c] gy[r][o], gb[o] = sum_r gy[r][o]; CTA (chunk, slice) accumulates the
// o-slice [o0, o0+os) over its chunk of rows in shared memory and writes a partial.
__global__ void __launch_bounds__(kThreads)
lin_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partW,
                 float* __restrict__ partB, int64_t rows, int Cin, int Cout, int os,
                 int64_t tiles_per_chunk) {
  extern __shared__ __align__(16) float sm[];
  const int ldc = (Cin + 3) & ~3;
  float* ACC = sm;                        // [ldc][os]   (c-major rows, o contiguous)
  float* X = ACC + ldc * os;              // [32][ldc]
  float* G = X + kTileFacets * ldc;       // [32][os]
  const int chunk = blockIdx.x, slice = blockIdx.y;
  const int o0 = slice * os;
  const int on = min(os, Cout - o0);
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  const int64_t t_begin = chunk * tiles_per_chunk;
  const int64_t t_end = min(ntiles, t_begin + tiles_per_chunk);
  for (int e = threadIdx.x; e < ldc * os; e += kThreads) ACC[e] = 0.f;
  float gb_acc = 0.f;
  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const int64_t r0 = tile * kTileFacets;
    __syncthreads();
    for (int e = threadIdx.x; e < kTileFacets * ldc; e += kThreads) {
      const int f = e / ldc, c = e % ldc;
      X[e] = (r0 + f < rows && c < Cin) ? __ldg(x + (r0 + f) * Cin + c) : 0.f;
    }
    for (int e = threadIdx.x; e < kTileFacets * os; e += kThreads) {
      const int f = e / os, o = e % os;
      G[e] = (r0 + f < rows && o < on) ? __ldg(gy + (r0 + f) * Cout + o0 + o) : 0.f;
    }
    __syncthreads();
    const int nbc = ldc / 4, nbo = os / 4;
    for (int blk = threadIdx.x; blk < nbc * nbo; blk += kThreads) {
      const int bc = blk / nbo, bo = blk % nbo;
      float a[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
#pragma unroll 4
      for (int f = 0; f < kTileFacets; ++f) {
        const float4 xv = *reinterpret_cast<const float4*>(X + f * ldc + 4 * bc);
        const float4 g = *reinterpret_cast<const float4*>(G + f * os + 4 * bo);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i][0] = fmaf(xa[i], g.x, a[i][0]);
          a[i][1] = fmaf(xa[i], g.y, a[i][1]);
          a[i][2] = fmaf(xa[i], g.z, a[i][2]);
          a[i][3] = fmaf(xa[i], g.w, a[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4* dst = reinterpret_cast<float4*>(ACC + (4 * bc + i) * os + 4 * bo);
        float4 cur = *dst;
        cur.x += a[i][0], cur.y += a[i][1], cur.z += a[i][2], cur.w += a[i][3];
        *dst = cur;
      }
    }
    if (threadIdx.x < on) {
      for (int f = 0; f < kTileFacets; ++f) gb_acc += G[f * os + threadIdx.x];
    }
  }
  __syncthreads();
  float* pw = partW + static_cast<int64_t>(chunk) * Cin * Cout;
  for (int e = threadIdx.x; e < Cin * on; e += kThreads) {
    const int c = e / on, o = e % on;
    pw[static_cast<int64_t>(c) * Cout + o0 + o] = ACC[c * os + o];
  }
  if (threadIdx.x < on) partB[static_cast<int64_t>(chunk) * Cout + o0 + threadIdx.x] = gb_acc;
}


// Weight gradient of a NARROW layer (Cout <= 4, e.g. the 1024 -> 3 projection of the head): the input rows are the
// big operand (rows x Cin floats, read once), so thread t owns four input channels, streams its 16 bytes of every
// row of the chunk and keeps the 4 x Cout sums in registers; the chunk's gy rows sit in shared memory.  Same
// partial layout and fixed summation order per chunk as lin_bwd_w_kernel.
constexpr int kNarrowRows = 128;
__global__ void __launch_bounds__(256)
lin_bwd_w_narrow_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partW,
                        float* __restrict__ partB, int64_t rows, int Cin, int Cout, int64_t rows_per_chunk) {
  __shared__ float G[kNarrowRows * 4];
  const int chunk = blockIdx.x;
  const int64_t r_begin = chunk * rows_per_chunk, r_end = min(rows, r_begin + rows_per_chunk);
  float gb_acc = 0.f;
  float* pw = partW + static_cast<int64_t>(chunk) * Cin * Cout;
  for (int cb = 0; cb < Cin; cb += 4 * 256) {
    // every thread runs every row block (barriers inside); threads past Cin only help staging
    const int c0 = cb + 4 * static_cast<int>(threadIdx.x);
    const bool own = c0 < Cin;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int o = 0; o < 4; ++o) acc[i][o] = 0.f;
    for (int64_t rb = r_begin; rb < r_end; rb += kNarrowRows) {
      const int nr = (r_end - rb < kNarrowRows) ? static_cast<int>(r_end - rb) : kNarrowRows;
      __syncthreads();
      for (int e = threadIdx.x; e < nr * 4; e += 256) {
        const int f = e >> 2, o = e & 3;
        G[e] = (o < Cout) ? __ldg(gy + (rb + f) * Cout + o) : 0.f;
      }
      __syncthreads();
      if (own) {
#pragma unroll 4
        for (int f = 0; f < nr; ++f) {
          const float4 h = __ldg(reinterpret_cast<const float4*>(x + (rb + f) * Cin + c0));
          const float4 g = *reinterpret_cast<const float4*>(G + 4 * f);
          const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(hv[i], g.x, acc[i][0]);
            acc[i][1] = fmaf(hv[i], g.y, acc[i][1]);
            acc[i][2] = fmaf(hv[i], g.z, acc[i][2]);
            acc[i][3] = fmaf(hv[i], g.w, acc[i][3]);
          }
        }
      }
      if (cb == 0 && static_cast<int>(threadIdx.x) < Cout)
        for (int f = 0; f < nr; ++f) gb_acc += G[4 * f + threadIdx.x];
    }
    if (own) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        for (int o = 0; o < Cout; ++o) pw[static_cast<int64_t>(c0 + i) * Cout + o] = acc[i][o];
    }
  }
  if (static_cast<int>(threadIdx.x) < Cout) partB[static_cast<int64_t>(chunk) * Cout + threadIdx.x] = gb_acc;
}

struct LinPlan {
  int os, nslices, chunks;
  int64_t tiles_per_chunk;
};

static void lin_plan(int64_t rows, int Cin, int Cout, LinPlan* pl) {
  const size_t ldc = (Cin + 3) & ~3;
  int os = (Cout + 3) & ~3;
  if (os > 128) os = 128;
  auto smem = [&](int o) { return (ldc * o + kTileFacets * ldc + kTileFacets * o) * 4; };
  while (os > 4 && smem(os) > 200 * 1024) os -= 4;
  pl->os = os;
  pl->nslices = (Cout + os - 1) / os;
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  int64_t chunks = (2 * num_sms()) / pl->nslices;
  if (chunks < 1) chunks = 1;
  if (chunks > ntiles) chunks = ntiles;
  if (chunks < 1) chunks = 1;
  pl->tiles_per_chunk = (ntiles + chunks - 1) / chunks;
  if (pl->tiles_per_chunk < 1) pl->tiles_per_chunk = 1;
  pl->chunks = static_cast<int>((ntiles + pl->tiles_per_chunk - 1) / pl->tiles_per_chunk);
  if (pl->chunks < 1) pl->chunks = 1;
}

}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_lin_fwd(const float* x, const float* W, const float* b, float* y, int64_t rows, int Cin,
                int Cout, int act, float alpha, void* stream) {
  FGC_REQUIRE(x && W && b && y && rows >= 0 && Cin > 0 && Cout > 0, "lin_fwd: bad arguments");
  if (rows == 0) return FGC_OK;
  if (Cout <= 4 && Cin % 4 == 0 && Cin >= 128) {
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    auto kern = Cout == 1 ? lin_fwd_narrow_kernel<1> : Cout == 2 ? lin_fwd_narrow_kernel<2>
                : Cout == 3 ? lin_fwd_narrow_kernel<3> : lin_fwd_narrow_kernel<4>;
    kern<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream)>>>(x, W, b, y, rows, Cin, act, alpha);
    FGC_LAUNCHED("lin_fwd_narrow_kernel");
    return FGC_OK;
  }
  const size_t smem = (static_cast<size_t>(kTileFacets) * ((Cin + 3) & ~3) + kChunkK * 128) * 4;
  FGC_UNSUPPORTED(smem > 227 * 1024, "lin_fwd: Cin = %d too large", Cin);
  FGC_CUDA(cudaFuncSetAttribute(lin_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  int64_t grid = static_cast<int64_t>(num_sms()) * 2;
  if (grid > ntiles) grid = ntiles;
  lin_fwd_kernel<<<static_cast<unsigned>(grid), kThreads, smem, as_stream(stream)>>>(
      x, W, b, y, rows, Cin, Cout, act, alpha);
  FGC_LAUNCHED("lin_fwd_kernel");
  return FGC_OK;
}

size_t fgc_mlp_head_workspace(int64_t rows, int Cin, int H, int Cout) {
  return mlp_head_tc_supported(rows, Cin, H, Cout) ? mlp_head_tc_workspace() : 0;
}

int fgc_mlp_head_fwd(const float* x, const float* W1, const float* b1, const float* W2,
                     const float* b2, float* y, int64_t rows, int Cin, int H, int Cout, float alpha,
                     void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x && W1 && b1 && W2 && b2 && y && rows >= 0 && Cin > 0 && H > 0 && Cout > 0,
              "mlp_head_fwd: bad arguments");
  FGC_UNSUPPORTED(Cout > kHeadMaxOut, "mlp_head_fwd: Cout <= %d supported", kHeadMaxOut);
  if (rows == 0) return FGC_OK;
  if (mlp_head_tc_supported(rows, Cin, H, Cout))
    return launch_mlp_head_tc(x, W1, b1, W2, b2, y, rows, alpha, workspace, workspace_bytes, as_stream(stream));
  const size_t smem = (static_cast<size_t>(kTileFacets) * ((Cin + 3) & ~3) + kChunkK * 128 +
                       kTileFacets * 32 * kHeadMaxOut) * 4;
  FGC_UNSUPPORTED(smem > 227 * 1024, "mlp_head_fwd: Cin = %d too large", Cin);
  FGC_CUDA(cudaFuncSetAttribute(mlp_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  int64_t grid = static_cast<int64_t>(num_sms()) * 2;
  if (grid > ntiles) grid = ntiles;
  mlp_head_kernel<<<static_cast<unsigned>(grid), kThreads, smem, as_stream(stream)>>>(
      x, W1, b1, W2, b2, y, rows, Cin, H, Cout, alpha);
  FGC_LAUNCHED("mlp_head_kernel");
  return FGC_OK;
}

size_t fgc_lin_bwd_workspace(int64_t rows, int Cin, int Cout) {
  LinPlan pl;
  lin_plan(rows, Cin, Cout, &pl);
  const size_t nW = static_cast<size_t>(Cin) * Cout;
  return ws_bytes(nW, 4) + ws_bytes(static_cast<size_t>(pl.chunks) * nW, 4) +
         ws_bytes(static_cast<size_t>(pl.chunks) * Cout, 4) + 1024;
}

int fgc_lin_bwd(const float* gy, const float* x, const float* W, float* gx, float* gW, float* gb,
                int64_t rows, int Cin, int Cout, void* workspace, size_t workspace_bytes,
                void* stream) {
  FGC_REQUIRE(gy && x && W && gW && gb && rows > 0 && Cin > 0 && Cout > 0, "lin_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  LinPlan pl;
  lin_plan(rows, Cin, Cout, &pl);
  const size_t nW = static_cast<size_t>(Cin) * Cout;
  Workspace ws(workspace, workspace_bytes);
  float* Wt = ws.take<float>(nW);
  float* partW = ws.take<float>(static_cast<size_t>(pl.chunks) * nW);
  float* partB = ws.take<float>(static_cast<size_t>(pl.chunks) * Cout);
  FGC_REQUIRE(ws.ok(), "lin_bwd: workspace too small");
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  if (gx) {
    transpose2d_kernel<<<static_cast<unsigned>((nW + 255) / 256), 256, 0, st>>>(W, Wt, Cin, Cout);
    FGC_LAUNCHED("transpose2d_kernel");
    if (Cin == 32 && Cout % 128 == 0) {
      int64_t blocks = (rows + 63) / 64;
      const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
      if (blocks > cap) blocks = cap;
      lin_bwd_x_wide_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(gy, Wt, gx, rows, Cout);
      FGC_LAUNCHED("lin_bwd_x_wide_kernel");
    } else {
    const size_t smem = (static_cast<size_t>(kTileFacets) * ((Cout + 3) & ~3) + kChunkK * 128) * 4;
    FGC_UNSUPPORTED(smem > 227 * 1024, "lin_bwd: Cout = %d too large", Cout);
    FGC_CUDA(cudaFuncSetAttribute(lin_bwd_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = static_cast<int64_t>(num_sms()) * 2;
    if (grid > ntiles) grid = ntiles;
    lin_bwd_x_kernel<<<static_cast<unsigned>(grid), kThreads, smem, st>>>(gy, Wt, gx, rows, Cin, Cout);
    FGC_LAUNCHED("lin_bwd_x_kernel");
    }
  }
  if (Cout <= 4 && Cin % 4 == 0 && Cin >= 256) {
    const int64_t rows_per_chunk = pl.tiles_per_chunk * kTileFacets;
    lin_bwd_w_narrow_kernel<<<pl.chunks, 256, 0, st>>>(x, gy, partW, partB, rows, Cin, Cout, rows_per_chunk);
    FGC_LAUNCHED("lin_bwd_w_narrow_kernel");
  } else {
    const size_t ldc = (Cin + 3) & ~3;
    const size_t smem = (ldc * pl.os + kTileFacets * ldc + kTileFacets * pl.os) * 4;
    FGC_UNSUPPORTED(smem > 227 * 1024, "lin_bwd: Cin = %d too large", Cin);
    FGC_CUDA(cudaFuncSetAttribute(lin_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(pl.chunks, pl.nslices);
    lin_bwd_w_kernel<<<grid, kThreads, smem, st>>>(x, gy, partW, partB, rows, Cin, Cout, pl.os,
                                                   pl.tiles_per_chunk);
    FGC_LAUNCHED("lin_bwd_w_kernel");
  }
  int rc = launch_reduce_partials(partW, gW, nW, pl.chunks, nW, st);
  if (rc) return rc;
  return launch_reduce_partials(partB, gb, Cout, pl.chunks, Cout, st);
}

}  // extern "C"
