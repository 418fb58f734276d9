// Forward facet-graph convolution (replaces reference Code/model.py:427-504, :74-95, :380-405).
//
// Pipeline per call (all on the caller's stream):
//   1. assign_logits_kernel : uvx[r, 0:M] = u.x_r + c,  uvx[r, M:2M] = v.x_r          (one pass over x)
//   2. transpose_w_kernel   : Wt[(m,c)][o] = W0[m][o][c]                              (tiny)
//   3. conv_fwd_kernel      : per tile of 32 facets
//        phase 1  warp-per-facet: soft assignments (lane per neighbour slot), then the
//                 q-weighted aggregation s[m][c] = sum_k q[k][m] x_{j_k}[c] with coalesced
//                 row loads (lanes over channels); s goes to shared memory
//        phase 2  tile GEMM y = s . Wt streamed through shared memory in 32-row chunks,
//                 fused epilogue (1/cnt, masked bias, leaky ReLU)
// The [B,N,K,M*Cout] tensor the reference materialises never exists.
#include <stdlib.h>

#include "conv_common.cuh"
#include "conv_launch.cuh"
#include "tc_common.cuh"

namespace fgc {

// ------------------------------------------------------------------ assignment logits pre-pass
// uvx[r, 0:M] = u . x_r[Ca0:Ca0+Ca] + c ; uvx[r, M:2M] = v . x_r[Ca0:Ca0+Ca].  Thread per row: the row
// streams through registers four channels at a time, u|v are broadcast from shared memory.
template <int OP>  // OP >= 2M accumulators
__global__ void __launch_bounds__(128)
assign_logits_kernel(const float* __restrict__ x, const float* __restrict__ u,
                     const float* __restrict__ v, const float* __restrict__ c,
                     float* __restrict__ uvx, int64_t rows, int Cin, int Ca0, int Ca, int M) {
  extern __shared__ __align__(16) float sm[];
  const int O = 2 * M;
  const int Ca4 = (Ca + 3) & ~3;
  float* uv = sm;  // [O][Ca4], zero padded
  for (int e = threadIdx.x; e < O * Ca4; e += blockDim.x) {
    const int o = e / Ca4, cc = e % Ca4;
    uv[e] = cc < Ca ? ((o < M) ? u[o * Ca + cc] : v[(o - M) * Ca + cc]) : 0.f;
  }
  __syncthreads();
  const bool vec = (Cin % 4 == 0) && (Ca0 % 4 == 0) && (Ca % 4 == 0);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc[OP];
#pragma unroll
    for (int o = 0; o < OP; ++o) acc[o] = 0.f;
    const float* xr = x + r * Cin + Ca0;
    for (int c0 = 0; c0 < Ca4; c0 += 4) {
      float4 xv;
      if (vec) {
        xv = __ldg(reinterpret_cast<const float4*>(xr + c0));
      } else {
        xv.x = c0 < Ca ? __ldg(xr + c0) : 0.f;
        xv.y = c0 + 1 < Ca ? __ldg(xr + c0 + 1) : 0.f;
        xv.z = c0 + 2 < Ca ? __ldg(xr + c0 + 2) : 0.f;
        xv.w = c0 + 3 < Ca ? __ldg(xr + c0 + 3) : 0.f;
      }
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        if (o < O) {
          const float4 w = *reinterpret_cast<const float4*>(uv + o * Ca4 + c0);
          acc[o] = fmaf(xv.x, w.x, fmaf(xv.y, w.y, fmaf(xv.z, w.z, fmaf(xv.w, w.w, acc[o]))));
        }
      }
    }
    float* dst = uvx + r * O;
#pragma unroll
    for (int o = 0; o < OP; ++o)
      if (o < O) dst[o] = acc[o] + (o < M ? __ldg(c + o) : 0.f);
  }
}

bool assign_logits_absmax_supported(const fgc_conv_shape* s) {
  static const bool slow = getenv("FGC_DISABLE_FAST_LOGITS") != nullptr;
  return !slow && logits_fast_supported(s->Cin, s->Ca0, s->Ca, s->M) && s->Ca0 == 0 && s->Ca == s->Cin;
}

int launch_assign_logits(const fgc_conv_shape* s, const float* x, const float* u, const float* v,
                         const float* c, float* uvx, cudaStream_t st, unsigned* maxbits) {
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  static const bool slow = getenv("FGC_DISABLE_FAST_LOGITS") != nullptr;
  if (!slow && logits_fast_supported(s->Cin, s->Ca0, s->Ca, s->M))
    return launch_assign_logits_fast(x, u, v, c, uvx, rows, s->Cin, s->Ca0, s->Ca, s->M, st,
                                     assign_logits_absmax_supported(s) ? maxbits : nullptr);
  FGC_REQUIRE(maxbits == nullptr, "assign_logits: fused absmax needs the warp-cooperative kernel");
  const int O = 2 * s->M;
  const size_t smem = static_cast<size_t>(O) * ((s->Ca + 3) & ~3) * 4;
  int64_t blocks = (rows + 127) / 128;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const unsigned g = static_cast<unsigned>(blocks);
  if (O <= 8) assign_logits_kernel<8><<<g, 128, smem, st>>>(x, u, v, c, uvx, rows, s->Cin, s->Ca0, s->Ca, s->M);
  else if (O <= 16) assign_logits_kernel<16><<<g, 128, smem, st>>>(x, u, v, c, uvx, rows, s->Cin, s->Ca0, s->Ca, s->M);
  else if (O <= 18) assign_logits_kernel<18><<<g, 128, smem, st>>>(x, u, v, c, uvx, rows, s->Cin, s->Ca0, s->Ca, s->M);
  else assign_logits_kernel<32><<<g, 128, smem, st>>>(x, u, v, c, uvx, rows, s->Cin, s->Ca0, s->Ca, s->M);
  FGC_LAUNCHED("assign_logits_kernel");
  return FGC_OK;
}

// ------------------------------------------------------------------ weight transpose
__global__ void transpose_w_kernel(const float* __restrict__ W0, float* __restrict__ Wt, int M,
                                   int Cout, int Cw) {
  const int total = M * Cout * Cw;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int o = e % Cout;
    const int mc = e / Cout;
    const int m = mc / Cw, cc = mc % Cw;
    Wt[e] = W0[(static_cast<int64_t>(m) * Cout + o) * Cw + cc];
  }
}

int launch_transpose_w(const float* W0, float* Wt, int M, int Cout, int Cw, cudaStream_t st) {
  const int total = M * Cout * Cw;
  transpose_w_kernel<<<(total + 255) / 256, 256, 0, st>>>(W0, Wt, M, Cout, Cw);
  FGC_LAUNCHED("transpose_w_kernel");
  return FGC_OK;
}

// ------------------------------------------------------------------ forward kernel
template <int MP>
__host__ __device__ constexpr int q_stride() { return QStride<MP>::value; }

// dynamic shared memory carve-up (floats): S | Bs | qs | nbr | inv | flag
__host__ __device__ inline int fwd_lda(int M, int Cw) { return ((M * Cw + 3) & ~3); }

template <int MP, int NC>
__global__ void __launch_bounds__(kThreads)
conv_fwd_kernel(const ConvFwdParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int QS = QStride<MP>::value;
  const int KK = p.M * p.Cw;
  const int lda = fwd_lda(p.M, p.Cw);
  float* S = sm;                                   // [32][lda]
  float* Bs = S + kTileFacets * lda;               // [32][128]
  float* qs_all = Bs + kChunkK * 128;              // [8][32][QS]
  int* nbr_all = reinterpret_cast<int*>(qs_all + kWarps * 32 * QS);  // [8][32]
  float* inv = reinterpret_cast<float*>(nbr_all + kWarps * 32);      // [32]
  float* flag = inv + kTileFacets;                                   // [32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = qs_all + warp * 32 * QS;
  int* nbr = nbr_all + warp * 32;
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileFacets;
    // ---------------- phase 1: assignments + aggregation, warp per facet
    for (int f = warp; f < kTileFacets; f += kWarps) {
      const int64_t r = r0 + f;
      float acc[MP][NC];
#pragma unroll
      for (int m = 0; m < MP; ++m)
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[m][i] = 0.f;
      int cnt = 0;
      if (r < p.rows) {
        const int64_t base = (r / p.N) * p.N;
        cnt = facet_assign<MP>(p.adj, p.uvx, r, base, p.N, p.K, p.M, qs, nbr, lane);
        aggregate_rows<MP, NC>(p.x, p.Cin, p.Cw, p.K, qs, nbr, lane, acc);
      }
      float* Sr = S + f * lda;
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        if (m < p.M) {
#pragma unroll
          for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            if (c < p.Cw) Sr[m * p.Cw + c] = acc[m][i];
          }
        }
      }
      if (lane < lda - KK) Sr[KK + lane] = 0.f;
      if (lane == 0) {
        inv[f] = cnt ? 1.f / static_cast<float>(cnt) : 0.f;
        flag[f] = (cnt > 0 || !p.bias_mask) ? 1.f : 0.f;
      }
      __syncwarp();
    }
    // ---------------- phase 2: contraction y = S . Wt, 128 output columns at a time
    for (int o0 = 0; o0 < p.Cout; o0 += 128) {
      const int ncols = min(128, p.Cout - o0);
      const TileGemmMap mp(ncols);
      float acc[4][4];
      tile_gemm(S, lda, KK, p.Wt, p.Cout, o0, ncols, Bs, mp, acc);  // syncs inside
      if (mp.ty < mp.TY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = mp.ty + mp.TY * i;
          const int64_t r = r0 + f;
          if (i < mp.RF && f < kTileFacets && r < p.rows) {
            const float sc = inv[f], fl = flag[f];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int o = o0 + 4 * mp.tx + j;
              if (o < p.Cout) {
                float yv = fmaf(sc, acc[i][j], fl * __ldg(p.b + o));
                if (p.act == FGC_ACT_LRELU) yv = lrelu_f(yv, p.alpha);
                p.y[r * p.Cout + o] = yv;
              }
            }
          }
        }
      }
    }
    __syncthreads();  // S / inv reused by the next tile
  }
}

template <int MP>
static size_t fwd_smem_bytes(int M, int Cw) {
  constexpr int QS = QStride<MP>::value;
  return (static_cast<size_t>(kTileFacets) * fwd_lda(M, Cw) + kChunkK * 128 + kWarps * 32 * QS +
          kWarps * 32 + 2 * kTileFacets) * 4;
}

template <int MP, int NC>
static int launch_conv_fwd_t(const ConvFwdParams& p, cudaStream_t st) {
  const size_t smem = fwd_smem_bytes<MP>(p.M, p.Cw);
  if (smem > 227 * 1024) {
    set_error("conv_fwd: M*Cw = %d needs %zu bytes of shared memory (> 227 KB)", p.M * p.Cw, smem);
    return FGC_ERR_UNSUPPORTED;
  }
  FGC_CUDA(cudaFuncSetAttribute(conv_fwd_kernel<MP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  int occ = 1;
  FGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_fwd_kernel<MP, NC>, kThreads, smem));
  if (occ < 1) occ = 1;
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;
  int64_t grid = static_cast<int64_t>(num_sms()) * occ;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  conv_fwd_kernel<MP, NC><<<static_cast<unsigned>(grid), kThreads, smem, st>>>(p);
  FGC_LAUNCHED("conv_fwd_kernel");
  return FGC_OK;
}

int launch_conv_fwd(const ConvFwdParams& p, cudaStream_t st) {
#define FGC_CALL(MPV, NCV) return launch_conv_fwd_t<MPV, NCV>(p, st)
  FGC_DISPATCH_MP_NC(pick_mp(p.M), pick_nc(p.Cw), FGC_CALL);
#undef FGC_CALL
  return FGC_OK;
}

// ------------------------------------------------------------------ narrow input layer (thread per facet)
// The network's first layer (6 -> 32, M = 9: reference Code/model.py:858) gathers 24-byte rows: a warp per
// facet leaves 26 of 32 lanes idle in the aggregation.  Here one thread owns one facet end to end -- own and
// neighbour logits inline (no uvx pre-pass), softmax, s[M][CIN] in registers, contraction against a
// transposed weight image in shared memory read by broadcast -- same arithmetic order per facet for every
// launch geometry.
#ifndef FGC_SMALL_BLOCKS
#define FGC_SMALL_BLOCKS 4
#endif
template <int M, int CIN, int COUT>
__global__ void __launch_bounds__(128, FGC_SMALL_BLOCKS)
conv_fwd_small_kernel(const float* __restrict__ x, const int32_t* __restrict__ adj, const float* __restrict__ W0,
                      const float* __restrict__ b, const float* __restrict__ u, const float* __restrict__ v,
                      const float* __restrict__ c, float* __restrict__ y, int64_t rows, int N, int K, int bias_mask,
                      int act, float alpha, float* __restrict__ ypool, unsigned* __restrict__ ymax) {
  __shared__ __align__(16) float Wt[M * CIN * COUT];   // [(m, c)][o]
  __shared__ __align__(8) float us[M * CIN], vs[M * CIN], cs[M];
  for (int e = threadIdx.x; e < M * CIN * COUT; e += blockDim.x) {
    const int o = e % COUT, mc = e / COUT;
    Wt[e] = W0[(static_cast<size_t>(mc / CIN) * COUT + o) * CIN + mc % CIN];
  }
  for (int e = threadIdx.x; e < M * CIN; e += blockDim.x) us[e] = u[e], vs[e] = v[e];
  if (threadIdx.x < M) cs[threadIdx.x] = c[threadIdx.x];
  __syncthreads();
  // warp-uniform trip count (the fused pooling / max|y| outputs shuffle across the warp): lanes past the end compute on
  // row rows - 1 and store nothing
  for (int64_t rw = static_cast<int64_t>(blockIdx.x) * blockDim.x + (threadIdx.x & ~31); rw < rows;
       rw += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool valid = rw + (threadIdx.x & 31) < rows;
    const int64_t r = valid ? rw + (threadIdx.x & 31) : rows - 1;
    const int64_t base = (r / N) * N;
    static_assert(CIN % 2 == 0, "channel pairs");
    float xn[CIN], own[M];
    float2 s2[M][CIN / 2];          // s[m][i] in channel pairs: the aggregation runs on packed fp32x2 FMAs
#pragma unroll
    for (int i = 0; i < CIN; ++i) xn[i] = __ldg(x + r * CIN + i);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float a = cs[m];
#pragma unroll
      for (int i = 0; i < CIN; ++i) a = fmaf(us[m * CIN + i], xn[i], a);
      own[m] = a;
#pragma unroll
      for (int i = 0; i < CIN / 2; ++i) s2[m][i] = make_float2(0.f, 0.f);
    }
    int cnt = 0;
    for (int k = 0; k < K; ++k) {
      const int id = __ldg(adj + r * K + k);
      if (id == 0) continue;          // padding: contributes nothing and is not counted
      ++cnt;
      if (id < 0 || id > N) continue; // id outside the patch: a counted neighbour that contributes zero
      const float* xr = x + (base + id - 1) * CIN;
      // rows of CIN = 6 floats are 8-byte aligned: three 8-byte loads per neighbour (every load of a thread-per-facet
      // gather touches 32 different lines, so the load count is what the LSU pays for)
      float xj[CIN];
      float2 xj2[CIN / 2];
#pragma unroll
      for (int i = 0; i < CIN / 2; ++i) {
        xj2[i] = __ldg(reinterpret_cast<const float2*>(xr) + i);
        xj[2 * i] = xj2[i].x, xj[2 * i + 1] = xj2[i].y;
      }
      float q[M];
      float mx = -3.4e38f;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        float a = own[m];     // (the dot products on packed FMAs, two partial sums per weight, measured slower)
#pragma unroll
        for (int i = 0; i < CIN; ++i) a = fmaf(vs[m * CIN + i], xj[i], a);
        q[m] = a;
        mx = fmaxf(mx, a);
      }
      float sum = 0.f;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        q[m] = __expf(q[m] - mx);
        sum += q[m];
      }
      const float rs = 1.f / sum;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float qq = q[m] * rs;
        const float2 qq2 = make_float2(qq, qq);
#pragma unroll
        for (int i = 0; i < CIN / 2; ++i) tc::ffma2(s2[m][i], qq2, xj2[i]);
      }
    }
    // contraction on packed fp32x2 FMAs (two outputs per instruction; each output still one IEEE fma chain in (m, i) order)
    float2 acc2[COUT / 2];
#pragma unroll
    for (int o = 0; o < COUT / 2; ++o) acc2[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int i = 0; i < CIN; ++i) {
        const float sv1 = (i & 1) ? s2[m][i >> 1].y : s2[m][i >> 1].x;
        const float2 sv = make_float2(sv1, sv1);
        const float4* wr = reinterpret_cast<const float4*>(Wt + (m * CIN + i) * COUT);
#pragma unroll
        for (int o4 = 0; o4 < COUT / 4; ++o4) {
          const float4 w = wr[o4];
          tc::ffma2(acc2[2 * o4], sv, make_float2(w.x, w.y));
          tc::ffma2(acc2[2 * o4 + 1], sv, make_float2(w.z, w.w));
        }
      }
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT / 2; ++o) acc[2 * o] = acc2[o].x, acc[2 * o + 1] = acc2[o].y;
    const float inv = cnt ? 1.f / static_cast<float>(cnt) : 0.f;
    const float fl = (cnt > 0 || !bias_mask) ? 1.f : 0.f;
    float4* yr = reinterpret_cast<float4*>(y + r * COUT);
    float amax = 0.f;
#pragma unroll
    for (int o4 = 0; o4 < COUT / 4; ++o4) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float yv = fmaf(inv, acc[4 * o4 + j], fl * __ldg(b + 4 * o4 + j));
        if (act == FGC_ACT_LRELU) yv = lrelu_f(yv, alpha);
        o[j] = yv;
        amax = fmaxf(amax, fabsf(yv));
      }
      if (valid) yr[o4] = make_float4(o[0], o[1], o[2], o[3]);
      if (ypool != nullptr) {
        // custom_binary_tree_pooling (model.py:863): max over the four consecutive rows of a group = four lanes
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = fmaxf(o[j], __shfl_xor_sync(0xffffffffu, o[j], 1));
          o[j] = fmaxf(o[j], __shfl_xor_sync(0xffffffffu, o[j], 2));
        }
        if (valid && (threadIdx.x & 3) == 0)
          reinterpret_cast<float4*>(ypool + (r >> 2) * COUT)[o4] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    if (ymax != nullptr) {
      // max|y| per batch element (the image scale of the next layer): one atomic per warp when its rows share the element
      const int be = static_cast<int>(r / N);
      const int be0 = __shfl_sync(0xffffffffu, be, 0);
      const bool same = __all_sync(0xffffffffu, be == be0);
      if (!valid) amax = 0.f;
      if (same) {
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, sft));
        if ((threadIdx.x & 31) == 0) atomicMax(ymax + be0, __float_as_uint(amax));
      } else if (valid) {
        atomicMax(ymax + be, __float_as_uint(amax));
      }
    }
  }
}

bool conv_fwd_small_supported(const fgc_conv_shape* s) {
  static const bool disabled = getenv("FGC_DISABLE_SMALL") != nullptr;
  return !disabled && s->M == 9 && s->Cin == 6 && s->Cw == 6 && s->Ca0 == 0 && s->Ca == 6 && s->Cout == 32;
}

int launch_conv_fwd_small(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0, const float* b,
                          const float* u, const float* v, const float* c, float* y, int bias_mask, int act,
                          float alpha, cudaStream_t st, float* ypool, unsigned* ymax) {
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  FGC_REQUIRE(ypool == nullptr || s->N % 4 == 0, "conv_fwd_small: pooled output needs N %% 4 == 0");
  int64_t blocks = (rows + 127) / 128;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  conv_fwd_small_kernel<9, 6, 32><<<static_cast<unsigned>(blocks), 128, 0, st>>>(x, adj, W0, b, u, v, c, y, rows, s->N,
                                                                               s->K, bias_mask, act, alpha, ypool, ymax);
  FGC_LAUNCHED("conv_fwd_small_kernel");
  return FGC_OK;
}

// ------------------------------------------------------------------ debug / parity helpers
__global__ void gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ adj,
                                   float* __restrict__ out, int64_t rows, int N, int K, int C) {
  // one warp per (row, slot); pure copy => bit-exact
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = rows * K;
  for (int64_t e = w; e < total; e += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
    const int64_t r = e / K;
    const int id = adj[e];
    const int64_t base = (r / N) * N;
    const bool valid = id > 0 && id <= N;
    for (int c = lane; c < C; c += 32)
      out[e * C + c] = valid ? x[(base + id - 1) * C + c] : 0.f;
  }
}

template <int MP>
__global__ void __launch_bounds__(kThreads)
assignments_kernel(const int32_t* __restrict__ adj, const float* __restrict__ uvx,
                   float* __restrict__ q, int64_t rows, int N, int K, int M) {
  constexpr int QS = QStride<MP>::value;
  __shared__ __align__(16) float qs_all[kWarps * 32 * QS];
  __shared__ int nbr_all[kWarps * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = qs_all + warp * 32 * QS;
  int* nbr = nbr_all + warp * 32;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarps + warp; r < rows;
       r += static_cast<int64_t>(gridDim.x) * kWarps) {
    const int64_t base = (r / N) * N;
    facet_assign<MP>(adj, uvx, r, base, N, K, M, qs, nbr, lane);
    for (int e = lane; e < K * M; e += 32) q[r * K * M + e] = qs[(e / M) * QS + e % M];
    __syncwarp();
  }
}

int launch_assignments(const int32_t* adj, const float* uvx, float* q, int64_t rows, int N, int K,
                       int M, cudaStream_t st) {
  int64_t blocks = (rows + kWarps - 1) / kWarps;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  switch (pick_mp(M)) {
    case 4: assignments_kernel<4><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(adj, uvx, q, rows, N, K, M); break;
    case 8: assignments_kernel<8><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(adj, uvx, q, rows, N, K, M); break;
    case 9: assignments_kernel<9><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(adj, uvx, q, rows, N, K, M); break;
    default: assignments_kernel<16><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(adj, uvx, q, rows, N, K, M); break;
  }
  FGC_LAUNCHED("assignments_kernel");
  return FGC_OK;
}

int launch_gather_rows(const float* x, const int32_t* adj, float* out, int64_t rows, int N, int K,
                       int C, cudaStream_t st) {
  int64_t warps = rows * K;
  int64_t blocks = (warps + 7) / 8;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  gather_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, adj, out, rows, N, K, C);
  FGC_LAUNCHED("gather_rows_kernel");
  return FGC_OK;
}

}  // namespace fgc
