// Backward of the facet-graph convolution: what TF autodiff of reference Code/model.py:427-504
// computes, restructured so that no floating-point atomics / scatters are needed.
//
// With gz[n] = inv_cnt[n] * gy[n] and the assignment logits a[n,k,m] = uxc[n,m] + vx[j_k,m]:
//   bwd_src_kernel (source-centric, per facet n)
//       ds[n,m,:]  = W0[m]^T gz[n]                              (tile GEMM)
//       dq[n,k,m]  = ds[n,m,:] . x_{j_k}[0:Cw]
//       da[n,k,m]  = q (dq - sum_m' q dq)   -> da_edge[n,k,m]   (workspace)
//       d_uvx[n,0:M] = sum_k da[n,k,:]                           (grad of the own-row logit part)
//   bwd_tgt_kernel (target-centric, per facet j, pulls over the reversed adjacency)
//       t[j,m,:]   = sum_{(n,k)->j} q[n,k,m] gz[n,:]
//       gx[j,0:Cw] = sum_m W0[m]^T... i.e. sum_{m,o} t[j,m,o] W0[m,o,:]       (tile GEMM)
//       d_uvx[j,M:2M] = sum_{(n,k)->j} da_edge[n,k,:]            (grad of the neighbour logit part)
//   bwd_w_kernel   (source-centric)  gW0[m] = sum_n gz[n] (x) s[n,m,:],  gb = sum_n flag[n] gy[n]
//       accumulated per CTA in shared memory, written as per-CTA partials
//   logits_bwd_kernel  d_uvx -> gu, gv, gc partials and the logit-window part of gx
//   reduce_partials_kernel  fixed-order sum of the per-CTA partials
// Every summation order is fixed by the launch geometry => bit-reproducible run to run.
#include <stdlib.h>

#include <algorithm>

#include <cub/device/device_scan.cuh>

#include "conv_common.cuh"
#include "conv_launch.cuh"

namespace fgc {


// ------------------------------------------------------------------ weight permutation for ds
// Wd[o][(m,c)] = W0[m][o][c]
__global__ void permute_w_ds_kernel(const float* __restrict__ W0, float* __restrict__ Wd, int M,
                                    int Cout, int Cw) {
  const int total = M * Cout * Cw;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int cc = e % Cw;
    const int mo = e / Cw;
    const int m = mo / Cout, o = mo % Cout;
    Wd[(static_cast<int64_t>(o) * M + m) * Cw + cc] = W0[e];
  }
}

// multi-value warp reduction: every lane holds part[0..MP); on return lane L (L%4==0) holds in
// `out` the full sum of value index (L>>2) (for MP<=8), see callers.  Generic fallback: loop.
template <int MP>
__device__ __forceinline__ void reduce_to_smem(float (&part)[MP], int M, float* dst, int lane) {
  // dst[m] = sum over lanes of part[m]
  if constexpr (MP == 8 || MP == 9 || MP == 4 || MP == 16) {
    constexpr int P2 = (MP == 9) ? 8 : MP;  // power-of-two body
    float v[P2];
#pragma unroll
    for (int i = 0; i < P2; ++i) v[i] = part[i];
    int width = P2;
    int bit = 16;
    // halving steps: each step sends half of the values to the partner lane
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      if (width > 1) {
        const int half = width / 2;
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < P2 / 2; ++i) {
          if (i < half) {
            const float keep = upper ? v[i + half] : v[i];
            const float send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
        width = half;
        bit >>= 1;
      }
    }
    // now v[0] holds a partial over the lanes that differ in the consumed bits; finish
    float r = v[0];
    for (int b = bit; b > 0; b >>= 1) r += __shfl_xor_sync(0xffffffffu, r, b);
    // which value index does this lane hold?  bits consumed from 16 downwards select halves
    int idx = 0;
    {
      int w = P2, bb = 16;
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        if (w > 1) {
          w /= 2;
          if (lane & bb) idx += w;
          bb >>= 1;
        }
      }
    }
    const int lowmask = (P2 == 16) ? 1 : (P2 == 8 ? 3 : 7);
    if ((lane & lowmask) == 0 && idx < M) dst[idx] = r;
    if constexpr (MP == 9) {
      float r8 = warp_sum(part[8]);
      if (lane == 0 && M > 8) dst[8] = r8;
    }
  } else {
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      const float r = warp_sum(part[m]);
      if (lane == 0 && m < M) dst[m] = r;
    }
  }
}

// ------------------------------------------------------------------ source-centric pass
struct BwdSrcParams {
  const float* gy;
  const float* x;
  const int32_t* adj;
  const float* uvx;
  const float* Wd;     // [Cout][M*Cw]
  float* da_edge;      // [rows*K][M]
  float* d_uvx;        // [rows][2M]  (this kernel writes columns 0..M-1)
  float* inv_out;      // [rows]
  int64_t rows;
  int N, K, Cin, Cw, Cout, M;
};

template <int MP, int NC>
__global__ void __launch_bounds__(kThreads)
bwd_src_kernel(const BwdSrcParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int QS = QStride<MP>::value;
  const int KK = p.M * p.Cw;
  const int lda = (KK + 3) & ~3;
  const int ldg = (p.Cout + 3) & ~3;
  float* DS = sm;                                   // [32][lda]
  float* GZ = DS + kTileFacets * lda;               // [32][ldg]
  // the weight chunk of the GEMM phase and the per-warp assignment tables of the per-facet phase share one region:
  // block barriers separate the phases in both directions (after the GEMM, at the end of the tile loop), and the
  // 16 KB it saves let two CTAs of the 64-channel layers share an SM (119.8 -> 103.4 KB)
  float* Bs = GZ + kTileFacets * ldg;               // [32][128]
  float* qs_all = Bs;                               // [8][32][QS]
  float* dq_all = qs_all + kWarps * 32 * QS;        // [8][32][QS]
  constexpr int kShared = (kChunkK * 128 > 2 * kWarps * 32 * QS) ? kChunkK * 128 : 2 * kWarps * 32 * QS;
  int* nbr_all = reinterpret_cast<int*>(Bs + kShared);  // [8][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = qs_all + warp * 32 * QS;
  float* dqs = dq_all + warp * 32 * QS;
  int* nbr = nbr_all + warp * 32;
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileFacets;
    // ---- gz tile
    for (int f = warp; f < kTileFacets; f += kWarps) {
      const int64_t r = r0 + f;
      float iv = 0.f;
      if (r < p.rows) {
        int id = 0;
        if (lane < p.K) id = __ldg(p.adj + r * p.K + lane);
        const int cnt = __popc(__ballot_sync(0xffffffffu, id != 0));
        iv = cnt ? 1.f / static_cast<float>(cnt) : 0.f;
        if (lane == 0) p.inv_out[r] = iv;
      }
      for (int o = lane; o < ldg; o += 32)
        GZ[f * ldg + o] = (r < p.rows && o < p.Cout) ? iv * __ldg(p.gy + r * p.Cout + o) : 0.f;
    }
    // ---- ds = gz . Wd   (columns (m,c) in blocks of 128)
    for (int c0 = 0; c0 < KK; c0 += 128) {
      const int ncols = min(128, KK - c0);
      const TileGemmMap mp(ncols);
      float acc[4][4];
      tile_gemm(GZ, ldg, p.Cout, p.Wd, KK, c0, ncols, Bs, mp, acc);
      if (mp.ty < mp.TY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = mp.ty + mp.TY * i;
          if (i < mp.RF && f < kTileFacets) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = c0 + 4 * mp.tx + j;
              if (col < KK) DS[f * lda + col] = acc[i][j];
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- per facet: dq, softmax backward, da_edge, d_ux
    for (int f = warp; f < kTileFacets; f += kWarps) {
      const int64_t r = r0 + f;
      if (r >= p.rows) continue;  // warp-uniform
      const int64_t base = (r / p.N) * p.N;
      facet_assign<MP>(p.adj, p.uvx, r, base, p.N, p.K, p.M, qs, nbr, lane);
      const float* dsr = DS + f * lda;
      for (int k = 0; k < p.K; ++k) {
        const int j = nbr[k];
        float part[MP];
#pragma unroll
        for (int m = 0; m < MP; ++m) part[m] = 0.f;
        if (j >= 0) {
          const float* xr = p.x + static_cast<int64_t>(j) * p.Cin;
#pragma unroll
          for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            if (c < p.Cw) {
              const float xv = __ldg(xr + c);
#pragma unroll
              for (int m = 0; m < MP; ++m)
                if (m < p.M) part[m] = fmaf(dsr[m * p.Cw + c], xv, part[m]);
            }
          }
          reduce_to_smem<MP>(part, p.M, dqs + k * QS, lane);
        } else if (lane < QS) {
          dqs[k * QS + lane] = 0.f;
        }
      }
      __syncwarp();
      float dux[MP];
#pragma unroll
      for (int m = 0; m < MP; ++m) dux[m] = 0.f;
      if (lane < p.K) {
        float dot = 0.f;
#pragma unroll
        for (int m = 0; m < MP; ++m)
          if (m < p.M) dot = fmaf(qs[lane * QS + m], dqs[lane * QS + m], dot);
        float* de = p.da_edge + (r * p.K + lane) * p.M;
#pragma unroll
        for (int m = 0; m < MP; ++m) {
          if (m < p.M) {
            const float da = qs[lane * QS + m] * (dqs[lane * QS + m] - dot);
            de[m] = da;
            dux[m] = da;
          }
        }
      }
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        if (m < p.M) {
          const float t = warp_sum(dux[m]);
          if (lane == 0) p.d_uvx[r * 2 * p.M + m] = t;
        }
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ target-centric pass
struct BwdTgtParams {
  const float* gy;
  const float* uvx;
  const float* W0;     // natural layout = [(m,o)][c]
  const float* da_edge;
  const float* inv;    // [rows]
  const int32_t* rev_ptr;
  const int32_t* rev_edge;
  float* gx;           // [rows][Cin]
  float* d_uvx;        // writes columns M..2M-1
  int64_t rows;
  int N, K, Cin, Cw, Cout, M;
};

template <int MP, int NC>  // NC over Cout here
__global__ void __launch_bounds__(kThreads)
bwd_tgt_kernel(const BwdTgtParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int QS = QStride<MP>::value;
  const int KK = p.M * p.Cout;
  const int lda = (KK + 3) & ~3;
  float* T = sm;                                   // [32][lda]
  float* Bs = T + kTileFacets * lda;               // [32][128]
  float* qs_all = Bs + kChunkK * 128;              // [8][32][QS]
  int* nbr_all = reinterpret_cast<int*>(qs_all + kWarps * 32 * QS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = qs_all + warp * 32 * QS;
  int* nbr = nbr_all + warp * 32;
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileFacets;
    for (int f = warp; f < kTileFacets; f += kWarps) {
      const int64_t t = r0 + f;
      float acc[MP][NC];
      float dvx[MP];
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        dvx[m] = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) acc[m][i] = 0.f;
      }
      if (t < p.rows) {
        const int e0 = p.rev_ptr[t], e1 = p.rev_ptr[t + 1];
        const float* vx = p.uvx + t * 2 * p.M + p.M;
        for (int eb = e0; eb < e1; eb += 32) {
          const int n_in = min(32, e1 - eb);
          __syncwarp();
          if (lane < n_in) {
            const int e = p.rev_edge[eb + lane];
            const int64_t n = e / p.K;
            const float* ux = p.uvx + n * 2 * p.M;
            const float iv = __ldg(p.inv + n);
            float a[MP];
            float mx = -INFINITY;
#pragma unroll
            for (int m = 0; m < MP; ++m)
              if (m < p.M) {
                a[m] = __ldg(ux + m) + __ldg(vx + m);
                mx = fmaxf(mx, a[m]);
              }
            float sum = 0.f;
#pragma unroll
            for (int m = 0; m < MP; ++m)
              if (m < p.M) {
                a[m] = expf(a[m] - mx);
                sum += a[m];
              }
            const float rs = iv / sum;  // q * inv_cnt[n]  => rows of gy are weighted as gz
            const float* de = p.da_edge + static_cast<int64_t>(e) * p.M;
#pragma unroll
            for (int m = 0; m < QS; ++m) {
              qs[lane * QS + m] = (m < p.M && m < MP) ? a[m < MP ? m : 0] * rs : 0.f;
            }
#pragma unroll
            for (int m = 0; m < MP; ++m)
              if (m < p.M) dvx[m] += __ldg(de + m);
            nbr[lane] = static_cast<int>(n);
          }
          __syncwarp();
          aggregate_rows<MP, NC>(p.gy, p.Cout, p.Cout, n_in, qs, nbr, lane, acc);
        }
#pragma unroll
        for (int m = 0; m < MP; ++m) {
          if (m < p.M) {
            const float s = warp_sum(dvx[m]);
            if (lane == 0) p.d_uvx[t * 2 * p.M + p.M + m] = s;
          }
        }
      }
      float* Tr = T + f * lda;
#pragma unroll
      for (int m = 0; m < MP; ++m) {
        if (m < p.M) {
#pragma unroll
          for (int i = 0; i < NC; ++i) {
            const int o = lane + 32 * i;
            if (o < p.Cout) Tr[m * p.Cout + o] = acc[m][i];
          }
        }
      }
      if (lane < lda - KK) Tr[KK + lane] = 0.f;
    }
    // ---- gx[:, 0:Cw] = T . W0 viewed as [(m,o)][c]
    for (int c0 = 0; c0 < p.Cw; c0 += 128) {
      const int ncols = min(128, p.Cw - c0);
      const TileGemmMap mp(ncols);
      float acc[4][4];
      tile_gemm(T, lda, KK, p.W0, p.Cw, c0, ncols, Bs, mp, acc);
      if (mp.ty < mp.TY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = mp.ty + mp.TY * i;
          const int64_t r = r0 + f;
          if (i < mp.RF && f < kTileFacets && r < p.rows) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + 4 * mp.tx + j;
              if (c < p.Cw) p.gx[r * p.Cin + c] = acc[i][j];
            }
          }
        }
      }
    }
    // channels outside the contraction window start from zero (the logits pass adds to them)
    if (p.Cin > p.Cw) {
      const int extra = p.Cin - p.Cw;
      for (int e = threadIdx.x; e < kTileFacets * extra; e += kThreads) {
        const int64_t r = r0 + e / extra;
        if (r < p.rows) p.gx[r * p.Cin + p.Cw + e % extra] = 0.f;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ weight / bias gradient pass
struct BwdWParams {
  const float* gy;
  const float* x;
  const int32_t* adj;
  const float* uvx;
  float* partW;   // [chunks][M][Cout][Cw]
  float* partB;   // [chunks][Cout]
  int64_t rows;
  int N, K, Cin, Cw, Cout, M;
  int bias_mask;
  int os;         // output-channel slice width handled by one CTA (multiple of 4)
  int nslices;    // ceil(Cout / os)
  int64_t tiles_per_chunk;
};

template <int MP, int NC>
__global__ void __launch_bounds__(kThreads)
bwd_w_kernel(const BwdWParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int QS = QStride<MP>::value;
  const int KK = p.M * p.Cw;
  const int lda = (KK + 3) & ~3;
  const int os = p.os;
  float* ACC = sm;                                 // [os][lda]
  float* S = ACC + os * lda;                       // [32][lda]
  float* GZ = S + kTileFacets * lda;               // [32][os]
  float* qs_all = GZ + kTileFacets * os;           // [8][32][QS]
  int* nbr_all = reinterpret_cast<int*>(qs_all + kWarps * 32 * QS);  // [8][32]
  float* flag = reinterpret_cast<float*>(nbr_all + kWarps * 32);     // [32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = qs_all + warp * 32 * QS;
  int* nbr = nbr_all + warp * 32;
  const int chunk = blockIdx.x;
  const int slice = blockIdx.y;
  const int o0 = slice * os;
  const int on = min(os, p.Cout - o0);
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;
  const int64_t t_begin = chunk * p.tiles_per_chunk;
  const int64_t t_end = min(ntiles, t_begin + p.tiles_per_chunk);

  for (int e = threadIdx.x; e < os * lda; e += kThreads) ACC[e] = 0.f;
  float gb_acc = 0.f;

  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    const int64_t r0 = tile * kTileFacets;
    __syncthreads();
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
      const float iv = cnt ? 1.f / static_cast<float>(cnt) : 0.f;
      for (int o = lane; o < os; o += 32)
        GZ[f * os + o] = (r < p.rows && o < on) ? iv * __ldg(p.gy + r * p.Cout + o0 + o) : 0.f;
      if (lane == 0) flag[f] = (r < p.rows && (cnt > 0 || !p.bias_mask)) ? 1.f : 0.f;
    }
    __syncthreads();
    // ACC[o][kk] += sum_f GZ[f][o] * S[f][kk]; thread owns 4x4 blocks, strided over the block grid
    const int nbk = lda / 4, nbo = os / 4;
    for (int blk = threadIdx.x; blk < nbk * nbo; blk += kThreads) {
      const int bo = blk / nbk, bk = blk % nbk;
      float a[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
#pragma unroll 4
      for (int f = 0; f < kTileFacets; ++f) {
        const float4 g = *reinterpret_cast<const float4*>(GZ + f * os + 4 * bo);
        const float4 sv = *reinterpret_cast<const float4*>(S + f * lda + 4 * bk);
        const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i][0] = fmaf(gv[i], sv.x, a[i][0]);
          a[i][1] = fmaf(gv[i], sv.y, a[i][1]);
          a[i][2] = fmaf(gv[i], sv.z, a[i][2]);
          a[i][3] = fmaf(gv[i], sv.w, a[i][3]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4* dst = reinterpret_cast<float4*>(ACC + (4 * bo + i) * lda + 4 * bk);
        float4 cur = *dst;
        cur.x += a[i][0], cur.y += a[i][1], cur.z += a[i][2], cur.w += a[i][3];
        *dst = cur;
      }
    }
    // bias gradient: thread o accumulates its column over the tile's facets
    if (threadIdx.x < on) {
      for (int f = 0; f < kTileFacets; ++f) {
        const int64_t r = r0 + f;
        if (r < p.rows) gb_acc = fmaf(flag[f], __ldg(p.gy + r * p.Cout + o0 + threadIdx.x), gb_acc);
      }
    }
  }
  __syncthreads();
  // write this CTA's partial: partW[chunk][m][o0+o][c] = ACC[o][m*Cw+c]
  float* pw = p.partW + static_cast<int64_t>(chunk) * p.M * p.Cout * p.Cw;
  for (int e = threadIdx.x; e < on * KK; e += kThreads) {
    const int o = e / KK, kk = e % KK;
    const int m = kk / p.Cw, c = kk % p.Cw;
    pw[(static_cast<int64_t>(m) * p.Cout + o0 + o) * p.Cw + c] = ACC[o * lda + kk];
  }
  if (threadIdx.x < on) p.partB[static_cast<int64_t>(chunk) * p.Cout + o0 + threadIdx.x] = gb_acc;
}

// ------------------------------------------------------------------ logits backward
// d_uvx[rows][2M] -> (a) the logit-window part of gx, thread per row;
//                    (b) gu, gv, gc partial sums per CTA (fixed order), thread per 4 outputs.
struct LogitsBwdParams {
  const float* x;
  const float* d_uvx;  // [rows][2M]
  const float* u;
  const float* v;
  float* gx;           // += on the window [Ca0, Ca0+Ca)
  float* part;         // [chunks][2M*Ca + M]
  int64_t rows;
  int Cin, Ca0, Ca, M;
  int64_t rows_per_chunk;
};

constexpr int kLogitRows = 32;

template <int OP>
__global__ void __launch_bounds__(128)
logits_bwd_x_kernel(const LogitsBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int O = 2 * p.M;
  const int Ca4 = (p.Ca + 3) & ~3;
  float* uv = sm;  // [O][Ca4]
  for (int e = threadIdx.x; e < O * Ca4; e += blockDim.x) {
    const int o = e / Ca4, cc = e % Ca4;
    uv[e] = cc < p.Ca ? ((o < p.M) ? p.u[o * p.Ca + cc] : p.v[(o - p.M) * p.Ca + cc]) : 0.f;
  }
  __syncthreads();
  const bool vec = (p.Cin % 4 == 0) && (p.Ca0 % 4 == 0) && (p.Ca % 4 == 0);
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < p.rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float d[OP];
#pragma unroll
    for (int o = 0; o < OP; ++o) d[o] = (o < O) ? __ldg(p.d_uvx + r * O + o) : 0.f;
    float* gr = p.gx + r * p.Cin + p.Ca0;
    for (int c0 = 0; c0 < Ca4; c0 += 4) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        if (o < O) {
          const float4 w = *reinterpret_cast<const float4*>(uv + o * Ca4 + c0);
          a.x = fmaf(d[o], w.x, a.x), a.y = fmaf(d[o], w.y, a.y), a.z = fmaf(d[o], w.z, a.z), a.w = fmaf(d[o], w.w, a.w);
        }
      }
      if (vec) {
        float4 cur = *reinterpret_cast<float4*>(gr + c0);
        cur.x += a.x, cur.y += a.y, cur.z += a.z, cur.w += a.w;
        *reinterpret_cast<float4*>(gr + c0) = cur;
      } else {
        if (c0 < p.Ca) gr[c0] += a.x;
        if (c0 + 1 < p.Ca) gr[c0 + 1] += a.y;
        if (c0 + 2 < p.Ca) gr[c0 + 2] += a.z;
        if (c0 + 3 < p.Ca) gr[c0 + 3] += a.w;
      }
    }
  }
}

constexpr int kLogitOwn = 8;  // output quads per thread: 2M * ceil(Ca/4) <= 32 * 64 = 8 * 256
__global__ void __launch_bounds__(kThreads)
logits_bwd_p_kernel(const LogitsBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int O = 2 * p.M;
  const int Ca4 = (p.Ca + 3) & ~3, nq = Ca4 / 4;
  float* xs = sm;                           // [32][Ca4]
  float* ds = xs + kLogitRows * Ca4;        // [32][O]
  float4 acc[kLogitOwn];
  int oo[kLogitOwn], qq[kLogitOwn];
#pragma unroll
  for (int i = 0; i < kLogitOwn; ++i) {
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int e = threadIdx.x + i * kThreads;
    oo[i] = e < O * nq ? e / nq : -1;
    qq[i] = e % nq;
  }
  float gc_acc = 0.f;
  const int64_t rb = static_cast<int64_t>(blockIdx.x) * p.rows_per_chunk;
  const int64_t re = min(p.rows, rb + p.rows_per_chunk);
  for (int64_t r0 = rb; r0 < re; r0 += kLogitRows) {
    const int nr = (re - r0 < kLogitRows) ? static_cast<int>(re - r0) : kLogitRows;
    __syncthreads();
    for (int e = threadIdx.x; e < kLogitRows * Ca4; e += kThreads) {
      const int rr = e / Ca4, cc = e % Ca4;
      xs[e] = (rr < nr && cc < p.Ca) ? __ldg(p.x + (r0 + rr) * p.Cin + p.Ca0 + cc) : 0.f;
    }
    for (int e = threadIdx.x; e < kLogitRows * O; e += kThreads) {
      const int rr = e / O;
      ds[e] = (rr < nr) ? __ldg(p.d_uvx + (r0 + rr) * O + e % O) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kLogitOwn; ++i) {
      if (oo[i] >= 0) {
        float4 a = acc[i];
#pragma unroll 8
        for (int rr = 0; rr < kLogitRows; ++rr) {
          const float dv = ds[rr * O + oo[i]];
          const float4 xv = *reinterpret_cast<const float4*>(xs + rr * Ca4 + 4 * qq[i]);
          a.x = fmaf(dv, xv.x, a.x), a.y = fmaf(dv, xv.y, a.y), a.z = fmaf(dv, xv.z, a.z), a.w = fmaf(dv, xv.w, a.w);
        }
        acc[i] = a;
      }
    }
    if (threadIdx.x < p.M) {
      for (int rr = 0; rr < kLogitRows; ++rr) gc_acc += ds[rr * O + threadIdx.x];
    }
  }
  const int nout = O * p.Ca;
  float* out = p.part + static_cast<int64_t>(blockIdx.x) * (nout + p.M);
#pragma unroll
  for (int i = 0; i < kLogitOwn; ++i) {
    if (oo[i] >= 0) {
      const int cc = 4 * qq[i];
      const float av[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (cc + j < p.Ca) out[oo[i] * p.Ca + cc + j] = av[j];
    }
  }
  if (threadIdx.x < p.M) out[nout + threadIdx.x] = gc_acc;
}

// out[e] = sum_p part[p*stride + e]   (fixed order)
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out,
                                       int64_t n, int P, int64_t stride) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a = 0.f;
    for (int q = 0; q < P; ++q) a += part[q * stride + e];
    out[e] = a;
  }
}

// few outputs, many partials: one warp per output, lane l sums partials l, l+32, ... and the lanes are
// combined by a butterfly (fixed order: deterministic)
__global__ void reduce_partials_warp_kernel(const float* __restrict__ part, float* __restrict__ out,
                                            int64_t n, int P, int64_t stride) {
  const int lane = threadIdx.x & 31;
  const int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (e >= n) return;
  float a = 0.f;
  for (int q = lane; q < P; q += 32) a += part[q * stride + e];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[e] = a;
}

int launch_reduce_partials(const float* part, float* out, int64_t n, int P, int64_t stride,
                           cudaStream_t st) {
  if (n <= 0) return FGC_OK;
  if (n <= 4096 && P >= 64) {
    reduce_partials_warp_kernel<<<static_cast<unsigned>((n * 32 + 255) / 256), 256, 0, st>>>(part, out, n, P, stride);
    FGC_LAUNCHED("reduce_partials_kernel");
    return FGC_OK;
  }
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  reduce_partials_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(part, out, n, P, stride);
  FGC_LAUNCHED("reduce_partials_kernel");
  return FGC_OK;
}

// ------------------------------------------------------------------ reverse adjacency
__global__ void rev_count_kernel(const int32_t* __restrict__ adj, int32_t* __restrict__ cnt,
                                 int64_t rows, int N, int K) {
  const int64_t total = rows * K;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int id = adj[e];
    if (id > 0 && id <= N) {
      const int64_t base = ((e / K) / N) * N;
      atomicAdd(cnt + base + id - 1, 1);
    }
  }
}

__global__ void rev_fill_kernel(const int32_t* __restrict__ adj, const int32_t* __restrict__ ptr,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ edges,
                                int64_t rows, int N, int K) {
  const int64_t total = rows * K;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int id = adj[e];
    if (id > 0 && id <= N) {
      const int64_t t = ((e / K) / N) * N + id - 1;
      const int pos = atomicAdd(cursor + t, 1);
      edges[ptr[t] + pos] = static_cast<int32_t>(e);
    }
  }
}

// integer atomics make the fill order arbitrary; sorting every segment restores a fixed order
__global__ void rev_sort_kernel(const int32_t* __restrict__ ptr, int32_t* __restrict__ edges,
                                int64_t rows) {
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < rows;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int a = ptr[t], b = ptr[t + 1];
    for (int i = a + 1; i < b; ++i) {
      const int key = edges[i];
      int j = i - 1;
      while (j >= a && edges[j] > key) {
        edges[j + 1] = edges[j];
        --j;
      }
      edges[j + 1] = key;
    }
  }
}

size_t reverse_adj_workspace(int64_t rows) {
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<int32_t*>(nullptr),
                                static_cast<int32_t*>(nullptr), static_cast<int>(rows + 1));
  return ws_bytes(rows + 1, 4) * 2 + align_up(scan_bytes, 256) + 512;
}

int build_reverse_adj(const int32_t* adj, int B, int N, int K, int32_t* rev_ptr, int32_t* rev_edge,
                      int64_t* nnz_out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(B) * N;
  FGC_REQUIRE(rows * K < (1ll << 31), "reverse adjacency: B*N*K must be < 2^31");
  Workspace ws(workspace, workspace_bytes);
  int32_t* cnt = ws.take<int32_t>(rows + 1);
  int32_t* cursor = ws.take<int32_t>(rows + 1);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt, rev_ptr, static_cast<int>(rows + 1));
  char* scan_tmp = ws.take<char>(scan_bytes + 256);
  FGC_REQUIRE(ws.ok(), "reverse adjacency: workspace too small");
  FGC_CUDA(cudaMemsetAsync(cnt, 0, (rows + 1) * 4, st));
  FGC_CUDA(cudaMemsetAsync(cursor, 0, (rows + 1) * 4, st));
  int64_t blocks = (rows * K + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  rev_count_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(adj, cnt, rows, N, K);
  FGC_LAUNCHED("rev_count_kernel");
  FGC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, cnt, rev_ptr,
                                         static_cast<int>(rows + 1), st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  rev_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(adj, rev_ptr, cursor, rev_edge, rows, N, K);
  FGC_LAUNCHED("rev_fill_kernel");
  int64_t sblocks = (rows + 255) / 256;
  if (sblocks > cap) sblocks = cap;
  if (sblocks < 1) sblocks = 1;
  rev_sort_kernel<<<static_cast<unsigned>(sblocks), 256, 0, st>>>(rev_ptr, rev_edge, rows);
  FGC_LAUNCHED("rev_sort_kernel");
  if (nnz_out) {
    int32_t nnz = 0;
    FGC_CUDA(cudaMemcpyAsync(&nnz, rev_ptr + rows, 4, cudaMemcpyDeviceToHost, st));
    FGC_CUDA(cudaStreamSynchronize(st));
    *nnz_out = nnz;
  }
  return FGC_OK;
}

// ------------------------------------------------------------------ host-side orchestration
static size_t smem_src(int MP, int M, int Cw, int Cout) {
  const int QS = (MP + 3) / 4 * 4;
  const size_t lda = (M * Cw + 3) & ~3, ldg = (Cout + 3) & ~3;
  const size_t shared = std::max<size_t>(kChunkK * 128, 2 * kWarps * 32 * QS);   // weight chunk | assignment tables
  return (kTileFacets * lda + kTileFacets * ldg + shared + kWarps * 32) * 4;
}
static size_t smem_tgt(int MP, int M, int Cout) {
  const int QS = (MP + 3) / 4 * 4;
  const size_t lda = (M * Cout + 3) & ~3;
  return (kTileFacets * lda + kChunkK * 128 + kWarps * 32 * QS + kWarps * 32) * 4;
}
static size_t smem_w(int MP, int M, int Cw, int os) {
  const int QS = (MP + 3) / 4 * 4;
  const size_t lda = (M * Cw + 3) & ~3;
  return (os * lda + kTileFacets * lda + kTileFacets * os + kWarps * 32 * QS + kWarps * 32 + kTileFacets) * 4;
}

constexpr size_t kSmemLimit = 227 * 1024;

struct BwdPlan {
  int os, nslices, chunks;
  int64_t tiles_per_chunk;
  int lchunks;
  int64_t rows_per_lchunk;
};

static int make_plan(const fgc_conv_shape* s, BwdPlan* pl) {
  const int MP = pick_mp(s->M);
  int os = (s->Cout + 3) & ~3;
  while (os > 4 && smem_w(MP, s->M, s->Cw, os) > kSmemLimit - 2048) os -= 4;
  if (smem_w(MP, s->M, s->Cw, os) > kSmemLimit) return FGC_ERR_UNSUPPORTED;
  pl->os = os;
  pl->nslices = (s->Cout + os - 1) / os;
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  const int64_t ntiles = (rows + kTileFacets - 1) / kTileFacets;
  int64_t chunks = num_sms() / pl->nslices;
  if (chunks < 1) chunks = 1;
  if (chunks > ntiles) chunks = ntiles;
  if (chunks < 1) chunks = 1;
  pl->tiles_per_chunk = (ntiles + chunks - 1) / chunks;
  pl->chunks = static_cast<int>((ntiles + pl->tiles_per_chunk - 1) / pl->tiles_per_chunk);
  if (pl->chunks < 1) pl->chunks = 1;
  int64_t lch = static_cast<int64_t>(num_sms()) * 2;
  const int64_t lt = (rows + kLogitRows - 1) / kLogitRows;
  if (lch > lt) lch = lt;
  if (lch < 1) lch = 1;
  pl->rows_per_lchunk = ((lt + lch - 1) / lch) * kLogitRows;
  pl->lchunks = static_cast<int>((rows + pl->rows_per_lchunk - 1) / pl->rows_per_lchunk);
  if (pl->lchunks < 1) pl->lchunks = 1;
  return FGC_OK;
}

size_t conv_bwd_workspace(const fgc_conv_shape* s) {
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  BwdPlan pl;
  if (make_plan(s, &pl) != FGC_OK) return 0;
  const size_t nW = static_cast<size_t>(s->M) * s->Cout * s->Cw;
  size_t b = 0;
  b += ws_bytes(rows * 2 * s->M, 4);               // uvx
  b += ws_bytes(rows * 2 * s->M, 4);               // d_uvx
  b += ws_bytes(nW, 4);                            // Wd
  b += ws_bytes(rows * s->K * s->M, 4);            // da_edge
  b += ws_bytes(rows, 4);                          // inv
  const bool mma_ok = conv_mma_supported(s->Cin, s->Cw, s->Cout, s->M, s->K);
  size_t wparts = std::max<size_t>(pl.chunks, bwd_w_tc_grid(rows));
  if (mma_ok) wparts = std::max<size_t>(wparts, bwd_w_mma_grid(rows, s->M));
  const size_t bparts = mma_ok ? std::max<size_t>(wparts, prep_image_blocks()) : wparts;
  b += ws_bytes(wparts * nW, 4);        // partW
  b += ws_bytes(bparts * s->Cout, 4);   // partB
  b += ws_bytes(4, 4);                  // absmax scratch
  b += ws_bytes(static_cast<size_t>(pl.lchunks) * (2 * s->M * s->Ca + s->M), 4);  // logits partials
  b += ws_bytes(conv_fwd_tc_workspace(s->Cw > s->Cout ? s->Cw : s->Cout, s->M), 1);   // TC weight image
  if (conv_mma_supported(s->Cin, s->Cw, s->Cout, s->M, s->K)) b += 2 * (conv_mma_workspace(rows) + 256);  // gy, x images
  return b + 1024;
}

template <int MP, int NC>
static int run_src(const BwdSrcParams& p, cudaStream_t st) {
  const size_t smem = smem_src(MP, p.M, p.Cw, p.Cout);
  if (smem > kSmemLimit) {
    set_error("conv_bwd(src): shape needs %zu bytes of shared memory", smem);
    return FGC_ERR_UNSUPPORTED;
  }
  FGC_CUDA(cudaFuncSetAttribute(bwd_src_kernel<MP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  FGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bwd_src_kernel<MP, NC>, kThreads, smem));
  if (occ < 1) occ = 1;
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;
  int64_t grid = static_cast<int64_t>(num_sms()) * occ;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  bwd_src_kernel<MP, NC><<<static_cast<unsigned>(grid), kThreads, smem, st>>>(p);
  FGC_LAUNCHED("bwd_src_kernel");
  return FGC_OK;
}

template <int MP, int NC>
static int run_tgt(const BwdTgtParams& p, cudaStream_t st) {
  const size_t smem = smem_tgt(MP, p.M, p.Cout);
  if (smem > kSmemLimit) {
    set_error("conv_bwd(tgt): shape needs %zu bytes of shared memory", smem);
    return FGC_ERR_UNSUPPORTED;
  }
  FGC_CUDA(cudaFuncSetAttribute(bwd_tgt_kernel<MP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  FGC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bwd_tgt_kernel<MP, NC>, kThreads, smem));
  if (occ < 1) occ = 1;
  const int64_t ntiles = (p.rows + kTileFacets - 1) / kTileFacets;
  int64_t grid = static_cast<int64_t>(num_sms()) * occ;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  bwd_tgt_kernel<MP, NC><<<static_cast<unsigned>(grid), kThreads, smem, st>>>(p);
  FGC_LAUNCHED("bwd_tgt_kernel");
  return FGC_OK;
}

template <int MP, int NC>
static int run_w(const BwdWParams& p, int chunks, cudaStream_t st) {
  const size_t smem = smem_w(MP, p.M, p.Cw, p.os);
  FGC_CUDA(cudaFuncSetAttribute(bwd_w_kernel<MP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(chunks, p.nslices);
  bwd_w_kernel<MP, NC><<<grid, kThreads, smem, st>>>(p);
  FGC_LAUNCHED("bwd_w_kernel");
  return FGC_OK;
}

// set by the host-buffer entry point (c_api.cu): recorded on the stream as soon as gx is final
thread_local cudaEvent_t g_gx_ready_event = nullptr;

int conv_bwd(const fgc_conv_shape* s, const float* gy, const float* x, const int32_t* adj,
             const int32_t* rev_ptr, const int32_t* rev_edge, const float* W0, const float* u,
             const float* v, const float* c, float* gx, float* gW0, float* gb, float* gu, float* gv,
             float* gc, int bias_mask, void* workspace, size_t workspace_bytes, cudaStream_t st,
             const int32_t* radj, int Kr, const void* rplan, const void* fplan, const void* fwd_ws,
             size_t fwd_ws_bytes) {
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  BwdPlan pl;
  if (make_plan(s, &pl) != FGC_OK) {
    set_error("conv_bwd: M*Cw = %d too large for the shared-memory accumulators", s->M * s->Cw);
    return FGC_ERR_UNSUPPORTED;
  }
  const size_t nW = static_cast<size_t>(s->M) * s->Cout * s->Cw;
  const int nL = 2 * s->M * s->Ca + s->M;
  Workspace ws(workspace, workspace_bytes);
  float* uvx_own = ws.take<float>(rows * 2 * s->M);
  float* d_uvx = ws.take<float>(rows * 2 * s->M);
  float* Wd = ws.take<float>(nW);
  float* da_edge = ws.take<float>(rows * s->K * s->M);
  float* inv = ws.take<float>(rows);
  const bool mma_ok = conv_mma_supported(s->Cin, s->Cw, s->Cout, s->M, s->K);
  size_t wparts = std::max<size_t>(pl.chunks, bwd_w_tc_grid(rows));
  if (mma_ok) wparts = std::max<size_t>(wparts, bwd_w_mma_grid(rows, s->M));
  const size_t bparts = mma_ok ? std::max<size_t>(wparts, prep_image_blocks()) : wparts;
  float* partW = ws.take<float>(wparts * nW);
  float* partB = ws.take<float>(bparts * s->Cout);
  unsigned* maxbits = ws.take<unsigned>(4);
  float* partL = ws.take<float>(static_cast<size_t>(pl.lchunks) * nL);
  char* wimg = ws.take<char>(conv_fwd_tc_workspace(s->Cw > s->Cout ? s->Cw : s->Cout, s->M));
  static const bool mma_disabled = getenv("FGC_DISABLE_MMA") != nullptr || getenv("FGC_DISABLE_TC") != nullptr;
  const bool tgt_mma = !mma_disabled && rplan != nullptr && radj != nullptr &&
                       bwd_tgt_mma_supported(s->Cin, s->Cw, s->Cout, s->M, Kr);
  const bool src_mma = !mma_disabled && fplan != nullptr && bwd_src_mma_supported(s->Cin, s->Cw, s->Cout, s->M, s->K);
  char* gyimg = (tgt_mma || src_mma) ? ws.take<char>(conv_mma_workspace(rows)) : nullptr;
  char* ximg = src_mma ? ws.take<char>(conv_mma_workspace(rows)) : nullptr;
  FGC_REQUIRE(ws.ok(), "conv_bwd: workspace too small (%zu bytes given)", workspace_bytes);
  static const bool tc_disabled = getenv("FGC_DISABLE_TC") != nullptr;

  // Products of the planned forward (logits uvx, fp16 image of x and its scale) are reused when the
  // caller hands back the forward's workspace untouched (what an autograd context saves).
  const float* uvx = uvx_own;
  bool ximg_saved = false;
  int rc = FGC_OK;
  if (fwd_ws != nullptr && src_mma) {
    FwdSaved sv;
    rc = conv_fwd_saved_views(s, fwd_ws, fwd_ws_bytes, &sv);
    if (rc) return rc;
    uvx = sv.uvx, ximg = const_cast<char*>(sv.ximg), ximg_saved = true;
  } else {
    rc = launch_assign_logits(s, x, u, v, c, uvx_own, st);
    if (rc) return rc;
  }
  const int MP = pick_mp(s->M);
  const bool tc_all = !tc_disabled && s->Cout % 4 == 0 && bwd_tgt_tc_supported(s->Cw, s->Cout, s->M) &&
                      bwd_src_tc_supported(s->Cw, s->Cout, s->M, s->Cin);
  if (tc_all) {
    rc = launch_prep_w_image_t(W0, wimg, s->M, s->Cw, st);
    if (rc) return rc;
    const float* wunscale = reinterpret_cast<const float*>(wimg + static_cast<size_t>(s->M) * 2 * s->Cw * 128);
    if (gyimg != nullptr) {   // the gy image serves the source pass, the target pass and the weight gradient
      // (with a forward plan the same pass over gy also leaves the bias-gradient partials)
      if (src_mma) rc = launch_prep_image(gy, s->Cout, rows, gyimg, st, conv_plan_inv(fplan, rows, s->K, s->M), bias_mask, partB);
      else rc = launch_prep_image(gy, s->Cout, rows, gyimg, st);
      if (rc) return rc;
    }
    if (src_mma) {
      if (!ximg_saved) {
        rc = launch_prep_image(x, s->Cin, rows, ximg, st);
        if (rc) return rc;
      }
      rc = launch_bwd_src_mma(gy, uvx, adj, fplan, ximg, gyimg, wimg, da_edge, d_uvx, rows, s->N, s->K, s->M, st);
      if (rc) return rc;
      FGC_CUDA(cudaMemcpyAsync(inv, conv_plan_inv(fplan, rows, s->K, s->M), rows * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    } else {
      rc = launch_bwd_src_tc(gy, x, adj, uvx, wimg, wunscale, da_edge, d_uvx, inv, rows, s->N, s->K, s->Cin, s->M, st);
      if (rc) return rc;
    }
  } else {
    const int total = static_cast<int>(nW);
    permute_w_ds_kernel<<<(total + 255) / 256, 256, 0, st>>>(W0, Wd, s->M, s->Cout, s->Cw);
    FGC_LAUNCHED("permute_w_ds_kernel");
    BwdSrcParams p{gy, x, adj, uvx, Wd, da_edge, d_uvx, inv, rows, s->N, s->K, s->Cin, s->Cw, s->Cout, s->M};
#define FGC_CALL(MPV, NCV) rc = run_src<MPV, NCV>(p, st)
    FGC_DISPATCH_MP_NC(MP, pick_nc(s->Cw), FGC_CALL);
#undef FGC_CALL
    if (rc) return rc;
  }
  if (tgt_mma && tc_all) {
    if (s->Cin > s->Cw) FGC_CUDA(cudaMemsetAsync(gx, 0, rows * s->Cin * sizeof(float), st));
    rc = launch_bwd_tgt_mma(gy, uvx, da_edge, inv, rev_ptr, rev_edge, radj, Kr, rplan, gx, d_uvx, rows, s->N, s->Cin,
                            s->Cout, s->M, wimg, gyimg, st);
    if (rc) return rc;
  } else if (!tc_disabled && s->Cout % 4 == 0 && bwd_tgt_tc_supported(s->Cw, s->Cout, s->M)) {
    if (s->Cin > s->Cw) FGC_CUDA(cudaMemsetAsync(gx, 0, rows * s->Cin * sizeof(float), st));
    rc = launch_bwd_tgt_tc(gy, uvx, tc_all ? nullptr : W0, da_edge, inv, rev_ptr, rev_edge, gx, d_uvx, rows,
                           s->N, s->K, s->Cin, s->Cw, s->Cout, s->M, wimg, st);
    if (rc) return rc;
  } else {
    BwdTgtParams p{gy, uvx, W0, da_edge, inv, rev_ptr, rev_edge, gx, d_uvx, rows,
                   s->N, s->K, s->Cin, s->Cw, s->Cout, s->M};
#define FGC_CALL(MPV, NCV) rc = run_tgt<MPV, NCV>(p, st)
    FGC_DISPATCH_MP_NC(MP, pick_nc(s->Cout), FGC_CALL);
#undef FGC_CALL
    if (rc) return rc;
  }
  // ---- logit-window part of gx first: gx is final after this kernel (the host-buffer entry point
  //      starts its device->host copy here, overlapping the weight-gradient passes below)
  static const bool slow_logits = getenv("FGC_DISABLE_FAST_LOGITS") != nullptr;
  const bool fast_logits = !slow_logits && logits_fast_supported(s->Cin, s->Ca0, s->Ca, s->M);
  LogitsBwdParams lp{x, d_uvx, u, v, gx, partL, rows, s->Cin, s->Ca0, s->Ca, s->M, pl.rows_per_lchunk};
  const int O2 = 2 * s->M, Ca4 = (s->Ca + 3) & ~3;
  if (fast_logits) {
    rc = launch_logits_bwd_x_fast(d_uvx, u, v, gx, rows, s->Cin, s->Ca0, s->Ca, s->M, st);
    if (rc) return rc;
  } else {
    const size_t smem_x = static_cast<size_t>(O2) * Ca4 * 4;
    int64_t bx = (rows + 127) / 128;
    const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const unsigned g = static_cast<unsigned>(bx);
    if (O2 <= 8) logits_bwd_x_kernel<8><<<g, 128, smem_x, st>>>(lp);
    else if (O2 <= 16) logits_bwd_x_kernel<16><<<g, 128, smem_x, st>>>(lp);
    else if (O2 <= 18) logits_bwd_x_kernel<18><<<g, 128, smem_x, st>>>(lp);
    else logits_bwd_x_kernel<32><<<g, 128, smem_x, st>>>(lp);
    FGC_LAUNCHED("logits_bwd_x_kernel");
  }
  if (g_gx_ready_event != nullptr) FGC_CUDA(cudaEventRecord(g_gx_ready_event, st));
  int wchunks = pl.chunks, bchunks = -1;
  if (src_mma && tc_all) {
    rc = launch_bwd_w_mma(uvx, adj, fplan, ximg, gyimg, partW, rows, s->N, s->K, s->M, st);
    if (rc) return rc;
    wchunks = bwd_w_mma_grid(rows, s->M);
    bchunks = prep_image_blocks();
  } else if (!tc_disabled && bwd_w_tc_supported(s->Cw, s->Cout, s->M, s->Cin)) {
    // max|x| and max|gy| were already reduced for the fp16 images of the planned passes
    const unsigned* xmax = src_mma ? conv_mma_image_maxbits(ximg, rows) : nullptr;
    const unsigned* gmax = gyimg != nullptr && tc_all ? conv_mma_image_maxbits(gyimg, rows) : nullptr;
    rc = launch_bwd_w_tc(gy, x, adj, uvx, partW, partB, maxbits, xmax, gmax, rows, s->N, s->K, s->Cin, s->M, bias_mask,
                         st);
    if (rc) return rc;
    wchunks = bwd_w_tc_grid(rows);
  } else {
    BwdWParams p{gy, x, adj, uvx, partW, partB, rows, s->N, s->K, s->Cin, s->Cw, s->Cout, s->M,
                 bias_mask, pl.os, pl.nslices, pl.tiles_per_chunk};
#define FGC_CALL(MPV, NCV) rc = run_w<MPV, NCV>(p, pl.chunks, st)
    FGC_DISPATCH_MP_NC(MP, pick_nc(s->Cw), FGC_CALL);
#undef FGC_CALL
    if (rc) return rc;
  }
  if (fast_logits) {
    rc = launch_logits_bwd_p_fast(x, d_uvx, partL, rows, pl.rows_per_lchunk, pl.lchunks, s->Cin, s->Ca0, s->Ca, s->M,
                                  st);
    if (rc) return rc;
  } else {
    const size_t smem_p = (static_cast<size_t>(kLogitRows) * Ca4 + kLogitRows * O2) * 4;
    FGC_CUDA(cudaFuncSetAttribute(logits_bwd_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    logits_bwd_p_kernel<<<pl.lchunks, kThreads, smem_p, st>>>(lp);
    FGC_LAUNCHED("logits_bwd_p_kernel");
  }
  rc = launch_reduce_partials(partW, gW0, nW, wchunks, nW, st);
  if (rc) return rc;
  rc = launch_reduce_partials(partB, gb, s->Cout, bchunks > 0 ? bchunks : wchunks, s->Cout, st);
  if (rc) return rc;
  const int nUV = s->M * s->Ca;
  rc = launch_reduce_partials(partL, gu, nUV, pl.lchunks, nL, st);
  if (rc) return rc;
  rc = launch_reduce_partials(partL + nUV, gv, nUV, pl.lchunks, nL, st);
  if (rc) return rc;
  rc = launch_reduce_partials(partL + 2 * nUV, gc, s->M, pl.lchunks, nL, st);
  return rc;
}

}  // namespace fgc
