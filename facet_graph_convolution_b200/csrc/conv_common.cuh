// Device-side building blocks shared by the forward and backward facet-graph convolution
// kernels: the per-facet soft-assignment phase and the q-weighted row aggregation.
#pragma once

#include "common.cuh"

namespace fgc {

constexpr int kTileFacets = 32;   // facets per CTA tile
constexpr int kThreads = 256;     // 8 warps
constexpr int kWarps = kThreads / 32;
constexpr int kChunkK = 32;       // rows of the streamed weight chunk

template <int MP>
struct QStride {
  static constexpr int value = (MP + 3) / 4 * 4;
};

// Lane k (< K) of the calling warp evaluates the soft assignment of neighbour slot k of facet
// row r:  q[k,:] = softmax_m( uxc[r,:] + (valid ? vx[row(k),:] : 0) )   (model.py:74-95).
// qs[k*QS + m] receives q, nbr[k] the gathered row (global row index) or -1 for padding.
// Returns cnt = number of non-zero adjacency entries (model.py:436), warp-uniform.
template <int MP>
__device__ __forceinline__ int facet_assign(const int32_t* __restrict__ adj,
                                            const float* __restrict__ uvx, int64_t r, int64_t base,
                                            int N, int K, int M, float* qs, int* nbr, int lane,
                                            float scale = 1.f) {
  constexpr int QS = QStride<MP>::value;
  int id = 0;
  if (lane < K) id = __ldg(adj + r * K + lane);
  const int cnt = __popc(__ballot_sync(0xffffffffu, id != 0));
  if (lane < K) {
    const bool valid = (id > 0) && (id <= N);  // out-of-range ids are clamped to padding
    const int64_t row = valid ? base + id - 1 : -1;
    const float* ux = uvx + r * (2 * M);
    const float* vx = uvx + (valid ? row : r) * (2 * M) + M;
    float a[MP];
    float mx = -INFINITY;
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      if (m < M) {
        a[m] = __ldg(ux + m) + (valid ? __ldg(vx + m) : 0.f);
        mx = fmaxf(mx, a[m]);
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int m = 0; m < MP; ++m) {
      if (m < M) {
        a[m] = expf(a[m] - mx);
        sum += a[m];
      }
    }
    const float rs = scale / sum;
#pragma unroll
    for (int m = 0; m < QS; ++m) qs[lane * QS + m] = (m < M && m < MP) ? a[m < MP ? m : 0] * rs : 0.f;
    nbr[lane] = static_cast<int>(row);
  }
  __syncwarp();
  return cnt;
}

// acc[m][i] += sum_k q[k,m] * rows[nbr[k]][lane + 32 i]     (padding slots skipped: adds 0)
template <int MP, int NC>
__device__ __forceinline__ void aggregate_rows(const float* __restrict__ rows, int row_stride, int C,
                                               int nk, const float* qs, const int* nbr, int lane,
                                               float (&acc)[MP][NC]) {
  constexpr int QS = QStride<MP>::value;
  for (int k = 0; k < nk; ++k) {
    const int j = nbr[k];
    if (j < 0) continue;  // warp-uniform
    const float* xr = rows + static_cast<int64_t>(j) * row_stride;
    float xv[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = lane + 32 * i;
      xv[i] = (c < C) ? __ldg(xr + c) : 0.f;
    }
    float q[QS];
#pragma unroll
    for (int m4 = 0; m4 < QS; m4 += 4) {
      const float4 t = *reinterpret_cast<const float4*>(qs + k * QS + m4);
      q[m4] = t.x, q[m4 + 1] = t.y, q[m4 + 2] = t.z, q[m4 + 3] = t.w;
    }
#pragma unroll
    for (int m = 0; m < MP; ++m)
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[m][i] = fmaf(q[m], xv[i], acc[m][i]);
  }
}

// Dispatch helpers: MP = compile-time capacity for M, NC = channels per lane.
inline int pick_mp(int M) { return M <= 4 ? 4 : (M <= 8 ? 8 : (M == 9 ? 9 : 16)); }
inline int pick_nc(int C) { return C <= 32 ? 1 : (C <= 64 ? 2 : (C <= 128 ? 4 : 8)); }

#define FGC_DISPATCH_MP_NC(MPV, NCV, CALL)                       \
  do {                                                           \
    switch ((MPV) * 16 + (NCV)) {                                \
      case 4 * 16 + 1: { CALL(4, 1); } break;                    \
      case 4 * 16 + 2: { CALL(4, 2); } break;                    \
      case 4 * 16 + 4: { CALL(4, 4); } break;                    \
      case 4 * 16 + 8: { CALL(4, 8); } break;                    \
      case 8 * 16 + 1: { CALL(8, 1); } break;                    \
      case 8 * 16 + 2: { CALL(8, 2); } break;                    \
      case 8 * 16 + 4: { CALL(8, 4); } break;                    \
      case 8 * 16 + 8: { CALL(8, 8); } break;                    \
      case 9 * 16 + 1: { CALL(9, 1); } break;                    \
      case 9 * 16 + 2: { CALL(9, 2); } break;                    \
      case 9 * 16 + 4: { CALL(9, 4); } break;                    \
      case 9 * 16 + 8: { CALL(9, 8); } break;                    \
      case 16 * 16 + 1: { CALL(16, 1); } break;                  \
      case 16 * 16 + 2: { CALL(16, 2); } break;                  \
      case 16 * 16 + 4: { CALL(16, 4); } break;                  \
      case 16 * 16 + 8: { CALL(16, 8); } break;                  \
      default:                                                   \
        ::fgc::set_error("unsupported M/C combination");         \
        return FGC_ERR_UNSUPPORTED;                              \
    }                                                            \
  } while (0)

// Tile GEMM shared by forward and backward:  out[f][o] = sum_kk A[f][kk] * Bm[kk][o]
//   A   : shared memory, kTileFacets rows, row stride lda (multiple of 4), KKp columns (mult. of 4)
//   Bm  : global memory, row-major [KK][ldb]; rows >= KK are treated as zero
//   Bs  : shared staging buffer kChunkK x 128 floats
// Thread t owns outputs (facet ty + TY*i, column o0 + 4*tx + j) for i < 4, j < 4.
struct TileGemmMap {
  int TX, TY, RF, tx, ty;
  __device__ TileGemmMap(int ncols) {
    TX = (ncols + 3) / 4;
    TY = kThreads / TX;
    if (TY > kTileFacets) TY = kTileFacets;
    RF = (kTileFacets + TY - 1) / TY;
    tx = threadIdx.x % TX;
    ty = threadIdx.x / TX;
  }
};

__device__ __forceinline__ void tile_gemm(const float* A, int lda, int KK, const float* __restrict__ Bm,
                                          int ldb, int o0, int ncols, float* Bs,
                                          const TileGemmMap& mp, float (&acc)[4][4]) {
  const int KKp = (KK + 3) & ~3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool active = mp.ty < mp.TY;
  for (int kk0 = 0; kk0 < KKp; kk0 += kChunkK) {
    __syncthreads();
    for (int e = threadIdx.x; e < kChunkK * 128; e += kThreads) {
      const int kc = e >> 7, o = e & 127;
      const int kk = kk0 + kc;
      float v = 0.f;
      if (kk < KK && o < ncols) v = __ldg(Bm + static_cast<int64_t>(kk) * ldb + o0 + o);
      Bs[e] = v;
    }
    __syncthreads();
    if (active) {
      const int kend = min(kChunkK, KKp - kk0);
      for (int kc = 0; kc < kend; kc += 4) {
        float4 b4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          b4[u] = *reinterpret_cast<const float4*>(Bs + (kc + u) * 128 + 4 * mp.tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = mp.ty + mp.TY * i;
          if (i < mp.RF && f < kTileFacets) {
            const float4 a = *reinterpret_cast<const float4*>(A + f * lda + kk0 + kc);
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[i][0] = fmaf(av[u], b4[u].x, acc[i][0]);
              acc[i][1] = fmaf(av[u], b4[u].y, acc[i][1]);
              acc[i][2] = fmaf(av[u], b4[u].z, acc[i][2]);
              acc[i][3] = fmaf(av[u], b4[u].w, acc[i][3]);
            }
          }
        }
      }
    }
  }
}

}  // namespace fgc
