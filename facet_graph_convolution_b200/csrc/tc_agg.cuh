// Shared aggregation stage of the tensor-core kernels: soft assignments + q-weighted sum of gathered
// rows for the 4 facets of a warp (8 lanes per facet, each lane owns channels 4l..4l+3 and
// 32+4l..32+4l+3 of the 64-wide rows), packed fp32x2 FMAs.
#pragma once

#include "common.cuh"
#include "tc_common.cuh"

namespace fgc {

constexpr int MODE_FWD = 0;  // rows gathered through the adjacency (source-centric)
constexpr int MODE_TGT = 1;  // rows gathered through the reversed adjacency (target-centric)
constexpr int kQK = 16;      // neighbour slots per assignment round

struct AggSrc {
  const float* x;        // gathered rows (x for FWD, gy for TGT), row stride ldx, 64 channels used
  int ldx;
  const int32_t* adj;    // FWD: adj[rows][K]
  const float* uvx;      // [rows][2M]: own-row logits | neighbour logits
  int N, K;
  int64_t rows;
  // TGT only
  const int32_t* rev_ptr;
  const int32_t* rev_edge;
  const float* inv;      // inv_cnt of every source row
  const float* da_edge;  // [rows*K][M]
};

template <int M>
struct AggQ {
  static constexpr int MQ = (M + 3) & ~3;            // q row stride in floats
  static constexpr int QS_FLOATS = 4 * kQK * MQ;     // per warp
  static constexpr int NBR_INTS = 4 * kQK;           // per warp
};

// list bounds of this lane's facet and the warp-uniform number of list entries to walk
template <int MODE>
__device__ __forceinline__ int agg_list_bounds(const AggSrc& p, int64_t r, int& lst0, int& lst1) {
  lst0 = 0, lst1 = 0;
  if constexpr (MODE == MODE_FWD) {
    lst1 = p.K;
  } else {
    if (r < p.rows) {
      lst0 = __ldg(p.rev_ptr + r);
      lst1 = __ldg(p.rev_ptr + r + 1);
    }
  }
  int nround = lst1 - lst0;
  nround = max(nround, __shfl_xor_sync(0xffffffffu, nround, 8));
  nround = max(nround, __shfl_xor_sync(0xffffffffu, nround, 16));
  return nround;
}

// soft assignments of one round (list entries kb..kb+nk-1 of the warp's 4 facets): lane per
// (facet, slot) pair, pair = lane + 32h, facet = pair / 16.  qs[f][k][0..M) receives q (already
// multiplied by inv_cnt[source] in TGT mode), nbr[f][k] the gathered row, -1 = padding,
// -2 = non-zero id outside the patch (counts as a neighbour, contributes nothing).
template <int M, int MODE>
__device__ __forceinline__ void agg_assign_round(const AggSrc& p, int64_t wrow0, int kb, int nk, int lst0,
                                                 int lst1, float* qs, int* nbr, int lane, float (&dv)[2][M]) {
  constexpr int MQ = AggQ<M>::MQ;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = (lane >> 4) + 2 * h, k = lane & 15;
    const int64_t rf = wrow0 + f;
    const int f0 = __shfl_sync(0xffffffffu, lst0, f * 8);
    const int f1 = __shfl_sync(0xffffffffu, lst1, f * 8);
    int row = -1;
    float a[M];
    bool have = false;
    if (k < nk && rf < p.rows && f0 + kb + k < f1) {
      have = true;
      const float* ux;
      const float* vx;
      bool vvalid = true;
      if constexpr (MODE == MODE_FWD) {
        const int id = __ldg(p.adj + rf * p.K + kb + k);
        const int64_t base = (rf / p.N) * p.N;
        vvalid = id > 0 && id <= p.N;
        row = vvalid ? static_cast<int>(base + id - 1) : (id != 0 ? -2 : -1);
        ux = p.uvx + rf * (2 * M);
        vx = p.uvx + (vvalid ? static_cast<int64_t>(row) : rf) * (2 * M) + M;
      } else {
        const int e = __ldg(p.rev_edge + f0 + kb + k);
        row = e / p.K;  // source facet of the in-edge
        ux = p.uvx + static_cast<int64_t>(row) * (2 * M);
        vx = p.uvx + rf * (2 * M) + M;
        const float* de = p.da_edge + static_cast<int64_t>(e) * M;
#pragma unroll
        for (int m = 0; m < M; ++m) dv[h][m] += __ldg(de + m);
      }
#pragma unroll
      for (int m = 0; m < M; ++m) a[m] = __ldg(ux + m) + (vvalid ? __ldg(vx + m) : 0.f);
      float mx = a[0];
#pragma unroll
      for (int m = 1; m < M; ++m) mx = fmaxf(mx, a[m]);
      float sum = 0.f;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        a[m] = __expf(a[m] - mx);
        sum += a[m];
      }
      float rs = 1.f / sum;
      if constexpr (MODE == MODE_TGT) rs *= __ldg(p.inv + row);  // gy rows weighted as gz
#pragma unroll
      for (int m = 0; m < M; ++m) a[m] *= rs;
    }
    if (!have) {
#pragma unroll
      for (int m = 0; m < M; ++m) a[m] = 0.f;
    }
    float* qd = qs + (f * kQK + k) * MQ;
#pragma unroll
    for (int m = 0; m < MQ; ++m) qd[m] = (m < M) ? a[m < M ? m : 0] : 0.f;
    nbr[f * kQK + k] = row;
  }
}

// the two float4 of gathered row j this lane owns, as 4 channel pairs (zeros for padding)
__device__ __forceinline__ void agg_load_row(const AggSrc& p, int j, int gl, float2 (&xp)[4]) {
  float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
  if (j >= 0) {
    const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<int64_t>(j) * p.ldx);
    x0 = __ldg(xr + gl);
    x1 = __ldg(xr + 8 + gl);
  }
  xp[0] = make_float2(x0.x, x0.y), xp[1] = make_float2(x0.z, x0.w);
  xp[2] = make_float2(x1.x, x1.y), xp[3] = make_float2(x1.z, x1.w);
}

// acc[m][i] += sum over the facet's neighbour list of q[.,m] * row[channel pair i]
// wrow0 = global row of the warp's first facet; lanes 8g..8g+7 serve facet wrow0 + g.
// cnt   = number of non-padding list entries (FWD: non-zero adjacency ids).
// dv    = TGT: per-lane partial sums of da_edge (pair lane+32h belongs to facet (lane>>4)+2h).
template <int M, int MODE>
__device__ __forceinline__ void tc_aggregate(const AggSrc& p, int64_t wrow0, float* qs, int* nbr, int lane,
                                             float2 (&acc)[M][4], int& cnt, float (&dv)[2][M]) {
  constexpr int MQ = AggQ<M>::MQ;
  const int grp = lane >> 3, gl = lane & 7;
  int lst0, lst1;
  const int nround = agg_list_bounds<MODE>(p, wrow0 + grp, lst0, lst1);
  for (int kb = 0; kb < nround; kb += kQK) {
    const int nk = min(kQK, nround - kb);
    __syncwarp();
    agg_assign_round<M, MODE>(p, wrow0, kb, nk, lst0, lst1, qs, nbr, lane, dv);
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
      const int j = nbr[grp * kQK + k];
      cnt += (j != -1);
      float2 xp[4];
      agg_load_row(p, j, gl, xp);
      const float* qk = qs + (grp * kQK + k) * MQ;
      float q[MQ];
#pragma unroll
      for (int m4 = 0; m4 < MQ; m4 += 4) {
        const float4 t = *reinterpret_cast<const float4*>(qk + m4);
        q[m4] = t.x, q[m4 + 1] = t.y, q[m4 + 2] = t.z, q[m4 + 3] = t.w;
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float2 qq = make_float2(q[m], q[m]);
#pragma unroll
        for (int i = 0; i < 4; ++i) tc::ffma2(acc[m][i], qq, xp[i]);
      }
    }
  }
}

// fp16 hi + fp16 (residual * 2^11) split of two scaled fp32 values, packed as half2 words
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((v0 - hf.x) * 2048.f, (v1 - hf.y) * 2048.f);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

}  // namespace fgc
