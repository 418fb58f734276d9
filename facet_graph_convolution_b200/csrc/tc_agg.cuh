// Shared aggregation stage of the tensor-core kernels: soft assignments + q-weighted sum of gathered
// rows.  A group of kLPG lanes serves one facet (lane l of the group owns the float4 number l of the
// 64-channel row, plus l + kLPG, ... when kLPG < 16), a warp serves kFPW = 32 / kLPG facets and
// 32 / kFPW warps cover one 32-row pass.  Row loads are 16-byte, fully coalesced per group, issued
// kUnroll rows ahead of the packed fp32x2 FMAs to keep many L2 requests in flight.
#pragma once

#include "common.cuh"
#include "tc_common.cuh"

namespace fgc {

constexpr int MODE_FWD = 0;  // rows gathered through the adjacency (source-centric)
constexpr int MODE_TGT = 1;  // rows gathered through the reversed adjacency (target-centric)
constexpr int kQK = 16;      // neighbour slots per assignment round

constexpr int kLPG = 16;                 // lanes per facet
constexpr int kFPW = 32 / kLPG;          // facets per warp
constexpr int kF4 = 16 / kLPG;           // float4 per lane and row
constexpr int kCP = 2 * kF4;             // channel pairs per lane
constexpr int kAggW = 32 / kFPW;         // aggregator warps per 32-row pass
constexpr int kUnroll = (kLPG == 16) ? 8 : 4;
constexpr int kPairIters = (kFPW * kQK + 31) / 32;  // (facet, slot) pairs per lane and round

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
  int cw;                // aggregation channels actually present in a row (0 = all 64); the rest read as zero
  int upshift;           // FWD: x and uvx have (rows >> upshift) rows -- a fused custom_upsampling (repeat x 2^upshift)
};

template <int M>
struct AggQ {
  static constexpr int MQ = (M + 3) & ~3;            // q row stride in floats
  static constexpr int QS_FLOATS = kFPW * kQK * MQ;  // per warp
  static constexpr int NBR_INTS = kFPW * kQK;        // per warp
};

// word (two fp16) index inside a staged row of M*32 words for this lane's pair p of weight m
__device__ __forceinline__ int agg_word(int m, int gl, int p) { return m * 32 + 2 * (gl + kLPG * (p >> 1)) + (p & 1); }
// first channel of this lane's pair p
__device__ __forceinline__ int agg_channel(int gl, int p) { return 4 * (gl + kLPG * (p >> 1)) + 2 * (p & 1); }

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
#pragma unroll
  for (int o = kLPG; o < 32; o <<= 1) nround = max(nround, __shfl_xor_sync(0xffffffffu, nround, o));
  return nround;
}

// soft assignments of one round (list entries kb..kb+nk-1 of the warp's facets): lane per
// (facet, slot) pair, pair = lane + 32h, facet = pair / 16.  qs[f][k][0..M) receives q (already
// multiplied by inv_cnt[source] in TGT mode), nbr[f][k] the gathered row, -1 = padding,
// -2 = non-zero id outside the patch (counts as a neighbour, contributes nothing).
template <int M, int MODE>
__device__ __forceinline__ void agg_assign_round(const AggSrc& p, int64_t wrow0, int kb, int nk, int lst0,
                                                 int lst1, float* qs, int* nbr, int lane,
                                                 float (&dv)[kPairIters][M]) {
  constexpr int MQ = AggQ<M>::MQ;
#pragma unroll
  for (int h = 0; h < kPairIters; ++h) {
    const int f = (lane >> 4) + 2 * h, k = lane & 15;
    const int64_t rf = wrow0 + f;
    const int f0 = __shfl_sync(0xffffffffu, lst0, (f * kLPG) & 31);
    const int f1 = __shfl_sync(0xffffffffu, lst1, (f * kLPG) & 31);
    if (f >= kFPW) continue;  // (only when kFPW == 1)
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
        ux = p.uvx + (rf >> p.upshift) * (2 * M);
        vx = p.uvx + ((vvalid ? static_cast<int64_t>(row) : rf) >> p.upshift) * (2 * M) + M;
      } else {
        const int e = __ldg(p.rev_edge + f0 + kb + k);
        row = e / p.K;  // source facet of the in-edge
        ux = p.uvx + static_cast<int64_t>(row) * (2 * M);
        vx = p.uvx + rf * (2 * M) + M;
        const float* de = p.da_edge + static_cast<int64_t>(e) * M;
        if constexpr (M % 4 == 0) {
#pragma unroll
          for (int m = 0; m < M; m += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(de + m));
            dv[h][m] += t.x, dv[h][m + 1] += t.y, dv[h][m + 2] += t.z, dv[h][m + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int m = 0; m < M; ++m) dv[h][m] += __ldg(de + m);
        }
      }
      if constexpr (M % 4 == 0) {
#pragma unroll
        for (int m = 0; m < M; m += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(ux + m));
          float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
          if (vvalid) w = __ldg(reinterpret_cast<const float4*>(vx + m));
          a[m] = t.x + w.x, a[m + 1] = t.y + w.y, a[m + 2] = t.z + w.z, a[m + 3] = t.w + w.w;
        }
      } else {
#pragma unroll
        for (int m = 0; m < M; ++m) a[m] = __ldg(ux + m) + (vvalid ? __ldg(vx + m) : 0.f);
      }
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
    for (int m4 = 0; m4 < MQ; m4 += 4) {
      float4 t;
      t.x = (m4 < M) ? a[m4 < M ? m4 : 0] : 0.f;
      t.y = (m4 + 1 < M) ? a[m4 + 1 < M ? m4 + 1 : 0] : 0.f;
      t.z = (m4 + 2 < M) ? a[m4 + 2 < M ? m4 + 2 : 0] : 0.f;
      t.w = (m4 + 3 < M) ? a[m4 + 3 < M ? m4 + 3 : 0] : 0.f;
      *reinterpret_cast<float4*>(qd + m4) = t;
    }
    nbr[f * kQK + k] = row;
  }
}

// this lane's float4(s) of gathered row j as channel pairs (zeros for padding)
__device__ __forceinline__ void agg_load_row(const AggSrc& p, int j, int gl, float2 (&xp)[kCP]) {
#pragma unroll
  for (int i = 0; i < kF4; ++i) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j >= 0 && (p.cw == 0 || 4 * (gl + kLPG * i) < p.cw))
      t = __ldg(reinterpret_cast<const float4*>(p.x + static_cast<int64_t>(j >> p.upshift) * p.ldx) + gl + kLPG * i);
    xp[2 * i] = make_float2(t.x, t.y);
    xp[2 * i + 1] = make_float2(t.z, t.w);
  }
}

template <int M>
__device__ __forceinline__ void agg_load_q(const float* qk, float (&q)[AggQ<M>::MQ]) {
#pragma unroll
  for (int m4 = 0; m4 < AggQ<M>::MQ; m4 += 4) {
    const float4 t = *reinterpret_cast<const float4*>(qk + m4);
    q[m4] = t.x, q[m4 + 1] = t.y, q[m4 + 2] = t.z, q[m4 + 3] = t.w;
  }
}

// acc[m][i] += sum over the facet's neighbour list of q[.,m] * row[channel pair i]
// wrow0 = global row of the warp's first facet; lanes kLPG*g .. serve facet wrow0 + g.
// cnt   = number of non-padding list entries (FWD: non-zero adjacency ids).
// dv    = TGT: per-lane partial sums of da_edge (pair lane+32h belongs to facet (lane>>4)+2h).
template <int M, int MODE>
__device__ __forceinline__ void tc_aggregate(const AggSrc& p, int64_t wrow0, float* qs, int* nbr, int lane,
                                             float2 (&acc)[M][kCP], int& cnt, float (&dv)[kPairIters][M]) {
  constexpr int MQ = AggQ<M>::MQ;
  const int grp = lane / kLPG, gl = lane % kLPG;
  int lst0, lst1;
  const int nround = agg_list_bounds<MODE>(p, wrow0 + grp, lst0, lst1);
  for (int kb = 0; kb < nround; kb += kQK) {
    const int nk = min(kQK, nround - kb);
    __syncwarp();
    agg_assign_round<M, MODE>(p, wrow0, kb, nk, lst0, lst1, qs, nbr, lane, dv);
    __syncwarp();
    for (int k0 = 0; k0 < nk; k0 += kUnroll) {
      // issue the row loads of kUnroll neighbours before any of them is consumed
      float2 xp[kUnroll][kCP];
      int jj[kUnroll];
#pragma unroll
      for (int t = 0; t < kUnroll; ++t) {
        jj[t] = (k0 + t < nk) ? nbr[grp * kQK + k0 + t] : -1;
        agg_load_row(p, jj[t], gl, xp[t]);
      }
#pragma unroll
      for (int t = 0; t < kUnroll; ++t) {
        cnt += (jj[t] != -1);
        float q[MQ];
        agg_load_q<M>(qs + (grp * kQK + ((k0 + t < nk) ? k0 + t : 0)) * MQ, q);
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const float2 qq = make_float2(q[m], q[m]);
#pragma unroll
          for (int i = 0; i < kCP; ++i) tc::ffma2(acc[m][i], qq, xp[t][i]);
        }
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
