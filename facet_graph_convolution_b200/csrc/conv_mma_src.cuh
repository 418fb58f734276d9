// Source-centric backward pass on the staged tcgen05 pipeline (included by conv_mma.cu).
// Replaces the bwd_src pass of conv_bwd_tc.cu for the dense 64-channel, M = 8 layers:
//
//   ds'[n,m,:]  = W0[m]^T gy[n]                    stage A   tcgen05.mma  A = [Wt_h;Wt_l] (TMEM), B = gy image
//   dq[n,k,m]   = inv_cnt[n] ds'[n,m,:] . x_{j_k}  stage B   tcgen05.mma  A = X rows (smem), B = ds'
//                                                  (1/cnt is a per-facet scalar: applied by the pair threads)
//   da[n,k,m]   = q (dq - sum_m' q dq)             CUDA cores (softmax recomputed from staged logits)
//   outputs     da_edge[n,k,:], d_uvx[n,0:M] = sum_k da[n,k,:]
//
// Tile = 16 source facets, the forward tile plan gives its R distinct neighbour rows.  Per item
// (tile, chunk of <= 64 distinct rows) and per half h of the weight matrices (m = 4h .. 4h+3):
//   A_h   D_A[c (hi|lo lanes), (m', f hi | f lo)] = Wt[:, m-slice] . gy^T            16 MMAs, N = 32
//         (B operand = the tile's rows of the fp16 hi|lo image of gy, copied by the loaders with
//         cp.async straight into the K-major swizzled layout: no conversion, no fence in the pair warps)
//   dr_A  TMEM -> registers, hi/lo lanes of a channel combined by shuffle, fp16 hi/lo
//         -> shared DS_h[c][(m', f)]   (MN-major B operand of stage B, 128 B per channel row)
//   B_h   D_B[row (hi|lo lanes), (m', f)] = [Xh;Xl] . DS_h                           8 MMAs, N = 64
//   dr_B  TMEM -> registers, hi/lo lanes of a row combined by shuffle -> shared DQ_h[row][(m', f)] fp32
// then the pair threads pick dq[lidx] for their (facet, slot), form da and write the outputs.
// TMEM: weights 256 + D_A 128 + D_B 2 x 64 columns.
//
// Warp roles (20 warps): 0-3 drain A, 4-7 drain B (both aligned to the TMEM lane quadrants),
// 8-11 / 12-15 two pair groups (alternate items: softmax, da, outputs), 16-17 loaders,
// 18 stage-A issuer, 19 stage-B issuer.
#pragma once

namespace fgc {
namespace {

template <int M>
struct SrcCfg {
  static_assert(M == 8, "first instantiation: M = 8");
  static constexpr int TF = 128 / M;
  static constexpr int NX = 4;
  // ring slot: 128 operand rows of 128 B (lane-permuted hi/lo planes of the chunk's distinct rows) |
  // gy operand of the tile's own facets [32 rows: f hi | f lo][64 K = o] halves | neighbour logits |
  // header + pair records | own logits | 1/cnt of the tile's facets
  static constexpr int SL_X = 0;
  static constexpr int SL_GZ = 2 * kRC * 128;
  static constexpr int GZ_BYTES = 2 * TF * 128;
  static constexpr int SL_VL = SL_GZ + GZ_BYTES;
  static constexpr int SL_PR = SL_VL + kRC * M * 4;
  static constexpr int SL_UO = SL_PR + 16 + TF * 32 * 2;
  static constexpr int SL_INV = SL_UO + TF * M * 4;
  static constexpr int X_BUF = ((SL_INV + TF * 4) + 1023) / 1024 * 1024;
  static constexpr int DS_BYTES = 2 * kC * 128;          // hi | lo planes of [64 K rows = c][64 N = (m', f)]
  static constexpr int DQ_LD = 68;                       // fp32 row stride of DQ (padded: conflict-free drain stores)
  static constexpr int DQ_BYTES = kRC * DQ_LD * 4;       // [64 rows][64 (m', f)] fp32
  static constexpr int OFF_X = 0;
  static constexpr int OFF_DS = OFF_X + NX * X_BUF;      // ds: two halves
  static constexpr int OFF_DQ = OFF_DS + 2 * DS_BYTES;   // dq: [pair group][half]
  static constexpr int OFF_BAR = OFF_DQ + 4 * DQ_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;
  static constexpr int W_COL = 0, W_COLS = M * kC / 2;   // 256
  static constexpr int DA_COL = W_COL + W_COLS;          // 128 columns: 4 m' x (16 hi + 16 lo)
  static constexpr int DB_COL = DA_COL + 128;            // 2 x 64
  static constexpr int TMEM_COLS = 512;
  static_assert(DB_COL + 128 <= TMEM_COLS, "TMEM overflow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

struct SrcParams {
  CUtensorMap tmap;        // 2-D map of img for the TMA row gather (use_tma)
  int use_tma;
  const uint4* img;        // fp16 hi|lo image of x
  const float* xunscale;
  const uint4* gimg;       // fp16 hi|lo image of gy (scaled by 2^-eg)
  const float* gunscale;   // 2^eg
  const float* uvx;
  const uint8_t* ppair;
  const int32_t* prow;
  const int32_t* pR;
  const float* pinv;
  const uint4* wimg;       // transposed weight image (rows c hi|lo, K = (m, o))
  const float* wunscale;
  float* da_edge;          // [rows*K][M]
  float* d_uvx;            // [rows][2M], columns 0..M-1 written (single-chunk tiles)
  int64_t rows, ntiles;
  int N, K;
};

enum {
  S_X_FULL = 0,    // NX (32 cp.async arrivals)
  S_X_FREE = 4,    // NX (4 pair warps + stage-B commit)
  S_DA_FULL = 12,  // 1
  S_DA_FREE = 13,  // 4
  S_DS_FULL = 14,  // [half] (4)
  S_DS_FREE = 16,  // [half] (1, stage-B commit)
  S_DB_FULL = 18,  // [half] (1)
  S_DB_FREE = 20,  // [half] (4)
  S_DQ_FULL = 22,  // [pair group][half] (4)
  S_DQ_FREE = 26,  // [pair group][half] (4 pair warps)
  S_NUM = 30
  // Buffers written by one pair group and phases waited by one pair group are never shared between
  // the groups: an mbarrier waiter must not be a phase behind (a parity wait on a barrier that is
  // still in the previous phase returns at once).
};

template <int M, int KP, bool TMA>
__global__ void __launch_bounds__(kMmaThreads, 1)
bwd_src_mma_kernel(const __grid_constant__ SrcParams p) {
  using Cfg = SrcCfg<M>;
  constexpr int TF = Cfg::TF, NX = Cfg::NX;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S_NUM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = TF * p.K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NX; ++i) tc::mbar_init(&bars[S_X_FULL + i], 32), tc::mbar_init(&bars[S_X_FREE + i], 5);
    tc::mbar_init(&bars[S_DA_FULL], 1), tc::mbar_init(&bars[S_DA_FREE], 4);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars[S_DS_FULL + i], 4), tc::mbar_init(&bars[S_DS_FREE + i], 1);
      tc::mbar_init(&bars[S_DB_FULL + i], 1), tc::mbar_init(&bars[S_DB_FREE + i], 4);
    }
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[S_DQ_FULL + i], 4), tc::mbar_init(&bars[S_DQ_FREE + i], 4);
    tc::mbar_fence_init();
  }
  if (warp == 18) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < Cfg::OFF_BAR / 16; i += kMmaThreads) z[i] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp < 4) {
    // weight operand -> TMEM; lane 32q + 16h + i holds row h*64 + 16q + i of [Wt_h; Wt_l]
    const int row = ((lane >> 4) & 1) * kC + warp * 16 + (lane & 15);
#pragma unroll 1
    for (int mm = 0; mm < M; ++mm) {
      const uint4* src = p.wimg + (static_cast<size_t>(mm) * 2 * kC + row) * 8;
      uint32_t r[32];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 tq = __ldg(src + (u ^ (row & 7)));
        r[4 * u] = tq.x, r[4 * u + 1] = tq.y, r[4 * u + 2] = tq.z, r[4 * u + 3] = tq.w;
      }
      tc::tmem_st32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + Cfg::W_COL + mm * 32, r);
    }
    tc::tc_wait_st();
    tc::tc_fence_before_sync();
  }
  __syncthreads();
  tc::tc_fence_after_sync();

  // every role walks the same (tile, chunk) item sequence; R of the next tile is prefetched
  const int64_t step = gridDim.x;

  if (warp < 4) {
    // =========================================================== drain A: ds (TMEM) -> DS_h (smem)
    // lane 16h + i of quadrant q: h = 0 holds Wt_h rows of channel c = 16q + i, h = 1 the Wt_l rows.
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int hh = lane >> 4, c = warp * 16 + (lane & 15);
    int n = 0;     // (item, half) counter
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += step) {
      const int nch = chunks_of(Rn);
      if (tile + step < p.ntiles) Rn = __ldg(p.pR + tile + step);
      for (int ch = 0; ch < nch; ++ch) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h, ++n) {
          tc::mbar_wait(&bars[S_DA_FULL], n & 1);
          tc::tc_fence_after_sync();
          // this lane finalises facets 8*hh .. 8*hh+7 of channel c for the 4 weight matrices of the half
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int mq = 0; mq < 4; ++mq) {
            uint32_t d[32];
            tc::tmem_ld32(tmem + lane_base + Cfg::DA_COL + mq * 32, d);
            tc::tc_wait_ld();
            if (mq == 3) {
              tc::tc_fence_before_sync();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(&bars[S_DA_FREE]);
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float hi_lo = __uint_as_float(d[j]) + __uint_as_float(d[TF + j]);
              const float hi_up = __uint_as_float(d[8 + j]) + __uint_as_float(d[TF + 8 + j]);
              const float lo_lo = __uint_as_float(d[j]) * (1.f / 2048.f);
              const float lo_up = __uint_as_float(d[8 + j]) * (1.f / 2048.f);
              const float send = hh ? lo_lo : hi_up;
              const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
              v[j] = hh ? (recv + lo_up) : (hi_lo + recv);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a0 = v[2 * j], a1 = v[2 * j + 1];
              const float h0 = __uint_as_float(__float_as_uint(a0) & 0xFFFFE000u);
              const float h1 = __uint_as_float(__float_as_uint(a1) & 0xFFFFE000u);
              const __half2 hhv = __floats2half2_rn(h0, h1);
              const __half2 llv = __floats2half2_rn(a0 - h0, a1 - h1);
              hi[mq * 4 + j] = *reinterpret_cast<const uint32_t*>(&hhv);
              lo[mq * 4 + j] = *reinterpret_cast<const uint32_t*>(&llv);
            }
          }
          tc::mbar_wait(&bars[S_DS_FREE + h], ((n >> 1) & 1) ^ 1);
          // DS_h: MN-major B operand, K row = channel c (128 B: 64 N values (m', f)), 16-byte unit
          // u = 2 m' + hh holds facets 8hh..8hh+7 of weight m'
          uint8_t* ds = smem + Cfg::OFF_DS + h * Cfg::DS_BYTES + (c >> 3) * 1024 + (c & 7) * 128;
#pragma unroll
          for (int mq = 0; mq < 4; ++mq) {
            const int u = 2 * mq + hh;
            *reinterpret_cast<uint4*>(ds + ((u ^ (c & 7)) << 4)) =
                make_uint4(hi[4 * mq], hi[4 * mq + 1], hi[4 * mq + 2], hi[4 * mq + 3]);
            *reinterpret_cast<uint4*>(ds + kC * 128 + ((u ^ (c & 7)) << 4)) =
                make_uint4(lo[4 * mq], lo[4 * mq + 1], lo[4 * mq + 2], lo[4 * mq + 3]);
          }
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars[S_DS_FULL + h]);
        }
      }
    }
  } else if (warp < 8) {
    // =========================================================== drain B: dq (TMEM) -> DQ_h (smem, fp32)
    // operand row 32q + 16h + i of the slot holds the (h ? lo : hi) plane of distinct row 16q + i
    const int q = warp - 4;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int hh = lane >> 4, row = q * 16 + (lane & 15);
    int n = 0;
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += step) {
      const int nch = chunks_of(Rn);
      if (tile + step < p.ntiles) Rn = __ldg(p.pR + tile + step);
      for (int ch = 0; ch < nch; ++ch) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h, ++n) {
          const int nb = n >> 1;   // per-half use counter
          tc::mbar_wait(&bars[S_DB_FULL + h], nb & 1);
          tc::tc_fence_after_sync();
          uint32_t d0[32], d1[32];
          tc::tmem_ld32(tmem + lane_base + Cfg::DB_COL + h * 64, d0);
          tc::tmem_ld32(tmem + lane_base + Cfg::DB_COL + h * 64 + 32, d1);
          tc::tc_wait_ld();
          tc::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars[S_DB_FREE + h]);
          // hi-plane lane finalises columns 0..31, lo-plane lane columns 32..63
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float send = hh ? __uint_as_float(d0[j]) : __uint_as_float(d1[j]);
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            v[j] = (hh ? __uint_as_float(d1[j]) : __uint_as_float(d0[j])) + recv;
          }
          const int g = nb & 1, ng = nb >> 1;   // pair group of the item and its per-group use counter
          tc::mbar_wait(&bars[S_DQ_FREE + 2 * g + h], (ng & 1) ^ 1);
          float4* dst =
              reinterpret_cast<float4*>(smem + Cfg::OFF_DQ + (2 * g + h) * Cfg::DQ_BYTES + row * (Cfg::DQ_LD * 4) + hh * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars[S_DQ_FULL + 2 * g + h]);
        }
      }
    }
  } else if (warp < 16) {
    // =========================================================== pair groups
    const int grp = (warp - 8) >> 2;
    const int qt = threadIdx.x & 127;
    const int f = qt >> 3, s = qt & 7;
    const float unsc = __ldg(p.xunscale) * __ldg(p.wunscale) * __ldg(p.gunscale);
    int it = 0;
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += step) {
      const int64_t r = tile * TF + f;
      const bool rv = r < p.rows;
      const int nch = chunks_of(Rn);
      if (tile + step < p.ntiles) Rn = __ldg(p.pR + tile + step);
      for (int c = 0; c < nch; ++c, ++it) {
        const int xbuf = it % NX;
        const uint8_t* slot = smem + Cfg::OFF_X + xbuf * Cfg::X_BUF;
        if ((it & 1) != grp) continue;
        tc::mbar_wait(&bars[S_X_FULL + xbuf], (it / NX) & 1);
        const float dsc = unsc * *reinterpret_cast<const float*>(slot + Cfg::SL_INV + f * 4);
        // ---- soft assignments of this thread's pairs (every slot on its own: no multiplicity here)
        uint32_t rec[KP];
        float uo[M];
        {
          const uint16_t* pr = reinterpret_cast<const uint16_t*>(
              (c == 0) ? (slot + Cfg::SL_PR + 16) : (p.ppair + tile * (16 + P * 2) + 16));
#pragma unroll
          for (int j = 0; j < KP; ++j) {
            const int k = s + 8 * j;
            rec[j] = (rv && k < p.K) ? pr[f * p.K + k] : 0u;
          }
#pragma unroll
          for (int i = 0; i < M; i += 4) {
            float4 tq;
            if (c == 0) tq = *reinterpret_cast<const float4*>(slot + Cfg::SL_UO + f * (M * 4) + i * 4);
            else tq = rv ? __ldg(reinterpret_cast<const float4*>(p.uvx + r * (2 * M) + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            uo[i] = tq.x, uo[i + 1] = tq.y, uo[i + 2] = tq.z, uo[i + 3] = tq.w;
          }
        }
        float qv[KP][M];
        int col[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          const int lidx = rec[j] & 511;
          col[j] = ((rec[j] >> 15) && (lidx >> 6) == c) ? (lidx & 63) : -1;
          const float4* vp = reinterpret_cast<const float4*>(slot + Cfg::SL_VL + (lidx & 63) * (M * 4));
#pragma unroll
          for (int i = 0; i < M; i += 4) {
            const float4 tq = vp[i / 4];
            qv[j][i] = uo[i] + tq.x, qv[j][i + 1] = uo[i + 1] + tq.y, qv[j][i + 2] = uo[i + 2] + tq.z, qv[j][i + 3] = uo[i + 3] + tq.w;
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[S_X_FREE + xbuf]);   // staged data consumed
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          float mx = qv[j][0];
#pragma unroll
          for (int i = 1; i < M; ++i) mx = fmaxf(mx, qv[j][i]);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < M; ++i) {
            qv[j][i] = exp2f((qv[j][i] - mx) * 1.4426950408889634f);
            sum += qv[j][i];
          }
          const float rs = __fdividef(1.f, sum);
#pragma unroll
          for (int i = 0; i < M; ++i) qv[j][i] *= rs;
        }
        // ---- dq of both halves, da, outputs
        const int itg = it >> 1;   // this group's item counter
        float dux[M];
#pragma unroll
        for (int i = 0; i < M; ++i) dux[i] = 0.f;
        float dq[KP][M];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tc::mbar_wait(&bars[S_DQ_FULL + 2 * grp + h], itg & 1);
          const float* dqs = reinterpret_cast<const float*>(smem + Cfg::OFF_DQ + (2 * grp + h) * Cfg::DQ_BYTES);
#pragma unroll
          for (int j = 0; j < KP; ++j)
#pragma unroll
            for (int mq = 0; mq < 4; ++mq)
              dq[j][4 * h + mq] = (col[j] >= 0) ? dqs[col[j] * Cfg::DQ_LD + mq * TF + f] * dsc : 0.f;
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars[S_DQ_FREE + 2 * grp + h]);
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < M; ++i) dot = fmaf(qv[j][i], dq[j][i], dot);
          float da[M];
#pragma unroll
          for (int i = 0; i < M; ++i) {
            da[i] = qv[j][i] * (dq[j][i] - dot);
            if (col[j] >= 0) dux[i] += da[i];
          }
          if (col[j] >= 0) {
            float4* de = reinterpret_cast<float4*>(p.da_edge + (r * p.K + s + 8 * j) * M);
            de[0] = make_float4(da[0], da[1], da[2], da[3]);
            de[1] = make_float4(da[4], da[5], da[6], da[7]);
          }
        }
        if (nch == 1) {
          // own-logit gradient: sum over the facet's 8 slot lanes (fixed butterfly order)
#pragma unroll
          for (int i = 0; i < M; ++i) {
            dux[i] += __shfl_xor_sync(0xffffffffu, dux[i], 1);
            dux[i] += __shfl_xor_sync(0xffffffffu, dux[i], 2);
            dux[i] += __shfl_xor_sync(0xffffffffu, dux[i], 4);
          }
          if (s == 0 && rv) {
            float4* du = reinterpret_cast<float4*>(p.d_uvx + r * (2 * M));
            du[0] = make_float4(dux[0], dux[1], dux[2], dux[3]);
            du[1] = make_float4(dux[4], dux[5], dux[6], dux[7]);
          }
        }
      }
    }
  } else if (warp == 16 || warp == 17) {
    // =========================================================== loaders (cp.async, alternate items)
    const uint32_t xbase = tc::smem_u32(smem + Cfg::OFF_X);
    const int pr_bytes = 16 + P * 2;
    // operand row of distinct row l: 32 (l / 16) + (l % 16), lo plane 16 rows further
    auto op_row_off = [](int l) {
      const int rho = 32 * (l >> 4) + (l & 15);
      return (rho >> 3) * 1024 + (rho & 7) * 128;
    };
    auto load_ids = [&](int64_t tile, int c, int R, int (&rid)[2]) {
      const int rc = (tile < p.ntiles) ? min(kRC, R - c * kRC) : 0;
      const int32_t* rl = p.prow + tile * P + c * kRC;
      rid[0] = (lane < rc) ? __ldg(rl + lane) : -1;
      rid[1] = (lane + 32 < rc) ? __ldg(rl + lane + 32) : -1;
    };
    const int me = warp - 16;
    int it = 0;
    int64_t tile = blockIdx.x;
    int c = 0;
    int R = (tile < p.ntiles) ? __ldg(p.pR + tile) : 0;
    int Rn = (tile + step < p.ntiles) ? __ldg(p.pR + tile + step) : 0;
    auto advance = [&]() {
      ++c;
      if (c >= chunks_of(R)) {
        tile += step, c = 0, R = Rn;
        Rn = (tile + step < p.ntiles) ? __ldg(p.pR + tile + step) : 0;
      }
      ++it;
    };
    if (me == 1 && tile < p.ntiles) advance();
    int ids[2][2];
    if (tile < p.ntiles) load_ids(tile, c, R, ids[0]);
    auto issue = [&](const int (&cur)[2], int (&nxt)[2]) {
      const int64_t tile0 = tile;
      const int c0 = c, it0 = it, R0 = R;
      advance();
      if (tile < p.ntiles) advance();
      load_ids(tile, c, R, nxt);
      const int buf = it0 % NX;
      tc::mbar_wait(&bars[S_X_FREE + buf], ((it0 / NX) & 1) ^ 1);
      const uint32_t sl = xbase + buf * Cfg::X_BUF;
      if constexpr (TMA) {
        // row planes by TMA gather (see conv_mma.cu): lane j < 16 fetches rows 4j .. 4j+3 of the chunk
        const int rc0 = min(kRC, R0 - c0 * kRC);
        if (lane < 16 && 4 * lane < rc0) {
          int4 r4 = __ldg(reinterpret_cast<const int4*>(p.prow + tile0 * P + c0 * kRC) + lane);
          const int nv = rc0 - 4 * lane;
          if (nv < 2) r4.y = r4.x;
          if (nv < 3) r4.z = r4.x;
          if (nv < 4) r4.w = r4.x;
          const int l = 4 * lane;
          const uint32_t dh = sl + op_row_off(l);
          mbar_expect_tx(&bars[S_X_FULL + buf], 1024);
          tma_gather4(dh, &p.tmap, &bars[S_X_FULL + buf], 0, r4.x, r4.y, r4.z, r4.w);
          tma_gather4(dh + 2 * 1024, &p.tmap, &bars[S_X_FULL + buf], 64, r4.x, r4.y, r4.z, r4.w);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (cur[rr] >= 0) {
          const int l = lane + 32 * rr;
          const uint4* src = p.img + static_cast<int64_t>(cur[rr]) * 16;
          const uint32_t dh = sl + op_row_off(l), dl = dh + 2 * 1024;   // lo plane: operand row + 16
          const int sw = (32 * (l >> 4) + (l & 15)) & 7;                 // (row & 7) is the same for row + 16
          if constexpr (!TMA) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              cp_async16(dh + ((cc ^ sw) << 4), src + cc);
              cp_async16(dl + ((cc ^ sw) << 4), src + 8 + cc);
            }
          }
          const float* vsrc = p.uvx + static_cast<int64_t>(cur[rr]) * (2 * M) + M;
          const uint32_t vdst = sl + Cfg::SL_VL + l * (M * 4);
#pragma unroll
          for (int q = 0; q < M / 4; ++q) cp_async16(vdst + q * 16, vsrc + q * 4);
        }
      }
      {
        // per item: gy image rows (stage-A B operand, K-major rows n = f (hi) | TF + f (lo), 128B swizzle)
        // and 1/cnt of the tile's own facets; rows past the end are zero-filled
        const int64_t r0 = tile0 * TF;
#pragma unroll
        for (int q = lane; q < TF * 16; q += 32) {      // 16 x 16 B per image row
          const int ff = q >> 4, u16 = q & 15;
          const int nrow = (u16 >> 3) * TF + ff, cc = u16 & 7;
          const uint32_t dst = sl + Cfg::SL_GZ + (nrow >> 3) * 1024 + (nrow & 7) * 128 + ((cc ^ (nrow & 7)) << 4);
          const bool ok = r0 + ff < p.rows;
          cp_async16_zfill(dst, p.gimg + (ok ? (r0 + ff) * 16 + u16 : 0), ok);
        }
        if (lane < TF && r0 + lane < p.rows) cp_async4(sl + Cfg::SL_INV + lane * 4, p.pinv + r0 + lane);
      }
      if (c0 == 0) {
        const uint8_t* src = p.ppair + tile0 * pr_bytes;
        for (int q = lane * 16; q < pr_bytes; q += 512) cp_async16(sl + Cfg::SL_PR + q, src + q);
        const int64_t r = tile0 * TF + (lane >> 1);
        if ((lane >> 1) < TF && r < p.rows)
          cp_async16(sl + Cfg::SL_UO + (lane >> 1) * (M * 4) + (lane & 1) * 16, p.uvx + r * (2 * M) + (lane & 1) * 4);
      }
      cp_async_arrive_noinc(&bars[S_X_FULL + buf]);
    };
    while (tile < p.ntiles) {
      issue(ids[0], ids[1]);
      if (tile >= p.ntiles) break;
      issue(ids[1], ids[0]);
    }
  } else if (warp == 18) {
    // =========================================================== stage-A issuer: D_A = Wt[:, m-slice] . gz^T
    constexpr uint32_t idescA = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sb = tc::smem_u32(smem);
    int it = 0, n = 0;
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += step) {
      const int nch = chunks_of(Rn);
      if (tile + step < p.ntiles) Rn = __ldg(p.pR + tile + step);
      for (int c = 0; c < nch; ++c, ++it) {
        const int xbuf = it % NX;
        const uint32_t gzb = sb + Cfg::OFF_X + xbuf * Cfg::X_BUF + Cfg::SL_GZ;
        tc::mbar_wait(&bars[S_X_FULL + xbuf], (it / NX) & 1);
#pragma unroll 1
        for (int h = 0; h < 2; ++h, ++n) {
          tc::mbar_wait(&bars[S_DA_FREE], (n & 1) ^ 1);
          tc::tc_fence_after_sync();
          if (tc::elect_one()) {
#pragma unroll
            for (int mq = 0; mq < 4; ++mq) {
              const int mm = 4 * h + mq;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bd = tc::smem_desc_k_sw128(gzb + ks * 32);
                tc::mma_f16_ts(tmem + Cfg::DA_COL + mq * 32, tmem + Cfg::W_COL + mm * 32 + ks * 8, bd, idescA,
                               ks ? 1u : 0u);
              }
            }
            tc::tc_commit(&bars[S_DA_FULL]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================================================== stage-B issuer: D_B = [Xh;Xl] . DS_h
    // A = the slot's 128 operand rows (K-major, K = 64 channels), B = DS_h (MN-major, N = 64)
    constexpr uint32_t idescB = (1u << 4) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sb = tc::smem_u32(smem);
    int it = 0, n = 0;
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += step) {
      const int nch = chunks_of(Rn);
      if (tile + step < p.ntiles) Rn = __ldg(p.pR + tile + step);
      for (int c = 0; c < nch; ++c, ++it) {
        const int xbuf = it % NX;
        tc::mbar_wait(&bars[S_X_FULL + xbuf], (it / NX) & 1);
#pragma unroll 1
        for (int h = 0; h < 2; ++h, ++n) {
          const int nb = n >> 1;
          tc::mbar_wait(&bars[S_DS_FULL + h], nb & 1);
          tc::mbar_wait(&bars[S_DB_FREE + h], (nb & 1) ^ 1);
          tc::tc_fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t xa = sb + Cfg::OFF_X + xbuf * Cfg::X_BUF;
            const uint32_t dsh = sb + Cfg::OFF_DS + h * Cfg::DS_BYTES, dsl = dsh + kC * 128;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = tc::smem_desc_k_sw128(xa + ks * 32);
              const uint64_t bh = desc_mn_sw128(dsh + ks * 2048, 1024, 1024);
              const uint64_t bl = desc_mn_sw128(dsl + ks * 2048, 1024, 1024);
              tc::mma_f16_ss(tmem + Cfg::DB_COL + h * 64, ad, bh, idescB, ks ? 1u : 0u);
              tc::mma_f16_ss(tmem + Cfg::DB_COL + h * 64, ad, bl, idescB, 1u);
            }
            tc::tc_commit(&bars[S_DS_FREE + h]);
            tc::tc_commit(&bars[S_DB_FULL + h]);
            if (h == 1) tc::tc_commit(&bars[S_X_FREE + xbuf]);
          }
          __syncwarp();
        }
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 18) tc::tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// own-logit gradient of facets in multi-chunk tiles (> 64 distinct rows): d_uvx[n, 0:M] = sum_k da_edge[n,k,:]
// over the valid slots, in slot order
template <int M>
__global__ void src_dux_fix_kernel(const int32_t* __restrict__ pR, const int32_t* __restrict__ adj,
                                   const float* __restrict__ da_edge, float* __restrict__ d_uvx, int64_t rows,
                                   int N, int K, int TF) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  if (__ldg(pR + r / TF) <= kRC) return;
  float a[M];
#pragma unroll
  for (int i = 0; i < M; ++i) a[i] = 0.f;
  for (int k = 0; k < K; ++k) {
    const int id = __ldg(adj + r * K + k);
    if (id > 0 && id <= N) {
#pragma unroll
      for (int i = 0; i < M; ++i) a[i] += __ldg(da_edge + (r * K + k) * M + i);
    }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) d_uvx[r * (2 * M) + i] = a[i];
}

}  // namespace
}  // namespace fgc
