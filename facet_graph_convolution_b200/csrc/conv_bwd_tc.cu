// Tensor-core weight-gradient pass of the facet-graph convolution backward (dense 64-channel
// layers):  gW0[m][o][c] = sum_n gz[n][o] * s[n][m][c],   gb[o] = sum_n flag[n] * gy[n][o]
// with s recomputed by the shared aggregation stage (tc_agg.cuh).
//
// Persistent CTA, 13 warps, passes of 32 facets, two shared-memory stages:
//   warps 0-7  aggregate s for 4 facets each, scale by a global power of two, split into fp16
//              hi / lo*2^11 planes and write them as the UMMA A operand  A = s^T  (M = (m,c), K = facet)
//              in the canonical MN-major 128B-swizzled layout; gz goes to the B operand (N = o).
//   warp 12    issues tcgen05.mma kind::f16 (SS mode, both operands MN-major):
//                 D_hh[(m,c)][o] += s_hi^T gz_hi ;  D_x += s_hi^T gz_lo + s_lo^T gz_hi
//              the 4 x (64 + 64) accumulator columns fill the whole TMEM and stay resident for
//              the lifetime of the CTA -- no shared-memory accumulators, no atomics.
//   warps 8-11 read the accumulators once at the end and write this CTA's partial; a fixed-order
//              reduction over CTAs (reduce_partials_kernel) finishes gW0 / gb deterministically.
#include "conv_common.cuh"
#include "conv_launch.cuh"
#include "tc_agg.cuh"
#include "tc_common.cuh"

namespace fgc {

namespace {

constexpr int kAggWarpsW = kAggW;
constexpr int kEpiWarpsW = 4;
constexpr int kThreadsW = (kAggWarpsW + kEpiWarpsW + 1) * 32;
constexpr int kPassW = 32;

// |x| max as ordered uint bits (non-negative floats order like unsigned ints)
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// power-of-two scale s with |v| * s < 2^15 for every |v| <= bound * 2^extra_log2 ; returns 1/s too
__device__ __forceinline__ void pow2_scale(unsigned maxbits, int extra_log2, float& scale, float& unscale) {
  int E = static_cast<int>((maxbits >> 23) & 0xFF) + extra_log2;
  E = min(max(E, 20), 250);
  scale = __int_as_float((268 - E) << 23);    // 2^(141-E)
  unscale = __int_as_float((E - 14) << 23);   // 2^(E-141)
}

template <int M>
struct WCfg {
  static constexpr int MQ = AggQ<M>::MQ;
  static constexpr int A_PLANE = 4 * M * 1024;                 // [4 k-groups][M blocks][8 rows][128 B]
  static constexpr int B_PLANE = 4 * 1024;                     // [4 k-groups][8 rows][128 B]
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int OFF_Q = 2 * STAGE;
  static constexpr int OFF_NBR = OFF_Q + kAggWarpsW * AggQ<M>::QS_FLOATS * 4;
  static constexpr int OFF_BAR = OFF_NBR + kAggWarpsW * AggQ<M>::NBR_INTS * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  static constexpr int NBLK = M / 2;                           // M-blocks of 128 (m,c) rows
  static_assert(M % 2 == 0, "M must be even");
  static_assert(NBLK * 128 <= 512, "TMEM overflow");
  static_assert(STAGE % 1024 == 0, "stage must keep 1024-byte alignment");
};

struct WParams {
  AggSrc src;
  const float* gy;
  const unsigned* maxbits;   // [0] = max|x| bits, [1] = max|gy| bits
  float* partW;              // [grid][M][64][64]
  float* partB;              // [grid][64]
  int bias_mask;
  int64_t npasses;
};

enum { WB_FULL0 = 0, WB_FULL1, WB_EMPTY0, WB_EMPTY1, WB_DONE, WB_COUNT };

__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int M>
__global__ void __launch_bounds__(kThreadsW, 1)
bwd_w_tc_kernel(const WParams p) {
  using Cfg = WCfg<M>;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* qs_all = reinterpret_cast<float*>(smem + Cfg::OFF_Q);
  int* nbr_all = reinterpret_cast<int*>(smem + Cfg::OFF_NBR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WB_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[WB_FULL0], kAggWarpsW);
    tc::mbar_init(&bars[WB_FULL1], kAggWarpsW);
    tc::mbar_init(&bars[WB_EMPTY0], 1);
    tc::mbar_init(&bars[WB_EMPTY1], 1);
    tc::mbar_init(&bars[WB_DONE], 1);
    tc::mbar_fence_init();
  }
  if (warp == kAggWarpsW + kEpiWarpsW) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  float ss, s_un, sg, g_un;
  pow2_scale(__ldg(p.maxbits), 5, ss, s_un);      // |s| <= K * max|x|, K <= 32
  pow2_scale(__ldg(p.maxbits + 1), 0, sg, g_un);  // |gz| <= max|gy|

  if (warp < kAggWarpsW) {
    // =========================================================== aggregators
    float* qs = qs_all + warp * AggQ<M>::QS_FLOATS;
    int* nbr = nbr_all + warp * AggQ<M>::NBR_INTS;
    const int grp = lane / kLPG, gl = lane % kLPG;
    const int j = warp * kFPW + grp;         // row within the pass
    const int kg = j >> 3, jr = j & 7;
    float gb[2 * kCP];
#pragma unroll
    for (int i = 0; i < 2 * kCP; ++i) gb[i] = 0.f;
    int it = 0;
    for (int64_t pass = blockIdx.x; pass < p.npasses; pass += gridDim.x, ++it) {
      const int64_t r = pass * kPassW + j;
      float2 acc[M][kCP];
#pragma unroll
      for (int m = 0; m < M; ++m)
#pragma unroll
        for (int i = 0; i < kCP; ++i) acc[m][i] = make_float2(0.f, 0.f);
      int cnt = 0;
      float dv[kPairIters][M];
      tc_aggregate<M, MODE_FWD>(p.src, pass * kPassW + warp * kFPW, qs, nbr, lane, acc, cnt, dv);
      // gz row of this facet: the lane's float4(s) of gy
      float4 g[kF4];
#pragma unroll
      for (int i = 0; i < kF4; ++i) {
        g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < p.src.rows) g[i] = __ldg(reinterpret_cast<const float4*>(p.gy + r * 64) + gl + kLPG * i);
      }
      const float fl = (r < p.src.rows && (cnt > 0 || !p.bias_mask)) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < kF4; ++i) {
        gb[4 * i] = fmaf(fl, g[i].x, gb[4 * i]), gb[4 * i + 1] = fmaf(fl, g[i].y, gb[4 * i + 1]);
        gb[4 * i + 2] = fmaf(fl, g[i].z, gb[4 * i + 2]), gb[4 * i + 3] = fmaf(fl, g[i].w, gb[4 * i + 3]);
      }
      const float gsc = (cnt ? 1.f / static_cast<float>(cnt) : 0.f) * sg;

      const int st = it & 1;
      tc::mbar_wait(&bars[WB_EMPTY0 + st], ((it >> 1) & 1) ^ 1);
      uint8_t* base = smem + st * Cfg::STAGE;
      uint8_t* ah = base;
      uint8_t* al = base + Cfg::A_PLANE;
      uint8_t* bh = base + 2 * Cfg::A_PLANE;
      uint8_t* bl = bh + Cfg::B_PLANE;
      // element (row j, block m, channel c): ((kg*M + m)*8 + jr)*128 + ((c/8 ^ jr)*16) + (c%8)*2
      int uo[kF4];
#pragma unroll
      for (int i = 0; i < kF4; ++i) {
        const int c = 4 * (gl + kLPG * i);
        uo[i] = ((c >> 3) ^ jr) * 16 + (c & 7) * 2;
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        uint32_t h[kCP], l[kCP];
#pragma unroll
        for (int i = 0; i < kCP; ++i) split_pair(acc[m][i].x * ss, acc[m][i].y * ss, h[i], l[i]);
        const int rowoff = ((kg * M + m) * 8 + jr) * 128;
#pragma unroll
        for (int i = 0; i < kF4; ++i) {
          *reinterpret_cast<uint2*>(ah + rowoff + uo[i]) = make_uint2(h[2 * i], h[2 * i + 1]);
          *reinterpret_cast<uint2*>(al + rowoff + uo[i]) = make_uint2(l[2 * i], l[2 * i + 1]);
        }
      }
      {
        const int rowoff = (kg * 8 + jr) * 128;
#pragma unroll
        for (int i = 0; i < kF4; ++i) {
          uint32_t h0, l0, h1, l1;
          split_pair(g[i].x * gsc, g[i].y * gsc, h0, l0);
          split_pair(g[i].z * gsc, g[i].w * gsc, h1, l1);
          *reinterpret_cast<uint2*>(bh + rowoff + uo[i]) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(bl + rowoff + uo[i]) = make_uint2(l0, l1);
        }
      }
      tc::fence_proxy_async_smem();   // generic-proxy writes -> UMMA (async proxy) reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[WB_FULL0 + st]);
    }
    // ---- bias gradient: fixed-order reduction over the 32 facet slots of the CTA
    asm volatile("bar.sync 2, %0;" ::"n"(kAggWarpsW * 32) : "memory");
    float* red = qs_all;  // [32 slots][64 channels]
#pragma unroll
    for (int i = 0; i < kF4; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) red[j * 64 + 4 * (gl + kLPG * i) + q] = gb[4 * i + q];
    asm volatile("bar.sync 2, %0;" ::"n"(kAggWarpsW * 32) : "memory");
    if (threadIdx.x < 64) {
      float a = 0.f;
      for (int s = 0; s < 32; ++s) a += red[s * 64 + threadIdx.x];
      p.partB[static_cast<int64_t>(blockIdx.x) * 64 + threadIdx.x] = a;
    }
  } else if (warp < kAggWarpsW + kEpiWarpsW) {
    // =========================================================== final read-out
    const int quad = warp - kAggWarpsW;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    tc::mbar_wait_relaxed(&bars[WB_DONE], 0);
    tc::tc_fence_after_sync();
    const float un = s_un * g_un;
    float* pw = p.partW + static_cast<int64_t>(blockIdx.x) * M * 64 * 64;
    for (int mb = 0; mb < Cfg::NBLK; ++mb) {
      const int idx = mb * 128 + quad * 32 + lane;   // (m,c) row of this thread
      const int m = idx >> 6, c = idx & 63;
#pragma unroll 1
      for (int o0 = 0; o0 < 64; o0 += 32) {
        uint32_t dh[32], dx[32];
        tc::tmem_ld32(tmem + lane_base + mb * 128 + o0, dh);
        tc::tmem_ld32(tmem + lane_base + mb * 128 + 64 + o0, dx);
        tc::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = (__uint_as_float(dh[i]) + __uint_as_float(dx[i]) * (1.f / 2048.f)) * un;
          pw[(static_cast<int64_t>(m) * 64 + o0 + i) * 64 + c] = v;
        }
      }
    }
    tc::tc_fence_before_sync();
  } else {
    // =========================================================== MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sbase = tc::smem_u32(smem);
    int it = 0;
    for (int64_t pass = blockIdx.x; pass < p.npasses; pass += gridDim.x, ++it) {
      const int st = it & 1;
      tc::mbar_wait_relaxed(&bars[WB_FULL0 + st], (it >> 1) & 1);
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t ah = sbase + st * Cfg::STAGE, al = ah + Cfg::A_PLANE;
        const uint32_t bh = ah + 2 * Cfg::A_PLANE, bl = bh + Cfg::B_PLANE;
#pragma unroll 1
        for (int mb = 0; mb < Cfg::NBLK; ++mb) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t aoff = (ks * 2 * M + mb * 2) * 1024;   // k-group 2ks, MN-atom 2mb
            const uint64_t dah = smem_desc_mn_sw128(ah + aoff, 1024, M * 1024);
            const uint64_t dal = smem_desc_mn_sw128(al + aoff, 1024, M * 1024);
            const uint64_t dbh = smem_desc_mn_sw128(bh + ks * 2 * 1024, 1024, 1024);
            const uint64_t dbl = smem_desc_mn_sw128(bl + ks * 2 * 1024, 1024, 1024);
            const uint32_t first = (it | ks) ? 1u : 0u;
            tc::mma_f16_ss(tmem + mb * 128, dah, dbh, idesc, first);
            tc::mma_f16_ss(tmem + mb * 128 + 64, dah, dbl, idesc, first);
            tc::mma_f16_ss(tmem + mb * 128 + 64, dal, dbh, idesc, 1u);
          }
        }
        tc::tc_commit(&bars[WB_EMPTY0 + st]);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::tc_commit(&bars[WB_DONE]);
    __syncwarp();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == kAggWarpsW + kEpiWarpsW) tc::tmem_dealloc(tmem, 512);
}


// ====================================================================================================
// Source-centric backward pass on tensor cores:
//   ds[n,m,:]  = W0[m]^T gz[n]            tcgen05.mma, A' = [gz_hi ; gz_lo] in TMEM (32-row tiles),
//                                         B = the transposed weight image shared with bwd_tgt
//   dq[n,k,m]  = ds[n,m,:] . x_{j_k}      8 lanes per facet, packed FMAs + transpose-reduce shuffles
//   da         = q (dq - sum_m q dq)  ->  da_edge[n,k,:],  d_uvx[n,0:M] = sum_k da,  inv_cnt[n]
// Warps 0-7 aggregate, warp 8/9 own TMEM lane quadrants 0/1 (hi / lo rows: A' writer and D' reader),
// warp 10 issues the MMAs.  ds travels TMEM -> shared (fp32, 32 x 512) -> registers of the owning lanes.
constexpr int kSrcAggWarps = kAggW;
constexpr int kSrcThreads = (kSrcAggWarps + 3) * 32;
constexpr int kSrcTile = 32;

template <int M>
struct SCfg {
  static constexpr int KK = M * 64;
  static constexpr int NB = 128;
  static constexpr int W_BYTES = M * NB * 128;
  static constexpr int DS_PITCH = KK + 4;                  // floats, == 4 mod 32
  static constexpr int EX_PITCH = 33;
  static constexpr int OFF_W = 0;
  static constexpr int OFF_DS = OFF_W + W_BYTES;
  static constexpr int OFF_EX = OFF_DS + kSrcTile * DS_PITCH * 4;
  static constexpr int OFF_Q = OFF_EX + kSrcTile * EX_PITCH * 4;
  static constexpr int OFF_NBR = OFF_Q + kSrcAggWarps * AggQ<M>::QS_FLOATS * 4;
  static constexpr int OFF_BAR = OFF_NBR + kSrcAggWarps * AggQ<M>::NBR_INTS * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
  static constexpr int D_COL0 = 64;                        // two D' slots of 128 columns
  static_assert(DS_PITCH % 32 == 4, "ds pitch must be 4 mod 32 words");
};

struct SParams {
  AggSrc src;
  const float* gy;
  const uint4* wimg;       // transposed weight image (chunk m: rows c hi|lo, K = o)
  const float* wunscale;
  float* da_edge;
  float* d_uvx;
  float* inv_out;
  int64_t ntiles;
};

enum { SB_A_READY = 0, SB_A_FREE, SB_D_FULL0, SB_D_FULL1, SB_D_FREE0, SB_D_FREE1, SB_DS_FULL, SB_DS_FREE, SB_COUNT };

template <int M>
__global__ void __launch_bounds__(kSrcThreads, 1)
bwd_src_tc_kernel(const SParams p) {
  static_assert(M == 8, "the transpose-reduce below is written for M == 8");
  using Cfg = SCfg<M>;
  constexpr int MQ = AggQ<M>::MQ;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* DS = reinterpret_cast<float*>(smem + Cfg::OFF_DS);
  float* EX = reinterpret_cast<float*>(smem + Cfg::OFF_EX);
  float* qs_all = reinterpret_cast<float*>(smem + Cfg::OFF_Q);
  int* nbr_all = reinterpret_cast<int*>(smem + Cfg::OFF_NBR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + SB_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[SB_A_READY], 2);
    tc::mbar_init(&bars[SB_A_FREE], 1);
    tc::mbar_init(&bars[SB_D_FULL0], 1);
    tc::mbar_init(&bars[SB_D_FULL1], 1);
    tc::mbar_init(&bars[SB_D_FREE0], 2);
    tc::mbar_init(&bars[SB_D_FREE1], 2);
    tc::mbar_init(&bars[SB_DS_FULL], 1);
    tc::mbar_init(&bars[SB_DS_FREE], kSrcAggWarps);
    tc::mbar_fence_init();
  }
  if (warp == kSrcAggWarps + 2) tc::tmem_alloc(tmem_slot, 512);
  {
    uint4* wdst = reinterpret_cast<uint4*>(smem + Cfg::OFF_W);
    for (int i = threadIdx.x; i < Cfg::W_BYTES / 16; i += kSrcThreads) wdst[i] = __ldg(p.wimg + i);
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < kSrcAggWarps) {
    // =========================================================== aggregators
    float* qs = qs_all + warp * AggQ<M>::QS_FLOATS;
    int* nbr = nbr_all + warp * AggQ<M>::NBR_INTS;
    const int grp = lane / kLPG, gl = lane % kLPG;
    const int j = warp * kFPW + grp;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int64_t wrow0 = tile * kSrcTile + warp * kFPW;
      const int64_t r = wrow0 + grp;
      tc::mbar_wait(&bars[SB_DS_FULL], it & 1);
      float2 ds2[M][kCP];
#pragma unroll
      for (int m = 0; m < M; ++m)
#pragma unroll
        for (int i = 0; i < kF4; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(DS + j * Cfg::DS_PITCH + m * 64 + 4 * (gl + kLPG * i));
          ds2[m][2 * i] = make_float2(a.x, a.y), ds2[m][2 * i + 1] = make_float2(a.z, a.w);
        }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[SB_DS_FREE]);
      int lst0, lst1;
      const int nround = agg_list_bounds<MODE_FWD>(p.src, r, lst0, lst1);
      float dv[kPairIters][M];
      float dux = 0.f;
      for (int kb = 0; kb < nround; kb += kQK) {
        const int nk = min(kQK, nround - kb);
        __syncwarp();
        agg_assign_round<M, MODE_FWD>(p.src, wrow0, kb, nk, lst0, lst1, qs, nbr, lane, dv);
        __syncwarp();
        // loop-invariant addressing of this lane's output slot and assignment column
        const bool wr_lane = ((kLPG == 16) ? ((gl & 1) == 0) : true) && r < p.src.rows;
        const int mine_c = (kLPG == 16) ? (gl >> 1) : gl;
        float* de_base = p.da_edge + (r * p.src.K + kb) * M + mine_c;
        const float* q_base = qs + grp * kQK * MQ + mine_c;
        for (int k0 = 0; k0 < nk; k0 += kUnroll) {
          float2 xp[kUnroll][kCP];
#pragma unroll
          for (int t = 0; t < kUnroll; ++t)
            agg_load_row(p.src, (k0 + t < nk) ? nbr[grp * kQK + k0 + t] : -1, gl, xp[t]);
#pragma unroll
          for (int t = 0; t < kUnroll; ++t) {
            const int k = k0 + t;
            float v[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
              float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < kCP; ++i) tc::ffma2(s2, ds2[m][i], xp[t][i]);
              v[m] = s2.x + s2.y;
            }
            // transpose-reduce over the lanes of the facet: the 8 partial sums end up one per lane
            // (lane gl holds m = gl for 8-lane groups, m = gl >> 1 for 16-lane groups)
            int mine;
            float dq;
            if constexpr (kLPG == 16) {
              {
                const bool up = (gl & 8) != 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float keep = up ? v[i + 4] : v[i], send = up ? v[i] : v[i + 4];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
              }
              {
                const bool up = (gl & 4) != 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const float keep = up ? v[i + 2] : v[i], send = up ? v[i] : v[i + 2];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
              }
              {
                const bool up = (gl & 2) != 0;
                const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
                dq = keep + __shfl_xor_sync(0xffffffffu, send, 2);
              }
              dq += __shfl_xor_sync(0xffffffffu, dq, 1);
              mine = gl >> 1;
            } else {
              {
                const bool up = (gl & 4) != 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float keep = up ? v[i + 4] : v[i], send = up ? v[i] : v[i + 4];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
              }
              {
                const bool up = (gl & 2) != 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const float keep = up ? v[i + 2] : v[i], send = up ? v[i] : v[i + 2];
                  v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
              }
              {
                const bool up = (gl & 1) != 0;
                const float keep = up ? v[1] : v[0], send = up ? v[0] : v[1];
                dq = keep + __shfl_xor_sync(0xffffffffu, send, 1);
              }
              mine = gl;
            }
            (void)mine;
            const float qv = q_base[(k < nk ? k : 0) * MQ];
            float dot = qv * dq;
            if constexpr (kLPG == 16) {
              dot += __shfl_xor_sync(0xffffffffu, dot, 2);
              dot += __shfl_xor_sync(0xffffffffu, dot, 4);
              dot += __shfl_xor_sync(0xffffffffu, dot, 8);
            } else {
              dot += __shfl_xor_sync(0xffffffffu, dot, 1);
              dot += __shfl_xor_sync(0xffffffffu, dot, 2);
              dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            }
            const float da = qv * (dq - dot);
            if (k < nk) {          // kb + k < K always holds here (nk <= K - kb)
              if (wr_lane) de_base[k * M] = da;
              dux += da;
            }
          }
        }
      }
      if (r < p.src.rows && ((kLPG == 16) ? ((gl & 1) == 0) : true)) p.d_uvx[r * (2 * M) + ((kLPG == 16) ? (gl >> 1) : gl)] = dux;
    }
  } else if (warp < kSrcAggWarps + 2) {
    // =========================================================== TMEM lane owners (quad 0: hi, quad 1: lo)
    const int quad = warp - kSrcAggWarps;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const float wun = __ldg(p.wunscale);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int64_t r = tile * kSrcTile + lane;
      // ---- gz row -> A' (hi or lo plane), scaled by a power of two to [0.5, 1)
      float inv = 0.f, mx = 0.f;
      const float4* gr = reinterpret_cast<const float4*>(p.gy + r * 64);
      if (r < p.src.rows) {
        int cnt = 0;
        for (int k = 0; k < p.src.K; ++k) cnt += (__ldg(p.src.adj + r * p.src.K + k) != 0);
        inv = cnt ? 1.f / static_cast<float>(cnt) : 0.f;
        if (quad == 0) p.inv_out[r] = inv;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 t = __ldg(gr + i);
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
        }
        mx *= inv;
      }
      int E = (__float_as_int(mx) >> 23) & 0xFF;
      E = min(max(E, 16), 240);
      const float sc = __int_as_float((253 - E) << 23) * inv;
      const float unsc = __int_as_float((E + 1) << 23) * wun;
      tc::mbar_wait(&bars[SB_A_FREE], (it & 1) ^ 1);
      tc::tc_fence_after_sync();
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < p.src.rows) t = __ldg(gr + half * 8 + i);
          uint32_t h0, l0, h1, l1;
          split_pair(t.x * sc, t.y * sc, h0, l0);
          split_pair(t.z * sc, t.w * sc, h1, l1);
          w[2 * i] = quad == 0 ? h0 : l0;
          w[2 * i + 1] = quad == 0 ? h1 : l1;
        }
        tc::tmem_st16(tmem + lane_base + half * 16, w);
      }
      tc::tc_wait_st();
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[SB_A_READY]);
      // ---- D' chunks -> ds (fp32) in shared memory
      tc::mbar_wait_relaxed(&bars[SB_DS_FREE], (it & 1) ^ 1);
      for (int jm = 0; jm < M; ++jm) {
        const int cj = it * M + jm, slot = cj & 1;
        tc::mbar_wait(&bars[SB_D_FULL0 + slot], (cj >> 1) & 1);
        tc::tc_fence_after_sync();
        const uint32_t dcol = tmem + lane_base + Cfg::D_COL0 + slot * 128;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t d0[32], d1[32];
          tc::tmem_ld32(dcol + half * 32, d0);        // (.)Wh
          tc::tmem_ld32(dcol + 64 + half * 32, d1);   // (.)Wl
          tc::tc_wait_ld();
          if (quad == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              EX[lane * Cfg::EX_PITCH + i] =
                  __uint_as_float(d0[i]) * (1.f / 2048.f) + __uint_as_float(d1[i]) * (1.f / 4194304.f);
          }
          asm volatile("bar.sync 3, 64;" ::: "memory");
          if (quad == 0) {
            float* dst = DS + lane * Cfg::DS_PITCH + jm * 64 + half * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float o[4];
#pragma unroll
              for (int q = 0; q < 4; ++q)
                o[q] = (__uint_as_float(d0[i + q]) + __uint_as_float(d1[i + q]) * (1.f / 2048.f) +
                        EX[lane * Cfg::EX_PITCH + i + q]) * unsc;
              *reinterpret_cast<float4*>(dst + i) = make_float4(o[0], o[1], o[2], o[3]);
            }
          }
          asm volatile("bar.sync 3, 64;" ::: "memory");
        }
        tc::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[SB_D_FREE0 + slot]);
      }
      if (quad == 0) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[SB_DS_FULL]);
      }
    }
  } else {
    // =========================================================== MMA issuer
    const uint32_t idesc = tc::idesc_f16(128, Cfg::NB);
    const uint32_t wbase = tc::smem_u32(smem + Cfg::OFF_W);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      tc::mbar_wait_relaxed(&bars[SB_A_READY], it & 1);
      for (int jm = 0; jm < M; ++jm) {
        const int cj = it * M + jm, slot = cj & 1;
        tc::mbar_wait(&bars[SB_D_FREE0 + slot], ((cj >> 1) & 1) ^ 1);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bdesc = tc::smem_desc_k_sw128(wbase + jm * (Cfg::NB * 128) + ks * 32);
            tc::mma_f16_ts(tmem + Cfg::D_COL0 + slot * 128, tmem + ks * 8, bdesc, idesc, ks ? 1u : 0u);
          }
          tc::tc_commit(&bars[SB_D_FULL0 + slot]);
          if (jm == M - 1) tc::tc_commit(&bars[SB_A_FREE]);
        }
        __syncwarp();
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == kSrcAggWarps + 2) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool bwd_w_tc_supported(int Cw, int Cout, int M, int Cin) { return Cw == 64 && Cout == 64 && M == 8 && Cin % 4 == 0; }

int bwd_w_tc_grid(int64_t rows) {
  const int64_t npasses = (rows + kPassW - 1) / kPassW;
  int64_t g = num_sms();
  if (g > npasses) g = npasses;
  return static_cast<int>(g < 1 ? 1 : g);
}

// partW[grid][M*64*64], partB[grid][64]; maxbits: 2 uints of scratch
// xmax / gmax: device words holding the float bits of max|x| / max|gy| when a previous pass already
// reduced them (nullptr: reduce here)
int launch_bwd_w_tc(const float* gy, const float* x, const int32_t* adj, const float* uvx, float* partW,
                    float* partB, unsigned* maxbits, const unsigned* xmax, const unsigned* gmax, int64_t rows, int N,
                    int K, int Cin, int M, int bias_mask, cudaStream_t st) {
  const int ab = num_sms() * 4;
  if (xmax != nullptr) {
    FGC_CUDA(cudaMemcpyAsync(maxbits, xmax, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
  } else {
    FGC_CUDA(cudaMemsetAsync(maxbits, 0, sizeof(unsigned), st));
    absmax_kernel<<<ab, 256, 0, st>>>(x, rows * Cin, maxbits);
    FGC_LAUNCHED("absmax_kernel");
  }
  if (gmax != nullptr) {
    FGC_CUDA(cudaMemcpyAsync(maxbits + 1, gmax, sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
  } else {
    FGC_CUDA(cudaMemsetAsync(maxbits + 1, 0, sizeof(unsigned), st));
    absmax_kernel<<<ab, 256, 0, st>>>(gy, rows * 64, maxbits + 1);
    FGC_LAUNCHED("absmax_kernel");
  }
  WParams p{};
  p.src = AggSrc{x, Cin, adj, uvx, N, K, rows, nullptr, nullptr, nullptr, nullptr};
  p.gy = gy, p.maxbits = maxbits, p.partW = partW, p.partB = partB, p.bias_mask = bias_mask;
  p.npasses = (rows + kPassW - 1) / kPassW;
  if (M != 8) {
    set_error("bwd_w_tc: unsupported M");
    return FGC_ERR_UNSUPPORTED;
  }
  using Cfg = WCfg<8>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  FGC_CUDA(cudaFuncSetAttribute(bwd_w_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  bwd_w_tc_kernel<8><<<bwd_w_tc_grid(rows), kThreadsW, Cfg::SMEM_BYTES, st>>>(p);
  FGC_LAUNCHED("bwd_w_tc_kernel");
  return FGC_OK;
}

}  // namespace fgc

namespace fgc {

bool bwd_src_tc_supported(int Cw, int Cout, int M, int Cin) { return Cw == 64 && Cout == 64 && M == 8 && Cin % 4 == 0; }

// wimg/wunscale: the transposed weight image prepared for bwd_tgt (prep_w_image_kernel, transposed = 1)
int launch_bwd_src_tc(const float* gy, const float* x, const int32_t* adj, const float* uvx, const void* wimg,
                      const float* wunscale, float* da_edge, float* d_uvx, float* inv_out, int64_t rows, int N,
                      int K, int Cin, int M, cudaStream_t st) {
  if (M != 8) {
    set_error("bwd_src_tc: unsupported M");
    return FGC_ERR_UNSUPPORTED;
  }
  using Cfg = SCfg<8>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  SParams p{};
  p.src = AggSrc{x, Cin, adj, uvx, N, K, rows, nullptr, nullptr, nullptr, nullptr};
  p.gy = gy, p.wimg = static_cast<const uint4*>(wimg), p.wunscale = wunscale;
  p.da_edge = da_edge, p.d_uvx = d_uvx, p.inv_out = inv_out;
  p.ntiles = (rows + kSrcTile - 1) / kSrcTile;
  FGC_CUDA(cudaFuncSetAttribute(bwd_src_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  int64_t grid = num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid < 1) grid = 1;
  bwd_src_tc_kernel<8><<<static_cast<unsigned>(grid), kSrcThreads, Cfg::SMEM_BYTES, st>>>(p);
  FGC_LAUNCHED("bwd_src_tc_kernel");
  return FGC_OK;
}

}  // namespace fgc
