// "Dense-assignment" tensor-core facet-graph convolution (second-generation dense path).
// Replaces reference Code/model.py:427-504 for 64-channel dense layers; same math as conv_fwd_tc.cu
// but BOTH contractions run on tcgen05 and every neighbour row is fetched once per tile:
//
//   tile = TF = 128/M facets.  A caller-owned *tile plan* (built once per adjacency, like the
//   reversed adjacency) lists the R distinct neighbour rows a tile touches and, per (facet,slot),
//   the local index of its row and the multiplicity of duplicates.
//
//   stage 1   S[(f,m), c] = sum_r Q[(f,m), r] * X[r, c]          tcgen05.mma  M=128 N=64 K=R
//             Q = soft assignments scattered into a dense [128 x R] fp16 hi/lo tile (K-major,
//             128B swizzle) by CUDA cores; X = the R distinct rows of a pre-split fp16 hi|lo image
//             of x, copied global->shared with cp.async straight into the UMMA MN-major layout.
//             hi.hi + lo.hi + hi.lo accumulate in one fp32 TMEM accumulator.
//   drain     S (TMEM) -> registers -> fp16 hi/lo -> shared, laid out as the B operand of stage 2
//   stage 2   Y^T[o, f] = sum_{m,c} [Wh;Wl][o,(m,c)] * [Sh|Sl][(m,c), f]   tcgen05.mma M=128 N=32 K=M*64
//             A = the resident swizzled weight image shared with conv_fwd_tc.cu.
//   epilogue  y = act( scale_f * (Wh.Sh + Wh.Sl + 2^-11 Wl.Sh) + flag_f * b )
//
// Warp roles (one persistent CTA per SM, 20 warps): 0-3 drain (S: TMEM -> fp16 hi/lo -> smem),
// 4-7 epilogue (Y: TMEM -> global) -- both sets aligned to the TMEM lane quadrants; the drain warps
// never touch global memory, so their proxy fence (a MEMBAR) only waits for their own shared stores --
// 8-11 / 12-15 two assignment groups (softmax, Q scatter; alternate chunks, one Q buffer each),
// 16-17 row loaders (cp.async, alternate chunks), 18 stage-1 MMA issuer, 19 stage-2 MMA issuer.
#include "conv_common.cuh"
#include "conv_launch.cuh"
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace fgc {

namespace {

constexpr int kC = 64;              // aggregation channels
constexpr int kRC = 64;             // distinct rows per chunk (one K atom of Q)
constexpr int kMmaThreads = 20 * 32;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// copies 16 bytes when ok, writes 16 zero bytes otherwise (src-size 0: the source is not read)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when all cp.async of this thread so far have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// TMA gather of four rows of the fp16 image (2-D tensor map: 128 halves per row = hi plane | lo plane, box
// 64 x 1, 128-byte swizzle): rows r0..r3 of plane `col` (0 or 64) land as four consecutive swizzled 128-byte
// rows at dst; 512 bytes are counted on the mbarrier.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

// MN-major, 128-byte-swizzled operand descriptor (rows of 64 MN elements = 128 B, 8-row K groups)
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int M, int COUT>
struct MCfg {
  static_assert(M == 8 && COUT == 64, "first instantiation: M = 8, Cout = 64");
  static constexpr int TF = 128 / M;                 // facets per tile; tile row = m * TF + f
  static constexpr int NB3 = 2 * TF;                 // stage-2 N: [Sh | Sl]
  static constexpr int NX = 4;                       // row-chunk ring depth
  static constexpr int X_PLANE = kRC * 128;          // one plane (hi or lo) of a row chunk
  // ring slot: X hi | X lo | neighbour logits of the chunk's rows | tile header + pair records | own logits
  static constexpr int SL_VL = 2 * X_PLANE;          // [kRC][M] fp32
  static constexpr int SL_PR = SL_VL + kRC * M * 4;  // 16-byte header (R) + TF*K uint16
  static constexpr int SL_UO = SL_PR + 16 + TF * 32 * 2;
  static constexpr int SL_VI = SL_UO + TF * M * 4;   // [kRC] fp32: 1/cnt of the chunk's rows (target-centric mode)
  static constexpr int X_BUF = ((SL_VI + kRC * 4) + 1023) / 1024 * 1024;
  static constexpr int Q_PLANE = 128 * 128;          // [128 rows][64 K] halves
  static constexpr int Q_BUF = 2 * Q_PLANE;          // hi | lo
  static constexpr int B3_ATOM = NB3 * 128;          // [NB3 rows][64 K] halves
  static constexpr int B3_BUF = M * B3_ATOM;
  static constexpr int GZ_BUF = 2 * TF * 128;        // weight-gradient mode: gy image rows of the tile's facets,
  static constexpr int OFF_X = 0;                    //   [TF rows hi | TF rows lo][64 o] halves (MN-major B operand);
                                                     //   buffer b shares the full/free barriers of B3 buffer b
  static constexpr int OFF_Q = OFF_X + NX * X_BUF;
  static constexpr int OFF_B3 = OFF_Q + 2 * Q_BUF;
  static constexpr int OFF_GZ = OFF_B3 + 2 * B3_BUF;
  static constexpr int OFF_EX = OFF_GZ + 2 * GZ_BUF;
  static constexpr int OFF_BAR = OFF_EX;
  static constexpr int SMEM_BYTES = OFF_BAR + 512;
  // TMEM: the weight operand [Wh;Wl] lives here for the CTA's lifetime (A of stage 2, 2 halves/column)
  static constexpr int W_COL = 0;
  static constexpr int W_COLS = M * kC / 2;          // 256
  static constexpr int D1_COL = W_COL + W_COLS;      // two stage-1 accumulators of 64 columns
  static constexpr int D3_COL = D1_COL + 2 * 64;     // two stage-2 accumulators of NB3 columns
  // weight-gradient mode: no resident weights; D1 at 0, the M/2 gW accumulators of 64 columns behind it
  static constexpr int GW_D1_COL = 0;
  static constexpr int GW_COL = 128;
  static constexpr int TMEM_COLS = 512;
  static_assert(D3_COL + 2 * NB3 <= TMEM_COLS, "TMEM overflow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

struct MmaParams {
  CUtensorMap tmap;        // 2-D map of img for the TMA row gather (use_tma)
  int use_tma;
  const uint4* img;        // [rows][16]: 8 x 16 B hi | 8 x 16 B lo (fp16, scaled by 2^-ex)
  const float* xunscale;   // 2^ex
  const float* uvx;        // [rows][2M]
  const int32_t* adj;      // [rows][K]
  const uint8_t* ppair;    // [ntiles][16 + TF*K*2]: header {R} + records: local row index (bits 0-8) |
                           // multiplicity << 9 (bits 9-14; 0 for repeated ids) | valid id << 15
  const int32_t* prow;     // [ntiles][TF*K]: distinct rows of the tile (global row ids)
  const int32_t* pR;       // [ntiles]
  const float* pinv;       // [rows]: 1/cnt or 0
  const uint4* wimg;       // swizzled fp16 image of [Wh | Wl] (prep_w_image_kernel)
  const float* wunscale;
  const float* b;
  float* y;
  int64_t rows, ntiles;
  int N, K, ldy, bias_mask, act;
  float alpha;
  int trace;
  // logits: a = uvx[centre][uo_off : uo_off+M] + uvx[other][vl_off : vl_off+M]
  //   forward         centre = facet, other = neighbour:  uo_off = 0, vl_off = M
  //   target-centric  centre = target, other = source:    uo_off = M, vl_off = 0, q *= inv_src[other]
  int uo_off, vl_off;
  const float* inv_src;    // nullptr in forward mode
  int mode;                // 0: y = act(inv*scale*acc + flag*b)   1: y = scale*acc (gx of the target pass)
                           // 2: weight gradient -- stage 2 is gW0[m] += (inv S_m)^T gy over the CTA's tiles,
                           //    y = this CTA's partial gW0 [M][COUT][64] (partW + blockIdx.x * M*COUT*64)
  const uint4* gimg;       // mode 2: fp16 hi|lo image of gy
  const float* gunscale;   // mode 2
};

enum {
  B_X_FULL = 0,            // NX
  B_X_FREE = 4,            // NX
  B_Q_FULL = 8,            // 2
  B_Q_FREE = 10,           // 2
  B_D1_FULL = 12,          // 2
  B_D1_FREE = 14,          // 2
  B_B3_FULL = 16,          // 2
  B_B3_FREE = 18,          // 2
  B_D3_FULL = 20,          // 2
  B_D3_FREE = 22,          // 2
  B_GW_DONE = 24,          // 1 (mode 2; last stage-2 commit)
  B_NUM = 25
};

// optional pipeline trace (FGC_MMA_TRACE=1): clock64 stamps of CTA 0, first 32 tiles, 4 roles x 8 events
__device__ long long g_mma_trace[4 * 32 * 8];
#define FGC_TR(role, t, ev)                                                        \
  do {                                                                             \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (t) < 32)                       \
      g_mma_trace[((role) * 32 + (t)) * 8 + (ev)] = clock64();                     \
  } while (0)

__device__ __forceinline__ int chunks_of(int R) { return R <= kRC ? 1 : (R + kRC - 1) / kRC; }

template <int M, int COUT, int KP, bool TMA>
__global__ void __launch_bounds__(kMmaThreads, 1)
conv_mma_kernel(const __grid_constant__ MmaParams p) {
  using Cfg = MCfg<M, COUT>;
  constexpr int TF = Cfg::TF;
  constexpr int NX = Cfg::NX;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_NUM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = TF * p.K;  // (facet, slot) pairs per tile

  if (threadIdx.x == 0) {
    for (int i = 0; i < NX; ++i) tc::mbar_init(&bars[B_X_FULL + i], 32), tc::mbar_init(&bars[B_X_FREE + i], 5);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars[B_Q_FULL + i], 4), tc::mbar_init(&bars[B_Q_FREE + i], 1);
      tc::mbar_init(&bars[B_D1_FULL + i], 1), tc::mbar_init(&bars[B_D1_FREE + i], 4);
      tc::mbar_init(&bars[B_B3_FULL + i], 4), tc::mbar_init(&bars[B_B3_FREE + i], 1);
      tc::mbar_init(&bars[B_D3_FULL + i], 1), tc::mbar_init(&bars[B_D3_FREE + i], 4);
    }
    tc::mbar_init(&bars[B_GW_DONE], 1);
    tc::mbar_fence_init();
  }
  if (warp == 18) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  {
    // stale shared memory may hold NaN bit patterns: rows beyond R are multiplied by zero columns of Q
    uint4* z = reinterpret_cast<uint4*>(smem + Cfg::OFF_X);
    for (int i = threadIdx.x; i < (Cfg::OFF_EX - Cfg::OFF_X) / 16; i += kMmaThreads) z[i] = make_uint4(0, 0, 0, 0);
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  const bool gw_mode = p.mode == 2;
  const uint32_t d1_col = gw_mode ? Cfg::GW_D1_COL : Cfg::D1_COL;
  if (warp < 4 && !gw_mode) {
    // weight operand -> TMEM, column = K pair; un-swizzle the shared-memory image (16-byte units XOR
    // row & 7) on the way.  Lane 32q + 16h + i holds row h*COUT + 16q + i of [Wh;Wl]: the hi and lo
    // rows of output channel 16q + i sit in the same warp, so the epilogue combines them by shuffle.
    {
      const int row = ((lane >> 4) & 1) * COUT + warp * 16 + (lane & 15);
#pragma unroll 1
      for (int mm = 0; mm < M; ++mm) {
        const uint4* src = p.wimg + (static_cast<size_t>(mm) * 2 * COUT + row) * 8;
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
  }
  __syncthreads();
  tc::tc_fence_after_sync();

  if (warp < 4) {
    // =========================================================== drain: S (TMEM) -> fp16 hi/lo -> B3
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int trow = warp * 32 + lane;           // tile row m * TF + f
    const int m = trow / TF, f = trow % TF;
    const int nh = f, nl = TF + f;               // B3 rows of the hi / lo value of facet f
    const int b3h_off = m * Cfg::B3_ATOM + (nh >> 3) * 1024 + (nh & 7) * 128;
    const int b3l_off = m * Cfg::B3_ATOM + (nl >> 3) * 1024 + (nl & 7) * 128;
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      // weight-gradient mode: S rows are scaled by 1/cnt of their facet (gz = inv gy, folded into S)
      // and the drain warps also stage the gy image rows of the tile (thread -> facet tid/8, 16-byte unit
      // tid%8 of the hi and of the lo plane): loaded here, long before the fence below
      float rsc = 1.f;
      uint4 gzh = make_uint4(0, 0, 0, 0), gzl = gzh;
      if (gw_mode) {
        rsc = (tile * TF + f < p.rows) ? __ldg(p.pinv + tile * TF + f) : 0.f;
        const int64_t gr = tile * TF + (threadIdx.x >> 3);
        if (gr < p.rows) {
          gzh = __ldg(p.gimg + gr * 16 + (threadIdx.x & 7));
          gzl = __ldg(p.gimg + gr * 16 + 8 + (threadIdx.x & 7));
        }
      }
      tc::mbar_wait(&bars[B_D1_FULL + buf], (t >> 1) & 1);
      if (warp == 0) FGC_TR(0, t, 0);
      tc::tc_fence_after_sync();
      uint8_t* b3 = smem + Cfg::OFF_B3 + buf * Cfg::B3_BUF;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tc::tmem_ld32(tmem + lane_base + d1_col + buf * 64 + half * 32, v);
        tc::tc_wait_ld();
        if (gw_mode) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * rsc);
        }
        if (half == 1) {
          tc::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars[B_D1_FREE + buf]);
        }
        // fp16 hi (truncated to 11 significant bits, exactly representable) + fp16 residual
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = __uint_as_float(v[2 * i]), a1 = __uint_as_float(v[2 * i + 1]);
          const float h0 = __uint_as_float(v[2 * i] & 0xFFFFE000u), h1 = __uint_as_float(v[2 * i + 1] & 0xFFFFE000u);
          const __half2 hh = __floats2half2_rn(h0, h1);
          const __half2 ll = __floats2half2_rn(a0 - h0, a1 - h1);
          hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        if (half == 0) {
          if (warp == 0) FGC_TR(0, t, 1);
          tc::mbar_wait(&bars[B_B3_FREE + buf], ((t >> 1) & 1) ^ 1);
          if (warp == 0) FGC_TR(0, t, 2);
          if (gw_mode) {
            uint8_t* gz = smem + Cfg::OFF_GZ + buf * Cfg::GZ_BUF;
            const int kh = threadIdx.x >> 3, kl = TF + kh, cc = threadIdx.x & 7;
            *reinterpret_cast<uint4*>(gz + (kh >> 3) * 1024 + (kh & 7) * 128 + ((cc ^ (kh & 7)) << 4)) = gzh;
            *reinterpret_cast<uint4*>(gz + (kl >> 3) * 1024 + (kl & 7) * 128 + ((cc ^ (kl & 7)) << 4)) = gzl;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int jj = half * 4 + j;
          *reinterpret_cast<uint4*>(b3 + b3h_off + ((jj ^ (nh & 7)) << 4)) =
              make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          *reinterpret_cast<uint4*>(b3 + b3l_off + ((jj ^ (nl & 7)) << 4)) =
              make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[B_B3_FULL + buf]);
      if (warp == 0) FGC_TR(0, t, 3);
    }
  } else if (warp < 8) {
    // =========================================================== epilogue: Y (TMEM) -> global
    const int q = warp - 4;                      // TMEM lane quadrant
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const float wun = gw_mode ? 1.f : __ldg(p.wunscale), xun = __ldg(p.xunscale);
    static_assert(TF == 16, "epilogue written for 16-facet tiles");
    // lane 16h + i of quadrant q: h = 0 holds Wh.(Sh|Sl) of channel o = 16q + i, h = 1 holds Wl.Sh.
    // After one shuffle round lane (h, i) owns facets 8h .. 8h+7 of channel o.
    const int hh = lane >> 4, o = q * 16 + (lane & 15);
    const float bo = (p.mode == 0) ? __ldg(p.b + o) : 0.f;
    const float sc0 = xun * wun;
    float* const ybase = p.y + o;
    const int64_t ldy = p.ldy;
    const bool unmasked = !p.bias_mask;
    int t = 0;
    if (gw_mode) {
      // weight-gradient mode: one read-out per CTA.  TMEM lane 64 parity + c, column 64 i + o holds
      // gW0[2i + parity][o][c] in image units; partial of this CTA -> p.y + blockIdx.x * M*COUT*64
      const float usc = xun * __ldg(p.gunscale);
      const int row = q * 32 + lane, par = row >> 6, cch = row & 63;
      float* out = p.y + static_cast<size_t>(blockIdx.x) * (M * COUT * kC);
      tc::mbar_wait(&bars[B_GW_DONE], 0);
      tc::tc_fence_after_sync();
#pragma unroll 1
      for (int i = 0; i < M / 2; ++i) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t d[32];
          tc::tmem_ld32(tmem + lane_base + Cfg::GW_COL + i * 64 + half * 32, d);
          tc::tc_wait_ld();
          float* dst = out + (static_cast<size_t>(2 * i + par) * COUT + half * 32) * kC + cch;
#pragma unroll
          for (int oo = 0; oo < 32; ++oo) dst[oo * kC] = __uint_as_float(d[oo]) * usc;
        }
      }
      tc::tc_fence_before_sync();
    } else
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      const int64_t r0 = tile * TF + 8 * hh;
      const int nv = (p.rows - r0 >= 8) ? 8 : static_cast<int>(p.rows - r0);   // valid facets of this lane's 8 (<= 0: none)
      float inv[8];
      if (p.mode == 0) {
        if (nv == 8) {   // r0 is a multiple of 8: two aligned 16-byte loads
          const float4 i0 = __ldg(reinterpret_cast<const float4*>(p.pinv + r0));
          const float4 i1 = __ldg(reinterpret_cast<const float4*>(p.pinv + r0) + 1);
          inv[0] = i0.x, inv[1] = i0.y, inv[2] = i0.z, inv[3] = i0.w;
          inv[4] = i1.x, inv[5] = i1.y, inv[6] = i1.z, inv[7] = i1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) inv[j] = (j < nv) ? __ldg(p.pinv + r0 + j) : 0.f;
        }
      }
      tc::mbar_wait(&bars[B_D3_FULL + buf], (t >> 1) & 1);
      if (q == 0) FGC_TR(0, t, 4);
      tc::tc_fence_after_sync();
      uint32_t d[32];
      tc::tmem_ld32(tmem + lane_base + Cfg::D3_COL + buf * Cfg::NB3, d);
      tc::tc_wait_ld();
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[B_D3_FREE + buf]);
      // all eight exchanges are issued before the first is consumed
      float keep[8], recv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float hi_lo = __uint_as_float(d[j]) + __uint_as_float(d[TF + j]);         // facets 0..7  (hi lanes)
        const float hi_up = __uint_as_float(d[8 + j]) + __uint_as_float(d[TF + 8 + j]); // facets 8..15 (hi lanes)
        const float lo_lo = __uint_as_float(d[j]) * (1.f / 2048.f);                     // facets 0..7  (lo lanes)
        const float lo_up = __uint_as_float(d[8 + j]) * (1.f / 2048.f);                 // facets 8..15 (lo lanes)
        keep[j] = hh ? lo_up : hi_lo;
        recv[j] = hh ? lo_lo : hi_up;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) recv[j] = __shfl_xor_sync(0xffffffffu, recv[j], 16);
      float yv[8];
      if (p.mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // hi lanes: (Wh.Sh + Wh.Sl) + Wl.Sh/2048; lo lanes: the same sum in the same order
          const float acc = hh ? (recv[j] + keep[j]) : (keep[j] + recv[j]);
          const float fl = (inv[j] > 0.f || unmasked) ? bo : 0.f;
          yv[j] = fmaf(inv[j] * sc0, acc, fl);
        }
        if (p.act == FGC_ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) yv[j] = lrelu_f(yv[j], p.alpha);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) yv[j] = sc0 * (hh ? (recv[j] + keep[j]) : (keep[j] + recv[j]));
      }
      float* yp = ybase + r0 * ldy;
      if (nv == 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) yp[j * ldy] = yv[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nv) yp[j * ldy] = yv[j];
      }
      if (q == 0) FGC_TR(0, t, 5);
    }
  } else if (warp < 16) {
    // =========================================================== assignments: softmax + Q scatter
    // Everything these warps read (pair records, own logits, neighbour logits of the chunk's distinct
    // rows) is staged in the ring slot by the loader: no global load -- and so no load latency and
    // no MEMBAR stall -- sits between two Q tiles.
    const int grp = (warp - 8) >> 2;             // group g handles items with it % 2 == g, Q buffer g
    const int qt = threadIdx.x & 127;            // 0..127 within the group
    const int f = qt >> 3, s = qt & 7;           // facet of the tile, slot lane: slots s, s+8, ...
    const bool tracer = (warp == 8);
    int it = 0;
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const bool rv = tile * TF + f < p.rows;
      uint32_t rec[KP];
      float uo[M];
      const int nch = chunks_of(Rn);
      if (tile + gridDim.x < p.ntiles) Rn = __ldg(p.pR + tile + gridDim.x);
      bool have = false;
      for (int c = 0; c < nch; ++c, ++it) {
        if ((it & 1) != grp) continue;
        const int xbuf = it % NX, qb = grp;
        const int itg = it >> 1;                 // this group's item counter
        const uint8_t* slot = smem + Cfg::OFF_X + xbuf * Cfg::X_BUF;
        if (tracer) FGC_TR(1, itg, 3);
        tc::mbar_wait(&bars[B_X_FULL + xbuf], (it / NX) & 1);
        if (tracer) FGC_TR(1, itg, 4);
        if (!have) {
          // pair records and own logits are staged with chunk 0 only; a group that first meets the
          // tile at a later chunk reads them from global memory (rare: tiles with > 64 distinct rows)
          have = true;
          if (c == 0) {
            const uint16_t* pr = reinterpret_cast<const uint16_t*>(slot + Cfg::SL_PR + 16);
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              const int k = s + 8 * j;
              rec[j] = (rv && k < p.K) ? pr[f * p.K + k] : 0u;
            }
            const float4* up = reinterpret_cast<const float4*>(slot + Cfg::SL_UO + f * (M * 4));
#pragma unroll
            for (int i = 0; i < M; i += 4) {
              const float4 tq = up[i / 4];
              uo[i] = tq.x, uo[i + 1] = tq.y, uo[i + 2] = tq.z, uo[i + 3] = tq.w;
            }
          } else {
            const uint16_t* pr = reinterpret_cast<const uint16_t*>(p.ppair + tile * (16 + P * 2) + 16);
#pragma unroll
            for (int j = 0; j < KP; ++j) {
              const int k = s + 8 * j;
              rec[j] = (rv && k < p.K) ? __ldg(pr + f * p.K + k) : 0u;
            }
#pragma unroll
            for (int i = 0; i < M; i += 4) {
              float4 tq = make_float4(0.f, 0.f, 0.f, 0.f);
              if (rv) tq = __ldg(reinterpret_cast<const float4*>(p.uvx + (tile * TF + f) * (2 * M) + p.uo_off + i));
              uo[i] = tq.x, uo[i + 1] = tq.y, uo[i + 2] = tq.z, uo[i + 3] = tq.w;
            }
          }
        }
        if (tracer) FGC_TR(1, itg, 5);
        uint32_t qhp[KP][M / 2], qlp[KP][M / 2];
        int col[KP];
        // branch-free: every lane evaluates all KP pairs (clamped row for the inactive ones), so
        // the KP softmax chains interleave; inactive pairs are dropped at the scatter
        float a[KP][M];
        float rs[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          const int mult = (rec[j] >> 9) & 63, lidx = rec[j] & 511;
          col[j] = (mult && (lidx >> 6) == c) ? (lidx & 63) : -1;
          const float4* vp = reinterpret_cast<const float4*>(slot + Cfg::SL_VL + (lidx & 63) * (M * 4));
#pragma unroll
          for (int i = 0; i < M; i += 4) {
            const float4 tq = vp[i / 4];
            a[j][i] = uo[i] + tq.x, a[j][i + 1] = uo[i + 1] + tq.y, a[j][i + 2] = uo[i + 2] + tq.z, a[j][i + 3] = uo[i + 3] + tq.w;
          }
          rs[j] = static_cast<float>(mult);
          if (p.inv_src != nullptr) rs[j] *= *reinterpret_cast<const float*>(slot + Cfg::SL_VI + (lidx & 63) * 4);
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          float mx = a[j][0];
#pragma unroll
          for (int i = 1; i < M; ++i) mx = fmaxf(mx, a[j][i]);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < M; ++i) {
            a[j][i] = exp2f((a[j][i] - mx) * 1.4426950408889634f);
            sum += a[j][i];
          }
          rs[j] = __fdividef(rs[j], sum);
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) {
#pragma unroll
          for (int i = 0; i < M; i += 2) {
            const float q0 = a[j][i] * rs[j], q1 = a[j][i + 1] * rs[j];
            const __half2 hh = __floats2half2_rn(q0, q1);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(q0 - hf.x, q1 - hf.y);
            qhp[j][i / 2] = *reinterpret_cast<const uint32_t*>(&hh);
            qlp[j][i / 2] = *reinterpret_cast<const uint32_t*>(&ll);
          }
        }
        // the staged data of this slot is consumed
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_X_FREE + xbuf]);
        if (tracer) FGC_TR(1, itg, 0);
        tc::mbar_wait(&bars[B_Q_FREE + qb], (itg & 1) ^ 1);
        if (tracer) FGC_TR(1, itg, 1);
        uint8_t* qh = smem + Cfg::OFF_Q + qb * Cfg::Q_BUF;
        uint8_t* ql = qh + Cfg::Q_PLANE;
        // zero both planes (consecutive threads, consecutive 16-byte units), then scatter
        {
          uint4* z = reinterpret_cast<uint4*>(qh);
#pragma unroll
          for (int j = 0; j < Cfg::Q_BUF / 16 / 128; ++j) z[qt + 128 * j] = make_uint4(0, 0, 0, 0);
        }
        if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          if (col[j] >= 0) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
              const int row = i * TF + f;    // tile row of (facet f, weight i)
              const int off = (row >> 3) * 1024 + (row & 7) * 128 + (((col[j] >> 3) ^ (row & 7)) << 4) + (col[j] & 7) * 2;
              const uint32_t wh = qhp[j][i / 2], wl = qlp[j][i / 2];
              *reinterpret_cast<uint16_t*>(qh + off) = static_cast<uint16_t>((i & 1) ? (wh >> 16) : (wh & 0xFFFF));
              *reinterpret_cast<uint16_t*>(ql + off) = static_cast<uint16_t>((i & 1) ? (wl >> 16) : (wl & 0xFFFF));
            }
          }
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_Q_FULL + qb]);
        if (tracer) FGC_TR(1, itg, 2);
      }
    }
  } else if (warp == 16 || warp == 17) {
    // =========================================================== row loaders (cp.async)
    // Per (tile, chunk): the chunk's distinct rows (fp16 hi|lo image) into the UMMA layout, their
    // neighbour logits, and for chunk 0 the tile's pair records and own logits.  Completion is
    // signalled by cp.async.mbarrier.arrive.noinc (no wait, no fence: the loader never blocks on its
    // own copies); row ids of the next chunk are fetched while the current copies are issued.
    // Lane l owns rows l and l + 32 of the chunk (two row ids in registers, all 16-byte units of the
    // row copied by the same lane: the per-copy address arithmetic is a constant offset).
    const uint32_t xbase = tc::smem_u32(smem + Cfg::OFF_X);
    const int64_t step = gridDim.x;
    const int pr_bytes = 16 + P * 2;
    const uint32_t row_off = (lane >> 3) * 1024 + (lane & 7) * 128;   // row l inside a plane; row l+32: +4096
    auto load_ids = [&](int64_t tile, int c, int R, int (&rid)[2]) {
      const int rc = (tile < p.ntiles) ? min(kRC, R - c * kRC) : 0;
      const int32_t* rl = p.prow + tile * P + c * kRC;
      rid[0] = (lane < rc) ? __ldg(rl + lane) : -1;
      rid[1] = (lane + 32 < rc) ? __ldg(rl + lane + 32) : -1;
    };
    // Two loader warps take alternate items (it % 2); each walks the whole (tile, chunk) sequence.
    const int me = warp - 16;
    int it = 0;
    int64_t tile = blockIdx.x;
    int c = 0;
    int R = (tile < p.ntiles) ? __ldg(p.pR + tile) : 0;
    int Rn = (tile + step < p.ntiles) ? __ldg(p.pR + tile + step) : 0;
    auto advance = [&]() {   // next item of the sequence
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
      // my next item is two steps ahead
      const int64_t tile0 = tile;
      const int c0 = c, it0 = it, R0 = R;
      advance();
      if (tile < p.ntiles) advance();
      load_ids(tile, c, R, nxt);
      const int buf = it0 % NX;
      if (me == 0) FGC_TR(2, it0 >> 1, 0);
      tc::mbar_wait(&bars[B_X_FREE + buf], ((it0 / NX) & 1) ^ 1);
      if (me == 0) FGC_TR(2, it0 >> 1, 1);
      const uint32_t sl = xbase + buf * Cfg::X_BUF;
      if constexpr (TMA) {
        // row planes by TMA gather: lane j < 16 fetches rows 4j .. 4j+3 of the chunk (their ids are the
        // L1-hot words load_ids read one item ago), hi plane then lo plane, 512 bytes each
        const int rc0 = min(kRC, R0 - c0 * kRC);
        if (lane < 16 && 4 * lane < rc0) {
          int4 r4 = __ldg(reinterpret_cast<const int4*>(p.prow + tile0 * P + c0 * kRC) + lane);
          const int nv = rc0 - 4 * lane;           // rows past the end repeat the first one (zero columns of Q)
          if (nv < 2) r4.y = r4.x;
          if (nv < 3) r4.z = r4.x;
          if (nv < 4) r4.w = r4.x;
          const uint32_t dst = sl + (lane >> 1) * 1024 + (lane & 1) * 512;
          mbar_expect_tx(&bars[B_X_FULL + buf], 1024);
          tma_gather4(dst, &p.tmap, &bars[B_X_FULL + buf], 0, r4.x, r4.y, r4.z, r4.w);
          tma_gather4(dst + Cfg::X_PLANE, &p.tmap, &bars[B_X_FULL + buf], 64, r4.x, r4.y, r4.z, r4.w);
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (cur[rr] >= 0) {
          const uint4* src = p.img + static_cast<int64_t>(cur[rr]) * 16;
          const uint32_t dst = sl + row_off + rr * 4096;
          if constexpr (!TMA) {
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              cp_async16(dst + ((cc ^ (lane & 7)) << 4), src + cc);
              cp_async16(dst + Cfg::X_PLANE + ((cc ^ (lane & 7)) << 4), src + 8 + cc);
            }
          }
          const float* vsrc = p.uvx + static_cast<int64_t>(cur[rr]) * (2 * M) + p.vl_off;
          const uint32_t vdst = sl + Cfg::SL_VL + (lane + 32 * rr) * (M * 4);
#pragma unroll
          for (int q = 0; q < M / 4; ++q) cp_async16(vdst + q * 16, vsrc + q * 4);
          if (p.inv_src != nullptr) cp_async4(sl + Cfg::SL_VI + (lane + 32 * rr) * 4, p.inv_src + cur[rr]);
        }
      }
      if (c0 == 0) {
        const uint8_t* src = p.ppair + tile0 * pr_bytes;
        for (int q = lane * 16; q < pr_bytes; q += 512) cp_async16(sl + Cfg::SL_PR + q, src + q);
        const int64_t r = tile0 * TF + (lane >> 1);
        if ((lane >> 1) < TF && r < p.rows)
          cp_async16(sl + Cfg::SL_UO + (lane >> 1) * (M * 4) + (lane & 1) * 16,
                     p.uvx + r * (2 * M) + p.uo_off + (lane & 1) * 4);
      }
      cp_async_arrive_noinc(&bars[B_X_FULL + buf]);
      if (me == 0) FGC_TR(2, it0 >> 1, 2);
    };
    while (tile < p.ntiles) {
      issue(ids[0], ids[1]);
      if (tile >= p.ntiles) break;
      issue(ids[1], ids[0]);
    }
  } else if (warp == 18) {
    // =========================================================== stage-1 MMA issuer
    // A = Q (smem, K-major), B = X rows (smem, MN-major)
    constexpr uint32_t idesc1 = (1u << 4) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sb = tc::smem_u32(smem);
    int t = 0, it = 0;
    int Rn = (blockIdx.x < p.ntiles) ? __ldg(p.pR + blockIdx.x) : 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int R = Rn;
      if (tile + gridDim.x < p.ntiles) Rn = __ldg(p.pR + tile + gridDim.x);
      const int nch = chunks_of(R);
      const int dbuf = t & 1;
      FGC_TR(3, t, 0);
      tc::mbar_wait(&bars[B_D1_FREE + dbuf], ((t >> 1) & 1) ^ 1);
      FGC_TR(3, t, 1);
      for (int c = 0; c < nch; ++c, ++it) {
        const int xbuf = it % NX, qb = it & 1;
        const int rc = min(kRC, R - c * kRC);
        const int nks = max(1, (rc + 15) >> 4);
        tc::mbar_wait(&bars[B_X_FULL + xbuf], (it / NX) & 1);
        FGC_TR(3, t, 7);
        tc::mbar_wait(&bars[B_Q_FULL + qb], (it >> 1) & 1);
        FGC_TR(3, t, 2);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t xh = sb + Cfg::OFF_X + xbuf * Cfg::X_BUF, xl = xh + Cfg::X_PLANE;
          const uint32_t qh = sb + Cfg::OFF_Q + qb * Cfg::Q_BUF, ql = qh + Cfg::Q_PLANE;
          const uint32_t d1 = tmem + d1_col + dbuf * 64;
#pragma unroll 1
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t aqh = tc::smem_desc_k_sw128(qh + ks * 32);
            const uint64_t aql = tc::smem_desc_k_sw128(ql + ks * 32);
            const uint64_t bxh = desc_mn_sw128(xh + ks * 2048, 1024, 1024);
            const uint64_t bxl = desc_mn_sw128(xl + ks * 2048, 1024, 1024);
            tc::mma_f16_ss(d1, aqh, bxh, idesc1, (c | ks) ? 1u : 0u);
            tc::mma_f16_ss(d1, aql, bxh, idesc1, 1u);
            tc::mma_f16_ss(d1, aqh, bxl, idesc1, 1u);
          }
          tc::tc_commit(&bars[B_X_FREE + xbuf]);
          tc::tc_commit(&bars[B_Q_FREE + qb]);
          if (c == nch - 1) tc::tc_commit(&bars[B_D1_FULL + dbuf]);
        }
        __syncwarp();
      }
      FGC_TR(3, t, 3);
    }
  } else {
    // =========================================================== stage-2 MMA issuer
    // A = [Wh;Wl] (TMEM), B = S (smem, K-major)
    constexpr uint32_t idesc3 = (1u << 4) | ((static_cast<uint32_t>(Cfg::NB3) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sb = tc::smem_u32(smem);
    int t = 0;
    if (gw_mode) {
      // weight gradient: D_i[(m parity, c), o] += sum_f (inv S)[f, 2i + parity, c] gy[f, o], i = 0..M/2-1
      // A = two B3 atoms read MN-major (M = 128: 64 channels of weight 2i, then of 2i+1; K = 16 facets),
      // B = the gy rows (MN-major, N = 64); hi.hi + lo.hi + hi.lo into one fp32 accumulator that stays
      // in TMEM for the CTA's lifetime.
      constexpr uint32_t idescW = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
        const int buf = t & 1;
        tc::mbar_wait(&bars[B_B3_FULL + buf], (t >> 1) & 1);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t b3 = sb + Cfg::OFF_B3 + buf * Cfg::B3_BUF;
          const uint32_t gz = sb + Cfg::OFF_GZ + buf * Cfg::GZ_BUF;
          const uint64_t bgh = desc_mn_sw128(gz, 1024, 1024);
          const uint64_t bgl = desc_mn_sw128(gz + TF * 128, 1024, 1024);
#pragma unroll
          for (int i = 0; i < M / 2; ++i) {
            const uint64_t ash = desc_mn_sw128(b3 + 2 * i * Cfg::B3_ATOM, Cfg::B3_ATOM, 1024);
            const uint64_t asl = desc_mn_sw128(b3 + 2 * i * Cfg::B3_ATOM + TF * 128, Cfg::B3_ATOM, 1024);
            const uint32_t dw = tmem + Cfg::GW_COL + i * 64;
            tc::mma_f16_ss(dw, ash, bgh, idescW, t ? 1u : 0u);
            tc::mma_f16_ss(dw, asl, bgh, idescW, 1u);
            tc::mma_f16_ss(dw, ash, bgl, idescW, 1u);
          }
          tc::tc_commit(&bars[B_B3_FREE + buf]);
          if (tile + gridDim.x >= p.ntiles) tc::tc_commit(&bars[B_GW_DONE]);
        }
        __syncwarp();
      }
    } else
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int buf = t & 1;
      FGC_TR(3, t, 4);
      tc::mbar_wait(&bars[B_B3_FULL + buf], (t >> 1) & 1);
      tc::mbar_wait(&bars[B_D3_FREE + buf], ((t >> 1) & 1) ^ 1);
      FGC_TR(3, t, 5);
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t b3 = sb + Cfg::OFF_B3 + buf * Cfg::B3_BUF;
#pragma unroll 1
        for (int mm = 0; mm < M; ++mm) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bd = tc::smem_desc_k_sw128(b3 + mm * Cfg::B3_ATOM + ks * 32);
            tc::mma_f16_ts(tmem + Cfg::D3_COL + buf * Cfg::NB3, tmem + Cfg::W_COL + mm * 32 + ks * 8, bd, idesc3,
                           (mm | ks) ? 1u : 0u);
          }
        }
        tc::tc_commit(&bars[B_B3_FREE + buf]);
        tc::tc_commit(&bars[B_D3_FULL + buf]);
      }
      __syncwarp();
      FGC_TR(3, t, 6);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 18) tc::tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------ fp16 hi|lo image of x
// img[r] = [fp16(x_r * s) (64) | fp16(x_r * s - hi) (64)], s = 2^(126-E) with E the exponent of max|x|
// Optionally (partB != nullptr, ldx == 64) also the bias gradient of the layer the rows are the gy of:
// partB[block][64] = sum over the block's rows of flag_r * x_r, flag_r = (pinv_r > 0 or no bias mask);
// fixed thread -> row assignment and a fixed-order block reduction: deterministic.
__global__ void __launch_bounds__(256)
prep_x_image_kernel(const float* __restrict__ x, int ldx, int64_t rows, const unsigned* __restrict__ maxbits,
                    uint4* __restrict__ img, float* __restrict__ xunscale, const float* __restrict__ pinv,
                    int bias_mask, float* __restrict__ partB) {
  __shared__ float red[256 * 9];
  float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int E = static_cast<int>((__ldg(maxbits) >> 23) & 0xFF);
  E = min(max(E, 16), 240);
  const float sc = __int_as_float((253 - E) << 23);
  if (blockIdx.x == 0 && threadIdx.x == 0) xunscale[0] = __int_as_float((E + 1) << 23);
  const int64_t total = rows * 8;  // one thread per 8 channels
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i >> 3;
    const int j = static_cast<int>(i & 7);
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + 2 * j);
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + 2 * j + 1);
    if (partB != nullptr) {
      const float fl = (!bias_mask || __ldg(pinv + r) > 0.f) ? 1.f : 0.f;
      bs[0] += fl * a.x, bs[1] += fl * a.y, bs[2] += fl * a.z, bs[3] += fl * a.w;
      bs[4] += fl * b.x, bs[5] += fl * b.y, bs[6] += fl * b.z, bs[7] += fl * b.w;
    }
    const float v[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2 hh = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(v[2 * q] - hf.x, v[2 * q + 1] - hf.y);
      hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    img[r * 16 + j] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    img[r * 16 + 8 + j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (partB != nullptr) {
    // thread t always works on channels 8 (t % 8) .. 8 (t % 8) + 7 (block and grid strides are multiples of 8)
#pragma unroll
    for (int q = 0; q < 8; ++q) red[threadIdx.x * 9 + q] = bs[q];
    __syncthreads();
    if (threadIdx.x < 64) {
      const int jj = threadIdx.x >> 3, q = threadIdx.x & 7;
      float acc = 0.f;
      for (int w = 0; w < 32; ++w) acc += red[(w * 8 + jj) * 9 + q];
      partB[static_cast<size_t>(blockIdx.x) * 64 + threadIdx.x] = acc;
    }
  }
}

__global__ void absmax2_kernel(const float* __restrict__ x, int64_t n4, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// ------------------------------------------------------------------ tile plan
// One block per tile, one thread per (facet, slot) pair.  Distinct rows are numbered in ascending
// row order; duplicates inside a facet's list collapse onto the first occurrence with a multiplicity
// (q depends only on the (facet, row) pair, so k equal ids contribute k * q).
__global__ void __launch_bounds__(512)
build_conv_plan_kernel(const int32_t* __restrict__ adj, int64_t rows, int N, int K, int TF,
                       uint8_t* __restrict__ ppair, int32_t* __restrict__ prow, int32_t* __restrict__ pR,
                       float* __restrict__ pinv, unsigned long long* __restrict__ total_rows) {
  extern __shared__ int32_t sm[];
  const int P = TF * K;
  int32_t* ids = sm;          // [P] global row or -1
  int32_t* first = sm + P;    // [P] first occurrence in the tile
  int32_t* cnt = sm + 2 * P;  // [TF]
  const int pidx = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int f = pidx / K, k = pidx % K;
  const int64_t r = tile * TF + f;
  if (pidx < TF) cnt[pidx] = 0;
  __syncthreads();
  int32_t g = -1;
  if (pidx < P && r < rows) {
    const int id = __ldg(adj + r * K + k);
    if (id != 0) atomicAdd(&cnt[f], 1);
    if (id > 0 && id <= N) g = static_cast<int32_t>((r / N) * N + id - 1);
  }
  if (pidx < P) ids[pidx] = g;
  __syncthreads();
  bool first_tile = g >= 0, first_facet = g >= 0;
  int mult = 0;
  if (g >= 0) {
    for (int q = 0; q < pidx; ++q)
      if (ids[q] == g) {
        first_tile = false;
        if (q >= f * K) first_facet = false;
      }
    for (int q = f * K; q < (f + 1) * K; ++q) mult += (ids[q] == g);
  }
  if (pidx < P) first[pidx] = first_tile ? 1 : 0;
  const int R = __syncthreads_count(first_tile);
  int lidx = 0;
  if (g >= 0) {
    for (int q = 0; q < P; ++q) lidx += (first[q] && ids[q] < g);
  }
  if (first_tile) prow[tile * P + lidx] = g;
  uint8_t* blk = ppair + tile * (16 + 2 * P);   // header {R, 0, 0, 0} + P pair records
  if (pidx < P)
    reinterpret_cast<uint16_t*>(blk + 16)[pidx] = (g >= 0) ? static_cast<uint16_t>(lidx | ((first_facet ? mult : 0) << 9) | 0x8000) : 0;
  if (pidx < 4) reinterpret_cast<int32_t*>(blk)[pidx] = (pidx == 0) ? R : 0;
  if (pidx == 0) {
    pR[tile] = R;
    atomicAdd(total_rows, static_cast<unsigned long long>(R));   // integer: order-independent
  }
  if (pidx < TF && tile * TF + pidx < rows) pinv[tile * TF + pidx] = cnt[pidx] ? 1.f / static_cast<float>(cnt[pidx]) : 0.f;
}

struct PlanLayout {
  int64_t rows, ntiles;
  int K, TF;
  size_t off_R, off_inv, off_row, off_pair, total;
  PlanLayout(int64_t rows_, int K_, int M) : rows(rows_), K(K_), TF(128 / M) {
    ntiles = (rows + TF - 1) / TF;
    size_t o = 256;  // header
    off_R = o, o = align_up(o + static_cast<size_t>(ntiles) * 4, 256);
    off_inv = o, o = align_up(o + static_cast<size_t>(rows) * 4, 256);
    off_row = o, o = align_up(o + static_cast<size_t>(ntiles) * TF * K * 4, 256);
    off_pair = o, o = align_up(o + static_cast<size_t>(ntiles) * (16 + TF * K * 2), 256);
    total = o;
  }
};

}  // namespace
}  // namespace fgc

#include "conv_mma_src.cuh"

namespace fgc {

bool conv_mma_supported(int Cin, int Cw, int Cout, int M, int K) {
  return Cw == 64 && Cout == 64 && M == 8 && K <= 32 && Cin % 4 == 0;
}

size_t conv_plan_bytes(int64_t rows, int K, int M) {
  if (M < 1 || 128 / M < 1) return 0;
  return PlanLayout(rows, K, M).total;
}

int build_conv_plan(const int32_t* adj, int B, int N, int K, int M, void* plan, size_t plan_bytes, cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(B) * N;
  FGC_UNSUPPORTED(M != 8, "conv plan: only M = 8 tiles are implemented (M=%d)", M);
  const PlanLayout L(rows, K, M);
  FGC_REQUIRE(plan_bytes >= L.total, "conv plan: buffer too small (%zu given, %zu needed)", plan_bytes, L.total);
  FGC_REQUIRE(L.TF * K <= 512, "conv plan: TF*K = %d exceeds 512", L.TF * K);
  char* base = static_cast<char*>(plan);
  // header: [0] sum over tiles of the distinct-row counts (int64), [1] number of tiles
  FGC_CUDA(cudaMemsetAsync(base, 0, 256, st));
  const long long nt = L.ntiles;
  FGC_CUDA(cudaMemcpyAsync(base + 8, &nt, sizeof(nt), cudaMemcpyHostToDevice, st));
  const int P = L.TF * K;
  const int threads = ((P + 31) / 32) * 32;
  build_conv_plan_kernel<<<static_cast<unsigned>(L.ntiles), threads, (2 * P + L.TF) * 4, st>>>(
      adj, rows, N, K, L.TF, reinterpret_cast<uint8_t*>(base + L.off_pair), reinterpret_cast<int32_t*>(base + L.off_row),
      reinterpret_cast<int32_t*>(base + L.off_R), reinterpret_cast<float*>(base + L.off_inv),
      reinterpret_cast<unsigned long long*>(base));
  FGC_LAUNCHED("build_conv_plan_kernel");
  return FGC_OK;
}

size_t conv_mma_workspace(int64_t rows) {
  return ws_bytes(static_cast<size_t>(rows) * 256, 1) + ws_bytes(64, 4);
}

// fp16 hi|lo image of the first 64 channels of x (row stride ld floats) into img_ws
// (conv_mma_workspace(rows) bytes): image at offset 0, then 16 words of scalars
// ([0] max|x| bits, [1] 2^ex un-scale).
struct ImgWs {
  uint4* img;
  unsigned* scal;
};
static ImgWs img_ws_views(void* img_ws, int64_t rows) {
  Workspace ws(img_ws, conv_mma_workspace(rows));
  ImgWs v;
  v.img = reinterpret_cast<uint4*>(ws.take<char>(static_cast<size_t>(rows) * 256));
  v.scal = ws.take<unsigned>(16);
  return v;
}
const unsigned* conv_mma_image_maxbits(const void* img_ws, int64_t rows) {
  return img_ws_views(const_cast<void*>(img_ws), rows).scal;
}
int prep_image_blocks() { return num_sms() * 8; }
int prep_image_reset(void* img_ws, int64_t rows, cudaStream_t st) {
  FGC_CUDA(cudaMemsetAsync(img_ws_views(img_ws, rows).scal, 0, 16 * sizeof(unsigned), st));
  return FGC_OK;
}
// ---- tensor map of an image workspace for the TMA row gather
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    if (getenv("FGC_DISABLE_TMA") != nullptr) return nullptr;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
typedef void (*ConvMmaKernel)(const MmaParams);
static ConvMmaKernel pick_conv_mma_kernel(int K, int tma) {
  if (tma) return K <= 16 ? conv_mma_kernel<8, 64, 2, true> : (K <= 24 ? conv_mma_kernel<8, 64, 3, true> : conv_mma_kernel<8, 64, 4, true>);
  return K <= 16 ? conv_mma_kernel<8, 64, 2, false> : (K <= 24 ? conv_mma_kernel<8, 64, 3, false> : conv_mma_kernel<8, 64, 4, false>);
}

// which passes gather their rows by TMA: bit 0 forward, 1 target pass, 2 weight gradient, 3 source pass
// (FGC_TMA_MODES overrides the default mask)
static bool tma_pass_enabled(int bit) {
  static const int mask = getenv("FGC_TMA_MODES") != nullptr ? atoi(getenv("FGC_TMA_MODES")) : 15;
  return (mask >> bit) & 1;
}
// true when *tm describes img as [rows][128 nunits halves] with a 64 x 1 box and the 128-byte swizzle
static bool make_img_tmap(CUtensorMap* tm, const void* img, int64_t rows, int pass_bit, int nunits = 1) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (enc == nullptr || !tma_pass_enabled(pass_bit)) return false;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(128 * nunits), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(256 * nunits)};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(img), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// image of the HMMA-aggregation family: [rows][nunits][hi 64 | lo 64] halves; bit 4 of FGC_TMA_MODES -- off by default:
// measured 0.564 ms against 0.526 ms with cp.async for 562 k rows 64 -> 32 (the gathers bypass L1, which serves 60 % of
// the row reads of neighbouring facets, and cost an expect_tx + four issues per plane)
bool make_hm_img_tmap(void* tm, const void* img, int64_t rows, int nunits) {
  return make_img_tmap(static_cast<CUtensorMap*>(tm), img, rows, 4, nunits);
}

int launch_prep_image(const float* x, int ld, int64_t rows, void* img_ws, cudaStream_t st, const float* pinv,
                      int bias_mask, float* partB, bool have_absmax) {
  const ImgWs v = img_ws_views(img_ws, rows);
  const int ab = prep_image_blocks();
  if (have_absmax) {   // scal[0] already holds max|x| (prep_image_reset + the fused reduction of assign_logits)
    FGC_REQUIRE(partB == nullptr || (ld == 64 && pinv != nullptr), "prep_image: bias partials need 64-channel rows");
    prep_x_image_kernel<<<ab, 256, 0, st>>>(x, ld, rows, v.scal, v.img, reinterpret_cast<float*>(v.scal + 1), pinv,
                                            bias_mask, partB);
    FGC_LAUNCHED("prep_x_image_kernel");
    return FGC_OK;
  }
  FGC_CUDA(cudaMemsetAsync(v.scal, 0, 16 * sizeof(unsigned), st));
  FGC_REQUIRE(partB == nullptr || (ld == 64 && pinv != nullptr), "prep_image: bias partials need 64-channel rows");
  // a row stride above 64 (concat tails): scan the whole tensor, a superset bound is still a valid scale
  absmax2_kernel<<<ab, 256, 0, st>>>(x, rows * (ld / 4), v.scal);
  FGC_LAUNCHED("absmax_kernel");
  prep_x_image_kernel<<<ab, 256, 0, st>>>(x, ld, rows, v.scal, v.img, reinterpret_cast<float*>(v.scal + 1), pinv,
                                          bias_mask, partB);
  FGC_LAUNCHED("prep_x_image_kernel");
  return FGC_OK;
}

// img_ws: conv_mma_workspace(rows) bytes; wimg_ws: the weight image workspace of conv_fwd_tc
int launch_conv_mma(const ConvFwdParams& p, const float* W0, const void* plan, void* img_ws, void* wimg_ws,
                    cudaStream_t st, bool have_absmax) {
  using Cfg = MCfg<8, 64>;
  const PlanLayout L(p.rows, p.K, p.M);
  const char* pb = static_cast<const char*>(plan);
  int rc0 = launch_prep_image(p.x, p.Cin, p.rows, img_ws, st, nullptr, 0, nullptr, have_absmax);
  if (rc0) return rc0;
  const ImgWs iv = img_ws_views(img_ws, p.rows);
  uint4* img = iv.img;
  unsigned* scal = iv.scal;
  const size_t wbytes = static_cast<size_t>(p.M) * 2 * p.Cout * 128;
  float* wunscale = reinterpret_cast<float*>(static_cast<char*>(wimg_ws) + wbytes);
  int rc = launch_prep_w_image(W0, wimg_ws, p.M, p.Cout, st);
  if (rc) return rc;
  MmaParams mp{};
  mp.img = img, mp.xunscale = reinterpret_cast<const float*>(scal + 1), mp.uvx = p.uvx, mp.adj = p.adj;
  mp.use_tma = make_img_tmap(&mp.tmap, img, p.rows, 0) ? 1 : 0;
  mp.ppair = reinterpret_cast<const uint8_t*>(pb + L.off_pair), mp.prow = reinterpret_cast<const int32_t*>(pb + L.off_row);
  mp.pR = reinterpret_cast<const int32_t*>(pb + L.off_R), mp.pinv = reinterpret_cast<const float*>(pb + L.off_inv);
  mp.wimg = static_cast<const uint4*>(wimg_ws), mp.wunscale = wunscale, mp.b = p.b, mp.y = p.y;
  mp.rows = p.rows, mp.ntiles = L.ntiles, mp.N = p.N, mp.K = p.K, mp.ldy = p.Cout, mp.bias_mask = p.bias_mask;
  mp.act = p.act, mp.alpha = p.alpha;
  mp.uo_off = 0, mp.vl_off = p.M, mp.inv_src = nullptr, mp.mode = 0;
  static const bool trace = getenv("FGC_MMA_TRACE") != nullptr;
  mp.trace = trace ? 1 : 0;
  void (*kern)(const MmaParams) = pick_conv_mma_kernel(p.K, mp.use_tma);
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  int64_t grid = num_sms();
  if (grid > L.ntiles) grid = L.ntiles;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), kMmaThreads, Cfg::SMEM_BYTES, st>>>(mp);
  FGC_LAUNCHED("conv_mma_kernel");
  return FGC_OK;
}


// ------------------------------------------------------------------ target-centric pass on the same kernel
// radj[t][0..Kr): 1-indexed (within the batch element) source facets of the in-edges of target t, in
// rev_edge order, 0 padded -- the reversed adjacency in the forward layout, so that the same tile
// plan builder and the same kernel serve the target-centric backward pass.
__global__ void build_radj_kernel(const int32_t* __restrict__ rev_ptr, const int32_t* __restrict__ rev_edge,
                                  int64_t rows, int N, int K, int Kr, int32_t* __restrict__ radj) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * Kr;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / Kr;
    const int k = static_cast<int>(i % Kr);
    const int e0 = __ldg(rev_ptr + t), e1 = __ldg(rev_ptr + t + 1);
    int32_t v = 0;
    if (e0 + k < e1) {
      const int64_t src = __ldg(rev_edge + e0 + k) / K;   // global source row
      v = static_cast<int32_t>(src - (t / N) * N) + 1;
    }
    radj[i] = v;
  }
}

// d_uvx[t, M:2M] = sum over the in-edges e of t of da_edge[e, 0:M]   (list order: deterministic)
template <int M>
__global__ void __launch_bounds__(256)
tgt_dv_kernel(const int32_t* __restrict__ rev_ptr, const int32_t* __restrict__ rev_edge,
              const float* __restrict__ da_edge, float* __restrict__ d_uvx, int64_t rows) {
  constexpr int LP = M / 4;   // lanes per target, one float4 each
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t t = gid / LP;
  const int part = static_cast<int>(gid % LP);
  if (t >= rows) return;
  const int e0 = __ldg(rev_ptr + t), e1 = __ldg(rev_ptr + t + 1);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = e0; e < e1; ++e) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(da_edge + static_cast<int64_t>(__ldg(rev_edge + e)) * M) + part);
    a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
  }
  *reinterpret_cast<float4*>(d_uvx + t * (2 * M) + M + 4 * part) = a;
}

__global__ void max_degree_kernel(const int32_t* __restrict__ rev_ptr, int64_t rows, int32_t* __restrict__ out) {
  int m = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = max(m, __ldg(rev_ptr + i + 1) - __ldg(rev_ptr + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// largest in-degree of the reversed adjacency -> *out (device)
int launch_max_degree(const int32_t* rev_ptr, int64_t rows, int32_t* out, cudaStream_t st) {
  FGC_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
  max_degree_kernel<<<num_sms() * 4, 256, 0, st>>>(rev_ptr, rows, out);
  FGC_LAUNCHED("max_degree_kernel");
  return FGC_OK;
}

int launch_build_radj(const int32_t* rev_ptr, const int32_t* rev_edge, int B, int N, int K, int Kr, int32_t* radj,
                      cudaStream_t st) {
  const int64_t rows = static_cast<int64_t>(B) * N;
  int64_t blocks = (rows * Kr + 255) / 256;
  if (blocks > static_cast<int64_t>(num_sms()) * 32) blocks = static_cast<int64_t>(num_sms()) * 32;
  if (blocks < 1) blocks = 1;
  build_radj_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(rev_ptr, rev_edge, rows, N, K, Kr, radj);
  FGC_LAUNCHED("build_radj_kernel");
  return FGC_OK;
}

bool bwd_tgt_mma_supported(int Cin, int Cw, int Cout, int M, int Kr) {
  return Cw == 64 && Cout == 64 && M == 8 && Kr <= 32 && Cin % 4 == 0;
}

// gx[:, 0:Cw] = sum_m t[., m, :] W0[m] with t[j, m, :] = sum over in-edges (n -> j) of q[n->j, m] inv_cnt[n] gy[n, :]
// and d_uvx[:, M:2M]; radj / rplan: reversed adjacency in forward layout and its tile plan.
// wimg: transposed weight image (launch_prep_w_image_t); img_ws: conv_mma_workspace(rows) bytes.
int launch_bwd_tgt_mma(const float* gy, const float* uvx, const float* da_edge, const float* inv,
                       const int32_t* rev_ptr, const int32_t* rev_edge, const int32_t* radj, int Kr, const void* rplan,
                       float* gx, float* d_uvx, int64_t rows, int N, int Cin, int Cout, int M, const void* wimg,
                       void* img_ws, cudaStream_t st) {
  using Cfg = MCfg<8, 64>;
  const PlanLayout L(rows, Kr, M);
  const char* pb = static_cast<const char*>(rplan);
  // img_ws holds the gy image prepared by launch_prep_image(gy, Cout, rows, img_ws)
  (void)gy;
  (void)Cout;
  const ImgWs iv = img_ws_views(img_ws, rows);
  uint4* img = iv.img;
  unsigned* scal = iv.scal;
  const size_t wbytes = static_cast<size_t>(M) * 2 * 64 * 128;
  MmaParams mp{};
  mp.img = img, mp.xunscale = reinterpret_cast<const float*>(scal + 1), mp.uvx = uvx, mp.adj = radj;
  mp.use_tma = make_img_tmap(&mp.tmap, img, rows, 1) ? 1 : 0;
  mp.ppair = reinterpret_cast<const uint8_t*>(pb + L.off_pair), mp.prow = reinterpret_cast<const int32_t*>(pb + L.off_row);
  mp.pR = reinterpret_cast<const int32_t*>(pb + L.off_R), mp.pinv = reinterpret_cast<const float*>(pb + L.off_inv);
  mp.wimg = static_cast<const uint4*>(wimg);
  mp.wunscale = reinterpret_cast<const float*>(static_cast<const char*>(wimg) + wbytes);
  mp.b = nullptr, mp.y = gx, mp.rows = rows, mp.ntiles = L.ntiles, mp.N = N, mp.K = Kr, mp.ldy = Cin;
  mp.bias_mask = 0, mp.act = FGC_ACT_NONE, mp.alpha = 0.f;
  mp.uo_off = M, mp.vl_off = 0, mp.inv_src = inv, mp.mode = 1;
  static const bool trace = getenv("FGC_MMA_TRACE") != nullptr;
  mp.trace = trace ? 1 : 0;
  void (*kern)(const MmaParams) = pick_conv_mma_kernel(Kr, mp.use_tma);
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  int64_t grid = num_sms();
  if (grid > L.ntiles) grid = L.ntiles;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), kMmaThreads, Cfg::SMEM_BYTES, st>>>(mp);
  FGC_LAUNCHED("bwd_tgt_mma_kernel");
  const int64_t threads = rows * (8 / 4);
  tgt_dv_kernel<8><<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(rev_ptr, rev_edge, da_edge, d_uvx, rows);
  FGC_LAUNCHED("tgt_dv_kernel");
  return FGC_OK;
}


// Weight gradient on the dense-assignment kernel (mode 2): gW0[m] = sum_n gz_n (x) s[n,m,:], with s rebuilt by
// stage 1 exactly as in the forward.  partW: [grid][M][64][64] per-CTA partials (reduced in fixed order by the
// caller); returns the grid in *grid_out.  ximg_ws / gyimg_ws: images prepared by launch_prep_image.
int bwd_w_mma_grid(int64_t rows, int M) {
  const int64_t ntiles = (rows + 128 / M - 1) / (128 / M);
  int64_t g = num_sms();
  if (g > ntiles) g = ntiles;
  return static_cast<int>(g < 1 ? 1 : g);
}
int launch_bwd_w_mma(const float* uvx, const int32_t* adj, const void* plan, void* ximg_ws, void* gyimg_ws,
                     float* partW, int64_t rows, int N, int K, int M, cudaStream_t st) {
  using Cfg = MCfg<8, 64>;
  const PlanLayout L(rows, K, M);
  const char* pb = static_cast<const char*>(plan);
  const ImgWs xi = img_ws_views(ximg_ws, rows), gi = img_ws_views(gyimg_ws, rows);
  MmaParams mp{};
  mp.img = xi.img, mp.xunscale = reinterpret_cast<const float*>(xi.scal + 1), mp.uvx = uvx, mp.adj = adj;
  mp.use_tma = make_img_tmap(&mp.tmap, xi.img, rows, 2) ? 1 : 0;
  mp.ppair = reinterpret_cast<const uint8_t*>(pb + L.off_pair), mp.prow = reinterpret_cast<const int32_t*>(pb + L.off_row);
  mp.pR = reinterpret_cast<const int32_t*>(pb + L.off_R), mp.pinv = reinterpret_cast<const float*>(pb + L.off_inv);
  mp.wimg = nullptr, mp.wunscale = nullptr, mp.b = nullptr, mp.y = partW;
  mp.rows = rows, mp.ntiles = L.ntiles, mp.N = N, mp.K = K, mp.ldy = 64, mp.bias_mask = 0;
  mp.act = FGC_ACT_NONE, mp.alpha = 0.f;
  mp.uo_off = 0, mp.vl_off = M, mp.inv_src = nullptr, mp.mode = 2;
  mp.gimg = gi.img, mp.gunscale = reinterpret_cast<const float*>(gi.scal + 1);
  mp.trace = 0;
  void (*kern)(const MmaParams) = pick_conv_mma_kernel(K, mp.use_tma);
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  kern<<<static_cast<unsigned>(bwd_w_mma_grid(rows, M)), kMmaThreads, Cfg::SMEM_BYTES, st>>>(mp);
  FGC_LAUNCHED("bwd_w_mma_kernel");
  return FGC_OK;
}

bool bwd_src_mma_supported(int Cin, int Cw, int Cout, int M, int K) { return conv_mma_supported(Cin, Cw, Cout, M, K); }

const float* conv_plan_inv(const void* plan, int64_t rows, int K, int M) {
  return reinterpret_cast<const float*>(static_cast<const char*>(plan) + PlanLayout(rows, K, M).off_inv);
}

// ds / dq / da of the source-centric pass (conv_mma_src.cuh).  ximg_ws / gyimg_ws: images prepared by
// launch_prep_image; plan: the forward tile plan; wimg: transposed weight image.
int launch_bwd_src_mma(const float* gy, const float* uvx, const int32_t* adj, const void* plan, void* ximg_ws,
                       void* gyimg_ws, const void* wimg, float* da_edge, float* d_uvx, int64_t rows, int N, int K,
                       int M, cudaStream_t st) {
  using Cfg = SrcCfg<8>;
  const PlanLayout L(rows, K, M);
  const char* pb = static_cast<const char*>(plan);
  const ImgWs xi = img_ws_views(ximg_ws, rows), gi = img_ws_views(gyimg_ws, rows);
  const size_t wbytes = static_cast<size_t>(M) * 2 * 64 * 128;
  SrcParams sp{};
  sp.img = xi.img, sp.xunscale = reinterpret_cast<const float*>(xi.scal + 1);
  sp.use_tma = make_img_tmap(&sp.tmap, xi.img, rows, 3) ? 1 : 0;
  (void)gy;   // the stage-A operand is the gy image prepared by launch_prep_image(gy, ...) in gyimg_ws
  sp.gimg = gi.img, sp.gunscale = reinterpret_cast<const float*>(gi.scal + 1);
  sp.uvx = uvx;
  sp.ppair = reinterpret_cast<const uint8_t*>(pb + L.off_pair), sp.prow = reinterpret_cast<const int32_t*>(pb + L.off_row);
  sp.pR = reinterpret_cast<const int32_t*>(pb + L.off_R), sp.pinv = reinterpret_cast<const float*>(pb + L.off_inv);
  sp.wimg = static_cast<const uint4*>(wimg);
  sp.wunscale = reinterpret_cast<const float*>(static_cast<const char*>(wimg) + wbytes);
  sp.da_edge = da_edge, sp.d_uvx = d_uvx, sp.rows = rows, sp.ntiles = L.ntiles, sp.N = N, sp.K = K;
  void (*kern)(const SrcParams) = bwd_src_mma_kernel<8, 4, false>;
  if (sp.use_tma) kern = K <= 16 ? bwd_src_mma_kernel<8, 2, true> : (K <= 24 ? bwd_src_mma_kernel<8, 3, true> : bwd_src_mma_kernel<8, 4, true>);
  else if (K <= 16) kern = bwd_src_mma_kernel<8, 2, false>;
  else if (K <= 24) kern = bwd_src_mma_kernel<8, 3, false>;
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  int64_t grid = num_sms();
  if (grid > L.ntiles) grid = L.ntiles;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), kMmaThreads, Cfg::SMEM_BYTES, st>>>(sp);
  FGC_LAUNCHED("bwd_src_mma_kernel");
  src_dux_fix_kernel<8><<<static_cast<unsigned>((rows + 255) / 256), 256, 0, st>>>(sp.pR, adj, da_edge, d_uvx, rows, N,
                                                                                  K, L.TF);
  FGC_LAUNCHED("src_dux_fix_kernel");
  return FGC_OK;
}

int debug_mma_trace(int64_t* out, int n) {
  long long host[4 * 32 * 8];
  FGC_CUDA(cudaDeviceSynchronize());
  FGC_CUDA(cudaMemcpyFromSymbol(host, g_mma_trace, sizeof(host)));
  for (int i = 0; i < n && i < 4 * 32 * 8; ++i) out[i] = host[i];
  return FGC_OK;
}

}  // namespace fgc
