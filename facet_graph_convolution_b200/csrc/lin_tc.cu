// Fused regression head on tcgen05: y = lrelu(x W1 + b1) W2 + b2 with x[rows,32], W1[32,1024], W2[1024,3]
// (reference Code/model.py:936-941 through custom_lin :763-769; 72 KFLOP per facet, the largest single FLOP
// item of the network).  The 1024-wide hidden activation never leaves the SM.
//
// One persistent CTA per SM, tiles of 128 rows, 13 warps:
//   warps 8-11 producers: thread = row.  The row is scaled by its own power of two (the scale is a per-lane
//              factor of the accumulator, undone in the epilogue) and split into fp16 hi + fp16 residual,
//              written as the K-major 128B-swizzled A operand (K = 32: the first 64 bytes of every row).
//   warp  12   MMA issuer: per tile and chunk of 256 hidden units  D[128 x 256] = Ah.Bh + Al.Bh + Ah.Bl
//              (tcgen05.mma kind::f16, M = 128, N = 256, two k-steps each; fp32 accumulation in TMEM).
//              B = the resident fp16 hi/lo image of W1^T (512 hidden units per launch: 128 KB of shared memory).
//   warps 0-7  epilogue: two warps per TMEM lane quadrant, each takes half of a chunk's columns: tcgen05.ld,
//              h = lrelu(scale * d + b1) and the three output FMAs on PAIRS of hidden units (packed fp32x2
//              FMAs; parameters staged as pair records, two 16-byte shared loads per pair); the partner warp's
//              partial sums cross through shared memory, always added in the same order.  The two chunk
//              accumulators ping-pong, so the MMAs of the next tile run under the epilogue of this one.
//              (With four thread-per-row epilogue warps and scalar FMAs the epilogue was the whole kernel:
//              ~8 instructions per hidden unit and row, 1.24 ms per 2.25 M rows against a 0.19 ms MMA floor.)
// The 1024 hidden units are two launches of 512 (the weight image of all 1024 does not fit shared memory):
// the first writes y = b2 + partial, the second adds its partial.
#include "conv_common.cuh"
#include "conv_launch.cuh"
#include "tc_common.cuh"

namespace fgc {

namespace {

constexpr int kHK = 32;          // input channels (MMA K)
constexpr int kHH = 1024;        // hidden units
constexpr int kHPass = 512;      // hidden units per launch
constexpr int kHChunk = 256;     // hidden units per MMA (N)
constexpr int kHTile = 128;      // rows per tile
constexpr int kHThreads = 13 * 32;
constexpr int kHEpi = 8;         // epilogue warps
constexpr int kHProd0 = 8;       // first producer warp
constexpr int kHIssuer = 12;

struct HeadCfg {
  static constexpr int B_PLANE = kHPass * 128;             // [512 rows = hidden units][64 K halves], 32 used
  static constexpr int A_PLANE = kHTile * 128;
  static constexpr int OFF_B = 0;                          // hi | lo
  static constexpr int OFF_A = OFF_B + 2 * B_PLANE;        // 2 buffers x (hi | lo)
  static constexpr int OFF_PRM = OFF_A + 4 * A_PLANE;      // [256 pairs] {b1 i, b1 i+1, W2[i,0], W2[i+1,0]} {W2[.,1] pair, W2[.,2] pair}
  static constexpr int OFF_RS = OFF_PRM + kHPass * 16;     // [4][128] row un-scales
  static constexpr int OFF_EX = OFF_RS + 4 * kHTile * 4;   // [2][128] float4: partial sums of the second warp of a quadrant
  static constexpr int OFF_BAR = OFF_EX + 2 * kHTile * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
};
static_assert(HeadCfg::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

enum { H_A_FULL = 0, H_A_FREE = 2, H_D_FULL = 4, H_D_FREE = 6, H_NUM = 8 };

struct HeadParams {
  const float* x;
  const uint4* wimg;       // this pass: hi plane then lo plane, B_PLANE bytes each, already swizzled
  const float* wunscale;
  const float* b1;         // + pass offset
  const float* W2;         // [1024][3], + pass offset rows
  const float* b2;
  float* y;
  int64_t rows, ntiles;
  int accumulate;          // second pass: y += partial
  float alpha;
};

// FAST_ACT: 0 <= alpha <= 1, lrelu(h) = max(h, alpha h) (a runtime flag made the compiler issue BOTH activation forms
// predicated -- nine issue slots per pair of hidden units where three do)
template <bool FAST_ACT>
__global__ void __launch_bounds__(kHThreads, 1)
mlp_head_tc_kernel(const HeadParams p) {
  using Cfg = HeadCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + H_NUM);
  float4* prm = reinterpret_cast<float4*>(smem + Cfg::OFF_PRM);
  float* rs = reinterpret_cast<float*>(smem + Cfg::OFF_RS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars[H_A_FULL + i], 4), tc::mbar_init(&bars[H_A_FREE + i], 1);
      tc::mbar_init(&bars[H_D_FULL + i], 1), tc::mbar_init(&bars[H_D_FREE + i], kHEpi);
    }
    tc::mbar_fence_init();
  }
  if (warp == kHIssuer) tc::tmem_alloc(tmem_slot, 512);
  {
    uint4* dst = reinterpret_cast<uint4*>(smem + Cfg::OFF_B);
    for (int i = threadIdx.x; i < 2 * Cfg::B_PLANE / 16; i += kHThreads) dst[i] = __ldg(p.wimg + i);
    // rows of A beyond K = 32 are never read; zero them once so that no NaN pattern sits in the operand
    uint4* az = reinterpret_cast<uint4*>(smem + Cfg::OFF_A);
    for (int i = threadIdx.x; i < 4 * Cfg::A_PLANE / 16; i += kHThreads) az[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < kHPass / 2; i += kHThreads) {   // pair records of hidden units 2i, 2i+1
      const int h0 = 2 * i, h1 = 2 * i + 1;
      prm[2 * i] = make_float4(__ldg(p.b1 + h0), __ldg(p.b1 + h1), __ldg(p.W2 + 3 * h0), __ldg(p.W2 + 3 * h1));
      prm[2 * i + 1] = make_float4(__ldg(p.W2 + 3 * h0 + 1), __ldg(p.W2 + 3 * h1 + 1), __ldg(p.W2 + 3 * h0 + 2), __ldg(p.W2 + 3 * h1 + 2));
    }
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // launched with programmatic stream serialization: everything above (weight image, parameter records: older than the
  // stream's previous kernel) may run under that kernel's tail; the rows and the partial y are read below
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < kHEpi) {
    // =========================================================== epilogue
    const int q = warp & 3, half = warp >> 2;            // TMEM lane quadrant, column half of every chunk
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int row = q * 32 + lane;
    const float wun = __ldg(p.wunscale);
    const float b20 = __ldg(p.b2), b21 = __ldg(p.b2 + 1), b22 = __ldg(p.b2 + 2);
    const float2 al2 = make_float2(p.alpha, p.alpha);
    float4* ex = reinterpret_cast<float4*>(smem + Cfg::OFF_EX);
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;
      float2 sc2 = a0;
      // The 2 chunks x 4 column blocks of this warp, with the TMEM load of block k+1 in flight while block k is
      // computed (load -> wait -> compute, as first written, exposed the TMEM latency eight times per tile).
      constexpr int kBlk = kHChunk / 64;                                // blocks per chunk
      auto block_addr = [&](int k) -> uint32_t {
        return tmem + lane_base + (k / kBlk) * kHChunk + half * (kHChunk / 2) + (k % kBlk) * 32;
      };
      auto compute = [&](const uint32_t (&d)[32], int k) {
        const float4* pp = prm + ((k / kBlk) * kHChunk + half * (kHChunk / 2) + (k % kBlk) * 32);   // two float4 per pair
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 r0 = pp[2 * i], r1 = pp[2 * i + 1];
          float2 h = make_float2(r0.x, r0.y);                     // b1 pair
          tc::ffma2(h, make_float2(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1])), sc2);
          if (FAST_ACT) {
            float2 ah = make_float2(0.f, 0.f);
            tc::ffma2(ah, h, al2);
            h.x = fmaxf(h.x, ah.x), h.y = fmaxf(h.y, ah.y);
          } else {
            h.x = lrelu_f(h.x, p.alpha), h.y = lrelu_f(h.y, p.alpha);
          }
          tc::ffma2(a0, h, make_float2(r0.z, r0.w));
          tc::ffma2(a1, h, make_float2(r1.x, r1.y));
          tc::ffma2(a2, h, make_float2(r1.z, r1.w));
        }
      };
      // after the loads of a chunk's last block have landed the accumulator may be overwritten
      auto release_chunk = [&](int c) {
        tc::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[H_D_FREE + c]);
      };
      uint32_t dA[32], dB[32];
      tc::mbar_wait(&bars[H_D_FULL + 0], t & 1);
      tc::tc_fence_after_sync();
      {
        const float sc = rs[(t & 3) * kHTile + row] * wun;
        sc2 = make_float2(sc, sc);
      }
      tc::tmem_ld32(block_addr(0), dA);
#pragma unroll 1
      for (int k = 0; k < 2 * kBlk; k += 2) {
        tc::tc_wait_ld();                                   // block k is in dA
        tc::tmem_ld32(block_addr(k + 1), dB);               // same chunk (kBlk is even)
        compute(dA, k);
        tc::tc_wait_ld();                                   // block k + 1 is in dB
        if (k + 2 == kBlk) {                                // chunk 0 fully read: free it, then wait for chunk 1
          release_chunk(0);
          tc::mbar_wait(&bars[H_D_FULL + 1], t & 1);
          tc::tc_fence_after_sync();
        }
        if (k + 2 < 2 * kBlk) tc::tmem_ld32(block_addr(k + 2), dA);
        else release_chunk(1);
        compute(dB, k + 1);
      }
      const float s0 = a0.x + a0.y, s1 = a1.x + a1.y, s2 = a2.x + a2.y;
      // the second warp of the quadrant hands its partial sums to the first (fixed order of addition)
      if (half == 1) ex[(t & 1) * kHTile + row] = make_float4(s0, s1, s2, 0.f);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      if (half == 0) {
        const float4 o = ex[(t & 1) * kHTile + row];
        const int64_t r = tile * kHTile + row;
        if (r < p.rows) {
          float* yr = p.y + 3 * r;
          if (p.accumulate) yr[0] += s0 + o.x, yr[1] += s1 + o.y, yr[2] += s2 + o.z;
          else yr[0] = (s0 + o.x) + b20, yr[1] = (s1 + o.y) + b21, yr[2] = (s2 + o.z) + b22;
        }
      }
    }
  } else if (warp < kHIssuer) {
    // =========================================================== producers: thread = row
    const int row = (warp - kHProd0) * 32 + lane;
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int b = t & 1;
      const int64_t r = tile * kHTile + row;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = (r < p.rows) ? __ldg(reinterpret_cast<const float4*>(p.x + r * kHK) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
      int E = (__float_as_int(mx) >> 23) & 0xFF;
      E = min(max(E, 16), 240);
      const float sc = __int_as_float((267 - E) << 23);   // 2^(140-E): |x| sc in [2^13, 2^14)
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float f[4] = {v[i].x * sc, v[i].y * sc, v[i].z * sc, v[i].w * sc};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const __half2 hh = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(f[2 * k] - hf.x, f[2 * k + 1] - hf.y);
          hi[2 * i + k] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[2 * i + k] = *reinterpret_cast<const uint32_t*>(&ll);
        }
      }
      tc::mbar_wait(&bars[H_A_FREE + b], ((t >> 1) & 1) ^ 1);
      uint8_t* ah = smem + Cfg::OFF_A + b * 2 * Cfg::A_PLANE + (row >> 3) * 1024 + (row & 7) * 128;
      uint8_t* al = ah + Cfg::A_PLANE;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint4*>(ah + ((u ^ (row & 7)) << 4)) = make_uint4(hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        *reinterpret_cast<uint4*>(al + ((u ^ (row & 7)) << 4)) = make_uint4(lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
      }
      rs[(t & 3) * kHTile + row] = __int_as_float((E - 13) << 23);   // 2^(E-140)
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[H_A_FULL + b]);
    }
  } else {
    // =========================================================== MMA issuer
    constexpr uint32_t idesc = tc::idesc_f16(128, kHChunk);
    const uint32_t sb = tc::smem_u32(smem);
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int b = t & 1;
      tc::mbar_wait(&bars[H_A_FULL + b], (t >> 1) & 1);
      const uint32_t ah = sb + Cfg::OFF_A + b * 2 * Cfg::A_PLANE, al = ah + Cfg::A_PLANE;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        tc::mbar_wait(&bars[H_D_FREE + c], (t & 1) ^ 1);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t bh = sb + Cfg::OFF_B + c * kHChunk * 128, bl = bh + Cfg::B_PLANE;
          const uint32_t d = tmem + c * kHChunk;
#pragma unroll
          for (int ks = 0; ks < kHK / 16; ++ks) {
            const uint64_t dah = tc::smem_desc_k_sw128(ah + ks * 32), dal = tc::smem_desc_k_sw128(al + ks * 32);
            const uint64_t dbh = tc::smem_desc_k_sw128(bh + ks * 32), dbl = tc::smem_desc_k_sw128(bl + ks * 32);
            tc::mma_f16_ss(d, dah, dbh, idesc, ks ? 1u : 0u);
            tc::mma_f16_ss(d, dal, dbh, idesc, 1u);
            tc::mma_f16_ss(d, dah, dbl, idesc, 1u);
          }
          tc::tc_commit(&bars[H_D_FULL + c]);
          if (c == 1) tc::tc_commit(&bars[H_A_FREE + b]);
        }
        __syncwarp();
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == kHIssuer) tc::tmem_dealloc(tmem, 512);
}

// Two-tile variant (default): the epilogue is bound by its parameter loads -- warp-broadcast LDS.128 still cost four LSU
// wavefronts each, 8 192 wavefront-cycles per 128-row tile and launch -- so the CTA keeps TWO tiles in flight (the two A
// buffers), the hidden units go through TMEM in chunks of 128 columns per tile, and every parameter load serves a row of
// each tile.  Same roles, barriers and arithmetic (per row: the same fma chain in the same order) as the kernel above.
template <bool FAST_ACT>
__global__ void __launch_bounds__(kHThreads, 1)
mlp_head_tc2_kernel(const HeadParams p) {
  using Cfg = HeadCfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + H_NUM);
  float4* prm = reinterpret_cast<float4*>(smem + Cfg::OFF_PRM);
  float* rs = reinterpret_cast<float*>(smem + Cfg::OFF_RS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars[H_A_FULL + i], 4), tc::mbar_init(&bars[H_A_FREE + i], 1);
      tc::mbar_init(&bars[H_D_FULL + i], 1), tc::mbar_init(&bars[H_D_FREE + i], kHEpi);
    }
    tc::mbar_fence_init();
  }
  if (warp == kHIssuer) tc::tmem_alloc(tmem_slot, 512);
  {
    uint4* dst = reinterpret_cast<uint4*>(smem + Cfg::OFF_B);
    for (int i = threadIdx.x; i < 2 * Cfg::B_PLANE / 16; i += kHThreads) dst[i] = __ldg(p.wimg + i);
    // rows of A beyond K = 32 are never read; zero them once so that no NaN pattern sits in the operand
    uint4* az = reinterpret_cast<uint4*>(smem + Cfg::OFF_A);
    for (int i = threadIdx.x; i < 4 * Cfg::A_PLANE / 16; i += kHThreads) az[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < kHPass / 2; i += kHThreads) {   // pair records of hidden units 2i, 2i+1
      const int h0 = 2 * i, h1 = 2 * i + 1;
      prm[2 * i] = make_float4(__ldg(p.b1 + h0), __ldg(p.b1 + h1), __ldg(p.W2 + 3 * h0), __ldg(p.W2 + 3 * h1));
      prm[2 * i + 1] = make_float4(__ldg(p.W2 + 3 * h0 + 1), __ldg(p.W2 + 3 * h1 + 1), __ldg(p.W2 + 3 * h0 + 2), __ldg(p.W2 + 3 * h1 + 2));
    }
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // launched with programmatic stream serialization: everything above (weight image, parameter records: older than the
  // stream's previous kernel) may run under that kernel's tail; the rows and the partial y are read below
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < kHEpi) {
    // =========================================================== epilogue
    const int q = warp & 3, half = warp >> 2;            // TMEM lane quadrant, column half of every chunk
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int row = q * 32 + lane;
    const float wun = __ldg(p.wunscale);
    const float b20 = __ldg(p.b2), b21 = __ldg(p.b2 + 1), b22 = __ldg(p.b2 + 2);
    const float2 al2 = make_float2(p.alpha, p.alpha);
    float4* ex = reinterpret_cast<float4*>(smem + Cfg::OFF_EX);
    const int my_tiles = static_cast<int>((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    constexpr int kC2 = 128;                                  // hidden units per chunk and tile
#pragma unroll 1
    for (int P = 0; 2 * P < my_tiles; ++P) {
      const int tA = 2 * P, tB = 2 * P + 1;
      const bool has_b = tB < my_tiles;
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, b0 = a0, b1v = a0, b2v = a0;
      float2 scA2 = make_float2(0.f, 0.f), scB2 = scA2;
#pragma unroll 1
      for (int c4 = 0; c4 < kHPass / kC2; ++c4) {
        const int s = c4 & 1, n = 2 * P + (c4 >> 1);
        tc::mbar_wait(&bars[H_D_FULL + s], n & 1);
        tc::tc_fence_after_sync();
        if (c4 == 0) {     // the scales were written before the producers' arrive that the MMAs of this pair waited for
          const float sa = rs[(tA & 3) * kHTile + row] * wun, sb_ = rs[(tB & 3) * kHTile + row] * wun;
          scA2 = make_float2(sa, sa), scB2 = make_float2(sb_, sb_);
        }
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const int col0 = half * (kC2 / 2) + j * 32;             // first hidden unit of this block inside the chunk
          uint32_t dA[32], dB[32];
          tc::tmem_ld32(tmem + lane_base + s * 2 * kC2 + col0, dA);
          tc::tmem_ld32(tmem + lane_base + s * 2 * kC2 + kC2 + col0, dB);
          tc::tc_wait_ld();
          if (j == 1) {
            tc::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars[H_D_FREE + s]);
          }
          const float4* pp = prm + (c4 * kC2 + col0);             // two float4 per pair of hidden units
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 r0 = pp[2 * i], r1 = pp[2 * i + 1];
            float2 hA = make_float2(r0.x, r0.y), hB = hA;           // b1 pair
            tc::ffma2(hA, make_float2(__uint_as_float(dA[2 * i]), __uint_as_float(dA[2 * i + 1])), scA2);
            tc::ffma2(hB, make_float2(__uint_as_float(dB[2 * i]), __uint_as_float(dB[2 * i + 1])), scB2);
            if (FAST_ACT) {
              float2 tA2 = make_float2(0.f, 0.f), tB2 = tA2;
              tc::ffma2(tA2, hA, al2);
              tc::ffma2(tB2, hB, al2);
              hA.x = fmaxf(hA.x, tA2.x), hA.y = fmaxf(hA.y, tA2.y);
              hB.x = fmaxf(hB.x, tB2.x), hB.y = fmaxf(hB.y, tB2.y);
            } else {
              hA.x = lrelu_f(hA.x, p.alpha), hA.y = lrelu_f(hA.y, p.alpha);
              hB.x = lrelu_f(hB.x, p.alpha), hB.y = lrelu_f(hB.y, p.alpha);
            }
            const float2 w0 = make_float2(r0.z, r0.w), w1 = make_float2(r1.x, r1.y), w2 = make_float2(r1.z, r1.w);
            tc::ffma2(a0, hA, w0), tc::ffma2(a1, hA, w1), tc::ffma2(a2, hA, w2);
            tc::ffma2(b0, hB, w0), tc::ffma2(b1v, hB, w1), tc::ffma2(b2v, hB, w2);
          }
        }
      }
      // the second warp of the quadrant hands its partial sums to the first (fixed order of addition)
      const float sA0 = a0.x + a0.y, sA1 = a1.x + a1.y, sA2 = a2.x + a2.y;
      const float sB0 = b0.x + b0.y, sB1 = b1v.x + b1v.y, sB2 = b2v.x + b2v.y;
      if (half == 1) {
        ex[row] = make_float4(sA0, sA1, sA2, 0.f);
        ex[kHTile + row] = make_float4(sB0, sB1, sB2, 0.f);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      if (half == 0) {
        const float4 oA = ex[row], oB = ex[kHTile + row];
        const int64_t tileA = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(tA) * gridDim.x;
        const int64_t rA = tileA * kHTile + row, rB = (tileA + gridDim.x) * kHTile + row;
        if (rA < p.rows) {
          float* yr = p.y + 3 * rA;
          if (p.accumulate) yr[0] += sA0 + oA.x, yr[1] += sA1 + oA.y, yr[2] += sA2 + oA.z;
          else yr[0] = (sA0 + oA.x) + b20, yr[1] = (sA1 + oA.y) + b21, yr[2] = (sA2 + oA.z) + b22;
        }
        if (has_b && rB < p.rows) {
          float* yr = p.y + 3 * rB;
          if (p.accumulate) yr[0] += sB0 + oB.x, yr[1] += sB1 + oB.y, yr[2] += sB2 + oB.z;
          else yr[0] = (sB0 + oB.x) + b20, yr[1] = (sB1 + oB.y) + b21, yr[2] = (sB2 + oB.z) + b22;
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // ex[] is reused by the next pair
    }
  } else if (warp < kHIssuer) {
    // =========================================================== producers: thread = row
    const int row = (warp - kHProd0) * 32 + lane;
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++t) {
      const int b = t & 1;
      const int64_t r = tile * kHTile + row;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = (r < p.rows) ? __ldg(reinterpret_cast<const float4*>(p.x + r * kHK) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
      int E = (__float_as_int(mx) >> 23) & 0xFF;
      E = min(max(E, 16), 240);
      const float sc = __int_as_float((267 - E) << 23);   // 2^(140-E): |x| sc in [2^13, 2^14)
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float f[4] = {v[i].x * sc, v[i].y * sc, v[i].z * sc, v[i].w * sc};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const __half2 hh = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(f[2 * k] - hf.x, f[2 * k + 1] - hf.y);
          hi[2 * i + k] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[2 * i + k] = *reinterpret_cast<const uint32_t*>(&ll);
        }
      }
      tc::mbar_wait(&bars[H_A_FREE + b], ((t >> 1) & 1) ^ 1);
      uint8_t* ah = smem + Cfg::OFF_A + b * 2 * Cfg::A_PLANE + (row >> 3) * 1024 + (row & 7) * 128;
      uint8_t* al = ah + Cfg::A_PLANE;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint4*>(ah + ((u ^ (row & 7)) << 4)) = make_uint4(hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
        *reinterpret_cast<uint4*>(al + ((u ^ (row & 7)) << 4)) = make_uint4(lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
      }
      rs[(t & 3) * kHTile + row] = __int_as_float((E - 13) << 23);   // 2^(E-140)
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[H_A_FULL + b]);
    }
  } else {
    // =========================================================== MMA issuer
    constexpr int kC2 = 128;
    constexpr uint32_t idesc = tc::idesc_f16(128, kC2);
    const uint32_t sb = tc::smem_u32(smem);
    const int my_tiles = static_cast<int>((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
#pragma unroll 1
    for (int P = 0; 2 * P < my_tiles; ++P) {
      const bool has_b = 2 * P + 1 < my_tiles;
      tc::mbar_wait(&bars[H_A_FULL + 0], P & 1);
      if (has_b) tc::mbar_wait(&bars[H_A_FULL + 1], P & 1);
#pragma unroll 1
      for (int c4 = 0; c4 < kHPass / kC2; ++c4) {
        const int s = c4 & 1, n = 2 * P + (c4 >> 1);
        tc::mbar_wait(&bars[H_D_FREE + s], (n & 1) ^ 1);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t bh = sb + Cfg::OFF_B + c4 * kC2 * 128, bl = bh + Cfg::B_PLANE;
#pragma unroll 1
          for (int tb = 0; tb < (has_b ? 2 : 1); ++tb) {
            const uint32_t ah = sb + Cfg::OFF_A + tb * 2 * Cfg::A_PLANE, al = ah + Cfg::A_PLANE;
            const uint32_t d = tmem + s * 2 * kC2 + tb * kC2;
#pragma unroll
            for (int ks = 0; ks < kHK / 16; ++ks) {
              const uint64_t dah = tc::smem_desc_k_sw128(ah + ks * 32), dal = tc::smem_desc_k_sw128(al + ks * 32);
              const uint64_t dbh = tc::smem_desc_k_sw128(bh + ks * 32), dbl = tc::smem_desc_k_sw128(bl + ks * 32);
              tc::mma_f16_ss(d, dah, dbh, idesc, ks ? 1u : 0u);
              tc::mma_f16_ss(d, dal, dbh, idesc, 1u);
              tc::mma_f16_ss(d, dah, dbl, idesc, 1u);
            }
          }
          tc::tc_commit(&bars[H_D_FULL + s]);
          if (c4 == kHPass / kC2 - 1) {
            tc::tc_commit(&bars[H_A_FREE + 0]);
            if (has_b) tc::tc_commit(&bars[H_A_FREE + 1]);
          }
        }
        __syncwarp();
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == kHIssuer) tc::tmem_dealloc(tmem, 512);
}

// image of W1^T for both passes: pass p, plane (hi, lo): [512 hidden units][128 B], K-major, 128B swizzle;
// values W1[k][512 p + n] * 2^(140 - E) with E the exponent of max|W1| (hi < 2^14, residual in the normal range)
__global__ void __launch_bounds__(1024)
prep_head_w_kernel(const float* __restrict__ W1, uint16_t* __restrict__ img, float* __restrict__ wunscale) {
  __shared__ float red[32];
  float mx = 0.f;
  for (int e = threadIdx.x; e < kHK * kHH; e += blockDim.x) mx = fmaxf(mx, fabsf(W1[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  int E = (__float_as_int(mx) >> 23) & 0xFF;
  E = min(max(E, 16), 240);
  const float sc = __int_as_float((267 - E) << 23);
  if (threadIdx.x == 0 && blockIdx.x == 0) wunscale[0] = __int_as_float((E - 13) << 23);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kHK * kHH; e += gridDim.x * blockDim.x) {
    const int k = e / kHH, hcol = e % kHH;           // coalesced read of W1[k][.]
    const int pass = hcol / kHPass, n = hcol % kHPass;
    const float v = W1[e] * sc;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const size_t plane_halves = static_cast<size_t>(kHPass) * 64;
    const size_t off = static_cast<size_t>(pass) * 2 * plane_halves + (n >> 3) * 512 + (n & 7) * 64 +
                       (((k >> 3) ^ (n & 7)) << 3) + (k & 7);
    img[off] = __half_as_ushort(h);
    img[off + plane_halves] = __half_as_ushort(l);
  }
}

}  // namespace

bool mlp_head_tc_supported(int64_t rows, int Cin, int H, int Cout) {
  static const bool disabled = getenv("FGC_DISABLE_TC") != nullptr;
  return !disabled && Cin == kHK && H == kHH && Cout == 3 && rows >= 2048;
}

size_t mlp_head_tc_workspace() { return 2 * 2 * static_cast<size_t>(HeadCfg::B_PLANE) + 1024; }

// The fp16 image of W1 depends only on the parameters: inference prepares it once (fgc_net_prepare) into a buffer of
// mlp_head_tc_workspace() bytes and passes it as `prepared` to launch_mlp_head_tc.
int launch_mlp_head_tc_prepare(const float* W1, void* prepared, size_t prepared_bytes, cudaStream_t st) {
  Workspace ws(prepared, prepared_bytes);
  char* img = ws.take<char>(2 * 2 * static_cast<size_t>(HeadCfg::B_PLANE));
  float* wunscale = ws.take<float>(4);
  FGC_REQUIRE(ws.ok(), "mlp_head: prepared buffer too small (%zu bytes given, %zu needed)", prepared_bytes, mlp_head_tc_workspace());
  prep_head_w_kernel<<<16, 1024, 0, st>>>(W1, reinterpret_cast<uint16_t*>(img), wunscale);
  FGC_LAUNCHED("prep_head_w_kernel");
  return FGC_OK;
}

int launch_mlp_head_tc(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y,
                       int64_t rows, float alpha, void* workspace, size_t workspace_bytes, cudaStream_t st,
                       const void* prepared) {
  const bool have = prepared != nullptr;
  Workspace ws(have ? const_cast<void*>(prepared) : workspace, have ? mlp_head_tc_workspace() : workspace_bytes);
  char* img = ws.take<char>(2 * 2 * static_cast<size_t>(HeadCfg::B_PLANE));
  float* wunscale = ws.take<float>(4);
  FGC_REQUIRE(ws.ok(), "mlp_head: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              mlp_head_tc_workspace());
  if (!have) {
    prep_head_w_kernel<<<16, 1024, 0, st>>>(W1, reinterpret_cast<uint16_t*>(img), wunscale);
    FGC_LAUNCHED("prep_head_w_kernel");
  }
  static const bool one_tile = getenv("FGC_HEAD_V1") != nullptr;   // the one-tile-per-CTA kernel, for comparisons
  const bool fast = alpha >= 0.f && alpha <= 1.f;
  auto kern = one_tile ? (fast ? mlp_head_tc_kernel<true> : mlp_head_tc_kernel<false>)
                       : (fast ? mlp_head_tc2_kernel<true> : mlp_head_tc2_kernel<false>);
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HeadCfg::SMEM_BYTES));
  HeadParams p{};
  p.x = x, p.wunscale = wunscale, p.b2 = b2, p.y = y, p.rows = rows, p.alpha = alpha;
  p.ntiles = (rows + kHTile - 1) / kHTile;
  int64_t grid = num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  for (int pass = 0; pass < kHH / kHPass; ++pass) {
    p.wimg = reinterpret_cast<const uint4*>(img + static_cast<size_t>(pass) * 2 * HeadCfg::B_PLANE);
    p.b1 = b1 + pass * kHPass, p.W2 = W2 + static_cast<size_t>(pass) * kHPass * 3, p.accumulate = pass > 0;
    {
      static const bool pdl = getenv("FGC_DISABLE_PDL") == nullptr;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(static_cast<unsigned>(grid)), cfg.blockDim = dim3(kHThreads), cfg.dynamicSmemBytes = HeadCfg::SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
      cfg.attrs = attr, cfg.numAttrs = 1;
      FGC_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    }
    FGC_LAUNCHED("mlp_head_tc_kernel");
  }
  return FGC_OK;
}

}  // namespace fgc
