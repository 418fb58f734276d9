// Facet-graph convolution forward for the network's dense layers (reference Code/model.py:427-504, called at
// :870-932 with M = 9; also the M = 8 benchmark layer): per-facet aggregation on the warp-level tensor path
// (mma.sync m16n8k16), contraction with W on tcgen05.  No tile plan, no per-tile row dedup: works on any adjacency.
//
//   stage 1 (16 aggregator warps, one facet at a time per warp)
//       S_n[m, c] = sum_k q[n,k,m] x_{j_k}[c]            A = q  (16 x 16: rows m, columns = neighbour slots)
//                                                         B = the 16 gathered rows of the fp16 hi|lo image of x
//       q is computed in the A-fragment layout (lane (g,t) owns row m = g and slots 2t,2t+1,2t+8,2t+9; the
//       softmax normaliser is a 3-step shuffle sum over the 8 g-lanes, the four slots side by side).  The gathered
//       rows go global -> shared with 16-byte cp.async (eight lanes per 128-byte line) into a per-warp 4 KB stage,
//       128B-swizzled, and come back as B fragments with ldmatrix.x4.trans; the copies of facet n+1 fly under the
//       MMAs, the drain and the next softmax.  q_hi.x_hi + q_lo.x_hi + q_hi.x_lo in fp32 accumulators = fp32-class
//       precision.  M = 9: rows 8..15 of A all carry q[.,8] (every lane then owns a copy of S[8,.] and converts 2 of
//       its 64 values); M = 8: rows 8..15 carry q_lo, so two MMAs per n-block instead of three.
//   drain   S -> fp16 hi (11 bits, exact) + fp16 residual -> shared memory, directly in the K-major 128B-swizzled
//       layout of the B operand of stage 2 ([hi rows of the tile's 32 facets | lo rows], one 64-element atom per
//       (weight pair, channel half) so that the 16-byte stores of a quarter warp hit 8 different bank groups).
//       Hand-over: one mbarrier arrive per warp and tile, no fence in the aggregators.
//   stage 2 (epilogue warp 0, one elected thread, after the generic->async proxy fence)
//       Y^T[(h,o), f] = sum_{m,c} [Wh;Wl][(h,o),(m,c)] [Sh|Sl][(m,c), f]
//       tcgen05.mma M = 128, N = 64, K = 64 M, A = the weight image resident in TMEM for the CTA's lifetime.
//   epilogue (4 warps on the TMEM lane quadrants)  y = act(inv_cnt * scale * (Wh.Sh + Wh.Sl + 2^-11 Wl.Sh) + flag * b),
//       optional max over groups of 4 rows (custom_binary_tree_pooling, model.py:863,875) and max|y| for the
//       image scale of the next layer.
// One launch covers 64 aggregation channels x up to 64 outputs; wider layers are sums of launches over channel
// blocks (bias in the first, activation / pooling in the last), as in conv_fwd_tc.cu.
#include "conv_launch.cuh"
#include <stdlib.h>

#include "tc_common.cuh"
#include <type_traits>

namespace fgc {

namespace {

constexpr int kHT = 32;                              // facets per tile
constexpr int kHAgg = 16;                            // aggregator warps
constexpr int kHFpw = kHT / kHAgg;                   // facets per warp and tile

template <int M>
struct HmCfg {
  static_assert(M == 8 || M == 9, "M = 8 or 9");
  static constexpr int NATOM = M;                    // K atoms of 64 elements: 8 for the (m < 8, unit) pairs, +1 for m = 8
  static constexpr int W_COLS = NATOM * 32;          // TMEM columns of the weight operand (2 halves per column)
  static constexpr int D_COL = 320;                  // two accumulators of ND columns
  static constexpr int ND = 2 * kHT;                 // stage-2 N: [Sh | Sl]
  static constexpr int ATOM_BYTES = ND * 128;
  static constexpr int B3_BUF = NATOM * ATOM_BYTES;
  static constexpr int OFF_ROW = 2 * B3_BUF;         // inv_cnt of the tile's facets, 4 tiles deep
  static constexpr int OFF_BAR = OFF_ROW + 4 * kHT * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static_assert(W_COLS <= D_COL && D_COL + 2 * ND <= 512, "TMEM overflow");
};


struct HmParams {
  alignas(64) unsigned char tmap[128];   // CUtensorMap of the whole image (second-generation kernel, TMA row gather)
  int unit_col;            // first column (in halves) of this launch's 64-channel unit inside an image row
  const uint4* img;        // fp16 image of this launch's 64-channel unit: row r at img + r * img_ld, [0..8) hi, [8..16) lo
  int img_ld;
  const float* xunscale;   // [B]: 2^ex of the image rows of every batch element
  const float* lg;         // [2][rows_img + 1][16]: pairs {vl'[m], vl'[8]} of every row, then pairs {uo'[m], uo'[8]} (prep_rows)
  const unsigned* flag;    // != 0: some row's logits spread over > 60 binary orders -> re-centre the softmax here
  const int32_t* adj;
  const uint32_t* wt;      // [128 TMEM lanes][W_COLS]
  const float* wunscale;
  const float* b;
  float* y;
  int ldy;
  float* ypool;            // optional: max over groups of 4 consecutive rows of the final output
  int ldp;
  unsigned* ymax;          // optional, [B]: atomicMax of the bits of |final output| per batch element (needs N % 16 == 0)
  int64_t rows, ntiles;
  int N, K, upshift;
  int bias_mask, act;
  float alpha;
  int add_bias, accumulate, apply_act;
  int cout;                // outputs of this launch (16 .. 64, multiple of 16)
  int single;              // B == 1: neighbour ids index the rows directly
  int trace;               // FGC_HM_TRACE: clock64 stamps of CTA 0 (tests/micro/hm_trace.py)
  int zrow;                // index of an all-zero row of the image and of uvx (padding / out-of-range slots read it)
};

__device__ __forceinline__ void hmma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fp16 pair of (a, b) and of the residuals
__device__ __forceinline__ void split_rn(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// hi truncated to 11 significant bits (exactly representable), lo = fp16 of the exact residual
__device__ __forceinline__ void split_trunc(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float h0 = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(h0, h1);
  const __half2 l = __floats2half2_rn(a - h0, b - h1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t w4(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

template <bool ZERO_C>
__device__ __forceinline__ void hm_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
  if (ZERO_C) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
  } else {
    hmma16816(d, a0, a1, a2, a3, b0, b1);
  }
}

// epilogue warp `warp` (TMEM lane quadrant `warp`): Y (TMEM) -> global, every tile of this CTA
// optional pipeline trace (FGC_HM_TRACE=1): clock64 stamps of CTA 0, first 64 tiles, 16 events per tile
__device__ long long g_hm_trace[64 * 16];
#define HM_TR(t, ev)                                                                  \
  do {                                                                                \
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (t) < 64) g_hm_trace[(t) * 16 + (ev)] = clock64(); \
  } while (0)

struct HmNoHook {
  __device__ __forceinline__ void operator()(int) const {}
};
// `before_tile(it)` runs ahead of the epilogue of the CTA's tile number it (the second-generation kernel issues stage 2 there)
template <class Cfg, int BAR_FULL, int BAR_FREE, class Hook = HmNoHook>
__device__ __forceinline__ void hm_epilogue(const HmParams& p, uint64_t* bars, const float* rowinv, uint32_t tmem, int warp,
                                            int lane, Hook before_tile = Hook()) {
  const int q = warp;
  const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
  // lane 16h + i of quadrant q: h = 0 holds Wh.(Sh|Sl) of channel o = 16q + i, h = 1 holds Wl.Sh.
  // After one shuffle round lane (h, i) owns facets 16h .. 16h+15 of channel o.
  const int hh = lane >> 4, o = q * 16 + (lane & 15);
  const float bo = p.add_bias ? __ldg(p.b + o) : 0.f;
  const float wun = __ldg(p.wunscale);
  const bool unmasked = !p.bias_mask;
  const bool aligned = (p.N & 15) == 0;   // a lane's 16 rows lie in one batch element
  const bool do_act = p.apply_act && p.act == FGC_ACT_LRELU;
  int it = 0;

  for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
    before_tile(it);
    const int buf = it & 1;
    const int64_t r0 = tile * kHT + 16 * hh;
    int64_t left = p.rows - r0;
    const int nv = left >= 16 ? 16 : (left < 0 ? 0 : static_cast<int>(left));
    // everything that does not depend on the accumulator is fetched before the wait: the scale(s) of the rows' batch
    // element(s) and, for accumulating launches, the sixteen partial sums (all loads in flight at once)
    float* yp = p.y + r0 * p.ldy + o;
    float yold[16];
    if (p.accumulate) {
#pragma unroll
      for (int j = 0; j < 16; ++j) yold[j] = (j < nv) ? __ldcg(yp + j * p.ldy) : 0.f;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) yold[j] = 0.f;
    }
    const int be0 = nv > 0 ? static_cast<int>(static_cast<int>(r0) / p.N) : 0;
    const float sca = __ldg(p.xunscale + be0) * wun;
    // rows of the next batch element start at j = cross (N >= 16: at most one boundary inside a lane's 16 rows)
    int cross = 16;
    float scb = sca;
    if (!aligned && nv > 0) {
      cross = static_cast<int>(static_cast<int64_t>(be0 + 1) * p.N - r0);
      if (cross < nv) scb = __ldg(p.xunscale + be0 + 1) * wun;
    }
    tc::mbar_wait_lazy(&bars[BAR_FULL + buf], (it >> 1) & 1);
    if (warp < 2) HM_TR(it, 4 + 2 * warp);
    tc::tc_fence_after_sync();
    uint32_t d0[32], d1[32];
    tc::tmem_ld32(tmem + lane_base + Cfg::D_COL + buf * Cfg::ND, d0);        // . Sh of facets 0..31
    tc::tmem_ld32(tmem + lane_base + Cfg::D_COL + buf * Cfg::ND + 32, d1);   // . Sl
    float inv[16];
    {
      const float4* ri = reinterpret_cast<const float4*>(rowinv + (it & 3) * kHT + 16 * hh);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = ri[j];
        inv[4 * j] = v.x, inv[4 * j + 1] = v.y, inv[4 * j + 2] = v.z, inv[4 * j + 3] = v.w;
      }
    }
    tc::tc_wait_ld();
    tc::tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&bars[BAR_FREE + buf]);
    float keep[16], send[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a_lo = __uint_as_float(d0[j]) + __uint_as_float(d1[j]);             // facets 0..15  (hi lanes)
      const float a_up = __uint_as_float(d0[16 + j]) + __uint_as_float(d1[16 + j]);   // facets 16..31 (hi lanes)
      const float l_lo = __uint_as_float(d0[j]) * (1.f / 2048.f);                     // facets 0..15  (lo lanes)
      const float l_up = __uint_as_float(d0[16 + j]) * (1.f / 2048.f);                // facets 16..31 (lo lanes)
      keep[j] = hh ? l_up : a_lo;
      send[j] = hh ? l_lo : a_up;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) send[j] = __shfl_xor_sync(0xffffffffu, send[j], 16);
    float yv[16];
    if (nv == 16 && cross >= 16 && p.N >= 16) {
      // the common case -- sixteen rows of one batch element -- without per-row branches (the branchy general loop
      // below was the larger part of an epilogue of ~6 000 cycles per tile, a third of it instruction-cache misses)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float accv = hh ? (send[j] + keep[j]) : (keep[j] + send[j]);   // (Wh.Sh + Wh.Sl) + Wl.Sh / 2048
        const float fl = (inv[j] > 0.f || unmasked) ? bo : 0.f;
        float v = fmaf(inv[j] * sca, accv, fl);
        v += yold[j];
        if (do_act) v = lrelu_f(v, p.alpha);
        yv[j] = v;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) yp[j * p.ldy] = yv[j];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float accv = hh ? (send[j] + keep[j]) : (keep[j] + send[j]);
        const float fl = (inv[j] > 0.f || unmasked) ? bo : 0.f;
        float sc0 = j < cross ? sca : scb;
        if (p.N < 16 && j > 0 && j < nv) sc0 = __ldg(p.xunscale + (r0 + j) / p.N) * wun;   // tiny elements: several boundaries
        float v = fmaf(inv[j] * sc0, accv, fl);
        v += yold[j];
        if (do_act) v = lrelu_f(v, p.alpha);
        yv[j] = v;
        if (j < nv) yp[j * p.ldy] = v;
      }
    }
    if (p.ypool != nullptr) {
      float* pp = p.ypool + (r0 >> 2) * p.ldp + o;
#pragma unroll
      for (int a = 0; a < 4; ++a)
        if (4 * a < nv) pp[a * p.ldp] = fmaxf(fmaxf(yv[4 * a], yv[4 * a + 1]), fmaxf(yv[4 * a + 2], yv[4 * a + 3]));
    }
    if (p.ymax != nullptr) {
      // max|y| per batch element: reduced over the 16 lanes that share the element, one atomic per half warp and
      // tile (max is order-independent: the atomics keep the result deterministic; thousands of same-address
      // atomics per launch serialise in L2 -- measured +180 us per 140 k rows when every lane issued its own)
      float m = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nv) m = fmaxf(m, fabsf(yv[j]));
#pragma unroll
      for (int s = 8; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
      if ((lane & 15) == 0 && nv > 0) atomicMax(p.ymax + be0, __float_as_uint(m));
    }
    if (warp < 2) HM_TR(it, 5 + 2 * warp);
  }
}

// =====================================================================================================================
// The kernel.  Second generation: the first (git history; profiles/r2_summary.md: register-staged 16-byte gathers with
// byte permutes, a proxy fence per warp and tile, 583 warp instructions and ~190 L1 data-pipe wavefronts per facet at an
// IPC of 0.45 per sub-partition, every aggregator warp running load -> wait -> compute -> fence) differed in this:
//   * the 16 gathered rows of a facet go global -> shared with cp.async (16-byte units, eight consecutive lanes copy one
//     128-byte line: 4 wavefronts per instruction instead of 16) into a 4 KB per-warp stage, 128B-swizzled, and come
//     back as B fragments with ldmatrix.x4.trans (natural channel order: no byte permutes, no register staging);
//   * the copies of facet n+1 are issued as soon as the fragments of facet n are in registers and land under its MMAs,
//     its drain and the softmax of facet n+1;
//   * aggregator warps never fence: they hand a tile over with one mbarrier arrive per warp (release), a dedicated
//     issuer warp waits (acquire), executes the generic->async proxy fence and issues stage 2 -- so copies in flight
//     never stall a hand-over (MEMBAR waits for every outstanding memory operation of its thread);
//   * ids are loaded once per facet (lane k < 16: slot k), row indices are computed once and spread by shuffles,
//     the neighbour count is one ballot.
constexpr int kH2Threads = (4 + kHAgg) * 32;         // 4 epilogue warps (warp 0 also issues stage 2) + aggregators:
                                                     // five warps per sub-partition, 96 registers each
constexpr int kH2StageBytes = 4096;                  // [hi: slots 0-7 | slots 8-15][lo: same], rows of 128 B

template <int M>
struct Hm2Cfg : HmCfg<M> {
  static constexpr int OFF_STAGE = 2 * HmCfg<M>::B3_BUF;
  static constexpr int OFF_ROW2 = OFF_STAGE + kHAgg * kH2StageBytes;
  static constexpr int OFF_BAR2 = OFF_ROW2 + 4 * kHT * 4;
  static constexpr int SMEM_BYTES2 = OFF_BAR2 + 512;
};

enum { H2_B3_FREE = 0, H2_D_FULL = 2, H2_D_FREE = 4, H2_B3_FULL = 6, H2_ROWS_HI = 8, H2_ROWS_LO = 8 + kHAgg, H2_NUM = 8 + 2 * kHAgg };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}

struct Hm2Pre {           // per-item inputs of the softmax
  float2 vl[4];           // neighbour logits of weights g and 8 of this lane's four slots (2t, 2t+1, 2t+8, 2t+9)
  float2 uo;              // own logits of weights g and 8
  int okm;                // bit i: slot i of this lane holds a valid neighbour
  int cnt;                // non-zero ids among the item's 16 slots
};

// soft assignments of the item in the A-fragment layout: a[0..3] = hi of rows g (a0,a2) / rows 8+ (a1,a3), a[4..7] = lo
template <int M>
__device__ __forceinline__ void hm2_softmax(const Hm2Pre& in, bool recentre, uint32_t (&a)[8]) {
  // the four slots of the lane side by side: every shuffle round is four independent shuffles (one slot after the other
  // is a chain of ~12 dependent shuffle / MUFU latencies per slot, which five warps per sub-partition do not cover)
  float ag[4], a8[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ag[i] = in.uo.x + in.vl[i].x;
    a8[i] = (M == 9) ? in.uo.y + in.vl[i].y : 0.f;
  }
  if (recentre) {   // warp-uniform: the pre-pass saw logits spread over > 60 binary orders
    float mx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) mx[i] = (M == 9) ? fmaxf(ag[i], a8[i]) : ag[i];
#pragma unroll
    for (int sh = 4; sh <= 16; sh <<= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], sh));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) ag[i] -= mx[i], a8[i] -= mx[i];
  }
  float eg[4], e8[4], z[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    eg[i] = ex2_approx(ag[i]);
    e8[i] = (M == 9) ? ex2_approx(a8[i]) : 0.f;
    z[i] = eg[i];
  }
#pragma unroll
  for (int sh = 4; sh <= 16; sh <<= 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] += __shfl_xor_sync(0xffffffffu, z[i], sh);
  }
  float qg[4], q8[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float rs = ((in.okm >> i) & 1) ? rcp_approx(z[i] + e8[i]) : 0.f;
    qg[i] = eg[i] * rs;
    q8[i] = e8[i] * rs;
  }
  split_rn(qg[0], qg[1], a[0], a[4]);
  split_rn(qg[2], qg[3], a[2], a[6]);
  if (M == 9) {
    split_rn(q8[0], q8[1], a[1], a[5]);
    split_rn(q8[2], q8[3], a[3], a[7]);
  } else {
    a[1] = a[4], a[3] = a[6];     // rows 8..15 carry q_lo
    a[5] = 0u, a[7] = 0u;
  }
}

// S of facet f (C fragments, natural channel order: acc[u][j] = row g, channel 8u + 2t + j; acc[u][2 + j] = row 8+) ->
// fp16 hi/lo rows of the stage-2 B operand.  K order of stage 2 (prep_wt_kernel writes the weights in the same order):
// atom a = 2 (m >> 1) + h, chunk 4 (m & 1) + t, element 2 (u & 3) + j  <->  weight m, channel 8 (4 h + (u & 3)) + 2 t + j;
// atom 8 = weight 8 in natural channel order.  The eight lanes of a quarter warp store to eight different chunks.
template <int M, bool HALF = false>
__device__ __forceinline__ void hm2_drain(uint8_t* b3, int f, int g, int t, const float (&acc)[8][4]) {
  using Cfg = HmCfg<M>;
  const int sw = f & 7;
  uint8_t* rh = b3 + (f >> 3) * 1024 + sw * 128;              // hi row f of an atom
  uint8_t* rl = rh + (kHT >> 3) * 1024;                       // lo row 32 + f
  const int chunk = (((g & 1) * 4 + t) ^ sw) << 4;
#pragma unroll
  for (int h = 0; h < (HALF ? 1 : 2); ++h) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v0 = acc[4 * h + j][0], v1 = acc[4 * h + j][1];
      if (M == 8) v0 += acc[4 * h + j][2], v1 += acc[4 * h + j][3];
      split_trunc(v0, v1, hi[j], lo[j]);
    }
    const int aoff = (2 * (g >> 1) + h) * Cfg::ATOM_BYTES + chunk;
    *reinterpret_cast<uint4*>(rh + aoff) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(rl + aoff) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  if (M == 9) {
    // row m = 8 (every lane holds a copy of its 16 channels): lane (g,t) converts channels 8g + 2t, +1
    float e0 = acc[0][2], e1 = acc[0][3];
#pragma unroll
    for (int i = 1; i < (HALF ? 4 : 8); ++i)
      if (g == i) e0 = acc[i][2], e1 = acc[i][3];
    uint32_t hi, lo;
    split_trunc(e0, e1, hi, lo);
    const int off = 8 * Cfg::ATOM_BYTES + ((g ^ sw) << 4) + t * 4;
    if (!HALF || g < 4) {
      *reinterpret_cast<uint32_t*>(rh + off) = hi;
      *reinterpret_cast<uint32_t*>(rl + off) = lo;
    }
  }
}

// HALF: the layer aggregates 32 channels (conv2 of the network): the upper half of the 64-channel unit is zeros, so its
// copies, fragments, MMAs, drain stores and stage-2 K atoms are skipped.
template <int M, int NG, bool TMA, bool HALF = false>
__global__ void __launch_bounds__(kH2Threads, 1)
conv_hm2_kernel(const __grid_constant__ HmParams p) {
  static_assert(!(HALF && TMA), "the half-unit instantiation uses the cp.async path");
  constexpr int NUP = HALF ? 2 : 4;   // pairs of 8-channel blocks with data
  using Cfg = Hm2Cfg<M>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + H2_NUM);
  float* invtab = reinterpret_cast<float*>(tmem_slot + 2);          // [33] 1 / cnt (0 for cnt = 0)
  float* rowinv = reinterpret_cast<float*>(smem + Cfg::OFF_ROW2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nepi = p.cout >> 4;   // epilogue warps with outputs
  const int my_tiles = static_cast<int>((p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&bars[H2_B3_FREE + i], 1);
      tc::mbar_init(&bars[H2_D_FULL + i], 1), tc::mbar_init(&bars[H2_D_FREE + i], nepi);
      tc::mbar_init(&bars[H2_B3_FULL + i], kHAgg);
    }
    for (int i = 0; i < 2 * kHAgg; ++i) tc::mbar_init(&bars[H2_ROWS_HI + i], 1);   // TMA: one plane of one warp's stage
    tc::mbar_fence_init();
  }
  // programmatic dependent launch: the next launch of the layer (or any kernel launched with the attribute) may start its
  // prologue on an SM as soon as this CTA has left it; it waits (griddepcontrol.wait) before it touches data
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 4) tc::tmem_alloc(tmem_slot, 512);
  if (warp == 5) invtab[lane + 1] = 1.f / static_cast<float>(lane + 1), invtab[0] = 0.f;
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // weight operand -> TMEM: lane 32 q + l of the image is TMEM lane 32 q + l, column = K pair; a warp reaches the TMEM
    // lanes of quadrant warp % 4 only.  The epilogue warps with outputs load their own quadrant while the aggregators are
    // already at work on the first tile (stage 2 is the first reader; quadrants without outputs -- lanes of D that nobody
    // reads -- are left as they are): the prologue, 147 KB per CTA from L2, is off the critical path of short launches.
    if (warp < nepi) {
      const uint32_t* src = p.wt + static_cast<size_t>(warp * 32 + lane) * Cfg::W_COLS;
#pragma unroll 1
      for (int c0 = 0; c0 < Cfg::W_COLS; c0 += 32) {
        uint32_t r[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 tq = __ldg(reinterpret_cast<const uint4*>(src + c0) + u);
          r[4 * u] = tq.x, r[4 * u + 1] = tq.y, r[4 * u + 2] = tq.z, r[4 * u + 3] = tq.w;
        }
        tc::tmem_st32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      }
      tc::tc_wait_st();
      tc::tc_fence_before_sync();
      asm volatile("bar.sync 1, %0;" ::"r"(nepi * 32) : "memory");
      tc::tc_fence_after_sync();
    }
  }
  // everything below reads what earlier kernels of the stream wrote (image, logit tables, scales, partial y): the weight
  // image above is the only input that is older (fgc_net_prepare / launch_conv_hm_weights of an earlier launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < 4) {
    // =========================================================== epilogue: Y (TMEM) -> global; warp 0 also issues stage 2
    if (warp == 0) {
      constexpr uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(Cfg::ND) >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t sb = tc::smem_u32(smem);
      auto issue_stage2 = [&](int it) {
        const int buf = it & 1;
        const uint32_t par = (it >> 1) & 1;
        tc::mbar_wait_lazy(&bars[H2_B3_FULL + buf], par);   // all 16 aggregator warps have stored their rows of S
        HM_TR(it, 2);
        tc::fence_proxy_async_smem();                          // ... and the tensor core may read them
        tc::mbar_wait(&bars[H2_D_FREE + buf], par ^ 1);
        HM_TR(it, 3);
        tc::tc_fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t b3 = sb + buf * Cfg::B3_BUF;
#pragma unroll 1
          for (int a = 0; a < Cfg::NATOM; ++a) {
            if (HALF && a < 8 && (a & 1)) continue;       // atoms of the upper channel half: all zero
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (HALF && a == 8 && ks >= 2) continue;    // weight 8, channels 32..63
              const uint64_t bd = tc::smem_desc_k_sw128(b3 + a * Cfg::ATOM_BYTES + ks * 32);
              tc::mma_f16_ts(tmem + Cfg::D_COL + buf * Cfg::ND, tmem + a * 32 + ks * 8, bd, idesc, (a | ks) ? 1u : 0u);
            }
          }
          tc::tc_commit(&bars[H2_B3_FREE + buf]);
          tc::tc_commit(&bars[H2_D_FULL + buf]);
        }
        __syncwarp();
      };
      // stage 2 of tile it + 1 is issued BEFORE the epilogue of tile it: its MMAs then run under this warp's epilogue
      // (issue -> epilogue -> issue in program order made the tile period T_mma + T_epilogue, which bound the kernel)
      auto before_tile = [&](int it) {
        if (it == 0) issue_stage2(0);
        if (it + 1 < my_tiles) issue_stage2(it + 1);
      };
      hm_epilogue<Cfg, H2_D_FULL, H2_D_FREE>(p, bars, rowinv, tmem, warp, lane, before_tile);
    } else if (warp < nepi) {
      hm_epilogue<Cfg, H2_D_FULL, H2_D_FREE>(p, bars, rowinv, tmem, warp, lane);
    }
  } else {
    // =========================================================== aggregators
    // Facet n of this warp: tile counter n / kHFpw, facet n % kHFpw of the warp's own; item m = n * NG + slot group.
    const int aw = warp - 4;
    const int g = lane >> 2, t = lane & 3;
    const int nitems = my_tiles * kHFpw * NG;
    const int rows32 = static_cast<int>(p.rows);
    const float2* lg_v = reinterpret_cast<const float2*>(p.lg) + g;                             // neighbour pair of row j: lg_v[8 j]
    const float2* lg_u = lg_v + static_cast<int64_t>(p.zrow + 1) * 8;                           // own pair of row r: lg_u[8 r]
    const bool recentre = __ldg(p.flag) != 0;
    const uint32_t stage = tc::smem_u32(smem + Cfg::OFF_STAGE + aw * kH2StageBytes);
    // copies: this lane moves 16-byte unit cc of the rows of slots cq, cq + 4, cq + 8, cq + 12 (hi plane, then lo plane)
    const int cq = lane >> 3, cc = lane & 7;
    const uint32_t cdst0 = stage + cq * 128 + ((cc ^ cq) << 4);
    const uint32_t cdst1 = stage + (cq + 4) * 128 + ((cc ^ (cq + 4)) << 4);
    const uint4* img_c = p.img + cc;
    // fragments: lane supplies row (lane & 7) of matrix lane >> 3 = (slots 0-7 | slots 8-15) x (unit 2 up | unit 2 up + 1)
    const uint32_t laddr = (stage + ((lane >> 3) & 1) * 1024 + (lane & 7) * 128) | ((((lane >> 4) ^ lane) & 7) << 4);
    auto row_of = [&](int m) -> int {   // global row of item m, -1 when there is none
      const int n = m / NG;
      const int r = (static_cast<int>(blockIdx.x) + (n / kHFpw) * static_cast<int>(gridDim.x)) * kHT + aw * kHFpw + (n % kHFpw);
      return (m < nitems && r < rows32) ? r : -1;
    };
    auto load_id = [&](int m, int r) -> int {   // id of slot (lane & 15) of item m (row r), 0 when there is none
      const int k = (m % NG) * 16 + (lane & 15);
      return (r >= 0 && k < p.K) ? __ldg(p.adj + static_cast<int64_t>(r) * p.K + k) : 0;
    };
    int base_cur = 0, base_next = p.N;     // batch element of the rows the issue stage walks (monotone)
    // row index of this lane's slot + the item's validity / count bits
    auto prep_rows = [&](int r, int id, Hm2Pre& o) -> int {
      int base = 0;
      if (!p.single && r >= 0) {
        while (r >= base_next) base_cur = base_next, base_next += p.N;
        base = base_cur;
      }
      const bool ok = static_cast<unsigned>(id - 1) < static_cast<unsigned>(p.N);
      const int row = ok ? ((base + id - 1) >> p.upshift) : p.zrow;   // zrow: all-zero image / logit row
      const unsigned okb = __ballot_sync(0xffffffffu, ok);
      const unsigned nzb = __ballot_sync(0xffffffffu, id != 0) & 0xFFFFu;
      o.okm = static_cast<int>(((okb >> (2 * t)) & 3u) | (((okb >> (2 * t + 8)) & 3u) << 2));
      o.cnt = __popc(nzb);
      return row;
    };
    // one plane (0 hi, 1 lo) of the item's 16 rows -> this warp's stage.  TMA: four gathers of four rows (512 B each) by
    // lanes 0..3, completion on the plane's mbarrier (phase = item parity); otherwise 16-byte cp.async, one commit group.
    auto copy_plane = [&](int row, int plane) {
      if constexpr (TMA) {
        int rr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = __shfl_sync(0xffffffffu, row, 4 * (lane & 3) + j);
        uint64_t* bar = &bars[(plane ? H2_ROWS_LO : H2_ROWS_HI) + aw];
        if (lane == 0) tc::mbar_arrive_expect_tx(bar, 2048);
        __syncwarp();
        if (lane < 4)
          tc::tma_gather4_rows(stage + plane * 2048 + lane * 512, p.tmap, bar, p.unit_col + plane * 64, rr[0], rr[1], rr[2],
                               rr[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = __shfl_sync(0xffffffffu, row, cq + 4 * i);
          if (!HALF || cc < 4)
            cp_async16(((i & 1) ? cdst1 : cdst0) + (i >> 1) * 1024 + plane * 2048,
                       img_c + static_cast<int64_t>(rr) * p.img_ld + plane * 8);
        }
        cp_async_commit();
      }
    };
    // the plane of item m has landed (cp.async: the groups complete in order hi(m), lo(m), hi(m+1), ...)
    auto wait_plane = [&](int m, int plane, bool next_hi_pending) {
      if constexpr (TMA) {
        tc::mbar_wait(&bars[(plane ? H2_ROWS_LO : H2_ROWS_HI) + aw], m & 1);
      } else {
        if (plane == 0 || next_hi_pending) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
    };
    auto issue_logits = [&](int r, int row, Hm2Pre& o) {
      o.uo = __ldg(lg_u + static_cast<int64_t>((r < 0 ? 0 : r) >> p.upshift) * 8);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = __shfl_sync(0xffffffffu, row, 2 * t + (i & 1) + 8 * (i >> 1));
        o.vl[i] = __ldg(lg_v + static_cast<int64_t>(rr) * 8);
      }
    };
    float acc[8][4];
    int cnt = 0;
    Hm2Pre P;
    int r1 = row_of(0);
    int idn = load_id(0, r1);
    {
      const int row = prep_rows(r1, idn, P);
      copy_plane(row, 0);
      copy_plane(row, 1);
      issue_logits(r1, row, P);
    }
    r1 = row_of(1);
    idn = load_id(1, r1);
    // One item.  In flight while it is computed: the copies of the next item's rows (hi plane from the moment this item's
    // hi fragments are in registers, lo plane likewise), the next item's logits (from before the MMAs) and the ids of the
    // item after that.
    auto step = [&](int m, auto first_c, auto last_c) {
      constexpr bool FIRST = decltype(first_c)::value, LAST = decltype(last_c)::value;
      uint32_t a[8];
      hm2_softmax<M>(P, recentre, a);
      cnt = FIRST ? P.cnt : cnt + P.cnt;
      const bool more = m + 1 < nitems;
      uint32_t bf[4][4];
      wait_plane(m, 0, true);
#pragma unroll
      for (int up = 0; up < NUP; ++up) ldsm_x4_t(laddr ^ (up << 5), bf[up]);
      __syncwarp();
      int row = 0;
      if (more) row = prep_rows(r1, idn, P);       // overwrites okm / cnt of P: both consumed above
      if (!TMA && more) copy_plane(row, 0);
      // the next item's logits are requested BEFORE this item's MMAs: a warp's facet is one long latency chain (~3 900
      // cycles with four aggregator warps per sub-partition) and, requested after the MMAs as first written, the L2 latency
      // of these loads showed at the top of the next softmax (0.522 -> 0.492 ms per 562 k rows)
      if (more) issue_logits(r1, row, P);
      // eight independent accumulators per round (back-to-back MMAs into one accumulator wait for each other)
#pragma unroll
      for (int u = 0; u < 2 * NUP; ++u)   // M = 8: rows g: q_hi.x_hi, rows g+8: q_lo.x_hi
        hm_mma<FIRST>(acc[u], a[0], a[1], a[2], a[3], bf[u >> 1][2 * (u & 1)], bf[u >> 1][2 * (u & 1) + 1]);
      // TMA writes through the async proxy: the gathers into the plane go out once every fragment register of the plane
      // has been consumed by an MMA, i.e. its ldmatrix reads are complete
      if (TMA && more) copy_plane(row, 0);
      if (M == 9) {
#pragma unroll
        for (int u = 0; u < 2 * NUP; ++u)
          hm_mma<false>(acc[u], a[4], a[5], a[6], a[7], bf[u >> 1][2 * (u & 1)], bf[u >> 1][2 * (u & 1) + 1]);
      }
      wait_plane(m, 1, more);
#pragma unroll
      for (int up = 0; up < NUP; ++up) ldsm_x4_t((laddr ^ (up << 5)) + 2048, bf[up]);
      __syncwarp();
      if (!TMA && more) copy_plane(row, 1);
      const int r2 = row_of(m + 2);
      idn = load_id(m + 2, r2);
#pragma unroll
      for (int up = 0; up < NUP; ++up) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int u = 2 * up + j;
          if (M == 9) hm_mma<false>(acc[u], a[0], a[1], a[2], a[3], bf[up][2 * j], bf[up][2 * j + 1]);
          else hm_mma<false>(acc[u], a[0], 0u, a[2], 0u, bf[up][2 * j], bf[up][2 * j + 1]);
        }
      }
      if (TMA && more) copy_plane(row, 1);
      r1 = r2;
      if (LAST) {
        const int n = m / NG, it = n / kHFpw, fi = n % kHFpw, buf = it & 1, f = aw * kHFpw + fi;
        if (fi == 0) {
#ifdef FGC_HM_TRACE_AGG   // aggregator stamps cost ~15 instructions per facet: only in builds with this flag
          if (aw == 0) HM_TR(it, 9);
#endif
          tc::mbar_wait(&bars[H2_B3_FREE + buf], ((it >> 1) & 1) ^ 1);
#ifdef FGC_HM_TRACE_AGG
          if (aw == 0) HM_TR(it, 10);
#endif
        }
        hm2_drain<M, HALF>(smem + buf * Cfg::B3_BUF, f, g, t, acc);
        if (lane == 0) rowinv[(it & 3) * kHT + f] = invtab[cnt];
        if (fi == kHFpw - 1) {
          __syncwarp();
#ifdef FGC_HM_TRACE_AGG
          if (aw == 0) HM_TR(it, 1);
          if (aw == kHAgg - 1) HM_TR(it, 8);
#endif
          if (lane == 0) tc::mbar_arrive(&bars[H2_B3_FULL + buf]);
        }
      }
    };
#pragma unroll 1
    for (int m = 0; m < nitems; m += NG) {
      if (NG == 1) {
        step(m, std::true_type{}, std::true_type{});
      } else {
        step(m, std::true_type{}, std::false_type{});
        step(m + 1, std::false_type{}, std::true_type{});
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ weight image in TMEM layout
// wt[img][lane][col]: lane 32q + 16h + i <-> row (h, o = 16q + i) of [Wh;Wl] of output block ob, column = K pair of
// aggregation unit u; K order: atom a < 8 holds weights m = 2(a>>1), +1 and the channel units of parity a&1
// (slot s of the atom: m = 2(a>>1) + (s>>2), unit = 2(s&3) + (a&1)), atom 8 holds m = 8.
__global__ void __launch_bounds__(1024)
prep_wt_kernel(const float* __restrict__ W0, uint32_t* __restrict__ wt, float* __restrict__ wunscale, int M, int Cout,
               int Cw, int CB, int nunits) {
  __shared__ float red[32];
  const int total = M * Cout * Cw;
  float mx = 0.f;
  for (int e = threadIdx.x; e < total; e += blockDim.x) mx = fmaxf(mx, fabsf(W0[e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  int E = (__float_as_int(mx) >> 23) & 0xFF;
  E = min(max(E, 16), 240);
  const float sc = __int_as_float((253 - E) << 23);
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) wunscale[0] = __int_as_float((E + 1) << 23);
  const int ob = blockIdx.y / nunits, u = blockIdx.y % nunits;
  const int wcols = M * 32;
  uint32_t* dst = wt + static_cast<size_t>(blockIdx.y) * 128 * wcols;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 128 * wcols; e += gridDim.x * blockDim.x) {
    const int ln = e / wcols, col = e % wcols;
    const int q = ln >> 5, h = (ln >> 4) & 1, o = 16 * q + (ln & 15);
    uint32_t word = 0;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int kpos = 2 * col + kk, a = kpos >> 6, s = (kpos & 63) >> 3, pos = kpos & 7;
      int m, c;
      if (a < 8) {
        m = 2 * (a >> 1) + (s >> 2);
        // hm2_drain's order: chunk s = 4 (m & 1) + t, element pos = 2 (u & 3) + j of atom 2 (m >> 1) + h holds channel
        // 8 (4 h + (u & 3)) + 2 t + j
        c = 8 * (4 * (a & 1) + (pos >> 1)) + 2 * (s & 3) + (pos & 1);
      } else {
        m = 8, c = kpos & 63;
      }
      float v = 0.f;
      if (o < CB && ob * CB + o < Cout && u * 64 + c < Cw && m < M)
        v = W0[(static_cast<size_t>(m) * Cout + ob * CB + o) * Cw + u * 64 + c] * sc;
      const __half hv = __float2half_rn(v);
      const __half out = h ? __float2half_rn((v - __half2float(hv)) * 2048.f) : hv;
      word |= static_cast<uint32_t>(__half_as_ushort(out)) << (16 * kk);
    }
    dst[e] = word;
  }
}

// ------------------------------------------------------------------ max|x| per batch element (image scale)
__global__ void __launch_bounds__(256)
hm_absmax_kernel(const float* __restrict__ x, int64_t n4, unsigned* __restrict__ out) {   // blockIdx.y: batch element
  float m = 0.f;
  x += static_cast<int64_t>(blockIdx.y) * n4 * 4;
  out += blockIdx.y;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// FGC_DISABLE_PDL: launch the convolution kernels fully stream-ordered (no overlap of a launch's prologue with the tail of
// the previous kernel)
bool hm_pdl_enabled() {
  static const bool v = getenv("FGC_DISABLE_PDL") == nullptr;
  return v;
}

size_t hm_img_bytes(int64_t rows_img, int nunits) { return static_cast<size_t>(rows_img) * nunits * 256; }
size_t hm_wt_bytes(int M, int nimg) { return static_cast<size_t>(nimg) * 128 * M * 32 * 4; }

}  // namespace

int debug_hm_trace(int64_t* out, int n) {
  long long host[64 * 16];
  FGC_CUDA(cudaMemcpyFromSymbol(host, g_hm_trace, sizeof(host)));
  for (int i = 0; i < n && i < 64 * 16; ++i) out[i] = host[i];
  return FGC_OK;
}

bool conv_hm_supported(int Cin, int Cw, int Cout, int M, int K) {
  return (M == 8 || M == 9) && K >= 1 && K <= 32 && Cin % 4 == 0 && Cw % 4 == 0 && (Cw == 32 || Cw == 64 || Cw == 128) &&
         (Cout == 32 || Cout == 64 || Cout == 128);
}

size_t conv_hm_weights_bytes(int Cw, int Cout, int M);
size_t conv_hm_workspace(int64_t rows_img, int Cw, int Cout, int M, int B) {
  const int nunits = (Cw + 63) / 64;
  return ws_bytes(hm_img_bytes(rows_img + 1, nunits), 1) + ws_bytes(static_cast<size_t>(rows_img + 1) * 32, 4) +
         ws_bytes(2 * static_cast<size_t>(B) + 16, 4) + ws_bytes(conv_hm_weights_bytes(Cw, Cout, M), 1);
}

size_t conv_hm_weights_bytes(int Cw, int Cout, int M) {
  const int nunits = (Cw + 63) / 64, CB = Cout < 64 ? Cout : 64, nob = Cout / CB;
  return align_up(hm_wt_bytes(M, nunits * nob), 256) + 256;
}

// Weight images of one layer (all output blocks x aggregation units) + the un-scale word behind them:
// constant across calls, so inference prepares them once (fgc_net_prepare).
int launch_conv_hm_weights(const float* W0, int M, int Cout, int Cw, void* wbuf, cudaStream_t st) {
  const int nunits = (Cw + 63) / 64, CB = Cout < 64 ? Cout : 64, nob = Cout / CB;
  float* wunscale = reinterpret_cast<float*>(static_cast<char*>(wbuf) + align_up(hm_wt_bytes(M, nunits * nob), 256));
  prep_wt_kernel<<<dim3(4, nunits * nob), 1024, 0, st>>>(W0, static_cast<uint32_t*>(wbuf), wunscale, M, Cout, Cw, CB, nunits);
  FGC_LAUNCHED("prep_w_image_kernel");
  return FGC_OK;
}

// The convolution proper on prepared operands (launch_prep_rows): img[rows_img + 1][nunits][16] / lg[2][rows_img + 1][16]
// (last row of each zero), flag = the pre-pass's spread flag,
// xunscale[B] = 2^ex of the image rows of every batch element, wbuf from launch_conv_hm_weights; ymax: [B] or null.
int launch_conv_hm_core(const void* img, const float* xunscale, const float* lg, const unsigned* flag, const int32_t* adj,
                        const void* wbuf,
                        const float* b, float* y, float* ypool, unsigned* ymax, int64_t rows, int N, int K, int M, int Cw,
                        int Cout, int upshift, int bias_mask, int act, float alpha, cudaStream_t st, const char* tag) {
  const int nunits = (Cw + 63) / 64, CB = Cout < 64 ? Cout : 64, nob = Cout / CB;
  const int64_t rows_img = rows >> upshift;
  FGC_REQUIRE(ypool == nullptr || (N % 4 == 0), "conv_hm: pooled output needs N %% 4 == 0");
  HmParams hp{};
  hp.img_ld = nunits * 16, hp.xunscale = xunscale, hp.lg = lg, hp.flag = flag, hp.adj = adj;
  hp.wunscale = reinterpret_cast<const float*>(static_cast<const char*>(wbuf) + align_up(hm_wt_bytes(M, nunits * nob), 256));
  hp.ldy = Cout, hp.ldp = Cout, hp.rows = rows, hp.ntiles = (rows + kHT - 1) / kHT;
  hp.N = N, hp.K = K, hp.upshift = upshift, hp.bias_mask = bias_mask, hp.act = act, hp.alpha = alpha;
  hp.cout = CB, hp.single = rows == N, hp.zrow = static_cast<int>(rows_img);
  {
    static const bool trace = getenv("FGC_HM_TRACE") != nullptr;
    hp.trace = trace ? 1 : 0;
  }
  // rows by TMA gather on request (FGC_TMA_MODES bit 4; slower than cp.async here, see make_hm_img_tmap)
  const bool tma = make_hm_img_tmap(hp.tmap, img, rows_img + 1, nunits);
  const bool half = Cw == 32 && M == 9 && !tma && getenv("FGC_HM_NO_HALF") == nullptr;
  auto kern = half ? (K <= 16 ? conv_hm2_kernel<9, 1, false, true> : conv_hm2_kernel<9, 2, false, true>)
              : tma ? (M == 9 ? (K <= 16 ? conv_hm2_kernel<9, 1, true> : conv_hm2_kernel<9, 2, true>)
                              : (K <= 16 ? conv_hm2_kernel<8, 1, true> : conv_hm2_kernel<8, 2, true>))
                    : (M == 9 ? (K <= 16 ? conv_hm2_kernel<9, 1, false> : conv_hm2_kernel<9, 2, false>)
                              : (K <= 16 ? conv_hm2_kernel<8, 1, false> : conv_hm2_kernel<8, 2, false>));
  const int smem = M == 9 ? Hm2Cfg<9>::SMEM_BYTES2 : Hm2Cfg<8>::SMEM_BYTES2;
  const int threads = kH2Threads;
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int64_t grid = num_sms();
  if (grid > hp.ntiles) grid = hp.ntiles;
  if (grid < 1) grid = 1;
  for (int ob = 0; ob < nob; ++ob)
    for (int u = 0; u < nunits; ++u) {
      const bool last = u == nunits - 1;
      hp.img = static_cast<const uint4*>(img) + u * 16;
      hp.unit_col = u * 128;
      hp.wt = static_cast<const uint32_t*>(wbuf) + static_cast<size_t>(ob * nunits + u) * 128 * M * 32;
      hp.b = b + ob * CB, hp.y = y + ob * CB;
      hp.ypool = (last && ypool != nullptr) ? ypool + ob * CB : nullptr;
      hp.ymax = last ? ymax : nullptr;
      hp.add_bias = u == 0, hp.accumulate = u > 0, hp.apply_act = last;
      {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(static_cast<unsigned>(grid)), cfg.blockDim = dim3(threads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = hm_pdl_enabled() ? 1 : 0;
        cfg.attrs = attr, cfg.numAttrs = 1;
        FGC_CUDA(cudaLaunchKernelEx(&cfg, kern, hp));
      }
      FGC_LAUNCHED(tag != nullptr ? tag : "conv_hm_kernel");
    }
  return FGC_OK;
}

// Generic entry (any caller of conv_fwd with plain feature assignment, Cw = Cin): p.x: rows >> upshift rows of Cin floats.
// ypool (optional): [rows / 4][Cout] max over groups of 4 rows of y; ymax (optional): [B] atomicMax targets for max|y| bits.
int launch_conv_hm(const ConvFwdParams& p, const float* W0, const float* u, const float* v, const float* c, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, int upshift, float* ypool, unsigned* ymax) {
  const int nunits = (p.Cw + 63) / 64;
  const int64_t rows_img = p.rows >> upshift;
  Workspace ws(workspace, workspace_bytes);
  const int B = static_cast<int>(p.rows / p.N), Nimg = p.N >> upshift;
  uint4* img = reinterpret_cast<uint4*>(ws.take<char>(hm_img_bytes(rows_img + 1, nunits)));
  float* lg = ws.take<float>((rows_img + 1) * 32);
  unsigned* scal = ws.take<unsigned>(2 * static_cast<size_t>(B) + 16);   // [B] max|x| bits, [B] 2^ex un-scales, flag
  float* xunscale = reinterpret_cast<float*>(scal + B);
  unsigned* flag = scal + 2 * B;
  char* wbuf = ws.take<char>(conv_hm_weights_bytes(p.Cw, p.Cout, p.M));
  FGC_REQUIRE(ws.ok(), "conv_hm: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              conv_hm_workspace(rows_img, p.Cw, p.Cout, p.M, B));
  FGC_CUDA(cudaMemsetAsync(scal, 0, (2 * static_cast<size_t>(B) + 16) * sizeof(unsigned), st));
  int rc = launch_absmax_bits(p.x, static_cast<int64_t>(Nimg) * (p.Cin / 4), B, scal, st);
  if (rc) return rc;
  rc = launch_prep_rows(p.x, p.Cin, p.Cin, nullptr, 0, 0, u, v, c, p.M, rows_img, Nimg, scal, nullptr, img, lg, xunscale, flag, st);
  if (rc) return rc;
  rc = launch_conv_hm_weights(W0, p.M, p.Cout, p.Cw, wbuf, st);
  if (rc) return rc;
  return launch_conv_hm_core(img, xunscale, lg, flag, p.adj, wbuf, p.b, p.y, ypool, ymax, p.rows, p.N, p.K, p.M, p.Cw, p.Cout,
                             upshift, p.bias_mask, p.act, p.alpha, st);
}

// max|x| of every batch element (n4_per_elem float4 each) -> atomicMax on out[b] (bits); out must have been zeroed
int launch_absmax_bits(const float* x, int64_t n4_per_elem, int B, unsigned* out, cudaStream_t st) {
  int bx = (num_sms() * 8 + B - 1) / B;
  if (static_cast<int64_t>(bx) * 256 > n4_per_elem) bx = static_cast<int>((n4_per_elem + 255) / 256);
  if (bx < 1) bx = 1;
  hm_absmax_kernel<<<dim3(bx, B), 256, 0, st>>>(x, n4_per_elem, out);
  FGC_LAUNCHED("absmax_kernel");
  return FGC_OK;
}

}  // namespace fgc
