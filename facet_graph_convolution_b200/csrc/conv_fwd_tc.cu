// Tensor-core forward facet-graph convolution for the dense layers (Cw = 64 aggregation channels).
// Replaces reference Code/model.py:427-504 for shapes where M*Cw*Cout is a dense contraction.
//
// One persistent CTA per SM, 13 warps, tiles of 64 facets:
//   warps 0-7   aggregators.  8 lanes per facet (lane owns 8 channels = two float4 of the row),
//               4 facets per warp and pass, 2 passes per tile.  Soft assignments are evaluated
//               lane-per-(facet,slot) into shared memory, then s[m][c] = sum_k q[k][m] x_{j_k}[c]
//               is accumulated with packed fp32x2 FMAs over coalesced 16-byte row loads.
//               Each finished row (M*64 fp32) is scaled by a power of two to [0.5,1), split into
//               fp16 hi + fp16 lo*2^11 (22 significant bits) and written to a 32-row staging buffer.
//   warps 8-11  movers + epilogue.  Thread-per-row copy staging -> TMEM (tcgen05.st): the tile's
//               A operand is A' = [s_hi ; s_lo] stacked on the 128 TMEM lanes, M*32 columns.
//               After the MMA they read D' (tcgen05.ld) and combine
//                   y = scale * (hi.Wh + 2^-11 (hi.Wl + lo.Wh) + 2^-22 lo.Wl) * inv_cnt + flag * b.
//   warp 12     MMA issuer: D'[128 x 2Cout] = A'[128 x M*64] . [Wh | Wl]  with tcgen05.mma
//               kind::f16 (A from TMEM, B = resident 128B-swizzled K-major smem image of W).
// fp16 hi/lo planes with exact power-of-two scaling carry ~2^-22 relative precision per element,
// i.e. fp32-class results (parity tests: max-abs <= 1e-5 vs the oracle).
#include "conv_common.cuh"
#include "conv_launch.cuh"
#include "tc_agg.cuh"
#include "tc_common.cuh"

namespace fgc {

namespace {

constexpr int kCw = 64;            // aggregation channels handled by this kernel
constexpr int kPrepWBlocks = 16;   // blocks of prep_w_image_kernel
constexpr int kTile = 64;          // facets per tile
constexpr int kPass = 32;          // rows per staging pass
constexpr int kAggWarps = kAggW;     // 32 / kFPW warps cover one pass
constexpr int kMoverWarps = 4;
constexpr int kTcThreads = (kAggWarps + kMoverWarps + 1) * 32;

template <int M, int COUT>
struct TcCfg {
  static constexpr int KK = M * kCw;                 // contraction length
  static constexpr int NB = 2 * COUT;                // MMA N: [Wh | Wl]
  static constexpr int A_COLS = KK / 2;              // TMEM columns of A' (2 halves per column)
  static constexpr int D_COL0 = (A_COLS + 31) & ~31; // first column of D'
  static constexpr int TMEM_COLS = 512;
  static constexpr int MQ = (M + 3) & ~3;            // q row stride in floats
  static constexpr int STG_PITCH = A_COLS + 4;       // words per staged row (pitch % 32 == 4)
  static constexpr int W_BYTES = M * NB * 128;       // M chunks of [NB rows][64 halves]
  static constexpr int EXW = 32;                     // epilogue exchange width (columns per round)
  static_assert(D_COL0 + NB <= TMEM_COLS, "TMEM overflow");
  static_assert(NB % 16 == 0 && NB <= 256, "invalid UMMA N");
  static_assert(STG_PITCH % 32 == 4, "staging pitch must be 4 mod 32 words");
  // shared memory carve-up (bytes)
  static constexpr int OFF_W = 0;
  static constexpr int OFF_STG_H = OFF_W + W_BYTES;
  static constexpr int OFF_STG_L = OFF_STG_H + kPass * STG_PITCH * 4;
  static constexpr int OFF_Q = OFF_STG_L + kPass * STG_PITCH * 4;
  static constexpr int OFF_NBR = OFF_Q + kAggWarps * AggQ<M>::QS_FLOATS * 4;
  static constexpr int OFF_EX = OFF_NBR + kAggWarps * AggQ<M>::NBR_INTS * 4;
  static constexpr int OFF_ROW = OFF_EX + kTile * (EXW + 1) * 4;   // rowscale[2][64], rowflag[2][64]
  static constexpr int OFF_BAR = OFF_ROW + 2 * 2 * kTile * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 128;
};

// MODE_FWD: rows gathered through the adjacency (source-centric forward).
// MODE_TGT: rows gathered through the reversed adjacency (target-centric backward: t = sum q*gz,
//           gx = t . W^T), see conv_bwd.cu.

struct TcParams {
  const float* x;        // gathered rows: x (FWD) or gy (TGT); row stride `Cin`
  const int32_t* adj;    // FWD: adj[rows][K]
  const float* uvx;
  const uint4* wimg;     // swizzled fp16 smem image of [Wh | Wl], W_BYTES
  const float* wunscale; // 2^-aw
  const float* b;
  float* y;              // output rows, row stride `ldy`
  int64_t rows;
  int N, K, Cin;
  int bias_mask, act;
  float alpha;
  int ldy;
  // TGT only
  const int32_t* rev_ptr;
  const int32_t* rev_edge;
  const float* inv;      // inv_cnt of every source row
  const float* da_edge;  // [rows*K][M]
  float* d_uvx;          // [rows][2M], columns M..2M-1 written here
  // FWD only: a layer wider than one launch covers (64 aggregation channels x COUT outputs) is a sum of
  // launches over channel blocks -- the first adds the bias, later ones accumulate into y, the last
  // applies the activation
  int add_bias, accumulate, apply_act;
  int cw;                // aggregation channels present in the rows of this launch (0 = 64)
  int upshift;           // FWD: gathered rows and logits are indexed by (row >> upshift)
};

// barrier indices
enum { B_STG_FULL = 0, B_STG_EMPTY, B_A_READY, B_A_FREE, B_MMA_DONE, B_D_FREE, B_COUNT };

template <int M, int COUT, int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_fwd_tc_kernel(const TcParams p) {
  using Cfg = TcCfg<M, COUT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* stg_h = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_STG_H);
  uint32_t* stg_l = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_STG_L);
  float* qs_all = reinterpret_cast<float*>(smem + Cfg::OFF_Q);
  int* nbr_all = reinterpret_cast<int*>(smem + Cfg::OFF_NBR);
  float* ex = reinterpret_cast<float*>(smem + Cfg::OFF_EX);
  float* rowscale = reinterpret_cast<float*>(smem + Cfg::OFF_ROW);
  float* rowflag = rowscale + 2 * kTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (p.rows + kTile - 1) / kTile;

  // ---------------- one-time setup
  if (threadIdx.x == 0) {
    tc::mbar_init(&bars[B_STG_FULL], kAggWarps);
    tc::mbar_init(&bars[B_STG_EMPTY], 2);
    tc::mbar_init(&bars[B_A_READY], 4);
    tc::mbar_init(&bars[B_A_FREE], 1);
    tc::mbar_init(&bars[B_MMA_DONE], 1);
    tc::mbar_init(&bars[B_D_FREE], kMoverWarps);
    tc::mbar_fence_init();
  }
  if (warp == kAggWarps + kMoverWarps) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  {
    // resident weight image: plain 16-byte copies (the global image is already swizzled)
    uint4* wdst = reinterpret_cast<uint4*>(smem + Cfg::OFF_W);
    for (int i = threadIdx.x; i < Cfg::W_BYTES / 16; i += kTcThreads) wdst[i] = __ldg(p.wimg + i);
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < kAggWarps) {
    // =========================================================== aggregators
    float* qs = qs_all + warp * AggQ<M>::QS_FLOATS;
    int* nbr = nbr_all + warp * AggQ<M>::NBR_INTS;
    const int grp = lane / kLPG;     // facet within the warp
    const int gl = lane % kLPG;      // lane within the facet's group
    uint32_t empty_parity = 1;       // producer convention: the first wait falls through
    int it = 0;
    const AggSrc src{p.x, p.Cin, p.adj, p.uvx, p.N, p.K, p.rows, p.rev_ptr, p.rev_edge, p.inv, p.da_edge, p.cw, p.upshift};
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      for (int pass = 0; pass < 2; ++pass) {
        const int prow = warp * kFPW + grp;                       // row within the pass (0..31)
        const int trow = pass * kPass + prow;                  // row within the tile
        const int64_t r = tile * kTile + trow;                 // global row of this lane's facet
        float2 acc[M][kCP];
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
          for (int i = 0; i < kCP; ++i) acc[m][i] = make_float2(0.f, 0.f);
        int cnt = 0;
        float dv[kPairIters][M];  // TGT: per-lane partial sums of da_edge -> d_uvx[:, M:2M]
#pragma unroll
        for (int h = 0; h < kPairIters; ++h)
#pragma unroll
          for (int m = 0; m < M; ++m) dv[h][m] = 0.f;
        const int64_t wrow0 = tile * kTile + pass * kPass + warp * kFPW;
        tc_aggregate<M, MODE>(src, wrow0, qs, nbr, lane, acc, cnt, dv);
        if constexpr (MODE == MODE_TGT) {
          // d_uvx[t, M + m] = sum of da_edge over the in-edges of t: reduce the 16 lanes of a facet
#pragma unroll
          for (int h = 0; h < kPairIters; ++h) {
            const int64_t rf = wrow0 + (lane >> 4) + 2 * h;
#pragma unroll
            for (int m = 0; m < M; ++m) {
              float t = dv[h][m];
              t += __shfl_xor_sync(0xffffffffu, t, 1);
              t += __shfl_xor_sync(0xffffffffu, t, 2);
              t += __shfl_xor_sync(0xffffffffu, t, 4);
              t += __shfl_xor_sync(0xffffffffu, t, 8);
              if ((lane & 15) == 0 && rf < p.rows) p.d_uvx[rf * (2 * M) + M + m] = t;
            }
          }
        }
        // ---- row scale: largest magnitude of the facet's M*64 values -> [0.5, 1)
        float mx = 0.f;
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
          for (int i = 0; i < kCP; ++i) mx = fmaxf(mx, fmaxf(fabsf(acc[m][i].x), fabsf(acc[m][i].y)));
#pragma unroll
        for (int o = 1; o < kLPG; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int E = (__float_as_int(mx) >> 23) & 0xFF;
        E = min(max(E, 16), 240);                                  // keep both scales normal
        const float sc = __int_as_float((253 - E) << 23);          // 2^(126-E)
        const float unsc = __int_as_float((E + 1) << 23);          // 2^(E-126)
        // ---- wait until the movers drained the previous pass, then stage hi / lo planes
        tc::mbar_wait(&bars[B_STG_EMPTY], empty_parity);
        empty_parity ^= 1;
        uint32_t* rh = stg_h + prow * Cfg::STG_PITCH;
        uint32_t* rl = stg_l + prow * Cfg::STG_PITCH;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          uint32_t h[kCP], l[kCP];
#pragma unroll
          for (int i = 0; i < kCP; ++i) split_pair(acc[m][i].x * sc, acc[m][i].y * sc, h[i], l[i]);
#pragma unroll
          for (int i = 0; i < kF4; ++i) {   // float4 i of the lane = channels 4(gl + kLPG i) .. +3
            *reinterpret_cast<uint2*>(rh + agg_word(m, gl, 2 * i)) = make_uint2(h[2 * i], h[2 * i + 1]);
            *reinterpret_cast<uint2*>(rl + agg_word(m, gl, 2 * i)) = make_uint2(l[2 * i], l[2 * i + 1]);
          }
        }
        if (gl == 0) {
          const float inv = (MODE == MODE_TGT) ? 1.f : (cnt ? 1.f / static_cast<float>(cnt) : 0.f);
          rowscale[(it & 1) * kTile + trow] = inv * unsc;
          rowflag[(it & 1) * kTile + trow] = (cnt > 0 || !p.bias_mask) ? 1.f : 0.f;
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars[B_STG_FULL]);
      }
    }
  } else if (warp < kAggWarps + kMoverWarps) {
    // =========================================================== movers + epilogue
    const int quad = warp - kAggWarps;                 // == warp % 4: TMEM lane quadrant
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    uint32_t full_parity = 0, afree_parity = 1, done_parity = 0;
    const float wun = __ldg(p.wunscale);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      // A' of the previous tile must have been consumed by the MMA
      tc::mbar_wait(&bars[B_A_FREE], afree_parity);
      afree_parity ^= 1;
      tc::tc_fence_after_sync();
      for (int pass = 0; pass < 2; ++pass) {
        tc::mbar_wait(&bars[B_STG_FULL], full_parity);
        full_parity ^= 1;
        const bool is_hi = (quad == pass), is_lo = (quad == 2 + pass);
        if (is_hi || is_lo) {
          const uint32_t* src = (is_hi ? stg_h : stg_l) + lane * Cfg::STG_PITCH;
#pragma unroll 1
          for (int c0 = 0; c0 < Cfg::A_COLS; c0 += 32) {
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 t = *reinterpret_cast<const uint4*>(src + c0 + 4 * i);
              r[4 * i] = t.x, r[4 * i + 1] = t.y, r[4 * i + 2] = t.z, r[4 * i + 3] = t.w;
            }
            tc::tmem_st32(tmem + lane_base + c0, r);
          }
          tc::tc_wait_st();
          tc::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive(&bars[B_STG_EMPTY]);
            tc::mbar_arrive(&bars[B_A_READY]);
          }
        }
      }
      // ---------------- epilogue of this tile
      tc::mbar_wait(&bars[B_MMA_DONE], done_parity);
      done_parity ^= 1;
      tc::tc_fence_after_sync();
      const int frow = (quad & 1) * 32 + lane;          // facet row within the tile
      const int64_t r = tile * kTile + frow;
      const float* rs = rowscale + (it & 1) * kTile;
      const float* rf = rowflag + (it & 1) * kTile;
      for (int c0 = 0; c0 < COUT; c0 += Cfg::EXW) {
        uint32_t d0[32], d1[32];
        tc::tmem_ld32(tmem + lane_base + Cfg::D_COL0 + c0, d0);          // (.)Wh columns
        tc::tmem_ld32(tmem + lane_base + Cfg::D_COL0 + COUT + c0, d1);   // (.)Wl columns
        tc::tc_wait_ld();
        if (quad >= 2) {   // lo rows: 2^-11 lo.Wh + 2^-22 lo.Wl
#pragma unroll
          for (int i = 0; i < 32; ++i)
            ex[frow * (Cfg::EXW + 1) + i] =
                __uint_as_float(d0[i]) * (1.f / 2048.f) + __uint_as_float(d1[i]) * (1.f / 4194304.f);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (quad < 2 && r < p.rows) {
          const float sc = rs[frow] * wun, fl = rf[frow];
          float* yr = p.y + r * p.ldy + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v = __uint_as_float(d0[i + j]) + __uint_as_float(d1[i + j]) * (1.f / 2048.f) +
                              ex[frow * (Cfg::EXW + 1) + i + j];
              float yv;
              if constexpr (MODE == MODE_FWD) {
                yv = sc * v;
                if (p.add_bias) yv = fmaf(fl, __ldg(p.b + c0 + i + j), yv);
                if (p.accumulate) yv += yr[i + j];
                if (p.apply_act && p.act == FGC_ACT_LRELU) yv = lrelu_f(yv, p.alpha);
              } else {
                yv = sc * v;
              }
              o[j] = yv;
            }
            *reinterpret_cast<float4*>(yr + i) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars[B_D_FREE]);
    }
  } else {
    // =========================================================== MMA issuer
    uint32_t ready_parity = 0, dfree_parity = 1;
    const uint32_t idesc = tc::idesc_f16(128, Cfg::NB);
    const uint32_t wbase = tc::smem_u32(smem + Cfg::OFF_W);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      tc::mbar_wait(&bars[B_A_READY], ready_parity);
      ready_parity ^= 1;
      tc::mbar_wait(&bars[B_D_FREE], dfree_parity);
      dfree_parity ^= 1;
      tc::tc_fence_after_sync();
      if (lane == 0) {
#pragma unroll 1
        for (int kc = 0; kc < M; ++kc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bdesc = tc::smem_desc_k_sw128(wbase + kc * (Cfg::NB * 128) + ks * 32);
            tc::mma_f16_ts(tmem + Cfg::D_COL0, tmem + kc * 32 + ks * 8, bdesc, idesc, (kc | ks) ? 1u : 0u);
          }
        }
        tc::tc_commit(&bars[B_A_FREE]);
        tc::tc_commit(&bars[B_MMA_DONE]);
      }
      __syncwarp();
    }
  }
  // ---------------- teardown
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == kAggWarps + kMoverWarps) tc::tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------ weight image preparation
// wimg: M chunks of [NB rows][64 halves] (128 B per row, 16-byte units XOR-swizzled by row & 7):
//   row n < COUT : fp16(W0[m][n][c] * 2^aw)                      (hi)
//   row n >= COUT: fp16((W*2^aw - hi) * 2^11)                    (lo)
// Every block reduces max|W| itself (the weights are a few hundred KB in L2: cheaper than a second launch
// or a grid barrier) -> power-of-two scale so that |W*2^aw| < 1; then the blocks share the image elements.
// The image may be a block of a wider layer: outputs o0 .. o0+COUT-1 of ldo, channels c0 .. c0+63 of ldc
// (ldo = COUT, ldc = 64, o0 = c0 = 0: the whole W0[M][COUT][64]).
__device__ __forceinline__ size_t w_src_index(int m, int o, int c, int ldo, int ldc, int o0, int c0) {
  return (static_cast<size_t>(m) * ldo + o0 + o) * ldc + c0 + c;
}
// blockIdx.y selects the block of a channel-block decomposition: image y = oh * ncc + ch lies img_stride
// bytes after the previous one (its un-scale word behind its M * 2 COUT * 128 image bytes) and covers
// outputs o0 + COUT oh .., channels c0 + 64 ch ..  (ncc = 1, gridDim.y = 1: a single image).
__global__ void __launch_bounds__(1024)
prep_w_image_kernel(const float* __restrict__ W0, uint16_t* __restrict__ wimg, float* __restrict__ wunscale,
                    int M, int COUT, int transposed, int ldo, int ldc, int o0, int c0, int ncc = 1,
                    size_t img_stride = 0) {
  __shared__ float red[32];
  if (gridDim.y > 1) {
    const int oh = blockIdx.y / ncc, ch = blockIdx.y % ncc;
    o0 += COUT * oh, c0 += kCw * ch;
    wimg = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(wimg) + img_stride * blockIdx.y);
    wunscale = reinterpret_cast<float*>(reinterpret_cast<char*>(wunscale) + img_stride * blockIdx.y);
  }
  const int total = M * COUT * kCw;
  float mx = 0.f;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int c = e % kCw, o = (e / kCw) % COUT, m = e / (kCw * COUT);
    const bool in = transposed || c0 + c < ldc;   // a 32-channel layer fills half of the 64-wide image
    mx = fmaxf(mx, in ? fabsf(transposed ? W0[e] : W0[w_src_index(m, o, c, ldo, ldc, o0, c0)]) : 0.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  int E = (__float_as_int(mx) >> 23) & 0xFF;
  E = min(max(E, 16), 240);
  const float sc = __int_as_float((253 - E) << 23);
  if (threadIdx.x == 0 && blockIdx.x == 0) wunscale[0] = __int_as_float((E + 1) << 23);
  const int NB = 2 * COUT;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    // image element (m, n = o, k = c); transposed: n = c of W0, k = o of W0 (B for gx = t . W^T)
    const int c = e % kCw;
    const int o = (e / kCw) % COUT;
    const int m = e / (kCw * COUT);
    const float v = (transposed ? W0[(static_cast<size_t>(m) * kCw + c) * COUT + o]
                                : (c0 + c < ldc ? W0[w_src_index(m, o, c, ldo, ldc, o0, c0)] : 0.f)) * sc;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn((v - __half2float(h)) * 2048.f);
    const int unit = c >> 3, within = c & 7;
    const size_t chunk = static_cast<size_t>(m) * NB * 64;  // halves
    const int nh = o, nl = COUT + o;
    wimg[chunk + nh * 64 + ((unit ^ (nh & 7)) << 3) + within] = __half_as_ushort(h);
    wimg[chunk + nl * 64 + ((unit ^ (nl & 7)) << 3) + within] = __half_as_ushort(l);
  }
}

template <int M, int COUT, int MODE>
int launch_tc(const TcParams& tp_in, void* wimg, float* wunscale, const float* W0, cudaStream_t st) {
  using Cfg = TcCfg<M, COUT>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
  if (W0 != nullptr) {   // nullptr: the caller already prepared the image
    prep_w_image_kernel<<<kPrepWBlocks, 1024, 0, st>>>(W0, static_cast<uint16_t*>(wimg), wunscale, M, COUT,
                                                       MODE == MODE_TGT ? 1 : 0, COUT, kCw, 0, 0);
    FGC_LAUNCHED("prep_w_image_kernel");
  }
  TcParams tp = tp_in;
  tp.wimg = static_cast<const uint4*>(wimg);
  tp.wunscale = wunscale;
  auto kern = conv_fwd_tc_kernel<M, COUT, MODE>;
  FGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int64_t ntiles = (tp.rows + kTile - 1) / kTile;
  int64_t grid = num_sms();
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), kTcThreads, Cfg::SMEM_BYTES, st>>>(tp);
  FGC_LAUNCHED(MODE == MODE_TGT ? "bwd_tgt_tc_kernel" : "conv_fwd_tc_kernel");
  return FGC_OK;
}

}  // namespace

// One launch covers 64 aggregation channels x COUT outputs with a resident weight image of
// M * 2 COUT * 128 B next to the staging buffers: (M, COUT) = (8, 64) -- the C2 layer -- and (9, 32).
// The network's M = 9 layers (reference Code/model.py:870-932) with 64 / 128 aggregation channels and
// 32 / 64 / 128 outputs are sums of (Cw / 64) x (Cout / 32) such launches over channel blocks.
bool conv_fwd_tc_supported(int Cw, int Cout, int M, int K) {
  if (K > 32) return false;
  if (M == 8) return Cw == kCw && Cout == 64;
  return M == 9 && (Cw == 32 || Cw == 64 || Cw == 128) && (Cout == 32 || Cout == 64 || Cout == 128);
}

static size_t tc_image_bytes(int Cout, int M) { return (static_cast<size_t>(M) * 2 * Cout * 128 + 512 + 255) / 256 * 256; }

size_t conv_fwd_tc_workspace(int Cout, int M, int Cw) {
  if (M == 9) return tc_image_bytes(32, M) * ((Cout + 31) / 32) * ((Cw + 63) / 64);
  return static_cast<size_t>(M) * 2 * Cout * 128 + 512;
}

// wimg_ws: conv_fwd_tc_workspace bytes (16-byte aligned); the tail of every image holds its scalar un-scale
int launch_conv_fwd_tc(const ConvFwdParams& p, const float* W0, void* wimg_ws, cudaStream_t st, int upshift) {
  TcParams tp{};
  tp.upshift = upshift;
  tp.x = p.x, tp.adj = p.adj, tp.uvx = p.uvx, tp.b = p.b, tp.y = p.y, tp.rows = p.rows;
  tp.N = p.N, tp.K = p.K, tp.Cin = p.Cin, tp.bias_mask = p.bias_mask, tp.act = p.act, tp.alpha = p.alpha;
  tp.ldy = p.Cout;
  tp.add_bias = 1, tp.accumulate = 0, tp.apply_act = 1;
  if (p.M == 8 && p.Cout == 64) {
    const size_t img = static_cast<size_t>(p.M) * 2 * p.Cout * 128;
    float* wunscale = reinterpret_cast<float*>(static_cast<char*>(wimg_ws) + img);
    return launch_tc<8, 64, MODE_FWD>(tp, wimg_ws, wunscale, W0, st);
  }
  if (p.M == 9) {
    const int nco = p.Cout / 32, ncc = (p.Cw + 63) / 64;
    const size_t img = static_cast<size_t>(p.M) * 2 * 32 * 128;
    // all weight images of the layer in one launch
    prep_w_image_kernel<<<dim3(kPrepWBlocks, nco * ncc), 1024, 0, st>>>(
        W0, static_cast<uint16_t*>(wimg_ws), reinterpret_cast<float*>(static_cast<char*>(wimg_ws) + img), p.M, 32, 0,
        p.Cout, p.Cw, 0, 0, ncc, tc_image_bytes(32, p.M));
    FGC_LAUNCHED("prep_w_image_kernel");
    for (int oh = 0; oh < nco; ++oh)
      for (int ch = 0; ch < ncc; ++ch) {
        char* base = static_cast<char*>(wimg_ws) + tc_image_bytes(32, p.M) * (oh * ncc + ch);
        float* wunscale = reinterpret_cast<float*>(base + img);
        TcParams t2 = tp;
        t2.x = p.x + 64 * ch, t2.b = p.b + 32 * oh, t2.y = p.y + 32 * oh;
        t2.add_bias = ch == 0, t2.accumulate = ch > 0, t2.apply_act = ch == ncc - 1;
        t2.cw = (p.Cw - 64 * ch >= 64) ? 0 : p.Cw - 64 * ch;
        const int rc = launch_tc<9, 32, MODE_FWD>(t2, base, wunscale, nullptr, st);
        if (rc) return rc;
      }
    return FGC_OK;
  }
  set_error("conv_fwd_tc: unsupported shape");
  return FGC_ERR_UNSUPPORTED;
}

// transposed weight image (chunk m: rows c hi|lo, K = o) shared by bwd_src_tc and bwd_tgt_tc
int launch_prep_w_image_t(const float* W0, void* wimg_ws, int M, int Cw, cudaStream_t st) {
  const size_t img = static_cast<size_t>(M) * 2 * Cw * 128;
  float* wunscale = reinterpret_cast<float*>(static_cast<char*>(wimg_ws) + img);
  prep_w_image_kernel<<<kPrepWBlocks, 1024, 0, st>>>(W0, static_cast<uint16_t*>(wimg_ws), wunscale, M, Cw, 1, Cw, kCw, 0, 0);
  FGC_LAUNCHED("prep_w_image_kernel");
  return FGC_OK;
}

// weight image in forward orientation (chunk m: rows o hi|lo, K = c), shared with conv_mma.cu
int launch_prep_w_image(const float* W0, void* wimg_ws, int M, int Cout, cudaStream_t st) {
  const size_t img = static_cast<size_t>(M) * 2 * Cout * 128;
  float* wunscale = reinterpret_cast<float*>(static_cast<char*>(wimg_ws) + img);
  prep_w_image_kernel<<<kPrepWBlocks, 1024, 0, st>>>(W0, static_cast<uint16_t*>(wimg_ws), wunscale, M, Cout, 0, Cout, kCw, 0, 0);
  FGC_LAUNCHED("prep_w_image_kernel");
  return FGC_OK;
}

// Target-centric backward pass on the same skeleton: gx[:, 0:Cw] = sum_m t[.,m,:] W0[m] with
// t[j,m,:] = sum over in-edges (n->j) of q[n->j,m] * inv_cnt[n] * gy[n,:], plus d_uvx[:, M:2M].
bool bwd_tgt_tc_supported(int Cw, int Cout, int M) { return Cw == 64 && Cout == kCw && M == 8; }

int launch_bwd_tgt_tc(const float* gy, const float* uvx, const float* W0, const float* da_edge,
                      const float* inv, const int32_t* rev_ptr, const int32_t* rev_edge, float* gx,
                      float* d_uvx, int64_t rows, int N, int K, int Cin, int Cw, int Cout, int M,
                      void* wimg_ws, cudaStream_t st) {
  const size_t img = static_cast<size_t>(M) * 2 * Cw * 128;
  float* wunscale = reinterpret_cast<float*>(static_cast<char*>(wimg_ws) + img);
  TcParams tp{};
  tp.x = gy, tp.uvx = uvx, tp.y = gx, tp.rows = rows, tp.N = N, tp.K = K, tp.Cin = Cout;  // gy row stride
  tp.ldy = Cin;
  tp.rev_ptr = rev_ptr, tp.rev_edge = rev_edge, tp.inv = inv, tp.da_edge = da_edge, tp.d_uvx = d_uvx;
  if (M == 8 && Cw == 64) return launch_tc<8, 64, MODE_TGT>(tp, wimg_ws, wunscale, W0, st);
  set_error("bwd_tgt_tc: unsupported shape");
  return FGC_ERR_UNSUPPORTED;
}

}  // namespace fgc
