// Warp-cooperative fast paths of the three logit kernels (bandwidth-bound passes over x):
//   assign_logits   uvx[r] = [u;v] . x_r[window] + [c;0]            reference Code/model.py:74-95
//   logits_bwd_x    gx[r, window] += d_uvx[r] . [u;v]
//   logits_bwd_p    gu, gv, gc = sum_r d_uvx[r] (x) x_r[window]      (fixed-order partial sums)
// LPR lanes share a row (lane owns 4 consecutive channels -> every row access is one coalesced
// 16-byte load per lane), the [u;v] coefficients of the lane's channels live in registers, and the
// per-row reduction over channels is a transpose-reduce over the LPR lanes.  The generic
// thread-per-row kernels in conv_fwd.cu / conv_bwd.cu remain for every other shape.
#include <cuda_fp16.h>

#include "conv_common.cuh"
#include "conv_launch.cuh"

namespace fgc {

namespace {

// v[0..OP) per lane, LPR lanes of a row -> each lane ends with OP/LPR sums: outputs gl*(OP/LPR)+i
template <int OP, int LPR>
__device__ __forceinline__ void transpose_reduce(float (&v)[OP], int gl) {
  int h = OP / 2;
#pragma unroll
  for (int bit = LPR / 2; bit >= 1; bit >>= 1, h >>= 1) {
    const bool up = (gl & bit) != 0;
#pragma unroll
    for (int i = 0; i < OP / 2; ++i) {
      if (i < h) {
        const float keep = up ? v[i + h] : v[i], send = up ? v[i] : v[i + h];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
}

struct LgParams {
  const float* x;
  const float* u;
  const float* v;
  const float* c;
  float* uvx;          // fwd: out; bwd: d_uvx in
  float* gx;
  float* part;
  int64_t rows, rows_per_chunk;
  int Cin, Ca0, Ca, M;
  unsigned* maxbits;   // assign_logits only, optional: atomicMax of the float bits of max|x| over the window
};

// coefficients of this lane's 4 channels: w[o][0..3], o < 2M (zero beyond Ca / 2M)
template <int OP>
__device__ __forceinline__ void load_coeffs(const LgParams& p, int gl, float (&w)[OP][4]) {
#pragma unroll
  for (int o = 0; o < OP; ++o)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = 4 * gl + j;
      float t = 0.f;
      if (o < 2 * p.M && cc < p.Ca) t = (o < p.M) ? __ldg(p.u + o * p.Ca + cc) : __ldg(p.v + (o - p.M) * p.Ca + cc);
      w[o][j] = t;
    }
}

template <int OP, int LPR>
__global__ void __launch_bounds__(256, (OP <= 16) ? 2 : 1)
assign_logits_warp_kernel(const LgParams p) {
  constexpr int RPW = 32 / LPR, OPL = OP / LPR;
  const int lane = threadIdx.x & 31, gl = lane % LPR, sub = lane / LPR;
  const int O = 2 * p.M;
  float w[OP][4];
  load_coeffs<OP>(p, gl, w);
  float cb[OPL];
#pragma unroll
  for (int i = 0; i < OPL; ++i) {
    const int o = gl * OPL + i;
    cb[i] = (o < p.M) ? __ldg(p.c + o) : 0.f;
  }
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const bool act = 4 * gl < p.Ca;
  constexpr int U = 4;   // row groups per iteration: all loads are in flight before the first is consumed
  float amax = 0.f;
  for (int64_t r0 = wid * RPW; r0 < p.rows; r0 += U * nw * RPW) {
    int64_t rr[U];
    float4 xs[U];
#pragma unroll
    for (int t = 0; t < U; ++t) {
      rr[t] = r0 + t * nw * RPW + sub;
      xs[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr[t] < p.rows && act) xs[t] = __ldg(reinterpret_cast<const float4*>(p.x + rr[t] * p.Cin + p.Ca0) + gl);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(xs[t].x), fabsf(xs[t].y)), fmaxf(fabsf(xs[t].z), fabsf(xs[t].w))));
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const float4 xv = xs[t];
      const int64_t r = rr[t];
      float a[OP];
#pragma unroll
      for (int o = 0; o < OP; ++o) a[o] = fmaf(xv.x, w[o][0], fmaf(xv.y, w[o][1], fmaf(xv.z, w[o][2], xv.w * w[o][3])));
      transpose_reduce<OP, LPR>(a, gl);
      if (r < p.rows) {
#pragma unroll
        for (int i = 0; i < OPL; ++i) {
          const int o = gl * OPL + i;
          if (o < O) p.uvx[r * O + o] = a[i] + cb[i];
        }
      }
    }
  }
  if (p.maxbits != nullptr) {   // max is order-independent: the atomic keeps the result deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) atomicMax(p.maxbits, __float_as_uint(amax));
  }
}

template <int OP, int LPR>
__global__ void __launch_bounds__(256, (OP <= 16) ? 2 : 1)
logits_bwd_x_warp_kernel(const LgParams p) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, gl = lane % LPR, sub = lane / LPR;
  const int O = 2 * p.M;
  float w[OP][4];
  load_coeffs<OP>(p, gl, w);
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const bool act = 4 * gl < p.Ca;
  constexpr int U = 2;   // row groups per iteration (loads of both are issued before either is used)
  for (int64_t r0 = wid * RPW; r0 < p.rows; r0 += U * nw * RPW) {
    float d[U][OP];
    float4 g[U];
    bool ok[U];
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const int64_t r = r0 + t * nw * RPW + sub;
      ok[t] = r < p.rows && act;
      g[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < OP; ++o) d[t][o] = 0.f;
      if (!ok[t]) continue;
      if ((O & 3) == 0) {
#pragma unroll
        for (int o = 0; o < OP; o += 4) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (o < O) q = __ldg(reinterpret_cast<const float4*>(p.uvx + r * O + o));
          d[t][o] = q.x, d[t][o + 1] = q.y, d[t][o + 2] = q.z, d[t][o + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int o = 0; o < OP; ++o) d[t][o] = (o < O) ? __ldg(p.uvx + r * O + o) : 0.f;
      }
      g[t] = *(reinterpret_cast<const float4*>(p.gx + r * p.Cin + p.Ca0) + gl);
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      if (!ok[t]) continue;
      const int64_t r = r0 + t * nw * RPW + sub;
      float4 gg = g[t];
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        gg.x = fmaf(d[t][o], w[o][0], gg.x), gg.y = fmaf(d[t][o], w[o][1], gg.y);
        gg.z = fmaf(d[t][o], w[o][2], gg.z), gg.w = fmaf(d[t][o], w[o][3], gg.w);
      }
      *(reinterpret_cast<float4*>(p.gx + r * p.Cin + p.Ca0) + gl) = gg;
    }
  }
}

// one CTA per chunk of rows; warps stride the chunk's rows, fixed-order reduction over the warps
template <int OP, int LPR>
__global__ void __launch_bounds__(256, (OP <= 16) ? 2 : 1)
logits_bwd_p_warp_kernel(const LgParams p) {
  constexpr int RPW = 32 / LPR;
  extern __shared__ float red[];   // [warps * RPW][OP * LPR * 4 + OP]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % LPR, sub = lane / LPR;
  const int nwarp = blockDim.x >> 5;
  const int O = 2 * p.M;
  float acc[OP][4];
  float gc[OP];
#pragma unroll
  for (int o = 0; o < OP; ++o) {
    acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
    gc[o] = 0.f;
  }
  const int64_t rb = static_cast<int64_t>(blockIdx.x) * p.rows_per_chunk;
  const int64_t re = min(p.rows, rb + p.rows_per_chunk);
  const bool act = 4 * gl < p.Ca;
  constexpr int U = 2;   // rows per iteration per lane group (loads of both issued first; fixed order of use)
  for (int64_t r0 = rb + warp * RPW; r0 < re; r0 += U * nwarp * RPW) {
    float4 xs[U];
    float d[U][OP];
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const int64_t r = r0 + t * nwarp * RPW + sub;
      xs[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o = 0; o < OP; ++o) d[t][o] = 0.f;
      if (r >= re) continue;
      if (act) xs[t] = __ldg(reinterpret_cast<const float4*>(p.x + r * p.Cin + p.Ca0) + gl);
      if ((O & 3) == 0) {
#pragma unroll
        for (int o = 0; o < OP; o += 4) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (o < O) q = __ldg(reinterpret_cast<const float4*>(p.uvx + r * O + o));
          d[t][o] = q.x, d[t][o + 1] = q.y, d[t][o + 2] = q.z, d[t][o + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int o = 0; o < OP; ++o) d[t][o] = (o < O) ? __ldg(p.uvx + r * O + o) : 0.f;
      }
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const float4 xv = xs[t];
#pragma unroll
      for (int o = 0; o < OP; ++o) {
        acc[o][0] = fmaf(d[t][o], xv.x, acc[o][0]), acc[o][1] = fmaf(d[t][o], xv.y, acc[o][1]);
        acc[o][2] = fmaf(d[t][o], xv.z, acc[o][2]), acc[o][3] = fmaf(d[t][o], xv.w, acc[o][3]);
        if (gl == 0) gc[o] += d[t][o];
      }
    }
  }
  // slot s = warp * RPW + sub holds this lane group's partials: [o][c] then gc[o]
  const int slot_floats = OP * LPR * 4 + OP;
  float* my = red + (warp * RPW + sub) * slot_floats;
#pragma unroll
  for (int o = 0; o < OP; ++o) {
    *reinterpret_cast<float4*>(my + (o * LPR + gl) * 4) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
    if (gl == 0) my[OP * LPR * 4 + o] = gc[o];
  }
  __syncthreads();
  const int nslot = nwarp * RPW;
  const int nout = O * p.Ca;
  float* out = p.part + static_cast<int64_t>(blockIdx.x) * (nout + p.M);
  for (int e = threadIdx.x; e < nout + p.M; e += blockDim.x) {
    int idx;
    if (e < nout) {
      const int o = e / p.Ca, cc = e % p.Ca;
      idx = (o * LPR + (cc >> 2)) * 4 + (cc & 3);
    } else {
      idx = OP * LPR * 4 + (e - nout);
    }
    float a = 0.f;
    for (int s = 0; s < nslot; ++s) a += red[s * slot_floats + idx];
    out[e] = a;
  }
}

// ------------------------------------------------------------------ fused pre-pass of the HMMA-aggregation forward
// One pass over the rows of a layer input (optionally the channel concatenation [xa | xb] of two tensors, reference
// Code/model.py:909,929, never materialised): assignment logits uvx[r] = [u;v] . x_r + [c;0] and the fp16 hi|lo image
// img[r][unit] = [fp16(x s) (64) | fp16(x s - hi) (64)] with s = 2^(126-E), E = exponent of max(max|xa|, max|xb|)
// per batch element (upper bounds the producing kernels leave behind; per element, so that a batch of patches gives
// every patch exactly the result of running it alone).  Row `rows` of both outputs is zeroed (what padding slots read).
struct PrepRowsParams {
  const float* xa;
  const float* xb;
  const float* u;
  const float* v;
  const float* c;
  const unsigned* maxa;
  const unsigned* maxb;
  uint4* img;
  float* uvx;
  float* xunscale;
  int64_t rows;
  int lda, Ca, ldb, Cb, M, nunits;
  int Nimg;   // rows per batch element: maxa / maxb / xunscale are per element
};

template <int OP, int LPR>
__global__ void __launch_bounds__(256, (OP <= 16) ? 2 : 1)
prep_rows_warp_kernel(const PrepRowsParams p) {
  constexpr int RPW = 32 / LPR, OPL = OP / LPR;
  const int lane = threadIdx.x & 31, gl = lane % LPR, sub = lane / LPR;
  const int O = 2 * p.M, Cin = p.Ca + p.Cb;
  const int c4 = 4 * gl;
  const bool act = c4 < Cin;                 // lane holds real channels
  const bool inimg = c4 < p.nunits * 64;     // lane holds image channels (zeros beyond Cin)
  float w[OP][4];
#pragma unroll
  for (int o = 0; o < OP; ++o)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = 0.f;
      if (o < O && c4 + j < Cin) t = (o < p.M) ? __ldg(p.u + o * Cin + c4 + j) : __ldg(p.v + (o - p.M) * Cin + c4 + j);
      w[o][j] = t;
    }
  float cb[OPL];
#pragma unroll
  for (int i = 0; i < OPL; ++i) {
    const int o = gl * OPL + i;
    cb[i] = (o < p.M) ? __ldg(p.c + o) : 0.f;
  }
  auto scale_exp = [&](int be) {   // exponent of the element's max |x| (clamped so that both scales stay normal)
    unsigned mb = __ldg(p.maxa + be);
    if (p.maxb != nullptr) mb = max(mb, __ldg(p.maxb + be));
    return min(max(static_cast<int>((mb >> 23) & 0xFF), 16), 240);
  };
  {
    const int nelem = static_cast<int>(p.rows / p.Nimg);
    for (int be = blockIdx.x * blockDim.x + threadIdx.x; be < nelem; be += gridDim.x * blockDim.x)
      p.xunscale[be] = __int_as_float((scale_exp(be) + 1) << 23);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < p.nunits * 16) p.img[p.rows * p.nunits * 16 + threadIdx.x] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < O) p.uvx[p.rows * O + threadIdx.x] = 0.f;
  }
  const bool from_a = c4 < p.Ca;
  const float* src = from_a ? p.xa + c4 : p.xb + (c4 - p.Ca);
  const int ld = from_a ? p.lda : p.ldb;
  uint8_t* const imgb = reinterpret_cast<uint8_t*>(p.img) + (c4 >> 6) * 256 + (c4 & 63) * 2;
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  constexpr int U = 4;   // row groups per iteration: all loads are in flight before the first is consumed
  for (int64_t r0 = wid * RPW; r0 < p.rows; r0 += U * nw * RPW) {
    int64_t rr[U];
    float4 xs[U];
#pragma unroll
    for (int t = 0; t < U; ++t) {
      rr[t] = r0 + t * nw * RPW + sub;
      xs[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr[t] < p.rows && act) xs[t] = __ldg(reinterpret_cast<const float4*>(src + rr[t] * ld));
    }
#pragma unroll
    for (int t = 0; t < U; ++t) {
      const float4 xv = xs[t];
      const int64_t r = rr[t];
      if (r < p.rows && inimg) {
        const float sc = __int_as_float((253 - scale_exp(static_cast<int>(r / p.Nimg))) << 23);
        const __half2 h0 = __floats2half2_rn(xv.x * sc, xv.y * sc), h1 = __floats2half2_rn(xv.z * sc, xv.w * sc);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(xv.x * sc - f0.x, xv.y * sc - f0.y);
        const __half2 l1 = __floats2half2_rn(xv.z * sc - f1.x, xv.w * sc - f1.y);
        uint8_t* d = imgb + r * (p.nunits * 256);
        *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(d + 128) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
      float a[OP];
#pragma unroll
      for (int o = 0; o < OP; ++o) a[o] = fmaf(xv.x, w[o][0], fmaf(xv.y, w[o][1], fmaf(xv.z, w[o][2], xv.w * w[o][3])));
      transpose_reduce<OP, LPR>(a, gl);
      if (r < p.rows) {
#pragma unroll
        for (int i = 0; i < OPL; ++i) {
          const int o = gl * OPL + i;
          if (o < O) p.uvx[r * O + o] = a[i] + cb[i];
        }
      }
    }
  }
}

template <int OP, int LPR>
int run_prep_rows(const PrepRowsParams& p, cudaStream_t st, const char* tag) {
  const int64_t rows_per_block = 8 * (32 / LPR);
  int64_t blocks = (p.rows + rows_per_block - 1) / rows_per_block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  prep_rows_warp_kernel<OP, LPR><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
  FGC_LAUNCHED(tag != nullptr ? tag : "prep_rows_kernel");
  return FGC_OK;
}

template <int OP, int LPR>
int run_assign(const LgParams& p, cudaStream_t st) {
  const int64_t rows_per_block = 8 * (32 / LPR);
  int64_t blocks = (p.rows + rows_per_block - 1) / rows_per_block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  assign_logits_warp_kernel<OP, LPR><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
  FGC_LAUNCHED("assign_logits_kernel");
  return FGC_OK;
}

template <int OP, int LPR>
int run_bwd_x(const LgParams& p, cudaStream_t st) {
  const int64_t rows_per_block = 8 * (32 / LPR);
  int64_t blocks = (p.rows + rows_per_block - 1) / rows_per_block;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  logits_bwd_x_warp_kernel<OP, LPR><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
  FGC_LAUNCHED("logits_bwd_x_kernel");
  return FGC_OK;
}

template <int OP, int LPR>
int run_bwd_p(const LgParams& p, int chunks, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(8 * (32 / LPR)) * (OP * LPR * 4 + OP) * 4;
  FGC_CUDA(cudaFuncSetAttribute(logits_bwd_p_warp_kernel<OP, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  logits_bwd_p_warp_kernel<OP, LPR><<<chunks, 256, smem, st>>>(p);
  FGC_LAUNCHED("logits_bwd_p_kernel");
  return FGC_OK;
}

// 0 = no fast path; otherwise lanes per row
int fast_lpr(int Cin, int Ca0, int Ca, int M) {
  if (Cin % 4 || Ca0 % 4 || Ca % 4 || M > 16) return 0;
  if (Ca <= 64) return 16;
  if (Ca <= 128) return 32;
  return 0;
}

}  // namespace

// rows of [xa (Ca channels, row stride lda) | xb (Cb, ldb; may be null)] -> uvx[rows + 1][2M], img[rows + 1][nunits][16]
int launch_prep_rows(const float* xa, int lda, int Ca, const float* xb, int ldb, int Cb, const float* u, const float* v,
                     const float* c, int M, int64_t rows, int Nimg, const unsigned* maxa, const unsigned* maxb, void* img,
                     float* uvx, float* xunscale, cudaStream_t st, const char* tag) {
  const int Cin = Ca + Cb, nunits = (Cin + 63) / 64;
  FGC_REQUIRE(Ca % 4 == 0 && Cb % 4 == 0 && Cin <= 128 && 2 * M <= 32 && lda % 4 == 0 && (xb == nullptr || ldb % 4 == 0),
              "prep_rows: unsupported shape (Ca=%d Cb=%d M=%d)", Ca, Cb, M);
  FGC_REQUIRE(Nimg > 0 && rows % Nimg == 0, "prep_rows: rows must be a multiple of the rows per batch element");
  PrepRowsParams p{xa, xb, u, v, c, maxa, maxb, static_cast<uint4*>(img), uvx, xunscale, rows, lda, Ca, ldb, Cb, M, nunits, Nimg};
  if (Cin <= 64) {
    if (2 * M <= 16) return run_prep_rows<16, 16>(p, st, tag);
    return run_prep_rows<32, 16>(p, st, tag);
  }
  return run_prep_rows<32, 32>(p, st, tag);
}

#define FGC_LG_DISPATCH(FN, ...)                                              \
  do {                                                                        \
    const int O = 2 * M;                                                      \
    if (lpr == 16) {                                                          \
      if (O <= 16) return FN<16, 16>(__VA_ARGS__);                            \
      return FN<32, 16>(__VA_ARGS__);                                         \
    }                                                                         \
    return FN<32, 32>(__VA_ARGS__);                                           \
  } while (0)

bool logits_fast_supported(int Cin, int Ca0, int Ca, int M) { return fast_lpr(Cin, Ca0, Ca, M) != 0; }

int launch_assign_logits_fast(const float* x, const float* u, const float* v, const float* c, float* uvx,
                              int64_t rows, int Cin, int Ca0, int Ca, int M, cudaStream_t st, unsigned* maxbits) {
  const int lpr = fast_lpr(Cin, Ca0, Ca, M);
  LgParams p{x, u, v, c, uvx, nullptr, nullptr, rows, 0, Cin, Ca0, Ca, M, maxbits};
  FGC_LG_DISPATCH(run_assign, p, st);
}

int launch_logits_bwd_x_fast(const float* d_uvx, const float* u, const float* v, float* gx, int64_t rows, int Cin,
                             int Ca0, int Ca, int M, cudaStream_t st) {
  const int lpr = fast_lpr(Cin, Ca0, Ca, M);
  LgParams p{nullptr, u, v, nullptr, const_cast<float*>(d_uvx), gx, nullptr, rows, 0, Cin, Ca0, Ca, M, nullptr};
  FGC_LG_DISPATCH(run_bwd_x, p, st);
}

int launch_logits_bwd_p_fast(const float* x, const float* d_uvx, float* part, int64_t rows, int64_t rows_per_chunk,
                             int chunks, int Cin, int Ca0, int Ca, int M, cudaStream_t st) {
  const int lpr = fast_lpr(Cin, Ca0, Ca, M);
  LgParams p{x, nullptr, nullptr, nullptr, const_cast<float*>(d_uvx), nullptr, part, rows, rows_per_chunk, Cin, Ca0, Ca, M, nullptr};
  FGC_LG_DISPATCH(run_bwd_p, p, chunks, st);
}

}  // namespace fgc
