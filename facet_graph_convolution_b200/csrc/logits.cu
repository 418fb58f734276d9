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
#include <stdlib.h>
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
// Code/model.py:909,929, never materialised).  Per row it writes
//   * the fp16 hi|lo image  img[r][unit] = [fp16(x s) (64) | fp16(x s - hi) (64)],  s = 2^(126-E), E = exponent of
//     max(max|xa|, max|xb|) of the row's batch element (upper bounds the producing kernels leave behind; per element,
//     so that a batch of patches gives every patch exactly the result of running it alone);
//   * the assignment logits (reference Code/model.py:74-95: u.x + c and v.x) in the form conv_hm.cu consumes:
//     lg[r][0..15] = pairs {vl'[m], vl'[8]}, m = 0..7 (the table the consumer GATHERS: 64 bytes per row, kept
//     apart from the own-row table so that gathered rows cost half an L1 line), lg[rows + 1 + r][0..15] = pairs
//     {uo'[m], uo'[8]} (read once per facet, streaming), with
//     uo' = (u.x + c - max_m) log2(e), vl' = (v.x - max_m) log2(e): exp2(uo' + vl') needs no further max.
//     *flag is set when a row's logits spread over more than 60 binary orders (the consumer then re-centres).
// The 2M dot products per row run on the warp-level tensor path: 16 rows x Cin as A fragments (the hi/lo split the
// image needs anyway), [u;v] as fp16 hi|lo B fragments in shared memory, x_hi.w_hi + x_lo.w_hi + x_hi.w_lo in fp32.
// Row `rows` of both outputs is zeroed (what padding slots read).
struct PrepRowsParams {
  const float* xa;
  const float* xb;
  const float* u;
  const float* v;
  const float* c;
  const unsigned* maxa;
  const unsigned* maxb;
  uint4* img;
  float* lg;
  float* xunscale;
  unsigned* flag;
  int64_t rows;
  int lda, Ca, ldb, Cb, M, nunits;
  int Nimg;   // rows per batch element: maxa / maxb / xunscale are per element
};

__device__ __forceinline__ void prep_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void prep_hmma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kPrepWarps = 8;

// Output columns of the logit GEMM: 0..7 = uo[0..7], 8..15 = vl[0..7], 16 = uo[8], 17 = vl[8] (M = 9), rest zero.
// k order inside a 16-channel step: k = 2t, 2t+1 <-> channels 4t, 4t+1; k = 2t+8, 2t+9 <-> channels 4t+2, 4t+3
// (lane (g,t) then loads one float4 per row and step).
__global__ void __launch_bounds__(kPrepWarps * 32)
prep_rows_mma_kernel(const PrepRowsParams p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the convolution launched next may start its prologue early
  extern __shared__ uint4 wfrag[];             // [Cin/16][3][32]: {b0 hi, b1 hi, b0 lo, b1 lo} of (step, n-block, lane)
  __shared__ float red[kPrepWarps];
  __shared__ float wsc_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int Cin = p.Ca + p.Cb, nsteps = Cin >> 4, M = p.M;
  // ---- weight fragments (every block builds its own: 2M x Cin values, L2-resident)
  {
    float mx = 0.f;
    for (int e = threadIdx.x; e < M * Cin; e += blockDim.x) mx = fmaxf(mx, fmaxf(fabsf(__ldg(p.u + e)), fabsf(__ldg(p.v + e))));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m2 = 0.f;
      for (int i = 0; i < kPrepWarps; ++i) m2 = fmaxf(m2, red[i]);
      int E = (__float_as_int(m2) >> 23) & 0xFF;
      E = min(max(E, 16), 240);
      wsc_s = __int_as_float((253 - E) << 23);   // |w| * wsc in [0.5, 1)
    }
    __syncthreads();
    const float wsc = wsc_s;
    for (int e = threadIdx.x; e < nsteps * 3 * 32; e += blockDim.x) {
      const int ln = e & 31, nb = (e >> 5) % 3, ks = e / 96;
      const int gg = ln >> 2, tt = ln & 3;
      const int col = nb * 8 + gg;             // output column this lane's B fragment feeds
      const float* src = nullptr;
      if (col < 8) src = p.u + col * Cin;
      else if (col < 16) src = p.v + (col - 8) * Cin;
      else if (col == 16 && M == 9) src = p.u + 8 * Cin;
      else if (col == 17 && M == 9) src = p.v + 8 * Cin;
      float w4[4] = {0.f, 0.f, 0.f, 0.f};
      if (src != nullptr && (col & 7) < M) {
#pragma unroll
        for (int j = 0; j < 4; ++j) w4[j] = __ldg(src + ks * 16 + 4 * tt + j) * wsc;
      }
      uint4 f;
      prep_split(w4[0], w4[1], f.x, f.z);
      prep_split(w4[2], w4[3], f.y, f.w);
      wfrag[e] = f;
    }
    __syncthreads();
  }
  // the parameters above are older than the stream's previous kernel; the rows and their max|x| are what it wrote
  asm volatile("griddepcontrol.wait;" ::: "memory");
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
    if (threadIdx.x < 16) p.lg[p.rows * 16 + threadIdx.x] = 0.f, p.lg[(2 * p.rows + 1) * 16 + threadIdx.x] = 0.f;
  }
  const float wun = 1.f / wsc_s;
  const float cu0 = (2 * t < M) ? __ldg(p.c + 2 * t) : 0.f, cu1 = (2 * t + 1 < M) ? __ldg(p.c + 2 * t + 1) : 0.f;
  const float c8 = (M == 9) ? __ldg(p.c + 8) : 0.f;
  constexpr float L2E = 1.4426950408889634f;
  bool spread = false;
  const int64_t nblk = (p.rows + 15) >> 4;
  for (int64_t blk = static_cast<int64_t>(blockIdx.x) * kPrepWarps + warp; blk < nblk;
       blk += static_cast<int64_t>(gridDim.x) * kPrepWarps) {
    const int64_t r0 = blk * 16 + g, r1 = r0 + 8;
    const bool ok0 = r0 < p.rows, ok1 = r1 < p.rows;
    const int e0 = ok0 ? scale_exp(static_cast<int>(r0 / p.Nimg)) : 127, e1 = ok1 ? scale_exp(static_cast<int>(r1 / p.Nimg)) : 127;
    const float sc0 = __int_as_float((253 - e0) << 23), sc1 = __int_as_float((253 - e1) << 23);
    float acc[3][4];
#pragma unroll
    for (int nb = 0; nb < 3; ++nb) acc[nb][0] = acc[nb][1] = acc[nb][2] = acc[nb][3] = 0.f;
    for (int u0 = 0; u0 < nsteps; u0 += 4) {          // one 64-channel unit (4 steps) at a time
      float4 xa0[4], xa1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = (u0 + j) * 16 + 4 * t;
        xa0[j] = make_float4(0.f, 0.f, 0.f, 0.f), xa1[j] = xa0[j];
        if (u0 + j < nsteps) {
          const float* s0 = ch < p.Ca ? p.xa + r0 * p.lda + ch : p.xb + r0 * p.ldb + (ch - p.Ca);
          const float* s1 = ch < p.Ca ? p.xa + r1 * p.lda + ch : p.xb + r1 * p.ldb + (ch - p.Ca);
          if (ok0) xa0[j] = __ldg(reinterpret_cast<const float4*>(s0));
          if (ok1) xa1[j] = __ldg(reinterpret_cast<const float4*>(s1));
        }
      }
      const int unit = u0 >> 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // a0 = row g (k 2t,2t+1), a1 = row g+8, a2 = row g (k 2t+8, 2t+9), a3 = row g+8
        uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
        prep_split(xa0[j].x * sc0, xa0[j].y * sc0, h0, l0);
        prep_split(xa1[j].x * sc1, xa1[j].y * sc1, h1, l1);
        prep_split(xa0[j].z * sc0, xa0[j].w * sc0, h2, l2);
        prep_split(xa1[j].z * sc1, xa1[j].w * sc1, h3, l3);
        // image: channels (u0+j)*16 + 4t .. +3 of this unit: 8 bytes per plane (written for padded channels too)
        const int cc = (j * 16 + 4 * t) * 2;   // byte offset inside the 128-byte plane
        if (ok0) {
          uint8_t* d = reinterpret_cast<uint8_t*>(p.img) + (r0 * p.nunits + unit) * 256 + cc;
          *reinterpret_cast<uint2*>(d) = make_uint2(h0, h2);
          *reinterpret_cast<uint2*>(d + 128) = make_uint2(l0, l2);
        }
        if (ok1) {
          uint8_t* d = reinterpret_cast<uint8_t*>(p.img) + (r1 * p.nunits + unit) * 256 + cc;
          *reinterpret_cast<uint2*>(d) = make_uint2(h1, h3);
          *reinterpret_cast<uint2*>(d + 128) = make_uint2(l1, l3);
        }
        if (u0 + j < nsteps) {
#pragma unroll
          for (int nb = 0; nb < 3; ++nb) {
            const uint4 f = wfrag[((u0 + j) * 3 + nb) * 32 + lane];
            prep_hmma(acc[nb], h0, h1, h2, h3, f.x, f.y);
            prep_hmma(acc[nb], l0, l1, l2, l3, f.x, f.y);
            prep_hmma(acc[nb], h0, h1, h2, h3, f.z, f.w);
          }
        }
      }
    }
    // ---- logits of rows r0 (c0,c1) and r1 (c2,c3): nb 0 = uo[2t,2t+1], nb 1 = vl[2t,2t+1], nb 2 (t = 0) = uo[8], vl[8]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = h ? r1 : r0;
      const bool ok = h ? ok1 : ok0;
      const float un = __int_as_float(((h ? e1 : e0) + 1) << 23) * wun;
      float u_a = fmaf(acc[0][2 * h], un, cu0), u_b = fmaf(acc[0][2 * h + 1], un, cu1);
      float v_a = acc[1][2 * h] * un, v_b = acc[1][2 * h + 1] * un;
      float u8 = fmaf(__shfl_sync(0xffffffffu, acc[2][2 * h], lane & ~3), un, c8);
      float v8 = __shfl_sync(0xffffffffu, acc[2][2 * h + 1], lane & ~3) * un;
      float mu = (2 * t + 1 < M) ? fmaxf(u_a, u_b) : ((2 * t < M) ? u_a : -INFINITY);
      float mv = (2 * t + 1 < M) ? fmaxf(v_a, v_b) : ((2 * t < M) ? v_a : -INFINITY);
      mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, 1)), mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, 2));
      mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, 1)), mv = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, 2));
      if (M == 9) mu = fmaxf(mu, u8), mv = fmaxf(mv, v8);
      u_a = (u_a - mu) * L2E, u_b = (u_b - mu) * L2E, u8 = (M == 9) ? (u8 - mu) * L2E : 0.f;
      v_a = (v_a - mv) * L2E, v_b = (v_b - mv) * L2E, v8 = (M == 9) ? (v8 - mv) * L2E : 0.f;
      if (2 * t >= M) u_a = 0.f, v_a = 0.f;
      if (2 * t + 1 >= M) u_b = 0.f, v_b = 0.f;
      if (ok) {
        spread |= (fminf(fminf(u_a, u_b), u8) < -60.f) || (fminf(fminf(v_a, v_b), v8) < -60.f);
        reinterpret_cast<float4*>(p.lg + r * 16)[t] = make_float4(v_a, v8, v_b, v8);   // pairs {vl'[2t], vl'[8]}, {vl'[2t+1], vl'[8]}
        reinterpret_cast<float4*>(p.lg + (p.rows + 1 + r) * 16)[t] = make_float4(u_a, u8, u_b, u8);
      }
    }
  }
  if (__any_sync(0xffffffffu, spread) && lane == 0) atomicOr(p.flag, 1u);
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

// rows of [xa (Ca channels, row stride lda) | xb (Cb, ldb; may be null)] -> lg[2][rows + 1][16], img[rows + 1][nunits][16];
// *flag (zeroed by the caller) receives 1 when the consumer must re-centre its softmax
bool prep_rows_supported(int Ca, int Cb, int M) {
  const int Cin = Ca + Cb;
  return (M == 8 || M == 9) && Ca % 16 == 0 && Cb % 16 == 0 && Cin >= 16 && Cin <= 128 && (Cin == 32 || Cin % 64 == 0);
}
int launch_prep_rows(const float* xa, int lda, int Ca, const float* xb, int ldb, int Cb, const float* u, const float* v,
                     const float* c, int M, int64_t rows, int Nimg, const unsigned* maxa, const unsigned* maxb, void* img,
                     float* lg, float* xunscale, unsigned* flag, cudaStream_t st, const char* tag) {
  const int Cin = Ca + Cb, nunits = (Cin + 63) / 64;
  FGC_REQUIRE(prep_rows_supported(Ca, Cb, M) && lda % 4 == 0 && (xb == nullptr || ldb % 4 == 0),
              "prep_rows: unsupported shape (Ca=%d Cb=%d M=%d)", Ca, Cb, M);
  FGC_REQUIRE(Nimg > 0 && rows % Nimg == 0, "prep_rows: rows must be a multiple of the rows per batch element");
  PrepRowsParams p{xa, xb, u, v, c, maxa, maxb, static_cast<uint4*>(img), lg, xunscale, flag, rows, lda, Ca, ldb, Cb, M, nunits, Nimg};
  const int64_t nblk = (rows + 15) / 16;
  int64_t blocks = (nblk + kPrepWarps - 1) / kPrepWarps;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = static_cast<size_t>(Cin / 16) * 3 * 32 * 16;
  {
    // programmatic stream serialization: the weight-fragment prologue above runs under the previous kernel's tail when that
    // kernel releases its dependents early (conv_hm2_kernel does); griddepcontrol.wait guards the first dependent read
    static const bool pdl = getenv("FGC_DISABLE_PDL") == nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(blocks)), cfg.blockDim = dim3(kPrepWarps * 32), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr, cfg.numAttrs = 1;
    FGC_CUDA(cudaLaunchKernelEx(&cfg, prep_rows_mma_kernel, p));
  }
  FGC_LAUNCHED(tag != nullptr ? tag : "prep_rows_kernel");
  return FGC_OK;
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
