// Pooling / unpooling / pointwise / permutation kernels around the facet-graph convolution
// (reference Code/model.py:779-830, tf.concat at :909,:929, host fancy-indexing at
// Code/dataClasses.py:142 and Code/train.py:117-121).  All HBM-bound streaming kernels:
// grid-stride, coalesced along the channel axis.
#include "common.cuh"

namespace fgc {

static inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 16) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(num_sms()) * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

#define FGC_GRID_STRIDE(i, n)                                                         \
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < (n); \
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)

__global__ void pool_max_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_out,
                                int C, int group) {
  FGC_GRID_STRIDE(i, n_out) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i % C);
    const float* src = x + (r * group) * C + c;
    float m = src[0];
    for (int g = 1; g < group; ++g) m = fmaxf(m, src[static_cast<int64_t>(g) * C]);
    y[i] = m;
  }
}

__global__ void pool_max_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                    const float* __restrict__ y, float* __restrict__ gx,
                                    int64_t n_out, int C, int group) {
  FGC_GRID_STRIDE(i, n_out) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i % C);
    const float* src = x + (r * group) * C + c;
    const float m = y[i];
    int ties = 0;
    for (int g = 0; g < group; ++g) ties += (src[static_cast<int64_t>(g) * C] == m);
    const float share = gy[i] / static_cast<float>(ties > 0 ? ties : 1);
    float* dst = gx + (r * group) * C + c;
    for (int g = 0; g < group; ++g)
      dst[static_cast<int64_t>(g) * C] = (src[static_cast<int64_t>(g) * C] == m) ? share : 0.f;
  }
}

// avg_ignore_zeros, `steps` pairwise levels (reference Code/model.py:792-814).  One thread per
// output row; a row counts as zero when all its C channels are exactly 0.
constexpr int kAizMaxC = 8;
constexpr int kAizMaxG = 16;
__global__ void pool_aiz_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows_out,
                                int C, int steps) {
  const int group = 1 << steps;
  FGC_GRID_STRIDE(r, rows_out) {
    float v[kAizMaxG][kAizMaxC];
    for (int g = 0; g < group; ++g)
      for (int c = 0; c < C; ++c) v[g][c] = x[(r * group + g) * C + c];
    int n = group;
    for (int s = 0; s < steps; ++s) {
      n >>= 1;
      for (int p = 0; p < n; ++p) {
        bool z0 = true, z1 = true;
        for (int c = 0; c < C; ++c) {
          z0 = z0 && (v[2 * p][c] == 0.f);
          z1 = z1 && (v[2 * p + 1][c] == 0.f);
        }
        for (int c = 0; c < C; ++c) {
          const float a = z0 ? v[2 * p + 1][c] : v[2 * p][c];
          const float b = z1 ? v[2 * p][c] : v[2 * p + 1][c];
          v[p][c] = (a + b) / 2.f;
        }
      }
    }
    for (int c = 0; c < C; ++c) y[r * C + c] = v[0][c];
  }
}

__global__ void upsample_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_out,
                                int C, int group) {
  FGC_GRID_STRIDE(i, n_out) {
    const int64_t r = i / C;
    y[i] = x[(r / group) * C + i % C];
  }
}

__global__ void upsample_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                    int64_t n_in, int C, int group) {
  FGC_GRID_STRIDE(i, n_in) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i % C);
    float a = 0.f;
    for (int g = 0; g < group; ++g) a += gy[(r * group + g) * C + c];
    gx[i] = a;
  }
}

__global__ void lrelu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n,
                             float alpha) {
  FGC_GRID_STRIDE(i, n) y[i] = lrelu_f(x[i], alpha);
}

__global__ void lrelu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ xpre,
                                 float* __restrict__ gx, int64_t n, float alpha) {
  FGC_GRID_STRIDE(i, n) {
    const float xv = xpre[i];
    gx[i] = gy[i] * (xv > 0.f ? 1.f : (xv < 0.f ? alpha : 0.f));
  }
}

__global__ void concat2_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               float* __restrict__ y, int64_t n, int Ca, int Cb) {
  const int C = Ca + Cb;
  FGC_GRID_STRIDE(i, n) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i % C);
    y[i] = c < Ca ? a[r * Ca + c] : b[r * Cb + c - Ca];
  }
}

__global__ void split2_kernel(const float* __restrict__ gy, float* __restrict__ ga,
                              float* __restrict__ gb, int64_t n, int Ca, int Cb) {
  const int C = Ca + Cb;
  FGC_GRID_STRIDE(i, n) {
    const int64_t r = i / C;
    const int c = static_cast<int>(i % C);
    if (c < Ca) ga[r * Ca + c] = gy[i];
    else gb[r * Cb + c - Ca] = gy[i];
  }
}

__global__ void gather_perm_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                   float* __restrict__ y, int64_t n, int C) {
  FGC_GRID_STRIDE(i, n) {
    const int64_t r = i / C;
    y[i] = x[static_cast<int64_t>(idx[r]) * C + i % C];
  }
}

}  // namespace fgc

using namespace fgc;

namespace fgc {
__global__ void push_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, const int64_t* __restrict__ ids,
                                 int64_t n, int C) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t off = ids[e / C] * C + e % C;
    dst[off] = src[off];
  }
}
}  // namespace fgc

extern "C" {

int fgc_pool_max(const float* x, float* y, int64_t rows_out, int C, int group, void* stream) {
  FGC_REQUIRE(x && y && rows_out >= 0 && C > 0 && group > 0, "pool_max: bad arguments");
  if (rows_out == 0) return FGC_OK;
  const int64_t n = rows_out * C;
  pool_max_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, y, n, C, group);
  FGC_LAUNCHED("pool_max_kernel");
  return FGC_OK;
}

int fgc_pool_max_bwd(const float* gy, const float* x, const float* y, float* gx, int64_t rows_out,
                     int C, int group, void* stream) {
  FGC_REQUIRE(gy && x && y && gx && C > 0 && group > 0, "pool_max_bwd: bad arguments");
  if (rows_out == 0) return FGC_OK;
  const int64_t n = rows_out * C;
  pool_max_bwd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(gy, x, y, gx, n, C, group);
  FGC_LAUNCHED("pool_max_bwd_kernel");
  return FGC_OK;
}

int fgc_pool_avg_ignore_zeros(const float* x, float* y, int B, int64_t rows_in, int C, int steps,
                              void* stream) {
  FGC_REQUIRE(x && y && B > 0 && C > 0 && steps >= 0, "pool_avg_ignore_zeros: bad arguments");
  FGC_UNSUPPORTED(C > kAizMaxC || (1 << steps) > kAizMaxG,
                  "pool_avg_ignore_zeros: C <= %d and 2^steps <= %d supported", kAizMaxC, kAizMaxG);
  FGC_REQUIRE(rows_in % (1 << steps) == 0, "pool_avg_ignore_zeros: rows not divisible by 2^steps");
  const int64_t rows_out = static_cast<int64_t>(B) * (rows_in >> steps);
  if (rows_out == 0) return FGC_OK;
  pool_aiz_kernel<<<grid_for(rows_out), 256, 0, as_stream(stream)>>>(x, y, rows_out, C, steps);
  FGC_LAUNCHED("pool_aiz_kernel");
  return FGC_OK;
}

int fgc_upsample(const float* x, float* y, int64_t rows_in, int C, int group, void* stream) {
  FGC_REQUIRE(x && y && C > 0 && group > 0, "upsample: bad arguments");
  const int64_t n = rows_in * group * C;
  if (n == 0) return FGC_OK;
  upsample_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, y, n, C, group);
  FGC_LAUNCHED("upsample_kernel");
  return FGC_OK;
}

int fgc_upsample_bwd(const float* gy, float* gx, int64_t rows_in, int C, int group, void* stream) {
  FGC_REQUIRE(gy && gx && C > 0 && group > 0, "upsample_bwd: bad arguments");
  const int64_t n = rows_in * C;
  if (n == 0) return FGC_OK;
  upsample_bwd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(gy, gx, n, C, group);
  FGC_LAUNCHED("upsample_bwd_kernel");
  return FGC_OK;
}

int fgc_lrelu(const float* x, float* y, int64_t n, float alpha, void* stream) {
  FGC_REQUIRE(x && y && n >= 0, "lrelu: bad arguments");
  if (n == 0) return FGC_OK;
  lrelu_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, y, n, alpha);
  FGC_LAUNCHED("lrelu_kernel");
  return FGC_OK;
}

int fgc_lrelu_bwd(const float* gy, const float* x_pre, float* gx, int64_t n, float alpha,
                  void* stream) {
  FGC_REQUIRE(gy && x_pre && gx && n >= 0, "lrelu_bwd: bad arguments");
  if (n == 0) return FGC_OK;
  lrelu_bwd_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(gy, x_pre, gx, n, alpha);
  FGC_LAUNCHED("lrelu_bwd_kernel");
  return FGC_OK;
}

int fgc_concat2(const float* a, const float* b, float* y, int64_t rows, int Ca, int Cb,
                void* stream) {
  FGC_REQUIRE(a && b && y && Ca > 0 && Cb > 0, "concat2: bad arguments");
  const int64_t n = rows * (Ca + Cb);
  if (n == 0) return FGC_OK;
  concat2_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(a, b, y, n, Ca, Cb);
  FGC_LAUNCHED("concat2_kernel");
  return FGC_OK;
}

int fgc_split2(const float* gy, float* ga, float* gb, int64_t rows, int Ca, int Cb, void* stream) {
  FGC_REQUIRE(gy && ga && gb && Ca > 0 && Cb > 0, "split2: bad arguments");
  const int64_t n = rows * (Ca + Cb);
  if (n == 0) return FGC_OK;
  split2_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(gy, ga, gb, n, Ca, Cb);
  FGC_LAUNCHED("split2_kernel");
  return FGC_OK;
}

// Rows ids[0..n) of src are stored at the same row indices of dst: dst is normally ANOTHER GPU's copy of the tensor
// (a peer-mapped pointer of a symmetric allocation), so these are plain stores over NVLink -- the exchange step of the
// vertex-sharded update (patches.vertex_update_edges_sharded, exchange = "p2p"): what a rank's peers read of its range
// after a sweep is pushed into their buffers, no collective call, no staging copy.
int fgc_push_rows(const float* src, float* dst, const int64_t* ids, int64_t n, int C, void* stream) {
  FGC_REQUIRE(src && dst && (ids || n == 0) && C > 0 && n >= 0, "push_rows: bad arguments");
  if (n == 0) return FGC_OK;
  push_rows_kernel<<<grid_for(n * C), 256, 0, as_stream(stream)>>>(src, dst, ids, n * C, C);
  FGC_LAUNCHED("push_rows_kernel");
  return FGC_OK;
}

int fgc_gather_perm(const float* x, const int32_t* idx, float* y, int64_t rows_out, int C,
                    void* stream) {
  FGC_REQUIRE(x && idx && y && C > 0, "gather_perm: bad arguments");
  const int64_t n = rows_out * C;
  if (n == 0) return FGC_OK;
  gather_perm_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, idx, y, n, C);
  FGC_LAUNCHED("gather_perm_kernel");
  return FGC_OK;
}

}  // extern "C"
