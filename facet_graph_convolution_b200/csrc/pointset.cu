// Point-set losses of the vertex-space trainers (reference Code/train.py:1332-1464: accuracyLoss, fullLoss,
// sampledAccuracyLoss): nearest-neighbour distances between the predicted vertices P0 and the ground-truth
// vertices P1, thresholded, averaged, times 1000 — and the gradient with respect to P0.
//
// The reference materialises dist[batch, n0, n1] and reduces it twice.  Here nothing of size n0 x n1 exists:
//   1. nearest_kernel   — every query point scans its share of the candidate set from shared memory and
//                         joins the other shares with ONE 64-bit atomicMin of (bits(d^2) << 32 | index):
//                         the minimum is exact and ties go to the lowest index whatever the arrival order,
//                         so the result is run-to-run identical.
//   2. events_kernel    — per query: distance, threshold, loss contribution, the P0 row the gradient
//                         belongs to and the gradient vector itself.
//   3. sum_kernel       — fixed-order sum of the contributions of each term.
//   4. scatter_kernel   — per P0 row, the events that name it, added in event order (no float atomics).
#include "common.cuh"

namespace fgc {
namespace {

constexpr int kQ = 256;        // queries per block (one per thread)
constexpr int kTile = 1024;    // candidates staged per shared-memory tile

// sample ids outside [0, n) never read out of bounds through the bare C ABI (the Python front end raises before the call)
__device__ __forceinline__ int64_t clamp_row(int32_t id, int64_t n) {
  return id < 0 ? 0 : (id >= n ? n - 1 : static_cast<int64_t>(id));
}

// queries a[b][ia ? ia[i] : i], candidates c[b][ic ? ic[j] : j]; block (x: query group, y: candidate share, z: batch)
__global__ void __launch_bounds__(kQ)
nearest_kernel(const float* __restrict__ a, const int32_t* __restrict__ ia, int nq, int64_t stride_a,
               const float* __restrict__ c, const int32_t* __restrict__ ic, int nc, int64_t stride_c,
               int share, unsigned long long* __restrict__ packed) {
  __shared__ float4 tile[kTile];
  const int b = blockIdx.z;
  const int q = blockIdx.x * kQ + threadIdx.x;
  const float* ab = a + static_cast<int64_t>(b) * stride_a * 3;
  const float* cb = c + static_cast<int64_t>(b) * stride_c * 3;
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (q < nq) {
    const int64_t r = ia ? clamp_row(ia[q], stride_a) : q;
    ax = ab[r * 3]; ay = ab[r * 3 + 1]; az = ab[r * 3 + 2];
  }
  const int j0 = blockIdx.y * share, j1 = min(nc, j0 + share);
  float best = __int_as_float(0x7f800000);
  int bi = 0x7fffffff;
  for (int t0 = j0; t0 < j1; t0 += kTile) {
    const int n = min(kTile, j1 - t0);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += kQ) {
      const int64_t r = ic ? clamp_row(ic[t0 + t], stride_c) : (t0 + t);
      tile[t] = make_float4(cb[r * 3], cb[r * 3 + 1], cb[r * 3 + 2], 0.f);
    }
    __syncthreads();
#pragma unroll 8
    for (int t = 0; t < n; ++t) {
      const float4 p = tile[t];
      const float dx = ax - p.x, dy = ay - p.y, dz = az - p.z;
      // the reference's sum of squares, left to right, products rounded before the additions
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d2 < best) { best = d2; bi = t0 + t; }   // strict: the first minimum of the share stays
    }
  }
  if (q < nq && bi != 0x7fffffff) {
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(best)) << 32) | static_cast<unsigned>(bi);
    atomicMin(packed + static_cast<int64_t>(b) * nq + q, key);
  }
}

// One event per (batch, query).  `query_is_p0`: the query is a row of P0 (precision term: the gradient goes to the
// query's row, direction query - candidate); otherwise the candidate is (completeness term: gradient to the
// nearest P0 row, direction candidate - query).  thr < 0: no threshold.
__global__ void events_kernel(const unsigned long long* __restrict__ packed, int nq, int batch,
                              const float* __restrict__ a, const int32_t* __restrict__ ia, int64_t stride_a,
                              const float* __restrict__ c, const int32_t* __restrict__ ic, int64_t stride_c,
                              int query_is_p0, int64_t n0, float thr, float scale,
                              float* __restrict__ val, int32_t* __restrict__ row, float* __restrict__ gvec) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= static_cast<int64_t>(batch) * nq) return;
  const int b = static_cast<int>(e / nq), q = static_cast<int>(e % nq);
  const unsigned long long key = packed[e];
  const float d = sqrtf(__uint_as_float(static_cast<unsigned>(key >> 32)));
  const int j = static_cast<int>(key & 0xffffffffu);
  const bool kept = thr < 0.f || d <= thr;
  val[e] = kept ? d : 0.f;
  const int64_t ra = ia ? clamp_row(ia[q], stride_a) : q, rc = ic ? clamp_row(ic[j], stride_c) : j;
  const float* pa = a + (static_cast<int64_t>(b) * stride_a + ra) * 3;
  const float* pc = c + (static_cast<int64_t>(b) * stride_c + rc) * 3;
  // d|x - y|/dx = (x - y)/|x - y|; at distance 0 the reference's quotient is 0/0, here the event carries no gradient
  const float s = (kept && d > 0.f) ? scale / d : 0.f;
  const float sx = query_is_p0 ? s : -s;
  gvec[e * 3 + 0] = sx * (pa[0] - pc[0]);
  gvec[e * 3 + 1] = sx * (pa[1] - pc[1]);
  gvec[e * 3 + 2] = sx * (pa[2] - pc[2]);
  row[e] = (s != 0.f) ? static_cast<int32_t>(static_cast<int64_t>(b) * n0 + (query_is_p0 ? ra : rc)) : -1;
}

// fixed-order sum of val[0..n) by one block: thread t adds elements t, t + 1024, ...; tree over the threads
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ val, int64_t n, float mul, float* __restrict__ out, int accumulate) {
  __shared__ float sh[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += val[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + mul * sh[0];
}

// gp0[r] = sum over the events with row == r, in event order
__global__ void __launch_bounds__(256)
scatter_kernel(const int32_t* __restrict__ row, const float* __restrict__ gvec, int64_t nev, int64_t rows, float* __restrict__ gp0) {
  __shared__ int32_t srow[2048];
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int64_t e0 = 0; e0 < nev; e0 += 2048) {
    const int n = (nev - e0 < 2048) ? static_cast<int>(nev - e0) : 2048;
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += 256) srow[t] = row[e0 + t];
    __syncthreads();
    for (int t = 0; t < n; ++t) {
      if (srow[t] == r) {
        gx += gvec[(e0 + t) * 3]; gy += gvec[(e0 + t) * 3 + 1]; gz += gvec[(e0 + t) * 3 + 2];
      }
    }
  }
  if (r < rows) { gp0[r * 3] = gx; gp0[r * 3 + 1] = gy; gp0[r * 3 + 2] = gz; }
}

struct Term {
  const float* a; const int32_t* ia; int nq; int64_t stride_a;     // queries
  const float* c; const int32_t* ic; int nc; int64_t stride_c;     // candidates
  int query_is_p0; float thr;
};

}  // namespace
}  // namespace fgc

using namespace fgc;

extern "C" {

size_t fgc_point_set_loss_workspace(int batch, int64_t n0, int64_t n1, int ns0, int ns1) {
  // events: the precision term has ns0 queries per batch element (n0 without a sample), the completeness term n1
  // (accuracyLoss) or ns1 (fullLoss): sized for the larger
  const size_t q0 = static_cast<size_t>(ns0 > 0 ? ns0 : n0), q1 = static_cast<size_t>(ns1 > n1 ? ns1 : n1);
  const size_t ev = static_cast<size_t>(batch) * (q0 + q1);
  return ws_bytes(ev, 8) + ws_bytes(ev, 4) * 2 + ws_bytes(ev * 3, 4) + 1024;
}

int fgc_point_set_loss(const float* p0, const float* p1, int batch, int64_t n0, int64_t n1, const int32_t* ind0, int ns0,
                       const int32_t* ind1, int ns1, int mode, float* loss, float* gp0, void* workspace,
                       size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(p0 && p1 && loss && batch > 0 && n0 > 0 && n1 > 0, "point_set_loss: bad arguments");
  FGC_REQUIRE(n0 < (1ll << 31) / batch && n1 < (1ll << 31) / batch, "point_set_loss: more than 2^31 points");
  FGC_REQUIRE(mode == 0 || mode == 1, "point_set_loss: mode %d (0 = accuracyLoss, 1 = fullLoss)", mode);
  if (!ind0) ns0 = static_cast<int>(n0);
  if (!ind1) ns1 = static_cast<int>(n1);
  FGC_REQUIRE(ns0 > 0 && ns1 > 0, "point_set_loss: empty sample");
  cudaStream_t st = as_stream(stream);
  // precision: sampled P0 against all of P1 (train.py:1355 / :1408); thresholds :1334 / :1375-1376
  // completeness: accuracyLoss — every P1 point against the SAMPLED P0 (:1357, dist is [batch, ns0, n1]);
  //               fullLoss    — the sampled P1 points against ALL of P0 (:1410, dist1 is [batch, n0, ns1])
  Term terms[2];
  terms[0] = {p0, ind0, ns0, n0, p1, nullptr, static_cast<int>(n1), n1, 1, mode == 0 ? 5.f : 5000.f};
  if (mode == 0) terms[1] = {p1, nullptr, static_cast<int>(n1), n1, p0, ind0, ns0, n0, 0, -1.f};
  else           terms[1] = {p1, ind1, ns1, n1, p0, nullptr, static_cast<int>(n0), n0, 0, 5000.f};
  const int64_t ev0 = static_cast<int64_t>(batch) * terms[0].nq, ev1 = static_cast<int64_t>(batch) * terms[1].nq;
  const int64_t nev = ev0 + ev1;
  Workspace ws(workspace, workspace_bytes);
  unsigned long long* packed = ws.take<unsigned long long>(nev);
  float* val = ws.take<float>(nev);
  int32_t* row = ws.take<int32_t>(nev);
  float* gvec = ws.take<float>(nev * 3);
  FGC_REQUIRE(ws.ok(), "point_set_loss: workspace too small");
  FGC_CUDA(cudaMemsetAsync(packed, 0xff, nev * 8, st));
  const int sms = num_sms();
  int64_t off = 0;
  for (int t = 0; t < 2; ++t) {
    const Term& T = terms[t];
    const int qb = (T.nq + kQ - 1) / kQ;
    // enough candidate shares to fill the machine twice, none shorter than 128 candidates (a 500-sample query set over
    // the ~10 k vertices of a mesh was 22 CTAs with whole-tile shares: ncu, profiles/r6_pointset_summary.md)
    constexpr int kMinShare = 128;
    int shares = max(1, min((T.nc + kMinShare - 1) / kMinShare, (2 * sms + qb * batch - 1) / (qb * batch)));
    int share = ((T.nc + shares - 1) / shares + kMinShare - 1) / kMinShare * kMinShare;
    shares = (T.nc + share - 1) / share;
    nearest_kernel<<<dim3(qb, shares, batch), kQ, 0, st>>>(T.a, T.ia, T.nq, T.stride_a, T.c, T.ic, T.nc, T.stride_c, share,
                                                          packed + off);
    FGC_LAUNCHED("nearest_kernel");
    const int64_t ev = static_cast<int64_t>(batch) * T.nq;
    const float mean = 1000.f / static_cast<float>(ev);   // 1000 * reduce_mean (train.py:1368 / :1422)
    events_kernel<<<static_cast<unsigned>((ev + 255) / 256), 256, 0, st>>>(packed + off, T.nq, batch, T.a, T.ia, T.stride_a, T.c, T.ic,
                                                                          T.stride_c, T.query_is_p0, n0, T.thr, mean, val + off,
                                                                          row + off, gvec + off * 3);
    FGC_LAUNCHED("events_kernel");
    sum_kernel<<<1, 1024, 0, st>>>(val + off, ev, mean, loss, t);
    FGC_LAUNCHED("sum_kernel");
    off += ev;
  }
  if (gp0) {
    const int64_t rows = static_cast<int64_t>(batch) * n0;
    scatter_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, st>>>(row, gvec, nev, rows, gp0);
    FGC_LAUNCHED("scatter_kernel");
  }
  return FGC_OK;
}

}  // extern "C"
