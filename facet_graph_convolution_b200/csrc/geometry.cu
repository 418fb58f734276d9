// Output normalisation (reference Code/utils.py:1700-1715), training loss (Code/train.py:1272-1294)
// and the normal-guided vertex position updates (Code/train.py:1467-1557, 1668-1798).
// HBM/latency-bound integer-index + fp32 work; reductions are two-stage with a fixed order.
#include "common.cuh"

namespace fgc {

constexpr int kRedThreads = 256;
constexpr int kRedBlocksMax = 1024;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  // fixed-order block reduction: warp shuffles, then warp 0 over the 8 warp sums
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in thread 0
}

static inline int red_blocks(int64_t n) {
  int64_t b = (n + kRedThreads - 1) / kRedThreads;
  if (b > kRedBlocksMax) b = kRedBlocksMax;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// out[slot] = sum of part[0..P)
__global__ void finalize_sum_kernel(const float* __restrict__ part, int P, float* __restrict__ out,
                                    int slot) {
  __shared__ float sh[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) a += part[i];
  const float r = block_sum(a, sh);
  if (threadIdx.x == 0) out[slot] = r;
}

// ------------------------------------------------------------------ normalizeTensor
__global__ void abs_sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ part) {
  __shared__ float sh[32];
  float a = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    a += fabsf(x[i]);
  const float r = block_sum(a, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// scal[0] = sum |x|
__global__ void normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                      int64_t rows, const float* __restrict__ scal) {
  const float eps = 1e-5f;
  const float mean = scal[0] / static_cast<float>(rows * 3);
  const float den = mean + eps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float a = x[3 * r] / den, b = x[3 * r + 1] / den, c = x[3 * r + 2] / den;
    const float nrm = sqrtf(eps + (a * a + b * b + c * c));
    const float inv = nrm > eps ? 1.f / (nrm + eps) : 0.f;
    y[3 * r] = a * inv, y[3 * r + 1] = b * inv, y[3 * r + 2] = c * inv;
  }
}

// gradient w.r.t. the rescaled rows xs = x/den; returns per-row g_xs
__device__ __forceinline__ void normalize_row_grad(const float* x, const float* gy, int64_t r,
                                                   float den, float (&xs)[3], float (&gxs)[3]) {
  const float eps = 1e-5f;
  xs[0] = x[3 * r] / den, xs[1] = x[3 * r + 1] / den, xs[2] = x[3 * r + 2] / den;
  const float nrm = sqrtf(eps + (xs[0] * xs[0] + xs[1] * xs[1] + xs[2] * xs[2]));
  if (nrm > eps) {
    const float inv = 1.f / (nrm + eps);
    const float dot = gy[3 * r] * xs[0] + gy[3 * r + 1] * xs[1] + gy[3 * r + 2] * xs[2];
    const float k = dot * inv * inv / nrm;
#pragma unroll
    for (int j = 0; j < 3; ++j) gxs[j] = inv * gy[3 * r + j] - k * xs[j];
  } else {
    gxs[0] = gxs[1] = gxs[2] = 0.f;
  }
}

__global__ void normalize_bwd_dot_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                         int64_t rows, const float* __restrict__ scal,
                                         float* __restrict__ part) {
  __shared__ float sh[32];
  const float den = scal[0] / static_cast<float>(rows * 3) + 1e-5f;
  float a = 0.f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float xs[3], gxs[3];
    normalize_row_grad(x, gy, r, den, xs, gxs);
    a += gxs[0] * x[3 * r] + gxs[1] * x[3 * r + 1] + gxs[2] * x[3 * r + 2];
  }
  const float rr = block_sum(a, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = rr;
}

// scal[0] = sum|x|, scal[1] = sum_k gxs_k x_k
__global__ void normalize_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                     float* __restrict__ gx, int64_t rows,
                                     const float* __restrict__ scal) {
  const float nall = static_cast<float>(rows * 3);
  const float den = scal[0] / nall + 1e-5f;
  const float s = 1.f / den;
  const float k = s * s * scal[1] / nall;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float xs[3], gxs[3];
    normalize_row_grad(x, gy, r, den, xs, gxs);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float xv = x[3 * r + j];
      const float sg = xv > 0.f ? 1.f : (xv < 0.f ? -1.f : 0.f);
      gx[3 * r + j] = s * gxs[j] - k * sg;
    }
  }
}

// ------------------------------------------------------------------ faceNormalsLoss
__device__ __forceinline__ bool loss_row(const float* fn, const float* gt, int64_t r, float* dot) {
  const float g0 = gt[3 * r], g1 = gt[3 * r + 1], g2 = gt[3 * r + 2];
  *dot = fn[3 * r] * g0 + fn[3 * r + 1] * g1 + fn[3 * r + 2] * g2;
  return (fabsf(g0) + fabsf(g1) + fabsf(g2)) <= 10e-4f;  // fake node
}

__global__ void loss_partial_kernel(const float* __restrict__ fn, const float* __restrict__ gt,
                                    int64_t rows, float* __restrict__ part_sum,
                                    float* __restrict__ part_cnt) {
  __shared__ float sh[32];
  const float lim = 0.9999999f;
  float a = 0.f, n = 0.f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float d;
    if (!loss_row(fn, gt, r, &d)) {
      a += 180.f * acosf(fminf(fmaxf(d, -lim), lim)) / 3.14159265358979323846f;
      n += 1.f;
    }
  }
  const float ra = block_sum(a, sh);
  const float rn = block_sum(n, sh);
  if (threadIdx.x == 0) part_sum[blockIdx.x] = ra, part_cnt[blockIdx.x] = rn;
}

// scal[0] = sum of angles, scal[1] = number of real rows
__global__ void loss_finish_kernel(const float* __restrict__ scal, float* __restrict__ loss) {
  loss[0] = scal[0] / scal[1];
}

__global__ void loss_grad_kernel(const float* __restrict__ fn, const float* __restrict__ gt,
                                 float* __restrict__ gfn, int64_t rows,
                                 const float* __restrict__ scal, float gscale) {
  const float lim = 0.9999999f;
  const float k = gscale * (180.f / 3.14159265358979323846f) / scal[1];
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float d;
    const bool fake = loss_row(fn, gt, r, &d);
    float g = 0.f;
    if (!fake && d > -lim && d < lim) g = -k * rsqrtf(1.f - d * d);
#pragma unroll
    for (int j = 0; j < 3; ++j) gfn[3 * r + j] = g * gt[3 * r + j];
  }
}

// ------------------------------------------------------------------ vertex update (edges)
__global__ void vertex_update_edges_kernel(const float* __restrict__ xin, float* __restrict__ xout,
                                           const float* __restrict__ normals,
                                           const int32_t* __restrict__ edge_map,
                                           const int32_t* __restrict__ v_edges, int64_t V, int64_t F,
                                           int64_t E, int max_edges, float lambda, int64_t v_begin = 0) {
  // vertices v_begin .. V - 1 (a rank of the sharded update sweeps its own range and reads every vertex of x_in)
  for (int64_t i = v_begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < V;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x0 = xin[3 * i], x1 = xin[3 * i + 1], x2 = xin[3 * i + 2];
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    const int32_t* ve = v_edges + i * max_edges;
    for (int s = 0; s < max_edges; ++s) {
      const int e = ve[s];
      if (e < 0 || e >= E) continue;  // padded slot: edge row 0 of the reference = zero normals
      const int4 em = *reinterpret_cast<const int4*>(edge_map + 4 * static_cast<int64_t>(e));
      const int vs[2] = {em.x, em.y};
      const int fs[2] = {em.z, em.w};
      float n[2][3];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const bool ok = fs[t] >= 0 && fs[t] < F;
#pragma unroll
        for (int j = 0; j < 3; ++j) n[t][j] = ok ? normals[3 * static_cast<int64_t>(fs[t]) + j] : 0.f;
      }
      float e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int64_t vw = vs[w];
        const float d0 = xin[3 * vw] - x0, d1 = xin[3 * vw + 1] - x1, d2 = xin[3 * vw + 2] - x2;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float dp = d0 * n[t][0] + d1 * n[t][1] + d2 * n[t][2];
          e0 += n[t][0] * dp, e1 += n[t][1] * dp, e2 += n[t][2] * dp;
        }
      }
      u0 += e0, u1 += e1, u2 += e2;
    }
    xout[3 * i] = x0 + lambda * u0;
    xout[3 * i + 1] = x1 + lambda * u1;
    xout[3 * i + 2] = x2 + lambda * u2;
  }
}

// ------------------------------------------------------------------ vertex update (multi-scale)
// centres[F_s][3]: face centres pooled `scale*steps` pairwise levels with avg_ignore_zeros
__global__ void face_centres_kernel(const float* __restrict__ x, const int32_t* __restrict__ faces,
                                    float* __restrict__ centres, int64_t Fs, int levels) {
  const int group = 1 << levels;  // fine faces per coarse face, <= 16
  for (int64_t F = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; F < Fs;
       F += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v[16][3];
    for (int g = 0; g < group; ++g) {
      const int64_t f = F * group + g;
      float c[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int vid = faces[3 * f + t];
        if (vid >= 0) {  // -1 => the reference's zero "fake vertex"
          c[0] += x[3 * static_cast<int64_t>(vid)];
          c[1] += x[3 * static_cast<int64_t>(vid) + 1];
          c[2] += x[3 * static_cast<int64_t>(vid) + 2];
        }
      }
      v[g][0] = c[0] / 3.f, v[g][1] = c[1] / 3.f, v[g][2] = c[2] / 3.f;
    }
    int n = group;
    for (int s = 0; s < levels; ++s) {
      n >>= 1;
      for (int p = 0; p < n; ++p) {
        const bool z0 = v[2 * p][0] == 0.f && v[2 * p][1] == 0.f && v[2 * p][2] == 0.f;
        const bool z1 = v[2 * p + 1][0] == 0.f && v[2 * p + 1][1] == 0.f && v[2 * p + 1][2] == 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float a = z0 ? v[2 * p + 1][j] : v[2 * p][j];
          const float b = z1 ? v[2 * p][j] : v[2 * p + 1][j];
          v[p][j] = (a + b) / 2.f;
        }
      }
    }
    centres[3 * F] = v[0][0], centres[3 * F + 1] = v[0][1], centres[3 * F + 2] = v[0][2];
  }
}

__global__ void vertex_update_ms_kernel(const float* __restrict__ xin, float* __restrict__ xout,
                                        const float* __restrict__ normals,
                                        const float* __restrict__ centres,
                                        const int32_t* __restrict__ v_faces, int64_t V, int64_t Fs,
                                        int max_faces, int shift) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < V;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x0 = xin[3 * i], x1 = xin[3 * i + 1], x2 = xin[3 * i + 2];
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    int numf = 0;
    const int32_t* vf = v_faces + i * max_faces;
    for (int s = 0; s < max_faces; ++s) {
      const int f = vf[s];
      numf += (f != -1);
      if (f < 0) continue;            // floor(-1 / 4^scale) = -1 => the zero-normal fake face
      const int64_t Fc = f >> shift;  // floor division for non-negative ids
      if (Fc >= Fs) continue;
      const float n0 = normals[3 * Fc], n1 = normals[3 * Fc + 1], n2 = normals[3 * Fc + 2];
      const float e0 = centres[3 * Fc] - x0, e1 = centres[3 * Fc + 1] - x1, e2 = centres[3 * Fc + 2] - x2;
      const float w = n0 * e0 + n1 * e1 + n2 * e2;
      u0 += w * n0, u1 += w * n1, u2 += w * n2;
    }
    const float lam = 1.f / static_cast<float>(numf);  // inf for unreferenced vertices, as the reference
    xout[3 * i] = x0 + lam * u0;
    xout[3 * i + 1] = x1 + lam * u1;
    xout[3 * i + 2] = x2 + lam * u2;
  }
}

// ------------------------------------------------------------------ vertex update (multi-scale), backward
// One sweep is  x'_v = x_v + lam_v sum_k (n_k . e_k) n_k,  e_k = c_k - x_v  (k: the faces around v at this scale).
// With g = dL/dx', h = lam g, a_k = n_k . h, w_k = n_k . e_k:
//   dL/dx_v = g - sum_k a_k n_k        dL/dn_k += a_k e_k + w_k h        dL/dc_k += a_k n_k
// The per-slot terms are written out and summed per coarse face / per vertex over index lists in a
// fixed order (no float atomics): the gradient is bit-reproducible.
__global__ void ms_bwd_slots_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                    const float* __restrict__ normals, const float* __restrict__ centres,
                                    const int32_t* __restrict__ v_faces, int64_t V, int64_t Fs, int max_faces,
                                    int shift, float* __restrict__ gx, float* __restrict__ slot_gn,
                                    float* __restrict__ slot_gc) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < V;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x0 = x[3 * i], x1 = x[3 * i + 1], x2 = x[3 * i + 2];
    const float g0 = g[3 * i], g1 = g[3 * i + 1], g2 = g[3 * i + 2];
    const int32_t* vf = v_faces + i * max_faces;
    int numf = 0;
    for (int s = 0; s < max_faces; ++s) numf += (vf[s] != -1);
    const float lam = 1.f / static_cast<float>(numf);
    const float h0 = lam * g0, h1 = lam * g1, h2 = lam * g2;
    float d0 = g0, d1 = g1, d2 = g2;
    for (int s = 0; s < max_faces; ++s) {
      const int f = vf[s];
      const int64_t Fc = f >> shift;
      if (f < 0 || Fc >= Fs) continue;       // slots outside the index list: never read
      const float n0 = normals[3 * Fc], n1 = normals[3 * Fc + 1], n2 = normals[3 * Fc + 2];
      const float e0 = centres[3 * Fc] - x0, e1 = centres[3 * Fc + 1] - x1, e2 = centres[3 * Fc + 2] - x2;
      const float w = n0 * e0 + n1 * e1 + n2 * e2;
      const float a = n0 * h0 + n1 * h1 + n2 * h2;
      d0 -= a * n0, d1 -= a * n1, d2 -= a * n2;
      float* sn = slot_gn + (i * max_faces + s) * 3;
      float* sc = slot_gc + (i * max_faces + s) * 3;
      sn[0] = a * e0 + w * h0, sn[1] = a * e1 + w * h1, sn[2] = a * e2 + w * h2;
      sc[0] = a * n0, sc[1] = a * n1, sc[2] = a * n2;
    }
    gx[3 * i] = d0, gx[3 * i + 1] = d1, gx[3 * i + 2] = d2;
  }
}

// per coarse face: gn += its slots' normal terms; the centre gradient goes down the avg_ignore_zeros tree
// (the tf.where masks of Code/model.py:799-809 route it) to the fine faces' centres: gfine[f] = dL/d(centre of f)
__global__ void ms_bwd_faces_kernel(const float* __restrict__ x, const int32_t* __restrict__ faces,
                                    const int32_t* __restrict__ slot_ptr, const int32_t* __restrict__ slot_id,
                                    const float* __restrict__ slot_gn, const float* __restrict__ slot_gc,
                                    float* __restrict__ gn, float* __restrict__ gfine, int64_t Fs, int levels) {
  const int group = 1 << levels;
  for (int64_t F = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; F < Fs;
       F += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int32_t q = slot_ptr[F]; q < slot_ptr[F + 1]; ++q) {
      const int64_t sl = slot_id[q];
      a0 += slot_gn[3 * sl], a1 += slot_gn[3 * sl + 1], a2 += slot_gn[3 * sl + 2];
      c0 += slot_gc[3 * sl], c1 += slot_gc[3 * sl + 1], c2 += slot_gc[3 * sl + 2];
    }
    gn[3 * F] += a0, gn[3 * F + 1] += a1, gn[3 * F + 2] += a2;
    // forward tree (as face_centres_kernel), every level kept: node (level s, index p) at tree[off_s + p]
    float v[31][3];
    unsigned zmask = 0;   // bit (off + p): node is all-zero
    for (int gI = 0; gI < group; ++gI) {
      const int64_t f = F * group + gI;
      float c[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int vid = faces[3 * f + t];
        if (vid >= 0) {
          c[0] += x[3 * static_cast<int64_t>(vid)];
          c[1] += x[3 * static_cast<int64_t>(vid) + 1];
          c[2] += x[3 * static_cast<int64_t>(vid) + 2];
        }
      }
      v[gI][0] = c[0] / 3.f, v[gI][1] = c[1] / 3.f, v[gI][2] = c[2] / 3.f;
    }
    int off = 0, n = group;
    for (int s = 0; s < levels; ++s) {
      const int nn = n >> 1;
      for (int p = 0; p < nn; ++p) {
        const float* u0 = v[off + 2 * p];
        const float* u1 = v[off + 2 * p + 1];
        const bool z0 = u0[0] == 0.f && u0[1] == 0.f && u0[2] == 0.f;
        const bool z1 = u1[0] == 0.f && u1[1] == 0.f && u1[2] == 0.f;
        if (z0) zmask |= 1u << (off + 2 * p);
        if (z1) zmask |= 1u << (off + 2 * p + 1);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float a = z0 ? u1[j] : u0[j];
          const float b = z1 ? u0[j] : u1[j];
          v[off + n + p][j] = (a + b) / 2.f;
        }
      }
      off += n;
      n = nn;
    }
    // backward: gradients overwrite the node values, root first
    v[off][0] = c0, v[off][1] = c1, v[off][2] = c2;
    for (int s = levels - 1; s >= 0; --s) {
      const int nn = n;       // nodes of level s + 1
      n <<= 1;                // nodes of level s
      off -= n;
      for (int p = 0; p < nn; ++p) {
        const bool z0 = (zmask >> (off + 2 * p)) & 1u, z1 = (zmask >> (off + 2 * p + 1)) & 1u;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float gh = v[off + n + p][j] / 2.f;   // d mean / d cline0 = d mean / d cline1
          // cline0 = z0 ? line1 : line0, cline1 = z1 ? line0 : line1
          v[off + 2 * p][j] = (z0 ? 0.f : gh) + (z1 ? gh : 0.f);
          v[off + 2 * p + 1][j] = (z0 ? gh : 0.f) + (z1 ? 0.f : gh);
        }
      }
    }
    for (int gI = 0; gI < group; ++gI) {
      const int64_t f = F * group + gI;
      gfine[3 * f] = v[gI][0], gfine[3 * f + 1] = v[gI][1], gfine[3 * f + 2] = v[gI][2];
    }
  }
}

// per vertex: the centres of the faces it is a corner of (index list over `faces`, ascending corner id)
__global__ void ms_bwd_verts_kernel(const int32_t* __restrict__ vert_ptr, const int32_t* __restrict__ vert_corner,
                                    const float* __restrict__ gfine, float* __restrict__ gx, int64_t V) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < V;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int32_t q = vert_ptr[i]; q < vert_ptr[i + 1]; ++q) {
      const int64_t f = vert_corner[q] / 3;
      a0 += gfine[3 * f] / 3.f, a1 += gfine[3 * f + 1] / 3.f, a2 += gfine[3 * f + 2] / 3.f;
    }
    gx[3 * i] += a0, gx[3 * i + 1] += a1, gx[3 * i + 2] += a2;
  }
}

static inline unsigned vgrid(int64_t n) {
  int64_t b = (n + 127) / 128;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace fgc

using namespace fgc;

extern "C" {

// ---- several patches of one batch: segment b = rows [b * stride_rows, b * stride_rows + counts[b])
constexpr int kSegBlocks = 32;   // partial sums per segment

__global__ void seg_abs_sum_kernel(const float* __restrict__ x, int64_t stride_rows, const int32_t* __restrict__ counts,
                                   float* __restrict__ part) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const int64_t n = static_cast<int64_t>(counts[b]) * 3;
  const float* xb = x + static_cast<int64_t>(b) * stride_rows * 3;
  float a = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    a += fabsf(xb[i]);
  const float r = block_sum(a, sh);
  if (threadIdx.x == 0) part[b * kSegBlocks + blockIdx.x] = r;
}

__global__ void seg_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t stride_rows,
                                     const int32_t* __restrict__ counts, const float* __restrict__ part) {
  const float eps = 1e-5f;
  const int b = blockIdx.y;
  const int64_t rows = counts[b];
  float tot = 0.f;
  for (int i = 0; i < kSegBlocks; ++i) tot += part[b * kSegBlocks + i];   // fixed order, same in every thread
  const float den = tot / static_cast<float>(rows * 3) + eps;
  const float* xb = x + static_cast<int64_t>(b) * stride_rows * 3;
  float* yb = y + static_cast<int64_t>(b) * stride_rows * 3;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < stride_rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (r >= rows) {   // padding rows of the batch element
      yb[3 * r] = 0.f, yb[3 * r + 1] = 0.f, yb[3 * r + 2] = 0.f;
      continue;
    }
    const float a = xb[3 * r] / den, bb = xb[3 * r + 1] / den, c = xb[3 * r + 2] / den;
    const float nrm = sqrtf(eps + (a * a + bb * bb + c * c));
    const float inv = nrm > eps ? 1.f / (nrm + eps) : 0.f;
    yb[3 * r] = a * inv, yb[3 * r + 1] = bb * inv, yb[3 * r + 2] = c * inv;
  }
}

size_t fgc_normalize_workspace(int64_t rows) { return ws_bytes(2 * kRedBlocksMax + 8, 4) + 256; }

int fgc_normalize_rows(const float* x, float* y, int64_t rows, void* workspace,
                       size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x && y && rows > 0, "normalize_rows: bad arguments");
  cudaStream_t st = as_stream(stream);
  Workspace ws(workspace, workspace_bytes);
  float* part = ws.take<float>(2 * kRedBlocksMax);
  float* scal = ws.take<float>(8);
  FGC_REQUIRE(ws.ok(), "normalize_rows: workspace too small");
  const int nb = red_blocks(rows * 3);
  abs_sum_kernel<<<nb, kRedThreads, 0, st>>>(x, rows * 3, part);
  FGC_LAUNCHED("abs_sum_kernel");
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(part, nb, scal, 0);
  FGC_LAUNCHED("finalize_sum_kernel");
  normalize_rows_kernel<<<red_blocks(rows), kRedThreads, 0, st>>>(x, y, rows, scal);
  FGC_LAUNCHED("normalize_rows_kernel");
  return FGC_OK;
}

int fgc_normalize_rows_segmented(const float* x, float* y, int B, int64_t stride_rows, const int32_t* counts,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x && y && counts && B > 0 && B <= 65535 && stride_rows > 0, "normalize_rows_segmented: bad arguments");
  cudaStream_t st = as_stream(stream);
  Workspace ws(workspace, workspace_bytes);
  float* part = ws.take<float>(static_cast<size_t>(B) * kSegBlocks);
  FGC_REQUIRE(ws.ok(), "normalize_rows_segmented: workspace too small (%zu bytes needed)",
              static_cast<size_t>(B) * kSegBlocks * 4 + 512);
  seg_abs_sum_kernel<<<dim3(kSegBlocks, B), kRedThreads, 0, st>>>(x, stride_rows, counts, part);
  FGC_LAUNCHED("seg_abs_sum_kernel");
  int nb = red_blocks(stride_rows);
  if (nb > 64) nb = 64;
  seg_normalize_kernel<<<dim3(nb, B), kRedThreads, 0, st>>>(x, y, stride_rows, counts, part);
  FGC_LAUNCHED("seg_normalize_kernel");
  return FGC_OK;
}

int fgc_normalize_rows_bwd(const float* gy, const float* x, float* gx, int64_t rows, void* workspace,
                           size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(gy && x && gx && rows > 0, "normalize_rows_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  Workspace ws(workspace, workspace_bytes);
  float* part = ws.take<float>(2 * kRedBlocksMax);
  float* scal = ws.take<float>(8);
  FGC_REQUIRE(ws.ok(), "normalize_rows_bwd: workspace too small");
  const int nb = red_blocks(rows * 3);
  abs_sum_kernel<<<nb, kRedThreads, 0, st>>>(x, rows * 3, part);
  FGC_LAUNCHED("abs_sum_kernel");
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(part, nb, scal, 0);
  FGC_LAUNCHED("finalize_sum_kernel");
  const int nr = red_blocks(rows);
  normalize_bwd_dot_kernel<<<nr, kRedThreads, 0, st>>>(x, gy, rows, scal, part);
  FGC_LAUNCHED("normalize_bwd_dot_kernel");
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(part, nr, scal, 1);
  FGC_LAUNCHED("finalize_sum_kernel");
  normalize_bwd_kernel<<<nr, kRedThreads, 0, st>>>(x, gy, gx, rows, scal);
  FGC_LAUNCHED("normalize_bwd_kernel");
  return FGC_OK;
}

int fgc_face_normals_loss(const float* fn, const float* gt, float* loss, float* gfn, int64_t rows,
                          float gscale, void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(fn && gt && loss && rows > 0, "face_normals_loss: bad arguments");
  cudaStream_t st = as_stream(stream);
  Workspace ws(workspace, workspace_bytes);
  float* part = ws.take<float>(2 * kRedBlocksMax);
  float* scal = ws.take<float>(8);
  FGC_REQUIRE(ws.ok(), "face_normals_loss: workspace too small");
  const int nb = red_blocks(rows);
  loss_partial_kernel<<<nb, kRedThreads, 0, st>>>(fn, gt, rows, part, part + kRedBlocksMax);
  FGC_LAUNCHED("loss_partial_kernel");
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(part, nb, scal, 0);
  FGC_LAUNCHED("finalize_sum_kernel");
  finalize_sum_kernel<<<1, kRedThreads, 0, st>>>(part + kRedBlocksMax, nb, scal, 1);
  FGC_LAUNCHED("finalize_sum_kernel");
  loss_finish_kernel<<<1, 1, 0, st>>>(scal, loss);
  FGC_LAUNCHED("loss_finish_kernel");
  if (gfn) {
    loss_grad_kernel<<<nb, kRedThreads, 0, st>>>(fn, gt, gfn, rows, scal, gscale);
    FGC_LAUNCHED("loss_grad_kernel");
  }
  return FGC_OK;
}

size_t fgc_vertex_update_workspace(int64_t V) { return ws_bytes(static_cast<size_t>(V) * 3, 4) + 256; }

int fgc_vertex_update_edges(const float* x_in, float* x_out, const float* normals,
                            const int32_t* edge_map, const int32_t* v_edges, int64_t V, int64_t F,
                            int64_t E, int max_edges, int iters, float lambda, void* workspace,
                            size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x_in && x_out && normals && edge_map && v_edges && V > 0 && iters >= 0 && max_edges > 0,
              "vertex_update_edges: bad arguments");
  cudaStream_t st = as_stream(stream);
  Workspace ws(workspace, workspace_bytes);
  float* tmp = ws.take<float>(static_cast<size_t>(V) * 3);
  FGC_REQUIRE(ws.ok(), "vertex_update_edges: workspace too small");
  if (iters == 0) {
    FGC_CUDA(cudaMemcpyAsync(x_out, x_in, V * 12, cudaMemcpyDeviceToDevice, st));
    return FGC_OK;
  }
  // ping-pong so that the last sweep lands in x_out
  const float* src = x_in;
  for (int it = 0; it < iters; ++it) {
    float* dst = ((iters - 1 - it) % 2 == 0) ? x_out : tmp;
    vertex_update_edges_kernel<<<vgrid(V), 128, 0, st>>>(src, dst, normals, edge_map, v_edges, V, F, E,
                                                         max_edges, lambda);
    FGC_LAUNCHED("vertex_update_edges_kernel");
    src = dst;
  }
  return FGC_OK;
}

// One Jacobi sweep of update_position2 over the vertices [v_begin, v_end): reads all of x_in, writes rows
// v_begin .. v_end - 1 of x_out.  The building block of the sharded update (patches.vertex_update_edges_sharded:
// every rank sweeps its range, one all-gather per sweep), bit-identical to the same sweep of fgc_vertex_update_edges.
int fgc_vertex_update_edges_range(const float* x_in, float* x_out, const float* normals, const int32_t* edge_map,
                                  const int32_t* v_edges, int64_t V, int64_t F, int64_t E, int max_edges,
                                  int64_t v_begin, int64_t v_end, float lambda, void* stream) {
  FGC_REQUIRE(x_in && x_out && normals && edge_map && v_edges && V > 0 && max_edges > 0, "vertex_update_edges_range: bad arguments");
  FGC_REQUIRE(0 <= v_begin && v_begin <= v_end && v_end <= V, "vertex_update_edges_range: range [%lld, %lld) outside [0, %lld)",
              static_cast<long long>(v_begin), static_cast<long long>(v_end), static_cast<long long>(V));
  FGC_REQUIRE(x_in != x_out, "vertex_update_edges_range: a Jacobi sweep cannot run in place");
  if (v_begin == v_end) return FGC_OK;
  cudaStream_t st = as_stream(stream);
  vertex_update_edges_kernel<<<vgrid(v_end - v_begin), 128, 0, st>>>(x_in, x_out, normals, edge_map, v_edges, v_end, F, E,
                                                                      max_edges, lambda, v_begin);
  FGC_LAUNCHED("vertex_update_edges_kernel");
  return FGC_OK;
}

size_t fgc_vertex_update_ms_workspace(int64_t V, int64_t N0) {
  return ws_bytes(static_cast<size_t>(V) * 3, 4) + ws_bytes(static_cast<size_t>(N0) * 3, 4) + 512;
}

int fgc_vertex_update_ms(const float* x_in, float* x_out, const float* normals, const int32_t* faces,
                         const int32_t* v_faces, int64_t V, int64_t N0, int max_faces, int scale,
                         int steps, int iters, void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x_in && x_out && normals && faces && v_faces && V > 0 && N0 > 0 && iters >= 0 &&
                  scale >= 0 && steps > 0,
              "vertex_update_ms: bad arguments");
  const int levels = scale * steps;
  FGC_UNSUPPORTED(levels > 4, "vertex_update_ms: at most 16 fine faces per coarse face");
  FGC_REQUIRE(N0 % (1ll << levels) == 0, "vertex_update_ms: N0 not divisible by 2^(scale*steps)");
  cudaStream_t st = as_stream(stream);
  Workspace ws(workspace, workspace_bytes);
  float* tmp = ws.take<float>(static_cast<size_t>(V) * 3);
  float* centres = ws.take<float>(static_cast<size_t>(N0) * 3);
  FGC_REQUIRE(ws.ok(), "vertex_update_ms: workspace too small");
  const int64_t Fs = N0 >> levels;
  if (iters == 0) {
    FGC_CUDA(cudaMemcpyAsync(x_out, x_in, V * 12, cudaMemcpyDeviceToDevice, st));
    return FGC_OK;
  }
  const float* src = x_in;
  for (int it = 0; it < iters; ++it) {
    float* dst = ((iters - 1 - it) % 2 == 0) ? x_out : tmp;
    face_centres_kernel<<<vgrid(Fs), 128, 0, st>>>(src, faces, centres, Fs, levels);
    FGC_LAUNCHED("face_centres_kernel");
    vertex_update_ms_kernel<<<vgrid(V), 128, 0, st>>>(src, dst, normals, centres, v_faces, V, Fs,
                                                      max_faces, levels);
    FGC_LAUNCHED("vertex_update_ms_kernel");
    src = dst;
  }
  return FGC_OK;
}

size_t fgc_vertex_update_ms_bwd_workspace(int64_t V, int64_t N0, int max_faces, int iters) {
  const size_t v3 = static_cast<size_t>(V) * 3;
  return ws_bytes(v3 * static_cast<size_t>(iters > 0 ? iters : 1), 4) + 2 * ws_bytes(v3, 4) +
         2 * ws_bytes(v3 * static_cast<size_t>(max_faces), 4) + 2 * ws_bytes(static_cast<size_t>(N0) * 3, 4) + 1024;
}

int fgc_vertex_update_ms_bwd(const float* x_in, const float* normals, const int32_t* faces, const int32_t* v_faces,
                             int64_t V, int64_t N0, int max_faces, int scale, int steps, int iters,
                             const int32_t* slot_ptr, const int32_t* slot_id, const int32_t* vert_ptr,
                             const int32_t* vert_corner, const float* g_out, float* g_in, float* g_normals,
                             void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x_in && normals && faces && v_faces && slot_ptr && slot_id && vert_ptr && vert_corner && g_out && g_in &&
                  g_normals && V > 0 && N0 > 0 && iters >= 0 && scale >= 0 && steps > 0 && max_faces > 0,
              "vertex_update_ms_bwd: bad arguments");
  const int levels = scale * steps;
  FGC_UNSUPPORTED(levels > 4, "vertex_update_ms_bwd: at most 16 fine faces per coarse face");
  FGC_REQUIRE(N0 % (1ll << levels) == 0, "vertex_update_ms_bwd: N0 not divisible by 2^(scale*steps)");
  cudaStream_t st = as_stream(stream);
  const int64_t Fs = N0 >> levels;
  const size_t v3 = static_cast<size_t>(V) * 3;
  Workspace ws(workspace, workspace_bytes);
  float* traj = ws.take<float>(v3 * static_cast<size_t>(iters > 0 ? iters : 1));   // x before sweep t
  float* ga = ws.take<float>(v3);
  float* gb = ws.take<float>(v3);
  float* slot_gn = ws.take<float>(v3 * max_faces);
  float* slot_gc = ws.take<float>(v3 * max_faces);
  float* centres = ws.take<float>(static_cast<size_t>(N0) * 3);
  float* gfine = ws.take<float>(static_cast<size_t>(N0) * 3);
  FGC_REQUIRE(ws.ok(), "vertex_update_ms_bwd: workspace too small");
  FGC_CUDA(cudaMemsetAsync(g_normals, 0, static_cast<size_t>(Fs) * 12, st));
  if (iters == 0) {
    FGC_CUDA(cudaMemcpyAsync(g_in, g_out, v3 * 4, cudaMemcpyDeviceToDevice, st));
    return FGC_OK;
  }
  // the forward trajectory again (the same two kernels as fgc_vertex_update_ms: the same bits)
  FGC_CUDA(cudaMemcpyAsync(traj, x_in, v3 * 4, cudaMemcpyDeviceToDevice, st));
  for (int it = 0; it + 1 < iters; ++it) {
    face_centres_kernel<<<vgrid(Fs), 128, 0, st>>>(traj + it * v3, faces, centres, Fs, levels);
    FGC_LAUNCHED("face_centres_kernel");
    vertex_update_ms_kernel<<<vgrid(V), 128, 0, st>>>(traj + it * v3, traj + (it + 1) * v3, normals, centres, v_faces, V, Fs,
                                                      max_faces, levels);
    FGC_LAUNCHED("vertex_update_ms_kernel");
  }
  const float* g = g_out;
  for (int it = iters - 1; it >= 0; --it) {
    float* gnext = (it == 0) ? g_in : ((g == ga) ? gb : ga);
    const float* x = traj + it * v3;
    face_centres_kernel<<<vgrid(Fs), 128, 0, st>>>(x, faces, centres, Fs, levels);
    FGC_LAUNCHED("face_centres_kernel");
    ms_bwd_slots_kernel<<<vgrid(V), 128, 0, st>>>(x, g, normals, centres, v_faces, V, Fs, max_faces, levels, gnext, slot_gn,
                                                  slot_gc);
    FGC_LAUNCHED("ms_bwd_slots_kernel");
    ms_bwd_faces_kernel<<<vgrid(Fs), 128, 0, st>>>(x, faces, slot_ptr, slot_id, slot_gn, slot_gc, g_normals, gfine, Fs, levels);
    FGC_LAUNCHED("ms_bwd_faces_kernel");
    ms_bwd_verts_kernel<<<vgrid(V), 128, 0, st>>>(vert_ptr, vert_corner, gfine, gnext, V);
    FGC_LAUNCHED("ms_bwd_verts_kernel");
    g = gnext;
  }
  return FGC_OK;
}

}  // extern "C"
