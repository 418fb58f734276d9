// GPU builders of the index tensors the hot path consumes: pure integer
// work, bit-exact against the reference's host loops.
//
//   fgc_build_faces_adj     getFacesLargeAdj   reference Code/utils.py:243-295
//                           getVerticesFaces   reference Code/utils.py:370-395
//   fgc_build_edge_maps     getEdgeMap         reference Code/utils.py:91-183
//
// getFacesLargeAdj appends, vertex by vertex in increasing vertex id and over the ordered pairs
// (vf1 < vf2) of the faces around the vertex (faces in increasing id), f2 to the row of f1 and then f1
// to the row of f2 while the row has room.  Seen from one face f that is: for each of its vertices in
// increasing vertex id, every other face around the vertex in increasing face id -- (a, p) pairs with
// a < p precede (p, b) pairs in the enumeration -- truncated to K - 1 appends.  Edge-adjacent faces
// therefore appear twice.  The builders below first make the vertex -> faces CSR (count, scan, fill,
// per-vertex sort) and then write every row independently: no atomics on the outputs, no order races.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "conv_common.cuh"
#include "conv_launch.cuh"

namespace fgc {

namespace {

__global__ void vf_count_kernel(const int32_t* __restrict__ faces, int64_t nf, int64_t nv, int32_t* __restrict__ cnt,
                                int32_t* __restrict__ bad) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * nf;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (faces[3 * (i / 3)] < 0) continue;   // fake node row (-1, -1, -1): skipped, as getVerticesFaces does
    const int v = faces[i];
    if (v < 0 || v >= nv) {
      atomicAdd(bad, 1);
      continue;
    }
    atomicAdd(cnt + v, 1);
  }
}

__global__ void vf_fill_kernel(const int32_t* __restrict__ faces, int64_t nf, int64_t nv, const int32_t* __restrict__ ptr,
                               int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < 3 * nf;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (faces[3 * (i / 3)] < 0) continue;
    const int v = faces[i];
    if (v < 0 || v >= nv) continue;
    list[ptr[v] + atomicAdd(cursor + v, 1)] = static_cast<int32_t>(i / 3);
  }
}

// lists are short (vertex valence): insertion sort per list restores the increasing order
__global__ void list_sort_kernel(const int32_t* __restrict__ ptr, int32_t* __restrict__ list, int64_t n) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < n;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = ptr[v], e = ptr[v + 1];
    for (int i = b + 1; i < e; ++i) {
      const int32_t x = list[i];
      int j = i - 1;
      while (j >= b && list[j] > x) {
        list[j + 1] = list[j];
        --j;
      }
      list[j + 1] = x;
    }
  }
}

__global__ void faces_adj_kernel(const int32_t* __restrict__ faces, const int32_t* __restrict__ ptr,
                                 const int32_t* __restrict__ list, int64_t nf, int64_t nv, int K,
                                 int32_t* __restrict__ adj) {
  for (int64_t f = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; f < nf;
       f += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int32_t* row = adj + f * K;
    row[0] = static_cast<int32_t>(f + 1);
    int n = 1;
    int v[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
    if (v[0] >= 0) {
      // the face's vertices in increasing vertex id (a repeated id is visited once)
      if (v[0] > v[1]) { const int t = v[0]; v[0] = v[1]; v[1] = t; }
      if (v[1] > v[2]) { const int t = v[1]; v[1] = v[2]; v[2] = t; }
      if (v[0] > v[1]) { const int t = v[0]; v[0] = v[1]; v[1] = t; }
      for (int c = 0; c < 3 && n < K; ++c) {
        if (v[c] < 0 || v[c] >= nv || (c > 0 && v[c] == v[c - 1])) continue;
        for (int i = ptr[v[c]]; i < ptr[v[c] + 1] && n < K; ++i) {
          const int32_t g = list[i];
          if (g != f) row[n++] = g + 1;
        }
      }
    }
    for (; n < K; ++n) row[n] = 0;
  }
}

__global__ void padded_lists_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ list, int64_t n, int width,
                                    int32_t* __restrict__ out, int32_t* __restrict__ overflow) {
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < n;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = ptr[v], d = ptr[v + 1] - b;
    if (d > width) atomicMax(overflow, d);
    for (int i = 0; i < width; ++i) out[v * width + i] = i < d ? list[b + i] : -1;
  }
}

// ---- edges
// half-edge h = 3 f + s, s = 0: (v1, v2), 1: (v1, v3), 2: (v2, v3)   (reference slot order)
__device__ __forceinline__ void half_edge(const int32_t* faces, int64_t h, int& a, int& b) {
  const int64_t f = h / 3;
  const int s = static_cast<int>(h % 3);
  a = faces[3 * f + (s == 2 ? 1 : 0)];
  b = faces[3 * f + (s == 0 ? 1 : 2)];
}

__global__ void edge_keys_kernel(const int32_t* __restrict__ faces, int64_t nf, int64_t nv, uint64_t* __restrict__ keys,
                                 int32_t* __restrict__ vals) {
  for (int64_t h = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; h < 3 * nf;
       h += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int a, b;
    half_edge(faces, h, a, b);
    const uint64_t lo = static_cast<uint64_t>(min(a, b)), hi = static_cast<uint64_t>(max(a, b));
    keys[h] = lo * static_cast<uint64_t>(nv) + hi;
    vals[h] = static_cast<int32_t>(h);
  }
}

// sorted by (key, h) [stable radix sort of h-ordered input]: segment head = first appearance of the edge,
// segment tail = its last later appearance.  first[h] = 1 marks the half-edges that create an edge.
__global__ void edge_mark_kernel(const uint64_t* __restrict__ skeys, const int32_t* __restrict__ svals, int64_t n,
                                 int32_t* __restrict__ first, int32_t* __restrict__ last_of_first) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const bool head = i == 0 || skeys[i] != skeys[i - 1];
    if (!head) continue;
    int64_t j = i;
    while (j + 1 < n && skeys[j + 1] == skeys[i]) ++j;   // edges have few incident faces
    first[svals[i]] = 1;
    last_of_first[svals[i]] = (j > i) ? svals[j] : -1;
  }
}

__global__ void edge_rows_kernel(const int32_t* __restrict__ faces, int64_t nf, int64_t nv, const int32_t* __restrict__ first,
                                 const int32_t* __restrict__ eid, const int32_t* __restrict__ last_of_first,
                                 int32_t* __restrict__ e_map, int32_t* __restrict__ vcnt) {
  for (int64_t h = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; h < 3 * nf;
       h += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (!first[h]) continue;
    int a, b;
    half_edge(faces, h, a, b);
    int32_t* row = e_map + 4 * static_cast<int64_t>(eid[h]);
    row[0] = a, row[1] = b, row[2] = static_cast<int32_t>(h / 3);
    row[3] = last_of_first[h] >= 0 ? last_of_first[h] / 3 : -1;
    if (a >= 0 && a < nv) atomicAdd(vcnt + a, 1);
    if (b >= 0 && b < nv) atomicAdd(vcnt + b, 1);
  }
}

__global__ void ve_fill_kernel(const int32_t* __restrict__ e_map, int64_t E, int64_t nv, const int32_t* __restrict__ ptr,
                               int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < 2 * E;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = e_map[4 * (i >> 1) + (i & 1)];
    if (v < 0 || v >= nv) continue;
    list[ptr[v] + atomicAdd(cursor + v, 1)] = static_cast<int32_t>(i >> 1);
  }
}

// ---- per-face input features [normal | barycentre / bbox diagonal]  (reference Code/utils.py:63-68 with the
// two-pass normalize of :26-35, and :1264-1294).  The reference works in float64 NumPy; the arithmetic here is
// double too, rounded once to fp32 (the dtype the placeholders cast to, train.py:52-56).
__device__ __forceinline__ unsigned f2ord(float f) {   // order-preserving float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void bbox_kernel(const float* __restrict__ verts, int64_t nv, unsigned* __restrict__ mm) {
  unsigned lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nv;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const unsigned o = f2ord(verts[3 * v + j]);
      lo[j] = min(lo[j], o), hi[j] = max(hi[j], o);
    }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[j] = min(lo[j], __shfl_xor_sync(0xffffffffu, lo[j], o));
      hi[j] = max(hi[j], __shfl_xor_sync(0xffffffffu, hi[j], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(mm + j, lo[j]);
      atomicMax(mm + 3 + j, hi[j]);
    }
  }
}

__global__ void face_features_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, int64_t nf,
                                     int64_t nv, const unsigned* __restrict__ mm, int normalize, float* __restrict__ out) {
  double diag = 1.0;
  if (normalize) {
    double d2 = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double e = static_cast<double>(ord2f(mm[3 + j])) - static_cast<double>(ord2f(mm[j]));
      d2 += e * e;
    }
    diag = sqrt(d2);
  }
  for (int64_t f = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; f < nf;
       f += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float* o = out + 6 * f;
    const int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
    if (a < 0 || a >= nv || b < 0 || b >= nv || c < 0 || c >= nv) {   // fake node: zero features
#pragma unroll
      for (int j = 0; j < 6; ++j) o[j] = 0.f;
      continue;
    }
    double p[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) p[0][j] = verts[3ll * a + j], p[1][j] = verts[3ll * b + j], p[2][j] = verts[3ll * c + j];
    const double e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    const double e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
    double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {   // normalize(): a * (1 / (|a| + 1e-8)), applied twice
      const double inv = 1.0 / (sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 0.00000001);
      n[0] *= inv, n[1] *= inv, n[2] *= inv;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      o[j] = static_cast<float>(n[j]);
      o[3 + j] = static_cast<float>((p[0][j] / diag + p[1][j] / diag + p[2][j] / diag) / 3);
    }
  }
}

unsigned grid_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 32;
  if (b > cap) b = cap;
  return static_cast<unsigned>(b < 1 ? 1 : b);
}

size_t scan_bytes_for(int64_t n) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, static_cast<int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                static_cast<int>(n));
  return align_up(b, 256) + 256;
}

size_t sort_bytes_for(int64_t n) {
  size_t b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b, static_cast<uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                  static_cast<int32_t*>(nullptr), static_cast<int32_t*>(nullptr), static_cast<int>(n));
  return align_up(b, 256) + 256;
}

}  // namespace
}  // namespace fgc

using namespace fgc;

extern "C" {

size_t fgc_faces_adj_workspace(int64_t nf, int64_t nv) {
  if (nf <= 0 || nv <= 0) return 0;
  return ws_bytes(nv + 1, 4) * 3 + ws_bytes(3 * nf, 4) + scan_bytes_for(nv + 1) + 1024;
}

int fgc_build_faces_adj(const int32_t* faces, int64_t nf, int64_t nv, int K, int32_t* adj, int32_t* v_faces,
                        int kv, void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(faces && nf > 0 && nv > 0 && 3 * nf < (1ll << 31), "build_faces_adj: bad arguments");
  FGC_REQUIRE(adj == nullptr || (K >= 1 && K <= 1024), "build_faces_adj: K=%d outside 1..1024", K);
  FGC_REQUIRE(v_faces == nullptr || kv >= 1, "build_faces_adj: kv must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Workspace ws(workspace, workspace_bytes);
  int32_t* cnt = ws.take<int32_t>(nv + 1);
  int32_t* ptr = ws.take<int32_t>(nv + 1);
  int32_t* cursor = ws.take<int32_t>(nv + 1);   // cursor[nv]: bad-index / overflow flag
  int32_t* list = ws.take<int32_t>(3 * nf);
  const size_t sb = scan_bytes_for(nv + 1);
  char* scan_tmp = ws.take<char>(sb);
  FGC_REQUIRE(ws.ok(), "build_faces_adj: workspace too small (%zu bytes given)", workspace_bytes);
  FGC_CUDA(cudaMemsetAsync(cnt, 0, (nv + 1) * 4, st));
  FGC_CUDA(cudaMemsetAsync(cursor, 0, (nv + 1) * 4, st));
  int32_t* flag = cursor + nv;
  vf_count_kernel<<<grid_for(3 * nf), 256, 0, st>>>(faces, nf, nv, cnt, flag);
  FGC_LAUNCHED("vf_count_kernel");
  size_t sbytes = sb;
  FGC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, sbytes, cnt, ptr, static_cast<int>(nv + 1), st));
  vf_fill_kernel<<<grid_for(3 * nf), 256, 0, st>>>(faces, nf, nv, ptr, cursor, list);
  FGC_LAUNCHED("vf_fill_kernel");
  list_sort_kernel<<<grid_for(nv), 256, 0, st>>>(ptr, list, nv);
  FGC_LAUNCHED("list_sort_kernel");
  if (adj != nullptr) {
    faces_adj_kernel<<<grid_for(nf), 256, 0, st>>>(faces, ptr, list, nf, nv, K, adj);
    FGC_LAUNCHED("faces_adj_kernel");
  }
  int32_t h_flag = 0;
  FGC_CUDA(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st));
  FGC_CUDA(cudaStreamSynchronize(st));
  FGC_REQUIRE(h_flag == 0, "build_faces_adj: %d vertex ids outside 0..nv-1", h_flag);
  if (v_faces != nullptr) {
    FGC_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    padded_lists_kernel<<<grid_for(nv), 256, 0, st>>>(ptr, list, nv, kv, v_faces, flag);
    FGC_LAUNCHED("padded_lists_kernel");
    FGC_CUDA(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st));
    FGC_CUDA(cudaStreamSynchronize(st));
    FGC_REQUIRE(h_flag == 0, "build_faces_adj: a vertex has %d faces, more than kv=%d", h_flag, kv);
  }
  return FGC_OK;
}

int fgc_face_features(const float* verts, const int32_t* faces, int64_t nf, int64_t nv, int normalize,
                      float* features, void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(verts && faces && features && nf > 0 && nv > 0, "face_features: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Workspace ws(workspace, workspace_bytes);
  unsigned* mm = ws.take<unsigned>(8);
  FGC_REQUIRE(ws.ok(), "face_features: workspace too small (needs 1 KB)");
  FGC_CUDA(cudaMemsetAsync(mm, 0xFF, 3 * sizeof(unsigned), st));
  FGC_CUDA(cudaMemsetAsync(mm + 3, 0, 3 * sizeof(unsigned), st));
  if (normalize) {
    bbox_kernel<<<grid_for(nv), 256, 0, st>>>(verts, nv, mm);
    FGC_LAUNCHED("bbox_kernel");
  }
  face_features_kernel<<<grid_for(nf), 256, 0, st>>>(verts, faces, nf, nv, mm, normalize, features);
  FGC_LAUNCHED("face_features_kernel");
  return FGC_OK;
}

size_t fgc_edge_maps_workspace(int64_t nf, int64_t nv) {
  if (nf <= 0 || nv <= 0) return 0;
  const int64_t n = 3 * nf;
  return 2 * ws_bytes(n, 8) + 5 * ws_bytes(n, 4) + 3 * ws_bytes(nv + 1, 4) + sort_bytes_for(n) + scan_bytes_for(n + 1) +
         scan_bytes_for(nv + 1) + 2048;
}

int fgc_build_edge_maps(const int32_t* faces, int64_t nf, int64_t nv, int max_edges, int32_t* e_map,
                        int64_t* num_edges, int32_t* v_edges, void* workspace, size_t workspace_bytes,
                        void* stream) {
  FGC_REQUIRE(faces && e_map && num_edges && nf > 0 && nv > 0 && 3 * nf < (1ll << 31) - 1,
              "build_edge_maps: bad arguments");
  FGC_REQUIRE(v_edges == nullptr || max_edges >= 1, "build_edge_maps: max_edges must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n = 3 * nf;
  Workspace ws(workspace, workspace_bytes);
  uint64_t* keys = ws.take<uint64_t>(n);
  uint64_t* skeys = ws.take<uint64_t>(n);
  int32_t* vals = ws.take<int32_t>(n);
  int32_t* svals = ws.take<int32_t>(n);
  int32_t* first = ws.take<int32_t>(n + 1);
  int32_t* eid = ws.take<int32_t>(n + 1);
  int32_t* lof = ws.take<int32_t>(n);
  int32_t* vcnt = ws.take<int32_t>(nv + 1);
  int32_t* vptr = ws.take<int32_t>(nv + 1);
  int32_t* cursor = ws.take<int32_t>(nv + 1);
  const size_t sortb = sort_bytes_for(n), scanb = scan_bytes_for(n + 1), scanv = scan_bytes_for(nv + 1);
  char* sort_tmp = ws.take<char>(sortb);
  char* scan_tmp = ws.take<char>(scanb);
  char* scanv_tmp = ws.take<char>(scanv);
  FGC_REQUIRE(ws.ok(), "build_edge_maps: workspace too small (%zu bytes given)", workspace_bytes);
  edge_keys_kernel<<<grid_for(n), 256, 0, st>>>(faces, nf, nv, keys, vals);
  FGC_LAUNCHED("edge_keys_kernel");
  size_t b = sortb;
  FGC_CUDA(cub::DeviceRadixSort::SortPairs(sort_tmp, b, keys, skeys, vals, svals, static_cast<int>(n), 0, 64, st));
  FGC_CUDA(cudaMemsetAsync(first, 0, (n + 1) * 4, st));
  edge_mark_kernel<<<grid_for(n), 256, 0, st>>>(skeys, svals, n, first, lof);
  FGC_LAUNCHED("edge_mark_kernel");
  b = scanb;
  FGC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, b, first, eid, static_cast<int>(n + 1), st));   // edge id = rank by first appearance
  int32_t E = 0;
  FGC_CUDA(cudaMemcpyAsync(&E, eid + n, 4, cudaMemcpyDeviceToHost, st));
  FGC_CUDA(cudaMemsetAsync(vcnt, 0, (nv + 1) * 4, st));
  FGC_CUDA(cudaMemsetAsync(cursor, 0, (nv + 1) * 4, st));
  edge_rows_kernel<<<grid_for(n), 256, 0, st>>>(faces, nf, nv, first, eid, lof, e_map, vcnt);
  FGC_LAUNCHED("edge_rows_kernel");
  FGC_CUDA(cudaStreamSynchronize(st));
  *num_edges = E;
  if (v_edges != nullptr) {
    // per-vertex edge lists in increasing edge id; the first 3 nf ints of the (now free) key buffer hold them
    int32_t* list = reinterpret_cast<int32_t*>(keys);
    b = scanv;
    FGC_CUDA(cub::DeviceScan::ExclusiveSum(scanv_tmp, b, vcnt, vptr, static_cast<int>(nv + 1), st));
    ve_fill_kernel<<<grid_for(2 * static_cast<int64_t>(E)), 256, 0, st>>>(e_map, E, nv, vptr, cursor, list);
    FGC_LAUNCHED("ve_fill_kernel");
    list_sort_kernel<<<grid_for(nv), 256, 0, st>>>(vptr, list, nv);
    FGC_LAUNCHED("list_sort_kernel");
    int32_t* flag = cursor + nv;
    padded_lists_kernel<<<grid_for(nv), 256, 0, st>>>(vptr, list, nv, max_edges, v_edges, flag);
    FGC_LAUNCHED("padded_lists_kernel");
    int32_t h_flag = 0;
    FGC_CUDA(cudaMemcpyAsync(&h_flag, flag, 4, cudaMemcpyDeviceToHost, st));
    FGC_CUDA(cudaStreamSynchronize(st));
    FGC_REQUIRE(h_flag == 0, "build_edge_maps: a vertex has %d edges, more than max_edges=%d", h_flag, max_edges);
  }
  return FGC_OK;
}

}  // extern "C"
