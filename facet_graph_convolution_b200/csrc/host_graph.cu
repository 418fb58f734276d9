// Host-side graph preprocessing: the sequential greedy pairing of the reference's graph coarsening
// (Code/lib/coarsening.py:135-194 `metis_one_level`).  The visiting order makes every decision depend on
// all earlier ones, so this runs on the CPU; it is O(nnz) instead of interpreted Python.  No device work.
#include <vector>

#include "common.cuh"

namespace {

template <typename T>
int64_t greedy_pairing(const int32_t* row, const int32_t* col, const float* val, int64_t nnz,
                       const int64_t* order, const float* weights, int32_t n, int32_t* cluster_id, T* total) {
  std::vector<int64_t> start(static_cast<size_t>(n), 0);
  std::vector<int32_t> len(static_cast<size_t>(n), 0);
  std::vector<uint8_t> paired(static_cast<size_t>(n), 0);
  for (int64_t e = 0; e < nnz; ++e) {
    if (len[row[e]]++ == 0) start[row[e]] = e;
  }
  T assoc = 0;
  int32_t next = 0;
  for (int32_t i = 0; i < n; ++i) {
    const int64_t t = order[i];
    if (paired[t]) continue;
    paired[t] = 1;
    T best = 0;
    int32_t mate = -1;
    const T wt = T(1) / static_cast<T>(weights[t]);
    for (int64_t e = start[t], e1 = start[t] + len[t]; e < e1; ++e) {
      const int32_t c = col[e];
      if (paired[c]) continue;  // scores 0, never beats `best`
      const T score = static_cast<T>(val[e]) * (wt + T(1) / static_cast<T>(weights[c]));
      if (score > best) best = score, mate = c;
    }
    cluster_id[t] = next;
    if (mate >= 0) cluster_id[mate] = next, paired[mate] = 1;
    assoc += best;
    ++next;
  }
  *total = assoc;
  return next;
}

}  // namespace

extern "C" int fgc_greedy_pairing(const int32_t* row, const int32_t* col, const float* val, int64_t nnz,
                                  const int64_t* order, int64_t n_order, const float* weights, int32_t n,
                                  int precision, int32_t* cluster_id, double* total_assoc, int32_t* n_clusters) {
  FGC_REQUIRE(row && col && val && order && weights && cluster_id && total_assoc && n_clusters && nnz > 0 && n > 0,
              "greedy_pairing: bad arguments");
  FGC_REQUIRE(precision == 32 || precision == 64, "greedy_pairing: precision must be 32 or 64");
  FGC_REQUIRE(n_order >= n, "greedy_pairing: visiting order has %lld entries for %d rows", (long long)n_order, n);
  for (int64_t e = 0; e < nnz; ++e) {
    FGC_REQUIRE(row[e] >= 0 && row[e] < n && col[e] >= 0 && col[e] < n && (e == 0 || row[e] >= row[e - 1]),
                "greedy_pairing: entry %lld is out of range or rows are not sorted", (long long)e);
  }
  for (int32_t i = 0; i < n; ++i)
    FGC_REQUIRE(order[i] >= 0 && order[i] < n, "greedy_pairing: order[%d] = %lld is not a row", i, (long long)order[i]);
  for (int32_t i = 0; i < n; ++i) cluster_id[i] = 0;
  if (precision == 32) {
    float t = 0.f;
    *n_clusters = static_cast<int32_t>(greedy_pairing<float>(row, col, val, nnz, order, weights, n, cluster_id, &t));
    *total_assoc = t;
  } else {
    double t = 0.0;
    *n_clusters = static_cast<int32_t>(greedy_pairing<double>(row, col, val, nnz, order, weights, n, cluster_id, &t));
    *total_assoc = t;
  }
  return FGC_OK;
}
