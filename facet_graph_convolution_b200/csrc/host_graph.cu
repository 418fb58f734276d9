// Host-side graph preprocessing: the sequential greedy pairing of the reference's graph coarsening
// (Code/lib/coarsening.py:135-194 `metis_one_level`) and the breadth-first patch growth of its patch
// extraction (Code/utils.py:1508-1696 `getGraphPatch_wMask`).  In both the visiting order makes every
// decision depend on all earlier ones, so they run on the CPU, O(edges) instead of interpreted Python.
// No device work.
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

// Breadth-first growth of one patch over the facet graph adj[n][K] (1-based, 0 = end of row, column 0 the
// node itself).  Nodes get patch-local ids in discovery order; neighbours already claimed by earlier patches
// (mask != 0) join as context but are only expanded once the unclaimed region is exhausted and the patch is
// still smaller than min_patch.  adj_out[capacity][K] receives the patch-local lists (1-based, 0-padded):
// rows expanded during growth keep their neighbours in the original columns, rows finished afterwards list
// only neighbours inside the patch, compacted.  *next_seed = an unclaimed node seen just outside, or -1.
extern "C" int fgc_grow_patch(const int32_t* adj, int64_t n, int K, int64_t nodes_num, int64_t seed,
                              const uint8_t* mask, int64_t min_patch, int32_t* adj_out, int64_t capacity,
                              int64_t* old_index, int64_t* patch_nodes, int64_t* next_seed) {
  FGC_REQUIRE(adj && mask && adj_out && old_index && patch_nodes && next_seed && n > 0 && K >= 2,
              "grow_patch: bad arguments");
  FGC_REQUIRE(seed >= 0 && seed < n, "grow_patch: seed %lld outside the graph", (long long)seed);
  const int64_t need = (nodes_num > min_patch ? nodes_num : min_patch) + K;
  FGC_REQUIRE(capacity >= need, "grow_patch: adj_out holds %lld rows, %lld needed", (long long)capacity, (long long)need);
  std::vector<int64_t> local(static_cast<size_t>(n), -1);
  std::vector<int64_t> inner, border;  // FIFO queues (head indices below)
  size_t ih = 0, bh = 0;
  int64_t count = 0;
  for (int64_t i = 0; i < capacity * K; ++i) adj_out[i] = 0;
  auto claim = [&](int64_t node) {
    local[node] = count;
    old_index[count] = node;
    ++count;
  };
  auto nb = [&](int64_t node, int k) -> int64_t {
    const int32_t id = adj[node * K + k];
    return (id >= 1 && id <= n) ? id - 1 : -1 - (id != 0);  // -1 end of row, -2 id outside the graph
  };
  // expand: every neighbour joins the patch; written at its original column
  auto expand = [&](int64_t node, bool sort_by_mask) -> int {
    const int64_t me = local[node];
    adj_out[me * K] = static_cast<int32_t>(me + 1);
    for (int k = 1; k < K; ++k) {
      const int64_t j = nb(node, k);
      if (j == -1) break;
      if (j < 0) return 1;
      if (local[j] < 0) {
        claim(j);
        if (sort_by_mask && mask[j]) border.push_back(j); else inner.push_back(j);
      }
      adj_out[me * K + k] = static_cast<int32_t>(local[j] + 1);
    }
    return 0;
  };
  claim(seed);
  inner.push_back(seed);
  int bad = 0;
  while (count < nodes_num && ih < inner.size() && !bad) bad = expand(inner[ih++], true);
  if (count < min_patch) {
    while (count < min_patch && bh < border.size() && !bad) bad = expand(border[bh++], false);
    while (count < min_patch && ih < inner.size() && !bad) bad = expand(inner[ih++], false);
  }
  FGC_REQUIRE(!bad, "grow_patch: adjacency holds an id outside 0..n");
  int64_t nxt = -1;
  auto finish = [&](int64_t node) -> int {
    const int64_t me = local[node];
    adj_out[me * K] = static_cast<int32_t>(me + 1);
    int col = 1;
    for (int k = 1; k < K; ++k) {
      const int64_t j = nb(node, k);
      if (j == -1) break;
      if (j < 0) return 1;
      if (local[j] < 0) {
        if (!mask[j]) nxt = j;
        continue;
      }
      adj_out[me * K + col++] = static_cast<int32_t>(local[j] + 1);
    }
    return 0;
  };
  while (ih < inner.size() && !bad) bad = finish(inner[ih++]);
  while (bh < border.size() && !bad) bad = finish(border[bh++]);
  FGC_REQUIRE(!bad, "grow_patch: adjacency holds an id outside 0..n");
  *patch_nodes = count;
  *next_seed = nxt;
  return FGC_OK;
}
