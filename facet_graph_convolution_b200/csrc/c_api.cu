// extern "C" boundary of libfacetconv_b200.so (declared in include/facetconv_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "conv_launch.cuh"

namespace fgc {

std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------ opt-in kernel profiler
bool g_prof_on = false;
namespace {
struct ProfMark {
  const char* name;
  cudaEvent_t ev;
};
std::vector<ProfMark> g_marks;
std::vector<cudaEvent_t> g_pool;
size_t g_pool_next = 0;
cudaStream_t g_prof_stream = nullptr;
cudaEvent_t prof_event() {
  if (g_pool_next == g_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    g_pool.push_back(e);
  }
  return g_pool[g_pool_next++];
}
}  // namespace
void prof_set_stream(cudaStream_t st) { g_prof_stream = st; }
void prof_mark(const char* name) {
  cudaEvent_t e = prof_event();
  if (!e) return;
  cudaEventRecord(e, g_prof_stream);
  g_marks.push_back({name, e});
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static int check_shape(const fgc_conv_shape* s, const char* who) {
  FGC_REQUIRE(s != nullptr, "%s: shape is NULL", who);
  FGC_REQUIRE(s->B > 0 && s->N > 0, "%s: B and N must be positive (B=%d N=%d)", who, s->B, s->N);
  FGC_REQUIRE(s->K > 0 && s->K <= FGC_MAX_K, "%s: K=%d outside 1..%d", who, s->K, FGC_MAX_K);
  FGC_REQUIRE(s->M > 0 && s->M <= FGC_MAX_M, "%s: M=%d outside 1..%d", who, s->M, FGC_MAX_M);
  FGC_REQUIRE(s->Cin > 0 && s->Cin <= FGC_MAX_C, "%s: Cin=%d outside 1..%d", who, s->Cin, FGC_MAX_C);
  FGC_REQUIRE(s->Cout > 0 && s->Cout <= FGC_MAX_C, "%s: Cout=%d outside 1..%d", who, s->Cout, FGC_MAX_C);
  FGC_REQUIRE(s->Cw > 0 && s->Cw <= s->Cin, "%s: Cw=%d outside 1..Cin", who, s->Cw);
  FGC_REQUIRE(s->Ca > 0 && s->Ca0 >= 0 && s->Ca0 + s->Ca <= s->Cin,
              "%s: logit window [%d,%d) outside the row", who, s->Ca0, s->Ca0 + s->Ca);
  FGC_REQUIRE(static_cast<int64_t>(s->B) * s->N * s->K < (1ll << 31), "%s: B*N*K must be < 2^31", who);
  return FGC_OK;
}

static bool use_tc(const fgc_conv_shape* s) {
  static const bool disabled = getenv("FGC_DISABLE_TC") != nullptr;
  if (disabled || s->Cin % 4 != 0 || !conv_fwd_tc_supported(s->Cw, s->Cout, s->M, s->K)) return false;
  // a layer that needs several channel-block launches only pays off with enough 64-facet tiles to fill
  // the SMs (measured: 1 250 rows x 8 launches 0.28 ms vs 0.18 ms on the FFMA path; 5 000 rows x 4 launches
  // 0.16 vs 0.30 ms)
  const int launches = ((s->Cw + 63) / 64) * ((s->Cout + 31) / 32);
  return s->M == 8 || launches == 1 || static_cast<int64_t>(s->B) * s->N >= 4096;
}

static bool use_mma(const fgc_conv_shape* s) {
  static const bool disabled = getenv("FGC_DISABLE_MMA") != nullptr || getenv("FGC_DISABLE_TC") != nullptr;
  return !disabled && conv_mma_supported(s->Cin, s->Cw, s->Cout, s->M, s->K);
}

// HMMA-aggregation forward (conv_hm.cu): the dense layers without a tile plan, fused upsampling included
static bool use_hm(const fgc_conv_shape* s) {
  static const bool disabled = getenv("FGC_DISABLE_HM") != nullptr || getenv("FGC_DISABLE_TC") != nullptr;
  // plain feature assignment over the whole row (the network's layers; the position / window variants keep their kernels)
  return !disabled && conv_hm_supported(s->Cin, s->Cw, s->Cout, s->M, s->K) && s->Cw == s->Cin && s->Ca0 == 0 &&
         s->Ca == s->Cin && prep_rows_supported(s->Cin, 0, s->M) && static_cast<int64_t>(s->B) * s->N >= 64;
}

static size_t conv_fwd_workspace(const fgc_conv_shape* s) {
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  return (use_hm(s) ? conv_hm_workspace(rows, s->Cw, s->Cout, s->M, s->B) + 256 : 0) + ws_bytes((rows + 1) * 2 * s->M, 4) + ws_bytes(static_cast<size_t>(s->M) * s->Cout * s->Cw, 4) +
         ws_bytes(conv_fwd_tc_workspace(s->Cout, s->M, s->Cw), 1) + (use_mma(s) ? conv_mma_workspace(rows) + 256 : 0) + 512;
}

static int conv_fwd(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0,
                    const float* b, const float* u, const float* v, const float* c, float* y,
                    int bias_mask, int act, float alpha, void* workspace, size_t workspace_bytes,
                    cudaStream_t st, const void* plan = nullptr, int upshift = 0) {
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  if (upshift > 0) {
    // fused custom_upsampling: x holds (rows >> upshift) rows, row r of the layer reads row r >> upshift.
    // Only the tensor-core forwards without a tile plan index their gathers this way.
    FGC_UNSUPPORTED(!(use_hm(s) || (s->M == 9 && use_tc(s))) || plan != nullptr || upshift > 4 ||
                        (s->N & ((1 << upshift) - 1)),
                    "conv_fwd_up: shape has no fused-upsampling path");
    Workspace wsu(workspace, workspace_bytes);
    float* uvx_c = wsu.take<float>((rows + 1) * 2 * s->M);
    if (use_hm(s)) {
      const size_t hb = conv_hm_workspace(rows >> upshift, s->Cw, s->Cout, s->M, s->B);
      char* hws = wsu.take<char>(hb);
      FGC_REQUIRE(wsu.ok(), "conv_fwd_up: workspace too small");
      ConvFwdParams ph{x, adj, nullptr, nullptr, b, y, rows, s->N, s->K, s->Cin, s->Cw, s->Cout, s->M, bias_mask, act, alpha};
      return launch_conv_hm(ph, W0, u, v, c, hws, hb, st, upshift);
    }
    wsu.take<float>(static_cast<size_t>(s->M) * s->Cout * s->Cw);
    char* wimg_u = wsu.take<char>(conv_fwd_tc_workspace(s->Cout, s->M, s->Cw));
    FGC_REQUIRE(wsu.ok(), "conv_fwd_up: workspace too small");
    fgc_conv_shape sc = *s;
    sc.B = 1, sc.N = static_cast<int>(rows >> upshift);     // the logits pass runs over the coarse rows
    int rcu = launch_assign_logits(&sc, x, u, v, c, uvx_c, st);
    if (rcu) return rcu;
    ConvFwdParams pu{x, adj, uvx_c, nullptr, b, y, rows, s->N, s->K, s->Cin, s->Cw, s->Cout, s->M, bias_mask, act, alpha};
    return launch_conv_fwd_tc(pu, W0, wimg_u, st, upshift);
  }
  if (conv_fwd_small_supported(s))   // the 6 -> 32 input layer: thread per facet, logits inline, no workspace
    return launch_conv_fwd_small(s, x, adj, W0, b, u, v, c, y, bias_mask, act, alpha, st);
  Workspace ws(workspace, workspace_bytes);
  float* uvx = ws.take<float>((rows + 1) * 2 * s->M);   // one spare row: the HMMA path's zero row
  float* Wt = ws.take<float>(static_cast<size_t>(s->M) * s->Cout * s->Cw);
  FGC_REQUIRE(ws.ok(), "conv_fwd: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              conv_fwd_workspace(s));
  ConvFwdParams p{x, adj, uvx, Wt, b, y, rows, s->N, s->K, s->Cin, s->Cw, s->Cout, s->M,
                  bias_mask, act, alpha};
  int rc = FGC_OK;
  if (plan != nullptr && use_mma(s)) {
    char* wimg = ws.take<char>(conv_fwd_tc_workspace(s->Cout, s->M, s->Cw));
    char* img = ws.take<char>(conv_mma_workspace(rows));
    FGC_REQUIRE(ws.ok(), "conv_fwd: workspace too small for the planned tensor-core path");
    // max|x| for the image scale rides on the logits pass when that pass reads the whole row
    const bool fused = assign_logits_absmax_supported(s);
    if (fused) {
      rc = prep_image_reset(img, rows, st);
      if (rc) return rc;
    }
    rc = launch_assign_logits(s, x, u, v, c, uvx, st,
                              fused ? const_cast<unsigned*>(conv_mma_image_maxbits(img, rows)) : nullptr);
    if (rc) return rc;
    return launch_conv_mma(p, W0, plan, img, wimg, st, fused);
  }
  if (use_hm(s)) {
    const size_t hb = conv_hm_workspace(rows, s->Cw, s->Cout, s->M, s->B);
    char* hws = ws.take<char>(hb);
    FGC_REQUIRE(ws.ok(), "conv_fwd: workspace too small for the HMMA-aggregation path");
    return launch_conv_hm(p, W0, u, v, c, hws, hb, st);
  }
  rc = launch_assign_logits(s, x, u, v, c, uvx, st);
  if (rc) return rc;
  if (use_tc(s)) {
    char* wimg = ws.take<char>(conv_fwd_tc_workspace(s->Cout, s->M, s->Cw));
    FGC_REQUIRE(ws.ok(), "conv_fwd: workspace too small for the tensor-core path");
    return launch_conv_fwd_tc(p, W0, wimg, st);
  }
  rc = launch_transpose_w(W0, Wt, s->M, s->Cout, s->Cw, st);
  if (rc) return rc;
  return launch_conv_fwd(p, st);
}

// shared with conv_bwd.cu
int conv_fwd_saved_views(const fgc_conv_shape* s, const void* fwd_ws, size_t fwd_ws_bytes, FwdSaved* out) {
  // mirrors the take() order of conv_fwd's planned branch
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  FGC_REQUIRE(use_mma(s), "conv_bwd: a saved forward workspace needs the planned tensor-core path");
  Workspace ws(const_cast<void*>(fwd_ws), fwd_ws_bytes);
  out->uvx = ws.take<float>((rows + 1) * 2 * s->M);
  ws.take<float>(static_cast<size_t>(s->M) * s->Cout * s->Cw);
  ws.take<char>(conv_fwd_tc_workspace(s->Cout, s->M, s->Cw));
  out->ximg = ws.take<char>(conv_mma_workspace(rows));
  FGC_REQUIRE(ws.ok(), "conv_bwd: saved forward workspace too small (%zu bytes given, %zu needed)", fwd_ws_bytes,
              conv_fwd_workspace(s));
  return FGC_OK;
}

// ------------------------------------------------------------------ host-buffer path
struct HostCache {
  int device = -1;
  char* dev = nullptr;
  size_t cap = 0;
  cudaStream_t stream = nullptr;      // compute
  cudaStream_t s_in = nullptr;        // host -> device copies
  cudaStream_t s_out = nullptr;       // device -> host copies
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
static thread_local HostCache g_hc;

static void host_destroy_streams() {
  if (g_hc.stream) cudaStreamDestroy(g_hc.stream);
  if (g_hc.s_in) cudaStreamDestroy(g_hc.s_in);
  if (g_hc.s_out) cudaStreamDestroy(g_hc.s_out);
  for (auto& e : g_hc.ev)
    if (e) cudaEventDestroy(e);
}

static int host_reserve(int device, size_t bytes) {
  if (g_hc.device != device) {
    if (g_hc.dev) {
      cudaSetDevice(g_hc.device);
      cudaFree(g_hc.dev);
    }
    host_destroy_streams();
    g_hc = HostCache();
  }
  FGC_CUDA(cudaSetDevice(device));
  g_hc.device = device;
  if (!g_hc.stream) {
    FGC_CUDA(cudaStreamCreateWithFlags(&g_hc.stream, cudaStreamNonBlocking));
    FGC_CUDA(cudaStreamCreateWithFlags(&g_hc.s_in, cudaStreamNonBlocking));
    FGC_CUDA(cudaStreamCreateWithFlags(&g_hc.s_out, cudaStreamNonBlocking));
    for (auto& e : g_hc.ev) FGC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (bytes > g_hc.cap) {
    if (g_hc.dev) FGC_CUDA(cudaFree(g_hc.dev));
    g_hc.dev = nullptr;
    g_hc.cap = 0;
    FGC_CUDA(cudaMalloc(&g_hc.dev, bytes));
    g_hc.cap = bytes;
  }
  return FGC_OK;
}

// Every exit of a host-buffer entry point (error exits included) leaves no copy from / to the caller's buffers in
// flight on the cached streams, and the caller's current device as it was.
struct HostCallGuard {
  int prev = -1;
  HostCallGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      prev = -1;
      cudaGetLastError();
    }
  }
  ~HostCallGuard() {
    if (g_hc.s_in) cudaStreamSynchronize(g_hc.s_in);
    if (g_hc.stream) cudaStreamSynchronize(g_hc.stream);
    if (g_hc.s_out) cudaStreamSynchronize(g_hc.s_out);
    g_gx_ready_event = nullptr;
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_version(void) { return 102; }   // 102: point-set loss, backward of the multi-scale vertex update
const char* fgc_last_error(void) { return g_err; }
uint64_t fgc_launch_count(void) { return g_launches.load(); }

int fgc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fgc_profile_begin(void* stream) {
  g_marks.clear();
  g_pool_next = 0;
  g_prof_on = true;
  prof_set_stream(reinterpret_cast<cudaStream_t>(stream));
  prof_mark("<begin>");
  return FGC_OK;
}

int fgc_profile_end(char* buf, size_t buf_bytes) {
  g_prof_on = false;
  if (g_marks.empty()) return FGC_OK;
  FGC_CUDA(cudaEventSynchronize(g_marks.back().ev));
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (size_t i = 1; i < g_marks.size(); ++i) {
    float ms = 0.f;
    if (g_marks[i].name[0] == '<') continue;  // a re-begin marker
    FGC_CUDA(cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev));
    auto it = acc.find(g_marks[i].name);
    if (it == acc.end()) {
      order.push_back(g_marks[i].name);
      acc[g_marks[i].name] = {ms, 1};
    } else {
      it->second.first += ms;
      it->second.second += 1;
    }
  }
  size_t off = 0;
  if (buf && buf_bytes) buf[0] = 0;
  for (const auto& nm : order) {
    char line[256];
    const int n = snprintf(line, sizeof(line), "%s %.6f %d\n", nm.c_str(), acc[nm].first, acc[nm].second);
    if (buf && off + n + 1 < buf_bytes) {
      memcpy(buf + off, line, n + 1);
      off += n;
    }
  }
  g_marks.clear();
  g_pool_next = 0;
  return FGC_OK;
}

size_t fgc_conv_fwd_workspace(const fgc_conv_shape* s) {
  if (check_shape(s, "conv_fwd_workspace")) return 0;
  return conv_fwd_workspace(s);
}

size_t fgc_conv_bwd_workspace(const fgc_conv_shape* s) {
  if (check_shape(s, "conv_bwd_workspace")) return 0;
  return conv_bwd_workspace(s);
}

int fgc_conv_fwd(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0,
                 const float* b, const float* u, const float* v, const float* c, float* y,
                 int bias_mask, int act, float alpha, void* workspace, size_t workspace_bytes,
                 void* stream) {
  int rc = check_shape(s, "conv_fwd");
  if (rc) return rc;
  FGC_REQUIRE(x && adj && W0 && b && u && v && c && y, "conv_fwd: NULL tensor pointer");
  FGC_REQUIRE(act == FGC_ACT_NONE || act == FGC_ACT_LRELU, "conv_fwd: unknown activation %d", act);
  return conv_fwd(s, x, adj, W0, b, u, v, c, y, bias_mask, act, alpha, workspace, workspace_bytes,
                  as_stream(stream));
}

int fgc_conv_fwd_up_supported(const fgc_conv_shape* s, int upshift) {
  if (check_shape(s, "conv_fwd_up_supported")) return 0;
  return ((use_hm(s) || (s->M == 9 && use_tc(s))) && upshift > 0 && upshift <= 4 && (s->N & ((1 << upshift) - 1)) == 0) ? 1 : 0;
}

int fgc_conv_fwd_up(const fgc_conv_shape* s, const float* x_coarse, const int32_t* adj, const float* W0,
                    const float* b, const float* u, const float* v, const float* c, float* y, int bias_mask,
                    int act, float alpha, int upshift, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_shape(s, "conv_fwd_up");
  if (rc) return rc;
  FGC_REQUIRE(x_coarse && adj && W0 && b && u && v && c && y && upshift > 0, "conv_fwd_up: bad arguments");
  FGC_REQUIRE(act == FGC_ACT_NONE || act == FGC_ACT_LRELU, "conv_fwd_up: unknown activation %d", act);
  return conv_fwd(s, x_coarse, adj, W0, b, u, v, c, y, bias_mask, act, alpha, workspace, workspace_bytes,
                  as_stream(stream), nullptr, upshift);
}

int fgc_debug_trace(int64_t* out, int n) {
  if (getenv("FGC_HM_TRACE") != nullptr) return debug_hm_trace(out, n);   // conv_hm.cu's pipeline stamps
  return debug_mma_trace(out, n);
}

size_t fgc_conv_plan_bytes(int B, int N, int K, int M) {
  if (B <= 0 || N <= 0 || K <= 0 || K > FGC_MAX_K || M != 8) return 0;
  return conv_plan_bytes(static_cast<int64_t>(B) * N, K, M);
}

int fgc_build_conv_plan(const int32_t* adj, int B, int N, int K, int M, void* plan, size_t plan_bytes,
                        void* stream) {
  FGC_REQUIRE(adj && plan && B > 0 && N > 0 && K > 0 && K <= FGC_MAX_K, "build_conv_plan: bad arguments");
  return build_conv_plan(adj, B, N, K, M, plan, plan_bytes, as_stream(stream));
}

int fgc_conv_fwd_planned(const fgc_conv_shape* s, const float* x, const int32_t* adj, const void* plan,
                         const float* W0, const float* b, const float* u, const float* v, const float* c,
                         float* y, int bias_mask, int act, float alpha, void* workspace,
                         size_t workspace_bytes, void* stream) {
  int rc = check_shape(s, "conv_fwd_planned");
  if (rc) return rc;
  FGC_REQUIRE(x && adj && W0 && b && u && v && c && y, "conv_fwd_planned: NULL tensor pointer");
  FGC_REQUIRE(act == FGC_ACT_NONE || act == FGC_ACT_LRELU, "conv_fwd_planned: unknown activation %d", act);
  return conv_fwd(s, x, adj, W0, b, u, v, c, y, bias_mask, act, alpha, workspace, workspace_bytes,
                  as_stream(stream), plan);
}

size_t fgc_reverse_adj_workspace(int B, int N, int K) {
  return reverse_adj_workspace(static_cast<int64_t>(B) * N);
}

int fgc_build_reverse_adj(const int32_t* adj, int B, int N, int K, int32_t* rev_ptr, int32_t* rev_edge,
                          int64_t* nnz_out, void* workspace, size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(adj && rev_ptr && rev_edge && B > 0 && N > 0 && K > 0, "build_reverse_adj: bad arguments");
  return build_reverse_adj(adj, B, N, K, rev_ptr, rev_edge, nnz_out, workspace, workspace_bytes,
                           as_stream(stream));
}

int fgc_conv_bwd(const fgc_conv_shape* s, const float* gy, const float* x, const int32_t* adj,
                 const int32_t* rev_ptr, const int32_t* rev_edge, const float* W0, const float* u,
                 const float* v, const float* c, float* gx, float* gW0, float* gb, float* gu, float* gv,
                 float* gc, int bias_mask, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_shape(s, "conv_bwd");
  if (rc) return rc;
  FGC_REQUIRE(gy && x && adj && rev_ptr && rev_edge && W0 && u && v && c && gx && gW0 && gb && gu &&
                  gv && gc,
              "conv_bwd: NULL tensor pointer");
  return conv_bwd(s, gy, x, adj, rev_ptr, rev_edge, W0, u, v, c, gx, gW0, gb, gu, gv, gc, bias_mask,
                  workspace, workspace_bytes, as_stream(stream));
}

int fgc_build_reverse_padded(const int32_t* rev_ptr, const int32_t* rev_edge, int B, int N, int K, int Kr,
                             int32_t* radj, void* stream) {
  FGC_REQUIRE(rev_ptr && rev_edge && radj && B > 0 && N > 0 && K > 0 && Kr > 0 && Kr <= FGC_MAX_K,
              "build_reverse_padded: bad arguments");
  return launch_build_radj(rev_ptr, rev_edge, B, N, K, Kr, radj, as_stream(stream));
}

int fgc_conv_bwd_planned(const fgc_conv_shape* s, const float* gy, const float* x, const int32_t* adj,
                         const void* plan, const int32_t* rev_ptr, const int32_t* rev_edge, const int32_t* radj,
                         int Kr, const void* rplan, const float* W0, const float* u, const float* v, const float* c,
                         float* gx, float* gW0, float* gb, float* gu, float* gv, float* gc, int bias_mask,
                         const void* fwd_workspace, size_t fwd_workspace_bytes, void* workspace,
                         size_t workspace_bytes, void* stream) {
  int rc = check_shape(s, "conv_bwd_planned");
  if (rc) return rc;
  FGC_REQUIRE(gy && x && adj && rev_ptr && rev_edge && W0 && u && v && c && gx && gW0 && gb && gu && gv && gc,
              "conv_bwd_planned: NULL tensor pointer");
  FGC_REQUIRE((radj == nullptr) == (rplan == nullptr), "conv_bwd_planned: radj and rplan go together");
  FGC_REQUIRE(fwd_workspace == nullptr || plan != nullptr, "conv_bwd_planned: a saved forward workspace needs the plan");
  return conv_bwd(s, gy, x, adj, rev_ptr, rev_edge, W0, u, v, c, gx, gW0, gb, gu, gv, gc, bias_mask, workspace,
                  workspace_bytes, as_stream(stream), radj, Kr, rplan, plan, fwd_workspace, fwd_workspace_bytes);
}

int fgc_gather_rows(const float* x, const int32_t* adj, float* out, int B, int N, int K, int C,
                    void* stream) {
  FGC_REQUIRE(x && adj && out && B > 0 && N > 0 && K > 0 && C > 0, "gather_rows: bad arguments");
  return launch_gather_rows(x, adj, out, static_cast<int64_t>(B) * N, N, K, C, as_stream(stream));
}

int fgc_assignments(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* u,
                    const float* v, const float* c, float* q, void* workspace, size_t workspace_bytes,
                    void* stream) {
  int rc = check_shape(s, "assignments");
  if (rc) return rc;
  FGC_REQUIRE(x && adj && u && v && c && q, "assignments: NULL tensor pointer");
  const int64_t rows = static_cast<int64_t>(s->B) * s->N;
  Workspace ws(workspace, workspace_bytes);
  float* uvx = ws.take<float>(rows * 2 * s->M);
  FGC_REQUIRE(ws.ok(), "assignments: workspace too small");
  rc = launch_assign_logits(s, x, u, v, c, uvx, as_stream(stream));
  if (rc) return rc;
  return launch_assignments(adj, uvx, q, rows, s->N, s->K, s->M, as_stream(stream));
}

// ------------------------------------------------------------------ host-buffer entry points
int fgc_conv_fwd_host(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0,
                      const float* b, const float* u, const float* v, const float* c, float* y,
                      int bias_mask, int act, float alpha, int device) {
  int rc = check_shape(s, "conv_fwd_host");
  if (rc) return rc;
  FGC_REQUIRE(x && adj && W0 && b && u && v && c && y, "conv_fwd_host: NULL pointer");
  const size_t rows = static_cast<size_t>(s->B) * s->N;
  const size_t nx = rows * s->Cin, nadj = rows * s->K, ny = rows * s->Cout;
  const size_t nW = static_cast<size_t>(s->M) * s->Cout * s->Cw, nu = static_cast<size_t>(s->M) * s->Ca;
  const size_t wsb = conv_fwd_workspace(s);
  size_t total = 0;
  auto place = [&](size_t bytes) { size_t o = total; total = align_up(total + bytes, 256); return o; };
  const size_t ox = place(nx * 4), oadj = place(nadj * 4), oy = place(ny * 4), oW = place(nW * 4),
               ob = place(s->Cout * 4), ou = place(nu * 4), ov = place(nu * 4), oc = place(s->M * 4),
               ows = place(wsb);
  HostCallGuard guard;
  rc = host_reserve(device, total);
  if (rc) return rc;
  char* d = g_hc.dev;
  cudaStream_t st = g_hc.stream;
  FGC_CUDA(cudaMemcpyAsync(d + ox, x, nx * 4, cudaMemcpyHostToDevice, st));
  FGC_CUDA(cudaMemcpyAsync(d + oadj, adj, nadj * 4, cudaMemcpyHostToDevice, st));
  FGC_CUDA(cudaMemcpyAsync(d + oW, W0, nW * 4, cudaMemcpyHostToDevice, st));
  FGC_CUDA(cudaMemcpyAsync(d + ob, b, s->Cout * 4, cudaMemcpyHostToDevice, st));
  FGC_CUDA(cudaMemcpyAsync(d + ou, u, nu * 4, cudaMemcpyHostToDevice, st));
  FGC_CUDA(cudaMemcpyAsync(d + ov, v, nu * 4, cudaMemcpyHostToDevice, st));
  FGC_CUDA(cudaMemcpyAsync(d + oc, c, s->M * 4, cudaMemcpyHostToDevice, st));
  rc = conv_fwd(s, (float*)(d + ox), (int32_t*)(d + oadj), (float*)(d + oW), (float*)(d + ob),
                (float*)(d + ou), (float*)(d + ov), (float*)(d + oc), (float*)(d + oy), bias_mask, act,
                alpha, d + ows, wsb, st);
  if (rc) return rc;
  FGC_CUDA(cudaMemcpyAsync(y, d + oy, ny * 4, cudaMemcpyDeviceToHost, st));
  FGC_CUDA(cudaStreamSynchronize(st));
  return FGC_OK;
}

int fgc_conv_fwd_bwd_host(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* gy,
                          const float* W0, const float* b, const float* u, const float* v,
                          const float* c, float* y, float* gx, float* gW0, float* gb, float* gu,
                          float* gv, float* gc, int bias_mask, int device) {
  int rc = check_shape(s, "conv_fwd_bwd_host");
  if (rc) return rc;
  FGC_REQUIRE(x && adj && gy && W0 && b && u && v && c && y && gx && gW0 && gb && gu && gv && gc,
              "conv_fwd_bwd_host: NULL pointer");
  const size_t rows = static_cast<size_t>(s->B) * s->N;
  const size_t nx = rows * s->Cin, nadj = rows * s->K, ny = rows * s->Cout;
  const size_t nW = static_cast<size_t>(s->M) * s->Cout * s->Cw, nu = static_cast<size_t>(s->M) * s->Ca;
  const size_t wsf = conv_fwd_workspace(s), wsb = conv_bwd_workspace(s);
  const size_t wsr = reverse_adj_workspace(rows);
  // caches of the planned tensor-core path, rebuilt per call from adj (they only depend on adj) while x
  // is still on the wire: forward tile plan, reversed adjacency in forward layout (<= 32 slots) + its plan
  const bool mma = use_mma(s);
  const int kKrMax = 32;
  const size_t pbf = mma ? conv_plan_bytes(rows, s->K, s->M) : 0;
  const size_t pbr = mma ? conv_plan_bytes(rows, kKrMax, s->M) : 0;
  size_t total = 0;
  auto place = [&](size_t bytes) { size_t o = total; total = align_up(total + bytes, 256); return o; };
  const size_t ox = place(nx * 4), oadj = place(nadj * 4), ogy = place(ny * 4), oy = place(ny * 4),
               ogx = place(nx * 4), oW = place(nW * 4), ob = place(s->Cout * 4), ou = place(nu * 4),
               ov = place(nu * 4), oc = place(s->M * 4), ogW = place(nW * 4), ogb = place(s->Cout * 4),
               ogu = place(nu * 4), ogv = place(nu * 4), ogc = place(s->M * 4),
               orp = place((rows + 1) * 4), ore = place(nadj * 4), orw = place(wsr), owf = place(wsf),
               owb = place(wsb), opf = place(pbf), opr = place(pbr), ora = place(mma ? rows * kKrMax * 4 : 0),
               odg = place(256);
  HostCallGuard guard;
  rc = host_reserve(device, total);
  if (rc) return rc;
  // Three streams: copies in, kernels, copies out.  The adjacency goes first (the reverse adjacency
  // is built while x is still arriving), gy arrives under the forward, y leaves under the backward
  // and gx leaves as soon as it is final (before the weight-gradient passes).
  char* d = g_hc.dev;
  cudaStream_t st = g_hc.stream, si = g_hc.s_in, so = g_hc.s_out;
  cudaEvent_t e_adj = g_hc.ev[0], e_x = g_hc.ev[1], e_gy = g_hc.ev[2], e_y = g_hc.ev[3], e_gx = g_hc.ev[4],
              e_done = g_hc.ev[5];
  FGC_CUDA(cudaMemcpyAsync(d + oadj, adj, nadj * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaEventRecord(e_adj, si));
  FGC_CUDA(cudaMemcpyAsync(d + oW, W0, nW * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaMemcpyAsync(d + ob, b, s->Cout * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaMemcpyAsync(d + ou, u, nu * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaMemcpyAsync(d + ov, v, nu * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaMemcpyAsync(d + oc, c, s->M * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaMemcpyAsync(d + ox, x, nx * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaEventRecord(e_x, si));
  FGC_CUDA(cudaMemcpyAsync(d + ogy, gy, ny * 4, cudaMemcpyHostToDevice, si));
  FGC_CUDA(cudaEventRecord(e_gy, si));
  FGC_CUDA(cudaStreamWaitEvent(st, e_adj, 0));
  // the reverse adjacency has its own scratch (orw) so that it can run before the forward
  rc = build_reverse_adj((int32_t*)(d + oadj), s->B, s->N, s->K, (int32_t*)(d + orp),
                         (int32_t*)(d + ore), nullptr, d + orw, wsr, st);
  if (rc) return rc;
  const void* fplan = nullptr;
  const void* rplan = nullptr;
  const int32_t* radj = nullptr;
  int Kr = 0;
  if (mma) {
    // The plans pay per distinct row of a tile: with little neighbour sharing (random adjacency) the
    // per-facet gather kernels are faster, so the plan is dropped above kMaxMeanRows rows per tile.
    const double kMaxMeanRows = 112.0;
    long long hdr[2] = {0, 0};
    int deg = 0;
    rc = build_conv_plan((int32_t*)(d + oadj), s->B, s->N, s->K, s->M, d + opf, pbf, st);
    if (rc) return rc;
    rc = launch_max_degree((int32_t*)(d + orp), rows, (int32_t*)(d + odg), st);
    if (rc) return rc;
    FGC_CUDA(cudaMemcpyAsync(hdr, d + opf, sizeof(hdr), cudaMemcpyDeviceToHost, st));
    FGC_CUDA(cudaMemcpyAsync(&deg, d + odg, sizeof(deg), cudaMemcpyDeviceToHost, st));
    FGC_CUDA(cudaStreamSynchronize(st));   // x is still arriving on the copy stream
    if (hdr[1] > 0 && static_cast<double>(hdr[0]) / hdr[1] <= kMaxMeanRows) fplan = d + opf;
    Kr = deg < 8 ? 8 : (deg + 7) / 8 * 8;
    if (Kr <= kKrMax) {
      rc = launch_build_radj((int32_t*)(d + orp), (int32_t*)(d + ore), s->B, s->N, s->K, Kr, (int32_t*)(d + ora), st);
      if (rc) return rc;
      rc = build_conv_plan((int32_t*)(d + ora), s->B, s->N, Kr, s->M, d + opr, pbr, st);
      if (rc) return rc;
      FGC_CUDA(cudaMemcpyAsync(hdr, d + opr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
      FGC_CUDA(cudaStreamSynchronize(st));
      if (hdr[1] > 0 && static_cast<double>(hdr[0]) / hdr[1] <= kMaxMeanRows) {
        rplan = d + opr;
        radj = (int32_t*)(d + ora);
      }
    }
  }
  FGC_CUDA(cudaStreamWaitEvent(st, e_x, 0));
  rc = conv_fwd(s, (float*)(d + ox), (int32_t*)(d + oadj), (float*)(d + oW), (float*)(d + ob),
                (float*)(d + ou), (float*)(d + ov), (float*)(d + oc), (float*)(d + oy), bias_mask,
                FGC_ACT_NONE, 0.f, d + owf, wsf, st, fplan);
  if (rc) return rc;
  FGC_CUDA(cudaEventRecord(e_y, st));
  FGC_CUDA(cudaStreamWaitEvent(so, e_y, 0));
  FGC_CUDA(cudaMemcpyAsync(y, d + oy, ny * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaStreamWaitEvent(st, e_gy, 0));
  g_gx_ready_event = e_gx;
  rc = conv_bwd(s, (float*)(d + ogy), (float*)(d + ox), (int32_t*)(d + oadj), (int32_t*)(d + orp),
                (int32_t*)(d + ore), (float*)(d + oW), (float*)(d + ou), (float*)(d + ov),
                (float*)(d + oc), (float*)(d + ogx), (float*)(d + ogW), (float*)(d + ogb),
                (float*)(d + ogu), (float*)(d + ogv), (float*)(d + ogc), bias_mask, d + owb, wsb, st, radj, Kr,
                rplan, fplan, fplan != nullptr ? d + owf : nullptr, wsf);
  g_gx_ready_event = nullptr;
  if (rc) return rc;
  FGC_CUDA(cudaStreamWaitEvent(so, e_gx, 0));
  FGC_CUDA(cudaMemcpyAsync(gx, d + ogx, nx * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaEventRecord(e_done, st));
  FGC_CUDA(cudaStreamWaitEvent(so, e_done, 0));
  FGC_CUDA(cudaMemcpyAsync(gW0, d + ogW, nW * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaMemcpyAsync(gb, d + ogb, s->Cout * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaMemcpyAsync(gu, d + ogu, nu * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaMemcpyAsync(gv, d + ogv, nu * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaMemcpyAsync(gc, d + ogc, s->M * 4, cudaMemcpyDeviceToHost, so));
  FGC_CUDA(cudaStreamSynchronize(so));
  FGC_CUDA(cudaStreamSynchronize(st));
  FGC_CUDA(cudaStreamSynchronize(si));
  return FGC_OK;
}

void fgc_host_release(void) {
  if (g_hc.dev) {
    cudaSetDevice(g_hc.device);
    cudaFree(g_hc.dev);
  }
  host_destroy_streams();
  g_hc = HostCache();
}

}  // extern "C"
