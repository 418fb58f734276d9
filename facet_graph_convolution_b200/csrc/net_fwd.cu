// The reference network's inference forward as ONE C-ABI call (reference Code/model.py:837-946,
// get_model_reg_multi_scale with multiScale = False): 8 facet-graph convolutions, 2 max-pools, 2 up-samplings,
// 2 channel concatenations and the 2-layer regression head, with what the per-layer entry points cannot do fused:
//   * custom_binary_tree_pooling (model.py:863,875) is a second output of the producing convolution's epilogue;
//   * custom_upsampling (model.py:902,923) is an index shift in the consuming convolution's gather;
//   * tf.concat (model.py:909,929) is never materialised: the pre-pass of the consuming layer reads two tensors;
//   * every layer input is read ONCE by a pre-pass that writes its assignment logits and its fp16 hi|lo image, scaled
//     by max|x| of the batch element, which the producing kernel left behind (no separate reduction);
//   * the tcgen05 weight images depend only on the parameters: fgc_net_prepare builds them once per checkpoint.
// Batch elements are independent patches (per-patch adjacency ids, per-patch scales): a batch gives every patch
// exactly the rows it would get alone.
#include "conv_launch.cuh"

namespace fgc {
namespace {

constexpr int kNetM = 9;
constexpr int kNetConvs = 8;
// creation order (model.py:853-941): conv1 conv2 conv3 dconv3 upconv2 dconv2 upconv1 dconv1, then the head
constexpr int kCin[kNetConvs] = {6, 32, 64, 128, 128, 128, 64, 64};
constexpr int kCout[kNetConvs] = {32, 64, 128, 128, 64, 64, 32, 32};
constexpr int kNetParams = kNetConvs * 5 + 4;
// names the launches of a layer carry in the library profiler (fgc_profile_begin/_end): bench.py's per-layer table
const char* const kLayerTag[kNetConvs] = {"conv1", "conv2", "conv3", "dconv3", "upconv2", "dconv2", "upconv1", "dconv1"};
const char* const kPrepTag[kNetConvs] = {"prep_conv1", "prep_conv2", "prep_conv3", "prep_dconv3", "prep_upconv2",
                                         "prep_dconv2", "prep_upconv1", "prep_dconv1"};

struct PreparedLayout {
  size_t off[kNetConvs];
  size_t head;    // fp16 image of the head's W1 (mlp_head_tc_workspace() bytes)
  size_t total;
  PreparedLayout() {
    size_t o = 256;
    off[0] = 0;
    for (int l = 1; l < kNetConvs; ++l) {
      off[l] = o;
      o += align_up(conv_hm_weights_bytes(kCin[l], kCout[l], kNetM), 256);
    }
    head = o;
    o += align_up(mlp_head_tc_workspace(), 256);
    total = o;
  }
};

struct ConvP {
  const float *W0, *b, *u, *c, *v;
};
ConvP conv_params(const float* const* params, int l) {
  return ConvP{params[5 * l], params[5 * l + 1], params[5 * l + 2], params[5 * l + 3], params[5 * l + 4]};
}

struct NetWs {
  float *h1, *p1, *h2, *p2, *h3, *d3, *u2, *d2, *u1, *d1, *lg, *xunscale;
  char *img, *head;
  unsigned* mx;   // [8][B] max|.| bits: 0 h1, 1 h2, 2 h3, 3 d3, 4 u2, 5 d2, 6 u1
  bool ok;
};
size_t head_ws_bytes(int64_t R0) { return fgc_mlp_head_workspace(R0, 32, 1024, 3) + 256; }
NetWs carve(void* p, size_t n, int B, int64_t R0) {
  Workspace ws(p, n);
  const int64_t R1 = R0 / 4, R2 = R0 / 16;
  NetWs w;
  w.h1 = ws.take<float>(R0 * 32), w.p1 = ws.take<float>(R1 * 32), w.h2 = ws.take<float>(R1 * 64);
  w.p2 = ws.take<float>(R2 * 64), w.h3 = ws.take<float>(R2 * 128), w.d3 = ws.take<float>(R2 * 128);
  w.u2 = ws.take<float>(R1 * 64), w.d2 = ws.take<float>(R1 * 64), w.u1 = ws.take<float>(R0 * 32);
  w.d1 = ws.take<float>(R0 * 32);
  w.lg = ws.take<float>((R0 + 1) * 32);
  w.img = ws.take<char>(static_cast<size_t>(R0 + 1) * 256);       // largest image: dconv1, R0 rows x 1 unit (= R1 x 2 x 2)
  w.xunscale = ws.take<float>(B);
  w.mx = ws.take<unsigned>(8 * static_cast<size_t>(B) + 16);   // + the pre-pass's spread flag
  w.head = ws.take<char>(head_ws_bytes(R0));
  w.ok = ws.ok();
  return w;
}

}  // namespace
}  // namespace fgc

using namespace fgc;

extern "C" {

int fgc_net_param_count(void) { return kNetParams; }

size_t fgc_net_prepared_bytes(void) { return PreparedLayout().total; }

int fgc_net_prepare(const float* const* params, int nparams, void* prepared, size_t prepared_bytes, void* stream) {
  FGC_REQUIRE(params != nullptr && nparams == kNetParams, "net_prepare: %d parameter tensors expected in creation order, got %d",
              kNetParams, nparams);
  const PreparedLayout L;
  FGC_REQUIRE(prepared != nullptr && prepared_bytes >= L.total, "net_prepare: buffer too small (%zu given, %zu needed)",
              prepared_bytes, L.total);
  cudaStream_t st = as_stream(stream);
  for (int l = 1; l < kNetConvs; ++l) {
    int rc = launch_conv_hm_weights(params[5 * l], kNetM, kCout[l], kCin[l], static_cast<char*>(prepared) + L.off[l], st);
    if (rc) return rc;
  }
  return launch_mlp_head_tc_prepare(params[40], static_cast<char*>(prepared) + L.head, mlp_head_tc_workspace(), st);
}

size_t fgc_net_fwd_workspace(int B, int N0, int K) {
  if (B <= 0 || N0 <= 0 || K <= 0) return 0;
  const int64_t R0 = static_cast<int64_t>(B) * N0;
  const int64_t R1 = R0 / 4, R2 = R0 / 16;
  size_t n = 0;
  for (int64_t c : {R0 * 32, R1 * 32, R1 * 64, R2 * 64, R2 * 128, R2 * 128, R1 * 64, R1 * 64, R0 * 32, R0 * 32,
                    (R0 + 1) * 32})
    n += ws_bytes(static_cast<size_t>(c), 4);
  n += ws_bytes(static_cast<size_t>(R0 + 1) * 256, 1) + ws_bytes(B, 4) + ws_bytes(8 * static_cast<size_t>(B) + 16, 4) +
       ws_bytes(head_ws_bytes(R0), 1);
  return n + 1024;
}

int fgc_net_fwd(int B, int N0, int K, const float* x, const int32_t* adj0, const int32_t* adj1, const int32_t* adj2,
                const float* const* params, int nparams, const void* prepared, float* y, void* workspace,
                size_t workspace_bytes, void* stream) {
  FGC_REQUIRE(x && adj0 && adj1 && adj2 && params && prepared && y, "net_fwd: NULL pointer");
  FGC_REQUIRE(nparams == kNetParams, "net_fwd: %d parameter tensors expected in creation order, got %d", kNetParams, nparams);
  FGC_REQUIRE(B > 0 && N0 >= 16 && N0 % 16 == 0, "net_fwd: N0 = %d must be a positive multiple of 16 (two x4 poolings)", N0);
  FGC_REQUIRE(K >= 1 && K <= 32, "net_fwd: K = %d outside 1..32", K);
  const int64_t R0 = static_cast<int64_t>(B) * N0, R1 = R0 / 4, R2 = R0 / 16;
  FGC_REQUIRE(R0 * K < (1ll << 31), "net_fwd: B*N0*K must be < 2^31");
  const int N1 = N0 / 4, N2 = N0 / 16;
  const NetWs w = carve(workspace, workspace_bytes, B, R0);
  FGC_REQUIRE(w.ok, "net_fwd: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              fgc_net_fwd_workspace(B, N0, K));
  cudaStream_t st = as_stream(stream);
  const PreparedLayout L;
  const char* prep = static_cast<const char*>(prepared);
  const float alpha = 0.1f;   // model.py:841
  FGC_CUDA(cudaMemsetAsync(w.mx, 0, (8 * static_cast<size_t>(B) + 16) * sizeof(unsigned), st));
  unsigned* flag = w.mx + 8 * static_cast<size_t>(B);
  unsigned* mx[8];
  for (int i = 0; i < 8; ++i) mx[i] = w.mx + static_cast<size_t>(i) * B;
  int rc;
  // ---- Level 0: conv1 (6 -> 32, gather-bound: thread per facet, logits inline) + lrelu; pool and max|h1| fused
  {
    const ConvP p = conv_params(params, 0);
    fgc_conv_shape s{B, N0, K, 6, 6, 0, 6, 32, kNetM};
    rc = launch_conv_fwd_small(&s, x, adj0, p.W0, p.b, p.u, p.v, p.c, w.h1, 1, FGC_ACT_LRELU, alpha, st, w.p1, mx[0]);
    if (rc) return rc;   // pooled copy and max|h1| (bounds max|p1| too) come from the kernel's epilogue
  }
  // one dense layer: pre-pass over [xa | xb] (rows_in rows of Nin per element), then the convolution on `adj`
  auto layer = [&](int l, const float* xa, int Ca, const unsigned* ma, const float* xb, int Cb, const unsigned* mb,
                   int64_t rows_in, int Nin, const int32_t* adj, int64_t rows, int N, int upshift, int act, float* out,
                   float* pooled, unsigned* omax) -> int {
    const ConvP p = conv_params(params, l);
    int r = launch_prep_rows(xa, Ca, Ca, xb, Cb, Cb, p.u, p.v, p.c, kNetM, rows_in, Nin, ma, mb, w.img, w.lg, w.xunscale, flag,
                             st, kPrepTag[l]);
    if (r) return r;
    return launch_conv_hm_core(w.img, w.xunscale, w.lg, flag, adj, prep + L.off[l], p.b, out, pooled, omax, rows, N, K, kNetM,
                               Ca + Cb, kCout[l], upshift, 1, act, alpha, st, kLayerTag[l]);
  };
  // ---- Level 1: conv2 on pool(h1); its epilogue also writes pool(h2)
  rc = layer(1, w.p1, 32, mx[0], nullptr, 0, nullptr, R1, N1, adj1, R1, N1, 0, FGC_ACT_LRELU, w.h2, w.p2, mx[1]);
  if (rc) return rc;
  // ---- Level 2
  rc = layer(2, w.p2, 64, mx[1], nullptr, 0, nullptr, R2, N2, adj2, R2, N2, 0, FGC_ACT_LRELU, w.h3, nullptr, mx[2]);
  if (rc) return rc;
  rc = layer(3, w.h3, 128, mx[2], nullptr, 0, nullptr, R2, N2, adj2, R2, N2, 0, FGC_ACT_LRELU, w.d3, nullptr, mx[3]);
  if (rc) return rc;
  // ---- Level 1: upconv2 on repeat(d3, 4) (index shift, no activation: model.py:902-905), dconv2 on [upconv2 | h2]
  rc = layer(4, w.d3, 128, mx[3], nullptr, 0, nullptr, R2, N2, adj1, R1, N1, 2, FGC_ACT_NONE, w.u2, nullptr, mx[4]);
  if (rc) return rc;
  rc = layer(5, w.u2, 64, mx[4], w.h2, 64, mx[1], R1, N1, adj1, R1, N1, 0, FGC_ACT_LRELU, w.d2, nullptr, mx[5]);
  if (rc) return rc;
  // ---- Level 0: upconv1 on repeat(d2, 4), dconv1 on [upconv1 | h1]
  rc = layer(6, w.d2, 64, mx[5], nullptr, 0, nullptr, R1, N1, adj0, R0, N0, 2, FGC_ACT_NONE, w.u1, nullptr, mx[6]);
  if (rc) return rc;
  rc = layer(7, w.u1, 32, mx[6], w.h1, 32, mx[0], R0, N0, adj0, R0, N0, 0, FGC_ACT_LRELU, w.d1, nullptr, nullptr);
  if (rc) return rc;
  // ---- regression head 32 -> 1024 -> 3 (model.py:936-941), hidden activation never materialised
  if (mlp_head_tc_supported(R0, 32, 1024, 3))   // W1's image comes prepared
    return launch_mlp_head_tc(w.d1, params[40], params[41], params[42], params[43], y, R0, alpha, w.head, head_ws_bytes(R0), st,
                              prep + L.head);
  return fgc_mlp_head_fwd(w.d1, params[40], params[41], params[42], params[43], y, R0, 32, 1024, 3, alpha, w.head,
                          head_ws_bytes(R0), stream);
}

}  // extern "C"
