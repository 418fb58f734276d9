// Host-side launch interfaces shared between the convolution translation units and c_api.cu.
#pragma once

#include "common.cuh"

namespace fgc {

struct ConvFwdParams {
  const float* x;
  const int32_t* adj;
  const float* uvx;
  const float* Wt;  // [M*Cw][Cout]
  const float* b;
  float* y;
  int64_t rows;
  int N, K, Cin, Cw, Cout, M;
  int bias_mask, act;
  float alpha;
};

int launch_assign_logits(const fgc_conv_shape* s, const float* x, const float* u, const float* v,
                         const float* c, float* uvx, cudaStream_t st, unsigned* maxbits = nullptr);
// warp-cooperative fast paths of the logit kernels (logits.cu)
bool logits_fast_supported(int Cin, int Ca0, int Ca, int M);
int launch_assign_logits_fast(const float* x, const float* u, const float* v, const float* c, float* uvx,
                              int64_t rows, int Cin, int Ca0, int Ca, int M, cudaStream_t st, unsigned* maxbits = nullptr);
int launch_logits_bwd_x_fast(const float* d_uvx, const float* u, const float* v, float* gx, int64_t rows, int Cin,
                             int Ca0, int Ca, int M, cudaStream_t st);
int launch_logits_bwd_p_fast(const float* x, const float* d_uvx, float* part, int64_t rows, int64_t rows_per_chunk,
                             int chunks, int Cin, int Ca0, int Ca, int M, cudaStream_t st);
int launch_transpose_w(const float* W0, float* Wt, int M, int Cout, int Cw, cudaStream_t st);
int launch_conv_fwd(const ConvFwdParams& p, cudaStream_t st);
int launch_assignments(const int32_t* adj, const float* uvx, float* q, int64_t rows, int N, int K,
                       int M, cudaStream_t st);
int launch_gather_rows(const float* x, const int32_t* adj, float* out, int64_t rows, int N, int K,
                       int C, cudaStream_t st);
// tensor-core (tcgen05) forward for the dense shapes
bool conv_fwd_tc_supported(int Cw, int Cout, int M, int K);
size_t conv_fwd_tc_workspace(int Cout, int M, int Cw = 64);
bool conv_fwd_small_supported(const fgc_conv_shape* s);
int launch_conv_fwd_small(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0, const float* b,
                          const float* u, const float* v, const float* c, float* y, int bias_mask, int act,
                          float alpha, cudaStream_t st, float* ypool = nullptr, unsigned* ymax = nullptr);
// ypool (optional): [rows / 4][32] max over groups of 4 rows; ymax (optional): [B] atomicMax targets for max|y| bits
int launch_conv_fwd_tc(const ConvFwdParams& p, const float* W0, void* wimg_ws, cudaStream_t st, int upshift = 0);
bool bwd_tgt_tc_supported(int Cw, int Cout, int M);
int launch_bwd_tgt_tc(const float* gy, const float* uvx, const float* W0, const float* da_edge,
                      const float* inv, const int32_t* rev_ptr, const int32_t* rev_edge, float* gx,
                      float* d_uvx, int64_t rows, int N, int K, int Cin, int Cw, int Cout, int M,
                      void* wimg_ws, cudaStream_t st);
int launch_prep_w_image(const float* W0, void* wimg_ws, int M, int Cout, cudaStream_t st);
// dense-assignment tcgen05 path (conv_mma.cu): tile plan + forward
bool conv_mma_supported(int Cin, int Cw, int Cout, int M, int K);
size_t conv_plan_bytes(int64_t rows, int K, int M);
int build_conv_plan(const int32_t* adj, int B, int N, int K, int M, void* plan, size_t plan_bytes, cudaStream_t st);
size_t conv_mma_workspace(int64_t rows);
int debug_mma_trace(int64_t* out, int n);
int debug_hm_trace(int64_t* out, int n);
// 128-byte CUtensorMap of an fp16 hi|lo image with `nunits` 64-channel units per row (false: no encoder / disabled)
bool make_hm_img_tmap(void* tm, const void* img, int64_t rows, int nunits);
int prep_image_blocks();
int launch_prep_image(const float* x, int ld, int64_t rows, void* img_ws, cudaStream_t st,
                      const float* pinv = nullptr, int bias_mask = 0, float* partB = nullptr, bool have_absmax = false);
// true when launch_assign_logits(..., maxbits) reduces max|x| over exactly the channels the image uses
bool assign_logits_absmax_supported(const fgc_conv_shape* s);
int prep_image_reset(void* img_ws, int64_t rows, cudaStream_t st);   // zeroes the scale words; returns them via conv_mma_image_maxbits
int bwd_w_mma_grid(int64_t rows, int M);
int launch_bwd_w_mma(const float* uvx, const int32_t* adj, const void* plan, void* ximg_ws, void* gyimg_ws,
                     float* partW, int64_t rows, int N, int K, int M, cudaStream_t st);
bool bwd_src_mma_supported(int Cin, int Cw, int Cout, int M, int K);
const float* conv_plan_inv(const void* plan, int64_t rows, int K, int M);
int launch_bwd_src_mma(const float* gy, const float* uvx, const int32_t* adj, const void* plan, void* ximg_ws,
                       void* gyimg_ws, const void* wimg, float* da_edge, float* d_uvx, int64_t rows, int N, int K,
                       int M, cudaStream_t st);
int launch_max_degree(const int32_t* rev_ptr, int64_t rows, int32_t* out, cudaStream_t st);
int launch_build_radj(const int32_t* rev_ptr, const int32_t* rev_edge, int B, int N, int K, int Kr, int32_t* radj,
                      cudaStream_t st);
bool bwd_tgt_mma_supported(int Cin, int Cw, int Cout, int M, int Kr);
int launch_bwd_tgt_mma(const float* gy, const float* uvx, const float* da_edge, const float* inv,
                       const int32_t* rev_ptr, const int32_t* rev_edge, const int32_t* radj, int Kr, const void* rplan,
                       float* gx, float* d_uvx, int64_t rows, int N, int Cin, int Cout, int M, const void* wimg,
                       void* img_ws, cudaStream_t st);
int launch_conv_mma(const ConvFwdParams& p, const float* W0, const void* plan, void* img_ws, void* wimg_ws,
                    cudaStream_t st, bool have_absmax = false);
int launch_prep_w_image_t(const float* W0, void* wimg_ws, int M, int Cw, cudaStream_t st);
bool bwd_src_tc_supported(int Cw, int Cout, int M, int Cin);
int launch_bwd_src_tc(const float* gy, const float* x, const int32_t* adj, const float* uvx, const void* wimg,
                      const float* wunscale, float* da_edge, float* d_uvx, float* inv_out, int64_t rows, int N,
                      int K, int Cin, int M, cudaStream_t st);
bool bwd_w_tc_supported(int Cw, int Cout, int M, int Cin);
int bwd_w_tc_grid(int64_t rows);
int launch_bwd_w_tc(const float* gy, const float* x, const int32_t* adj, const float* uvx, float* partW,
                    float* partB, unsigned* maxbits, const unsigned* xmax, const unsigned* gmax, int64_t rows, int N,
                    int K, int Cin, int M, int bias_mask, cudaStream_t st);
const unsigned* conv_mma_image_maxbits(const void* img_ws, int64_t rows);
// products the planned forward leaves in its workspace (c_api.cu: conv_fwd), reusable by the backward
struct FwdSaved {
  const float* uvx;   // [rows][2M] assignment logits
  const char* ximg;   // conv_mma_workspace(rows) bytes: fp16 hi|lo image of x + its scale words
};
int conv_fwd_saved_views(const fgc_conv_shape* s, const void* fwd_ws, size_t fwd_ws_bytes, FwdSaved* out);
extern thread_local cudaEvent_t g_gx_ready_event;
size_t conv_bwd_workspace(const fgc_conv_shape* s);
int conv_bwd(const fgc_conv_shape* s, const float* gy, const float* x, const int32_t* adj,
             const int32_t* rev_ptr, const int32_t* rev_edge, const float* W0, const float* u,
             const float* v, const float* c, float* gx, float* gW0, float* gb, float* gu, float* gv,
             float* gc, int bias_mask, void* workspace, size_t workspace_bytes, cudaStream_t st,
             const int32_t* radj = nullptr, int Kr = 0, const void* rplan = nullptr, const void* fplan = nullptr,
             const void* fwd_ws = nullptr, size_t fwd_ws_bytes = 0);
size_t reverse_adj_workspace(int64_t rows);
int build_reverse_adj(const int32_t* adj, int B, int N, int K, int32_t* rev_ptr, int32_t* rev_edge,
                      int64_t* nnz_out, void* workspace, size_t workspace_bytes, cudaStream_t st);
int launch_reduce_partials(const float* part, float* out, int64_t n, int P, int64_t stride,
                           cudaStream_t st);
// HMMA-aggregation + tcgen05-contraction forward (conv_hm.cu): any adjacency, no tile plan
bool conv_hm_supported(int Cin, int Cw, int Cout, int M, int K);
size_t conv_hm_workspace(int64_t rows_img, int Cw, int Cout, int M, int B);
int launch_conv_hm(const ConvFwdParams& p, const float* W0, const float* u, const float* v, const float* c, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, int upshift = 0, float* ypool = nullptr, unsigned* ymax = nullptr);
size_t conv_hm_weights_bytes(int Cw, int Cout, int M);
int launch_conv_hm_weights(const float* W0, int M, int Cout, int Cw, void* wbuf, cudaStream_t st);
int launch_conv_hm_core(const void* img, const float* xunscale, const float* lg, const unsigned* flag, const int32_t* adj,
                        const void* wbuf,
                        const float* b, float* y, float* ypool, unsigned* ymax, int64_t rows, int N, int K, int M, int Cw,
                        int Cout, int upshift, int bias_mask, int act, float alpha, cudaStream_t st,
                        const char* tag = nullptr);   // tag: name the launches carry in the library profiler
int launch_absmax_bits(const float* x, int64_t n4_per_elem, int B, unsigned* out, cudaStream_t st);
// fused pre-pass (logits.cu): logits + fp16 image of [xa | xb] in one pass
int launch_prep_rows(const float* xa, int lda, int Ca, const float* xb, int ldb, int Cb, const float* u, const float* v,
                     const float* c, int M, int64_t rows, int Nimg, const unsigned* maxa, const unsigned* maxb, void* img,
                     float* lg, float* xunscale, unsigned* flag, cudaStream_t st, const char* tag = nullptr);
bool prep_rows_supported(int Ca, int Cb, int M);
// fused regression head on tcgen05 (lin_tc.cu)
bool mlp_head_tc_supported(int64_t rows, int Cin, int H, int Cout);
size_t mlp_head_tc_workspace();
int launch_mlp_head_tc(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* y,
                       int64_t rows, float alpha, void* workspace, size_t workspace_bytes, cudaStream_t st,
                       const void* prepared = nullptr);   // prepared: from launch_mlp_head_tc_prepare, or null
int launch_mlp_head_tc_prepare(const float* W1, void* prepared, size_t prepared_bytes, cudaStream_t st);

}  // namespace fgc
