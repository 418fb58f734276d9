/*
 * facetconv_b200.h -- C ABI of the B200-native facet-graph convolution hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference
 * tree Elensil/Facet_Graph_Convolution).  The reference is Python/TensorFlow and has no
 * FFI of its own; INTEGRATION.md shows the ctypes stub a maintainer would add to
 * Code/model.py to route custom_conv2d & friends through this library.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers (cudaMalloc / torch CUDA storage), fp32
 *     row-major contiguous, indices int32, unless the function name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls only enqueue work, they never synchronise (the _host variants do);
 *   - inputs are borrowed and never written; outputs must not alias inputs;
 *   - return value 0 = success, non-zero = error; fgc_last_error() gives the message
 *     (thread-local).  No CPU fallback exists: without a CUDA device every compute
 *     entry point fails with FGC_ERR_CUDA.
 *   - adjacency layout: adj[B,N,K] int32, 1-indexed neighbour ids, 0 = padding,
 *     column 0 = the facet itself (reference Code/utils.py:243-295, 1799-1827).
 */
#ifndef FACETCONV_B200_H_
#define FACETCONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FGC_API __attribute__((visibility("default")))
#else
#define FGC_API
#endif

#define FGC_OK 0
#define FGC_ERR_ARG 1
#define FGC_ERR_CUDA 2
#define FGC_ERR_UNSUPPORTED 3

#define FGC_ACT_NONE 0
#define FGC_ACT_LRELU 1 /* relu(x) - alpha*relu(-x), reference Code/model.py:828-830 */

#define FGC_MAX_K 32  /* neighbour slots per facet (reference default K_faces = 23) */
#define FGC_MAX_M 16  /* weight matrices per layer (reference network: 9)          */
#define FGC_MAX_C 256 /* channels per row                                           */

FGC_API int fgc_version(void);
FGC_API const char* fgc_last_error(void);
/* number of CUDA devices visible (0 when there is no driver/GPU); never fails */
FGC_API int fgc_device_count(void);
/* kernels launched by this library in the calling process since load (all threads) */
FGC_API uint64_t fgc_launch_count(void);

/* Opt-in per-kernel timing for the bench harness (not thread-safe; one stream at a time):
 * between _begin and _end every kernel this library launches on `stream` is bracketed by CUDA
 * events; _end synchronises on the last one and writes lines "kernel_name total_ms launches\n". */
FGC_API int fgc_profile_begin(void* stream);
FGC_API int fgc_profile_end(char* buf, size_t buf_bytes);

/* ---------------------------------------------------------------- facet-graph convolution
 *
 * Replaces reference Code/model.py:427-504 (custom_conv2d) together with :74-95
 * (get_weight_assigments) and :380-405 (get_slices/get_patches); with the channel
 * windows below it also covers the variants at :97-124, :610-696 and :699-760.
 *
 *   uvx[r, 0:M]  = u . x_r[ca0:ca0+ca] + c        (own-row logit part)
 *   uvx[r, M:2M] = v . x_r[ca0:ca0+ca]            (neighbour logit part)
 *   q[n,k,:]     = softmax_m( uvx[n,0:M] + (adj[n,k] ? uvx[adj[n,k]-1, M:2M] : 0) )
 *   s[n,m,:]     = sum_k q[n,k,m] * x_{adj[n,k]-1}[0:cw]          (padding adds 0)
 *   y[n,:]       = act( inv_cnt[n] * sum_m W0[m] s[n,m,:] + (cnt[n]>0 || !bias_mask) * b )
 *
 * Translation-invariant assignments (model.py:97-124) are the same formula with
 * v = -u.  "position for assignment" (model.py:610-696): ca0 = 0, ca = Cin, cw = Cin-3.
 * "only position" (model.py:699-760): ca0 = Cin-3, ca = 3, cw = Cin-3, bias_mask = 0.
 */
typedef struct fgc_conv_shape {
  int32_t B;    /* batch (patches); rows of different batch elements never mix */
  int32_t N;    /* facets per batch element */
  int32_t K;    /* neighbour slots, <= FGC_MAX_K */
  int32_t Cin;  /* row width of x */
  int32_t Cw;   /* channels [0,Cw) enter the aggregation / contraction; W0 is [M,Cout,Cw] */
  int32_t Ca0;  /* channels [Ca0,Ca0+Ca) enter the assignment logits; u,v are [M,Ca] */
  int32_t Ca;
  int32_t Cout;
  int32_t M;    /* <= FGC_MAX_M */
} fgc_conv_shape;

/* bytes of device scratch fgc_conv_fwd / fgc_conv_bwd need for this shape (upper bound) */
FGC_API size_t fgc_conv_fwd_workspace(const fgc_conv_shape* s);
FGC_API size_t fgc_conv_bwd_workspace(const fgc_conv_shape* s);

/* forward; y[B,N,Cout].  v may not be NULL (pass -u for translation invariance). */
FGC_API int fgc_conv_fwd(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0,
                 const float* b, const float* u, const float* v, const float* c, float* y,
                 int bias_mask, int act, float alpha, void* workspace, size_t workspace_bytes,
                 void* stream);

/* forward with a fused custom_upsampling (reference Code/model.py:817-825 followed by :427-504, as the network
 * does at :902-905 and :923-926): x_coarse has (B*N) >> upshift rows and row r of the layer reads row
 * r >> upshift (features and assignment logits), so the repeated tensor is never materialised.  Same values,
 * bit for bit, as fgc_conv_fwd on the repeated input.  fgc_conv_fwd_up_supported tells whether the shape
 * has this path (the M = 9 layers on the tcgen05 forward); workspace as fgc_conv_fwd_workspace(s). */
FGC_API int fgc_conv_fwd_up_supported(const fgc_conv_shape* s, int upshift);
FGC_API int fgc_conv_fwd_up(const fgc_conv_shape* s, const float* x_coarse, const int32_t* adj, const float* W0,
                    const float* b, const float* u, const float* v, const float* c, float* y,
                    int bias_mask, int act, float alpha, int upshift, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Tile plan (caller-owned cache, built once per adjacency like the reverse adjacency below; pure
 * index work on adj, bit-exact): for every tile of 128/M consecutive facets the list of distinct
 * neighbour rows it touches and, per (facet, slot), the row's local index and the multiplicity of
 * repeated ids.  With a plan the dense layers run the dense-assignment tcgen05 path, in which each
 * neighbour row is fetched once per tile.  fgc_conv_plan_bytes returns 0 for shapes without a
 * planned path (then pass plan = NULL / use fgc_conv_fwd).  fgc_conv_fwd_planned computes exactly
 * what fgc_conv_fwd computes (same reference lines, Code/model.py:427-504). */
FGC_API size_t fgc_conv_plan_bytes(int B, int N, int K, int M);
FGC_API int fgc_build_conv_plan(const int32_t* adj, int B, int N, int K, int M, void* plan,
                        size_t plan_bytes, void* stream);
FGC_API int fgc_conv_fwd_planned(const fgc_conv_shape* s, const float* x, const int32_t* adj,
                         const void* plan, const float* W0, const float* b, const float* u,
                         const float* v, const float* c, float* y, int bias_mask, int act,
                         float alpha, void* workspace, size_t workspace_bytes, void* stream);

/* diagnostics: with FGC_MMA_TRACE set in the environment the planned kernel records clock64 stamps of
 * its pipeline roles for CTA 0 (4 roles x 32 tiles x 8 events); this copies them out (synchronises). */
FGC_API int fgc_debug_trace(int64_t* out, int n);

/* Reverse adjacency (caller-owned cache, built once per adjacency; replaces the scatter of
 * TF's gather gradient, UnsortedSegmentSum): for every target row t = b*N + j the list of
 * edge ids e = (b*N + n)*K + k with adj[b,n,k] == j+1, ascending.  rev_ptr[B*N+1],
 * rev_edge[nnz] with nnz <= B*N*K (allocate B*N*K).  *nnz_out (host) receives nnz; this
 * call synchronises the stream. */
FGC_API size_t fgc_reverse_adj_workspace(int B, int N, int K);
FGC_API int fgc_build_reverse_adj(const int32_t* adj, int B, int N, int K, int32_t* rev_ptr,
                          int32_t* rev_edge, int64_t* nnz_out, void* workspace,
                          size_t workspace_bytes, void* stream);

/* backward of fgc_conv_fwd (act = NONE; the activation gradient is applied by the caller or by
 * fgc_lrelu_bwd).  Deterministic: no floating-point atomics, fixed summation order.
 * Outputs: gx[B,N,Cin], gW0[M,Cout,Cw], gb[Cout], gu[M,Ca], gv[M,Ca], gc[M] (all overwritten). */
FGC_API int fgc_conv_bwd(const fgc_conv_shape* s, const float* gy, const float* x, const int32_t* adj,
                 const int32_t* rev_ptr, const int32_t* rev_edge, const float* W0, const float* u,
                 const float* v, const float* c, float* gx, float* gW0, float* gb, float* gu,
                 float* gv, float* gc, int bias_mask, void* workspace, size_t workspace_bytes,
                 void* stream);

/* Target-centric pass on the planned path.  fgc_build_reverse_padded writes the reversed adjacency
 * in the forward layout: radj[B,N,Kr], 1-indexed source facets of every target's in-edges in
 * rev_edge order, 0 padded (Kr >= the largest in-degree, <= FGC_MAX_K).  A tile plan built on radj
 * (fgc_build_conv_plan(radj, B, N, Kr, M, ...)) lets fgc_conv_bwd_planned run the gx pass on the
 * dense-assignment tensor-core kernel, and the forward plan of adj lets it run the source-centric
 * pass (ds, dq, da) on the staged tcgen05 pipeline; with plan = radj = rplan = NULL it equals
 * fgc_conv_bwd.
 * fwd_workspace (optional, needs plan): the workspace buffer fgc_conv_fwd_planned ran in for the same
 * shape, x, u, v, c, untouched since -- what an autograd context saves for backward.  The backward then
 * reuses the forward's assignment logits and fp16 image of x instead of recomputing them (same bits,
 * three kernels fewer).  NULL: everything is recomputed from x. */
FGC_API int fgc_build_reverse_padded(const int32_t* rev_ptr, const int32_t* rev_edge, int B, int N, int K,
                             int Kr, int32_t* radj, void* stream);
FGC_API int fgc_conv_bwd_planned(const fgc_conv_shape* s, const float* gy, const float* x,
                         const int32_t* adj, const void* plan /* forward tile plan or NULL */,
                         const int32_t* rev_ptr, const int32_t* rev_edge, const int32_t* radj,
                         int Kr, const void* rplan, const float* W0,
                         const float* u, const float* v, const float* c, float* gx, float* gW0,
                         float* gb, float* gu, float* gv, float* gc, int bias_mask,
                         const void* fwd_workspace, size_t fwd_workspace_bytes, void* workspace,
                         size_t workspace_bytes, void* stream);

/* debug/parity helper: the gathered-neighbour tensor concat([0],x)[adj] -> out[B,N,K,C]
 * (reference Code/model.py:380-399).  Bit-exact by construction. */
FGC_API int fgc_gather_rows(const float* x, const int32_t* adj, float* out, int B, int N, int K, int C,
                    void* stream);
/* debug/parity helper: assignments q[B,N,K,M] (reference Code/model.py:74-95) */
FGC_API int fgc_assignments(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* u,
                    const float* v, const float* c, float* q, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- pooling / unpooling / pointwise
 * reference Code/model.py:779-825 (custom_binary_tree_pooling, custom_upsampling), :828-830 */
FGC_API int fgc_pool_max(const float* x, float* y, int64_t rows_out, int C, int group,
                 void* stream); /* y[r] = max over x[group*r .. group*r+group-1] */
/* gradient of reduce_max as TF defines it: split equally among tied maxima */
FGC_API int fgc_pool_max_bwd(const float* gy, const float* x, const float* y, float* gx, int64_t rows_out,
                     int C, int group, void* stream);
FGC_API int fgc_pool_avg_ignore_zeros(const float* x, float* y, int B, int64_t rows_in, int C, int steps,
                              void* stream);
FGC_API int fgc_upsample(const float* x, float* y, int64_t rows_in, int C, int group, void* stream);
FGC_API int fgc_upsample_bwd(const float* gy, float* gx, int64_t rows_in, int C, int group, void* stream);
FGC_API int fgc_lrelu(const float* x, float* y, int64_t n, float alpha, void* stream);
FGC_API int fgc_lrelu_bwd(const float* gy, const float* x_pre, float* gx, int64_t n, float alpha,
                  void* stream);
/* concat along channels: y[r] = [a[r,0:Ca] | b[r,0:Cb]]  (tf.concat at model.py:909,929) */
FGC_API int fgc_concat2(const float* a, const float* b, float* y, int64_t rows, int Ca, int Cb,
                void* stream);
FGC_API int fgc_split2(const float* gy, float* ga, float* gb, int64_t rows, int Ca, int Cb, void* stream);
/* rows permutation gather: y[r] = x[idx[r]] (host fancy indexing at dataClasses.py:142,
 * train.py:117-121) */
FGC_API int fgc_gather_perm(const float* x, const int32_t* idx, float* y, int64_t rows_out, int C,
                    void* stream);
/* dst[ids[i]][0..C) = src[ids[i]][0..C) for i < n: rows pushed into another buffer at the same indices.  dst may be a
 * peer GPU's mapping of the same tensor (symmetric memory): the per-sweep exchange of the vertex-sharded
 * update_position2 (reference Code/train.py:1467-1557 runs it on one device) as direct NVLink stores. */
FGC_API int fgc_push_rows(const float* src, float* dst, const int64_t* ids, int64_t n, int C, void* stream);

/* ---------------------------------------------------------------- per-facet linear layers
 * reference Code/model.py:763-769 (custom_lin): y = x @ W + b, W[Cin,Cout] */
FGC_API int fgc_lin_fwd(const float* x, const float* W, const float* b, float* y, int64_t rows, int Cin,
                int Cout, int act, float alpha, void* stream);
FGC_API int fgc_lin_bwd(const float* gy, const float* x, const float* W, float* gx, float* gW, float* gb,
                int64_t rows, int Cin, int Cout, void* workspace, size_t workspace_bytes,
                void* stream);
FGC_API size_t fgc_lin_bwd_workspace(int64_t rows, int Cin, int Cout);
/* fused regression head (model.py:936-941): y = lrelu(x@W1+b1) @ W2 + b2 without materialising
 * the hidden activation.  W1[Cin,H], W2[H,Cout], Cout <= 4.  The network's shape (Cin = 32, H = 1024,
 * Cout = 3) runs on tcgen05 from 2 048 rows on and then needs fgc_mlp_head_workspace() bytes for the
 * fp16 image of W1; workspace may be NULL otherwise. */
FGC_API size_t fgc_mlp_head_workspace(int64_t rows, int Cin, int H, int Cout);
FGC_API int fgc_mlp_head_fwd(const float* x, const float* W1, const float* b1, const float* W2,
                     const float* b2, float* y, int64_t rows, int Cin, int H, int Cout, float alpha,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- the network's inference forward
 * reference Code/model.py:837-946, get_model_reg_multi_scale(x, adjs, keep_prob, multiScale=False): the
 * 3-level U-Net as one call.  x[B,N0,6], adj0[B,N0,K], adj1[B,N0/4,K], adj2[B,N0/16,K] (reference layout:
 * 1-indexed, 0 = padding, ids local to the batch element), N0 a multiple of 16; params = the
 * fgc_net_param_count() = 44 parameter tensors in the reference's variable-creation order
 * (W0[9,Cout,Cin], b, u, c, v for conv1 conv2 conv3 dconv3 upconv2 dconv2 upconv1 dconv1, then W[32,1024], b,
 * W[1024,3], b of the head: model.py:853-941); y[B,N0,3] = the network output BEFORE normalizeTensor.
 * fgc_net_prepare builds the tensor-core weight images once per set of parameters into a caller-owned
 * buffer of fgc_net_prepared_bytes() bytes.  Pooling, up-sampling and concatenation are fused into the
 * convolutions; batch elements are independent patches. */
FGC_API int fgc_net_param_count(void);
FGC_API size_t fgc_net_prepared_bytes(void);
FGC_API int fgc_net_prepare(const float* const* params, int nparams, void* prepared, size_t prepared_bytes, void* stream);
FGC_API size_t fgc_net_fwd_workspace(int B, int N0, int K);
FGC_API int fgc_net_fwd(int B, int N0, int K, const float* x, const int32_t* adj0, const int32_t* adj1, const int32_t* adj2,
                const float* const* params, int nparams, const void* prepared, float* y, void* workspace,
                size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- output normalisation & loss
 * reference Code/utils.py:1700-1715 (normalizeTensor) over x[rows,3] of ONE patch: global
 * mean-abs rescale, then row L2 normalisation with the three 1e-5 epsilons. */
FGC_API size_t fgc_normalize_workspace(int64_t rows);
FGC_API int fgc_normalize_rows(const float* x, float* y, int64_t rows, void* workspace,
                       size_t workspace_bytes, void* stream);
/* the same per patch for a batch of padded patches: element b of x[B][stride_rows][3] is normalised over its
 * first counts[b] rows (device int32) with its own global mean; the padding rows of y are zeroed.
 * workspace: B * 128 + 512 bytes. */
FGC_API int fgc_normalize_rows_segmented(const float* x, float* y, int B, int64_t stride_rows, const int32_t* counts,
                                 void* workspace, size_t workspace_bytes, void* stream);
FGC_API int fgc_normalize_rows_bwd(const float* gy, const float* x, float* gx, int64_t rows, void* workspace,
                           size_t workspace_bytes, void* stream);
/* reference Code/train.py:1272-1294 (faceNormalsLoss): loss[0] = mean angle in degrees over real
 * rows; optional gradient w.r.t. fn scaled by gscale. */
FGC_API int fgc_face_normals_loss(const float* fn, const float* gt, float* loss, float* gfn /*nullable*/,
                          int64_t rows, float gscale, void* workspace, size_t workspace_bytes,
                          void* stream);

/* reference Code/train.py:1332-1370 (accuracyLoss, mode 0) and :1373-1424 (fullLoss, mode 1); sampledAccuracyLoss
 * (:1428-1464) is mode 0 over the batch flattened into one point set with ind0 = NULL.
 *   p0[batch][n0][3] predicted points, p1[batch][n1][3] ground truth; ind0[ns0] / ind1[ns1] int32 sample rows (device;
 *   NULL = every row; ind1 is read by mode 1 only).
 *   loss[0] = 1000 * (mean over (batch, sample) of the thresholded nearest-point distance P0 -> P1
 *                     + mean of the nearest-point distance P1 -> P0), thresholds 5 / none (mode 0), 5000 / 5000 (mode 1);
 *   gp0[batch][n0][3] (nullable) = d loss / d p0, accumulated in a fixed order (bit-reproducible).
 * Nothing of size n0 x n1 is materialised. */
FGC_API size_t fgc_point_set_loss_workspace(int batch, int64_t n0, int64_t n1, int ns0, int ns1);
FGC_API int fgc_point_set_loss(const float* p0, const float* p1, int batch, int64_t n0, int64_t n1,
                       const int32_t* ind0 /*nullable*/, int ns0, const int32_t* ind1 /*nullable*/, int ns1,
                       int mode, float* loss, float* gp0 /*nullable*/, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---------------------------------------------------------------- vertex position updates
 * reference Code/train.py:1467-1557 (update_position2): `iters` Jacobi sweeps,
 *   x_i += (1/18) sum_{e in v_edges[i]} sum_{w in (v1,v2)(e)} sum_{f in (f1,f2)(e)} n_f (n_f.(x_w - x_i))
 * x_in/x_out[V,3], normals[F,3], edge_map[E,4] = (v1,v2,f1,f2 | -1), v_edges[V,max_edges] | -1.
 * workspace holds the ping-pong buffer. */
FGC_API size_t fgc_vertex_update_workspace(int64_t V);
FGC_API int fgc_vertex_update_edges(const float* x_in, float* x_out, const float* normals,
                            const int32_t* edge_map, const int32_t* v_edges, int64_t V, int64_t F,
                            int64_t E, int max_edges, int iters, float lambda, void* workspace,
                            size_t workspace_bytes, void* stream);
/* ONE sweep of the same update over the vertex range [v_begin, v_end): reads every row of x_in, writes rows
 * v_begin .. v_end-1 of x_out (x_out != x_in).  Lets N ranks shard the vertices of a large scan (BASELINE config C5,
 * reference Code/train.py:1467-1557 run on one device there): each sweeps its range, one all-gather per sweep. */
FGC_API int fgc_vertex_update_edges_range(const float* x_in, float* x_out, const float* normals,
                                  const int32_t* edge_map, const int32_t* v_edges, int64_t V, int64_t F,
                                  int64_t E, int max_edges, int64_t v_begin, int64_t v_end, float lambda,
                                  void* stream);
/* reference Code/train.py:1668-1798 (update_position_MS + updateFacesCenter) for ONE scale:
 * `iters` sweeps of  x_v += (1/#faces_v) sum_{f in v_faces[v]} n_F (n_F.(c_F - x_v)),
 * F = f >> (2*scale) (floor, -1 stays padding), c = face centres pooled `scale` times with
 * avg_ignore_zeros.  faces[N0,3] vertex ids (-1 rows = fake nodes), normals[N0 >> 2*scale, 3]. */
FGC_API size_t fgc_vertex_update_ms_workspace(int64_t V, int64_t N0);
FGC_API int fgc_vertex_update_ms(const float* x_in, float* x_out, const float* normals,
                         const int32_t* faces, const int32_t* v_faces, int64_t V, int64_t N0,
                         int max_faces, int scale, int steps, int iters, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Backward of fgc_vertex_update_ms (what TensorFlow's autodiff derives for the sweeps of Code/train.py:1724-1758 inside
 * trainAccuracyNet / trainDoubleLossNet, :771-781, :1089-1102): g_out = dL/dx_out -> g_in[V,3] = dL/dx_in and
 * g_normals[N0 >> levels, 3] = dL/dnormals (overwritten).  The forward trajectory is recomputed into the workspace
 * (iters * V * 12 bytes).  Index lists, built once per mesh by the caller, make every sum a fixed-order gather:
 *   slot_ptr[Fs+1] / slot_id[]       the v_faces slots (v * max_faces + k) that map to coarse face F, ascending;
 *   vert_ptr[V+1]  / vert_corner[]   the corners (f * 3 + t) of `faces` that are vertex v, ascending. */
FGC_API size_t fgc_vertex_update_ms_bwd_workspace(int64_t V, int64_t N0, int max_faces, int iters);
FGC_API int fgc_vertex_update_ms_bwd(const float* x_in, const float* normals, const int32_t* faces, const int32_t* v_faces,
                             int64_t V, int64_t N0, int max_faces, int scale, int steps, int iters,
                             const int32_t* slot_ptr, const int32_t* slot_id, const int32_t* vert_ptr,
                             const int32_t* vert_corner, const float* g_out, float* g_in, float* g_normals,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- index builders
 * GPU versions of the host loops that produce the index tensors above; pure integer work, outputs
 * bit-identical to the reference's (pinned by tests/golden/index_layouts.npz, which the reference
 * functions generated).  faces[nf][3] int32 vertex ids (rows of -1 = fake nodes are skipped).
 *
 * fgc_build_faces_adj: adj[nf][K] in getFacesLargeAdj layout (reference Code/utils.py:243-295: row f =
 * f+1, then the other faces around each vertex of f -- vertices in increasing id, faces in increasing id,
 * edge-adjacent faces therefore twice -- cut after K-1 appends, 0 padded) and/or v_faces[nv][kv] in
 * getVerticesFaces layout (Code/utils.py:370-395: faces around each vertex, increasing, -1 padded).
 * Either output may be NULL.  Synchronises the stream (returns an error for vertex ids outside
 * 0..nv-1 or a vertex with more than kv faces).
 *
 * fgc_build_edge_maps: e_map[E][4] = (va, vb, first face, last later face | -1) with edges numbered by
 * first appearance over faces and slots (v1v2, v1v3, v2v3), va/vb as in that first appearance, and
 * v_edges[nv][max_edges] = edge ids around each vertex, increasing, -1 padded (getEdgeMap,
 * Code/utils.py:91-183).  e_map must hold 3*nf rows; *num_edges (host) receives E.  Synchronises. */
/* features[nf][6] = [unit face normal | barycentre of the vertices divided by the bounding-box diagonal]
 * (reference computeFacesNormals Code/utils.py:63-68 with the two-pass normalize of :26-35, and
 * getTrianglesBarycenter :1264-1294; double arithmetic like the reference's NumPy, one rounding to fp32);
 * rows of faces with an id outside 0..nv-1 (fake nodes) get zeros.  workspace: 1 KB. */
FGC_API int fgc_face_features(const float* verts, const int32_t* faces, int64_t nf, int64_t nv, int normalize,
                      float* features, void* workspace, size_t workspace_bytes, void* stream);
FGC_API size_t fgc_faces_adj_workspace(int64_t nf, int64_t nv);
FGC_API int fgc_build_faces_adj(const int32_t* faces, int64_t nf, int64_t nv, int K, int32_t* adj,
                        int32_t* v_faces, int kv, void* workspace, size_t workspace_bytes, void* stream);
FGC_API size_t fgc_edge_maps_workspace(int64_t nf, int64_t nv);
FGC_API int fgc_build_edge_maps(const int32_t* faces, int64_t nf, int64_t nv, int max_edges, int32_t* e_map,
                        int64_t* num_edges, int32_t* v_edges, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Host-only (no device work, HOST pointers): the greedy heavy-edge pairing of one coarsening level, reference
 * Code/lib/coarsening.py:135-194 `metis_one_level`.  (row, col, val)[nnz] is the weighted adjacency sorted by
 * row, n = row[nnz-1] + 1 nodes, order[0..n) the visiting order, weights[n] the node degrees.  A node visited
 * unpaired takes the unpaired neighbour with the largest val * (1/weights[node] + 1/weights[neighbour]) (first
 * one on ties, none when every score is 0) and both get the next cluster id.  precision 32 evaluates scores
 * and their sum in float (what the reference's expressions give on float32 inputs under NumPy >= 2), 64 in
 * double (NumPy 1.x promotion).  cluster_id[n], *total_assoc = sum of the winning scores, *n_clusters. */
FGC_API int fgc_greedy_pairing(const int32_t* row, const int32_t* col, const float* val, int64_t nnz,
                       const int64_t* order, int64_t n_order, const float* weights, int32_t n, int precision,
                       int32_t* cluster_id, double* total_assoc, int32_t* n_clusters);

/* Host-only: breadth-first growth of one patch, reference Code/utils.py:1508-1696 `getGraphPatch_wMask`.
 * adj[n][K] 1-based / 0-padded with the node itself in column 0, mask[n] != 0 for nodes earlier patches own.
 * Grows from `seed` until nodes_num nodes are reached (owned nodes join but wait in a second queue that is
 * only expanded while the patch is below min_patch), then closes the rows still queued with the neighbours
 * that are inside.  adj_out[capacity][K] (capacity >= max(nodes_num, min_patch) + K) gets patch-local lists,
 * old_index[capacity] the original ids in local order, *patch_nodes the size, *next_seed a not-owned node
 * adjacent to the patch or -1. */
FGC_API int fgc_grow_patch(const int32_t* adj, int64_t n, int K, int64_t nodes_num, int64_t seed,
                   const uint8_t* mask, int64_t min_patch, int32_t* adj_out, int64_t capacity,
                   int64_t* old_index, int64_t* patch_nodes, int64_t* next_seed);

/* ---------------------------------------------------------------- host-buffer entry points
 * What a non-torch FFI binding calls: pinned or pageable HOST pointers in, HOST pointers out;
 * the call allocates device buffers (cached per thread), copies H2D, runs the kernels on
 * `device`, copies D2H and synchronises.  Used for the end-to-end (`e2e`) measurement. */
FGC_API int fgc_conv_fwd_host(const fgc_conv_shape* s, const float* x, const int32_t* adj, const float* W0,
                      const float* b, const float* u, const float* v, const float* c, float* y,
                      int bias_mask, int act, float alpha, int device);
FGC_API int fgc_conv_fwd_bwd_host(const fgc_conv_shape* s, const float* x, const int32_t* adj,
                          const float* gy, const float* W0, const float* b, const float* u,
                          const float* v, const float* c, float* y, float* gx, float* gW0, float* gb,
                          float* gu, float* gv, float* gc, int bias_mask, int device);
FGC_API void fgc_host_release(void); /* frees the per-thread device cache of the _host entry points */

#ifdef __cplusplus
}
#endif
#endif /* FACETCONV_B200_H_ */
