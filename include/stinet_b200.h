/* stinet_b200.h -- C ABI of the B200-native STINet hot path (libstinet_b200.so).
 *
 * The reference (johnpeterflynn/surface-texture-inpainting-net) has NO FFI of its own: its hot path reaches native
 * code only through torch_geometric / torch_scatter / ATen.  Each entry point below therefore names the reference
 * call site whose arithmetic it replaces (paths relative to the reference repo root).  The Python modules in
 * stinet_b200/models mirror the reference's module API and call these functions through ctypes (stinet_b200/_abi.py).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch allocates inputs, outputs and workspaces);
 *    the library never allocates, frees or synchronises, and launches only on the given stream;
 *  - matrices are row-major fp32 with an explicit leading dimension `ld*` counted in ELEMENTS (so column slices of a
 *    wider buffer can be passed); index arrays produced by the library are int32; index arrays coming from the
 *    reference's data pipeline are int64 (edge_index, trace maps) and are narrowed once by stinet_csr_build;
 *  - return value: 0 = ok, <0 = error (STINET_ERR_*); stinet_last_error() gives a thread-local message;
 *    no exception crosses the ABI.  Functions are re-entrant and stream-ordered; the library has no mutable global
 *    state apart from per-kernel function attributes set once.
 *  - `status` (nullable, int32[1] on the device): kernels OR a bit into it on data-dependent errors
 *    (bit0: index out of range) because reporting them through the return value would need a sync.
 */
#ifndef STINET_B200_H
#define STINET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STINET_ABI_VERSION 1

#define STINET_OK 0
#define STINET_ERR_ARG (-1)          /* null pointer, negative size, bad enum */
#define STINET_ERR_UNSUPPORTED (-2)  /* shape / alignment the kernels do not cover */
#define STINET_ERR_CUDA (-3)         /* launch failed; message holds cudaGetErrorString */
#define STINET_ERR_WORKSPACE (-4)    /* workspace too small */

typedef void* stinet_stream_t; /* cudaStream_t */

enum { STINET_REDUCE_ADD = 0, STINET_REDUCE_MEAN = 1, STINET_REDUCE_MAX = 2 };
/* arithmetic of the dense layers (storage at the ABI is fp32, accumulation is always fp32):
 *   FP32      tcgen05 kind::tf32, every product as hi*hi + hi*lo + lo*hi of TF32-rounded halves (fp32-class error,
 *             the reference's 1e-5 bar); operands TMA cannot address (row pitch % 16 B != 0) run on FFMA tiles
 *   BF16      operands cast to bf16 in the workspace, tcgen05 kind::f16 (2e-2 per operator)
 *   BF16X3    operands split into two bf16 planes x = hi + lo in the workspace, tcgen05 kind::f16 on
 *             hi*hi + hi*lo + lo*hi (~2^-16 per product: keeps a whole network within the 2e-2 bar)
 *   FP32_SIMT FFMA tiles only (cross-check of the tensor-core paths)
 *   TF32      one tcgen05 kind::tf32 pass (~1e-3) */
enum { STINET_PREC_FP32 = 0, STINET_PREC_BF16 = 1, STINET_PREC_FP32_SIMT = 2, STINET_PREC_TF32 = 3,
       STINET_PREC_BF16X3 = 4 };
enum { STINET_ACT_NONE = 0, STINET_ACT_ELU = 1 };

int stinet_abi_version(void);
const char* stinet_last_error(void);
/* 1 iff the running device is compute capability 10.x (the only target this library is built for). */
int stinet_device_ok(void);
/* process-wide number of kernels this library has launched so far (bench.py's `gpu_launches`) */
long long stinet_launch_count(void);

/* ---- structure: CSR builder (no reference counterpart; replaces the COO edge_index that PyG's
 * MessagePassing.propagate consumes, call sites models/modules/edge_conv_filter.py:57, sage_conv_filter.py:75 and
 * the trace maps consumed by torch_scatter at models/surfacetextureinpaintingnet.py:384,386,422).
 * Groups the positions 0..n_items-1 by key[pos] in [0,n_rows), STABLY (original order kept inside a row, which
 * reproduces torch_scatter's CPU summation / tie-break order):
 *   rowptr[n_rows+1], perm[n_items] (positions grouped by key), col[n_items] = (int32) other[perm[k]] (if other != NULL),
 *   key32[n_items] = (int32) key[pos] (if key32 != NULL). */
size_t stinet_csr_workspace_bytes(int64_t n_rows, int64_t n_items);
int stinet_csr_build(const int64_t* key, const int64_t* other, int64_t n_items, int64_t n_rows, int32_t* rowptr,
                     int32_t* perm, int32_t* col, int32_t* key32, int32_t* status, void* workspace,
                     size_t workspace_bytes, stinet_stream_t stream);
/* tpos_s[k_s] = position in the by-target CSR of the edge at position k_s of the by-source CSR (eid_t / eid_s = the
 * two `perm` outputs of stinet_csr_build over the same edge list); scratch: n_items int32. */
int stinet_csr_cross_positions(const int32_t* eid_t, const int32_t* eid_s, int64_t n_items, int32_t* tpos_s,
                               int32_t* scratch, stinet_stream_t stream);
/* Block-diagonal batching of structures built once per sample (the offsets are those of
 * HierarchicalData.__inc__, utils/data_utils.py:29-42, which the reference applies to the COO tensors in collate):
 *   dst[dst_off[p] + i] = src[p][i] + add[p]   for i < len[p], p < n_parts.
 * `src`, `len`, `dst_off`, `add` are HOST arrays (they travel in the launch parameters); src[p] and dst are device
 * pointers.  rowptr parts pass len = n_rows (the last part n_rows + 1) and add = edge offset; col / member parts pass
 * add = vertex offset. */
int stinet_concat_i32(const int32_t* const* src, const int64_t* len, const int64_t* dst_off, const int32_t* add,
                      int n_parts, int32_t* dst, stinet_stream_t stream);
/* ---- aggregation over CSR rows (replaces PyG propagate's scatter(msg, edge_index[1], reduce=...) =
 * torch_scatter.scatter_{sum,mean,max}; call sites edge_conv_filter.py:57 (mean), sage_conv_filter.py:75 (mean),
 * utils/metrics/graph_metrics.py:12 (add)).  out[i,:] = reduce_{k in row i} x[col[k],:]; mean = sum/max(deg,1);
 * rows without entries give 0.  max: arg[i,c] = ORIGINAL edge id (eid[k]) of the first edge attaining the maximum,
 * n_items for empty rows (torch_scatter CPU semantics). */
int stinet_aggregate_fwd(const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                         int64_t n_rows, int64_t n_items, int64_t channels, int reduce, float* out, int64_t ldo,
                         int32_t* arg, stinet_stream_t stream);
/* dx[j,:] = sum over out-edges (j->i) of g[i,:] * w, w = 1 (add), 1/max(deg_t(i),1) (mean), [arg[i,c]==eid] (max).
 * rowptr_s/col_s/eid_s: CSR grouped by SOURCE (col_s = target of each out-edge); rowptr_t: CSR by target (degrees). */
int stinet_aggregate_bwd(const float* g, int64_t ldg, const int32_t* rowptr_s, const int32_t* col_s,
                         const int32_t* eid_s, const int32_t* rowptr_t, const int32_t* arg, int64_t n_src_rows,
                         int64_t channels, int reduce, float* dx, int64_t lddx, stinet_stream_t stream);

/* ---- fused EdgeConv message stage (replaces x_j/x_i gathers + cat + Linear + ReLU + scatter-mean of PyG
 * EdgeConv.message/aggregate, models/modules/edge_conv_filter.py:46-57, edge_conv_translation_invariance.py:19-21,
 * in the hoisted form  nn.0([x_i || x_j-x_i]) = P_i + Q_j,  P = X(Wa-Wb)^T + b, Q = X Wb^T):
 *   hid[i,:] = (1/max(deg_i,1)) * sum_{j->i} relu(P[i,:] + Q[j,:]) */
int stinet_edge_message_fwd(const float* P, int64_t ldp, const float* Q, int64_t ldq, const int32_t* rowptr_t,
                            const int32_t* col_t, int64_t n_rows, int64_t hidden, float* hid, int64_t ldh,
                            stinet_stream_t stream);
/* dP[i,:] = (1/deg_i) * sum_{j->i} dhid[i,:] * [P_i+Q_j > 0]                    (CSR by target) */
int stinet_edge_message_bwd_target(const float* P, int64_t ldp, const float* Q, int64_t ldq, const float* dhid,
                                   int64_t ldd, const int32_t* rowptr_t, const int32_t* col_t, int64_t n_rows,
                                   int64_t hidden, float* dP, int64_t lddp, stinet_stream_t stream);
/* dQ[j,:] = sum_{j->i} (1/deg_i) * dhid[i,:] * [P_i+Q_j > 0]                    (CSR by source) */
int stinet_edge_message_bwd_source(const float* P, int64_t ldp, const float* Q, int64_t ldq, const float* dhid,
                                   int64_t ldd, const int32_t* rowptr_t, const int32_t* rowptr_s,
                                   const int32_t* col_s, int64_t n_rows, int64_t hidden, float* dQ, int64_t lddq,
                                   stinet_stream_t stream);

/* The same stage with the ReLU decisions saved for backward (hidden % 4 == 0, 16-byte aligned rows): forward also
 * writes mask[k * hidden/4 + c4] = one byte whose low nibble holds the decision bits [P_i+Q_j > 0] of channels
 * 4*c4..4*c4+3 for the edge at POSITION k of the by-target CSR (n_edges * hidden/4 bytes, written and re-read as a
 * stream).  The backward kernels read the masks instead of re-evaluating P_i + Q_j: dP needs no neighbour rows at
 * all, dQ gathers dhid only; tpos_s[k_s] = by-target position of by-source entry k_s (stinet_csr_cross_positions).
 * Results are bit-identical to the recomputing entry points above. */
int stinet_edge_message_fwd_mask(const float* P, int64_t ldp, const float* Q, int64_t ldq, const int32_t* rowptr_t,
                                 const int32_t* col_t, int64_t n_rows, int64_t hidden, float* hid, int64_t ldh,
                                 void* mask, stinet_stream_t stream);
int stinet_edge_message_bwd_target_mask(const float* dhid, int64_t ldd, const int32_t* rowptr_t, const void* mask,
                                        int64_t n_rows, int64_t hidden, float* dP, int64_t lddp,
                                        stinet_stream_t stream);
int stinet_edge_message_bwd_source_mask(const float* dhid, int64_t ldd, const int32_t* rowptr_t,
                                        const int32_t* rowptr_s, const int32_t* col_s, const int32_t* tpos_s,
                                        const void* mask, int64_t n_rows, int64_t hidden, float* dQ, int64_t lddq,
                                        stinet_stream_t stream);

/* Parameters of the hoisted first layer from nn.0's own W [hidden, kin] / b [hidden] (kin = 2*din, or din for
 * EdgeConvTransInv):  Wcat [2*hidden, din] = [Wa - Wb ; Wb]  (trans_inv: [-W ; W]),  bcat [2*hidden] = [b ; 0]
 * (b / bcat nullable together), and the matching gradient fold  dW = [dP-part | dQ-part - dP-part], db = dbcat[:hidden].
 * Wcat / dWcat are dense (ld = din). */
int stinet_edgeconv_hoist_fwd(const float* W, int64_t ldw, const float* b, int64_t hidden, int64_t din, int trans_inv,
                              float* Wcat, float* bcat, stinet_stream_t stream);
int stinet_edgeconv_hoist_bwd(const float* dWcat, const float* dbcat, int64_t hidden, int64_t din, int trans_inv,
                              float* dW, int64_t ldw, float* db, stinet_stream_t stream);

/* ---- trace-map pooling / unpooling (replaces SurfaceTextureInpaintingNet._pooling / _unpooling,
 * models/surfacetextureinpaintingnet.py:382-391 = torch_scatter.scatter_max / scatter_mean / x[trace]).
 * Cluster CSR: rowptr_c[n_coarse+1], member[n_fine] = fine vertex ids grouped by cluster in ascending order
 * (stinet_csr_build with key = trace).  pool-max: arg[c,ch] = LOWEST fine id attaining the max, n_fine if empty. */
int stinet_pool_max_fwd(const float* x, int64_t ldx, const int32_t* rowptr_c, const int32_t* member, int64_t n_fine,
                        int64_t n_coarse, int64_t channels, float* out, int64_t ldo, int32_t* arg,
                        stinet_stream_t stream);
int stinet_pool_max_bwd(const float* g, int64_t ldg, const int32_t* arg, const int32_t* trace32, int64_t n_fine,
                        int64_t channels, float* dx, int64_t lddx, stinet_stream_t stream);
int stinet_pool_mean_fwd(const float* x, int64_t ldx, const int32_t* rowptr_c, const int32_t* member,
                         int64_t n_coarse, int64_t channels, float* out, int64_t ldo, stinet_stream_t stream);
int stinet_pool_mean_bwd(const float* g, int64_t ldg, const int32_t* rowptr_c, const int32_t* trace32,
                         int64_t n_fine, int64_t channels, float* dx, int64_t lddx, stinet_stream_t stream);
/* integer variant used for the per-vertex graph id (`batch = scatter_max(batch, trace)`, :422); empty cluster -> 0 */
int stinet_pool_max_i32(const int32_t* v, const int32_t* rowptr_c, const int32_t* member, int64_t n_coarse,
                        int32_t* out, stinet_stream_t stream);
/* out[i, col_off:col_off+channels] = xc[trace32[i], :]  (unpool; a col_off/ldo pair gives the skip-concat variant of
 * models/singleconvmeshnet.py:140-141) */
int stinet_unpool_fwd(const float* xc, int64_t ldc, const int32_t* trace32, int64_t n_fine, int64_t channels,
                      float* out, int64_t ldo, stinet_stream_t stream);
/* dxc[c,:] = sum_{i in cluster c} g[i,:]   (ascending member order; replaces index_add_ atomics) */
int stinet_unpool_bwd(const float* g, int64_t ldg, const int32_t* rowptr_c, const int32_t* member, int64_t n_coarse,
                      int64_t channels, float* dxc, int64_t ldd, stinet_stream_t stream);

/* ---- per-graph instance norm (replaces FastInstanceNorm.forward, models/modules/fastinstancenorm.py:42-107)
 * Rows are partitioned twice, exactly as the reference does: `slice_ptr` (= torch.linspace(0,N,B+1), :53) bounds the
 * ranges the sums run over, `gid[r]` (= batch[r]; NULL = all zero) selects which mean / rstd a row uses (:73,:99),
 * and `cnt[s]` (= degree(batch).clamp(min=1), :60) is the divisor.  biased variance, eps inside the sqrt.
 *   mean[s,c] = sum_{r in slice s} x[r,c] / cnt[s];  var[s,c] = sum_{r in slice s} (x[r,c]-mean[gid[r],c])^2 / cnt[s]
 *   rstd = 1/sqrt(var+eps). */
size_t stinet_segnorm_workspace_bytes(int64_t max_seg_rows, int64_t channels, int64_t n_seg);
/* max_seg_rows: host-known upper bound of the slice lengths (sizes the grid without a device read).
 * gid == NULL: a row uses the statistics of the slice that contains it. */
int stinet_segnorm_stats(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, int64_t n_seg,
                         int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt, const int32_t* gid,
                         float eps, float* mean, float* rstd, void* workspace, size_t workspace_bytes,
                         stinet_stream_t stream);
/* One-call forward for batches whose slices ARE the graphs (equal-size batches, and batch=None = one slice; the only
 * cases the reference trains on):  out = residual + act((x - mean[s]) * rstd[s]),  mean / rstd [n_seg, channels] are
 * written for the backward.  Slices of at most 16384 rows run as ONE kernel (a thread-block cluster per (slice,
 * 32-channel slab), partial sums exchanged through distributed shared memory in rank order, rows kept in registers
 * between the passes when they fit); longer slices run stats + a slice-indexed apply kernel. */
int stinet_segnorm_fwd(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, int64_t n_seg,
                       int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt, float eps,
                       const float* residual, int64_t ldr, int act, float* out, int64_t ldo, float* mean, float* rstd,
                       float* amax_out, void* out_hi, void* out_lo, int64_t ldp, const float* res_amax, int32_t* out_exp,
                       void* workspace, size_t workspace_bytes, stinet_stream_t stream);
/* out_hi / out_lo / ldp / out_exp (nullable group): the same pass also writes `out` as fp16 operand planes for the dense
 * layer that reads it next (no split pass later).  Their scale comes from the bound max|residual| + sqrt(longest slice)
 * (res_amax: float[1] holding max|residual| or a bound of it; required when a residual is given).  Vector path only.
 * amax_out (nullable, float[1], here and in stinet_segnorm_bwd): the kernels also leave max|out| (max|dx|) there -- the
 * plane scale of the dense layer that reads the result next (stinet_f16_split) -- at no extra pass over the data. */
/* out = residual + act((x - mean[g]) * rstd[g])   (residual nullable; act = STINET_ACT_*): the tail of
 * GraphResnetBlock.forward, models/surfacetextureinpaintingnet.py:510-521.  g = gid[r] (NULL: segment 0);
 * mean == rstd == NULL: identity norm (norm_type 'none', :257-263). */
int stinet_segnorm_apply(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, const int32_t* gid,
                         const float* mean, const float* rstd, const float* residual, int64_t ldr, int act,
                         float* out, int64_t ldo, stinet_stream_t stream);
/* backward of out = act(norm(x)):  dx = rstd*(dz - mean_s(dz) - yhat*mean_s(dz*yhat)), dz = dout*act'(yhat).
 * Requires slices == true segments (gid constant on every slice); otherwise STINET_ERR_UNSUPPORTED is the
 * caller's job to raise (the library cannot see it without a sync).  gid == NULL: a row uses the statistics of the
 * slice that contains it (single cluster kernel for slices of at most 16384 rows, as in stinet_segnorm_fwd).
 * dout_amax (nullable, float[1]): max|dout|, taken by the pass that reads dout anyway -- dout is also the gradient of the
 * block's shortcut branch (x' = shortcut(x) + ..., :521), whose dense layer then needs no reduction pass of its own. */
int stinet_segnorm_bwd(const float* x, int64_t ldx, const float* dout, int64_t ldg, int64_t n_rows, int64_t channels,
                       int64_t n_seg, int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt,
                       const int32_t* gid, const float* mean, const float* rstd, int act, float* dx, int64_t lddx,
                       float* amax_out, float* dout_amax, void* workspace, size_t workspace_bytes,
                       stinet_stream_t stream);

/* ---- affine segmented norms: the alternative norm_type modules of the network and the BatchNorm1d
 * over the EDGES inside SingleConvMeshNet's message MLP, all of the form
 *     y = gamma * (x - alpha * m[g]) * r[g] + beta,   r = 1 / sqrt(v + eps),   m = slice sum / cnt
 *   kind 0: v = slice sum of (x - m)^2 / cnt    -- torch_geometric BatchNorm == nn.BatchNorm1d over node rows
 *           (models/surfacetextureinpaintingnet.py:236-241 BatchNorm2Param; edge_conv_filter.py:34-44), one slice = the batch
 *   kind 1: v = slice sum of x^2 / cnt          -- SingleBatchGraphNorm (models/modules/singlebatchgroupnorm.py:44-71), which
 *           takes the second moment of the UN-shifted x (:66-68); slices = the reference's linspace slices, cnt = their length
 * alpha / gamma / beta are per-channel parameters (NULL: 1 / 1 / 0).  Statistics are the deterministic two-stage column
 * reductions of stinet_segnorm_stats (no atomics).  gid (int32 per row -> segment; NULL only with one segment).
 *   fwd    writes y, mean[n_seg, C], rstd[n_seg, C]
 *   apply  the elementwise part alone, with given statistics (eval mode: running mean / 1/sqrt(running_var + eps))
 *   bwd    dx, dgamma = sum dy (x - alpha m) r, dbeta = sum dy, dalpha = -gamma sum_s m r sum_slice dy  (any of the four
 *          outputs may be NULL); requires every row's gid to be the slice that contains it
 *   bn_running_update  running_mean / running_var of nn.BatchNorm1d after one training step (unbiased variance) */
size_t stinet_affnorm_workspace_bytes(int64_t max_seg_rows, int64_t channels, int64_t n_seg);
int stinet_affnorm_fwd(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, int64_t n_seg, int64_t max_seg_rows,
                       const int32_t* slice_ptr, const float* cnt, const int32_t* gid, int kind, float eps,
                       const float* alpha, const float* gamma, const float* beta, float* out, int64_t ldo, float* mean,
                       float* rstd, void* workspace, size_t workspace_bytes, stinet_stream_t stream);
int stinet_affnorm_apply(const float* x, int64_t ldx, int64_t n_rows, int64_t channels, const int32_t* gid,
                         const float* mean, const float* rstd, const float* alpha, const float* gamma, const float* beta,
                         float* out, int64_t ldo, stinet_stream_t stream);
int stinet_affnorm_bwd(const float* x, int64_t ldx, const float* dy, int64_t ldg, int64_t n_rows, int64_t channels,
                       int64_t n_seg, int64_t max_seg_rows, const int32_t* slice_ptr, const float* cnt, const int32_t* gid,
                       int kind, const float* mean, const float* rstd, const float* alpha, const float* gamma, float* dx,
                       int64_t lddx, float* dgamma, float* dbeta, float* dalpha, void* workspace, size_t workspace_bytes,
                       stinet_stream_t stream);
int stinet_bn_running_update(const float* mean, const float* rstd, int64_t n_rows, float eps, float momentum,
                             int64_t channels, float* running_mean, float* running_var, stinet_stream_t stream);

/* ---- dense layers (replace torch.nn.Linear inside the message MLP, the shortcut and the head:
 * models/modules/edge_conv_filter.py:46-55, models/surfacetextureinpaintingnet.py:504-505,356-358).
 * fwd:   C[M,N]  = A[M,K] W[N,K]^T + bias[N] * (rowmask ? rowmask[m] > 0 : 1)
 * dgrad: dA[M,K] = dC[M,N] W[N,K]
 * wgrad: dW[N,K] = dC[M,N]^T A[M,K];  dbias[N] = sum_m dC[m,:] * (rowmask ? rowmask[m] > 0 : 1)   (dbias nullable) */
size_t stinet_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int precision);
int stinet_linear_fwd(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                      const int32_t* rowmask, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int precision,
                      void* workspace, size_t workspace_bytes, stinet_stream_t stream);
int stinet_linear_dgrad(const float* dC, int64_t ldc, const float* W, int64_t ldw, float* dA, int64_t lda, int64_t M,
                        int64_t N, int64_t K, int precision, void* workspace, size_t workspace_bytes,
                        stinet_stream_t stream);
int stinet_linear_wgrad(const float* dC, int64_t ldc, const float* A, int64_t lda, const int32_t* rowmask, float* dW,
                        int64_t ldw, float* dbias, int64_t M, int64_t N, int64_t K, int precision, void* workspace,
                        size_t workspace_bytes, stinet_stream_t stream);

/* ---- dense layers on operand PLANES (the fp32-parity path of the dense layers; same reference call sites as above).
 * A matrix x is handed to the tensor cores as two fp16 planes of the scaled matrix x 2^s:
 *     hi = fp16(x 2^s),  lo = fp16((x 2^s - hi) 2^11),  exp = -s            (hi + lo 2^-11 = x 2^s to 22 significand bits)
 * with s picked from amax = max|x| so that amax 2^s lies in [2^14, 2^15): fp16's range is used in full, every element
 * down to amax 2^-28 keeps its 22 bits and smaller ones an absolute error below amax 2^-50.  `amax` may be any upper
 * bound of max|x| within a factor ~2^16 (producers that know a bound skip the reduction).  The GEMMs evaluate
 * hi*hi + (hi*lo + lo*hi) 2^-11 on tcgen05 kind::f16 with fp32 accumulation (two TMEM accumulators, promotion to
 * round-to-nearest registers every 128 reduction elements) and undo the scales in the epilogue: fp32-class results
 * (the 1e-5 parity bar) at the full 16-bit tensor rate.  passes = 3: as described; passes = 1: hi planes only
 * (11 significand bits, the reduced-precision mode; lo pointers may be NULL).
 * Planes are row-major with pitch `ldp` (elements, a multiple of 8), laid out like the matrix they were split from.
 *   f16_amax   amax[0] = max|x| (as an fp32 bit pattern; NaN propagates).  Resets amax first.
 *   f16_split  reads amax[0], writes hi, lo and exp_out[0] = -s.
 *   colsum     dbias[N] = sum_m dC[m,:] * (rowmask ? rowmask[m] > 0 : 1), deterministic two-stage column sums
 *              (workspace: stinet_gemm_workspace_bytes(M, N, 1, STINET_PREC_FP32)). */
int stinet_f16_amax(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* amax, stinet_stream_t stream);
int stinet_f16_split(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* amax, void* hi, void* lo,
                     int64_t ldp, int32_t* exp_out, stinet_stream_t stream);
/* stinet_f16_split and stinet_colsum in ONE pass over x (the backward of a Linear with bias reads dC for both): planes as
 * f16_split, colsum[n] = sum_m x[m,n] * (rowmask ? rowmask[m] > 0 : 1).  cols % 8 == 0, 16-byte aligned rows.
 * workspace: stinet_gemm_workspace_bytes(rows, cols, 1, STINET_PREC_FP32). */
int stinet_f16_split_colsum(const float* x, int64_t ldx, int64_t rows, int64_t cols, const float* amax,
                            const int32_t* rowmask, void* hi, void* lo, int64_t ldp, int32_t* exp_out, float* colsum,
                            void* workspace, size_t workspace_bytes, stinet_stream_t stream);
/* workspace for the three calls below: stinet_gemm_workspace_bytes(M, N, K, STINET_PREC_FP32).
 * amax_out (nullable, float[1]): the GEMM's epilogue also leaves max|C| (C, dA after bias) there, so that the consumer of
 * the result can pick its plane scale without another pass over it. */
int stinet_linear_fwd_f16(const void* A_hi, const void* A_lo, int64_t lda, const int32_t* a_exp, const void* W_hi,
                          const void* W_lo, int64_t ldw, const int32_t* w_exp, const float* bias,
                          const int32_t* rowmask, float* C, int64_t ldc, float* amax_out, int64_t M, int64_t N, int64_t K,
                          int passes, void* workspace, size_t workspace_bytes, stinet_stream_t stream);
int stinet_linear_dgrad_f16(const void* dC_hi, const void* dC_lo, int64_t ldc, const int32_t* c_exp, const void* W_hi,
                            const void* W_lo, int64_t ldw, const int32_t* w_exp, float* dA, int64_t lda, float* amax_out,
                            int64_t M, int64_t N, int64_t K, int passes, void* workspace, size_t workspace_bytes,
                            stinet_stream_t stream);
int stinet_linear_wgrad_f16(const void* dC_hi, const void* dC_lo, int64_t ldc, const int32_t* c_exp, const void* A_hi,
                            const void* A_lo, int64_t lda, const int32_t* a_exp, float* dW, int64_t ldw, int64_t M,
                            int64_t N, int64_t K, int passes, void* workspace, size_t workspace_bytes,
                            stinet_stream_t stream);
int stinet_colsum(const float* dC, int64_t ldc, const int32_t* rowmask, int64_t M, int64_t N, float* dbias,
                  void* workspace, size_t workspace_bytes, stinet_stream_t stream);

/* colsum over a matrix held as planes: out[n] = sum_m x[m,n]  (workspace as stinet_colsum) */
int stinet_colsum_planes(const void* hi, const void* lo, int64_t ldp, const int32_t* exp, int64_t M, int64_t N, float* out,
                         void* workspace, size_t workspace_bytes, stinet_stream_t stream);

/* ---- fused EdgeConv message stage on operand planes (same reference call sites as stinet_edge_message_*): the hidden
 * activations hid and the gradient dPQ = [dP | dQ] are only ever read by tensor-core GEMMs, so these variants write their
 * fp16 planes directly (no fp32 matrix, no split pass).  The plane scale comes from an upper bound of the result that is
 * known before the first element is written:  0 <= hid <= 2 max|PQ|  (pq_amax: the producing GEMM's amax_out) and
 * |dPQ| <= max|dhid| * max(1, dq_factor)  (dhid_amax: the dgrad GEMM's amax_out; dq_factor = max_j sum_{j->i} 1/deg_i, a
 * property of the edge set computed once by stinet_csr_dq_factor).  mask as in stinet_edge_message_fwd_mask (NULL in
 * fwd_planes: inference, no decisions stored).  dpq planes are [n_rows, 2*hidden] with pitch ldp. */
int stinet_csr_dq_factor(const int32_t* rowptr_t, const int32_t* rowptr_s, const int32_t* col_s, int64_t n, float* out,
                         stinet_stream_t stream);
int stinet_edge_message_fwd_planes(const float* P, int64_t ldp, const float* Q, int64_t ldq, const int32_t* rowptr_t,
                                   const int32_t* col_t, int64_t n_rows, int64_t hidden, const float* pq_amax,
                                   void* hid_hi, void* hid_lo, int64_t ldh, int32_t* hid_exp, void* mask,
                                   stinet_stream_t stream);
int stinet_edge_message_bwd_planes(float* dhid, int64_t ldd, const float* dhid_amax, const float* dq_factor,
                                   const int32_t* rowptr_t, const int32_t* rowptr_s, const int32_t* col_s,
                                   const int32_t* tpos_s, const void* mask, int64_t n_rows, int64_t hidden, void* dpq_hi,
                                   void* dpq_lo, int64_t ldp, int32_t* dpq_exp, float* dp_colsum, void* workspace,
                                   size_t workspace_bytes, stinet_stream_t stream);
/* dp_colsum (nullable, float[hidden]): also sum_i dP[i,:], the bias gradient of the hoisted first Linear, accumulated by the
 * kernel that produces dP (per-CTA partials in `workspace`, fixed-order second stage; for hidden > 256 taken from the finished
 * planes instead).  dhid is CONSUMED: its rows are overwritten with dhid[i,:] / deg_i on the way (the second kernel gathers the
 * scaled rows once per out-edge). */
size_t stinet_edge_message_bwd_workspace_bytes(int64_t n_rows, int64_t hidden);

/* ---- operand planes of ALL dense-layer weights of a network in two launches (per step: weights change with every optimizer
 * step, so their planes are re-split once per forward and shared by fwd / dgrad / wgrad of that step).  The caller keeps a
 * table of entries in device memory -- built on the host with stinet_weight_entry_fill (stinet_weight_entry_bytes() bytes
 * each, chunk0 = running sum of the returned chunk counts), copied to the device once -- and calls refresh every step.
 *   kind 0: planes of W [rows, cols];  kind 1: planes of the hoisted first layer [Wa - Wb ; Wb] of EdgeConv (W = [Wa | Wb],
 *   rows = H, cols = din; models/modules/edge_conv_filter.py:46-57) and bcat = [b ; 0];  kind 2: [-W ; W] of EdgeConvTransInv.
 *   Planes of kind 1 / 2 have 2*rows rows.  amax / exp: one float / int32 slot per entry. */
size_t stinet_weight_entry_bytes(void);
long long stinet_weight_entry_fill(void* host_entry, const float* w, int64_t ldw, const float* b, int64_t rows, int64_t cols,
                                 int kind, void* hi, void* lo, int64_t ldp, float* bcat, float* amax, int32_t* exp,
                                 int64_t chunk0);
int stinet_weight_planes_refresh(const void* device_table, int n_entries, int64_t n_chunks, float* amax_slots,
                                 stinet_stream_t stream);

/* ---- the tail of the network and the trainer's loss
 *   head        out[N,3] = tanh(h[N,C] W^T + b), W [3,C] -- reference models/surfacetextureinpaintingnet.py:466-469
 *               (final_linear2 + Tanh); the 3-wide Linear is evaluated in registers (C in {8,...,256}, a power of two).
 *               bwd: dh = (dout (1 - out^2)) W, dW, db (two-stage sums; dh nullable).
 *   masked_l1   loss[0] = mean over N x channels of |where(mask > 0, out, color) - color| * 0.99^mask, mask [N] float
 *               -- trainers/inpainting3d_trainer.py:127-137 (torch.where + L1Loss(reduction='none') + pow + mean);
 *               bwd: dout = gloss[0] * [mask > 0] * 0.99^mask * sign(out - color) / (N channels). */
size_t stinet_head_workspace_bytes(int64_t n_rows, int64_t channels);
int stinet_head_fwd(const float* h, int64_t ldh, const float* W, const float* b, int64_t n_rows, int64_t channels, float* out,
                    stinet_stream_t stream);
int stinet_head_bwd(const float* h, int64_t ldh, const float* W, const float* out, const float* dout, int64_t n_rows,
                    int64_t channels, float* dh, int64_t lddh, float* dW, float* db, void* workspace, size_t workspace_bytes,
                    stinet_stream_t stream);
size_t stinet_masked_l1_workspace_bytes(int64_t n_rows);
int stinet_masked_l1_fwd(const float* out, const float* color, const float* mask, int64_t n_rows, int64_t channels,
                         float* loss, void* workspace, size_t workspace_bytes, stinet_stream_t stream);
int stinet_masked_l1_bwd(const float* out, const float* color, const float* mask, const float* gloss, int64_t n_rows,
                         int64_t channels, float* dout, stinet_stream_t stream);

/* ---- integer sort primitives and hierarchy construction (replace the Python loops and np.unique calls
 * of preprocessing/graph_level_generation.py:194-244 `vertex_clustering`).  All results are bit-identical to the reference's:
 * integers by construction, coordinates because members are summed sequentially in ascending vertex id in the input dtype.
 *   sort_pairs_u64      stable LSD radix sort (8 bits per pass, ceil(key_bits / 8) passes) of (key, value) pairs; vals_in NULL
 *                       = the positions 0..n-1 (argsort), vals_out NULL = keys only.  keys_in is not modified.
 *   csr_degree_order    order[k] = k-th row of a CSR by DESCENDING degree, ties by ascending row (north_star's degree sort)
 *   voxel_bins          bins[i,a] = floor_divide(coords[i,a], voxel) with numpy's float semantics (:207); minmax[0..2] =
 *                       per-axis minimum, [3..5] = maximum (int64)
 *   voxel_keys          keys[i] = lexicographic rank key of bin i inside the bounding box (the row order of np.unique(axis=0))
 *   unique_sorted_u64   ids[i] = index of key i among the distinct keys < limit (keys >= limit are ignored), count[0] = number
 *                       of distinct keys < limit      (:208-209 return_inverse, after the sort)
 *   cluster_finish      trace[idx_sorted[i]] = ids[i];  start[c] = first sorted position of cluster c, start[n_coarse] = n
 *   cluster_centroids   out[c] = float32(mean of coords[members of c]) (:238-242)
 *   coarse_edge_keys    key[e] = trace[src] * n_coarse + trace[dst], self loops (and out-of-range ends: status bit 0) = n_coarse^2
 *   coarse_edges_emit   the distinct keys < n_coarse^2 as the [2, n_out] coarse edge set sorted by (vertex, neighbour) (:215-228) */
size_t stinet_sort_workspace_bytes(int64_t n);
int stinet_sort_pairs_u64(const uint64_t* keys_in, const int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out, int64_t n,
                          int key_bits, void* workspace, size_t workspace_bytes, stinet_stream_t stream);
size_t stinet_csr_degree_order_workspace_bytes(int64_t n_rows);
int stinet_csr_degree_order(const int32_t* rowptr, int64_t n_rows, int32_t* order, void* workspace, size_t workspace_bytes,
                            stinet_stream_t stream);
int stinet_voxel_bins(const void* coords, int is_f64, int64_t n, double voxel, int64_t* bins, int64_t* minmax,
                      stinet_stream_t stream);
int stinet_voxel_keys(const int64_t* bins, const int64_t* minmax, int64_t n, uint64_t* keys, stinet_stream_t stream);
size_t stinet_unique_workspace_bytes(int64_t n);
int stinet_unique_sorted_u64(const uint64_t* keys_sorted, int64_t n, uint64_t limit, int32_t* ids, int32_t* count,
                             void* workspace, size_t workspace_bytes, stinet_stream_t stream);
int stinet_cluster_finish(const int32_t* idx_sorted, const int32_t* ids, int64_t n, int64_t n_coarse, int64_t* trace,
                          int32_t* start, stinet_stream_t stream);
int stinet_cluster_centroids(const void* coords, int is_f64, const int32_t* idx_sorted, const int32_t* start,
                             int64_t n_coarse, float* out, stinet_stream_t stream);
int stinet_coarse_edge_keys(const int64_t* src, const int64_t* dst, int64_t n_edges, const int64_t* trace, int64_t n_fine,
                            int64_t n_coarse, uint64_t* keys, int32_t* status, stinet_stream_t stream);
int stinet_coarse_edges_emit(const uint64_t* keys_sorted, const int32_t* ids, int64_t n_edges, int64_t n_coarse, int64_t n_out,
                             int64_t* out, stinet_stream_t stream);

/* ---- per-step graph metrics (replace utils/metrics/graph_metrics.py:6-72 as called from
 * trainers/inpainting3d_trainer.py:254-263) on the level-0 CSR by target.  Scalar results are written to `out` on the
 * device (float[1], psnr float[2] = {score, rows used}); reductions are deterministic (double partials, fixed order).
 *   laplace            out[i,c] = sum_{j->i} x[j,c] - deg_i x[i,c]                     (GraphLaplaceOperator.forward :10-13)
 *   laplace_variance   biased variance over vertices of laplace(0.299 r + 0.587 g + 0.114 b)   (GraphLaplaceVariance :24-30)
 *   total_variation    sum_edges sum_c |x[src,c] - x[dst,c]| / (n * channels)            (:33-37)
 *   psnr               -10 log10(mean(((x - y) / data_range)^2) + 1e-8) over rows with mask > 0 (mask NULL: all rows),
 *                      i.e. psnr(x[mask > 0], y[mask > 0]) without the boolean-index copies (:40-72) */
size_t stinet_metrics_workspace_bytes(int64_t n);
int stinet_graph_laplace(const float* x, int64_t ldx, const int32_t* rowptr_t, const int32_t* col_t, int64_t n,
                         int64_t channels, float* out, int64_t ldo, stinet_stream_t stream);
int stinet_graph_laplace_variance(const float* x, int64_t ldx, const int32_t* rowptr_t, const int32_t* col_t, int64_t n,
                                  float* out, void* workspace, size_t workspace_bytes, stinet_stream_t stream);
int stinet_graph_total_variation(const float* x, int64_t ldx, const int32_t* rowptr_t, const int32_t* col_t, int64_t n,
                                 int64_t channels, float* out, void* workspace, size_t workspace_bytes,
                                 stinet_stream_t stream);
int stinet_psnr(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* mask, int64_t n,
                int64_t channels, float data_range, float* out, void* workspace, size_t workspace_bytes,
                stinet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STINET_B200_H */
