/*
 * nw_sm100.h — C ABI of libnw_sm100.so: the B200 (sm_100a) implementation of the
 * Nadaraya-Watson head hot path of alanqrwang/nwhead.
 *
 * The reference has no FFI layer (it is pure PyTorch).  Each entry point below replaces the
 * torch library calls of one reference function; the citation names that function
 * (paths relative to the reference repository root).  The Python drop-in package
 * `nwhead_b200` binds these with ctypes (nwhead_b200/_abi.py); INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and nothing synchronises with the host;
 *   - the library allocates no device memory: scratch is passed in by the caller and sized with
 *     the matching *_plan / *_workspace_bytes query;
 *   - return value: NW_OK (0) or a negative NW_ERR_* code; nw_last_error() returns a
 *     thread-local, NUL-terminated description of the last failure on the calling thread;
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry point
 *     returns NW_ERR_CUDA / NW_ERR_UNSUPPORTED.
 */
#ifndef NW_SM100_H_
#define NW_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NW_ABI_VERSION 3

#if defined(__GNUC__)
#define NW_API __attribute__((visibility("default")))
#else
#define NW_API
#endif

#define NW_OK 0
#define NW_ERR_INVALID (-1)     /* bad argument (shape, alignment, enum) */
#define NW_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define NW_ERR_UNSUPPORTED (-3) /* device is not sm_100, or shape outside the kernel's limits */
#define NW_ERR_WORKSPACE (-4)   /* caller-provided scratch too small */

/* similarity kernels — nwhead/kernel.py:80-97 (get_kernel) */
#define NW_KIND_EUCLIDEAN 0   /* -cdist(x, y)                       nwhead/kernel.py:13-15 */
#define NW_KIND_HYPERSPHERE 1 /* -cdist(normalize(x), normalize(y)) nwhead/kernel.py:17-21 */
#define NW_KIND_COSINE 2      /* normalize(x) . normalize(y)        nwhead/kernel.py:23-28 */
#define NW_KIND_DOT 3         /* x . y                              nwhead/kernel.py:30-33 */
#define NW_KIND_CLIP 4        /* exp(logit_scale) * cosine          nwhead/kernel.py:35-44 */

/* score epilogue of the fused tensor-core forward */
#define NW_EPI_EUCLID 0 /* score = -sqrt(max(|q|^2 + |s|^2 - 2 q.s, 0)) */
#define NW_EPI_LINEAR 1 /* score = scale * q.s */

/* operand precision of the tensor-core path */
#define NW_PREC_BF16 1   /* one bf16 value per feature */
#define NW_PREC_BF16X3 3 /* hi/lo bf16 split, 3 products: ~fp32 accuracy at 3x the FLOPs */

/* per-pair outputs of nw_forward_emit */
#define NW_EMIT_SCORES 0    /* out[b, j] = score(b, j)  (the kernel(x, y) matrix, nwhead/kernel.py:13-44) */
#define NW_EMIT_INFLUENCE 1 /* out[b, j] = support influence (util/metric.py:47) with w = softmax_j score */
#define NW_EMIT_BLOCK_BEST 2 /* out[blk, b] = max over support rows [64 blk, 64 blk + 64) of score(b, j) */

/* row layouts written by nw_rows_to_bf16 */
#define NW_ROWS_BANK 0  /* support rows:  [hi]  or [hi | hi | lo] */
#define NW_ROWS_QUERY 1 /* query rows:    [hi]  or [hi | lo | hi] */

NW_API const char* nw_last_error(void);
NW_API int nw_abi_version(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x), else NW_ERR_*. */
NW_API int nw_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Support bank (K0) — replaces the CPU fp32 bank assembled by NWNet._compute_all_support_feats
 * (nwhead/nw.py:213-243) and stored by SupportSetEval.build_infer_iters (nwhead/support.py:113-120),
 * and the per-call `.to(device)` of the whole bank in NWNet.predict (nwhead/nw.py:156).
 * ------------------------------------------------------------------------------------------ */

/* bf16 elements per stored row for feature width d: precision*d rounded up to a multiple of 64.
 * bf16 operands (bank and prepared queries) are stored K-BLOCK-MAJOR: a matrix of n rows is the 3-D array
 * [row_elems / 64][n][64], element (row, col) at ((col / 64) * n + row) * 64 + col % 64.  Every 64-element
 * k-block of a row is one 128-byte TMA swizzle row and every (row tile, k-block) box the fused forward loads
 * is contiguous in HBM. */
NW_API int nw_row_elems(int d, int precision);

/* labels_i64[perm[i]] -> int32, validating 0 <= label < C (F.one_hot would raise, nwhead/nw.py:276)
 * and that the gathered sequence is non-decreasing.  status_out[0] = #out-of-range labels,
 * status_out[1] = #descents (0 means class-sorted).  perm may be NULL (identity). */
NW_API int nw_labels_to_i32(const int64_t* labels_i64, const int64_t* perm, int64_t n, int n_classes,
                     int32_t* labels_out, int32_t* status_out, void* stream);

/* offsets[c] = first row of class c in a class-sorted label vector, offsets[C] = n. */
NW_API int nw_class_offsets(const int32_t* labels_sorted, int64_t n, int n_classes, int32_t* offsets, void* stream);

/* mean over rows of a row-major fp32 matrix (used to centre euclidean banks before bf16 rounding;
 * distances are translation invariant).  workspace: nw_column_mean_workspace_bytes(d). */
NW_API size_t nw_column_mean_workspace_bytes(int d);
NW_API int nw_column_mean(const float* rows, int64_t n, int d, int64_t ld, float* mean_out, void* workspace,
                   size_t workspace_bytes, void* stream);

/* out[i, :] = bf16 layout of f(rows[perm[i], :]) with f = optional centring (x - center) followed by
 * optional L2 normalisation x / max(|x|, 1e-12) (F.normalize, nwhead/kernel.py:19-20,25-26,41-42);
 * sqnorm_out[i] = squared norm of the values the tensor cores will see.
 * out holds n * row_elems bf16 values, row_elems = nw_row_elems(d, precision), in the k-block-major layout
 * described above; padding columns are zero. */
NW_API int nw_rows_to_bf16(const float* rows, int64_t n, int d, int64_t ld, const int64_t* perm, const float* center,
                    int normalize, int layout, int precision, void* out_bf16, int row_elems,
                    float* sqnorm_out, void* stream);

/* Query conversion fused with replication across the GPUs that share a sharded bank (nwhead_b200/dist.py; no
 * reference counterpart — the reference has no multi-GPU code).  This rank converts ITS n query rows (NW_ROWS_QUERY
 * layout) and stores them as rows [row_offset, row_offset + n) of EVERY destination buffer: out_bf16_host[r] is rank
 * r's (row_elems / 64, n_total, 64) query buffer and sqnorm_host[r] its (n_total) norm vector, peer-mapped over
 * NVLink (host arrays of device pointers, n_dest <= 16).  Replaces an fp32 all-gather of the queries followed by
 * n_dest-fold redundant conversion.  The caller separates these stores from the consuming forward (a signal
 * barrier across the ranks on the same stream). */
NW_API int nw_rows_to_bf16_peers(const float* rows, int64_t n, int d, int64_t ld, const float* center, int normalize,
                          int precision, void* const* out_bf16_host, float* const* sqnorm_host, int n_dest,
                          int64_t n_total, int64_t row_offset, int row_elems, void* stream);

/* resid_sq_out[i] = squared norm of what nw_rows_to_bf16 (normalize = 0) discards from row i: |x - hi|^2 for
 * NW_PREC_BF16, |x - hi - lo|^2 for NW_PREC_BF16X3, x = rows[i, :] - center.  By the triangle inequality the
 * distance between two rounded rows differs from the true one by at most the sum of their residual norms: the
 * error certificate of the exact neighbour search behind NWNet.get_neighbors (nwhead/nw.py:245-249). */
NW_API int nw_rounding_residual(const float* rows, int64_t n, int d, int64_t ld, const float* center, int precision,
                         float* resid_sq_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused tensor-core forward (K1) — replaces, for a shared 2-D support with N > 25,
 * NWHead.forward (nwhead/nw.py:266-289): one_hot (276), expand (277-279), kernel (283:
 * nwhead/kernel.py:13-44), softmax (285), bmm with the one-hot labels (287), log(.+1e-12) (289).
 * The (B, N) score matrix never reaches HBM.
 * ------------------------------------------------------------------------------------------ */
typedef struct nw_forward_plan_t {
  int q_tiles;         /* query tiles: ceil(B / 128), or ceil(B / 256) when cta_pair */
  int s_tiles;         /* ceil(N / 256) */
  int chunks;          /* contiguous support chunks the bank is cut into */
  int tiles_per_chunk; /* support tiles per chunk */
  int grid;            /* persistent CTAs launched */
  int cta_pair;        /* 1: CTA pairs (cluster of 2, tcgen05 cta_group::2, 256x256 tiles); 0: single CTAs */
  int64_t side_elems;  /* floats of scratch `side` required: chunks * B * 8 + 2 * chunks (the per-tile producer gates) */
} nw_forward_plan_t;

NW_API int nw_forward_plan(int n_query, int64_t n_support, nw_forward_plan_t* plan_out);

/* Epilogue warp sets the class-LSE forward prefers for rows of row_elems bf16 (1, 2 or 4): short GEMMs are bound by
 * the epilogue's two MUFU operations per score, so several independent sets of 4 epilogue warps share a tile, each
 * with its own class-LSE table.  The caller offers room for the further tables by passing
 * side_elems >= plan.side_elems + (sets - 1) * n_query * n_classes; with less, fewer sets run (same results). */
NW_API int nw_forward_epilogue_sets(int row_elems);

/* Measurement hook (bench.py `sustained.sm_mhz_in_kernel`; no reference counterpart).  While a buffer is set, every
 * fused-forward launch whose grid fits ADDS, per CTA i, the SM cycles (clock64) and the nanoseconds (globaltimer)
 * its epilogue role was alive to buf[2 i] and buf[2 i + 1] (uint64, caller-zeroed device memory): cycles / ns is
 * the SM clock the kernel actually ran at, integrated over the launches.  buf = NULL switches it off.
 * Process-wide, not thread-safe. */
NW_API int nw_forward_set_clock_probe(void* buf_u64, int64_t capacity_ctas);

/* class_lse[b, c] = log sum_{j : labels[j] == c} exp(score(b, j)); -inf for classes with no support
 * row in this bank (or bank shard).  Inputs are the bf16 layouts of nw_rows_to_bf16.
 * labels must be class-sorted int32.  q_sqnorm / s_sqnorm are only read for NW_EPI_EUCLID.
 * side: scratch of at least plan.side_elems floats.  With room for (sets - 1) * n_query * n_classes more floats
 * (sets = nw_forward_epilogue_sets(row_elems)) the kernel runs that many independent epilogue warp sets (each with
 * its own table, the further ones placed in `side`) and combines them afterwards — short GEMMs are otherwise
 * epilogue-bound. */
NW_API int nw_forward_class_lse(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm, int n_query,
                         const void* bank_bf16, const float* s_sqnorm, const int32_t* labels, int64_t n_support,
                         int row_elems, int n_classes, float* class_lse, float* side, int64_t side_elems,
                         void* stream);

/* Bank-sharded variant with the exchange fused into the kernel (new; the reference has no multi-GPU code).
 * tables_host: HOST array of n_tables DEVICE pointers to (B, C) class-LSE tables, one per rank in rank order —
 * this GPU's own table and the peer GPUs' tables mapped into this process (NVLink P2P / symmetric memory).
 * Every class-LSE entry this shard owns is stored from the epilogue (peer stores overlap the MMAs):
 *   rows_per_table == 0 : to ALL tables            (all-gather: every rank ends with the whole table)
 *   rows_per_table  > 0 : to table[row / rows_per_table] only (all-to-all: rank r ends with complete rows
 *                         [r*rows_per_table, (r+1)*rows_per_table) and finalises / returns just those).
 * This replaces the all-reduce; the caller only needs a cross-GPU barrier before reading its table.  Tables are
 * NOT cleared by this call: fill them with -inf once; classes owned by no shard then stay -inf, and every class
 * must be owned by exactly one shard (class-aligned sharding). */
NW_API int nw_forward_class_lse_peers(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm,
                               int n_query, const void* bank_bf16, const float* s_sqnorm, const int32_t* labels,
                               int64_t n_support, int row_elems, int n_classes, float* const* tables_host,
                               int n_tables, int rows_per_table, float* side, int64_t side_elems, void* stream);

/* Dense per-pair output through the same TMA + tcgen05 mainloop (tensor-core replacement of the (B, N) matrices
 * the reference materialises): out[b, j] for every query b and bank row j, row stride ld_out floats.
 *   NW_EMIT_SCORES    : the similarity kernel itself — `kernel(x, y)` of nwhead/kernel.py:13-44 as used by
 *                       NWNet.get_neighbors (nwhead/nw.py:248) and KNN (nwhead/utils.py:187), bf16-operand accuracy.
 *   NW_EMIT_INFLUENCE : support_influence (util/metric.py:23-50) computed FROM FEATURES: w[b,j] = exp(score -
 *                       row_lse[b]) is formed in registers and never written; needs row_lse (B) = logsumexp_j
 *                       score(b, :), p_query (B) = softmax mass of the query's own class, qlabel (B) and the bank
 *                       labels.  4 bytes per pair of HBM traffic instead of 8 + the weight matrix.
 *   NW_EMIT_BLOCK_BEST: candidate search for an exact top-k (nwhead/nw.py:245-249 at scale): out is
 *                       (ceil(N / 64), ld_out >= B), out[blk, b] = best score of query b inside block blk; the
 *                       k nearest supports of a query lie in the blocks whose best score is within the operand
 *                       rounding error of the k-th best block.
 * labels may be NULL for NW_EMIT_SCORES and NW_EMIT_BLOCK_BEST. */
NW_API int nw_forward_emit(int epilogue, float scale, const void* q_bf16, const float* q_sqnorm, int n_query,
                    const void* bank_bf16, const float* s_sqnorm, const int32_t* labels, int64_t n_support,
                    int row_elems, int emit_kind, const float* row_lse, const float* p_query,
                    const int32_t* qlabel, float* out, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core backward of NWHead.forward against a large shared support (the autograd of
 * nwhead/nw.py:266-289 as triggered at train.py:414) in three tensor-core passes:
 *   1. nw_backward_coefficients, orientation 0 and / or 1: recompute the scores (same GEMM as the forward) and turn
 *      each into w(b, j) = dL/dscore(b, j) [/ distance], rounded to bf16, stored as the A operand of step 2;
 *   2. nw_dense_products: grad_q = W S - rowsum(W) q  and  grad_s = W^t Q - colsum(W) s   (euclidean), or
 *      grad_q = W S, grad_s = W^t Q (linear scores), with split-K for skinny problems.
 * With P(b, c) = exp(class_lse[b, c] - row_lse[b]) and g = dL/dlogp, the caller supplies the table
 *   T(b, c) = g(b, c) / (P(b, c) + 1e-12) - sum_c' g(b, c') P(b, c') / (P(b, c') + 1e-12)
 * (orientation 0: T as (B, table_ld); orientation 1: its transpose (C, table_ld >= B)).
 *
 * rows / cols: the two operands in the fused forward's k-block-major bf16 layout with their squared norms
 * (euclidean).  Orientation 0: rows = queries (row_lse (B) required), cols = class-sorted bank (col_labels int32
 * required).  Orientation 1: rows = bank (row_labels required), cols = queries (col_lse required).
 * out_bf16: (ceil(n_cols / 64), n_rows, 64) bf16, every element written (padding columns as zeros):
 * out[j / 64][r][j % 64] = w(r, j).
 * row_sums (n_rows): sum_j of the ROUNDED w(r, j) — rowsum(W) / colsum(W) of the formulas above, summed in a
 * fixed order (bitwise reproducible).  workspace: nw_backward_coefficients_workspace_elems(n_rows, n_cols) floats. */
NW_API int64_t nw_backward_coefficients_workspace_elems(int64_t n_rows, int64_t n_cols);
NW_API int nw_backward_coefficients(int epilogue, float scale, int orientation, const void* rows_bf16,
                             const float* rows_sqnorm, int64_t n_rows, const void* cols_bf16,
                             const float* cols_sqnorm, int64_t n_cols, int row_elems, const float* row_lse,
                             const int32_t* row_labels, const float* col_lse, const int32_t* col_labels,
                             const float* table, int64_t table_ld, void* out_bf16, float* row_sums,
                             float* workspace, int64_t workspace_elems, void* stream);

/* grad_s of the tensor-core backward in one launch (products + last step):
 *   out[dst(c)][r] = sum_k a[r][k] * b[c][k] - col_sub[c] * rows_t(r, c),   r < n_out_cols, c < n_b
 * a = Q^t (k_elems/64, n_a, 64): the features as rows, K = the queries; b = W^t (k_elems/64, n_b, 64): the supports as
 * rows; rows_t = the transposed bank (ceil(n_b / 64), n_a, 64) with col_sub = colsum(W) (both NULL for linear
 * scores); dst_rows int32 (n_b): bank row -> row of the caller's support tensor (NULL: identity).  out fp32
 * (n_b, ld_out >= n_out_cols).  The big operand streams once as the kernel's "bank"; nothing but the gradient is
 * written. */
NW_API int nw_dense_products_transposed(const void* a_bf16, int64_t n_a, const void* b_bf16, int64_t n_b, int k_elems,
                                 const float* col_sub, const void* rows_t_bf16, const int32_t* dst_rows,
                                 int n_out_cols, float* out, int64_t ld_out, void* stream);

/* Transposed operand: in (kblocks, n_rows, 64) bf16 k-block-major -> out (ceil(n_rows / 64), kblocks * 64, 64),
 * out[r / 64][f][r % 64] = in[f / 64][r][f % 64], zero rows appended up to a multiple of 64 (the B operand of
 * nw_dense_products when the contraction runs over the ROWS: S^t for grad_q, Q^t for grad_s). */
NW_API int nw_transpose_kblocks(const void* in_bf16, int64_t n_rows, int kblocks, void* out_bf16, void* stream);

/* Last step of a gradient: out[dst(r)][c] = raw[r][c] - row_sums[r] * rows_bf16[c / 64][r][c % 64], c < d, with
 * dst(r) = perm ? perm[r] : r (bank row -> row of the caller's support tensor).  row_sums == NULL: copy / permute
 * only (linear scores).  The subtracted rows are the STORED (centred, rounded) operand rows, so the two terms of a
 * close (query, support) pair cancel. */
NW_API int nw_backward_finish(const float* raw, int64_t ld_raw, const void* rows_bf16, const float* row_sums,
                       const int64_t* perm, int64_t n_rows, int d, float* out, int64_t ld_out, void* stream);

/* out[ks][r][c] = sum over the k-blocks of K slice ks of a[r][k] * b[c][k]; a (k_elems/64, n_a, 64) and
 * b (k_elems/64, n_b, 64) bf16 k-block-major, out fp32 row-major (n_a, ld_out >= n_b) per slice, slices
 * slice_stride floats apart.  The number of slices actually written is ceil(kblocks / ceil(kblocks / kslices))
 * (no slice is empty); the caller sums them. */
NW_API int nw_dense_products(const void* a_bf16, int64_t n_a, const void* b_bf16, int64_t n_b, int k_elems,
                      int kslices, float* out, int64_t ld_out, int64_t slice_stride, void* stream);

/* logp[b, c] = log( exp(class_lse[b,c] - logsumexp_c class_lse[b,:]) + 1e-12 )  (nwhead/nw.py:285-289).
 * With a sharded bank, all-reduce class_lse with MAX across ranks first (each class is owned by one
 * rank, so the merge is exact). */
NW_API int nw_logp_from_class_lse(const float* class_lse, int n_query, int n_classes, float* logp, void* stream);

/* row_lse[b] = logsumexp_c class_lse[b,:] (the softmax normaliser over ALL supports) and, when p_query is not
 * NULL, p_query[b] = exp(class_lse[b, qlabel[b]] - row_lse[b]) = softmaxes[b, y_b] of util/metric.py:45 — the
 * per-query inputs of nw_forward_emit(NW_EMIT_INFLUENCE).  qlabel must be in [0, C). */
NW_API int nw_row_stats(const float* class_lse, int n_query, int n_classes, const int32_t* qlabel, float* row_lse,
                 float* p_query, void* stream);

/* Exact merge of two class-LSE tables (generic row-sharded banks): a = log(exp(a) + exp(b)). */
NW_API int nw_class_lse_merge(float* a, const float* b, int64_t n_elems, void* stream);

/* ------------------------------------------------------------------------------------------
 * Direct fp32 path — exact-difference scores, any shape, 2-D or per-query 3-D support, with
 * gradients.  Replaces NWHead.forward + its autograd for episodic training (nwhead/nw.py:162-211,
 * train.py:412-415) and the direct `kernel(x, y)` call of NWNet.get_neighbors (nwhead/nw.py:248).
 * torch.cdist itself switches to exact differences for <= 25 rows.
 * ------------------------------------------------------------------------------------------ */

/* scores[b, j] = kernel(q[b], s[j]) (support_batched = 0, s is (N, d)) or kernel(q[b], s[b, j])
 * (support_batched = 1, s is (B, N, d)).  scale: exp(logit_scale) for NW_KIND_CLIP, ignored otherwise. */
NW_API int nw_direct_scores(int kind, float scale, const float* q, int n_query, int d, const float* s,
                     int64_t n_support, int support_batched, float* scores, void* stream);

/* softmax over the support axis + label aggregation + log: logp (B, C), row_lse (B) = logsumexp_j
 * scores[b, :] (saved for backward).  labels int64 as handed to F.one_hot: (N) or (B, N).
 * status_flag: caller-zeroed int32, set to 1 (plain store, never cleared) when a label is outside [0, C) —
 * what F.one_hot rejects at nwhead/nw.py:276.  It may be device memory or device-visible (mapped / UVA pinned)
 * HOST memory, so the caller can poll it without a copy.  Such a label contributes to no class, in the forward
 * and in nw_direct_backward alike (no out-of-bounds access is possible). */
NW_API int nw_direct_aggregate(const float* scores, const int64_t* labels, int labels_batched, int n_query,
                        int64_t n_support, int n_classes, float* logp, float* row_lse, int32_t* status_flag,
                        void* stream);

/* scores + aggregate in one call; a single fused launch when n_support <= 1024 (episodic training). */
NW_API int nw_direct_forward(int kind, float scale, const float* q, int n_query, int d, const float* s,
                      int64_t n_support, int support_batched, const int64_t* labels, int labels_batched,
                      int n_classes, float* scores, float* logp, float* row_lse, int32_t* status_flag,
                      void* stream);

/* d + n_classes limit of nw_direct_backward (one feature row + one class table staged in shared memory) */
#define NW_DIRECT_BACKWARD_MAX_D_PLUS_C 49152

/* closed-form backward.  workspace: nw_direct_backward_workspace_elems(...) floats.
 * grad_q (B, d) and grad_s ((N, d) or (B, N, d)) may each be NULL when not needed.
 * grad_scale_rows (B) receives per-query partial sums of d/d(logit_scale) for NW_KIND_CLIP (may be
 * NULL).  Coincident points contribute zero gradient, as torch.cdist's backward does. */
NW_API int64_t nw_direct_backward_workspace_elems(int n_query, int d, int64_t n_support, int support_batched);
NW_API int nw_direct_backward(int kind, float scale, const float* q, int n_query, int d, const float* s,
                       int64_t n_support, int support_batched, const int64_t* labels, int labels_batched,
                       int n_classes, const float* scores, const float* row_lse, const float* logp,
                       const float* grad_out, float* workspace, float* grad_q, float* grad_s,
                       float* grad_scale_rows, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cluster mode (K3) — replaces compute_clusters(embeddings, labels, n_clusters=1)
 * (nwhead/utils.py:218-246; sklearn KMeans with one cluster is the class mean).
 * out[c, :] = mean of the rows of class c (zeros for an empty class); rows are addressed through
 * perm (NULL = identity) so that offsets describe a class-sorted order.
 * workspace: nw_class_centroids_workspace_bytes(n_classes, d).
 * ------------------------------------------------------------------------------------------ */
NW_API size_t nw_class_centroids_workspace_bytes(int n_classes, int d);
NW_API int nw_class_centroids(const float* rows, int d, int64_t ld, const int64_t* perm, const int32_t* offsets,
                       int n_classes, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* k-means assignment step (K3b) — the inner loop of compute_clusters(embeddings, labels, n_clusters=k > 1)
 * (nwhead/utils.py:230: one scikit-learn KMeans fit per class).  group[i] is the class of row i (0 <= group <
 * n_groups); centroids is (n_groups * k, d) with the k centroids of a class stored consecutively.  Row i is
 * compared with the k centroids of its own class only (squared euclidean distance, exact fp32 differences):
 * assign_out[i] = group[i] * k + argmin_j (lowest j on ties), dist_out[i] = that squared distance (may be
 * NULL).  order (may be NULL = identity) lists the rows in class-sorted order: rows are visited in that order,
 * two per warp, so that neighbours share their class's centroid loads; results do not depend on it.
 * The update step is nw_class_centroids over the rows ordered by assign_out. */
NW_API int nw_kmeans_assign(const float* rows, int d, int64_t ld, const int32_t* group, const int64_t* order,
                     int64_t n_rows, const float* centroids, int k, int32_t* assign_out, float* dist_out,
                     void* stream);

/* k-means++ seeding with scikit-learn's arithmetic, for all classes at once (compute_clusters with n_clusters > 1,
 * nwhead/utils.py:230 -> sklearn KMeans(random_state=0): _kmeans_plusplus).
 * nw_kmeans_seed_dist: dist_out[t, i] = squared distance of row i to candidate row cand_rows[group[i], t] of its own
 *   class, evaluated like sklearn: both rows centred with the class mean in float32 (X -= X.mean(0)), the squared
 *   distance accumulated in float64, rounded to float32 and clipped at 0 (_euclidean_distances_upcast).
 *   class_mean (n_groups, d); cand_rows (n_groups, n_cand) int64 row indices; dist_out (n_cand, n_rows); order as in
 *   nw_kmeans_assign; n_cand <= 8.
 * nw_kmeans_seed_pick: pos_out[c, t] = row at the first class-sorted position of class c whose running float32 sum
 *   of closest[] (sequential, numpy's cumsum) reaches targets[c, t] (float64) — np.searchsorted(np.cumsum(...), v),
 *   clipped to the class's last row.  offsets (n_classes + 1) class-sorted positions; order maps positions to rows
 *   (NULL = identity); n_targets <= 8. */
NW_API int nw_kmeans_seed_dist(const float* rows, int d, int64_t ld, const int32_t* group, const int64_t* order,
                        int64_t n_rows, const float* class_mean, const int64_t* cand_rows, int n_cand,
                        float* dist_out, void* stream);
NW_API int nw_kmeans_seed_pick(const float* closest, const int64_t* order, const int32_t* offsets, int n_classes,
                        const double* targets, int n_targets, int64_t* pos_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * support_influence (K4) — replaces the per-query Python loop of util/metric.py:23-50.
 * ------------------------------------------------------------------------------------------ */

/* out[r] = argmax_c onehot[r, c] (first maximum, as torch.argmax) — util/metric.py:42-43. */
NW_API int nw_onehot_argmax(const float* onehot, int64_t n_rows, int n_classes, int32_t* out, void* stream);

/* out[b, g, j] = log( (p_b - p_b * w[b, j]) / (p_b - w[b, j] * [slabel[g, j] == qlabel[b]]) ),
 * p_b = softmaxes[b, qlabel[b]]   (util/metric.py:45-47).  n_label_sets = 1 for (N, C) slabels,
 * = B for the documented (B, N, C) slabels, whose argmax broadcasts to a (B, B, N) result. */
NW_API int nw_support_influence(const float* softmaxes, const int32_t* qlabel, const float* sweights,
                         const int32_t* slabel, int n_query, int n_label_sets, int64_t n_support,
                         int n_classes, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Neighbour ranking (K5) — replaces torch.argsort(scores, descending=True) of NWNet.get_neighbors
 * (nwhead/nw.py:248-249) and KNN.__call__ (nwhead/utils.py:187-189).
 * idx_out (R, k) int64 column indices of the k largest scores of each row, best first; equal scores
 * are ordered by ascending index.  k == n_cols gives the full ranking.
 * workspace: nw_rank_rows_workspace_bytes(n_rows, n_cols).
 * ------------------------------------------------------------------------------------------ */
NW_API size_t nw_rank_rows_workspace_bytes(int n_rows, int64_t n_cols);
NW_API int nw_rank_rows(const float* scores, int n_rows, int64_t n_cols, int64_t k, int64_t* idx_out, void* workspace,
                 size_t workspace_bytes, void* stream);

/* Exact top-k refinement — the device side of SupportBank.topk_exact (exact k nearest supports without the (B, N)
 * score matrix; same ranking as NWNet.get_neighbors' dense fp32 path, nwhead/nw.py:245-249, ties by ascending
 * source index).  For every query b that is not yet done: a candidate budget m <= m_cap is sized from the error
 * bounds of the reduced-precision pass, the rows of its m best 64-row bank blocks (block_order[b, 0..m), ranked by the
 * NW_EMIT_BLOCK_BEST pass) are gathered from the fp32 source rows (through perm: bank row -> source row, NULL =
 * identity), scored exactly (the per-pair arithmetic of nw_direct_scores), ranked, and the query is certified: with
 * beta = block_best_sorted[b, m] (pass score of the best block left out), no row outside the candidates can reach
 * the exact k-th candidate score (bounds from q_sqnorm, the measured rounding residuals resid_q / *resid_max and
 * *smax_sq = max squared norm of the bank rows).  Certified queries get idx_out[b, 0..k) (source indices, best
 * first) and done[b] = 1; the others add 1 to *n_pending.  One launch, no host synchronisation.
 * m_cap <= 64, k <= 64 m_cap; order_stride = entries per query in block_order / block_best_sorted (> m_cap unless
 * m_cap == n_blocks; >= k lets the kernel size m below m_cap). */
NW_API int nw_topk_refine(const float* q, int n_query, int d, const float* source_rows, int64_t n_rows,
                   const int64_t* perm, const int64_t* block_order, const float* block_best_sorted, int order_stride,
                   int m_cap, int64_t n_blocks, int k, const float* q_sqnorm, const float* resid_q, const float* smax_sq,
                   const float* resid_max, int precision, int32_t* done, int64_t* idx_out, int32_t* n_pending,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NW_SM100_H_ */
