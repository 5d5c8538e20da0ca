/* scb200.h -- C ABI of libscb200.so, the B200 (sm_100a) implementation of the
 * sparsify-clip contrastive-loss hot path.
 *
 * The reference (noostale/sparsify-clip) has no FFI / plugin interface: its hot path is
 * six Python functions executed by PyTorch eager ops.  The drop-in
 * boundary is therefore the Python signatures in sparsify_clip_b200/losses.py; those
 * call the entry points below through ctypes.  Each entry point cites the reference
 * lines whose arithmetic it replaces (paths relative to the reference repo).
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless noted;
 *  - the library never allocates or frees device memory and keeps no global state
 *    beyond the lazily resolved cuTensorMapEncodeTiled driver entry point;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs
 *    no host synchronisation and is CUDA-graph capturable;
 *  - return value: 0 ok; <0 invalid argument (SCB_E_*); >0 a cudaError_t.
 *    scb_last_error() returns a thread-local message for the last failure;
 *  - matrices are row-major [n, D] with leading dimension ld (elements);
 *  - dtype: SCB_F32 / SCB_BF16 / SCB_F16;
 *  - path:  SCB_PATH_SIMT  fp32 CUDA-core kernels (any dtype, any D; the exact path used
 *                          for fp32 inputs and the "tf32-off" parity gate),
 *           SCB_PATH_TC    TMA + tcgen05/TMEM tensor-core kernels (bf16/fp16, D % 8 == 0,
 *                          16-byte aligned rows);
 *  - "jparts": the column sweep of a pass may be split into `jparts` contiguous parts so
 *    that the grid fills 148 SMs; each part writes its own partial result and the
 *    *_finalize / *_combine entry points add the partials in a fixed order
 *    (deterministic, no float atomics).  A pass writes nsub sub-partials per part
 *    for the per-row statistics (scb_pass_plan reports jparts and nsub).
 */
#ifndef SCB200_H_
#define SCB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_F32 0
#define SCB_BF16 1
#define SCB_F16 2

#define SCB_PATH_SIMT 0
#define SCB_PATH_TC 1

#define SCB_E_ARG (-1)       /* bad pointer / size / alignment */
#define SCB_E_DTYPE (-2)     /* dtype not supported on this path */
#define SCB_E_SHAPE (-3)     /* shape not supported on this path */
#define SCB_E_DRIVER (-4)    /* cuTensorMapEncodeTiled unavailable / failed */

/* ABI revision of this header.  scb_version() returns the revision the loaded library was BUILT against; a binding
 * must refuse a library whose revision differs (argument lists may have changed). */
#define SCB_ABI_VERSION 203
int scb_version(void);
/* Device-resident temperature.  Every entry that takes the logit scale 1/tau as a host float (`scale`, and the
 * coefficients derived from it) also takes `const float* scale_dev` as its last argument before the stream: when non-NULL
 * the effective scale is scale * (*scale_dev), read on the device -- pass scale = 1 and a pointer to 1/tau.  Nothing then
 * synchronises with the host, and a captured CUDA graph follows a temperature that changes between replays. */
const char* scb_last_error(void);
/* Launch plan of a B x B pass with nA rows against nB columns (host-only arithmetic, no CUDA call):
 * *jparts = how many contiguous parts the column sweep is split into so that the work items fill
 * n_sm SMs evenly; *nsub = per-row statistic sub-partials each part writes (1 SIMT, 2 TC, 4 TC on a
 * CTA pair or a cluster of 4).  grad != 0 for the passes that also produce a [nA, D] output. */
int scb_pass_plan(int path, int64_t nA, int64_t nB, int D, int grad, int n_sm, int* jparts, int* nsub);
/* debug/tuning knobs for the TC path: bit0 = keep the weight tile in TMEM (TS-mode MMA) instead of
 * shared memory; bit1 = run the gradient passes on CTA pairs (cluster of 2, weight tile shared through
 * distributed shared memory) when 256 < D <= 512; bit2 = run them on clusters of 4 with cta_group::2
 * MMAs (two row blocks x two output halves) when 256 < D <= 1024 and nA > 128; bit3 = with bit2, give the last rows of a
 * large pass to CTA pairs running concurrently on the SMs that clusters of 4 cannot use; bit4 = with bit2, passes with
 * 512 < D <= 768 keep all output columns in TMEM (one S buffer) instead of running once per group of 512 output columns
 * (768 < D <= 1024 always runs in two groups); bit5 = use the row-block-aligned work split of the cluster-of-4 kernel
 * whatever the operand size (by default only when the column operand exceeds 64 MB, i.e. does not stay in L2).
 * Default: bits 0-4 set.  Returns the previous value. */
int scb_set_tc_flags(int flags);
/* Work split of the cluster-of-4 gradient kernel (host-only arithmetic, no CUDA call): the (256-row block, 128-column
 * tile) space of n_rp x n_jb items, linearised row-block-major, is cut into *n_used contiguous spans of *span items, one
 * per cluster; *pmax = the largest number of spans that touch one row block (= output partial slots).  align != 0: spans
 * of whole row blocks when that costs at most 6 % of balance (2: whatever it costs) -- the split used when the column
 * operand does not stay in L2. */
int scb_quad_plan(int64_t n_rp, int64_t n_jb, int n_clusters, int align, int* n_used, int64_t* span, int* pmax);
/* Which kernel a TC gradient pass over nA rows of width D runs on the current device:
 * 0 = single CTA (k_tc_pass), 1 = CTA pair (k_tc_pair), 2 = cluster of 4 (k_tc_quad).  In *units (may be
 * null): how many of those run concurrently (SMs, pairs, clusters).  The first call on a device asks the
 * driver how many clusters of 4 fit (cudaOccupancyMaxActiveClusters); scb_pass_plan uses the same answer. */
int scb_grad_kernel_kind(int64_t nA, int D, int n_sm, int* units);

/* ------------------------------------------------------------------ row-wise kernels */

/* out[i] = sum_d X[i,d]^2 (fp32).  Used for d2_ij = n_i + n_j - 2 x_i.x_j, the Gram form
 * of torch.pdist at sparsify_clip.py:161. */
int scb_row_sqnorm(const void* X, int64_t n, int D, int64_t ld, int dtype, float* out, void* stream);

/* out[i] = A[i,:] . B[i,:] (fp32): the diagonal logits of sparsify_clip.py:119 before /tau. */
int scb_row_dot(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype,
                float* out, void* stream);

/* L_align, sparsify_clip.py:186-187 with alpha = 2:
 * row_out[i] = ||x_i - y_i||^2.  The mean is taken by scb_sum + a host scalar. */
int scb_lalign_rows(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                    float* row_out, void* stream);
/* dX[i,:] (+)= scale * (x_i - y_i), dY[i,:] (+)= -scale * (x_i - y_i); scale = host_scale *
 * (dev_scale ? *dev_scale : 1).  Gradient of the line above (2/B folded into scale). */
int scb_lalign_bwd(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                   float host_scale, const float* dev_scale, int accumulate, float* dX, float* dY, void* stream);

/* c = normalize((a+b)/2, eps=1e-12): compute_centroids_only (sparsify_clip.py:334-355)
 * followed by F.normalize at its call sites (:803-805 ...).  C_out has dtype `out_dtype`;
 * inv_norm[i] = 1/max(||m_i||,1e-12). */
int scb_centroid_fwd(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype,
                     void* C_out, int out_dtype, float* inv_norm, void* stream);
/* given dC (fp32 [n,D]): dm = (dC - c (c.dC)) * inv_norm; dA (+)= scale*dm/2; dB likewise. */
int scb_centroid_bwd(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype,
                     const float* dC, const float* inv_norm, float host_scale, const float* dev_scale,
                     int accumulate, float* dA, float* dB, void* stream);

/* pre-loss normalise e/||e|| (no eps), sparsify_clip.py:772-773, and its backward
 * dx = (g - xhat (xhat.g)) / ||x||. */
int scb_normalize_fwd(const void* X, int64_t n, int D, int64_t ld, int dtype, void* Y, int out_dtype,
                      float* inv_norm, void* stream);
int scb_normalize_bwd(const void* X, int64_t n, int D, int64_t ld, int dtype, const float* dY,
                      const float* inv_norm, float* dX, void* stream);

/* deterministic two-stage sum of n floats into out[0] (scratch >= 1024 floats). */
int scb_sum(const float* x, int64_t n, float* scratch, float* out, void* stream);

/* ------------------------------------------------------------- B x B passes (never materialised) */

/* Row log-sum-exp partials of S = scale * A.Bm^T  (sparsify_clip.py:119-120 + the
 * log_softmax inside F.cross_entropy at :127/:129; call with (I,T) for rows and (T,I)
 * for columns).  Writes part_m/part_l [jparts*nsub][nA] (log2-domain running max and sum). */
int scb_lse_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB,
                 int dtype, float scale, int jparts, float* part_m, float* part_l, int path, const float* scale_dev, void* stream);
/* lse[i] = natural-log LSE from nparts partials. */
int scb_lse_combine(const float* part_m, const float* part_l, int nparts, int64_t n, float* lse, void* stream);

/* Fused row + column LSE (tensor-core path): ONE sweep over S = scale * A.Bm^T yields the row partials of
 * scb_lse_pass and, from the same tiles, partial column sums per 32-row strip:
 *   sum_{i in strip p} 2^{y_ij} = col_sum[p][j] * 2^{col_ref[p][j/32]},  y = scale*log2(e) * A_i.Bm_j
 * col_sum is [4*ceil(nA/128)][nB], col_ref [4*ceil(nA/128)][ceil(nB/32)]; part_m/part_l are [jparts*4][nA] here.  This replaces the second sweep
 * (rows of S^T) of F.cross_entropy(logits.t(), ...) at sparsify_clip.py:129.  The partial sums are exact while the
 * logits spread by less than 2^100 inside a 32 x 32 block; scb_lse2_spread_flag evaluates a sufficient norm bound ON
 * THE DEVICE (flag = 1: not guaranteed) and scb_lse_pass_cond / scb_lse_combine_cond run the exact second sweep only
 * when that flag is set -- no host synchronisation, CUDA-graph capturable. */
int scb_lse2_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                  float scale, int jparts, float* part_m, float* part_l, float* col_ref, float* col_sum, const float* scale_dev, void* stream);
/* lse[j] = natural-log column LSE from the nparts = 4*ceil(nA/128) strip partials. */
int scb_colstat_combine(const float* col_ref, const float* col_sum, int nparts, int64_t n, float* lse, void* stream);
/* the same fold kept as a pair: sum over the partials = sum_out[j] * 2^ref_out[j] (log2 domain).  Row-sharded runs
 * all-gather these pairs (each rank sweeps its rows against all columns) and fold them once more. */
int scb_colstat_partial(const float* col_ref, const float* col_sum, int nparts, int64_t n, float* ref_out, float* sum_out,
                        void* stream);
/* That second fold, after the packed gather of a sharded step (losses.py, exchange step 2).  pack = [world][stride]
 * fp32, one row per rank: at off_exact the exact column LSE of the rank's own n_loc columns, at off_ref / off_sum its
 * (reference, sum) pair for every one of the world * n_loc columns.  col_lse[j] = the exact value when *flag != 0
 * (the norm bound armed the second sweep on every rank), else ln2 * (M + log2 sum_r sum_r[j] 2^(ref_r[j] - M)). */
int scb_lse2_fold_ranks(const float* pack, int world, int64_t stride, int64_t n_loc, int64_t off_exact, int64_t off_ref,
                        int64_t off_sum, const int* flag, float* col_lse, void* stream);
/* *flag = (2 * scale * log2(e) * max_i |A_i| * max_j |B_j| >= 90) from the squared row norms (scb_row_sqnorm). */
int scb_lse2_spread_flag(const float* sqnA, int64_t nA, const float* sqnB, int64_t nB, float scale, int* flag, const float* scale_dev, void* stream);
/* scb_lse_pass / scb_lse_combine that do nothing unless *run_flag != 0 (device pointer). */
int scb_lse_pass_cond(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                      float scale, int jparts, float* part_m, float* part_l, const int* run_flag, const float* scale_dev, void* stream);
int scb_lse_combine_cond(const float* part_m, const float* part_l, int nparts, int64_t n, float* lse, const int* run_flag,
                         void* stream);

/* Backward of the anchor loss w.r.t. the rows of A (sparsify_clip.py:110-132; the
 * recompute replaces autograd's saved B x B tensors):
 *   out[p][i,:]  = sum_{j in part p, j != i+diag_off} (e^{s_ij - row_lse_i} + e^{s_ij - col_lse_j}) Bm[j,:]
 *   ws[p*nsub+q][i] = sum_j (same weights, diagonal INCLUDED) * (A_i.Bm_j)        (for d/dtau)
 * s_ij = scale*A_i.Bm_j.  Call with (I,T,row_lse=r,col_lse=c) for dI and (T,I,c,r) for dT.
 * ws may be NULL. */
int scb_anchor_grad_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB,
                         int dtype, float scale, const float* row_lse, const float* col_lse, int64_t diag_off,
                         int jparts, float* out, float* ws, int path, const float* scale_dev, void* stream);
/* dA[i,:] (+)= s * ( sum_p out[p][i,:] + (e^{sc*diag_i - row_lse_i} + e^{sc*diag_i - col_lse_i} - 2) * V[i,:] ),
 * s = host_scale * (dev_scale ? *dev_scale : 1); V = the paired rows of the other modality. */
int scb_anchor_grad_finalize(const float* out, int jparts, int64_t n, int D, const void* V, int64_t ldV, int dtype,
                             const float* row_lse, const float* col_lse_rows, const float* diag, float scale,
                             float host_scale, const float* dev_scale, int accumulate, float* dA, const float* scale_dev, void* stream);

/* Single-pass L_unif forward+backward core (sparsify_clip.py:159-164; replaces torch.pdist,
 * the 7 element-wise passes over its output and _pdist_backward):
 *   w_ij = exp(-t (n_i + n_j - 2 x_i.x_j)),  w_ij = 0 where global row == global column
 *   U[p][i,:] = sum_{j in part p} w~_ij Xall[j,:]      (w~ = w rounded to the MMA operand type on TC)
 *   rq[.][i]  = sum_j w~_ij         rs[.][i] = sum_j w_ij (fp32)
 * Xr = the nR local rows (global row index = row_offset + i), Xall = all nAll rows. */
int scb_lunif_pass(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                   int dtype, float t, const float* sqn_r, const float* sqn_all, int64_t row_offset,
                   int jparts, float* U, float* rq, float* rs, int path, void* stream);
/* forward-only variant (no gradient needed): rs only. */
int scb_lunif_sum_pass(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                       int dtype, float t, const float* sqn_r, const float* sqn_all, int64_t row_offset,
                       int jparts, float* rs, int path, void* stream);
/* dX[i,:] (+)= s * ( (sum_q rq[q][i]) * x_i - sum_p U[p][i,:] ),  s = host_scale * *dev_scale
 * (dev_scale carries -2t/Ssum, which depends on a device-side reduction). */
int scb_lunif_grad_finalize(const float* U, int jparts, const float* rq, int nparts_rq, int64_t n, int D,
                            const void* X, int64_t ld, int dtype, float host_scale, const float* dev_scale,
                            int accumulate, float* dX, void* stream);

/* Fused gradient finaliser of the composed loss (the ladder at sparsify_clip.py:775-938 adds the terms; here
 * their gradients w.r.t. one operand are combined in a single pass, written in `out_dtype`):
 *   dX[i,:] = gs * ( a_coef * ( sum_p a_out[p][i,:] + (e^{sc*diag_i - row_lse_i} + e^{sc*diag_i - col_lse_i} - 2) * Y[i,:] )
 *                  + uc     * ( (sum_q rq[q][i]) * X[i,:] - sum_p u_out[p][i,:] )
 *                  + l_coef * ( X[i,:] - Y[i,:] )
 *                  + e_coef * extra[i,:] )
 * gs = dev_scale ? *dev_scale : 1;  uc = u_coef * (u_dev_coef ? *u_dev_coef : 1).  a_out / u_out / extra may be NULL
 * (term absent); l_coef == 0 skips L_align.  X = the operand's rows, Y = the paired rows of the other modality;
 * extra = a finished contiguous fp32 [n, D] gradient term (the centroid chain of sparsify_clip.py:353/804: L_unif of
 * the normalised centroids pulled back through the normalisation by scb_centroid_bwd).
 * unit_src / unit_inv (both or neither): the pre-loss normalise of sparsify_clip.py:772-773 fused in.  X then holds the
 * NORMALISED rows E_i / ||E_i|| the loss was evaluated on, unit_src the un-normalised rows E (dtype unit_dtype, row stride
 * ld_unit), unit_inv[i] = 1 / ||E_i|| (scb_normalize_fwd), and the pass writes the gradient w.r.t. E:
 *   dE[i,:] = ( dX[i,:] - e_i (e_i . dX[i,:]) ) * unit_inv[i],   e_i = E_i * unit_inv[i] in fp32
 * i.e. scb_normalize_bwd applied to the fp32 dX before the output conversion.  Needs D <= 2048 (16-byte aligned rows, D % 8
 * == 0) or D <= 256 otherwise: a row is reduced inside one thread block. */
int scb_grad_combine(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                     const float* a_out, int a_jparts, const float* row_lse, const float* col_lse_rows,
                     const float* diag, float scale, float a_coef, const float* u_out, int u_jparts,
                     const float* rq, int rq_parts, float u_coef, const float* u_dev_coef, float l_coef,
                     const float* extra, float e_coef, const float* dev_scale, void* dX, int out_dtype,
                     int64_t ldOut, const float* scale_dev, const void* unit_src, int64_t ld_unit, int unit_dtype,
                     const float* unit_inv, void* stream);

/* Scalar assembly of the composed loss (the additions of the ladder, sparsify_clip.py:778-938, on the partial sums):
 * parts = [sum_i row_lse, sum_j col_lse, sum_i I_i.T_i, sum_i |I_i - T_i|^2, rs(img), rs(txt), rs(cen)] (device),
 *   *loss = c_anchor (parts[0] + parts[1] - two_scale parts[2]) + c_align parts[3]
 *           + sum_k w_k log( (parts[4+k] / 2) / pair_norm ),      inv_ssum[k] = 2 / parts[4+k]   (0 when w_k == 0)
 * c_anchor = w_anchor / 2B, two_scale = 2 / tau, c_align = w_align / B, pair_norm = B (B - 1) / 2; a zero weight
 * skips its term.  One launch instead of ~20 one-element element-wise launches of the host framework. */
int scb_loss_assemble(const float* parts, float c_anchor, float two_scale, float c_align, float w_unif_img,
                      float w_unif_txt, float w_unif_cen, float pair_norm, float* loss, float* inv_ssum, const float* scale_dev, void* stream);

/* sparsify_loss (sparsify_clip.py:166-176), forward: row partial sums of
 * (x_i.x_j - (2 delta_ij - 1))^2 over j; rs [jparts*nsub][nR]. */
int scb_sparsify_sum_pass(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                          int dtype, int64_t row_offset, int jparts, float* rs, int path, void* stream);

/* ------------------------------------------------------------------ cold variants and evaluation-side consumers */

/* out[d] = scale * sum_i (X[i,d] - Y[i,d])   (Y may be NULL).  centroid_alignment_loss (sparsify_clip.py:487-505) and
 * compute_gap (:418-436) are |mean I - mean T|_p; the mean off-diagonal cosine (:438-457) is
 * (|sum_i x_i|^2 - sum_i |x_i|^2) / (N (N - 1)).  scratch: [scratch_rows][D] fp32, scratch_rows >= 1 (more rows = more
 * parallelism; fixed-order two-stage sum). */
int scb_col_sum(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype, float scale,
                float* scratch, int scratch_rows, float* out, void* stream);
/* out[D][D] = scale * sum_i (x_i - mu)(x_i - mu)^T   (mu may be NULL), fp32 FMA.  The covariance of the W2 uniformity
 * metric (uniformity.py:26, :70, :106, :150, :188; sparsify_clip.py:466) and X^T X of sparsify_loss's backward.
 * scratch: [scratch_parts][D][D] fp32. */
int scb_gram_dd(const void* X, int64_t n, int D, int64_t ld, int dtype, const float* mu, float scale, float* scratch,
                int scratch_parts, float* out, void* stream);
/* out[n][D] (fp32, contiguous) = X[n][D] . M[D][D] */
int scb_rows_times_dd(const void* X, int64_t n, int D, int64_t ld, int dtype, const float* M, float* out, void* stream);
/* out[i] = scale * sum_p parts[p][i], fixed order */
int scb_sum_parts(const float* parts, int nparts, int64_t n, float scale, float* out, void* stream);
/* rank[q] = #{e : S[l, e] > S[l, gt[q]]}, l = line ? line[q] : q, for n_lines queries over the lines of a score matrix
 * given with explicit strides (rows: stride_line = ld, stride_elem = 1; columns: the transpose): the position of the
 * ground truth in the descending sort of compute_metric_ret (sparsify_clip.py:374-378, :396-400). */
int scb_rank_count(const void* S, int64_t n_lines, int64_t n_elem, int64_t stride_line, int64_t stride_elem, int dtype,
                   const int64_t* line, const int64_t* gt, int* rank, void* stream);
/* The same rank from the FEATURES: cnt[part*nsub + s][i] partial counts of #{j != i + diag_off : A_i . Bm_j > gt_score[i]}
 * (sum them with scb_sum_parts); the N x N similarity of sparsify_clip.py:628 is produced tile by tile and never stored.
 * jparts / nsub from scb_pass_plan(path, nA, nB, D, grad = 0, ...). */
int scb_rank_count_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                        const float* gt_score, int64_t diag_off, int jparts, float* cnt, int path, void* stream);

/* ------------------------------------------------------------------ SM-free all-gather inside one node
 * Every rank pushes its row shard into every peer's gather buffer with the copy engines and then writes an epoch flag;
 * the consumer waits on its own flags with a one-warp kernel.  The buffers are the one thing the library allocates
 * (explicit alloc / open / close, like a communicator handle): scb_peer_alloc -> device memory + a 64-byte CUDA IPC
 * handle to hand to the other processes; scb_peer_open maps a peer's buffer (peer access enabled lazily);
 * scb_peer_close(ptr, opened) unmaps (opened = 1) or frees (opened = 0). */
int scb_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int scb_peer_open(const unsigned char* handle64, void** ptr);
int scb_peer_close(void* ptr, int opened);
/* One gather epoch of one role.  State per rank: a device int `epoch`, and in the IPC-shared buffer two int arrays
 * arrived[world], done[world].  All calls are stream-ordered and read the epoch on the device (graph-replayable).
 * scb_peer_begin (consumer's stream): ++epoch; wait until every peer released the previous epoch (done[p] >= epoch - 1).
 * scb_peer_push (any stream ordered after begin): copy `bytes` from src to each dst[k] (k < n, copy engines), then store
 *   epoch into *arrived_words[p] for every p < world (arrived_words / done_words: DEVICE arrays of `world` peer-mapped
 *   pointers to this rank's word on each rank).
 * scb_peer_wait: until arrived[p] >= epoch for every p.   scb_peer_release: store epoch into *done_words[p]. */
int scb_peer_begin(int* epoch, const int* done, int world, void* stream);
int scb_peer_push(const void* src, int64_t bytes, void* const* dst, int n, const int* epoch, int* const* arrived_words,
                  int world, void* stream);
/* scb_peer_push with the copies done by the SMs (16-byte stores over NVLink, every destination written by one kernel;
 * the last CTA publishes the flags): for the gather a step starts with, when no sweep occupies the SMs yet.
 * counter: zero-initialised device word owned by the role (every launch leaves it at zero).  bytes % 16 == 0. */
int scb_peer_push_sm(const void* src, int64_t bytes, void* const* dst, int n, const int* epoch, int* const* arrived_words,
                     int world, unsigned* counter, void* stream);
/* the copies of scb_peer_push alone (several streams may share a large shard's pushes; scb_peer_push with n = 0 sends the flags) */
int scb_peer_copy(const void* src, int64_t bytes, void* const* dst, int n, void* stream);
int scb_peer_wait(const int* epoch, const int* arrived, int world, void* stream);
int scb_peer_release(const int* epoch, int* const* done_words, int world, void* stream);
/* the same for up to 8 roles in one launch (host arrays of the per-role arguments) */
int scb_peer_release_many(const int* const* epochs, int* const* const* done_words, int n_roles, int world, void* stream);

/* debug: device buffer of 2 x 4 x 4096 x 2 uint64 that the CTA-pair kernel fills with a per-role timeline of
 * cluster 0 (tag, tile, clock64) on the following launches; NULL switches it off (tools/pair_trace.py). */
int scb_debug_pair_trace(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* SCB200_H_ */
