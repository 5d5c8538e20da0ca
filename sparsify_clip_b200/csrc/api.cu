// api.cu -- extern "C" entry points of the B x B passes declared in include/scb200.h:
// argument validation and dispatch to the SIMT (simt_pass.cu) or tensor-core (tc_pass.cu) path.
#include "common.cuh"

int scb_simt_lse(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, int, float*, float*, const float*,
                 cudaStream_t);
int scb_simt_anchor_grad(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*,
                         const float*, int64_t, int, float*, float*, const float*, cudaStream_t);
int scb_simt_lunif(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*, const float*,
                   int64_t, int, float*, float*, float*, cudaStream_t);
int scb_simt_sparsify_sum(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, int64_t, int, float*, cudaStream_t);
int scb_tc_lse(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, int, float*, float*, const int*,
               const float*, cudaStream_t);
int scb_tc_lse2(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, int, float*, float*, float*, float*,
                const float*, cudaStream_t);
int scb_tc_anchor_grad(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*, const float*,
                       int64_t, int, float*, float*, const float*, cudaStream_t);
int scb_tc_lunif(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*, const float*, int64_t,
                 int, float*, float*, float*, cudaStream_t);
int scb_tc_sparsify_sum(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, int64_t, int, float*, cudaStream_t);
int scb_simt_rank_count(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, const float*, int64_t, int, float*,
                        cudaStream_t);
int scb_tc_rank_count(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, const float*, int64_t, int, float*,
                      cudaStream_t);
int scb_tc_set_flags(int);
int scb_tc_grad_kernel(int64_t nA, int D, int grad);
int scb_quad_clusters();
void scb_pair_span_plan(int64_t n_rb, int64_t n_jb, int n_sm, int* n_pairs, int64_t* span, int* pmax);
void scb_quad_span_plan(int64_t n_rp, int64_t n_jb, int n_clusters, int align, int* n_used, int64_t* span, int* pmax);
struct QuadSplit { int64_t rows_quad; int side_pairs; int jparts; };
QuadSplit scb_quad_split(int64_t nA, int64_t nB, int D, int n_sm);

#define SCB_PASS_CHECKS(A, nA, Bm, nB, D, ldA, ldB, dtype, jparts, path)                                           \
  SCB_CHECK_ARG(scb_dtype_ok(dtype), SCB_E_DTYPE, "%s: unsupported dtype %d", __func__, (int)(dtype));             \
  SCB_CHECK_ARG((path) == SCB_PATH_SIMT || (path) == SCB_PATH_TC, SCB_E_ARG, "%s: unknown path %d", __func__, (int)(path)); \
  SCB_CHECK_ARG((nA) >= 0 && (nB) >= 0 && (D) > 0 && (ldA) >= (D) && (ldB) >= (D), SCB_E_ARG,                      \
                "%s: bad shape nA=%lld nB=%lld D=%d ldA=%lld ldB=%lld", __func__, (long long)(nA), (long long)(nB), \
                (int)(D), (long long)(ldA), (long long)(ldB));                                                     \
  SCB_CHECK_ARG(((A) && (Bm)) || (nA) == 0, SCB_E_ARG, "%s: null operand", __func__);                              \
  SCB_CHECK_ARG((jparts) >= 1, SCB_E_ARG, "%s: jparts must be >= 1", __func__)

// Split the column sweep so that (row blocks x column groups x parts) fills the execution units evenly.
// cost model: rounds of `n_units` concurrent work items x (tiles per item + `overhead` tiles of prologue/drain:
// 1 for the single-CTA kernels, 4 for the CTA-pair kernel whose software pipeline is three steps deep).
static int scb_choose_jparts(int64_t n_rb, int nsplit, int64_t n_jb, int n_units, double overhead = 1.0, int max_parts = 16) {
  int best = 1;
  double best_cost = -1.0;
  const int hi = (int)(n_jb < max_parts ? n_jb : max_parts);
  for (int jp = 1; jp <= (hi < 1 ? 1 : hi); ++jp) {
    const int64_t rounds = (n_rb * nsplit * jp + n_units - 1) / n_units;
    const double cost = (double)rounds * ((double)((n_jb + jp - 1) / jp) + overhead);
    if (best_cost < 0.0 || cost < best_cost - 1e-9) { best = jp; best_cost = cost; }
  }
  return best;
}

extern "C" int scb_pass_plan(int path, int64_t nA, int64_t nB, int D, int grad, int n_sm, int* jparts, int* nsub) {
  SCB_CHECK_ARG(path == SCB_PATH_SIMT || path == SCB_PATH_TC, SCB_E_ARG, "pass_plan: unknown path %d", path);
  SCB_CHECK_ARG(nA >= 0 && nB >= 0 && D > 0 && n_sm > 0 && jparts && nsub, SCB_E_ARG, "pass_plan: bad argument");
  if (path == SCB_PATH_TC) {
    const int kch = (D + 63) / 64;
    const int64_t n_rb = (nA + 127) / 128, n_jb = (nB + 127) / 128;
    const int kern = n_sm >= 2 ? scb_tc_grad_kernel(nA, D, grad) : 0;
    if (kern == 2) {            // equal contiguous spans of (256-row block, tile) per cluster of 4 (tc_quad.cu), the last
      *jparts = scb_quad_split(nA, nB, D, n_sm).jparts;      // rows on CTA pairs; cluster count = a property of the device
      *nsub = 4;
    } else if (kern == 1) {     // equal contiguous spans of (row block, tile) per CTA pair
      int np = 0;
      int64_t span = 0;
      scb_pair_span_plan(n_rb, n_jb, n_sm, &np, &span, jparts);
      *nsub = 4;
    } else {
      *jparts = scb_choose_jparts(n_rb, grad ? (kch + 3) / 4 : 1, n_jb, n_sm);
      *nsub = 2;
    }
  } else {
    const int64_t tiles = ((nA + 31) / 32) * (grad ? (D + 127) / 128 : 1);
    int64_t jp = (2 * (int64_t)n_sm) / (tiles > 0 ? tiles : 1);
    const int64_t cap = (nB + 31) / 32;
    if (jp > cap) jp = cap;
    *jparts = (int)(jp < 1 ? 1 : jp);
    *nsub = 1;
  }
  return 0;
}
extern "C" int scb_set_tc_flags(int flags) { return scb_tc_set_flags(flags); }
extern "C" int scb_grad_kernel_kind(int64_t nA, int D, int n_sm, int* units) {
  const int kern = n_sm >= 2 ? scb_tc_grad_kernel(nA, D, 1) : 0;
  if (units) *units = kern == 2 ? scb_quad_clusters() : (kern == 1 ? n_sm / 2 : n_sm);
  return kern;
}

extern "C" int scb_lse_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                            float scale, int jparts, float* part_m, float* part_l, int path, const float* scale_dev,
                            void* stream) {
  SCB_PASS_CHECKS(A, nA, Bm, nB, D, ldA, ldB, dtype, jparts, path);
  SCB_CHECK_ARG((part_m && part_l) || nA == 0, SCB_E_ARG, "lse_pass: null output");
  SCB_CHECK_ARG(scale > 0.f && nB > 0, SCB_E_ARG, "lse_pass: needs scale > 0 and nB > 0");
  cudaStream_t s = (cudaStream_t)stream;
  return path == SCB_PATH_TC ? scb_tc_lse(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, jparts, part_m, part_l, nullptr, scale_dev, s)
                             : scb_simt_lse(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, jparts, part_m, part_l, scale_dev, s);
}

// the same sweep launched conditionally: the kernel returns at once unless *run_flag != 0 (tensor-core path only)
extern "C" int scb_lse_pass_cond(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB,
                                 int dtype, float scale, int jparts, float* part_m, float* part_l, const int* run_flag,
                                 const float* scale_dev, void* stream) {
  SCB_PASS_CHECKS(A, nA, Bm, nB, D, ldA, ldB, dtype, jparts, SCB_PATH_TC);
  SCB_CHECK_ARG((part_m && part_l && run_flag) || nA == 0, SCB_E_ARG, "lse_pass_cond: null argument");
  SCB_CHECK_ARG(scale > 0.f && nB > 0, SCB_E_ARG, "lse_pass_cond: needs scale > 0 and nB > 0");
  return scb_tc_lse(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, jparts, part_m, part_l, run_flag, scale_dev, (cudaStream_t)stream);
}

extern "C" int scb_lse2_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                             float scale, int jparts, float* part_m, float* part_l, float* col_ref, float* col_sum,
                             const float* scale_dev, void* stream) {
  SCB_PASS_CHECKS(A, nA, Bm, nB, D, ldA, ldB, dtype, jparts, SCB_PATH_TC);
  SCB_CHECK_ARG((part_m && part_l && col_ref && col_sum) || nA == 0, SCB_E_ARG, "lse2_pass: null output");
  SCB_CHECK_ARG(scale > 0.f && nB > 0, SCB_E_ARG, "lse2_pass: needs scale > 0 and nB > 0");
  return scb_tc_lse2(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, jparts, part_m, part_l, col_ref, col_sum, scale_dev,
                     (cudaStream_t)stream);
}

extern "C" int scb_anchor_grad_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB,
                                    int dtype, float scale, const float* row_lse, const float* col_lse, int64_t diag_off,
                                    int jparts, float* out, float* ws, int path, const float* scale_dev, void* stream) {
  SCB_PASS_CHECKS(A, nA, Bm, nB, D, ldA, ldB, dtype, jparts, path);
  SCB_CHECK_ARG((row_lse && col_lse && out) || nA == 0, SCB_E_ARG, "anchor_grad_pass: null argument");
  SCB_CHECK_ARG(scale > 0.f && nB > 0, SCB_E_ARG, "anchor_grad_pass: needs scale > 0 and nB > 0");
  cudaStream_t s = (cudaStream_t)stream;
  return path == SCB_PATH_TC
             ? scb_tc_anchor_grad(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, row_lse, col_lse, diag_off, jparts, out, ws, scale_dev, s)
             : scb_simt_anchor_grad(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, row_lse, col_lse, diag_off, jparts, out, ws, scale_dev,
                                    s);
}

extern "C" int scb_lunif_pass(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                              int dtype, float t, const float* sqn_r, const float* sqn_all, int64_t row_offset, int jparts,
                              float* U, float* rq, float* rs, int path, void* stream) {
  SCB_PASS_CHECKS(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, jparts, path);
  SCB_CHECK_ARG((sqn_r && sqn_all && U && rq && rs) || nR == 0, SCB_E_ARG, "lunif_pass: null argument");
  SCB_CHECK_ARG(nAll > 0, SCB_E_ARG, "lunif_pass: empty column side");
  cudaStream_t s = (cudaStream_t)stream;
  return path == SCB_PATH_TC
             ? scb_tc_lunif(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, t, sqn_r, sqn_all, row_offset, jparts, U, rq, rs, s)
             : scb_simt_lunif(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, t, sqn_r, sqn_all, row_offset, jparts, U, rq, rs, s);
}

extern "C" int scb_lunif_sum_pass(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                                  int dtype, float t, const float* sqn_r, const float* sqn_all, int64_t row_offset, int jparts,
                                  float* rs, int path, void* stream) {
  SCB_PASS_CHECKS(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, jparts, path);
  SCB_CHECK_ARG((sqn_r && sqn_all && rs) || nR == 0, SCB_E_ARG, "lunif_sum_pass: null argument");
  SCB_CHECK_ARG(nAll > 0, SCB_E_ARG, "lunif_sum_pass: empty column side");
  cudaStream_t s = (cudaStream_t)stream;
  return path == SCB_PATH_TC
             ? scb_tc_lunif(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, t, sqn_r, sqn_all, row_offset, jparts, nullptr, nullptr, rs, s)
             : scb_simt_lunif(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, t, sqn_r, sqn_all, row_offset, jparts, nullptr, nullptr, rs, s);
}

extern "C" int scb_sparsify_sum_pass(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                                     int dtype, int64_t row_offset, int jparts, float* rs, int path, void* stream) {
  SCB_PASS_CHECKS(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, jparts, path);
  SCB_CHECK_ARG(rs || nR == 0, SCB_E_ARG, "sparsify_sum_pass: null output");
  SCB_CHECK_ARG(nAll > 0, SCB_E_ARG, "sparsify_sum_pass: empty column side");
  cudaStream_t s = (cudaStream_t)stream;
  return path == SCB_PATH_TC ? scb_tc_sparsify_sum(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, row_offset, jparts, rs, s)
                             : scb_simt_sparsify_sum(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, row_offset, jparts, rs, s);
}

// cnt[part][i] = #{j != i + diag_off : A_i . Bm_j > gt_score[i]}: the rank of row i's ground-truth pair among the columns,
// from the features (the N x N score matrix of sparsify_clip.py:628 is never materialised)
extern "C" int scb_rank_count_pass(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB,
                                   int dtype, const float* gt_score, int64_t diag_off, int jparts, float* cnt, int path,
                                   void* stream) {
  SCB_PASS_CHECKS(A, nA, Bm, nB, D, ldA, ldB, dtype, jparts, path);
  SCB_CHECK_ARG((gt_score && cnt) || nA == 0, SCB_E_ARG, "rank_count_pass: null argument");
  SCB_CHECK_ARG(nB > 0, SCB_E_ARG, "rank_count_pass: empty column side");
  cudaStream_t s = (cudaStream_t)stream;
  return path == SCB_PATH_TC ? scb_tc_rank_count(A, nA, Bm, nB, D, ldA, ldB, dtype, gt_score, diag_off, jparts, cnt, s)
                             : scb_simt_rank_count(A, nA, Bm, nB, D, ldA, ldB, dtype, gt_score, diag_off, jparts, cnt, s);
}
