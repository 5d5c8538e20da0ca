// simt_pass.cu -- SCB_PATH_SIMT: fp32 CUDA-core implementation of the B x B passes.
// Exact path for fp32 inputs (and for shapes the tensor-core path does not take): every
// product and accumulation is an fp32 FMA, exponentials use exp2f.  Same partial-output
// protocol as the tcgen05 path (tc_pass.cu), so the finalisers in rowwise.cu are shared.
//
// One CTA owns 32 rows of A and a 128-wide slice of the output columns, and sweeps its
// part of the columns of S in 32-wide tiles:  S tile (fp32 registers) -> weights ->
// shared memory -> rank-32 update of the 32x128 output slice.  Each warp owns 4 rows, so
// every per-row statistic is a warp shuffle reduction.  S is never written to memory.
#include "common.cuh"

namespace {

enum { M_LSE = 0, M_ANCHOR_GRAD = 1, M_LUNIF_GRAD = 2, M_LUNIF_SUM = 3, M_SPARSIFY_SUM = 4, M_RANK_COUNT = 6 };

constexpr int TR = 32;    // rows per CTA
constexpr int TJ = 32;    // columns of S per tile
constexpr int KC = 32;    // K chunk
constexpr int DC = 128;   // output columns per CTA

struct SimtParams {
  const void* A; const void* Bm;
  int64_t nA, nB, ldA, ldB;
  int D, dtype, jparts;
  float p0;                 // LSE/anchor: scale*log2e ; lunif: t*log2e
  const float* p0_dev;      // optional device multiplier of p0 (1/tau of a device-resident temperature)
  const float* rowvec;      // anchor: row lse (natural log) ; lunif: sq norms of the A rows
  const float* colvec;      // anchor: col lse ; lunif: sq norms of the Bm rows
  int64_t diag_off;         // column j is "the diagonal" of local row i when j == i + diag_off
  float* out;               // [jparts][nA][D]
  float* s0;                // LSE: part_m ; anchor: ws ; lunif: rq ; sums: rs
  float* s1;                // LSE: part_l ; lunif grad: rs
};

template <int MODE>
__global__ void __launch_bounds__(256) k_simt_pass(const SimtParams P) {
  constexpr bool GRAD = (MODE == M_ANCHOR_GRAD || MODE == M_LUNIF_GRAD);
  const float p0_eff = P.p0_dev ? P.p0 * __ldg(P.p0_dev) : P.p0;
  __shared__ float As[TR][KC + 1];
  __shared__ float Bs[TJ][KC + 1];
  __shared__ float Ws[GRAD ? TR : 1][TJ + 1];
  __shared__ float Vs[GRAD ? TJ : 1][GRAD ? DC : 1];

  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * TR;
  const int c0 = blockIdx.y * DC;
  const int part = blockIdx.z;
  const int64_t n_jt = (P.nB + TJ - 1) / TJ;
  const int64_t jt_lo = n_jt * part / P.jparts, jt_hi = n_jt * (part + 1) / P.jparts;

  // per-row constants / running statistics (identical in all lanes of the warp)
  float rowc[4], st0[4], st1[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int64_t gi = r0 + ty * 4 + rr;
    const bool ok = gi < P.nA;
    rowc[rr] = 0.f;
    if (MODE == M_ANCHOR_GRAD) rowc[rr] = ok ? P.rowvec[gi] * SCB_LOG2E : 0.f;
    if (MODE == M_LUNIF_GRAD || MODE == M_LUNIF_SUM) rowc[rr] = ok ? P.rowvec[gi] : 0.f;
    if (MODE == M_RANK_COUNT) rowc[rr] = ok ? P.rowvec[gi] : INFINITY;
    st0[rr] = (MODE == M_LSE) ? -INFINITY : 0.f;
    st1[rr] = 0.f;
  }
  float oacc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) oacc[a][b] = 0.f;

  for (int64_t jt = jt_lo; jt < jt_hi; ++jt) {
    const int64_t j0 = jt * TJ;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < P.D; k0 += KC) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + 256 * i, r = e >> 5, k = e & 31;
        const int64_t ga = r0 + r, gb = j0 + r;
        const bool kin = (k0 + k) < P.D;
        As[r][k] = (kin && ga < P.nA) ? scb_ld(P.A, P.dtype, ga * P.ldA + k0 + k) : 0.f;
        Bs[r][k] = (kin && gb < P.nB) ? scb_ld(P.Bm, P.dtype, gb * P.ldB + k0 + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const float b = Bs[tx][k];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) acc[rr] = fmaf(As[ty * 4 + rr][k], b, acc[rr]);
      }
    }
    // ---- weights for this 32x32 tile
    const int64_t gj = j0 + tx;
    const bool jok = gj < P.nB;
    float colc = 0.f;
    if (MODE == M_ANCHOR_GRAD) colc = jok ? P.colvec[gj] * SCB_LOG2E : 0.f;
    if (MODE == M_LUNIF_GRAD || MODE == M_LUNIF_SUM) colc = jok ? P.colvec[gj] : 0.f;
    float w[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int64_t gi = r0 + ty * 4 + rr;
      const bool is_diag = (gj == gi + P.diag_off);
      const float g = acc[rr];
      if (MODE == M_LSE) {
        const float y = jok ? g * p0_eff : -INFINITY;
        const float tmax = scb_warp_max(y);
        const float mnew = fmaxf(st0[rr], tmax);
        float e = (mnew == -INFINITY) ? 0.f : exp2f(y - mnew);
        e = scb_warp_sum(e);
        const float corr = (st0[rr] == -INFINITY) ? 0.f : exp2f(st0[rr] - mnew);
        st1[rr] = st1[rr] * corr + e;
        st0[rr] = mnew;
      } else if (MODE == M_ANCHOR_GRAD) {
        const float y = g * p0_eff;
        float ww = jok ? (exp2f(y - rowc[rr]) + exp2f(y - colc)) : 0.f;
        st0[rr] += scb_warp_sum(ww * g);           // sum_j w_ij (a_i.b_j), diagonal included
        w[rr] = is_diag ? 0.f : ww;
      } else if (MODE == M_LUNIF_GRAD || MODE == M_LUNIF_SUM) {
        const float d2 = fmaxf(rowc[rr] + colc - 2.f * g, 0.f);
        const float ww = (jok && !is_diag) ? exp2f(-p0_eff * d2) : 0.f;
        const float s = scb_warp_sum(ww);
        st0[rr] += s;
        st1[rr] += s;
        w[rr] = ww;
      } else if (MODE == M_RANK_COUNT) {
        st0[rr] += scb_warp_sum((jok && !is_diag && g > rowc[rr]) ? 1.f : 0.f);
      } else {  // M_SPARSIFY_SUM
        const float e = g - (is_diag ? 1.f : -1.f);
        st0[rr] += scb_warp_sum(jok ? e * e : 0.f);
      }
    }
    if (GRAD) {
      __syncthreads();
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) Ws[ty * 4 + rr][tx] = w[rr];
#pragma unroll
      for (int i = 0; i < (TJ * DC) / 256; ++i) {
        const int e = tid + 256 * i, j = e / DC, d = e % DC;
        const int64_t gb = j0 + j;
        Vs[j][d] = (gb < P.nB && (c0 + d) < P.D) ? scb_ld(P.Bm, P.dtype, gb * P.ldB + c0 + d) : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < TJ; ++k) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = Vs[k][tx + 32 * c];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float ww = Ws[ty * 4 + rr][k];
#pragma unroll
          for (int c = 0; c < 4; ++c) oacc[rr][c] = fmaf(ww, v[c], oacc[rr][c]);
        }
      }
    }
  }

  // ---- write partials
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int64_t gi = r0 + ty * 4 + rr;
    if (gi >= P.nA) continue;
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c0 + tx + 32 * c;
        if (d < P.D) P.out[((int64_t)part * P.nA + gi) * P.D + d] = oacc[rr][c];
      }
    }
    if (tx == 0 && blockIdx.y == 0) {
      const int64_t o = (int64_t)part * P.nA + gi;
      if (P.s0) P.s0[o] = st0[rr];
      if (P.s1) P.s1[o] = st1[rr];
    }
  }
}

template <int MODE>
int launch(const SimtParams& P, cudaStream_t s) {
  if (P.nA == 0) return 0;
  constexpr bool GRAD = (MODE == M_ANCHOR_GRAD || MODE == M_LUNIF_GRAD);
  dim3 grid((unsigned)((P.nA + TR - 1) / TR), GRAD ? (unsigned)((P.D + DC - 1) / DC) : 1u, (unsigned)P.jparts);
  k_simt_pass<MODE><<<grid, 256, 0, s>>>(P);
  SCB_CHECK_LAUNCH("simt_pass");
  return 0;
}

}  // namespace

// entry points used by api.cu -------------------------------------------------------------
int scb_simt_lse(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                 float scale, int jparts, float* pm, float* pl, const float* scale_dev, cudaStream_t s) {
  SimtParams P{A, Bm, nA, nB, ldA, ldB, D, dtype, jparts, scale * SCB_LOG2E, scale_dev, nullptr, nullptr, INT64_MIN / 2, nullptr, pm, pl};
  return launch<M_LSE>(P, s);
}
int scb_simt_anchor_grad(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                         float scale, const float* row_lse, const float* col_lse, int64_t diag_off, int jparts,
                         float* out, float* ws, const float* scale_dev, cudaStream_t s) {
  SimtParams P{A, Bm, nA, nB, ldA, ldB, D, dtype, jparts, scale * SCB_LOG2E, scale_dev, row_lse, col_lse, diag_off, out, ws, nullptr};
  return launch<M_ANCHOR_GRAD>(P, s);
}
int scb_simt_lunif(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll, int dtype,
                   float t, const float* sqn_r, const float* sqn_all, int64_t row_offset, int jparts, float* U,
                   float* rq, float* rs, cudaStream_t s) {
  SimtParams P{Xr, Xall, nR, nAll, ldR, ldAll, D, dtype, jparts, t * SCB_LOG2E, nullptr, sqn_r, sqn_all, row_offset, U, rq, rs};
  return U ? launch<M_LUNIF_GRAD>(P, s) : launch<M_LUNIF_SUM>(P, s);
}
int scb_simt_rank_count(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                        const float* gt_score, int64_t diag_off, int jparts, float* cnt, cudaStream_t s) {
  SimtParams P{A, Bm, nA, nB, ldA, ldB, D, dtype, jparts, 0.f, nullptr, gt_score, nullptr, diag_off, nullptr, cnt, nullptr};
  return launch<M_RANK_COUNT>(P, s);
}
int scb_simt_sparsify_sum(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                          int dtype, int64_t row_offset, int jparts, float* rs, cudaStream_t s) {
  SimtParams P{Xr, Xall, nR, nAll, ldR, ldAll, D, dtype, jparts, 0.f, nullptr, nullptr, nullptr, row_offset, nullptr, rs, nullptr};
  return launch<M_SPARSIFY_SUM>(P, s);
}
