// rowwise.cu -- O(B*D) kernels of the loss path: norms, diagonal logits, L_align, centroids,
// normalise, deterministic sums and the per-term gradient finalisers.  One warp per row,
// 16-byte vector loads when the layout allows (VEC), scalar otherwise.  HBM-bound; each
// input row is read exactly once per kernel.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
void scb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* scb_last_error(void) { return g_err; }
extern "C" int scb_version(void) { return SCB_ABI_VERSION; }

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

inline dim3 row_grid(int64_t n) { return dim3((unsigned)((n + kWarpsPerBlock - 1) / kWarpsPerBlock)); }

inline bool vec_ok(const void* p, int64_t ld, int D) { return scb_aligned16(p) && (ld % 8 == 0) && (D % 8 == 0); }

__device__ __forceinline__ float eff_scale(float host_scale, const float* dev_scale) {
  return dev_scale ? host_scale * __ldg(dev_scale) : host_scale;
}

// Generic per-row visitor: calls f(d, a[8], b[8], nvalid) over the row in a warp-strided fashion.
template <bool VEC, bool TWO, class F>
__device__ __forceinline__ void for_row(const void* A, int64_t ldA, const void* B, int64_t ldB, int dtype,
                                        int64_t row, int D, int lane, F f) {
  if (VEC) {
    for (int d = lane * 8; d < D; d += 256) {
      float a[8], b[8];
      scb_ld8(A, dtype, row * ldA + d, a);
      if (TWO) scb_ld8(B, dtype, row * ldB + d, b);
      f(d, a, b, 8);
    }
  } else {
    for (int d = lane; d < D; d += 32) {
      float a[8], b[8];
      a[0] = scb_ld(A, dtype, row * ldA + d);
      b[0] = TWO ? scb_ld(B, dtype, row * ldB + d) : 0.f;
      f(d, a, b, 1);
    }
  }
}

// ------------------------------------------------------------------ reductions per row
// MODE 0: sum a^2 ; 1: sum a*b ; 2: sum (a-b)^2
template <bool VEC, int MODE>
__global__ void __launch_bounds__(kThreads) k_row_reduce(const void* __restrict__ A, const void* __restrict__ B,
                                                         int64_t n, int D, int64_t ldA, int64_t ldB, int dtype,
                                                         float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n) return;
  float acc = 0.f;
  for_row<VEC, MODE != 0>(A, ldA, B, ldB, dtype, row, D, lane, [&](int, float(&a)[8], float(&b)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        if (MODE == 0) acc = fmaf(a[i], a[i], acc);
        if (MODE == 1) acc = fmaf(a[i], b[i], acc);
        if (MODE == 2) { float t = a[i] - b[i]; acc = fmaf(t, t, acc); }
      }
  });
  acc = scb_warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

template <int MODE>
int launch_row_reduce(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype, float* out,
                      cudaStream_t s) {
  if (n == 0) return 0;
  bool v = vec_ok(A, ldA, D) && (MODE == 0 || vec_ok(B, ldB, D));
  if (v) k_row_reduce<true, MODE><<<row_grid(n), kThreads, 0, s>>>(A, B, n, D, ldA, ldB, dtype, out);
  else k_row_reduce<false, MODE><<<row_grid(n), kThreads, 0, s>>>(A, B, n, D, ldA, ldB, dtype, out);
  SCB_CHECK_LAUNCH("row_reduce");
  return 0;
}

// ------------------------------------------------------------------ L_align backward
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_lalign_bwd(const void* __restrict__ X, const void* __restrict__ Y,
                                                         int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                                                         float host_scale, const float* __restrict__ dev_scale,
                                                         int accumulate, float* __restrict__ dX, float* __restrict__ dY) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n) return;
  const float s = eff_scale(host_scale, dev_scale);
  for_row<VEC, true>(X, ldX, Y, ldY, dtype, row, D, lane, [&](int d, float(&a)[8], float(&b)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const float g = s * (a[i] - b[i]);
        const int64_t o = row * (int64_t)D + d + i;
        if (dX) dX[o] = accumulate ? dX[o] + g : g;
        if (dY) dY[o] = accumulate ? dY[o] - g : -g;
      }
  });
}

// ------------------------------------------------------------------ centroid / normalise forward
// CENT: m = (a+b)/2 with eps clamp 1e-12 (F.normalize); else m = a, no eps (x / x.norm()).
template <bool VEC, bool CENT>
__global__ void __launch_bounds__(kThreads) k_unit_fwd(const void* __restrict__ A, const void* __restrict__ B, int64_t n,
                                                       int D, int64_t ldA, int64_t ldB, int dtype, void* __restrict__ C,
                                                       int out_dtype, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n) return;
  float acc = 0.f;
  for_row<VEC, CENT>(A, ldA, B, ldB, dtype, row, D, lane, [&](int, float(&a)[8], float(&b)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) { float m = CENT ? 0.5f * (a[i] + b[i]) : a[i]; acc = fmaf(m, m, acc); }
  });
  acc = scb_warp_sum(acc);
  const float nrm = sqrtf(acc);
  const float inv = CENT ? 1.f / fmaxf(nrm, 1e-12f) : 1.f / nrm;
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for_row<VEC, CENT>(A, ldA, B, ldB, dtype, row, D, lane, [&](int d, float(&a)[8], float(&b)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) { float m = CENT ? 0.5f * (a[i] + b[i]) : a[i]; scb_st(C, out_dtype, row * (int64_t)D + d + i, m * inv); }
  });
}

// backward of c = m * inv:  dm = (g - c (c.g)) * inv ;  CENT: dA,dB (+)= s*dm/2 ; else dA = dm
template <bool VEC, bool CENT>
__global__ void __launch_bounds__(kThreads) k_unit_bwd(const void* __restrict__ A, const void* __restrict__ B, int64_t n,
                                                       int D, int64_t ldA, int64_t ldB, int dtype,
                                                       const float* __restrict__ dC, const float* __restrict__ inv_norm,
                                                       float host_scale, const float* __restrict__ dev_scale,
                                                       int accumulate, float* __restrict__ dA, float* __restrict__ dB) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n) return;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for_row<VEC, CENT>(A, ldA, B, ldB, dtype, row, D, lane, [&](int d, float(&a)[8], float(&b)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        float c = (CENT ? 0.5f * (a[i] + b[i]) : a[i]) * inv;
        dot = fmaf(c, dC[row * (int64_t)D + d + i], dot);
      }
  });
  dot = scb_warp_sum(dot);
  const float s = eff_scale(host_scale, dev_scale) * (CENT ? 0.5f : 1.f) * inv;
  for_row<VEC, CENT>(A, ldA, B, ldB, dtype, row, D, lane, [&](int d, float(&a)[8], float(&b)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const int64_t o = row * (int64_t)D + d + i;
        float c = (CENT ? 0.5f * (a[i] + b[i]) : a[i]) * inv;
        float g = s * (dC[o] - c * dot);
        if (dA) dA[o] = accumulate ? dA[o] + g : g;
        if (CENT && dB) dB[o] = accumulate ? dB[o] + g : g;
      }
  });
}

// ------------------------------------------------------------------ deterministic sum
constexpr int kSumBlocks = 1024;
__global__ void __launch_bounds__(256) k_sum_stage1(const float* __restrict__ x, int64_t n, float* __restrict__ scratch) {
  // contiguous chunk per block, fixed strided order inside -> bitwise reproducible
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(lo + per, n);
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc += x[i];
  __shared__ float sm[8];
  acc = scb_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i];
    scratch[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_sum_stage2(const float* __restrict__ scratch, int nb, float* __restrict__ out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < nb; i += 256) acc += scratch[i];
  __shared__ float sm[8];
  acc = scb_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i];
    out[0] = t;
  }
}

// ------------------------------------------------------------------ LSE combine
__global__ void __launch_bounds__(256) k_lse_combine(const float* __restrict__ pm, const float* __restrict__ pl, int nparts,
                                                     int64_t n, float* __restrict__ lse) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float M = -INFINITY;
  for (int p = 0; p < nparts; ++p) M = fmaxf(M, pm[(int64_t)p * n + i]);
  float L = 0.f;
  for (int p = 0; p < nparts; ++p) {
    const float m = pm[(int64_t)p * n + i];
    if (m != -INFINITY) L += pl[(int64_t)p * n + i] * exp2f(m - M);
  }
  lse[i] = (M + log2f(L)) * SCB_LN2;   // log2 domain -> natural log
}

// same, but only when *run_flag != 0 (the conditional column sweep behind the fused row+column LSE pass)
__global__ void __launch_bounds__(256) k_lse_combine_cond(const float* __restrict__ pm, const float* __restrict__ pl, int nparts,
                                                          int64_t n, float* __restrict__ lse, const int* __restrict__ run_flag) {
  if (*run_flag == 0) return;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float M = -INFINITY;
  for (int p = 0; p < nparts; ++p) M = fmaxf(M, pm[(int64_t)p * n + i]);
  float L = 0.f;
  for (int p = 0; p < nparts; ++p) {
    const float m = pm[(int64_t)p * n + i];
    if (m != -INFINITY) L += pl[(int64_t)p * n + i] * exp2f(m - M);
  }
  lse[i] = (M + log2f(L)) * SCB_LN2;
}

// Fold of the per-strip column partials of the fused pass: sum_p csum[p][j] 2^{cref[p][j/32]}.
// A block owns 32 columns; its 8 warps split the partials (warp w takes p = w, w+8, ...: every load is one coalesced
// 128-byte row of csum), then the 8 (reference, sum) pairs per column meet in shared memory.  With one thread per column
// walking all 4*n_rb partials the fold was latency-bound: 325 us at c3 for 134 MB (now HBM-bound).
// FINAL: write the natural-log column LSE; else keep the (reference, sum) pair in the log2 domain.
template <bool FINAL>
__global__ void __launch_bounds__(256) k_colstat_fold(const float* __restrict__ cref, const float* __restrict__ csum, int nparts,
                                                      int64_t n, float* __restrict__ out0, float* __restrict__ out1) {
  // one block = 128 columns: every lane owns 4 consecutive columns (one 16-byte load per partial row when the layout
  // allows: 512 contiguous bytes per warp and row instead of 128), the 8 warps split the partial rows
  __shared__ float shM[8][128], shL[8][128];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t j0 = (int64_t)blockIdx.x * 128 + lane * 4;
  const int64_t nref = (n + 31) / 32;
  const int64_t rcol = j0 >> 5;                       // the 4 columns of a lane share one 32-column reference block
  const bool vec = (n % 4 == 0) && (j0 + 3 < n);
  float M[4], L[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { M[c] = -INFINITY; L[c] = 0.f; }
  if (j0 < n) {
    for (int p = w; p < nparts; p += 8) {
      const float r = __ldg(cref + (int64_t)p * nref + rcol);
      float v[4];
      if (vec) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(csum + (int64_t)p * n + j0));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = (j0 + c < n) ? __ldcs(csum + (int64_t)p * n + j0 + c) : 0.f;
      }
      if (r == -INFINITY) continue;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (v[c] == 0.f) continue;
        if (r > M[c]) { L[c] *= exp2f(M[c] - r); M[c] = r; }
        L[c] += v[c] * exp2f(r - M[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) { shM[w][lane * 4 + c] = M[c]; shL[w][lane * 4 + c] = L[c]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = threadIdx.x;
    const int64_t j = (int64_t)blockIdx.x * 128 + col;
    if (j < n) {
      float Mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) Mx = fmaxf(Mx, shM[k][col]);
      float Ls = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (shM[k][col] != -INFINITY) Ls += shL[k][col] * exp2f(shM[k][col] - Mx);
      if (FINAL) {
        out0[j] = (Mx + log2f(Ls)) * SCB_LN2;
      } else {
        out0[j] = Mx;
        out1[j] = Ls;
      }
    }
  }
}

// Scalar assembly of the composed loss from the step's partial sums (one thread; replaces ~20 one-element
// element-wise launches of the host framework per step).  p = [sum r, sum c, sum diag, sum |x-y|^2, rsI, rsT, rsC].
__global__ void k_loss_assemble(const float* __restrict__ p, float ca, float two_scale, float cl, float wi, float wt, float wc,
                                float pair_norm, float* __restrict__ loss, float* __restrict__ inv_ssum,
                                const float* __restrict__ scale_dev) {
  float L = 0.f;
  if (ca != 0.f) L += ca * (p[0] + p[1] - eff_scale(two_scale, scale_dev) * p[2]);
  if (cl != 0.f) L += cl * p[3];
  const float w[3] = {wi, wt, wc};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    inv_ssum[k] = 0.f;
    if (w[k] != 0.f) {
      const float ss = 0.5f * p[4 + k];
      L += w[k] * logf(ss / pair_norm);          // B == 1: log(0 / 0) = nan, as the reference
      inv_ssum[k] = 1.f / ss;
    }
  }
  *loss = L;
}

// Fold of the per-rank column partials after the packed gather of a sharded step.  pack = [world][stride] fp32, one row
// per rank: at off_exact the exact column LSE of that rank's own n_loc columns (valid when *flag != 0), at off_ref /
// off_sum that rank's (reference, sum) partial of every one of the world * n_loc columns (log2 domain).
__global__ void __launch_bounds__(256) k_fold_ranks(const float* __restrict__ pack, int world, int64_t stride, int64_t n_loc,
                                                    int64_t off_exact, int64_t off_ref, int64_t off_sum,
                                                    const int* __restrict__ flag, float* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n_loc * world) return;
  if (*flag != 0) {
    out[j] = pack[(j / n_loc) * stride + off_exact + j % n_loc];
    return;
  }
  float Mx = -INFINITY;
  for (int r = 0; r < world; ++r) Mx = fmaxf(Mx, pack[r * stride + off_ref + j]);
  float Ls = 0.f;
  for (int r = 0; r < world; ++r) {
    const float m = pack[r * stride + off_ref + j];
    if (m != -INFINITY) Ls += pack[r * stride + off_sum + j] * exp2f(m - Mx);
  }
  out[j] = (Mx + log2f(Ls)) * SCB_LN2;
}

// flag = 1 when the logits can spread by more than `bound` (log2 units) inside one block of the fused pass:
// 2 * scale * log2e * max_i |a_i| * max_j |b_j| >= bound   (sqn = squared row norms)
__global__ void __launch_bounds__(1024) k_spread_flag(const float* __restrict__ sqnA, int64_t nA, const float* __restrict__ sqnB,
                                                      int64_t nB, float scale, const float* __restrict__ scale_dev, float bound,
                                                      int* __restrict__ flag) {
  __shared__ float sh[2][32];
  float ma = 0.f, mb = 0.f;
  for (int64_t i = threadIdx.x; i < nA; i += 1024) ma = fmaxf(ma, sqnA[i]);
  for (int64_t i = threadIdx.x; i < nB; i += 1024) mb = fmaxf(mb, sqnB[i]);
  ma = scb_warp_max(ma); mb = scb_warp_max(mb);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ma; sh[1][threadIdx.x >> 5] = mb; }
  __syncthreads();
  if (threadIdx.x < 32) {
    ma = scb_warp_max(sh[0][threadIdx.x]); mb = scb_warp_max(sh[1][threadIdx.x]);
    if (threadIdx.x == 0) {
      const float spread = 2.f * eff_scale(scale, scale_dev) * SCB_LOG2E * sqrtf(ma) * sqrtf(mb);
      *flag = (spread >= bound || !(spread == spread)) ? 1 : 0;      // NaN/inf norms -> take the exact sweep
    }
  }
}

// ------------------------------------------------------------------ gradient finalisers
// dA[i,:] (+)= s * ( sum_p out[p][i,:] + dcoef_i * V[i,:] )
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_anchor_fin(const float* __restrict__ out, int jparts, int64_t n, int D,
                                                         const void* __restrict__ V, int64_t ldV, int dtype,
                                                         const float* __restrict__ row_lse,
                                                         const float* __restrict__ col_lse_rows,
                                                         const float* __restrict__ diag, float scale, float host_scale,
                                                         const float* __restrict__ dev_scale, int accumulate,
                                                         float* __restrict__ dA, const float* __restrict__ scale_dev) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n) return;
  const float sd = scale_dev ? __ldg(scale_dev) : 1.f;       // a device-resident 1/tau multiplies both uses of the scale
  const float s = eff_scale(host_scale, dev_scale) * sd;
  const float sii = scale * sd * diag[row];
  const float dcoef = expf(sii - row_lse[row]) + expf(sii - col_lse_rows[row]) - 2.f;
  for_row<VEC, false>(V, ldV, nullptr, 0, dtype, row, D, lane, [&](int d, float(&a)[8], float(&)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const int64_t o = row * (int64_t)D + d + i;
        float acc = 0.f;
        for (int p = 0; p < jparts; ++p) acc += out[(int64_t)p * n * D + o];
        const float g = s * fmaf(dcoef, a[i], acc);
        dA[o] = accumulate ? dA[o] + g : g;
      }
  });
}

// dX[i,:] (+)= s * ( rq_i * x_i - sum_p U[p][i,:] )
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_lunif_fin(const float* __restrict__ U, int jparts,
                                                        const float* __restrict__ rq, int nparts_rq, int64_t n, int D,
                                                        const void* __restrict__ X, int64_t ld, int dtype,
                                                        float host_scale, const float* __restrict__ dev_scale,
                                                        int accumulate, float* __restrict__ dX) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= n) return;
  const float s = eff_scale(host_scale, dev_scale);
  float r = 0.f;
  for (int p = 0; p < nparts_rq; ++p) r += rq[(int64_t)p * n + row];
  for_row<VEC, false>(X, ld, nullptr, 0, dtype, row, D, lane, [&](int d, float(&a)[8], float(&)[8], int nv) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nv) {
        const int64_t o = row * (int64_t)D + d + i;
        float acc = 0.f;
        for (int p = 0; p < jparts; ++p) acc += U[(int64_t)p * n * D + o];
        const float g = s * fmaf(r, a[i], -acc);
        dX[o] = accumulate ? dX[o] + g : g;
      }
  });
}

// ------------------------------------------------------------------ fused gradient combine
// One pass over an operand that applies every term of the composed loss at once:
//   dX[i,:] = gs * ( a_coef * ( sum_p a_out[p][i,:] + dcoef_i * Y[i,:] )          anchor  (dcoef_i = P_ii + Q_ii - 2)
//                  + u_coef * ( (sum_q rq[q][i]) * X[i,:] - sum_p u_out[p][i,:] )  L_unif
//                  + l_coef * ( X[i,:] - Y[i,:] )                                  L_align
//                  + e_coef * extra[i,:] )                                         a finished fp32 term (centroid chain)
// Each thread owns 8 consecutive columns of one row (16-byte loads of every partial, one 16/32-byte store), so the
// kernel streams at HBM speed; the per-row scalars are recomputed per thread (a few L1-resident loads).
struct CombineArgs {
  const void* X; const void* Y; int64_t n; int D; int64_t ldX, ldY; int dtype;
  const float* a_out; int a_jparts; const float* row_lse; const float* col_lse_rows; const float* diag; float scale;
  float a_coef;
  const float* u_out; int u_jparts; const float* rq; int rq_parts; float u_coef; const float* u_dev_coef;
  float l_coef;
  const float* extra; float e_coef;
  const float* dev_scale;
  void* dX; int out_dtype; int64_t ldOut;
  const float* scale_dev;      // optional device multiplier of `scale` and `a_coef` (1/tau of a device-resident temperature)
  // optional pull-back through the pre-loss normalise X = E / ||E|| (sparsify_clip.py:772-773): the un-normalised rows
  // and their 1 / ||E||; the kernel then writes dE = (g - e (e . g)) / ||E|| with e = E / ||E|| in fp32
  const void* unit_src; int64_t ld_unit; int unit_dtype; const float* unit_inv;
};

// every term of `row`, columns [d, d + W): g before the output conversion (everything of the combine but the store)
template <bool VEC>
__device__ __forceinline__ void combine_terms(const CombineArgs& a, int64_t row, int d, float (&g)[8]) {
  constexpr int W = VEC ? 8 : 1;
  float x[8], y[8];
  const bool need_y = (a.a_out != nullptr) || (a.l_coef != 0.f);
  if (VEC) {
    scb_ld8(a.X, a.dtype, row * a.ldX + d, x);
    if (need_y) scb_ld8(a.Y, a.dtype, row * a.ldY + d, y);
  } else {
    x[0] = scb_ld(a.X, a.dtype, row * a.ldX + d);
    y[0] = need_y ? scb_ld(a.Y, a.dtype, row * a.ldY + d) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < W; ++i) g[i] = a.l_coef != 0.f ? a.l_coef * (x[i] - y[i]) : 0.f;
  const int64_t o = row * (int64_t)a.D + d;
  if (a.a_out) {
    const float sd = a.scale_dev ? __ldg(a.scale_dev) : 1.f;
    const float sii = a.scale * sd * __ldg(a.diag + row);
    const float dcoef = expf(sii - __ldg(a.row_lse + row)) + expf(sii - __ldg(a.col_lse_rows + row)) - 2.f;
    float acc[8];
#pragma unroll
    for (int i = 0; i < W; ++i) acc[i] = dcoef * y[i];
    for (int p = 0; p < a.a_jparts; ++p) {
      const float* src = a.a_out + (int64_t)p * a.n * a.D + o;
      if (VEC) {
        const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src)), v1 = __ldcs(reinterpret_cast<const float4*>(src) + 1);
        acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
        acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
      } else {
        acc[0] += __ldcs(src);
      }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) g[i] = fmaf(a.a_coef * sd, acc[i], g[i]);
  }
  if (a.u_out) {
    const float uc = a.u_dev_coef ? a.u_coef * __ldg(a.u_dev_coef) : a.u_coef;
    float r = 0.f;
    for (int q = 0; q < a.rq_parts; ++q) r += __ldg(a.rq + (int64_t)q * a.n + row);
    float acc[8];
#pragma unroll
    for (int i = 0; i < W; ++i) acc[i] = r * x[i];
    for (int p = 0; p < a.u_jparts; ++p) {
      const float* src = a.u_out + (int64_t)p * a.n * a.D + o;
      if (VEC) {
        const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src)), v1 = __ldcs(reinterpret_cast<const float4*>(src) + 1);
        acc[0] -= v0.x; acc[1] -= v0.y; acc[2] -= v0.z; acc[3] -= v0.w;
        acc[4] -= v1.x; acc[5] -= v1.y; acc[6] -= v1.z; acc[7] -= v1.w;
      } else {
        acc[0] -= __ldcs(src);
      }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) g[i] = fmaf(uc, acc[i], g[i]);
  }
  if (a.extra) {
    const float* src = a.extra + o;
    if (VEC) {
      const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src)), v1 = __ldcs(reinterpret_cast<const float4*>(src) + 1);
      g[0] = fmaf(a.e_coef, v0.x, g[0]); g[1] = fmaf(a.e_coef, v0.y, g[1]);
      g[2] = fmaf(a.e_coef, v0.z, g[2]); g[3] = fmaf(a.e_coef, v0.w, g[3]);
      g[4] = fmaf(a.e_coef, v1.x, g[4]); g[5] = fmaf(a.e_coef, v1.y, g[5]);
      g[6] = fmaf(a.e_coef, v1.z, g[6]); g[7] = fmaf(a.e_coef, v1.w, g[7]);
    } else {
      g[0] = fmaf(a.e_coef, __ldcs(src), g[0]);
    }
  }
  if (a.dev_scale) {
    const float gs = __ldg(a.dev_scale);
#pragma unroll
    for (int i = 0; i < W; ++i) g[i] *= gs;
  }
}

template <bool VEC>
__device__ __forceinline__ void combine_store(const CombineArgs& a, int64_t row, int d, const float (&g)[8]) {
  const int64_t oo = row * a.ldOut + d;
  if (VEC) {
    if (a.out_dtype == SCB_F32) {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.dX) + oo);
      dst[0] = make_float4(g[0], g[1], g[2], g[3]);
      dst[1] = make_float4(g[4], g[5], g[6], g[7]);
    } else {
      uint4 pk;
      uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (a.out_dtype == SCB_BF16) {
          __nv_bfloat162 h = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        } else {
          __half2 h = __floats2half2_rn(g[2 * i], g[2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.dX) + oo) = pk;
    }
  } else {
    scb_st(a.dX, a.out_dtype, oo, g[0]);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_grad_combine(const CombineArgs a) {
  constexpr int W = VEC ? 8 : 1;
  const int per_row = (a.D + W - 1) / W;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = idx / per_row;
  if (row >= a.n) return;
  const int d = (int)(idx - row * per_row) * W;
  float g[8];
  combine_terms<VEC>(a, row, d, g);
  combine_store<VEC>(a, row, d, g);
}

// The same pass with the backward of the pre-loss normalise fused in: a row is owned by `tpr` consecutive threads (whole
// warps) of one block, so that e . g is a block-local reduction.  dE = (g - e (e . g)) / ||E||.
template <bool VEC>
__global__ void __launch_bounds__(256) k_grad_combine_unit(const CombineArgs a, int tpr) {
  constexpr int W = VEC ? 8 : 1;
  const int per_row = (a.D + W - 1) / W;
  const int rows_per_block = 256 / tpr;
  const int r_local = (int)threadIdx.x / tpr, t_in_row = (int)threadIdx.x - r_local * tpr;
  const int64_t row = (int64_t)blockIdx.x * rows_per_block + r_local;
  const bool active = r_local < rows_per_block && row < a.n && t_in_row < per_row;
  const int d = t_in_row * W;
  float g[8], e[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g[i] = e[i] = 0.f;
  float inv = 0.f, dot = 0.f;
  if (active) {
    combine_terms<VEC>(a, row, d, g);
    inv = __ldg(a.unit_inv + row);
    if (VEC) scb_ld8(a.unit_src, a.unit_dtype, row * a.ld_unit + d, e);
    else e[0] = scb_ld(a.unit_src, a.unit_dtype, row * a.ld_unit + d);
#pragma unroll
    for (int i = 0; i < W; ++i) {
      e[i] *= inv;
      dot = fmaf(e[i], g[i], dot);
    }
  }
  dot = scb_warp_sum(dot);
  __shared__ float wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (!active) return;
  const int w0 = r_local * (tpr >> 5);
  float tot = 0.f;
  for (int w = 0; w < (tpr >> 5); ++w) tot += wsum[w0 + w];      // fixed order: bit-reproducible
#pragma unroll
  for (int i = 0; i < W; ++i) g[i] = (g[i] - e[i] * tot) * inv;
  combine_store<VEC>(a, row, d, g);
}

}  // namespace

// ============================================================================ C ABI
#define SCB_COMMON_CHECKS(ptr_ok, n, D, dtype)                                           \
  SCB_CHECK_ARG(scb_dtype_ok(dtype), SCB_E_DTYPE, "%s: unsupported dtype %d", __func__, (int)(dtype)); \
  SCB_CHECK_ARG((n) >= 0 && (D) > 0, SCB_E_ARG, "%s: bad shape n=%lld D=%d", __func__, (long long)(n), (int)(D)); \
  SCB_CHECK_ARG((ptr_ok) || (n) == 0, SCB_E_ARG, "%s: null pointer", __func__)

extern "C" int scb_row_sqnorm(const void* X, int64_t n, int D, int64_t ld, int dtype, float* out, void* stream) {
  SCB_COMMON_CHECKS(X && out, n, D, dtype);
  return launch_row_reduce<0>(X, nullptr, n, D, ld, 0, dtype, out, (cudaStream_t)stream);
}

extern "C" int scb_row_dot(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype, float* out,
                           void* stream) {
  SCB_COMMON_CHECKS(A && B && out, n, D, dtype);
  return launch_row_reduce<1>(A, B, n, D, ldA, ldB, dtype, out, (cudaStream_t)stream);
}

extern "C" int scb_lalign_rows(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                               float* row_out, void* stream) {
  SCB_COMMON_CHECKS(X && Y && row_out, n, D, dtype);
  return launch_row_reduce<2>(X, Y, n, D, ldX, ldY, dtype, row_out, (cudaStream_t)stream);
}

extern "C" int scb_lalign_bwd(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                              float host_scale, const float* dev_scale, int accumulate, float* dX, float* dY, void* stream) {
  SCB_COMMON_CHECKS(X && Y && (dX || dY), n, D, dtype);
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(X, ldX, D) && vec_ok(Y, ldY, D))
    k_lalign_bwd<true><<<row_grid(n), kThreads, 0, s>>>(X, Y, n, D, ldX, ldY, dtype, host_scale, dev_scale, accumulate, dX, dY);
  else
    k_lalign_bwd<false><<<row_grid(n), kThreads, 0, s>>>(X, Y, n, D, ldX, ldY, dtype, host_scale, dev_scale, accumulate, dX, dY);
  SCB_CHECK_LAUNCH("lalign_bwd");
  return 0;
}

extern "C" int scb_centroid_fwd(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype, void* C_out,
                                int out_dtype, float* inv_norm, void* stream) {
  SCB_COMMON_CHECKS(A && B && C_out, n, D, dtype);
  SCB_CHECK_ARG(scb_dtype_ok(out_dtype), SCB_E_DTYPE, "centroid_fwd: bad out dtype");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(A, ldA, D) && vec_ok(B, ldB, D))
    k_unit_fwd<true, true><<<row_grid(n), kThreads, 0, s>>>(A, B, n, D, ldA, ldB, dtype, C_out, out_dtype, inv_norm);
  else
    k_unit_fwd<false, true><<<row_grid(n), kThreads, 0, s>>>(A, B, n, D, ldA, ldB, dtype, C_out, out_dtype, inv_norm);
  SCB_CHECK_LAUNCH("centroid_fwd");
  return 0;
}

extern "C" int scb_centroid_bwd(const void* A, const void* B, int64_t n, int D, int64_t ldA, int64_t ldB, int dtype,
                                const float* dC, const float* inv_norm, float host_scale, const float* dev_scale,
                                int accumulate, float* dA, float* dB, void* stream) {
  SCB_COMMON_CHECKS(A && B && dC && inv_norm && (dA || dB), n, D, dtype);
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(A, ldA, D) && vec_ok(B, ldB, D))
    k_unit_bwd<true, true><<<row_grid(n), kThreads, 0, s>>>(A, B, n, D, ldA, ldB, dtype, dC, inv_norm, host_scale, dev_scale, accumulate, dA, dB);
  else
    k_unit_bwd<false, true><<<row_grid(n), kThreads, 0, s>>>(A, B, n, D, ldA, ldB, dtype, dC, inv_norm, host_scale, dev_scale, accumulate, dA, dB);
  SCB_CHECK_LAUNCH("centroid_bwd");
  return 0;
}

extern "C" int scb_normalize_fwd(const void* X, int64_t n, int D, int64_t ld, int dtype, void* Y, int out_dtype, float* inv_norm,
                                 void* stream) {
  SCB_COMMON_CHECKS(X && Y, n, D, dtype);
  SCB_CHECK_ARG(scb_dtype_ok(out_dtype), SCB_E_DTYPE, "normalize_fwd: bad out dtype");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(X, ld, D))
    k_unit_fwd<true, false><<<row_grid(n), kThreads, 0, s>>>(X, nullptr, n, D, ld, 0, dtype, Y, out_dtype, inv_norm);
  else
    k_unit_fwd<false, false><<<row_grid(n), kThreads, 0, s>>>(X, nullptr, n, D, ld, 0, dtype, Y, out_dtype, inv_norm);
  SCB_CHECK_LAUNCH("normalize_fwd");
  return 0;
}

extern "C" int scb_normalize_bwd(const void* X, int64_t n, int D, int64_t ld, int dtype, const float* dY, const float* inv_norm,
                                 float* dX, void* stream) {
  SCB_COMMON_CHECKS(X && dY && inv_norm && dX, n, D, dtype);
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(X, ld, D))
    k_unit_bwd<true, false><<<row_grid(n), kThreads, 0, s>>>(X, nullptr, n, D, ld, 0, dtype, dY, inv_norm, 1.f, nullptr, 0, dX, nullptr);
  else
    k_unit_bwd<false, false><<<row_grid(n), kThreads, 0, s>>>(X, nullptr, n, D, ld, 0, dtype, dY, inv_norm, 1.f, nullptr, 0, dX, nullptr);
  SCB_CHECK_LAUNCH("normalize_bwd");
  return 0;
}

extern "C" int scb_sum(const float* x, int64_t n, float* scratch, float* out, void* stream) {
  SCB_CHECK_ARG(x && scratch && out && n >= 0, SCB_E_ARG, "scb_sum: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int nb = (int)((n + 4095) / 4096);
  if (nb < 1) nb = 1;
  if (nb > kSumBlocks) nb = kSumBlocks;
  k_sum_stage1<<<nb, 256, 0, s>>>(x, n, scratch);
  k_sum_stage2<<<1, 256, 0, s>>>(scratch, nb, out);
  SCB_CHECK_LAUNCH("scb_sum");
  return 0;
}

extern "C" int scb_lse_combine(const float* part_m, const float* part_l, int nparts, int64_t n, float* lse, void* stream) {
  SCB_CHECK_ARG(part_m && part_l && lse && nparts > 0 && n >= 0, SCB_E_ARG, "lse_combine: bad argument");
  if (n == 0) return 0;
  k_lse_combine<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part_m, part_l, nparts, n, lse);
  SCB_CHECK_LAUNCH("lse_combine");
  return 0;
}

extern "C" int scb_anchor_grad_finalize(const float* out, int jparts, int64_t n, int D, const void* V, int64_t ldV, int dtype,
                                        const float* row_lse, const float* col_lse_rows, const float* diag, float scale,
                                        float host_scale, const float* dev_scale, int accumulate, float* dA, const float* scale_dev, void* stream) {
  SCB_COMMON_CHECKS(out && V && row_lse && col_lse_rows && diag && dA, n, D, dtype);
  SCB_CHECK_ARG(jparts > 0, SCB_E_ARG, "anchor_grad_finalize: jparts");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(V, ldV, D))
    k_anchor_fin<true><<<row_grid(n), kThreads, 0, s>>>(out, jparts, n, D, V, ldV, dtype, row_lse, col_lse_rows, diag, scale, host_scale, dev_scale, accumulate, dA, scale_dev);
  else
    k_anchor_fin<false><<<row_grid(n), kThreads, 0, s>>>(out, jparts, n, D, V, ldV, dtype, row_lse, col_lse_rows, diag, scale, host_scale, dev_scale, accumulate, dA, scale_dev);
  SCB_CHECK_LAUNCH("anchor_grad_finalize");
  return 0;
}

extern "C" int scb_lunif_grad_finalize(const float* U, int jparts, const float* rq, int nparts_rq, int64_t n, int D, const void* X,
                                       int64_t ld, int dtype, float host_scale, const float* dev_scale, int accumulate,
                                       float* dX, void* stream) {
  SCB_COMMON_CHECKS(U && rq && X && dX, n, D, dtype);
  SCB_CHECK_ARG(jparts > 0 && nparts_rq > 0, SCB_E_ARG, "lunif_grad_finalize: parts");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok(X, ld, D))
    k_lunif_fin<true><<<row_grid(n), kThreads, 0, s>>>(U, jparts, rq, nparts_rq, n, D, X, ld, dtype, host_scale, dev_scale, accumulate, dX);
  else
    k_lunif_fin<false><<<row_grid(n), kThreads, 0, s>>>(U, jparts, rq, nparts_rq, n, D, X, ld, dtype, host_scale, dev_scale, accumulate, dX);
  SCB_CHECK_LAUNCH("lunif_grad_finalize");
  return 0;
}

extern "C" int scb_grad_combine(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype,
                                const float* a_out, int a_jparts, const float* row_lse, const float* col_lse_rows,
                                const float* diag, float scale, float a_coef, const float* u_out, int u_jparts,
                                const float* rq, int rq_parts, float u_coef, const float* u_dev_coef, float l_coef,
                                const float* extra, float e_coef, const float* dev_scale, void* dX, int out_dtype,
                                int64_t ldOut, const float* scale_dev, const void* unit_src, int64_t ld_unit, int unit_dtype,
                                const float* unit_inv, void* stream) {
  SCB_COMMON_CHECKS(X && dX, n, D, dtype);
  SCB_CHECK_ARG(scb_dtype_ok(out_dtype) && ldOut >= D, SCB_E_ARG, "grad_combine: bad output layout");
  SCB_CHECK_ARG(!a_out || (Y && row_lse && col_lse_rows && diag && a_jparts > 0), SCB_E_ARG, "grad_combine: anchor term");
  SCB_CHECK_ARG(!u_out || (rq && rq_parts > 0 && u_jparts > 0), SCB_E_ARG, "grad_combine: L_unif term");
  SCB_CHECK_ARG(l_coef == 0.f || Y, SCB_E_ARG, "grad_combine: L_align term needs Y");
  if (n == 0) return 0;
  SCB_CHECK_ARG(!unit_src == !unit_inv, SCB_E_ARG, "grad_combine: the fused normalise needs both the rows and 1/norm");
  SCB_CHECK_ARG(!unit_src || (scb_dtype_ok(unit_dtype) && ld_unit >= D), SCB_E_ARG, "grad_combine: bad layout of the un-normalised rows");
  CombineArgs a{X, Y, n, D, ldX, ldY, dtype, a_out, a_jparts, row_lse, col_lse_rows, diag, scale, a_coef, u_out, u_jparts,
                rq, rq_parts, u_coef, u_dev_coef, l_coef, extra, e_coef, dev_scale, dX, out_dtype, ldOut, scale_dev,
                unit_src, ld_unit, unit_dtype, unit_inv};
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = vec_ok(X, ldX, D) && (!Y || vec_ok(Y, ldY, D)) && scb_aligned16(dX) && ldOut % 8 == 0 &&
                   (!extra || scb_aligned16(extra)) && (!unit_src || vec_ok(unit_src, ld_unit, D));
  if (unit_src) {          // a row per `tpr` threads of one block (whole warps): D <= 2048 vectorised, D <= 256 otherwise
    const int per_row_u = vec ? D / 8 : D;
    const int tpr = ((per_row_u + 31) / 32) * 32;
    SCB_CHECK_ARG(tpr <= 256, SCB_E_SHAPE, "grad_combine: the fused normalise needs D <= 2048 (D <= 256 for unaligned rows), D=%d", D);
    const int64_t rows_per_block = 256 / tpr;
    const unsigned grid_u = (unsigned)((n + rows_per_block - 1) / rows_per_block);
    if (vec) k_grad_combine_unit<true><<<grid_u, 256, 0, s>>>(a, tpr);
    else k_grad_combine_unit<false><<<grid_u, 256, 0, s>>>(a, tpr);
    SCB_CHECK_LAUNCH("grad_combine (fused normalise)");
    return 0;
  }
  const int per_row = vec ? D / 8 : D;
  const int64_t total = n * per_row;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (vec) k_grad_combine<true><<<grid, 256, 0, s>>>(a);
  else k_grad_combine<false><<<grid, 256, 0, s>>>(a);
  SCB_CHECK_LAUNCH("grad_combine");
  return 0;
}

extern "C" int scb_lse_combine_cond(const float* part_m, const float* part_l, int nparts, int64_t n, float* lse,
                                    const int* run_flag, void* stream) {
  SCB_CHECK_ARG(part_m && part_l && lse && run_flag && nparts > 0 && n >= 0, SCB_E_ARG, "lse_combine_cond: bad argument");
  if (n == 0) return 0;
  k_lse_combine_cond<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part_m, part_l, nparts, n, lse, run_flag);
  SCB_CHECK_LAUNCH("lse_combine_cond");
  return 0;
}

extern "C" int scb_colstat_combine(const float* col_ref, const float* col_sum, int nparts, int64_t n, float* lse, void* stream) {
  SCB_CHECK_ARG(col_ref && col_sum && lse && nparts > 0 && n >= 0, SCB_E_ARG, "colstat_combine: bad argument");
  if (n == 0) return 0;
  k_colstat_fold<true><<<(unsigned)((n + 127) / 128), 256, 0, (cudaStream_t)stream>>>(col_ref, col_sum, nparts, n, lse, nullptr);
  SCB_CHECK_LAUNCH("colstat_combine");
  return 0;
}

extern "C" int scb_lse2_spread_flag(const float* sqnA, int64_t nA, const float* sqnB, int64_t nB, float scale, int* flag,
                                    const float* scale_dev, void* stream) {
  SCB_CHECK_ARG(sqnA && sqnB && flag && nA >= 0 && nB >= 0 && scale > 0.f, SCB_E_ARG, "lse2_spread_flag: bad argument");
  k_spread_flag<<<1, 1024, 0, (cudaStream_t)stream>>>(sqnA, nA, sqnB, nB, scale, scale_dev, 90.f, flag);
  SCB_CHECK_LAUNCH("lse2_spread_flag");
  return 0;
}

extern "C" int scb_loss_assemble(const float* parts, float c_anchor, float two_scale, float c_align, float w_unif_img,
                                 float w_unif_txt, float w_unif_cen, float pair_norm, float* loss, float* inv_ssum,
                                 const float* scale_dev, void* stream) {
  SCB_CHECK_ARG(parts && loss && inv_ssum, SCB_E_ARG, "loss_assemble: bad argument");
  k_loss_assemble<<<1, 1, 0, (cudaStream_t)stream>>>(parts, c_anchor, two_scale, c_align, w_unif_img, w_unif_txt, w_unif_cen,
                                                     pair_norm, loss, inv_ssum, scale_dev);
  SCB_CHECK_LAUNCH("loss_assemble");
  return 0;
}

extern "C" int scb_lse2_fold_ranks(const float* pack, int world, int64_t stride, int64_t n_loc, int64_t off_exact,
                                   int64_t off_ref, int64_t off_sum, const int* flag, float* col_lse, void* stream) {
  SCB_CHECK_ARG(pack && flag && col_lse && world > 0 && n_loc >= 0, SCB_E_ARG, "lse2_fold_ranks: bad argument");
  const int64_t B = n_loc * world;
  SCB_CHECK_ARG(off_exact >= 0 && off_exact + n_loc <= stride && off_ref >= 0 && off_ref + B <= stride && off_sum >= 0 &&
                    off_sum + B <= stride,
                SCB_E_ARG, "lse2_fold_ranks: offsets outside the packed row");
  if (B == 0) return 0;
  k_fold_ranks<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pack, world, stride, n_loc, off_exact, off_ref,
                                                                              off_sum, flag, col_lse);
  SCB_CHECK_LAUNCH("lse2_fold_ranks");
  return 0;
}

extern "C" int scb_colstat_partial(const float* col_ref, const float* col_sum, int nparts, int64_t n, float* ref_out,
                                   float* sum_out, void* stream) {
  SCB_CHECK_ARG(col_ref && col_sum && ref_out && sum_out && nparts > 0 && n >= 0, SCB_E_ARG, "colstat_partial: bad argument");
  if (n == 0) return 0;
  k_colstat_fold<false><<<(unsigned)((n + 127) / 128), 256, 0, (cudaStream_t)stream>>>(col_ref, col_sum, nparts, n, ref_out, sum_out);
  SCB_CHECK_LAUNCH("colstat_partial");
  return 0;
}
