// tc_pass.cu -- SCB_PATH_TC: the B x B passes on the 5th-gen tensor cores (sm_100a).
//
// One persistent, warp-specialised CTA per SM (384 threads, 1 CTA/SM):
//   warp 0      TMA producer   cp.async.bulk.tensor 2D loads of [128 rows x 64 cols] 16-bit tiles
//                              (128B swizzle) into a ring of 16 KB shared-memory slots
//   warp 1      MMA issuer     one lane issues tcgen05.mma (kind::f16, fp32 accumulate in TMEM)
//   warp 2      TMEM allocator (all 512 columns; 1 CTA per SM by construction)
//   warps 4-11  epilogue       tcgen05.ld S tile -> exp2 / online LSE / masks in registers ->
//                              16-bit weight tile (shared memory, or TMEM for the TS-mode MMA)
//
// A work item is (row block of 128 rows of A, output column group of <=256 columns, column part).
// For every 128-column block j of the sweep:
//   MMA1   S[128x128]   = A_rb . Bm_j^T          (K = D, both operands K-major, A stationary in smem
//                                                  when D <= 512, streamed otherwise)
//   epi    W[128x128]   = f(S)                    (mode-specific, never leaves the SM)
//   MMA2   OUT[128xN2] += W . Bm_j[:, group]      (K = 128; W is the K-major A operand, the SAME
//                                                  Bm tile layout is re-read as an MN-major B operand)
// MMA1 of tile t+1 is issued before MMA2 of tile t, so the tensor pipe works on the next S tile
// while the epilogue warps transform the current one (S is double-buffered in TMEM).
// TMEM columns: OUT [0,256) | S0 [256,384) | S1 [384,512).
//
// Column-tail / row-tail / D-tail handling relies on TMA zero fill plus masks in the epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

enum { M_LSE = 0, M_ANCHOR_GRAD = 1, M_LUNIF_GRAD = 2, M_LUNIF_SUM = 3, M_SPARSIFY_SUM = 4, M_LSE2 = 5, M_RANK_COUNT = 6 };
// M_LSE2: row LSE as M_LSE plus, from the SAME S tiles, per-(row block, column) partial column sums
//   colsum[strip][j] * 2^colref[strip][j/32] = sum_{i in 32-row strip} 2^{y_ij}   (y = scale*log2e * a_i.b_j)
// so that the column LSE needs no second sweep over S^T.  Exact while the spread of y inside a 32 x 32 block stays
// below kColBias (the caller checks a norm bound on the device and falls back to the second sweep otherwise).

constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
// M_LSE2 runs 16 epilogue warps (32 columns each): its epilogue -- a 31-step shuffle butterfly per 32 x 32 block on
// top of the exponentials -- is latency-bound with 8 warps (measured 0.97 ms vs 0.73 ms for the plain LSE sweep)
__host__ __device__ constexpr int epi_warps(int mode) { return mode == 5 ? 16 : 8; }
constexpr int kSlotBytes = 128 * 64 * 2;  // one [128 x 64] 16-bit tile ("chunk")
constexpr int kPairBytes = 2 * kSlotBytes; // ring unit: two chunks under one full / one empty barrier
constexpr int kMaxStages = 7;              // pair-slots
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColOut = 0, kColS0 = 256;
constexpr float kColBias = 100.f;          // M_LSE2: exponent bias of the per-block column sums (log2 units)

struct TcParams {
  int64_t nA, nB;
  int D, kch, nsplit, n_rb, n_jb, jparts;
  int a_stationary, nstage, g_in_tmem, fmt;  // a_stationary: K-chunks of the row block resident in shared memory (kch = all,
                                             // 0 = none, else even).  fmt: 1 bf16, 0 fp16 (both MMA operands must share it:
                                             // a mixed-format kind::f16 descriptor is an illegal instruction)
  float p0;                                  // LSE/anchor: scale*log2e ; lunif: t*log2e
  const float* p0_dev;                       // optional device multiplier of p0 (1/tau of a device-resident temperature)
  const float* rowvec;
  const float* colvec;
  int64_t diag_off;
  float* out;  // [jparts][nA][D]
  float* s0;   // LSE: part_m ; anchor: ws ; lunif grad: rq ; sums: rs      [jparts*2][nA]
  float* s1;   // LSE: part_l ; lunif grad: rs
  float* c0;   // LSE2: column reference   [n_rb*4][ceil(nB/32)]  (one per 32-row strip and 32-column chunk)
  float* c1;   // LSE2: column partial sum [n_rb*4][nB]
  const int* run_flag;   // when non-null the kernel does nothing unless *run_flag != 0 (device-side conditional fallback)
};

// barrier indices inside the barrier block
enum {
  BAR_FULL = 0,                      // [kMaxStages]
  BAR_EMPTY = kMaxStages,            // [kMaxStages]
  BAR_A_FULL = 2 * kMaxStages,
  BAR_A_EMPTY,
  BAR_S_FULL,                        // [2]
  BAR_S_EMPTY = BAR_S_FULL + 2,      // [2]
  BAR_G_FULL = BAR_S_EMPTY + 2,
  BAR_G_EMPTY,
  BAR_OUT_FULL,
  BAR_OUT_EMPTY,
  BAR_COUNT
};

// Ring of `n` pair-slots.  `bits` holds the parity to wait for, per slot.
struct Ring {
  uint32_t slot, bits;
  __device__ __forceinline__ uint32_t take(uint32_t n) {
    const uint32_t s = slot;
    slot = (slot + 1 == n) ? 0u : slot + 1;
    return s;
  }
  __device__ __forceinline__ uint32_t parity_then_flip(uint32_t s) {
    const uint32_t p = (bits >> s) & 1u;
    bits ^= (1u << s);
    return p;
  }
};

// KCH: number of 64-wide K chunks when known at compile time (8 for D = 512: the kc loops unroll and the
// stationary-A descriptors become immediates), 0 = generic.
template <int MODE, int KCH>
__global__ void __launch_bounds__(128 + 32 * epi_warps(MODE), 1)
k_tc_pass(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams P) {
  constexpr bool GRAD = (MODE == M_ANCHOR_GRAD || MODE == M_LUNIF_GRAD);
  constexpr bool NEED_COLVEC = (MODE == M_ANCHOR_GRAD || MODE == M_LUNIF_GRAD || MODE == M_LUNIF_SUM);

  extern __shared__ uint8_t smem_raw[];
  const float p0_eff = P.p0_dev ? P.p0 * __ldg(P.p0_dev) : P.p0;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  // resident K-chunks of the row block: all of them (D <= 512), or the first n_as (even) while the rest streams with the
  // column tiles (D > 512: a 128 x 768 row block is 192 KB), or none
  const int n_as = KCH ? KCH : P.a_stationary;
  const uint32_t a_bytes = (uint32_t)n_as * kSlotBytes;
  const uint32_t sm_a = smem_base;
  const uint32_t sm_ring = sm_a + a_bytes;
  const uint32_t sm_g = sm_ring + (uint32_t)P.nstage * kPairBytes;
  const int kch = KCH ? KCH : P.kch;
  const uint32_t nps = (uint32_t)P.nstage;
  const uint32_t g_bytes = (GRAD && !P.g_in_tmem) ? 2u * kSlotBytes : 0u;
  const uint32_t sm_cbuf = sm_g + g_bytes;            // 2 x 128 floats
  const uint32_t sm_bar = sm_cbuf + 1024u;            // BAR_COUNT x 8 bytes
  const uint32_t sm_tmem_ptr = sm_bar + BAR_COUNT * 8u;
  auto bar = [&](int i) -> uint32_t { return sm_bar + 8u * (uint32_t)i; };
  // generic pointers for the few plain loads/stores
  uint8_t* gen_base = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  float* cbuf = reinterpret_cast<float*>(gen_base + (sm_cbuf - smem_base));
  if (P.run_flag && *P.run_flag == 0) return;         // conditional launch: nothing to do
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sm_tmem_ptr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsplit_eff = GRAD ? P.nsplit : 1;
  const int n_items = P.n_rb * nsplit_eff * P.jparts;
  constexpr int EW = epi_warps(MODE);        // epilogue warps
  constexpr int CW = 512 / EW;               // S-tile columns per epilogue warp (64 or 32)
  const uint32_t s_empty_count = (GRAD && P.g_in_tmem) ? 1u : (uint32_t)EW;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) { ptx::mbar_init(bar(BAR_FULL + i), 1); ptx::mbar_init(bar(BAR_EMPTY + i), 1); }
    ptx::mbar_init(bar(BAR_A_FULL), 1);
    ptx::mbar_init(bar(BAR_A_EMPTY), 1);
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(bar(BAR_S_FULL + b), 1); ptx::mbar_init(bar(BAR_S_EMPTY + b), s_empty_count); }
    ptx::mbar_init(bar(BAR_G_FULL), 8);
    ptx::mbar_init(bar(BAR_G_EMPTY), 1);
    ptx::mbar_init(bar(BAR_OUT_FULL), 1);
    ptx::mbar_init(bar(BAR_OUT_EMPTY), 8);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(sm_tmem_ptr, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  // =========================================================================== TMA producer
  // The whole warp runs the (warp-uniform) control flow; one elected lane issues the async copies.
  if (warp == 0) {
    Ring ring{0u, 0xFFFFFFFFu};
    uint32_t a_empty_par = 1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int split = item % nsplit_eff;
      const int rb = (item / nsplit_eff) % P.n_rb;
      const int jp = item / (nsplit_eff * P.n_rb);
      const int jb_lo = (int)((int64_t)P.n_jb * jp / P.jparts), jb_hi = (int)((int64_t)P.n_jb * (jp + 1) / P.jparts);
      const int nt = jb_hi - jb_lo;
      const int gch = min(4, kch - 4 * split);
      if (n_as > 0) {
        ptx::mbar_wait(bar(BAR_A_EMPTY), a_empty_par, 100);
        a_empty_par ^= 1;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(bar(BAR_A_FULL), (uint32_t)n_as * kSlotBytes);
          for (int kc = 0; kc < n_as; ++kc) ptx::tma_load_2d(sm_a + kc * kSlotBytes, &tmA, kc * 64, rb * 128, bar(BAR_A_FULL));
        }
        __syncwarp();
      }
      // one pair-slot: up to two [128 x 64] chunks (x = first column, consecutive 64-wide chunks) of rows y..y+127
      auto load_pair = [&](const CUtensorMap* tm, int x, int y, int n, int tag) {
        const uint32_t s = ring.take(nps);
        ptx::mbar_wait(bar(BAR_EMPTY + s), ring.parity_then_flip(s), tag);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(bar(BAR_FULL + s), (uint32_t)n * kSlotBytes);
          ptx::tma_load_2d(sm_ring + s * kPairBytes, tm, x, y, bar(BAR_FULL + s));
          if (n > 1) ptx::tma_load_2d(sm_ring + s * kPairBytes + kSlotBytes, tm, x + 64, y, bar(BAR_FULL + s));
        }
        __syncwarp();
      };
      auto load_v = [&](int tt) {   // V tile of column block tt: <=2 pair-slots (128 output columns each)
        for (int c0 = 0; c0 < gch; c0 += 2)
          load_pair(&tmB, (4 * split + c0) * 64, (jb_lo + tt) * 128, min(2, gch - c0), 110);
      };
      for (int t = 0; t < nt; ++t) {
        for (int kc = 0; kc < kch; kc += 2) {
          if (kc >= n_as) load_pair(&tmA, kc * 64, rb * 128, min(2, kch - kc), 120);
          load_pair(&tmB, kc * 64, (jb_lo + t) * 128, min(2, kch - kc), 121);
        }
        if (GRAD && t >= 1) load_v(t - 1);
      }
      if (GRAD && nt > 0) load_v(nt - 1);
    }
  }
  // =========================================================================== MMA issuer
  // Converged warp; elect.sync picks the one lane that issues tcgen05.mma / tcgen05.commit.  Descriptor
  // words are warp-uniform (uniform registers); the loop body is kept minimal because this single warp must
  // issue 8 MMAs (512 tensor-pipe cycles) faster than the pipe retires them.
  else if (warp == 1) {
    Ring ring{0u, 0u};
    uint32_t a_full_par = 0, out_empty_par = 1;
    uint32_t gt1 = 0, gt2 = 0;  // global tile counters of issued MMA1 / MMA2
    const uint32_t idesc1 = ptx::idesc_f16(128, 128, P.fmt, P.fmt, 0, 0);
    const uint32_t a_lo0 = ptx::desc_lo(sm_a, 16);
    const uint32_t ring_lo0 = ptx::desc_lo(sm_ring, 16);
    const uint32_t ring_v_lo0 = ptx::desc_lo(sm_ring, kSlotBytes);   // same address, LBO = one chunk (MN-major V)
    constexpr uint32_t kChunkLo = kSlotBytes >> 4, kPairLo = kPairBytes >> 4;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int split = item % nsplit_eff;
      const int jp = item / (nsplit_eff * P.n_rb);
      const int jb_lo = (int)((int64_t)P.n_jb * jp / P.jparts), jb_hi = (int)((int64_t)P.n_jb * (jp + 1) / P.jparts);
      const int nt = jb_hi - jb_lo;
      const int gch = min(4, kch - 4 * split);
      if (n_as > 0) {
        ptx::mbar_wait(bar(BAR_A_FULL), a_full_par, 200);
        a_full_par ^= 1;
      }
      auto mma1 = [&](bool last) {
        const uint32_t b = gt1 & 1u;
        ptx::mbar_wait(bar(BAR_S_EMPTY + b), ((gt1 >> 1) & 1u) ^ 1u, 210);
        const uint32_t d_tmem = tmem_base + kColS0 + 128u * b;
        auto kgroup = [&](int kc, int nk) {      // 8 MMAs (two 64-wide K chunks) behind one full-barrier wait
          uint32_t alo, sa = 0;
          const bool a_res = kc < n_as;
          if (a_res) {
            alo = a_lo0 + (uint32_t)kc * kChunkLo;
          } else {
            sa = ring.take(nps);
            ptx::mbar_wait(bar(BAR_FULL + sa), ring.parity_then_flip(sa), 211);
            alo = ring_lo0 + sa * kPairLo;
          }
          const uint32_t sb = ring.take(nps);
          ptx::mbar_wait(bar(BAR_FULL + sb), ring.parity_then_flip(sb), 212);
          ptx::tc_fence_after();
          const uint32_t blo = ring_lo0 + sb * kPairLo;
          if (ptx::elect_one()) {
#pragma unroll
            for (uint32_t i = 0; i < 2; ++i) {
              if ((int)i < nk) {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)   // +32 B per K=16 step inside the 128B-swizzled row; +16 KB per chunk
                  ptx::umma_ss(d_tmem, ptx::desc_join(alo + i * kChunkLo + 2u * k), ptx::desc_join(blo + i * kChunkLo + 2u * k),
                               idesc1, (uint32_t)((kc | (int)i | (int)k) != 0));
              }
            }
            if (!a_res) ptx::umma_commit(bar(BAR_EMPTY + sa));
            ptx::umma_commit(bar(BAR_EMPTY + sb));
          }
          __syncwarp();
        };
        if constexpr (KCH > 0) {
#pragma unroll
          for (int kc = 0; kc < KCH; kc += 2) kgroup(kc, 2);
        } else {
#pragma unroll 1
          for (int kc = 0; kc < kch; kc += 2) kgroup(kc, min(2, kch - kc));
        }
        if (ptx::elect_one()) {
          ptx::umma_commit(bar(BAR_S_FULL + b));
          if (last && n_as > 0) ptx::umma_commit(bar(BAR_A_EMPTY));
        }
        __syncwarp();
        ++gt1;
      };
      auto mma2 = [&](bool first, bool last) {
        const uint32_t b = gt2 & 1u;
        ptx::mbar_wait(bar(BAR_G_FULL), gt2 & 1u, 220);
        if (first) {
          ptx::mbar_wait(bar(BAR_OUT_EMPTY), out_empty_par, 221);
          out_empty_par ^= 1;
        }
        const uint32_t g_tmem = tmem_base + kColS0 + 128u * b;
        const uint32_t glo = ptx::desc_lo(sm_g, 16);
#pragma unroll
        for (int c0 = 0; c0 < 4; c0 += 2) {      // output columns [64 c0, 64 c0 + 64 n) of this group
          if (c0 >= gch) break;
          const int n = min(2, gch - c0);
          const uint32_t sv = ring.take(nps);
          ptx::mbar_wait(bar(BAR_FULL + sv), ring.parity_then_flip(sv), 222);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + kColOut + 64u * (uint32_t)c0;
          const uint32_t idesc2 = ptx::idesc_f16(128, 64 * n, P.fmt, P.fmt, 0, 1);
          // V tile re-read as an MN-major operand: +2048 B (16 K-rows) per step, 64-wide blocks one chunk apart
          const uint32_t vlo = ring_v_lo0 + sv * kPairLo;
          const bool last_group = (c0 + 2 >= gch);
          if (ptx::elect_one()) {
#pragma unroll
            for (uint32_t ks = 0; ks < 8; ++ks) {
              const uint64_t bdesc = ptx::desc_join(vlo + 128u * ks);
              const uint32_t accum = (uint32_t)!(first && ks == 0);
              if (P.g_in_tmem)
                ptx::umma_ts(d_tmem, g_tmem + (ks >> 2) * 64u + (ks & 3u) * 8u, bdesc, idesc2, accum);
              else
                ptx::umma_ss(d_tmem, ptx::desc_join(glo + (ks >> 2) * kChunkLo + (ks & 3u) * 2u), bdesc, idesc2, accum);
            }
            ptx::umma_commit(bar(BAR_EMPTY + sv));
            if (last_group) {
              if (P.g_in_tmem) ptx::umma_commit(bar(BAR_S_EMPTY + b));
              else ptx::umma_commit(bar(BAR_G_EMPTY));
              if (last) ptx::umma_commit(bar(BAR_OUT_FULL));
            }
          }
          __syncwarp();
        }
        ++gt2;
      };
      for (int t = 0; t < nt; ++t) {
        mma1(t == nt - 1);
        if (GRAD && t >= 1) mma2(t == 1, false);
      }
      if (GRAD && nt > 0) mma2(nt == 1, true);
    }
  }
  // =========================================================================== epilogue warps
  else if (warp >= 4) {
    const int e = warp - 4;
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int h = e >> 2;         // which CW-column slice of the S tile
    const int rrow = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    uint32_t gt = 0, item_cnt = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_cnt) {
      const int split = item % nsplit_eff;
      const int rb = (item / nsplit_eff) % P.n_rb;
      const int jp = item / (nsplit_eff * P.n_rb);
      const int jb_lo = (int)((int64_t)P.n_jb * jp / P.jparts), jb_hi = (int)((int64_t)P.n_jb * (jp + 1) / P.jparts);
      const int nt = jb_hi - jb_lo;
      const int gch = min(4, kch - 4 * split);
      const int64_t gi = (int64_t)rb * 128 + rrow;
      const bool row_ok = gi < P.nA;
      // per-row constants and running statistics
      float rowc = 0.f;
      if (MODE == M_ANCHOR_GRAD) rowc = row_ok ? P.rowvec[gi] * SCB_LOG2E : 0.f;
      if (MODE == M_LUNIF_GRAD || MODE == M_LUNIF_SUM) rowc = row_ok ? P.rowvec[gi] * p0_eff : 0.f;
      if (MODE == M_RANK_COUNT) rowc = row_ok ? P.rowvec[gi] : INFINITY;      // the score of the row's ground-truth pair
      float st0 = (MODE == M_LSE || MODE == M_LSE2) ? -INFINITY : 0.f, st1 = 0.f;
      const int64_t my_diag_col = gi + P.diag_off;  // column index that is "the diagonal" of this row

      for (int t = 0; t < nt; ++t, ++gt) {
        const uint32_t b = gt & 1u;
        const int jb = jb_lo + t;
        const int64_t col0 = (int64_t)jb * 128 + CW * h;     // first global column handled by this warp
        const bool tile_partial = ((int64_t)jb * 128 + 128) > P.nB;
        if (NEED_COLVEC) {
          const int idx = e * 32 + lane;
          if (idx < 128) {
            const int64_t gj = (int64_t)jb * 128 + idx;
            float cv = INFINITY;
            if (gj < P.nB) cv = (MODE == M_ANCHOR_GRAD) ? P.colvec[gj] * SCB_LOG2E : P.colvec[gj] * p0_eff;
            cbuf[b * 128 + idx] = cv;
          }
          ptx::named_bar_sync(1, kEpiThreads);
        }
        ptx::mbar_wait(bar(BAR_S_FULL + b), (gt >> 1) & 1u, 300);
        ptx::tc_fence_after();
        // does the diagonal cross the 32x64 block this warp handles?
        const int64_t drow0 = (int64_t)rb * 128 + 32 * q + P.diag_off;
        const bool diag_here = (drow0 < col0 + CW) && (drow0 + 32 > col0);

        uint32_t packed[32];
#pragma unroll
        for (int cc = 0; cc < CW / 32; ++cc) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + lane_addr + kColS0 + 128u * b + (uint32_t)(CW * h) + 32u * cc, v);
          ptx::tmem_ld_wait();
          const int64_t cbase = col0 + 32 * cc;
          const float* cb = cbuf + b * 128 + CW * h + 32 * cc;
          const int dcol = diag_here ? (int)(my_diag_col - cbase) : -1;   // local index of the diagonal, if any
          // column tail (last column block only): neutralise the zero-filled columns once, up front
          if (tile_partial && MODE != M_LUNIF_GRAD && MODE != M_LUNIF_SUM) {
            const int nvalid = (int)min((int64_t)32, max((int64_t)0, P.nB - cbase));
            const float dead = (MODE == M_LSE || MODE == M_LSE2) ? -INFINITY : -1e30f;
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c >= nvalid) v[c] = __float_as_uint(dead);
          }
          if (MODE == M_LSE) {
            float cmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) cmax = fmaxf(cmax, __uint_as_float(v[c]));
            if (cmax != -INFINITY) {
              const float mnew = fmaxf(st0, cmax * p0_eff);
              float sum = 0.f;
#pragma unroll
              for (int c = 0; c < 32; ++c) sum += scb_ex2(fmaf(__uint_as_float(v[c]), p0_eff, -mnew));
              st1 = st1 * scb_ex2(st0 - mnew) + sum;
              st0 = mnew;
            }
          } else if (MODE == M_LSE2) {
            // rows: online LSE with a chunk-local exponent reference; columns: the same exponentials, rescaled to
            // the block reference rho = max over the 32 x 32 block, summed over the 32 rows by a shuffle butterfly
            float cmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) cmax = fmaxf(cmax, __uint_as_float(v[c]));
            const float cm = cmax * p0_eff;                       // -inf when every column of the chunk is dead
            float ef[32];
            float sum = 0.f;
            const float cmf = (cmax != -INFINITY) ? cm : 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) { ef[c] = scb_ex2(fmaf(__uint_as_float(v[c]), p0_eff, -cmf)); sum += ef[c]; }
            if (cmax != -INFINITY) {
              const float mnew = fmaxf(st0, cm);
              st1 = st1 * scb_ex2(st0 - mnew) + sum * scb_ex2(cm - mnew);
              st0 = mnew;
            }
            const float rho = scb_warp_max(row_ok ? cm : -INFINITY);
            const float f = (row_ok && cmax != -INFINITY) ? scb_ex2(cm - rho + kColBias) : 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) ef[c] *= f;
            // butterfly: after the five steps lane L holds the sum over the 32 rows of column L of this chunk
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const bool up = (lane & sft) != 0;
#pragma unroll
              for (int k = 0; k < sft; ++k) {
                const float mine = up ? ef[k + sft] : ef[k];
                const float theirs = up ? ef[k] : ef[k + sft];
                ef[k] = mine + __shfl_xor_sync(0xffffffffu, theirs, sft);
              }
            }
            // one partial per (row block, 32-row strip): sums [n_rb*4][nB], references [n_rb*4][ceil(nB/32)]
            const int64_t prow = (int64_t)rb * 4 + q;
            if (cbase + lane < P.nB) P.c1[prow * P.nB + cbase + lane] = ef[0];
            if (lane == 0 && cbase < P.nB) P.c0[prow * ((P.nB + 31) / 32) + (cbase >> 5)] = rho - kColBias;
          } else if (MODE == M_RANK_COUNT) {
            // how many scores of this row beat its ground-truth score (dead columns hold -1e30; the ground-truth
            // column itself never counts, whatever the last bits of its recomputed score)
#pragma unroll
            for (int c = 0; c < 32; ++c) st0 += (__uint_as_float(v[c]) > rowc && c != dcol) ? 1.f : 0.f;
          } else if (MODE == M_SPARSIFY_SUM) {
            if (!tile_partial && !diag_here) {
#pragma unroll
              for (int c = 0; c < 32; ++c) { const float er = __uint_as_float(v[c]) + 1.f; st0 = fmaf(er, er, st0); }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const float g = __uint_as_float(v[c]);
                const float er = g - ((c == dcol) ? 1.f : -1.f);
                st0 += (g > -1e29f) ? er * er : 0.f;
              }
            }
          } else {
            float w[32];
            if (MODE == M_ANCHOR_GRAD) {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const float g = __uint_as_float(v[c]);
                const float y = g * p0_eff;
                const float ww = scb_ex2(y - rowc) + scb_ex2(y - cb[c]);   // dead columns: 0 + 0
                st0 = fmaf(ww, g, st0);
                w[c] = ww;
              }
            } else {  // lunif: exp2(2 p0 g - p0 n_i - p0 n_j); dead columns carry +inf in cb -> 0
              const float two_p0 = 2.f * p0_eff;
#pragma unroll
              for (int c = 0; c < 32; ++c) w[c] = scb_ex2(fmaf(__uint_as_float(v[c]), two_p0, -(rowc + cb[c])));
            }
            if (diag_here) {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (c == dcol) w[c] = 0.f;
            }
            if (MODE == M_LUNIF_SUM) {
#pragma unroll
              for (int c = 0; c < 32; ++c) st1 += w[c];
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const uint32_t pk = P.fmt ? ptx::pack_bf16(w[2 * c], w[2 * c + 1]) : ptx::pack_f16(w[2 * c], w[2 * c + 1]);
                packed[16 * cc + c] = pk;
                if (MODE == M_LUNIF_GRAD) {
                  st1 += w[2 * c] + w[2 * c + 1];
                  if (P.fmt) {
                    st0 += __uint_as_float(pk << 16) + __uint_as_float(pk & 0xffff0000u);
                  } else {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pk));
                    st0 += f.x + f.y;
                  }
                }
              }
            }
          }
        }
        // ---- hand the tile over
        if (!GRAD) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(BAR_S_EMPTY + b));
        } else if (P.g_in_tmem) {
          // weights overwrite the first 32 columns of this warp's own half of the S buffer
          ptx::tmem_st32(tmem_base + lane_addr + kColS0 + 128u * b + 64u * h, packed);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(BAR_G_FULL));
        } else {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(BAR_S_EMPTY + b));
          ptx::mbar_wait(bar(BAR_G_EMPTY), (gt & 1u) ^ 1u, 310);
          const uint32_t row_addr = sm_g + (uint32_t)h * kSlotBytes + (uint32_t)rrow * 128u;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            ptx::st_shared_v4(row_addr + (uint32_t)((u ^ (rrow & 7)) << 4), packed[4 * u], packed[4 * u + 1],
                              packed[4 * u + 2], packed[4 * u + 3]);
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar(BAR_G_FULL));
        }
      }  // tiles

      // ---- drain the output accumulator of this item
      if (GRAD && nt > 0) {
        ptx::mbar_wait(bar(BAR_OUT_FULL), item_cnt & 1u, 320);
        ptx::tc_fence_after();
        const int ncol_half = 32 * gch;  // columns of OUT handled by this warp
        float* orow = P.out + ((int64_t)jp * P.nA + gi) * P.D;
        for (int c0 = 0; c0 < ncol_half; c0 += 32) {
          uint32_t v[32];
          const int ocol = h * ncol_half + c0;
          ptx::tmem_ld32(tmem_base + lane_addr + kColOut + (uint32_t)ocol, v);
          ptx::tmem_ld_wait();
          const int d0 = 256 * split + ocol;
          if (row_ok) {
            if (d0 + 32 <= P.D) {
#pragma unroll
              for (int c = 0; c < 32; c += 4)
                *reinterpret_cast<float4*>(orow + d0 + c) = make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                                                        __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (d0 + c < P.D) orow[d0 + c] = __uint_as_float(v[c]);
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(BAR_OUT_EMPTY));
      }
      if (row_ok && split == 0) {
        const int64_t o = ((int64_t)jp * (EW / 4) + h) * P.nA + gi;
        if (MODE == M_LSE || MODE == M_LSE2) { P.s0[o] = st0; P.s1[o] = st1; }
        if (MODE == M_ANCHOR_GRAD && P.s0) P.s0[o] = st0;
        if (MODE == M_LUNIF_GRAD) { P.s0[o] = st0; P.s1[o] = st1; }
        if (MODE == M_LUNIF_SUM) P.s1[o] = st1;
        if (MODE == M_SPARSIFY_SUM || MODE == M_RANK_COUNT) P.s0[o] = st0;
      }
    }  // items
  }

  // =========================================================================== teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int D, int64_t ld, int dtype, int box_rows = 128) {
  PFN_encodeTiled enc = get_encode();
  SCB_CHECK_ARG(enc != nullptr, SCB_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, dtype == SCB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCB_CHECK_ARG(r == CUDA_SUCCESS, SCB_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

std::atomic<int> g_tc_flags{31};   // bit0: weight tile in TMEM (TS-mode MMA2) -- measured faster than the smem variant
                      // bit1: CTA-pair kernel (tc_pair.cu) for the gradient passes when 256 < D <= 512
                      // bit2: cluster-of-4 kernel (tc_quad.cu, cta_group::2 MMAs) for the same passes when nA > 128
                      // bit3: with bit2, the last rows of a large pass run on CTA pairs on the SMs clusters of 4 strand
                      // bit4: with bit2, 512 < D <= 768 runs the single-S-buffer variant (all 384 output columns of a pair
                      //       in TMEM, nothing recomputed) instead of two column-group launches
                      // bit5: (tests) the cluster-of-4 kernel's row-block-aligned span plan for any operand size

// Measured (profiles/r03q_lse_residency.log, one-sweep LSE at 32768^2 x 768 / 8192 x 65536 x 768 / 16384^2 x 1024):
// 0 resident 919 / 1009 / 824 TFLOP/s, 4: 915 / 1004 / 824, 6: 910 / 983 / 809, 8: 829 / 880 / 676 -- what these sweeps
// pull from L2 is NOT what bounds them (the single-CTA SS MMA reads all of an SM's shared-memory bandwidth); the deeper
// ring wins.  Default: none resident.
#ifndef SCB_PASS_ASTAT_BIG
#define SCB_PASS_ASTAT_BIG 0
#endif
// resident row-block chunks of the non-gradient sweeps when D > 512 (even; SCB_PASS_ASTAT in the environment overrides
// the compiled default -- a tuning knob, read once)
int pass_astat_big() {
  static const int v = [] {
    const char* e = getenv("SCB_PASS_ASTAT");
    int n = e ? atoi(e) : SCB_PASS_ASTAT_BIG;
    if (n < 0) n = 0;
    if (n > 8) n = 8;
    return n & ~1;
  }();
  return v;
}

template <int MODE>
int launch_tc(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
              TcParams P, cudaStream_t s) {
  constexpr bool GRAD = (MODE == M_ANCHOR_GRAD || MODE == M_LUNIF_GRAD);
  SCB_CHECK_ARG(dtype == SCB_BF16 || dtype == SCB_F16, SCB_E_DTYPE, "tensor-core path needs bf16/fp16 operands");
  SCB_CHECK_ARG(D % 8 == 0 && ldA % 8 == 0 && ldB % 8 == 0, SCB_E_SHAPE, "tensor-core path needs D and ld multiples of 8");
  SCB_CHECK_ARG(scb_aligned16(A) && scb_aligned16(Bm), SCB_E_ARG, "tensor-core path needs 16-byte aligned operands");
  if (nA == 0) return 0;
  SCB_CHECK_ARG(nB > 0, SCB_E_SHAPE, "empty column side");
  P.nA = nA; P.nB = nB; P.D = D;
  P.kch = (D + 63) / 64;
  P.nsplit = (P.kch + 3) / 4;
  P.n_rb = (int)((nA + 127) / 128);
  P.n_jb = (int)((nB + 127) / 128);
  SCB_CHECK_ARG(P.jparts >= 1 && P.jparts <= P.n_jb, SCB_E_ARG, "jparts=%d outside [1, %d]", P.jparts, P.n_jb);
  P.fmt = (dtype == SCB_BF16) ? 1 : 0;
  P.g_in_tmem = (GRAD && (g_tc_flags & 1)) ? 1 : 0;
  // Resident K-chunks of the row block.  D <= 512: all.  D > 512 (a 128 x 768 row block is 192 KB): the first
  // pass_astat_big() may stay while the rest streams with every column tile (fewer bytes from L2, shallower ring) --
  // measured not to pay, see SCB_PASS_ASTAT_BIG.  The gradient modes (only D > 1024 still runs them here) stream everything.
  P.a_stationary = (P.kch <= 8) ? P.kch : (GRAD ? 0 : pass_astat_big());
  const int budget = 232448 - 1024 /*align slack*/ - 1024 /*cbuf*/ - 1024 /*barriers*/ ;
  const int fixed = P.a_stationary * kSlotBytes + ((GRAD && !P.g_in_tmem) ? 2 * kSlotBytes : 0);
  int nstage = (budget - fixed) / kPairBytes;      // pair-slots of 32 KB
  if (nstage > kMaxStages) nstage = kMaxStages;
  SCB_CHECK_ARG(nstage >= 2, SCB_E_SHAPE, "not enough shared memory for the tile ring (D=%d)", D);
  P.nstage = nstage;
  const size_t smem = (size_t)fixed + (size_t)nstage * kPairBytes + 3 * 1024;

  CUtensorMap tmA, tmB;
  int rc = make_tmap(&tmA, A, nA, D, ldA, dtype);
  if (rc) return rc;
  rc = make_tmap(&tmB, Bm, nB, D, ldB, dtype);
  if (rc) return rc;

  const int num_sms = scb_num_sms();
  static std::atomic<unsigned long long> attr_done{0};
  {
    const cudaError_t e = scb_opt_in_smem(attr_done, 232448, k_tc_pass<MODE, 0>, k_tc_pass<MODE, 8>);
    if (e != cudaSuccess) { scb_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
  }
  const int n_items = P.n_rb * (GRAD ? P.nsplit : 1) * P.jparts;
  const int grid = n_items < num_sms ? n_items : num_sms;
  constexpr int threads = 128 + 32 * epi_warps(MODE);
  if (P.kch == 8) k_tc_pass<MODE, 8><<<grid, threads, smem, s>>>(tmA, tmB, P);
  else k_tc_pass<MODE, 0><<<grid, threads, smem, s>>>(tmA, tmB, P);
  SCB_CHECK_LAUNCH("tc_pass");
  return 0;
}

}  // namespace

int scb_tc_pair_set_dbg(int);
int scb_quad_clusters();
int scb_tc_set_flags(int flags) { const int o = g_tc_flags.exchange(flags & 63); scb_tc_pair_set_dbg(flags >> 6); return o; }
int scb_tc_flags_get() { return g_tc_flags.load(); }
int scb_make_tmap_2d(CUtensorMap* m, const void* base, int64_t rows, int D, int64_t ld, int dtype) {
  return make_tmap(m, base, rows, D, ld, dtype);
}
int scb_make_tmap_2d_box(CUtensorMap* m, const void* base, int64_t rows, int D, int64_t ld, int dtype, int box_rows) {
  return make_tmap(m, base, rows, D, ld, dtype, box_rows);
}
// Which kernel runs a gradient pass on the tensor-core path: 0 = single CTA (k_tc_pass; D <= 256 needs no split, D > 512
// sweeps once per 256-column output group), 1 = CTA pair (k_tc_pair), 2 = cluster of 4 with cta_group::2 MMAs (k_tc_quad).
// The planner (scb_pass_plan) and the launchers both ask here, so they cannot disagree.
int scb_tc_grad_kernel(int64_t nA, int D, int grad) {
  const int kch = (D + 63) / 64;
  const int f = g_tc_flags.load();
  if (!grad || kch <= 4 || kch > 16) return 0;
  if ((f & 4) && nA > 128 && scb_quad_clusters() > 0) return 2;     // clusters of 4: 256 < D <= 1024 (column groups)
  return ((f & 2) && kch <= 8) ? 1 : 0;                              // CTA pairs: 256 < D <= 512
}
int scb_tc_pair_anchor_grad(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*,
                            const float*, int64_t, int, float*, float*, const float*, cudaStream_t);
int scb_tc_pair_lunif(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*, const float*,
                      int64_t, int, float*, float*, float*, cudaStream_t);
int scb_tc_quad_anchor_grad(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*,
                            const float*, int64_t, int, float*, float*, const float*, cudaStream_t);
int scb_tc_quad_lunif(const void*, int64_t, const void*, int64_t, int, int64_t, int64_t, int, float, const float*, const float*,
                      int64_t, int, float*, float*, float*, cudaStream_t);

int scb_tc_lse(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype, float scale,
               int jparts, float* pm, float* pl, const int* run_flag, const float* scale_dev, cudaStream_t s) {
  TcParams P{};
  P.p0_dev = scale_dev;
  P.jparts = jparts; P.p0 = scale * SCB_LOG2E; P.diag_off = INT64_MIN / 2; P.s0 = pm; P.s1 = pl; P.run_flag = run_flag;
  return launch_tc<M_LSE>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s);
}
// rows AND columns from one sweep: col_sum [4 n_rb][nB], col_ref [4 n_rb][ceil(nB/32)], n_rb = ceil(nA / 128)
int scb_tc_lse2(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype, float scale,
                int jparts, float* pm, float* pl, float* col_ref, float* col_sum, const float* scale_dev, cudaStream_t s) {
  TcParams P{};
  P.p0_dev = scale_dev;
  P.jparts = jparts; P.p0 = scale * SCB_LOG2E; P.diag_off = INT64_MIN / 2; P.s0 = pm; P.s1 = pl; P.c0 = col_ref; P.c1 = col_sum;
  return launch_tc<M_LSE2>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s);
}
int scb_tc_anchor_grad(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                       float scale, const float* row_lse, const float* col_lse, int64_t diag_off, int jparts, float* out,
                       float* ws, const float* scale_dev, cudaStream_t s) {
  const int kern = scb_tc_grad_kernel(nA, D, 1);
  if (kern && (dtype == SCB_BF16 || dtype == SCB_F16) && D % 8 == 0 && ldA % 8 == 0 && ldB % 8 == 0 && scb_aligned16(A) &&
      scb_aligned16(Bm))
    return kern == 2
               ? scb_tc_quad_anchor_grad(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, row_lse, col_lse, diag_off, jparts, out, ws, scale_dev, s)
               : scb_tc_pair_anchor_grad(A, nA, Bm, nB, D, ldA, ldB, dtype, scale, row_lse, col_lse, diag_off, jparts, out, ws, scale_dev, s);
  TcParams P{};
  P.p0_dev = scale_dev;
  P.jparts = jparts; P.p0 = scale * SCB_LOG2E; P.rowvec = row_lse; P.colvec = col_lse; P.diag_off = diag_off;
  P.out = out; P.s0 = ws;
  return launch_tc<M_ANCHOR_GRAD>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s);
}
int scb_tc_lunif(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll, int dtype,
                 float t, const float* sqn_r, const float* sqn_all, int64_t row_offset, int jparts, float* U, float* rq,
                 float* rs, cudaStream_t s) {
  const int kern = U ? scb_tc_grad_kernel(nR, D, 1) : 0;
  if (kern && (dtype == SCB_BF16 || dtype == SCB_F16) && D % 8 == 0 && ldR % 8 == 0 && ldAll % 8 == 0 && scb_aligned16(Xr) &&
      scb_aligned16(Xall))
    return kern == 2
               ? scb_tc_quad_lunif(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, t, sqn_r, sqn_all, row_offset, jparts, U, rq, rs, s)
               : scb_tc_pair_lunif(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, t, sqn_r, sqn_all, row_offset, jparts, U, rq, rs, s);
  TcParams P{};
  P.jparts = jparts; P.p0 = t * SCB_LOG2E; P.rowvec = sqn_r; P.colvec = sqn_all; P.diag_off = row_offset;
  P.out = U; P.s0 = rq; P.s1 = rs;
  return U ? launch_tc<M_LUNIF_GRAD>(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, P, s)
           : launch_tc<M_LUNIF_SUM>(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, P, s);
}
int scb_tc_rank_count(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                      const float* gt_score, int64_t diag_off, int jparts, float* cnt, cudaStream_t s) {
  TcParams P{};
  P.jparts = jparts; P.rowvec = gt_score; P.diag_off = diag_off; P.s0 = cnt;
  return launch_tc<M_RANK_COUNT>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s);
}
int scb_tc_sparsify_sum(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll,
                        int dtype, int64_t row_offset, int jparts, float* rs, cudaStream_t s) {
  TcParams P{};
  P.jparts = jparts; P.diag_off = row_offset; P.s0 = rs;
  return launch_tc<M_SPARSIFY_SUM>(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, P, s);
}
