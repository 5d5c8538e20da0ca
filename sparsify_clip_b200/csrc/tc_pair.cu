// tc_pair.cu -- gradient passes on a CTA PAIR (thread-block cluster of 2, sm_100a) for 256 < D <= 512.
//
// Why a pair: the row-stationary output accumulator of a 128-row block is 128 x D fp32 = all 512 TMEM columns at
// D = 512, which leaves no room for the S tile.  k_tc_pass therefore sweeps the columns once per 256-column output
// group and recomputes S each time (3 contractions per sweep instead of 2).  Here the two CTAs of a cluster own
// the SAME 128 rows and ONE HALF of the output columns each; the S tiles are computed once, alternately by the
// two CTAs, and the 16-bit weight tile W = f(S) is handed to the peer through distributed shared memory:
//
//   tile t owned by CTA (t & 1):   MMA1  S = A_rb . Bm_t^T          (K = D, A stationary in smem)
//                                  epi   W = f(S)  -> own TMEM (in place, TS-mode operand)
//                                                  -> peer smem  (st.shared::cluster, 32 KB, K-major SW128)
//   every tile, both CTAs:         MMA2  OUT[:, half] += W . Bm_t[:, half]   (own W from TMEM, peer's W from smem)
//
// Per pair and per two tiles the tensor pipes execute 2 x (2048 + 1024 + 1024) cycles = exactly the algorithmic
// 2 contractions.  DSMEM traffic is 32 KB per CTA per two tiles (~8 B/clk of the measured ~20 B/clk).
//
// Roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue,
// warps 12-15 W senders (read the finished weight tile back from TMEM and push it to the peer with st.async, so that
// the DSMEM store back-pressure -- ~3000 cycles per 32 KB tile, measured -- never blocks the epilogue warps).
// Registers are rebalanced with setmaxnreg: 80 (control) / 168 (epilogue) / 96 (senders) = 64 K.
// MMA issue order per CTA (lag 2, so that the epilogue of tile t hides behind three other MMAs):
//   step t:  [MMA1(t) if own]  then  [MMA2(t-2)]
// TMEM columns: OUT [0,256) | S0 [256,384) | S1 [384,512)   (W overwrites the first 32 columns of each 64-col half)
// Shared memory: A kch x 16 KB | Wrecv 32 KB | ring nslots x 16 KB | column vector 1 KB | barriers
#include "common.cuh"
#include "ptx.cuh"

namespace {

enum { M_ANCHOR_GRAD = 1, M_LUNIF_GRAD = 2 };

constexpr int kThreads = 512;
constexpr int kEpiThreads = 256;
constexpr int kSlotBytes = 128 * 64 * 2;   // one [128 x 64] 16-bit chunk
constexpr int kMaxSlots = 8;
#ifndef SCB_PAIR_ASTAT
#define SCB_PAIR_ASTAT 6                   // tuning experiments may override it (-DSCB_PAIR_ASTAT=4: ring of 8 slots)
#endif
constexpr int kAStat = SCB_PAIR_ASTAT;     // K-chunks of A resident in smem; the rest stream with the column tile
                                           // (frees ring slots: the ring, not the tensor pipe, was the bottleneck)
constexpr int kSendPaceClk = 200;          // idle cycles between two 16-byte remote stores of a sender thread
constexpr int kPeerLag = 3;                // MMA2 of a peer tile is issued this many steps after the tile (odd:
                                           // it lands on my own steps); hides epilogue + DSMEM latency of the peer
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColOut = 0, kColS0 = 256;

struct PairParams {
  int64_t nA, nB;
  int64_t slot_rows;           // rows of one output partial slot (= nA unless this launch covers a row range of a larger pass)
  int D, kch, n_rb, n_jb, jparts, nslots, fmt;
  int64_t span;                // tiles of the linearised (row block, column tile) space per CTA pair
  float p0;
  const float* p0_dev;         // optional device multiplier of p0 (1/tau of a device-resident temperature)
  const float* rowvec;
  const float* colvec;
  int64_t diag_off;
  float* out;  // [jparts][nA][D]
  float* s0;   // anchor: ws ; lunif: rq     [jparts*4][nA]
  float* s1;   // lunif: rs
  unsigned long long* trace;   // debug timeline (scb_debug_pair_trace), normally null
  int dbg;                     // sender pace override in units of 50 cycles (bits 3..), 0 = kSendPaceClk.  Builds with
                               // -DSCB_PAIR_EXPERIMENTS also honour bit0 = send 1/16 of the W bytes (WRONG results,
                               // timing experiments only); the shipped library ignores it.
};
#ifdef SCB_PAIR_EXPERIMENTS
#define SCB_PAIR_SHORT_W(P) ((P).dbg & 1)
#else
#define SCB_PAIR_SHORT_W(P) 0
#endif

// Debug timeline (builds with -DSCB_PAIR_TRACE only; the shipped library compiles it out and keeps no mutable state):
// cluster 0 records (tag, tile, clock64) per role; kTraceCap events per (CTA rank, role).
constexpr int kTraceCap = 4096;
#ifdef SCB_PAIR_TRACE
struct Tracer {
  unsigned long long* base;
  uint32_t n;
  __device__ __forceinline__ void init(unsigned long long* trace, int pair_id, uint32_t crank, int role) {
    base = (trace && pair_id == 0) ? trace + ((size_t)(crank * 4 + role) * kTraceCap) * 2 : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void rec(uint32_t tag, uint32_t tile) {
    if (base && n < kTraceCap) {
      base[2 * n] = ((unsigned long long)tag << 32) | tile;
      base[2 * n + 1] = (unsigned long long)clock64();
      ++n;
    }
  }
};
#else
struct Tracer {
  __device__ __forceinline__ void init(unsigned long long*, int, uint32_t, int) {}
  __device__ __forceinline__ void rec(uint32_t, uint32_t) {}
};
#endif

enum {
  BAR_FULL = 0,                      // [kMaxSlots]
  BAR_EMPTY = kMaxSlots,             // [kMaxSlots]
  BAR_A_FULL = 2 * kMaxSlots,
  BAR_A_EMPTY,
  BAR_S_FULL,                        // [2]
  BAR_S_EMPTY = BAR_S_FULL + 2,      // [2]
  BAR_G_FULL = BAR_S_EMPTY + 2,      // [2]  own W tile stored in TMEM (8 epilogue warps)
  BAR_OUT_FULL = BAR_G_FULL + 2,
  BAR_OUT_EMPTY,
  BAR_W_FULL,                        // peer's W tile landed in my Wrecv (32 KB of st.async complete_tx)
  BAR_W_EMPTY,                       // peer consumed the W tile I sent (peer's tcgen05.commit, multicast)
  BAR_COUNT
};

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

// ---- cluster / DSMEM helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte asynchronous store into the peer's shared memory; completes 16 tx-bytes on the peer's mbarrier when the
// data has landed, so the sender needs no fence and no arrive (a release.cluster arrive behind generic
// st.shared::cluster stores costs MEMBAR.ALL.GPU + ERRBAR per warp: measured 50% of the epilogue time).
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                            uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar) : "memory");
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = ptx::globaltimer_ns();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (ptx::globaltimer_ns() - t0 > SCB_TC_WATCHDOG_NS) ptx::watchdog_fire(tag, parity);
  }
}
// arrive::one on the barrier at the same CTA-relative offset in every CTA of `mask`, once all tcgen05 operations
// issued so far by this thread have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

template <int MODE, int KCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_tc_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PairParams P) {
  extern __shared__ uint8_t smem_raw[];
  const float p0_eff = P.p0_dev ? P.p0 * __ldg(P.p0_dev) : P.p0;
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // same offset in both CTAs
  const int kch = KCH ? KCH : P.kch;
  const int n_astat = KCH ? (KCH < kAStat ? KCH : kAStat) : min(kch, kAStat);
  const int b1_slots = 2 * kch - n_astat;                                   // ring slots one S tile consumes
  const uint32_t sm_a = smem_base;
  const uint32_t sm_w = sm_a + (uint32_t)n_astat * kSlotBytes;              // Wrecv: 2 chunks
  const uint32_t sm_ring = sm_w + 2u * kSlotBytes;
  const uint32_t nslots = (uint32_t)P.nslots;
  const uint32_t sm_cbuf = sm_ring + nslots * kSlotBytes;                   // 2 x 128 floats
  const uint32_t sm_bar = sm_cbuf + 1024u;
  const uint32_t sm_tmem_ptr = sm_bar + BAR_COUNT * 8u;
  auto bar = [&](int i) -> uint32_t { return sm_bar + 8u * (uint32_t)i; };
  uint8_t* gen_base = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  float* cbuf = reinterpret_cast<float*>(gen_base + (sm_cbuf - smem_base));
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(gen_base + (sm_tmem_ptr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0 / 1: which half of the output columns this CTA owns
  const uint32_t peer = crank ^ 1u;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t total_tiles = (int64_t)P.n_rb * P.n_jb;
  const int gch = min(4, kch - 4 * (int)crank);      // 64-wide output chunks of my half
  const int gch_even = (gch + 1) & ~1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMaxSlots; ++i) { ptx::mbar_init(bar(BAR_FULL + i), 1); ptx::mbar_init(bar(BAR_EMPTY + i), 1); }
    ptx::mbar_init(bar(BAR_A_FULL), 1);
    ptx::mbar_init(bar(BAR_A_EMPTY), 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(bar(BAR_S_FULL + b), 1);
      ptx::mbar_init(bar(BAR_S_EMPTY + b), 5);   // MMA2 commit + the 4 sender warps (W read back from TMEM)
      ptx::mbar_init(bar(BAR_G_FULL + b), 8);
    }
    ptx::mbar_init(bar(BAR_OUT_FULL), 1);
    ptx::mbar_init(bar(BAR_OUT_EMPTY), 8);
    ptx::mbar_init(bar(BAR_W_FULL), 1);      // armed by my MMA warp with expect_tx(32 KB); bytes come from the peer
    ptx::mbar_init(bar(BAR_W_EMPTY), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(sm_tmem_ptr, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer's barriers are initialised before anything remote touches them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  // Work distribution: the (row block, column tile) space is linearised row-block-major and cut into equal contiguous
  // spans, one per CTA pair ("stream-K" over the column sweep): perfect balance for any shard size, and one pipeline
  // fill/drain per row block a pair touches instead of one per (row block, part) item.  A span decomposes into
  // segments (row block, tile range); a segment's output partial index `jp` is its ordinal inside its row block.
  // `item_cnt` alternates which CTA of the pair owns the even tiles.
#define SCB_PAIR_FOR_SEGMENTS()                                                                               \
  for (int64_t g = (int64_t)pair_id * P.span, g_end = (g + P.span < total_tiles ? g + P.span : total_tiles); g < g_end; ++item_cnt)
#define SCB_PAIR_ITEM_SETUP()                                                                              \
  const int rb = (int)(g / P.n_jb);                                                                        \
  const int jb_lo = (int)(g - (int64_t)rb * P.n_jb);                                                       \
  const int nt = (int)(((int64_t)(P.n_jb - jb_lo) < g_end - g) ? (int64_t)(P.n_jb - jb_lo) : g_end - g);   \
  const int jp = pair_id - (int)(((int64_t)rb * P.n_jb) / P.span);                                         \
  const int t_first = (int)((crank ^ (item_cnt & 1u)) & 1u);  /* my own tiles: t_first, t_first + 2, ... */ \
  const int n_own = (nt > t_first) ? (nt - t_first + 1) / 2 : 0;                                           \
  g += nt;                                                                                                 \
  (void)rb; (void)jp; (void)jb_lo; (void)n_own

  // (setmaxnreg sits at the top of each role branch so that the role's code is dominated by it: ptxas only
  //  budgets registers per region when that holds)
  // =========================================================================== TMA producer
  if (warp == 0) {
    setmaxnreg_dec<80>();
    Tracer tr; tr.init(P.trace, pair_id, crank, 0);
#ifdef SCB_PAIR_TRACE
    if (lane == 0) { tr.rec(0, (uint32_t)(ptx::globaltimer_ns() & 0xffffffffu)); tr.rec(0, (uint32_t)(ptx::globaltimer_ns() >> 32)); }
#endif
    Ring ring{0u, 0xFFFFFFFFu};
    uint32_t a_empty_par = 1, item_cnt = 0;
    SCB_PAIR_FOR_SEGMENTS() {
      SCB_PAIR_ITEM_SETUP();
      if (n_own > 0) {
        ptx::mbar_wait(bar(BAR_A_EMPTY), a_empty_par, 100);
        a_empty_par ^= 1;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(bar(BAR_A_FULL), (uint32_t)n_astat * kSlotBytes);
          for (int kc = 0; kc < n_astat; ++kc) ptx::tma_load_2d(sm_a + kc * kSlotBytes, &tmA, kc * 64, rb * 128, bar(BAR_A_FULL));
        }
        __syncwarp();
      }
      // one chunk slot; real == false burns the slot (keeps V pairs on even slots)
      auto load_chunk = [&](bool real, const CUtensorMap* tm, int x, int y, int tag) {
        const uint32_t s = ring.take(nslots);
        ptx::mbar_wait(bar(BAR_EMPTY + s), ring.parity_then_flip(s), tag);
        if (lane == 0) tr.rec((uint32_t)tag, (uint32_t)y);
        if (ptx::elect_one()) {
          if (real) {
            ptx::mbar_expect_tx(bar(BAR_FULL + s), (uint32_t)kSlotBytes);
            ptx::tma_load_2d(sm_ring + s * kSlotBytes, tm, x, y, bar(BAR_FULL + s));
          } else {
            ptx::mbar_arrive(bar(BAR_FULL + s));
          }
        }
        __syncwarp();
      };
      auto load_v = [&](int tt, int tag) {
        for (int c0 = 0; c0 < gch_even; ++c0) load_chunk(c0 < gch, &tmB, (4 * (int)crank + c0) * 64, (jb_lo + tt) * 128, tag);
      };
      // my steps are the steps of my own tiles: [MMA1(t)] [MMA2 own (t-2)] [MMA2 peer (t-kPeerLag)]
      for (int t = t_first; t < nt + kPeerLag + 1; t += 2) {
        if (t < nt) {
          for (int kc = 0; kc < kch; ++kc) {
            if (kc >= n_astat) load_chunk(true, &tmA, kc * 64, rb * 128, 119);
            load_chunk(true, &tmB, kc * 64, (jb_lo + t) * 128, 120);
          }
          if (b1_slots & 1) load_chunk(false, &tmB, 0, 0, 118);
        }
        if (t - 2 >= 0 && t - 2 < nt) load_v(t - 2, 121);
        if (t - kPeerLag >= 0 && t - kPeerLag < nt) load_v(t - kPeerLag, 122);
      }
    }
  }
  // =========================================================================== MMA issuer
  else if (warp == 1) {
    setmaxnreg_dec<80>();
    Tracer tr; tr.init(P.trace, pair_id, crank, 1);
    Ring ring{0u, 0u};
    uint32_t a_full_par = 0, out_empty_par = 1, item_cnt = 0;
    uint32_t k1 = 0, k2 = 0, kp = 0;   // issued MMA1 (own tiles), MMA2 on own tiles, MMA2 on peer tiles
    const uint32_t idesc1 = ptx::idesc_f16(128, 128, P.fmt, P.fmt, 0, 0);
    const uint32_t a_lo0 = ptx::desc_lo(sm_a, 16);
    const uint32_t w_lo0 = ptx::desc_lo(sm_w, 16);
    const uint32_t ring_lo0 = ptx::desc_lo(sm_ring, 16);
    const uint32_t ring_v_lo0 = ptx::desc_lo(sm_ring, kSlotBytes);   // MN-major V: 64-wide blocks one chunk apart
    constexpr uint32_t kChunkLo = kSlotBytes >> 4;
    SCB_PAIR_FOR_SEGMENTS() {
      SCB_PAIR_ITEM_SETUP();
      if (n_own > 0) {
        ptx::mbar_wait(bar(BAR_A_FULL), a_full_par, 200);
        a_full_par ^= 1;
      }
      int own_left = n_own;
      auto mma1 = [&]() {
        const uint32_t b = k1 & 1u;
        if (lane == 0) tr.rec(10, k1);
        ptx::mbar_wait(bar(BAR_S_EMPTY + b), ((k1 >> 1) & 1u) ^ 1u, 210);
        if (lane == 0) tr.rec(11, k1);
        const uint32_t d_tmem = tmem_base + kColS0 + 128u * b;
        auto kchunk = [&](int kc) {
          uint32_t alo, sa = 0;
          const bool a_streamed = kc >= n_astat;
          if (a_streamed) {
            sa = ring.take(nslots);
            ptx::mbar_wait(bar(BAR_FULL + sa), ring.parity_then_flip(sa), 211);
            alo = ring_lo0 + sa * kChunkLo;
          } else {
            alo = a_lo0 + (uint32_t)kc * kChunkLo;
          }
          const uint32_t s = ring.take(nslots);
          ptx::mbar_wait(bar(BAR_FULL + s), ring.parity_then_flip(s), 212);
          if (lane == 0) tr.rec(12, (uint32_t)kc);
          ptx::tc_fence_after();
          const uint32_t blo = ring_lo0 + s * kChunkLo;
          if (ptx::elect_one()) {
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
              ptx::umma_ss(d_tmem, ptx::desc_join(alo + 2u * k), ptx::desc_join(blo + 2u * k), idesc1,
                           (uint32_t)((kc | (int)k) != 0));
            if (a_streamed) ptx::umma_commit(bar(BAR_EMPTY + sa));
            ptx::umma_commit(bar(BAR_EMPTY + s));
          }
          __syncwarp();
        };
        if constexpr (KCH > 0) {
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc) kchunk(kc);
        } else {
#pragma unroll 1
          for (int kc = 0; kc < kch; ++kc) kchunk(kc);
        }
        if (b1_slots & 1) {     // padding slot (keeps the V pairs on even slots)
          const uint32_t s = ring.take(nslots);
          ptx::mbar_wait(bar(BAR_FULL + s), ring.parity_then_flip(s), 213);
          if (ptx::elect_one()) ptx::umma_commit(bar(BAR_EMPTY + s));
          __syncwarp();
        }
        --own_left;
        if (ptx::elect_one()) {
          ptx::umma_commit(bar(BAR_S_FULL + b));
          if (own_left == 0) ptx::umma_commit(bar(BAR_A_EMPTY));
        }
        __syncwarp();
        ++k1;
      };
      auto mma2 = [&](bool own, bool first, bool last) {
        uint32_t b = 0;
        if (lane == 0) tr.rec(own ? 20 : 30, own ? k2 : kp);
        if (own) {
          b = k2 & 1u;
          ptx::mbar_wait(bar(BAR_G_FULL + b), (k2 >> 1) & 1u, 220);
        } else {
          if (ptx::elect_one()) ptx::mbar_expect_tx(bar(BAR_W_FULL), SCB_PAIR_SHORT_W(P) ? 2048u : 2u * kSlotBytes);
          __syncwarp();
          ptx::mbar_wait(bar(BAR_W_FULL), kp & 1u, 225);
          ptx::fence_proxy_async_smem();
        }
        if (lane == 0) tr.rec(own ? 21 : 31, own ? k2 : kp);
        if (first) {
          ptx::mbar_wait(bar(BAR_OUT_EMPTY), out_empty_par, 221);
          out_empty_par ^= 1;
        }
        const uint32_t g_tmem = tmem_base + kColS0 + 128u * b;
#pragma unroll
        for (int c0 = 0; c0 < 4; c0 += 2) {      // output columns [64 c0, 64 c0 + 64 n) of my half
          if (c0 >= gch_even) break;
          const int n = min(2, gch - c0);
          const uint32_t sv = ring.take(nslots);
          ptx::mbar_wait(bar(BAR_FULL + sv), ring.parity_then_flip(sv), 222);
          const uint32_t sv1 = ring.take(nslots);
          ptx::mbar_wait(bar(BAR_FULL + sv1), ring.parity_then_flip(sv1), 223);
          if (lane == 0) tr.rec(own ? 22 : 32, (uint32_t)c0);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + kColOut + 64u * (uint32_t)c0;
          const uint32_t idesc2 = ptx::idesc_f16(128, 64 * n, P.fmt, P.fmt, 0, 1);
          const uint32_t vlo = ring_v_lo0 + sv * kChunkLo;
          const bool last_group = (c0 + 2 >= gch_even);
          if (ptx::elect_one()) {
#pragma unroll
            for (uint32_t ks = 0; ks < 8; ++ks) {
              const uint64_t bdesc = ptx::desc_join(vlo + 128u * ks);
              const uint32_t accum = (uint32_t)!(first && ks == 0);
              if (own)
                ptx::umma_ts(d_tmem, g_tmem + (ks >> 2) * 64u + (ks & 3u) * 8u, bdesc, idesc2, accum);
              else
                ptx::umma_ss(d_tmem, ptx::desc_join(w_lo0 + (ks >> 2) * kChunkLo + (ks & 3u) * 2u), bdesc, idesc2, accum);
            }
            ptx::umma_commit(bar(BAR_EMPTY + sv));
            ptx::umma_commit(bar(BAR_EMPTY + sv1));
            if (last_group) {
              if (own) ptx::umma_commit(bar(BAR_S_EMPTY + b));
              else umma_commit_mc(bar(BAR_W_EMPTY), (uint16_t)(1u << peer));
              if (last) ptx::umma_commit(bar(BAR_OUT_FULL));
            }
          }
          __syncwarp();
          if (lane == 0) tr.rec(own ? 23 : 33, (uint32_t)c0);
        }
        if (own) ++k2; else ++kp;
      };
      int m2_left = nt;          // the first MMA2 of the item overwrites OUT, the last one publishes it
      for (int t = t_first; t < nt + kPeerLag + 1; t += 2) {
        if (t < nt) mma1();
        if (t - 2 >= 0 && t - 2 < nt) { mma2(true, m2_left == nt, m2_left == 1); --m2_left; }
        if (t - kPeerLag >= 0 && t - kPeerLag < nt) { mma2(false, m2_left == nt, m2_left == 1); --m2_left; }
      }
    }
  }
  // =========================================================================== epilogue warps
  else if (warp < 4) {
    setmaxnreg_dec<80>();      // TMEM allocator + spare warp: nothing to do until teardown
  }
  else if (warp < 12) {
    setmaxnreg_inc<168>();
    const int e = warp - 4;
    const int q = warp & 3;       // TMEM lane quarter this warp may access
    const int h = e >> 2;         // which 64-column half of the S tile
    const int rrow = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    uint32_t ke = 0, item_cnt = 0;
    Tracer tr; tr.init(e == 0 ? P.trace : nullptr, pair_id, crank, 2);
    SCB_PAIR_FOR_SEGMENTS() {
      SCB_PAIR_ITEM_SETUP();
      const int64_t gi = (int64_t)rb * 128 + rrow;
      const bool row_ok = gi < P.nA;
      float rowc = 0.f;
      if (MODE == M_ANCHOR_GRAD) rowc = row_ok ? P.rowvec[gi] * SCB_LOG2E : 0.f;
      if (MODE == M_LUNIF_GRAD) rowc = row_ok ? P.rowvec[gi] * p0_eff : 0.f;
      float st0 = 0.f, st1 = 0.f;
      const int64_t my_diag_col = gi + P.diag_off;

      for (int t = t_first; t < nt; t += 2, ++ke) {
        const uint32_t b = ke & 1u;
        const int jb = jb_lo + t;
        const int64_t col0 = (int64_t)jb * 128 + 64 * h;
        const bool tile_partial = ((int64_t)jb * 128 + 128) > P.nB;
        {
          const int idx = e * 32 + lane;
          if (idx < 128) {
            const int64_t gj = (int64_t)jb * 128 + idx;
            float cv = INFINITY;
            if (gj < P.nB) cv = (MODE == M_ANCHOR_GRAD) ? P.colvec[gj] * SCB_LOG2E : P.colvec[gj] * p0_eff;
            cbuf[b * 128 + idx] = cv;
          }
          ptx::named_bar_sync(1, kEpiThreads);
        }
        if (lane == 0) tr.rec(40, ke);
        ptx::mbar_wait(bar(BAR_S_FULL + b), (ke >> 1) & 1u, 300);
        if (lane == 0) tr.rec(41, ke);
        ptx::tc_fence_after();
        const int64_t drow0 = (int64_t)rb * 128 + 32 * q + P.diag_off;
        const bool diag_here = (drow0 < col0 + 64) && (drow0 + 32 > col0);

        uint32_t packed[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + lane_addr + kColS0 + 128u * b + 64u * h + 32u * cc, v);
          ptx::tmem_ld_wait();
          const int64_t cbase = col0 + 32 * cc;
          const float* cb = cbuf + b * 128 + 64 * h + 32 * cc;
          const int dcol = diag_here ? (int)(my_diag_col - cbase) : -1;
          if (tile_partial && MODE == M_ANCHOR_GRAD) {
            const int nvalid = (int)min((int64_t)32, max((int64_t)0, P.nB - cbase));
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c >= nvalid) v[c] = __float_as_uint(-1e30f);
          }
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
        // ---- own copy: weights overwrite the first 32 columns of this warp's half of the S buffer (TS-mode MMA2)
        ptx::tmem_st32(tmem_base + lane_addr + kColS0 + 128u * b + 64u * h, packed);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(bar(BAR_G_FULL + b)); tr.rec(42, ke); }
      }  // own tiles

      // ---- drain my half of the output accumulator
      if (nt > 0) {
        ptx::mbar_wait(bar(BAR_OUT_FULL), item_cnt & 1u, 320);
        ptx::tc_fence_after();
        const int ncol_half = 32 * gch;  // columns of OUT handled by this warp
        float* orow = P.out + ((int64_t)jp * P.slot_rows + gi) * P.D;
        for (int c0 = 0; c0 < ncol_half; c0 += 32) {
          uint32_t v[32];
          const int ocol = h * ncol_half + c0;
          ptx::tmem_ld32(tmem_base + lane_addr + kColOut + (uint32_t)ocol, v);
          ptx::tmem_ld_wait();
          const int d0 = 256 * (int)crank + ocol;
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
      if (row_ok) {   // statistics over MY tiles only: 4 sub-partials per part (CTA rank x column half)
        const int64_t o = ((int64_t)jp * 4 + 2 * (int)crank + h) * P.slot_rows + gi;
        if (MODE == M_ANCHOR_GRAD && P.s0) P.s0[o] = st0;
        if (MODE == M_LUNIF_GRAD) { P.s0[o] = st0; P.s1[o] = st1; }
      }
      // A row block is covered by 1 .. jparts segments; the consumers sum all `jparts` partial slots, so the segment
      // that finishes the row block clears the slots nobody writes.
      if (row_ok && jb_lo + nt == P.n_jb) {
        const int ncol_half = 32 * gch;
        for (int sl = jp + 1; sl < P.jparts; ++sl) {
          float* orow = P.out + ((int64_t)sl * P.slot_rows + gi) * P.D + 256 * (int)crank + h * ncol_half;
          for (int c = 0; c < ncol_half; c += 4) {
            if (256 * (int)crank + h * ncol_half + c + 4 <= P.D) *reinterpret_cast<float4*>(orow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            else for (int cc2 = 0; cc2 < 4; ++cc2) if (256 * (int)crank + h * ncol_half + c + cc2 < P.D) orow[c + cc2] = 0.f;
          }
          const int64_t o = ((int64_t)sl * 4 + 2 * (int)crank + h) * P.slot_rows + gi;
          if (MODE == M_ANCHOR_GRAD && P.s0) P.s0[o] = 0.f;
          if (MODE == M_LUNIF_GRAD) { P.s0[o] = 0.f; P.s1[o] = 0.f; }
        }
      }
    }  // items
  }
  // =========================================================================== W senders
  else {
    setmaxnreg_dec<96>();
    const int q = warp & 3;
    const int rrow = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    const uint32_t peer_w = mapa(sm_w, peer);
    const uint32_t peer_w_full = mapa(bar(BAR_W_FULL), peer);
    uint32_t kx = 0, item_cnt = 0;
    Tracer tr; tr.init(q == 0 ? P.trace : nullptr, pair_id, crank, 3);
    SCB_PAIR_FOR_SEGMENTS() {
      SCB_PAIR_ITEM_SETUP();
      for (int t = t_first; t < nt; t += 2, ++kx) {
        const uint32_t b = kx & 1u;
        ptx::mbar_wait(bar(BAR_G_FULL + b), (kx >> 1) & 1u, 400);
        if (lane == 0) tr.rec(50, kx);
        ptx::tc_fence_after();
        uint32_t w0[32], w1[32];     // the two 64-column halves of my 32 rows of W (packed 16-bit pairs)
        ptx::tmem_ld32(tmem_base + lane_addr + kColS0 + 128u * b, w0);
        ptx::tmem_ld32(tmem_base + lane_addr + kColS0 + 128u * b + 64u, w1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar(BAR_S_EMPTY + b));
        // K-major SW128 image in the peer's Wrecv, once the peer has consumed the previous tile
        ptx::mbar_wait(bar(BAR_W_EMPTY), (kx & 1u) ^ 1u, 410);
        if (lane == 0) tr.rec(51, kx);
        const uint32_t row_addr = peer_w + (uint32_t)rrow * 128u;
        // Pacing: a back-to-back burst of 64 remote stores per warp delays the TMA traffic of both SMs (measured:
        // ~3000 cycles per tile pair); spreading the 32 KB over ~3000 cycles costs nothing (the peer consumes the
        // tile kPeerLag steps later) and was the best of {0, 100, 200, 300, 400+} cycles per store.
        const int pace = (P.dbg >> 3) ? (P.dbg >> 3) * 50 : kSendPaceClk;
        if (SCB_PAIR_SHORT_W(P)) {     // timing experiment: same handshake, 1/16 of the bytes
          st_async_v4(row_addr, w0[0], w0[1], w0[2], w0[3], peer_w_full);
          continue;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          st_async_v4(row_addr + (uint32_t)((u ^ (rrow & 7)) << 4), w0[4 * u], w0[4 * u + 1], w0[4 * u + 2], w0[4 * u + 3],
                      peer_w_full);
          if (pace) { const long long c0 = clock64(); while (clock64() - c0 < pace) {} }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          st_async_v4(row_addr + kSlotBytes + (uint32_t)((u ^ (rrow & 7)) << 4), w1[4 * u], w1[4 * u + 1], w1[4 * u + 2],
                      w1[4 * u + 3], peer_w_full);
          if (pace) { const long long c0 = clock64(); while (clock64() - c0 < pace) {} }
        }
        if (lane == 0) tr.rec(52, kx);
      }
    }
  }
#undef SCB_PAIR_ITEM_SETUP
#undef SCB_PAIR_FOR_SEGMENTS

  // =========================================================================== teardown
  // nobody leaves while the peer may still write into this CTA's shared memory or signal its barriers
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int scb_make_tmap_2d(CUtensorMap* m, const void* base, int64_t rows, int D, int64_t ld, int dtype);   // tc_pass.cu

namespace {

#ifdef SCB_PAIR_TRACE
std::atomic<unsigned long long*> g_pair_trace{nullptr};
#endif
#ifdef SCB_PAIR_EXPERIMENTS
std::atomic<int> g_pair_dbg{0};
#endif

}  // namespace

// Span plan shared with the planner (scb_pass_plan): pairs used, tiles per pair, and the largest number of segments
// any row block is cut into (= the number of output partial slots the caller must provide, "jparts").
void scb_pair_span_plan(int64_t n_rb, int64_t n_jb, int n_sm, int* n_pairs, int64_t* span, int* pmax) {
  const int64_t total = n_rb * n_jb;
  int64_t np = n_sm / 2;
  if (np > total) np = total;
  if (np < 1) np = 1;
  int64_t sp = (total + np - 1) / np;
  if (sp < (n_jb + 14) / 15) sp = (n_jb + 14) / 15;      // at most 16 partial slots per row block
  if (sp < 1) sp = 1;
  np = (total + sp - 1) / sp;
  int mx = 1;
  for (int64_t rb = 0; rb < n_rb; ++rb) {
    const int64_t first = (rb * n_jb) / sp, last = ((rb + 1) * n_jb - 1) / sp;
    if ((int)(last - first + 1) > mx) mx = (int)(last - first + 1);
  }
  *n_pairs = (int)(np < 1 ? 1 : np);
  *span = sp;
  *pmax = mx;
}

namespace {

template <int MODE>
int launch_pair(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                PairParams P, cudaStream_t s, int max_pairs = 0) {
  if (nA == 0) return 0;
  P.nA = nA; P.nB = nB; P.D = D;
  if (P.slot_rows == 0) P.slot_rows = nA;
#ifdef SCB_PAIR_TRACE
  P.trace = g_pair_trace.load();
#endif
#ifdef SCB_PAIR_EXPERIMENTS
  P.dbg = g_pair_dbg.load();
#endif
  P.kch = (D + 63) / 64;
  P.n_rb = (int)((nA + 127) / 128);
  P.n_jb = (int)((nB + 127) / 128);
  SCB_CHECK_ARG(P.kch > 4 && P.kch <= 8, SCB_E_SHAPE, "pair kernel needs 256 < D <= 512 (D=%d)", D);

  P.fmt = (dtype == SCB_BF16) ? 1 : 0;
  const int budget = 232448 - 1024 /*align slack*/ - 1024 /*cbuf*/ - 1024 /*barriers*/;
  const int n_astat = P.kch < kAStat ? P.kch : kAStat;
  int nslots = (budget - (n_astat + 2) * kSlotBytes) / kSlotBytes;
  nslots &= ~1;
  if (nslots > kMaxSlots) nslots = kMaxSlots;
  SCB_CHECK_ARG(nslots >= 4, SCB_E_SHAPE, "not enough shared memory for the chunk ring (D=%d)", D);
  P.nslots = nslots;
  const size_t smem = (size_t)(n_astat + 2 + nslots) * kSlotBytes + 3 * 1024;

  CUtensorMap tmA, tmB;
  int rc = scb_make_tmap_2d(&tmA, A, nA, D, ldA, dtype);
  if (rc) return rc;
  rc = scb_make_tmap_2d(&tmB, Bm, nB, D, ldB, dtype);
  if (rc) return rc;

  const int num_sms = scb_num_sms();
  static std::atomic<unsigned long long> attr_done{0};
  {
    const cudaError_t e = scb_opt_in_smem(attr_done, 232448, k_tc_pair<MODE, 0>, k_tc_pair<MODE, 8>);
    if (e != cudaSuccess) { scb_set_error("cudaFuncSetAttribute(pair): %s", cudaGetErrorString(e)); return (int)e; }
  }
  int n_pairs = 1, pmax = 1;
  scb_pair_span_plan(P.n_rb, P.n_jb, max_pairs > 0 ? 2 * max_pairs : num_sms, &n_pairs, &P.span, &pmax);
  SCB_CHECK_ARG(P.jparts >= pmax, SCB_E_ARG, "pair kernel: jparts=%d but the span plan needs %d partial slots (scb_pass_plan)",
                P.jparts, pmax);
  if (P.kch == 8) k_tc_pair<MODE, 8><<<2 * n_pairs, kThreads, smem, s>>>(tmA, tmB, P);
  else k_tc_pair<MODE, 0><<<2 * n_pairs, kThreads, smem, s>>>(tmA, tmB, P);
  SCB_CHECK_LAUNCH("tc_pair");
  return 0;
}

}  // namespace

int scb_tc_pair_anchor_grad(const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                            float scale, const float* row_lse, const float* col_lse, int64_t diag_off, int jparts,
                            float* out, float* ws, const float* scale_dev, cudaStream_t s) {
  PairParams P{};
  P.p0_dev = scale_dev;
  P.jparts = jparts; P.p0 = scale * SCB_LOG2E; P.rowvec = row_lse; P.colvec = col_lse; P.diag_off = diag_off;
  P.out = out; P.s0 = ws;
  return launch_pair<M_ANCHOR_GRAD>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s);
}
int scb_tc_pair_lunif(const void* Xr, int64_t nR, const void* Xall, int64_t nAll, int D, int64_t ldR, int64_t ldAll, int dtype,
                      float t, const float* sqn_r, const float* sqn_all, int64_t row_offset, int jparts, float* U,
                      float* rq, float* rs, cudaStream_t s) {
  PairParams P{};
  P.jparts = jparts; P.p0 = t * SCB_LOG2E; P.rowvec = sqn_r; P.colvec = sqn_all; P.diag_off = row_offset;
  P.out = U; P.s0 = rq; P.s1 = rs;
  return launch_pair<M_LUNIF_GRAD>(Xr, nR, Xall, nAll, D, ldR, ldAll, dtype, P, s);
}

// A row range of a larger pass on at most `max_pairs` CTA pairs (tc_quad.cu runs it on the SMs that clusters of 4
// cannot use): `row_base` rows precede A's first row in the pass, `slot_rows` = rows of the whole pass.
int scb_tc_pair_range(int mode, const void* A, int64_t nA, const void* Bm, int64_t nB, int D, int64_t ldA, int64_t ldB, int dtype,
                      float p0, const float* p0_dev, const float* rowvec, const float* colvec, int64_t diag_off, int jparts,
                      int64_t slot_rows, float* out, float* s0, float* s1, int max_pairs, cudaStream_t s) {
  PairParams P{};
  P.p0_dev = p0_dev;
  P.jparts = jparts; P.p0 = p0; P.rowvec = rowvec; P.colvec = colvec; P.diag_off = diag_off; P.slot_rows = slot_rows;
  P.out = out; P.s0 = s0; P.s1 = s1;
  return mode == M_ANCHOR_GRAD ? launch_pair<M_ANCHOR_GRAD>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s, max_pairs)
                               : launch_pair<M_LUNIF_GRAD>(A, nA, Bm, nB, D, ldA, ldB, dtype, P, s, max_pairs);
}

unsigned long long* scb_pair_trace_buffer() {
#ifdef SCB_PAIR_TRACE
  return g_pair_trace.load();
#else
  return nullptr;
#endif
}
// debug: timeline buffer for the next pair launches (2 CTAs x 4 roles x 4096 events x 2 words of 8 bytes), or null.
// Only builds with -DSCB_PAIR_TRACE record anything; the shipped library rejects the call.
extern "C" int scb_debug_pair_trace(void* buf) {
#ifdef SCB_PAIR_TRACE
  g_pair_trace.store(static_cast<unsigned long long*>(buf));
  return 0;
#else
  SCB_CHECK_ARG(buf == nullptr, SCB_E_ARG, "scb_debug_pair_trace: this build has no tracer (-DSCB_PAIR_TRACE)");
  return 0;
#endif
}
int scb_tc_pair_set_dbg(int v) {
#ifdef SCB_PAIR_EXPERIMENTS
  g_pair_dbg.store(v);
#else
  (void)v;
#endif
  return 0;
}
