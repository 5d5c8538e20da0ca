// Does incoming DSMEM traffic slow TMA loads into the same SM?  Every CTA streams 16 KB chunks from an L2-resident
// matrix (as tools/tma_stream_bench.cu, unicast, consumer delay 300 clk) while 4 "sender" warps of each CTA push a
// 32 KB tile to the peer CTA of a 2-cluster every `period` clocks, by
//   mode 0: nothing (reference)        mode 1: st.async.v4 (16-byte remote stores, complete_tx)
//   mode 2: cp.async.bulk shared::cta -> shared::cluster (2 x 16 KB from a local staging buffer)
//   mode 3: st.global.v4 of the tile + __threadfence (no DSMEM; the L2 route's writer side)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) { while (!try_wait(bar, par)) {} }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* m, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
constexpr int SLOT = 16384, NS = 6, TILE = 32768;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(224, 1)
k(const __grid_constant__ CUtensorMap tm, int ntiles, int delay, int period, int nsend, uint8_t* gscratch, long long* cycles, int* sent) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
  const uint32_t land = base + NS * SLOT;          // 32 KB landing zone (written by the peer)
  const uint32_t stage = land + TILE;              // 32 KB local staging (mode 2)
  const uint32_t bars = stage + TILE;              // full[NS], empty[NS], land_full
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (NS + i), 1); }
    mbar_init(bars + 8 * (2 * NS), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) {
      uint32_t slot = 0, par = 1;
      for (int t = 0; t < ntiles; ++t)
        for (int kc = 0; kc < 8; ++kc) {
          wait(bars + 8 * (NS + slot), par);
          mbar_expect_tx(bars + 8 * slot, SLOT);
          tma_load(base + slot * SLOT, &tm, kc * 64, t * 128, bars + 8 * slot);
          if (++slot == NS) { slot = 0; par ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t slot = 0, par = 0;
      for (int t = 0; t < ntiles; ++t)
        for (int kc = 0; kc < 8; ++kc) {
          wait(bars + 8 * slot, par);
          const long long c0 = clock64(); while (clock64() - c0 < delay) {}
          mbar_arrive(bars + 8 * (NS + slot));
          if (++slot == NS) { slot = 0; par ^= 1; }
        }
      cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (warp == 2) {   // receiver of the peer's tiles
    if (lane == 0 && (MODE == 1 || MODE == 2))
      for (int i = 0; i < nsend; ++i) {
        mbar_expect_tx(bars + 8 * (2 * NS), TILE);
        wait(bars + 8 * (2 * NS), i & 1);
      }
  } else {   // 4 sender warps
    const int st = threadIdx.x - 96;     // 0..127
    const uint32_t rland = mapa(land, rank ^ 1u), rbar = mapa(bars + 8 * (2 * NS), rank ^ 1u);
    int n = 0;
    uint32_t v0 = st, v1 = 1, v2 = 2, v3 = 3;
    long long next = clock64() + period;
    for (int it = 0; it < nsend && MODE != 0; ++it) {
      while (clock64() < next) {}
      next += period;
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t o = (uint32_t)(j * 128 + st) * 16u;
          asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(rland + o), "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(rbar) : "memory");
        }
      } else if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t o = (uint32_t)(j * 128 + st) * 16u;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stage + o), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (st == 0) {
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rland), "r"(stage), "r"(TILE), "r"(rbar) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      } else {
        uint4* g = reinterpret_cast<uint4*>(gscratch + (size_t)blockIdx.x * TILE);
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j * 128 + st] = make_uint4(v0, v1, v2, v3);
        __threadfence();
      }
      ++n; v0 += 3;
    }
    if (st == 0) sent[blockIdx.x] = n;
  }
  __syncthreads();
  cluster_sync();
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int MODE>
void run(const CUtensorMap& tm, int delay, int period, uint8_t* gs, long long* d, int* sent) {
  const int ntiles = 128, grid = 148;
  const size_t smem = NS * SLOT + 2 * TILE + 2048;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nsend = (int)((double)ntiles * 8 * (delay + 140) / period);
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<grid, 224, smem>>>(tm, ntiles, delay, period, nsend, gs, d, sent);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); exit(1); }
  }
  long long h[160]; int hs[160];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost); cudaMemcpy(hs, sent, grid * sizeof(int), cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("mode %d  delay %3d  send period %5d : %5.0f clk per chunk (%.1f B/clk/SM TMA), tiles sent per CTA %d (%.1f B/clk DSMEM/L2)\n", MODE, delay, period,
         mx / (ntiles * 8), (double)ntiles * 8 * SLOT / mx, hs[0], (double)hs[0] * TILE / mx);
}
int main() {
  const int64_t rows = 32768; const int D = 512;
  void* X; cudaMalloc(&X, rows * D * 2); cudaMemset(X, 1, rows * D * 2);
  long long* d; cudaMalloc(&d, 160 * sizeof(long long)); int* sent; cudaMalloc(&sent, 160 * sizeof(int));
  uint8_t* gs; cudaMalloc(&gs, 160 * TILE);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((PFN_enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  for (int delay : {200, 300}) {
    run<0>(tm, delay, 6000, gs, d, sent);
    for (int period : {6000, 3000}) {
      run<1>(tm, delay, period, gs, d, sent);
      run<2>(tm, delay, period, gs, d, sent);
      run<3>(tm, delay, period, gs, d, sent);
    }
  }
  return 0;
}
