// DSMEM handoff probe 2: a 32 KB tile is pushed to the peer CTA by NT sender threads, with the completion signal
// the consumer would need, then the peer acknowledges (remote arrive) before the next tile.  Modes:
//   0  st.shared::cluster.v4 + fence.proxy.async.shared::cluster + mbarrier.arrive.release.cluster (per warp)
//   1  st.async (complete_tx on the peer's barrier), no fence
//   2  st.shared::cluster.v4 + mbarrier.arrive.release.cluster only (consumer-side proxy fence)
// Prints cycles per tile for one-way and two-way traffic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
constexpr int TILE = 32768;
template <int MODE, int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT + 32, 1) k(int iters, int senders, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* dst = sm;  // landing buffer
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TILE);   // [0] full (data landed here), [1] empty (peer consumed what I sent)
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1u;
  const int nw = NT / 32;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(MODE == 1 ? 1 : nw));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const bool send = (int)rank < senders, recv = (int)peer < senders;
  const uint32_t rdst = mapa(smem_u32(dst), peer), rfull = mapa(smem_u32(&bars[0]), peer), rempty = mapa(smem_u32(&bars[1]), peer);
  const uint32_t lfull = smem_u32(&bars[0]), lempty = smem_u32(&bars[1]);
  long long t0 = clock64();
  if (threadIdx.x < NT) {           // sender threads
    if (send) {
      uint32_t v0 = threadIdx.x, v1 = 2, v2 = 3, v3 = 4;
      for (int it = 0; it < iters; ++it) {
        while (!try_wait(lempty, (it & 1) ^ 1)) {}
#pragma unroll
        for (int j = 0; j < TILE / 16 / NT; ++j) {
          const uint32_t o = (uint32_t)(j * NT + threadIdx.x) * 16u;
          if (MODE == 1)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(rdst + o), "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(rfull) : "memory");
          else
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rdst + o), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
        }
        if (MODE != 1) {
          if (MODE == 0) asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
          __syncwarp();
          if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rfull) : "memory");
        }
        v0 += 7;
      }
    }
  } else if (threadIdx.x == NT) {    // consumer thread
    if (recv) {
      for (int it = 0; it < iters; ++it) {
        if (MODE == 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lfull), "r"(TILE) : "memory");
        while (!try_wait(lfull, it & 1)) {}
        if (MODE == 2) asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rempty) : "memory");
      }
    }
  }
  __syncthreads();
  cluster_sync();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (dst[threadIdx.x] == 0x5a && iters < 0) printf("x");
}
template <int MODE, int NT>
void run(long long* d) {
  long long h[296];
  const int smem = TILE + 64;
  cudaFuncSetAttribute(k<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int senders = 1; senders <= 2; ++senders) {
    const int grid = 148, iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      k<MODE, NT><<<grid, NT + 32, smem>>>(iters, senders, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mode %d  sender threads %3d  senders/pair=%d : %.0f cycles per 32 KB tile incl. handshake (%.1f B/clk)\n", MODE, NT, senders, mx / iters, (double)iters * TILE / mx);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 296 * sizeof(long long));
  run<0, 128>(d); run<0, 256>(d); run<1, 128>(d); run<1, 256>(d); run<2, 128>(d); run<2, 256>(d);
  return 0;
}
