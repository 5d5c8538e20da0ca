// DSMEM bandwidth probe for the CTA-pair design: cluster of 2, each 32 KB "tile" is pushed into the peer's
// shared memory either with st.shared::cluster.v4 from 256 threads or with cp.async.bulk smem->peer smem.
// Prints bytes/clk per SM for one-way and two-way traffic.   nvcc -gencode arch=compute_100a,code=sm_100a
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
constexpr int TILE = 32768;

// mode 0: st.shared::cluster.v4 ; mode 1: cp.async.bulk shared::cta -> shared::cluster
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k(int iters, int senders, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* src = sm;            // 32 KB local source
  uint8_t* dst = sm + TILE;     // 32 KB landing buffer (written by the peer)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * TILE);
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t peer = rank ^ 1u;
  for (int i = threadIdx.x; i < TILE / 4; i += 256) reinterpret_cast<uint32_t*>(src)[i] = i * 2654435761u + rank;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  const bool send = (int)rank < senders;
  const uint32_t rdst = mapa(smem_u32(dst), peer), rbar = mapa(smem_u32(bar), peer);
  long long t0 = clock64();
  if (MODE == 0) {
    if (send) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < TILE / 16 / 256; ++j) {
          const int o = (j * 256 + threadIdx.x) * 16;
          uint4 v = *reinterpret_cast<const uint4*>(src + o);
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rdst + o), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
      }
    }
  } else {
    // receiver arms its own barrier for all incoming bytes of one tile, sender issues one bulk copy per tile
    uint32_t parity = 0;
    for (int it = 0; it < iters; ++it) {
      const bool receive = (int)peer < senders;
      if (threadIdx.x == 0 && receive)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(TILE) : "memory");
      cluster_sync();   // (coarse flow control: fine for a bandwidth probe with large tiles)
      if (threadIdx.x == 0 && send)
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(rdst), "r"(smem_u32(src)), "r"(TILE), "r"(rbar) : "memory");
      if (receive) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      }
      parity ^= 1;
    }
  }
  __syncthreads();
  cluster_sync();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (dst[threadIdx.x] == 0x5a && iters < 0) printf("x");
}

int main() {
  long long* d; cudaMalloc(&d, 296 * sizeof(long long));
  long long h[296];
  const int smem = 2 * TILE + 64;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode)
    for (int senders = 1; senders <= 2; ++senders)
      for (int grid : {2, 148}) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) k<0><<<grid, 256, smem>>>(iters, senders, d); else k<1><<<grid, 256, smem>>>(iters, senders, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%s  senders/pair=%d  grid=%3d : %.1f bytes/clk per sending SM (%.0f cycles per 32 KB tile)\n",
               mode == 0 ? "st.shared::cluster.v4 " : "cp.async.bulk to peer ", senders, grid, (double)iters * TILE / mx, mx / iters);
      }
  return 0;
}
