// L2 -> SM streaming probe for the B x B passes' access pattern: every CTA walks the SAME sequence of column tiles of an
// L2-resident [32768 x 512] bf16 matrix, loading [128 rows x 64 cols] boxes (128B swizzle) through a ring of 16 KB
// slots; the consumer only recycles the slots.  Reports bytes/clk per SM for unicast and for multicast inside clusters
// of 2 / 4 (each CTA issues 1/csz of the chunks with a multicast mask).   nvcc -gencode arch=compute_100a,code=sm_100a -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{.reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) { while (!try_wait(bar, par)) {} }
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* m, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_mc(uint32_t dst, const CUtensorMap* m, int x, int y, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(x), "r"(y), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int SLOT = 16384;
// CSZ = cluster size (1: unicast).  nslots ring slots.  Each of ntiles tiles = 8 chunks.  delay = clocks the consumer
// holds a chunk (emulates MMA time) before releasing its slot.
template <int CSZ>
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, int ntiles, int nslots, int delay, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t base = (smem_u32(sm) + 1023u) & ~1023u;
  const uint32_t bars = base + nslots * SLOT;   // full[nslots], empty[nslots]
  uint32_t rank = 0;
  if (CSZ > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslots; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (nslots + i), CSZ); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync();
  const long long t0 = clock64();
  if (warp == 0) {                       // producer
    if (lane == 0) {
      uint32_t slot = 0, par = 1;
      for (int t = 0; t < ntiles; ++t)
        for (int kc = 0; kc < 8; ++kc) {
          wait(bars + 8 * (nslots + slot), par);           // all CTAs of the cluster released this slot
          mbar_expect_tx(bars + 8 * slot, SLOT);
          if (CSZ == 1) tma_load(base + slot * SLOT, &tm, kc * 64, t * 128, bars + 8 * slot);
          else if ((kc % CSZ) == (int)rank) tma_load_mc(base + slot * SLOT, &tm, kc * 64, t * 128, bars + 8 * slot, (uint16_t)((1u << CSZ) - 1));
          if (++slot == (uint32_t)nslots) { slot = 0; par ^= 1; }
        }
    }
  } else {                               // consumer
    if (lane == 0) {
      uint32_t slot = 0, par = 0;
      for (int t = 0; t < ntiles; ++t)
        for (int kc = 0; kc < 8; ++kc) {
          wait(bars + 8 * slot, par);
          if (delay) { const long long c0 = clock64(); while (clock64() - c0 < delay) {} }
          if (CSZ == 1) mbar_arrive(bars + 8 * (nslots + slot));
          else for (int c = 0; c < CSZ; ++c) mbar_arrive_cluster(bars + 8 * (nslots + slot), c);
          if (++slot == (uint32_t)nslots) { slot = 0; par ^= 1; }
        }
    }
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CSZ>
void run(const CUtensorMap& tm, int nslots, int delay, long long* d, int grid) {
  const int ntiles = 256;
  const size_t smem = (size_t)nslots * SLOT + 2048;
  cudaFuncSetAttribute(k<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<CSZ>, tm, ntiles, nslots, delay, d);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("csz %d: %s\n", CSZ, cudaGetErrorString(e)); exit(1); }
    cudaEventElapsedTime(&ms, e0, e1);
  }
  long long h[160]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = (double)ntiles * 8 * SLOT;
  printf("cluster %d  slots %2d  delay %4d : %6.1f B/clk/SM  (%.0f clk per 16 KB chunk, %.3f ms, %.2f TB/s into %d SMs)\n", CSZ, nslots,
         delay, bytes / mx, mx / (ntiles * 8), ms, bytes * grid / (ms * 1e-3) / 1e12, grid);
}

int main() {
  const int64_t rows = 32768; const int D = 512;
  void* X; cudaMalloc(&X, rows * D * 2); cudaMemset(X, 1, rows * D * 2);
  long long* d; cudaMalloc(&d, 160 * sizeof(long long));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((PFN_enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  for (int nslots : {4, 6, 8, 12}) {
    run<1>(tm, nslots, 0, d, 148);
    run<2>(tm, nslots, 0, d, 148);
    run<4>(tm, nslots, 0, d, 148);
  }
  for (int delay : {256, 330, 384}) {
    for (int nslots : {4, 6, 8}) {
      run<1>(tm, nslots, delay, d, 148);
      run<2>(tm, nslots, delay, d, 148);
    }
  }
  return 0;
}
