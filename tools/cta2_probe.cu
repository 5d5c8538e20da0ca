// cta2_probe.cu -- issue rate and correctness of tcgen05.mma cta_group::2 (M=256 across a CTA pair, each CTA staging
// half of the B operand) against the single-CTA form the pass kernels use today.  Operands are written into shared
// memory by hand in the 128-byte-swizzle K-major layout the TMA produces, so the probe needs no tensor maps.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/cta2_probe tools/cta2_probe.cu
//   tools/cta2_probe            (prints max |err| and cycles per MMA for every variant)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../sparsify_clip_b200/csrc/ptx.cuh"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit2(uint32_t bar) {   // arrives on `bar` of both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// byte offset of element (row r, k) in a [rows][64] 16-bit tile, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2);
}

// A: [CTAS*128][64] bf16 (K contiguous), B: [N][64] bf16, D: [CTAS*128][N] fp32 = A * B^T accumulated `reps` times.
template <int CTAS, int N>
__global__ void __launch_bounds__(128, 1) k_probe(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                   float* __restrict__ D, long long* __restrict__ clk, int reps, int write_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 128 x 64 x 2 = 16 KB
  uint8_t* sB = smem + 16384;         // (N / CTAS) x 64 x 2
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CTAS;
  constexpr int NB = N / CTAS;        // rows of B this CTA stages

  for (int e = tid; e < 128 * 64; e += 128) {
    int r = e >> 6, k = e & 63;
    *(__nv_bfloat16*)(sA + sw128_off(r, k)) = A[(size_t)(crank * 128 + r) * 64 + k];
  }
  for (int e = tid; e < NB * 64; e += 128) {
    int r = e >> 6, k = e & 63;
    *(__nv_bfloat16*)(sB + sw128_off(r, k)) = B[(size_t)(crank * NB + r) * 64 + k];
  }
  if (tid == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar_done), 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async_smem();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (CTAS == 2) tmem_alloc2(ptx::smem_u32(&tmem_base), N); else ptx::tmem_alloc(ptx::smem_u32(&tmem_base), N);
  }
  ptx::tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;

  long long t0 = 0, t1 = 0;
  if (warp == 1 && crank == 0) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_f16(128 * CTAS, N, 1, 1, 0, 0);
      const uint64_t ad = ptx::desc_kmajor(ptx::smem_u32(sA)), bd = ptx::desc_kmajor(ptx::smem_u32(sB));
      t0 = clock64();
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (CTAS == 2) umma_ss2(tm, ad + 2 * kk, bd + 2 * kk, idesc, (rep | kk) != 0);
          else ptx::umma_ss(tm, ad + 2 * kk, bd + 2 * kk, idesc, (rep | kk) != 0);
        }
      }
      if (CTAS == 2) umma_commit2(ptx::smem_u32(&bar_done)); else ptx::umma_commit(ptx::smem_u32(&bar_done));
    }
    __syncwarp();
  }
  ptx::mbar_wait(ptx::smem_u32(&bar_done), 0, 1);
  t1 = clock64();
  ptx::tc_fence_after();
  if (warp == 1 && crank == 0) {
    long long dt = t1 - t0;   // only the elected lane has t0; take the max over the warp
    for (int o = 16; o; o >>= 1) {
      long long other = __shfl_xor_sync(0xffffffffu, t0, o);
      t0 = other > t0 ? other : t0;
    }
    dt = t1 - t0;
    if (lane == 0) clk[cluster_id] = dt;
  }
  if (write_out) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
      ptx::tmem_ld_wait();
      const size_t row = (size_t)cluster_id * (128 * CTAS) + crank * 128 + warp * 32 + lane;
#pragma unroll
      for (int c = 0; c < 32; ++c) D[row * N + c0 + c] = __uint_as_float(v[c]);
    }
  }
  ptx::tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (CTAS == 2) tmem_dealloc2(tm, N); else ptx::tmem_dealloc(tm, N);
  }
}

template <int CTAS, int N>
void run(int n_clusters, int reps, bool check) {
  const int M = 128 * CTAS;
  std::vector<__nv_bfloat16> hA((size_t)M * 64), hB((size_t)N * 64);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(1234);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dclk;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, (size_t)n_clusters * M * N * 4));
  CK(cudaMalloc(&dclk, n_clusters * sizeof(long long)));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 1024 + 16384 + (size_t)(N / CTAS) * 128;
  CK(cudaFuncSetAttribute(k_probe<CTAS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * CTAS);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const __nv_bfloat16 *cA = dA, *cB = dB;
  int w = check ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, k_probe<CTAS, N>, cA, cB, dD, dclk, reps, w));
  CK(cudaDeviceSynchronize());
  std::vector<long long> hclk(n_clusters);
  CK(cudaMemcpy(hclk.data(), dclk, n_clusters * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0, mn = 1ll << 62;
  for (auto c : hclk) { mx = c > mx ? c : mx; mn = c < mn ? c : mn; }
  double err = 0.0;
  if (check) {
    std::vector<float> hD((size_t)n_clusters * M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    for (int cl = 0; cl < n_clusters; ++cl)
      for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
          double acc = 0.0;
          for (int k = 0; k < 64; ++k) acc += (double)fA[(size_t)i * 64 + k] * fB[(size_t)j * 64 + k];
          double d = fabs(acc * reps - hD[((size_t)cl * M + i) * N + j]);
          err = d > err ? d : err;
        }
  }
  printf("cta_group::%d M=%3d N=%3d clusters=%3d mmas=%4d  cycles/MMA min %.1f max %.1f", CTAS, M, N, n_clusters, 4 * reps,
         (double)mn / (4 * reps), (double)mx / (4 * reps));
  if (check) printf("   max|err| %.3g", err);
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dclk);
}

// ---------------------------------------------------------------------------------------------------------------
// The data path of the planned pair kernel, one tile pair, K = 64:
//   MMA1 (cta_group::2, transposed): St_c[j, i] = sum_k X_c[j, k] A[i, k]   -- M side = CTA c's own column tile X_c
//        (128 rows, K-major), N side = the SHARED row block A, each CTA staging 64 of its 128 rows;
//   epilogue: thread j reads St_c[j, 0..127] and writes W_c[i, j] = bf16(0.25 St_c[j, i]) as an MN-major SW128 tile
//        (rows = K index j, 64 consecutive i per 128-byte row, two 64-wide i blocks 16 KB apart);
//   MMA2 (cta_group::1): OUT_c[i, d] = sum_j W_c[i, j] X_c[j, d]   -- A = W_c with a_major = MN, B = the SAME X_c bytes
//        read MN-major (rows = K index j, 64 d per row).
__global__ void __launch_bounds__(128, 1) k_chain(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ X,
                                                   float* __restrict__ St, float* __restrict__ OUT) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                  // [128 j][64] 16 KB
  uint8_t* sA = smem + 16384;          // [64 i][64]   8 KB (this CTA's half of the row block)
  uint8_t* sW = smem + 32768;          // 2 x [128 j][64 i] 32 KB, MN-major
  __shared__ __align__(8) uint64_t bar1, bar2;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  for (int e = tid; e < 128 * 64; e += 128) {
    int r = e >> 6, k = e & 63;
    *(__nv_bfloat16*)(sX + sw128_off(r, k)) = X[(size_t)(crank * 128 + r) * 64 + k];
  }
  for (int e = tid; e < 64 * 64; e += 128) {
    int r = e >> 6, k = e & 63;
    *(__nv_bfloat16*)(sA + sw128_off(r, k)) = A[(size_t)(crank * 64 + r) * 64 + k];
  }
  if (tid == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar1), 1);
    ptx::mbar_init(ptx::smem_u32(&bar2), 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async_smem();
  cluster_sync_all();
  if (warp == 0) tmem_alloc2(ptx::smem_u32(&tmem_base), 256);
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (warp == 1 && crank == 0) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_f16(256, 128, 1, 1, 0, 0);
      const uint64_t ad = ptx::desc_kmajor(ptx::smem_u32(sX)), bd = ptx::desc_kmajor(ptx::smem_u32(sA));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_ss2(tm, ad + 2 * kk, bd + 2 * kk, idesc, kk != 0);
      umma_commit2(ptx::smem_u32(&bar1));
    }
    __syncwarp();
  }
  ptx::mbar_wait(ptx::smem_u32(&bar1), 0, 1);
  ptx::tc_fence_after();
  const int j = warp * 32 + lane;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) St[((size_t)crank * 128 + j) * 128 + c0 + c] = __uint_as_float(v[c]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i0 = c0 + 8 * q;
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        w[u] = ptx::pack_bf16(0.25f * __uint_as_float(v[8 * q + 2 * u]), 0.25f * __uint_as_float(v[8 * q + 2 * u + 1]));
      const uint32_t addr = ptx::smem_u32(sW) + (uint32_t)((i0 >> 6) * 16384 + (j >> 3) * 1024 + (j & 7) * 128 +
                                                           ((((i0 & 63) >> 3) ^ (j & 7)) << 4));
      ptx::st_shared_v4(addr, w[0], w[1], w[2], w[3]);
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_f16(128, 64, 1, 1, 1, 1);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = ptx::desc_mnmajor(ptx::smem_u32(sW) + 2048u * ks, 16384);
        const uint64_t bd = ptx::desc_mnmajor(ptx::smem_u32(sX) + 2048u * ks, 16384);
        ptx::umma_ss(tm + 128, ad, bd, idesc, ks != 0);
      }
      ptx::umma_commit(ptx::smem_u32(&bar2));
    }
    __syncwarp();
  }
  ptx::mbar_wait(ptx::smem_u32(&bar2), 0, 2);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + 128 + c0, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c) OUT[((size_t)crank * 128 + j) * 64 + c0 + c] = __uint_as_float(v[c]);
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tm, 256);
}

void run_chain() {
  std::vector<__nv_bfloat16> hA(128 * 64), hX(256 * 64);
  std::vector<float> fA(hA.size()), fX(hX.size());
  srand(77);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hX.size(); ++i) { hX[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fX[i] = __bfloat162float(hX[i]); }
  __nv_bfloat16 *dA, *dX;
  float *dSt, *dOut;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dX, hX.size() * 2));
  CK(cudaMalloc(&dSt, 2 * 128 * 128 * 4));
  CK(cudaMalloc(&dOut, 2 * 128 * 64 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 1024 + 65536;
  CK(cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const __nv_bfloat16 *cA = dA, *cX = dX;
  CK(cudaLaunchKernelEx(&cfg, k_chain, cA, cX, dSt, dOut));
  CK(cudaDeviceSynchronize());
  std::vector<float> hSt(2 * 128 * 128), hOut(2 * 128 * 64);
  CK(cudaMemcpy(hSt.data(), dSt, hSt.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost));
  double e1 = 0.0, e2 = 0.0, mag = 0.0;
  std::vector<float> W(128 * 128);
  for (int c = 0; c < 2; ++c) {
    for (int j = 0; j < 128; ++j)
      for (int i = 0; i < 128; ++i) {
        double acc = 0.0;
        for (int k = 0; k < 64; ++k) acc += (double)fX[((size_t)c * 128 + j) * 64 + k] * fA[(size_t)i * 64 + k];
        e1 = fmax(e1, fabs(acc - hSt[((size_t)c * 128 + j) * 128 + i]));
        W[(size_t)i * 128 + j] = __bfloat162float(__float2bfloat16(0.25f * hSt[((size_t)c * 128 + j) * 128 + i]));
      }
    for (int i = 0; i < 128; ++i)
      for (int d = 0; d < 64; ++d) {
        double acc = 0.0;
        for (int j = 0; j < 128; ++j) acc += (double)W[(size_t)i * 128 + j] * fX[((size_t)c * 128 + j) * 64 + d];
        e2 = fmax(e2, fabs(acc - hOut[((size_t)c * 128 + i) * 64 + d]));
        mag = fmax(mag, fabs(acc));
      }
  }
  printf("chain: transposed cta_group::2 MMA1 max|err| %.3g; MN-major W -> MMA2 max|err| %.3g (max |OUT| %.3g)\n", e1, e2, mag);
  cudaFree(dA); cudaFree(dX); cudaFree(dSt); cudaFree(dOut);
}

}  // namespace

int main() {
  run_chain();
  run<1, 128>(1, 1, true);
  run<1, 256>(1, 1, true);
  run<2, 128>(1, 1, true);
  run<2, 256>(1, 1, true);
  run<2, 128>(3, 2, true);
  for (int ncl : {1, 74}) {
    run<1, 128>(ncl * 2, 256, false);
    run<1, 256>(ncl * 2, 256, false);
    run<2, 128>(ncl, 256, false);
    run<2, 256>(ncl, 256, false);
  }
  return 0;
}
