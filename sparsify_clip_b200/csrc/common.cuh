// common.cuh -- shared host/device helpers for libscb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scb200.h"

// ----------------------------------------------------------------------------- errors
void scb_set_error(const char* fmt, ...);

#define SCB_CHECK_ARG(cond, code, ...)                                                   \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      scb_set_error(__VA_ARGS__);                                                        \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)

#define SCB_CHECK_LAUNCH(what)                                                           \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      scb_set_error("%s: %s", (what), cudaGetErrorString(e__));                          \
      return (int)e__;                                                                   \
    }                                                                                    \
  } while (0)

// ----------------------------------------------------------------------------- per-device launch state
// Nothing below is keyed by "the first device seen": the SM count and the opt-in to > 48 KB of dynamic shared memory
// (cudaFuncSetAttribute is per device / context) are cached per CUDA device in lock-free slots, so a process that
// drives several GPUs (the reference wraps its model in DataParallel) gets the right value on each.
#include <atomic>
constexpr int kScbMaxDevices = 64;
static inline int scb_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
inline int scb_num_sms() {
  static std::atomic<int> cache[kScbMaxDevices];
  const int dev = scb_current_device();
  int n = (dev >= 0 && dev < kScbMaxDevices) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev >= 0 && dev < kScbMaxDevices) cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
// `done` is one bit mask per kernel family (a function-local static of the launcher); idempotent, so a race between two
// host threads only repeats the call
template <typename... K>
inline cudaError_t scb_opt_in_smem(std::atomic<unsigned long long>& done, int bytes, K... kernels) {
  const int dev = scb_current_device();
  const unsigned long long bit = (dev >= 0 && dev < kScbMaxDevices) ? (1ull << dev) : 0ull;
  if (bit && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  cudaError_t e = cudaSuccess;
  const cudaError_t rs[] = {cudaFuncSetAttribute(kernels, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)...};
  for (cudaError_t r : rs) if (r != cudaSuccess) e = r;
  if (e == cudaSuccess && bit) done.fetch_or(bit, std::memory_order_release);
  return e;
}

static inline bool scb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int scb_dtype_size(int dtype) { return dtype == SCB_F32 ? 4 : 2; }
static inline bool scb_dtype_ok(int dtype) { return dtype == SCB_F32 || dtype == SCB_BF16 || dtype == SCB_F16; }

#define SCB_LOG2E 1.4426950408889634f
#define SCB_LN2 0.6931471805599453f

// ----------------------------------------------------------------------------- device loads
// Scalar element load with a warp-uniform dtype switch.
__device__ __forceinline__ float scb_ld(const void* __restrict__ p, int dtype, int64_t idx) {
  if (dtype == SCB_F32) return __ldg(reinterpret_cast<const float*>(p) + idx);
  if (dtype == SCB_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
  return __half2float(reinterpret_cast<const __half*>(p)[idx]);
}

// 8 consecutive elements starting at idx (idx % 8 == 0, base 16-byte aligned, row stride % 8 == 0).
__device__ __forceinline__ void scb_ld8(const void* __restrict__ p, int dtype, int64_t idx, float (&v)[8]) {
  if (dtype == SCB_F32) {
    const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + idx);
    float4 a = __ldg(q), b = __ldg(q + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + idx));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    if (dtype == SCB_BF16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        float2 f = __half22float2(h);
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
      }
    }
  }
}

__device__ __forceinline__ void scb_st(void* p, int dtype, int64_t idx, float x) {
  if (dtype == SCB_F32) reinterpret_cast<float*>(p)[idx] = x;
  else if (dtype == SCB_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(x);
  else reinterpret_cast<__half*>(p)[idx] = __float2half_rn(x);
}

__device__ __forceinline__ float scb_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float scb_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float scb_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
