// metrics.cu -- the O(B*D) / O(B*D^2) kernels behind the cold loss variants and the evaluation-side consumers:
//   * column sums of X or X - Y            -> centroid_alignment_loss (sparsify_clip.py:487-505), compute_gap (:418-436),
//                                             mean off-diagonal cosine (:438-457) through |sum x|^2 - sum |x|^2
//   * D x D second-moment / covariance      -> W2 uniformity (uniformity.py:6-205, sparsify_clip.py:459-485) and
//                                             sparsify_loss backward: sum_j (x_i.x_j) x_j = x_i (X^T X), O(B D^2) instead
//                                             of a second B x B x D contraction (sparsify_clip.py:166-176)
//   * rows times a D x D matrix             -> X (X^T X) for that backward
//   * rank of the ground-truth entry of every row / column of a given score matrix -> compute_metric_ret (:357-416)
// All fp32 on CUDA cores with fixed-order two-stage reductions (bit-reproducible); none of them is on the training
// step's critical path (2 B D^2 is <= 0.1 % of one B x B x D contraction).
#include <string.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ column sums
// stage 1: block b sums rows [b*per, (b+1)*per) for 256 consecutive columns; stage 2: fixed-order sum over blocks
__global__ void __launch_bounds__(256) k_colsum_stage1(const void* __restrict__ X, const void* __restrict__ Y, int64_t n, int D,
                                                       int64_t ldX, int64_t ldY, int dtype, int64_t per,
                                                       float* __restrict__ scratch) {
  const int d = blockIdx.y * 256 + threadIdx.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(lo + per, n);
  float acc = 0.f;
  if (d < D) {
    for (int64_t r = lo; r < hi; ++r) {
      float v = scb_ld(X, dtype, r * ldX + d);
      if (Y) v -= scb_ld(Y, dtype, r * ldY + d);
      acc += v;
    }
    scratch[(int64_t)blockIdx.x * D + d] = acc;
  }
}
__global__ void __launch_bounds__(256) k_colsum_stage2(const float* __restrict__ scratch, int nb, int D, float scale,
                                                       float* __restrict__ out) {
  const int d = blockIdx.x * 256 + threadIdx.x;
  if (d >= D) return;
  float acc = 0.f;
  for (int b = 0; b < nb; ++b) acc += scratch[(int64_t)b * D + d];
  out[d] = acc * scale;
}

// ------------------------------------------------------------------ D x D second moment
// C[dm, dn] = sum_b (x[b, dm] - mu[dm]) (x[b, dn] - mu[dn]) over this block's K range; 64 x 64 output tile per CTA
constexpr int GT = 64, GK = 32;
__global__ void __launch_bounds__(256) k_gram_dd(const void* __restrict__ X, int64_t n, int D, int64_t ld, int dtype,
                                                 const float* __restrict__ mu, int64_t per, float* __restrict__ scratch) {
  __shared__ float As[GK][GT + 1];
  __shared__ float Bs[GK][GT + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  const int64_t lo = (int64_t)blockIdx.z * per, hi = min(lo + per, n);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t r0 = lo; r0 < hi; r0 += GK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < (GK * GT) / 256; ++i) {
      const int e = tid + 256 * i, k = e / GT, c = e % GT;
      const int64_t r = r0 + k;
      const bool rok = r < hi;
      const int dm = m0 + c, dn = n0 + c;
      As[k][c] = (rok && dm < D) ? scb_ld(X, dtype, r * ld + dm) - (mu ? mu[dm] : 0.f) : 0.f;
      Bs[k][c] = (rok && dn < D) ? scb_ld(X, dtype, r * ld + dn) - (mu ? mu[dn] : 0.f) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* out = scratch + (int64_t)blockIdx.z * D * D;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dm = m0 + ty + 16 * i, dn = n0 + tx + 16 * j;
      if (dm < D && dn < D) out[(int64_t)dm * D + dn] = acc[i][j];
    }
}
__global__ void __launch_bounds__(256) k_sum_parts(const float* __restrict__ scratch, int nparts, int64_t n, float scale,
                                                   float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int p = 0; p < nparts; ++p) acc += scratch[(int64_t)p * n + i];
  out[i] = acc * scale;
}

// ------------------------------------------------------------------ out[n x D] = X[n x D] . M[D x D]   (fp32)
__global__ void __launch_bounds__(256) k_rows_times_dd(const void* __restrict__ X, int64_t n, int D, int64_t ld, int dtype,
                                                       const float* __restrict__ M, float* __restrict__ out) {
  __shared__ float As[GT][GK + 1];    // [row][k]
  __shared__ float Bs[GK][GT + 1];    // [k][col]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * GT;
  const int c0 = blockIdx.y * GT;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < D; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < (GT * GK) / 256; ++i) {
      const int e = tid + 256 * i;
      const int r = e / GK, k = e % GK;
      As[r][k] = (r0 + r < n && k0 + k < D) ? scb_ld(X, dtype, (r0 + r) * ld + k0 + k) : 0.f;
      const int kk = e / GT, c = e % GT;
      Bs[kk][c] = (k0 + kk < D && c0 + c < D) ? M[(int64_t)(k0 + kk) * D + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[ty + 16 * i][k]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + ty + 16 * i;
      const int c = c0 + tx + 16 * j;
      if (r < n && c < D) out[r * D + c] = acc[i][j];
    }
}

// ------------------------------------------------------------------ rank of the ground truth in a score matrix
// line l of the matrix = elements base + l*stride_line + e*stride_elem, e < n_elem; rank[l] = #{e : s[l,e] > s[l,gt[l]]}
// (the position of the ground truth in a descending sort; ties do not count, as in a stable descending sort that
// happens to place the ground truth first among equals).  One warp per line.
__global__ void __launch_bounds__(256) k_rank_count(const void* __restrict__ S, int64_t n_lines, int64_t n_elem,
                                                    int64_t stride_line, int64_t stride_elem, int dtype,
                                                    const int64_t* __restrict__ line, const int64_t* __restrict__ gt,
                                                    int* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= n_lines) return;
  const int64_t l = line ? line[q] : q;
  const int64_t g = gt[q];
  const float ref = scb_ld(S, dtype, l * stride_line + g * stride_elem);
  int cnt = 0;
  for (int64_t e = lane; e < n_elem; e += 32) cnt += scb_ld(S, dtype, l * stride_line + e * stride_elem) > ref ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) rank[q] = cnt;
}

}  // namespace

extern "C" int scb_col_sum(const void* X, const void* Y, int64_t n, int D, int64_t ldX, int64_t ldY, int dtype, float scale,
                           float* scratch, int scratch_rows, float* out, void* stream) {
  SCB_CHECK_ARG(scb_dtype_ok(dtype), SCB_E_DTYPE, "col_sum: unsupported dtype %d", dtype);
  SCB_CHECK_ARG(X && out && scratch && n >= 0 && D > 0 && ldX >= D && (!Y || ldY >= D) && scratch_rows >= 1, SCB_E_ARG,
                "col_sum: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int nb = (int)((n + 127) / 128);
  if (nb > scratch_rows) nb = scratch_rows;
  if (nb < 1) nb = 1;
  const int64_t per = (n + nb - 1) / nb;
  k_colsum_stage1<<<dim3((unsigned)nb, (unsigned)((D + 255) / 256)), 256, 0, s>>>(X, Y, n, D, ldX, ldY, dtype, per, scratch);
  k_colsum_stage2<<<(unsigned)((D + 255) / 256), 256, 0, s>>>(scratch, nb, D, scale, out);
  SCB_CHECK_LAUNCH("col_sum");
  return 0;
}

extern "C" int scb_gram_dd(const void* X, int64_t n, int D, int64_t ld, int dtype, const float* mu, float scale, float* scratch,
                           int scratch_parts, float* out, void* stream) {
  SCB_CHECK_ARG(scb_dtype_ok(dtype), SCB_E_DTYPE, "gram_dd: unsupported dtype %d", dtype);
  SCB_CHECK_ARG(X && out && scratch && n >= 0 && D > 0 && ld >= D && scratch_parts >= 1, SCB_E_ARG, "gram_dd: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int np = (int)((n + 511) / 512);
  if (np > scratch_parts) np = scratch_parts;
  if (np < 1) np = 1;
  int64_t per = (n + np - 1) / np;
  per = (per + GK - 1) / GK * GK;
  const unsigned gt = (unsigned)((D + GT - 1) / GT);
  k_gram_dd<<<dim3(gt, gt, (unsigned)np), 256, 0, s>>>(X, n, D, ld, dtype, mu, per, scratch);
  const int64_t nn = (int64_t)D * D;
  k_sum_parts<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(scratch, np, nn, scale, out);
  SCB_CHECK_LAUNCH("gram_dd");
  return 0;
}

extern "C" int scb_sum_parts(const float* parts, int nparts, int64_t n, float scale, float* out, void* stream) {
  SCB_CHECK_ARG((parts && out) || n == 0, SCB_E_ARG, "sum_parts: null argument");
  SCB_CHECK_ARG(nparts >= 1 && n >= 0, SCB_E_ARG, "sum_parts: bad shape");
  if (n == 0) return 0;
  k_sum_parts<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parts, nparts, n, scale, out);
  SCB_CHECK_LAUNCH("sum_parts");
  return 0;
}

extern "C" int scb_rows_times_dd(const void* X, int64_t n, int D, int64_t ld, int dtype, const float* M, float* out,
                                 void* stream) {
  SCB_CHECK_ARG(scb_dtype_ok(dtype), SCB_E_DTYPE, "rows_times_dd: unsupported dtype %d", dtype);
  SCB_CHECK_ARG((X && M && out) || n == 0, SCB_E_ARG, "rows_times_dd: null argument");
  SCB_CHECK_ARG(n >= 0 && D > 0 && ld >= D, SCB_E_ARG, "rows_times_dd: bad shape");
  if (n == 0) return 0;
  k_rows_times_dd<<<dim3((unsigned)((n + GT - 1) / GT), (unsigned)((D + GT - 1) / GT)), 256, 0, (cudaStream_t)stream>>>(
      X, n, D, ld, dtype, M, out);
  SCB_CHECK_LAUNCH("rows_times_dd");
  return 0;
}

extern "C" int scb_rank_count(const void* S, int64_t n_lines, int64_t n_elem, int64_t stride_line, int64_t stride_elem,
                              int dtype, const int64_t* line, const int64_t* gt, int* rank, void* stream) {
  SCB_CHECK_ARG(scb_dtype_ok(dtype), SCB_E_DTYPE, "rank_count: unsupported dtype %d", dtype);
  SCB_CHECK_ARG((S && gt && rank) || n_lines == 0, SCB_E_ARG, "rank_count: null argument");
  SCB_CHECK_ARG(n_lines >= 0 && n_elem > 0, SCB_E_ARG, "rank_count: bad shape");
  if (n_lines == 0) return 0;
  k_rank_count<<<(unsigned)((n_lines + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, n_lines, n_elem, stride_line, stride_elem,
                                                                             dtype, line, gt, rank);
  SCB_CHECK_LAUNCH("rank_count");
  return 0;
}

// ===================================================================================================================
// SM-free row all-gather between the GPUs of one node.  NCCL's all-gather is a kernel: it takes
// SMs away from the persistent sweeps it is meant to overlap with (measured at 8 GPUs: a sweep that shares the machine
// with ncclDevKernel_AllGather_RING_LL runs 270 us instead of 217 us).  Here every rank PUSHES its shard into every peer's
// gather buffer with the copy engines (cudaMemcpyAsync on peer-mapped pointers over NVLink), followed by a 4-byte
// "epoch" flag per peer on the same stream; the only kernel is the consumer's one-warp wait on its own flag words.
// Buffers are exchanged once through CUDA IPC (explicit alloc / open / close: the one place where the library owns
// device memory, as the communicator-handle exception of the boundary contract allows).
// ===================================================================================================================
namespace {
__device__ __forceinline__ void spin_until_ge(const volatile int* w, int v) {
  long long t0 = 0;
  bool timed = false;
  while (*w < v) {
    if (!timed) { t0 = (long long)clock64(); timed = true; }
    if ((long long)clock64() - t0 > 60000000000ll) __trap();   // ~30 s: a peer that never arrives must end the job, not hang
    __nanosleep(100);
  }
}
// All four protocol kernels are ONE warp (lane = peer index) and read the epoch from a device counter, so a step that
// contains them can be replayed from a CUDA graph: nothing about the epoch is baked into the launch.
//   begin:   ++epoch; wait until every peer has RELEASED the previous contents of its gather buffer (done[p] >= epoch-1)
//   arrive:  (after the copy-engine pushes, same stream) arrived-word of every peer for my slot = epoch
//   wait:    until every peer's shard of this epoch has arrived in MY buffer
//   release: (after my last read) done-word of every peer for my slot = epoch
__global__ void k_peer_begin(int* __restrict__ epoch, const volatile int* __restrict__ done, int world) {
  int e = 0;
  if (threadIdx.x == 0) { e = *epoch + 1; *epoch = e; }
  e = __shfl_sync(0xffffffffu, e, 0);
  if ((int)threadIdx.x < world) spin_until_ge(done + threadIdx.x, e - 1);
}
__global__ void k_peer_signal(const int* __restrict__ epoch, int* const* __restrict__ words, int world) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(words[threadIdx.x]) = *epoch;
  }
}
__global__ void k_peer_wait(const int* __restrict__ epoch, const volatile int* __restrict__ arrived, int world) {
  if ((int)threadIdx.x < world) spin_until_ge(arrived + threadIdx.x, *epoch);
  __threadfence_system();
}
// The same push done by the SMs (plain 16-byte stores over NVLink) for the gather a step STARTS with: nothing else runs
// yet, so the SMs are free, and one copy engine moves a 4 MB shard at only ~360 GB/s, one peer after the other (measured at
// 8 GPUs: 7 x 12 us exposed in front of the first sweep).  Every thread loads its 16 bytes once and stores them to all n
// destinations; the last CTA to finish (device-scope counter, re-armed for the next launch) publishes the epoch flags
// after a system-scope fence.
struct PushDst {
  void* p[32];
};
__global__ void __launch_bounds__(256) k_peer_push_sm(const uint4* __restrict__ src, int64_t n16, const PushDst d, int n,
                                                      const int* __restrict__ epoch, int* const* __restrict__ words, int world,
                                                      unsigned* __restrict__ counter) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += 2 * stride) {
    const int64_t i2 = i + stride;
    const uint4 v0 = __ldg(src + i);
    uint4 v1 = make_uint4(0u, 0u, 0u, 0u);
    if (i2 < n16) v1 = __ldg(src + i2);
    for (int k = 0; k < n; ++k) {
      uint4* dst = reinterpret_cast<uint4*>(d.p[k]);
      dst[i] = v0;
      if (i2 < n16) dst[i2] = v1;
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *counter = 0u;
    __threadfence_system();
    if ((int)threadIdx.x < world) *reinterpret_cast<volatile int*>(words[threadIdx.x]) = *epoch;
  }
}
struct ReleaseMany {
  const int* epoch[8];
  int* const* words[8];
};
// several roles released by one launch: warp r handles role r
__global__ void k_peer_release_many(const ReleaseMany a, int world) {
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < world) {
    __threadfence_system();
    *reinterpret_cast<volatile int*>(a.words[r][lane]) = *a.epoch[r];
  }
}
}  // namespace

extern "C" int scb_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  SCB_CHECK_ARG(bytes > 0 && ptr && handle64, SCB_E_ARG, "peer_alloc: bad argument");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) { scb_set_error("peer_alloc: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int scb_peer_open(const unsigned char* handle64, void** ptr) {
  SCB_CHECK_ARG(handle64 && ptr, SCB_E_ARG, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { scb_set_error("peer_open: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
  return 0;
}
extern "C" int scb_peer_close(void* ptr, int opened) {
  const cudaError_t e = opened ? cudaIpcCloseMemHandle(ptr) : cudaFree(ptr);
  if (e != cudaSuccess) { scb_set_error("peer_close: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
  return 0;
}
// One gather epoch, sender side.  scb_peer_begin goes on the CONSUMER's stream (it advances the epoch the later
// scb_peer_wait on that stream compares against, and waits for the peers' release of the previous epoch);
// scb_peer_push may go on a side stream that waits for it: `bytes` from src to each dst[k] with the copy engines
// (peer-mapped or local pointers), then this rank's arrived-word on every peer.
// epoch: device int owned by this role on this rank; done: my done[world] words; arrived_words: DEVICE array of `world`
// pointers (peer-mapped) to arrived[my rank] on each rank.
extern "C" int scb_peer_begin(int* epoch, const int* done, int world, void* stream) {
  SCB_CHECK_ARG(epoch && done && world >= 1 && world <= 32, SCB_E_ARG, "peer_begin: bad argument");
  k_peer_begin<<<1, 32, 0, (cudaStream_t)stream>>>(epoch, done, world);
  SCB_CHECK_LAUNCH("peer_begin");
  return 0;
}
extern "C" int scb_peer_push(const void* src, int64_t bytes, void* const* dst, int n, const int* epoch,
                             int* const* arrived_words, int world, void* stream) {
  SCB_CHECK_ARG(src && dst && epoch && arrived_words && n >= 0 && bytes >= 0 && world >= 1 && world <= 32, SCB_E_ARG,
                "peer_push: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  for (int k = 0; k < n; ++k) {
    const cudaError_t e = cudaMemcpyAsync(dst[k], src, (size_t)bytes, cudaMemcpyDefault, s);
    if (e != cudaSuccess) { scb_set_error("peer_push (copy %d): %s", k, cudaGetErrorString(e)); return (int)e; }
  }
  k_peer_signal<<<1, 32, 0, s>>>(epoch, arrived_words, world);
  SCB_CHECK_LAUNCH("peer_push");
  return 0;
}
// scb_peer_push with the copies done by a kernel (see k_peer_push_sm).  counter: a zero-initialised device word owned by
// this role (left at zero by every launch).  bytes and every pointer must be multiples of 16.
extern "C" int scb_peer_push_sm(const void* src, int64_t bytes, void* const* dst, int n, const int* epoch,
                                int* const* arrived_words, int world, unsigned* counter, void* stream) {
  SCB_CHECK_ARG(src && dst && epoch && arrived_words && counter && n >= 1 && n <= 32 && bytes > 0 && world >= 1 && world <= 32,
                SCB_E_ARG, "peer_push_sm: bad argument");
  SCB_CHECK_ARG(bytes % 16 == 0 && scb_aligned16(src), SCB_E_ARG, "peer_push_sm: needs 16-byte granularity");
  PushDst d{};
  for (int k = 0; k < n; ++k) {
    SCB_CHECK_ARG(dst[k] && scb_aligned16(dst[k]), SCB_E_ARG, "peer_push_sm: destination %d is not 16-byte aligned", k);
    d.p[k] = dst[k];
  }
  const int64_t n16 = bytes / 16;
  int64_t blocks = (n16 + 511) / 512;                      // two elements per thread and pass
  const int64_t cap = 4 * (int64_t)scb_num_sms();
  if (blocks > cap) blocks = cap;
  k_peer_push_sm<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint4*>(src), n16, d, n, epoch,
                                                                     arrived_words, world, counter);
  SCB_CHECK_LAUNCH("peer_push_sm");
  return 0;
}
// the copies alone (to spread a large shard's pushes over several streams; scb_peer_push with n = 0 then sends the flag)
extern "C" int scb_peer_copy(const void* src, int64_t bytes, void* const* dst, int n, void* stream) {
  SCB_CHECK_ARG(src && dst && n >= 0 && bytes >= 0, SCB_E_ARG, "peer_copy: bad argument");
  for (int k = 0; k < n; ++k) {
    const cudaError_t e = cudaMemcpyAsync(dst[k], src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    if (e != cudaSuccess) { scb_set_error("peer_copy (copy %d): %s", k, cudaGetErrorString(e)); return (int)e; }
  }
  return 0;
}
// receiver side: stream-ordered wait until every rank's shard of the current epoch has landed in my buffer
extern "C" int scb_peer_wait(const int* epoch, const int* arrived, int world, void* stream) {
  SCB_CHECK_ARG(epoch && arrived && world >= 1 && world <= 32, SCB_E_ARG, "peer_wait: bad argument");
  k_peer_wait<<<1, 32, 0, (cudaStream_t)stream>>>(epoch, arrived, world);
  SCB_CHECK_LAUNCH("peer_wait");
  return 0;
}
// after the last read of my gather buffer: tell every peer it may overwrite its slot (done_words: device array of
// `world` pointers to done[my rank] on each rank)
extern "C" int scb_peer_release(const int* epoch, int* const* done_words, int world, void* stream) {
  SCB_CHECK_ARG(epoch && done_words && world >= 1 && world <= 32, SCB_E_ARG, "peer_release: bad argument");
  k_peer_signal<<<1, 32, 0, (cudaStream_t)stream>>>(epoch, done_words, world);
  SCB_CHECK_LAUNCH("peer_release");
  return 0;
}

// scb_peer_release for up to 8 roles in one launch (epochs[r], done_words[r] as for scb_peer_release)
extern "C" int scb_peer_release_many(const int* const* epochs, int* const* const* done_words, int n_roles, int world,
                                     void* stream) {
  SCB_CHECK_ARG(epochs && done_words && n_roles >= 1 && n_roles <= 8 && world >= 1 && world <= 32, SCB_E_ARG,
                "peer_release_many: bad argument");
  ReleaseMany a{};
  for (int r = 0; r < n_roles; ++r) { a.epoch[r] = epochs[r]; a.words[r] = done_words[r]; }
  k_peer_release_many<<<1, 32 * n_roles, 0, (cudaStream_t)stream>>>(a, world);
  SCB_CHECK_LAUNCH("peer_release_many");
  return 0;
}
