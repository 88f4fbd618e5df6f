// Launch plumbing and block/warp primitives shared by the sm_100a kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pqdet_b200.h"
#include "pq_math.cuh"

#define PQ_FULL 0xffffffffu

// Every ABI call selects the device explicitly (the library has its own runtime instance and
// nn.DataParallel-style callers use one thread per GPU) and reports launch errors as codes.  The
// selection is scoped: PyTorch reads the thread's current device from the same CUDA runtime, so the
// caller's device is put back on every return path (including the PQ_CUDA / PQ_LAUNCH_CHECK early
// returns) -- a call with device=1 from a thread sitting on device 0 must not move that thread.
namespace pq {
struct DeviceScope {
  int prev = -1;
  bool ok = false;
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
    ok = (prev == device) || (cudaSetDevice(device) == cudaSuccess);
    if (ok && prev == device) prev = -1;          // nothing to restore
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};
}  // namespace pq

#define PQ_ENTER(device)                                                   \
  pq::DeviceScope pq_device_scope_(device);                                \
  do {                                                                     \
    if (!pq_device_scope_.ok) return PQDET_ERR_CUDA;                       \
  } while (0)

#define PQ_LAUNCH_CHECK()                                                  \
  do {                                                                     \
    if (cudaGetLastError() != cudaSuccess) return PQDET_ERR_CUDA;          \
  } while (0)

#define PQ_CUDA(call)                                                      \
  do {                                                                     \
    if ((call) != cudaSuccess) return PQDET_ERR_CUDA;                      \
  } while (0)

// conf > thr needs sigmoid(x) > thr; x <= logit(thr) - margin can never pass (the margin is ~1e3 times the
// worst-case error of the fp32 sigmoid).  thr <= 0: no prefilter; thr >= 1: nothing passes.
static inline float logit_lo_for(float thr_f) {
  const double t = (double)thr_f;
  if (!(t > 0.0)) return -INFINITY;
  if (t >= 1.0) return INFINITY;
  const double lg = log(t / (1.0 - t));
  return (float)(lg - 1e-3 * (1.0 + fabs(lg)));
}

namespace pq {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// Streaming read: data that is touched once (raw heads, labels) should not displace L1 lines.
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

// float <-> order-preserving uint (for atomicMax on coordinates that may be negative)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Block-wide in-place bitonic sort (ascending) of P = 2^k uint64 keys; `a` may point to shared or
// global memory (the global variant is the any-size fallback).  All threads of the block call it.
__device__ __forceinline__ void bitonic_sort_block(uint64_t* a, int P) {
  const int half = P >> 1;
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < half; t += blockDim.x) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        uint64_t x = a[i], y = a[l];
        bool asc = (i & k) == 0;
        if ((x > y) == asc) {
          a[i] = y;
          a[l] = x;
        }
      }
      __syncthreads();
    }
  }
}

// Inclusive warp scan (sum) of an int.
__device__ __forceinline__ int warp_inclusive_sum(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(PQ_FULL, v, d);
    if (lane_id() >= d) v += n;
  }
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(PQ_FULL, v, d));
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(PQ_FULL, v, d);
  return v;
}

// ---- TMA (bulk async copy) + mbarrier helpers for streaming kernels whose tiles are contiguous runs ----------
// cp.async.bulk moves a 16-byte aligned, 16-byte multiple run between global and shared memory without touching
// registers; completion of a load is signalled on an mbarrier (transaction bytes), of a store through bulk groups.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity; traps instead of hanging if the bytes never arrive.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// One lane of a converged warp; the surrounding loop stays warp-uniform so that descriptors and addresses live in
// uniform registers (a loop run by `lane == 0` alone pays register -> uniform-register moves in front of every MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}


// ---- host side: cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda) ---------
typedef CUresult (*PqEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PqEncodeTiledFn encode_tiled_fn() {
  static PqEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (PqEncodeTiledFn)p;
  }();
  return fn;
}


// Device limits that the persistent kernels size themselves by, queried once per device.
struct DeviceLimits { int max_smem_optin, sms; };
inline bool device_limits(int device, DeviceLimits* out) {
  static DeviceLimits cache[64];
  static bool have[64];
  if (device >= 0 && device < 64 && have[device]) { *out = cache[device]; return true; }
  DeviceLimits d;
  if (cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return false;
  if (device >= 0 && device < 64) { cache[device] = d; have[device] = true; }
  *out = d;
  return true;
}

}  // namespace pq
