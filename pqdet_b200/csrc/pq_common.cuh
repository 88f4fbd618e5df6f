// Launch plumbing and block/warp primitives shared by the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pqdet_b200.h"
#include "pq_math.cuh"

#define PQ_FULL 0xffffffffu

// Every ABI call selects the device explicitly (the library has its own runtime instance and
// nn.DataParallel-style callers use one thread per GPU) and reports launch errors as codes.
#define PQ_ENTER(device)                                                   \
  do {                                                                     \
    if (cudaSetDevice(device) != cudaSuccess) return PQDET_ERR_CUDA;       \
  } while (0)

#define PQ_LAUNCH_CHECK()                                                  \
  do {                                                                     \
    if (cudaGetLastError() != cudaSuccess) return PQDET_ERR_CUDA;          \
  } while (0)

#define PQ_CUDA(call)                                                      \
  do {                                                                     \
    if ((call) != cudaSuccess) return PQDET_ERR_CUDA;                      \
  } while (0)

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

}  // namespace pq
