// Microbenchmark: SM-initiated reads of pinned host memory over PCIe (zero-copy), by request shape.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pcie_probe profiles/tools/pcie_probe.cu && /tmp/pcie_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float ldnc(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
// every warp reads `span` bytes (32..512) at pseudo-random line-aligned offsets, ILP requests in flight per lane
template <int ILP>
__global__ void rnd(const float* __restrict__ src, size_t nlines, int span_floats, int iters, float* sink) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  uint64_t s = warp * 0x9E3779B97F4A7C15ull + 12345;
  float acc = 0.f;
  const int groups = 32 / span_floats;               // independent spans per warp instruction when span < 32 floats
  for (int it = 0; it < iters; ++it) {
    float v[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      uint64_t r = s >> 20;
      size_t line;
      int off;
      if (span_floats >= 32) { line = r % nlines; off = lane; }
      else { int g = lane / span_floats; line = (r + (uint64_t)g * 0x9E3779B1ull) % nlines; off = lane % span_floats; }
      v[u] = ldnc(src + line * 32 + off);
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) acc += v[u];
  }
  (void)groups;
  if (acc == 123.456f) *sink = acc;
}
__global__ void seq(const float4* __restrict__ src, size_t n4, float* sink) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = src[i]; acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) *sink = acc;
}
int main() {
  const size_t bytes = 1ull << 30;
  float* h; CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  for (size_t i = 0; i < bytes / 4; i += 1024) h[i] = 1.f;
  float* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int blocks : {148, 592, 2368}) {
    seq<<<blocks, 256>>>((const float4*)h, bytes / 16 / 4, sink); CK(cudaDeviceSynchronize());
    cudaEventRecord(a); seq<<<blocks, 256>>>((const float4*)h, bytes / 16 / 4, sink); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b);
    printf("sequential float4, %4d blocks: %.2f GB/s\n", blocks, bytes / 4 / ms / 1e6);
  }
  const size_t nlines = bytes / 128;
  for (int span : {1, 8, 16, 32}) {             // floats per request: 4 B (one sector), 32 B, 64 B, 128 B
    for (int blocks : {148, 592, 2368}) {
      const int iters = 64;
      rnd<4><<<blocks, 256>>>(h, nlines, span, iters, sink); CK(cudaDeviceSynchronize());
      cudaEventRecord(a); rnd<4><<<blocks, 256>>>(h, nlines, span, iters, sink); cudaEventRecord(b); CK(cudaDeviceSynchronize());
      cudaEventElapsedTime(&ms, a, b);
      const double warps = blocks * 8.0, reqs = warps * iters * 4 * (32 / (span < 32 ? span : 32));
      printf("random span %3d B x %2d per warp-instr, %4d blocks: %.1f M req/s, %.2f GB/s useful\n", span * 4,
             32 / (span < 32 ? span : 32), blocks, reqs / ms / 1e3, reqs * span * 4 / ms / 1e6);
    }
  }
  return 0;
}
