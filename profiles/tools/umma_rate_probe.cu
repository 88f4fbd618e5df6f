// Probe: issue rate of tcgen05.mma kind::tf32 (M = 128, cta_group::1) for different shared-memory operand layouts.
// One CTA per SM, one thread issues `iters` MMAs back to back into the same accumulator and waits for the commit;
// reports cycles per MMA; `accumulators` > 1 rotates the MMAs over independent TMEM accumulators.  Operand values are irrelevant (shared memory holds small floats).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/umma_rate_probe profiles/tools/umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant {
  uint32_t a_lbo, a_sbo, a_lt, a_step;   // descriptor fields (16-byte units), layout type, byte step per K block
  uint32_t b_lbo, b_sbo, b_lt, b_step;
  uint32_t a_mn, b_mn, N;
};

__global__ void __launch_bounds__(768) rate(Variant V, int iters, int ksteps, int nacc, int extra, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float)((i * 37) % 101) * 0.01f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(s32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp >= 4) {          // optional pollers: warps that spin on a barrier that never completes while the MMAs run
    long long t0 = clock64();
    uint32_t ok = 0, n = 0;
    while (!ok && clock64() - t0 < 400000) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s32(&bar3)), "r"(0u) : "memory");
      ++n;
    }
    if (tid == 128 && blockIdx.x == 0) out[200] = n;
  }
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (V.a_mn << 15) | (V.b_mn << 16) | ((V.N >> 3) << 17) | (8u << 24);
    const uint32_t a0 = s32(sm), b0 = s32(sm) + 96 * 1024;
    uint64_t da[4], db[4];
    for (int kb = 0; kb < 4; ++kb) {
      const uint32_t a = a0 + kb * V.a_step, b = b0 + kb * V.b_step;
      da[kb] = (uint64_t)((a >> 4) & 0x3fffu) | ((uint64_t)(V.a_lbo & 0x3fffu) << 16) | ((uint64_t)(V.a_sbo & 0x3fffu) << 32) |
               (1ull << 46) | ((uint64_t)V.a_lt << 61);
      db[kb] = (uint64_t)((b >> 4) & 0x3fffu) | ((uint64_t)(V.b_lbo & 0x3fffu) << 16) | ((uint64_t)(V.b_sbo & 0x3fffu) << 32) |
               (1ull << 46) | ((uint64_t)V.b_lt << 61);
    }
    const uint32_t astep = 512u / (uint32_t)nacc;
    const uint32_t d0 = tmem, d1 = tmem + (1 % nacc) * astep, d2 = tmem + (2 % nacc) * astep, d3 = tmem + (3 % nacc) * astep;
#define MMA(D, A, B, ACC) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" \
                   ::"r"(D), "l"(A), "l"(B), "r"(idesc), "r"(ACC) : "memory")
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 4) {
      const uint32_t acc = it > 0;
      MMA(d0, da[0], db[0], acc); MMA(d1, da[1], db[1], acc);
      if (extra & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2)) : "memory");
      if (extra & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (extra & 4) {   // a wait that is already satisfied (parity 1 of a fresh barrier)
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(&bar3)), "r"(1u) : "memory");
        if (!ok) __trap();
      }
      MMA(d2, da[2], db[2], acc); MMA(d3, da[3], db[3], acc);
      if (extra & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s32(&bar)), "r"(0u) : "memory");
    }
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 256 * 8));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Named { const char* name; Variant v; } vs[] = {
    // A: 128 x 32 chunk (KC = 32), B: N x 32 chunk
    {"A K-major interleave, B K-major interleave, N=80", {128, 8, 0, 2 * 128 * 16, 80, 8, 0, 2 * 80 * 16, 0, 0, 80}},
    {"A MN SW128_32B,       B K-major interleave, N=80", {256, 32, 1, 1024, 80, 8, 0, 2 * 80 * 16, 1, 0, 80}},
    {"A MN SW128_32B,       B K-major SW128,      N=80", {256, 32, 1, 1024, 1, 64, 2, 32, 1, 0, 80}},
    {"A K-major SW128,      B K-major SW128,      N=80", {1, 64, 2, 32, 1, 64, 2, 32, 0, 0, 80}},
    {"A K-major SW128,      B K-major interleave, N=80", {1, 64, 2, 32, 80, 8, 0, 2 * 80 * 16, 0, 0, 80}},
    {"A K-major SW128,      B K-major SW128,      N=128", {1, 64, 2, 32, 1, 64, 2, 32, 0, 0, 128}},
    {"A K-major SW128,      B K-major SW128,      N=256", {1, 64, 2, 32, 1, 64, 2, 32, 0, 0, 256}},
    {"A MN SW128_32B,       B K-major SW128,      N=256", {256, 32, 1, 1024, 1, 64, 2, 32, 1, 0, 256}},
    {"A MN SW128_32B,       B MN SW128_32B,       N=80", {256, 32, 1, 1024, 256, 32, 1, 1024, 1, 1, 80}},
    {"A MN SW128_32B,       B K-major SW128,      N=16", {256, 32, 1, 1024, 1, 64, 2, 32, 1, 0, 16}},
  };
  for (int vi : {1, 3}) {
    auto& nv = vs[vi];
    for (int threads : {128, 256, 768}) {
      const int iters = 1024, nacc = 1, extra = 0;
      rate<<<148, threads, 200 * 1024>>>(nv.v, iters, 4, nacc, extra, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: failed %s\n", nv.name, cudaGetErrorString(e)); return 2; }
      long long h[148];
      CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
      long long mx = 0; for (auto x : h) mx = x > mx ? x : mx;
      long long npoll = 0;
      CK(cudaMemcpy(&npoll, d + 200, 8, cudaMemcpyDeviceToHost));
      printf("%-55s %2d polling warps (%lld try_waits each in 400k cycles): %7.1f cycles/MMA\n", nv.name, threads / 32 - 4 > 0 ? threads / 32 - 4 : 0,
             threads > 128 ? npoll : 0LL, (double)mx / iters);
    }
  }
  return 0;
}
