// Probe: tcgen05.mma kind::tf32 with the A operand MN-major.  CUTLASS (sm100_common.inl:92) says "for mn-major tf32
// operands, SW128_32B is the only available smem layout": 128-byte rows along MN, 4 K-rows per atom, 32-byte chunks
// XOR-swizzled with the row index (Swizzle<2,5,2>), layout type 1.  Staged (mode 1) by threads with the swizzle applied
// by hand and (mode 2) by 2-D TMA tensor-map loads with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; mode 0 = K-major control;
// modes 3/4 = plain SWIZZLE_128B (16-byte atoms), which yields zeros for tf32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/umma_mn_probe profiles/tools/umma_mn_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int M = 128, KC = 32, N = 80;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 22)) __trap();
  }
}

__global__ void __launch_bounds__(128) probe(const float* X, int ldx, int cell0, const float* Wt, float* out,
                                             const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2, int mode, uint32_t lbo,
                                             uint32_t sbo, uint32_t layout_type) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar_ld, bar_mma;
  __shared__ uint32_t tmem_s;
  unsigned char* sA = sm;                    // 16 KB, 1024-aligned
  unsigned char* sB = sm + 16384;            // KC/4 * N units
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_s)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar_ld)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  // B: K-major interleave: unit (n, j) at j*N + n
  for (int u = tid; u < (KC / 4) * N; u += 128) {
    const int j = u / N, n = u - j * N;
    float4 v = *reinterpret_cast<const float4*>(Wt + (size_t)n * KC + 4 * j);
    *reinterpret_cast<float4*>(sB + (size_t)(j * N + n) * 16) = v;
  }
  if (mode == 0) {          // K-major interleave control: unit (m, j) at j*128 + m holds channels 4j..4j+3 of cell m
    for (int e = tid; e < KC * M; e += 128) {
      const int r = e / M, m = e % M;
      *reinterpret_cast<float*>(sA + ((r >> 2) * M + m) * 16 + (r & 3) * 4) = X[(size_t)r * ldx + cell0 + m];
    }
  } else if (mode == 1 || mode == 3) {
    for (int e = tid; e < KC * M; e += 128) {
      const int r = e / M, m = e % M;
      const int g = m >> 5, w = m & 31;
      const uint32_t off = mode == 1 ? g * (KC * 128) + r * 128 + (((w >> 3) ^ (r & 3)) << 5) + (w & 7) * 4
                                     : g * (KC * 128) + r * 128 + (((w >> 2) ^ (r & 7)) << 4) + (w & 3) * 4;
      *reinterpret_cast<float*>(sA + off) = X[(size_t)r * ldx + cell0 + m];
    }
  } else if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar_ld)), "r"(16384u) : "memory");
    for (int g = 0; g < 4; ++g) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(s32(sA + g * (KC * 128))), "l"(mode == 2 ? &tmap : &tmap2), "r"(cell0 + 32 * g), "r"(0), "r"(s32(&bar_ld)) : "memory");
    }
  }
  if (mode == 2 || mode == 4) mwait(&bar_ld, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((mode ? 1u : 0u) << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int kb = 0; kb < KC / 8; ++kb) {
      const uint32_t a = s32(sA) + (mode ? kb * 1024 : 2 * kb * M * 16), b = s32(sB) + (2 * kb) * N * 16;
      const uint64_t da = (uint64_t)((a >> 4) & 0x3fffu) | ((uint64_t)(lbo & 0x3fffu) << 16) | ((uint64_t)(sbo & 0x3fffu) << 32) |
                          (1ull << 46) | ((uint64_t)layout_type << 61);
      const uint64_t db = (uint64_t)((b >> 4) & 0x3fffu) | ((uint64_t)(N & 0x3fffu) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
      const uint32_t acc = kb > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar_mma)) : "memory");
  }
  mwait(&bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(size_t)(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int LDX = 1024, CELL0 = 256;
  std::vector<float> hX((size_t)KC * LDX), hW((size_t)N * KC);
  srand(1);
  for (auto& v : hX) v = (float)(rand() % 2001 - 1000) / 1000.0f;
  for (auto& v : hW) v = (float)(rand() % 2001 - 1000) / 1000.0f;
  float *dX, *dW, *dO;
  CK(cudaMalloc(&dX, hX.size() * 4)); CK(cudaMalloc(&dW, hW.size() * 4)); CK(cudaMalloc(&dO, (size_t)M * N * 4));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice));
  std::vector<double> ref((size_t)M * N);
  double scale = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double s = 0, a = 0;
    for (int k = 0; k < KC; ++k) { double p = (double)hX[(size_t)k * LDX + CELL0 + m] * hW[(size_t)n * KC + k]; s += p; a += fabs(p); }
    ref[(size_t)m * N + n] = s; scale = fmax(scale, a);
  }
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
  cuuint64_t dims[2] = {(cuuint64_t)LDX, (cuuint64_t)KC}, strides[1] = {(cuuint64_t)LDX * 4};
  cuuint32_t box[2] = {32, KC}, estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUtensorMap tmap2; memset(&tmap2, 0, sizeof(tmap2));
  CUresult r2 = ((EncodeFn)fn)(&tmap2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d %d\n", (int)r, (int)r2);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  struct V { int mode; uint32_t lbo, sbo, lt; const char* name; } vs[] = {
    {0, 128, 8, 0, "K-major interleave control"},
    {1, KC * 8, 32, 1, "threads SW128_32B LBO=group SBO=512B"},
    {2, KC * 8, 32, 1, "TMA     SW128_32B LBO=group SBO=512B"},
    {1, 32, KC * 8, 1, "threads SW128_32B LBO=512B SBO=group"},
    {2, 32, KC * 8, 1, "TMA     SW128_32B LBO=512B SBO=group"},
    {3, KC * 8, 64, 2, "threads SW128 (16B atoms)"},
    {4, KC * 8, 64, 2, "TMA     SW128 (16B atoms)"},
  };
  std::vector<float> hO((size_t)M * N);
  for (auto& v : vs) {
    CK(cudaMemset(dO, 0, (size_t)M * N * 4));
    probe<<<1, 128, 32768>>>(dX, LDX, CELL0, dW, dO, tmap, tmap2, v.mode, v.lbo, v.sbo, v.lt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-50s launch failed: %s\n", v.name, cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0; int bad = 0;
    for (size_t i = 0; i < hO.size(); ++i) { double d = fabs(hO[i] - ref[i]); worst = fmax(worst, d); bad += d > 4e-3 * scale; }
    printf("%-50s max|err| %.3e (scale %.2f) bad %d/%d  out[0..2]=%g %g %g ref %g %g %g\n", v.name, worst, scale, bad, M * N,
           hO[0], hO[1], hO[2], ref[0], ref[1], ref[2]);
  }
  return 0;
}
