// Head 1x1 convolution + decode in one kernel: the `filters = A*(5+C), size = 1,
// activation = linear` convolution that feeds every [yolo] layer (model/cfg/regnetx-600m-fpn.cfg:646-651) is a
// GEMM  raw[cell, o] = sum_c X[c, cell] * Wt[o, c] + bias[o]  and the only contraction on the path, so it runs on
// the 5th-generation tensor cores: tcgen05.mma kind::tf32 (PyTorch's own convolutions use TF32 by default too),
// 128 cells x N channels per CTA, accumulator in TMEM; the epilogue reads the accumulator back with tcgen05.ld,
// adds the bias, applies Decode (model/parser.py:206-235) and writes the rows of the (B, N, 5+C) prediction, so the
// raw head never makes the round trip through HBM.
//
// Operand staging (cp.async; both operands use the un-swizzled "interleave" canonical layouts, in 16-byte units):
//   A = X tile, K-major: unit(m, j = k/4) = 4 consecutive channels of cell m, at j*128 + m        (SBO 8, LBO 128)
//   B = weights, K-major: unit(n, j = k/4) = 4 consecutive input channels of output n, at j*N + n (SBO 8, LBO N)
// One tcgen05.mma consumes K = 8 tf32 values; the K loop runs in chunks of 32 channels through two smem stages: the
// MMAs of a chunk (committed to that stage's mbarrier) run while the other stage is being filled.  The epilogue
// tile reuses the stages; all 8 warps read the accumulator (warps w and w+4 share TMEM lane quadrant w%4).
#include <string.h>

#include "pq_common.cuh"

namespace pq {

constexpr int kHcM = 128;       // cells per CTA = UMMA M
constexpr int kHcKC = 32;       // input channels per smem stage
constexpr int kHcThreads = 256;

struct HeadConvParams {
  const float* x;      // (B, Cin, H, W)
  const float* w;      // (A*(5+C), Cin)
  const float* bias;   // (A*(5+C)) or null
  float* out_dec;      // (B, rows_total, 5+C) or null
  float* out_raw;      // (B, A*(5+C), H, W) or null
  int B, Cin, H, W, A, C;
  int N;               // A*(5+C) rounded up to a multiple of 16
  int tmem_cols;       // power of two >= max(N, 32)
  float stride;
  int64_t rows_total, row_off;
};

__device__ __forceinline__ uint64_t hc_smem_desc(uint32_t smem_addr, uint32_t lbo_units, uint32_t sbo_units) {
  // cute::UMMA::SmemDescriptor: start address [0,14), LBO [16,30), SBO [32,46), version [46,48) = 1 (Blackwell),
  // base offset 0, layout type [61,64) = 0 (no swizzle); addresses and offsets in 16-byte units
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(lbo_units & 0x3fffu) << 16) |
         ((uint64_t)(sbo_units & 0x3fffu) << 32) | (1ull << 46);
}

__global__ void __launch_bounds__(kHcThreads)
head_conv_decode_kernel(const __grid_constant__ HeadConvParams P) {
  extern __shared__ __align__(128) unsigned char hsm[];
  __shared__ __align__(8) uint64_t mma_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int HW = P.H * P.W, ACH = P.A * (5 + P.C), ch = 5 + P.C, N = P.N;
  const int b = blockIdx.y;
  const int cell0 = blockIdx.x * kHcM;
  const int ncell = min(kHcM, HW - cell0);
  // two operand stages (A: kHcKC/4 x 128 units, B: kHcKC/4 x N units of 16 bytes); the epilogue tile reuses them
  const size_t a_bytes = (size_t)(kHcKC / 4) * kHcM * 16, b_bytes = (size_t)(kHcKC / 4) * N * 16;
  const size_t stage_bytes = a_bytes + b_bytes;
  float* tile = reinterpret_cast<float*>(hsm);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init_fence();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                         ((uint32_t)(kHcM >> 4) << 24);
  const float* xb = P.x + (size_t)b * P.Cin * HW;
  const bool vecB = ((P.Cin & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.w) & 15) == 0);
  const int nchunk = (P.Cin + kHcKC - 1) / kHcKC;
  // Operand staging with cp.async (no registers, zero fill outside the tensor): chunk i+1 is in flight while the
  // MMAs of chunk i are issued and run.
  auto stage_chunk = [&](int i) {
    const int kc0 = i * kHcKC, st = i & 1;
    const uint32_t sA = smem_u32(hsm + (size_t)st * stage_bytes);
    const uint32_t sB = smem_u32(hsm + (size_t)st * stage_bytes + a_bytes);
    for (int u = tid; u < (kHcKC / 4) * kHcM; u += kHcThreads) {
      const int j = u >> 7, m = u & (kHcM - 1);            // unit = 4 consecutive channels of one cell
      const int k = kc0 + 4 * j, cell = cell0 + m;
      const float* p = xb + (size_t)k * HW + cell;         // consecutive threads -> consecutive cells: coalesced
      const uint32_t d = sA + (uint32_t)(j * kHcM + m) * 16u;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = (cell < HW) && (k + e < P.Cin);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                     ::"r"(d + 4u * e), "l"(ok ? p + (size_t)e * HW : xb), "r"(ok ? 4u : 0u) : "memory");
      }
    }
    for (int u = tid; u < (kHcKC / 4) * N; u += kHcThreads) {
      const int j = u / N, n = u - j * N;
      const int k = kc0 + 4 * j;
      const float* p = P.w + (size_t)n * P.Cin + k;
      const uint32_t d = sB + (uint32_t)(j * N + n) * 16u;
      if (vecB) {
        const bool ok = (n < ACH) && (k < P.Cin);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;"
                     ::"r"(d), "l"(ok ? p : P.w), "r"(ok ? 16u : 0u) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = (n < ACH) && (k + e < P.Cin);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                       ::"r"(d + 4u * e), "l"(ok ? p + e : P.w), "r"(ok ? 4u : 0u) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_chunk(0);
  for (int i = 0; i < nchunk; ++i) {
    const int st = i & 1;
    if (i + 1 < nchunk) {
      // the MMAs that read the other stage (chunk i-1) must have finished before it is refilled
      if (i >= 1) mbar_wait(&mma_done[st ^ 1], (uint32_t)(((i - 1) >> 1) & 1));
      stage_chunk(i + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // chunk i has landed, chunk i+1 may be in flight
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    fence_async_smem();                       // writes of this thread -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(hsm + (size_t)st * stage_bytes), b0 = a0 + (uint32_t)a_bytes;
#pragma unroll
      for (int kb = 0; kb < kHcKC / 8; ++kb) {
        const uint64_t da = hc_smem_desc(a0 + (uint32_t)(2 * kb) * (uint32_t)kHcM * 16u, (uint32_t)kHcM, 8u);
        const uint64_t db = hc_smem_desc(b0 + (uint32_t)(2 * kb) * (uint32_t)N * 16u, (uint32_t)N, 8u);
        const uint32_t accumulate = (i > 0 || kb > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
      }
      // the commit makes the mbarrier track completion of everything issued so far (and implies the
      // before_thread_sync fence)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                   ::"r"(smem_u32(&mma_done[st])) : "memory");
    }
  }
  // the last commit covers every MMA issued before it
  {
    const int last = nchunk - 1;
    mbar_wait(&mma_done[last & 1], (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();                            // nobody is still staging operands: the tile may reuse the stages

  // ---- epilogue: accumulator row = cell, column = output channel; warps w and w+4 share TMEM lane quadrant
  // w%4 and split the columns ------------------------------------------------------------------------------
  const int ST = ACH | 1;
  {
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;                          // TMEM lane == row of the tile
    const int cell = cell0 + r;
    const int cy = cell / P.W, cx = cell - cy * P.W;
    const float gx = (float)cx + 0.5f, gy = (float)cy + 0.5f;
    const int nblk = N / 16, blk_lo = half ? (nblk + 1) / 2 : 0, blk_hi = half ? nblk : (nblk + 1) / 2;
    for (int blk = blk_lo; blk < blk_hi; ++blk) {
      const int c0 = blk * 16;
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (r < ncell) {
        int k = c0 % ch;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + i;
          if (c < ACH) {
            float raw = __uint_as_float(v[i]);
            if (P.bias) raw = PQ_ADD(raw, __ldg(P.bias + c));
            if (P.out_raw) P.out_raw[((size_t)b * ACH + c) * HW + cell] = raw;
            float o;
            if (k < 4) {
              const float e = expf(raw);
              const float g = (k & 1) ? gy : gx;
              o = PQ_MUL((k < 2) ? PQ_SUB(g, e) : PQ_ADD(g, e), P.stride);      // decode_coord
            } else {
              o = __frcp_rn(PQ_ADD(1.0f, expf(-raw)));                          // sigmoidf_
            }
            tile[r * ST + c] = o;
          }
          if (++k == ch) k = 0;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (P.out_dec) {
    float* dst = P.out_dec + ((size_t)b * P.rows_total + P.row_off + (size_t)cell0 * P.A) * ch;
    const int n = ncell * ACH;
    if (ST == ACH && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0)) {
      const float4* t4 = reinterpret_cast<const float4*>(tile);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int e = tid; e < (n >> 2); e += kHcThreads) d4[e] = t4[e];
    } else {
      for (int r = warp; r < ncell; r += kHcThreads / 32) {
        const float* trow = tile + r * ST;
        float* drow = dst + (size_t)r * ACH;
        for (int c = lane; c < ACH; c += 32) drow[c] = trow[c];
      }
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

}  // namespace pq

extern "C" int pqdet_head_conv_decode(const float* x, const float* weight, const float* bias, float* out_decoded,
                                      float* out_raw, int B, int Cin, int H, int W, int A, int C, float stride,
                                      int64_t out_rows_total, int64_t out_row_offset, int device, void* stream) {
  using namespace pq;
  if (B < 0 || Cin < 1 || H < 1 || W < 1 || A < 1 || C < 0) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;
  if (!x || !weight || (!out_decoded && !out_raw)) return PQDET_ERR_INVALID_ARG;
  const int ACH = A * (5 + C);
  if (ACH > 256 || B > 65535) return PQDET_ERR_UNSUPPORTED;
  if (out_decoded && (out_row_offset < 0 || out_row_offset + (int64_t)H * W * A > out_rows_total))
    return PQDET_ERR_INVALID_ARG;
  HeadConvParams P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.w = weight; P.bias = bias; P.out_dec = out_decoded; P.out_raw = out_raw;
  P.B = B; P.Cin = Cin; P.H = H; P.W = W; P.A = A; P.C = C;
  P.N = (ACH + 15) / 16 * 16;
  P.tmem_cols = 32;
  while (P.tmem_cols < P.N) P.tmem_cols <<= 1;
  P.stride = stride; P.rows_total = out_rows_total; P.row_off = out_row_offset;
  PQ_ENTER(device);
  const size_t stages = 2 * ((size_t)(kHcKC / 4) * kHcM * 16 + (size_t)(kHcKC / 4) * P.N * 16);
  const size_t tile_bytes = (size_t)kHcM * (ACH | 1) * sizeof(float);
  const size_t smem = stages > tile_bytes ? stages : tile_bytes;
  if (smem > 200 * 1024) return PQDET_ERR_UNSUPPORTED;
  PQ_CUDA(cudaFuncSetAttribute(head_conv_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((H * W + kHcM - 1) / kHcM, B);
  head_conv_decode_kernel<<<grid, kHcThreads, smem, (cudaStream_t)stream>>>(P);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
