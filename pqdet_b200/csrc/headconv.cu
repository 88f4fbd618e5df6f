// Head 1x1 convolution + decode in one kernel: the `filters = A*(5+C), size = 1,
// activation = linear` convolution that feeds every [yolo] layer (model/cfg/regnetx-600m-fpn.cfg:646-651) is a
// GEMM  raw[cell, o] = sum_c X[c, cell] * Wt[o, c] + bias[o]  and the only contraction on the path, so it runs on
// the 5th-generation tensor cores: tcgen05.mma kind::tf32 (PyTorch's own convolutions use TF32 by default too),
// 128 cells x N channels per tile, accumulator in TMEM; the epilogue reads the accumulator back with tcgen05.ld,
// adds the bias, applies Decode (model/parser.py:206-235) and writes the rows of the (B, N, 5+C) prediction, so the
// raw head never makes the round trip through HBM.
//
// Two kernels.
// `head_conv_decode_ws_kernel` (the fast path: H*W a multiple of 128, weights fit in shared memory) is persistent and
// warp specialised: one CTA per SM keeps the whole weight matrix resident in shared memory, warp 0 streams X tiles
// with TMA tensor-map loads (the NCHW planes are MN-major for this GEMM; tf32 takes an MN-major operand only in the
// "128-byte swizzle, 32-byte atom" layout = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, UMMA layout type 1) through an
// mbarrier ring, warp 1 issues the MMAs into one of two TMEM accumulators, and 4*wq epilogue warps decode the other
// accumulator and hand the finished (128 cells x A x (5+C)) tile - one contiguous run of the output - to a TMA bulk
// store.
// `head_conv_decode_kernel` is the general kernel (any shape), one tile per CTA.  Operand staging with cp.async, both
// operands in the un-swizzled "interleave" canonical layouts, in 16-byte units:
//   A = X tile, K-major: unit(m, j = k/4) = 4 consecutive channels of cell m, at j*128 + m        (SBO 8, LBO 128)
//   B = weights, K-major: unit(n, j = k/4) = 4 consecutive input channels of output n, at j*N + n (SBO 8, LBO N)
// One tcgen05.mma consumes K = 8 tf32 values; the K loop runs in chunks of 32 channels through two smem stages: the
// MMAs of a chunk (committed to that stage's mbarrier) run while the other stage is being filled.  The epilogue
// tile reuses the stages; all 8 warps read the accumulator (warps w and w+4 share TMEM lane quadrant w%4).
#include <stdlib.h>
#include <string.h>

#include "pq_common.cuh"

namespace pq {

constexpr int kHcM = 128;       // cells per CTA = UMMA M
#ifndef PQ_HC_KC
#define PQ_HC_KC 32
#endif
constexpr int kHcKC = PQ_HC_KC;  // input channels per smem stage of the general kernel
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
  int col_group;       // 0: the whole decoded tile is staged at once; else columns per epilogue pass (multiple of 32)
  float stride;
  int64_t rows_total, row_off;
  // HITS mode (see HeadConvWsParams)
  float* rec;
  int32_t* rec_count;
  int rec_cap;
  float logit_lo;
};

__device__ __forceinline__ uint64_t hc_smem_desc(uint32_t smem_addr, uint32_t lbo_units, uint32_t sbo_units) {
  // cute::UMMA::SmemDescriptor: start address [0,14), LBO [16,30), SBO [32,46), version [46,48) = 1 (Blackwell),
  // base offset 0, layout type [61,64) = 0 (no swizzle); addresses and offsets in 16-byte units
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(lbo_units & 0x3fffu) << 16) |
         ((uint64_t)(sbo_units & 0x3fffu) << 32) | (1ull << 46);
}

// One tf32 MMA, descriptors passed as 32-bit halves: only the low word (start address, LBO) changes between the MMAs
// of a tile, so the issue loop is one add per operand.  (Measured, profiles/tools/umma_rate_probe.cu: rebuilding the
// 64-bit descriptors per MMA costs ~190 cycles of dependent uniform-datapath arithmetic, 4x the MMA itself.)
__device__ __forceinline__ void hc_mma_tf32(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

// The thresholding epilogue shared by both kernels (pqdet_head_conv_hits): for anchor `a` of this warp's 32 cells read
// back only the objectness column, and - only when one of the cells passes the logit-space prefilter - the anchor's
// 4 box + C class columns; the passing cells append a record [row, objectness, 4 box, C class raw values].
// tacc = TMEM address of the warp's lane quadrant, column 0; col_limit = allocated accumulator columns.
template <typename BiasFn>
__device__ __forceinline__ void hc_emit_hits(uint32_t tacc, int a, int ch, int C, int col_limit, BiasFn bias_at,
                                             bool in_level, int lane, int64_t row, int img, float* rec,
                                             int32_t* rec_count, int rec_cap, float logit_lo) {
  const int c0 = a * ch;
  uint32_t vo;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(vo) : "r"(tacc + (uint32_t)(c0 + 4)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const float xo = PQ_ADD(__uint_as_float(vo), bias_at(c0 + 4));
  const bool pass = in_level && xo > logit_lo;
  const unsigned pm = __ballot_sync(PQ_FULL, pass);
  if (pm == 0u) return;                                      // warp-uniform
  int slot0 = 0;
  if (lane == 0) slot0 = atomicAdd(rec_count + img, __popc(pm));
  slot0 = __shfl_sync(PQ_FULL, slot0, 0);
  const int slot = slot0 + __popc(pm & ((1u << lane) - 1u));
  const bool store = pass && slot < rec_cap;
  float* rc = rec + ((size_t)img * rec_cap + (store ? slot : 0)) * (size_t)(6 + C);
  uint32_t v4[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v4[0]), "=r"(v4[1]), "=r"(v4[2]), "=r"(v4[3]) : "r"(tacc + (uint32_t)c0) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (store) {
    rc[0] = __int_as_float((int)row);
    rc[1] = xo;
#pragma unroll
    for (int i = 0; i < 4; ++i) rc[2 + i] = PQ_ADD(__uint_as_float(v4[i]), bias_at(c0 + i));
  }
  for (int k = 5; k < ch; k += 8) {                          // class columns, 8 at a time (reads may run past the anchor)
    if (c0 + k + 8 <= col_limit) {
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(tacc + (uint32_t)(c0 + k)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (store) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (k + i < ch) rc[1 + k + i] = PQ_ADD(__uint_as_float(v[i]), bias_at(c0 + k + i));
      }
    } else {                                                 // the last columns of a 256-column accumulator (COCO)
      for (int i = 0; i < 8 && k + i < ch; ++i) {
        uint32_t v1;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v1) : "r"(tacc + (uint32_t)(c0 + k + i)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (store) rc[1 + k + i] = PQ_ADD(__uint_as_float(v1), bias_at(c0 + k + i));
      }
    }
  }
}

template <bool HITS>
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
    // (the staging loops are what this kernel issues most: kept to pointer increments - a thread's cell is the same
    // in every iteration of the X loop (256 threads, 128 cells), and the W loop walks n inside j, no division)
    {
      const int m = tid & (kHcM - 1), cell = cell0 + m;
      const bool in = cell < HW;
      const bool whole = kc0 + kHcKC <= P.Cin;               // no zero-filled channel tail in this chunk
      const float* p = xb + (size_t)(kc0 + 4 * (tid >> 7)) * HW + cell;   // consecutive threads -> consecutive cells
      uint32_t d = sA + (uint32_t)((tid >> 7) * kHcM + m) * 16u;
      const size_t pstep = (size_t)(4 * (kHcThreads >> 7)) * HW;
#pragma unroll
      for (int j = tid >> 7; j < kHcKC / 4; j += kHcThreads >> 7, p += pstep, d += (uint32_t)((kHcThreads >> 7) * kHcM) * 16u) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = in && (whole || kc0 + 4 * j + e < P.Cin);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                       ::"r"(d + 4u * e), "l"(ok ? p + (size_t)e * HW : xb), "r"(ok ? 4u : 0u) : "memory");
        }
      }
    }
    for (int j = 0; j < kHcKC / 4; ++j) {
      const int k = kc0 + 4 * j;
      for (int n = tid; n < N; n += kHcThreads) {
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
      const uint64_t da = hc_smem_desc(a0, (uint32_t)kHcM, 8u), db = hc_smem_desc(b0, (uint32_t)N, 8u);
      uint32_t a_lo = (uint32_t)da, b_lo = (uint32_t)db;
      const uint32_t a_hi = (uint32_t)(da >> 32), b_hi = (uint32_t)(db >> 32);
#pragma unroll
      for (int kb = 0; kb < kHcKC / 8; ++kb) {
        hc_mma_tf32(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc, (i > 0 || kb > 0) ? 1u : 0u);
        a_lo += 2u * (uint32_t)kHcM;       // two 16-byte units of K further, in 16-byte address units
        b_lo += 2u * (uint32_t)N;
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
  if (HITS) {
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int a = half; a < P.A; a += 2)
      hc_emit_hits(tacc, a, ch, P.C, P.tmem_cols, [&](int c) { return P.bias ? __ldg(P.bias + c) : 0.0f; }, r < ncell,
                   lane, P.row_off + (int64_t)(cell0 + r) * P.A + a, b, P.rec, P.rec_count, P.rec_cap, P.logit_lo);
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;                          // TMEM lane == row of the tile
    const int cell = cell0 + r;
    const int cy = cell / P.W, cx = cell - cy * P.W;
    const float gx = (float)cx + 0.5f, gy = (float)cy + 0.5f;
    float* dst = P.out_dec ? P.out_dec + ((size_t)b * P.rows_total + P.row_off + (size_t)cell0 * P.A) * ch : nullptr;
    // The decoded tile is staged either whole (it then leaves as one contiguous run) or, where 128 x A(5+C) floats
    // would need more shared memory than the operand stages (255 channels: 130 KB, one CTA per SM), in passes of
    // col_group columns: decode a column group, barrier, copy its segment of every row out, barrier.
    const int GC = P.col_group ? P.col_group : N;
    const int STg = P.col_group ? (GC | 1) : ST;
    for (int g0 = 0; g0 < N; g0 += GC) {
      const int gb = g0 / 16, ge = min(N, g0 + GC) / 16;            // 16-column blocks of this pass
      const int nb = ge - gb, blk_lo = gb + (half ? (nb + 1) / 2 : 0), blk_hi = half ? ge : gb + (nb + 1) / 2;
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
          // two blocks of 8 columns, each with its exp / reciprocal chains interleaved (decode_block8: bit-identical to
          // the scalar functions); a block may run into the padding columns behind ACH, which decode_block8 does not
          // store
          int k = c0 % ch;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cb = c0 + 8 * h;
            if (cb < ACH) {
              float raw[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = cb + i;
                raw[i] = __uint_as_float(v[8 * h + i]);
                if (P.bias && c < ACH) raw[i] = PQ_ADD(raw[i], __ldg(P.bias + c));
                if (P.out_raw && c < ACH) P.out_raw[((size_t)b * ACH + c) * HW + cell] = raw[i];
              }
              decode_block8(raw, k, cb, ACH, ch, gx, gy, P.stride, tile + r * STg + (cb - g0));
            }
            k += 8;
            if (k >= ch) k -= ch;
            if (k >= ch) k -= ch;
          }
        }
      }
      if (!P.col_group) break;                              // staged whole: stored below
      __syncthreads();
      if (dst) {
        const int gc = min(GC, ACH - g0);                   // columns of this pass that exist
        for (int rr = warp; rr < ncell; rr += kHcThreads / 32) {
          const float* trow = tile + rr * STg;
          float* drow = dst + (size_t)rr * ACH + g0;
          for (int c = lane; c < gc; c += 32) drow[c] = trow[c];
        }
      }
      __syncthreads();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (P.out_dec && !(P.col_group && !HITS)) {
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


// ---------------------------------------------------------------------------------------------------------------
// Persistent, warp-specialised fast path
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWsMaxStages = 12;

struct HeadConvWsParams {
  const float* w;
  const float* bias;
  float* out_dec;
  float* out_raw;
  int B, Cin, HW, Wd, A, C;
  int N;               // A*(5+C) rounded up to 16
  int KC;              // channels per stage: 32, 16 or 8 (divides Cin)
  int stages;
  int wq;              // epilogue warps per TMEM lane quadrant
  int tile_bufs;       // 1 or 2 staging tiles for the output
  int all_bulk;        // every output tile is full and 16-byte aligned: all of them leave as bulk stores
  // Anchor split (weights of all anchors too large to stay resident: COCO at Cin 176): this launch computes ONE anchor
  // (A = 1, weights / bias of that anchor) of a head with A_out anchors; its rows are A_out rows apart in the
  // prediction, its raw planes at plane a_off * (5+C) of the raw head.  Otherwise A_out = A, a_off = 0.
  int A_out, a_off;
  int slice;           // 1: the staging tiles hold ONE anchor's rows of a tile (A * (5+C) * 512 bytes would not fit beside
                       // the resident weights: COCO's 255 channels); the epilogue then works anchor by anchor
  int units;           // 1: anchor-aligned work units in the epilogue (needs ACH + 7 <= buf_cols), 0: 8-column blocks
  int buf_cols;        // TMEM column stride between the two accumulators
  int tiles_per_img, ntiles;
  uint32_t magic_w;    // ceil(2^32 / W): cell / W by multiply-high (0 when W == 1)
  float stride;
  int64_t rows_total, row_off;
  // HITS mode (pqdet_head_conv_hits): instead of the decoded tile the epilogue appends one record per row whose
  // objectness logit passes the conservative prefilter: [row (int bits), objectness, 4 box, C class raw values]
  float* rec;          // (B, rec_cap, 6 + C)
  int32_t* rec_count;  // (B) appended records (may exceed rec_cap: the image then overflowed)
  int rec_cap;
  float logit_lo;
};

// Body of the persistent kernel: this CTA is number `cta` of the `ncta` that share the level described by P.
template <bool WANT_RAW, bool HITS = false, bool SLICE = false>
__device__ __forceinline__ void head_conv_ws_body(const HeadConvWsParams& P, const CUtensorMap* tmap_x, const int cta,
                                                  const int ncta) {
  extern __shared__ __align__(1024) unsigned char hsm_ws[];
  __shared__ __align__(8) uint64_t full_bar[kWsMaxStages], empty_bar[kWsMaxStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t utab[64];        // work units of one cell row: first channel | channel count << 16 (0 = box)
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int ACH = P.A * (5 + P.C), ch = 5 + P.C, N = P.N, KC = P.KC, S = P.stages;
  const int nchunk = P.Cin / KC;
  const uint32_t stage_bytes = (uint32_t)KC * 512u;
  // dynamic smem: [X ring, 1024-byte aligned][W: Cin/4 x N units][1 or 2 output tiles: 128 x ACH floats][bias]
  const uint32_t base = (smem_u32(hsm_ws) + 1023u) & ~1023u;
  unsigned char* aligned = hsm_ws + (base - smem_u32(hsm_ws));
  const uint32_t sX = base;
  const uint32_t sW = base + (uint32_t)S * stage_bytes;
  float* tile0 = reinterpret_cast<float*>(aligned + (size_t)S * stage_bytes + (size_t)P.Cin * N * 4);
  // + 16 bytes: a tile is written at the output's 16-byte phase; slice mode: one anchor's rows (128 x (5+C))
  const int tile_stride = (P.slice ? kHcM * ch : kHcM * ACH) + 4;
  float* sbias = tile0 + P.tile_bufs * tile_stride;         // N + 8 floats, zero beyond ACH or without a bias
  const int n_epi = 4 * P.wq * 32;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)(2 * P.buf_cols)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 4 * P.wq);
    mbar_init(&acc_empty[1], 4 * P.wq);
    mbar_init_fence();
  }
  // the weights stay resident: K-major, unit(n, j = k/4) at j*N + n, rows n >= ACH zero
  for (int u = tid; u < (P.Cin / 4) * N; u += blockDim.x) {
    const int j = u / N, n = u - j * N;
    const bool ok = n < ACH;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;"
                 ::"r"(sW + (uint32_t)u * 16u), "l"(ok ? P.w + (size_t)n * P.Cin + 4 * j : P.w), "r"(ok ? 16u : 0u)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int c = tid; c < N + 8; c += blockDim.x) sbias[c] = (P.bias && c < ACH) ? P.bias[c] : 0.0f;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  fence_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ---- producer: one thread keeps the ring full -------------------------------------------------------------
    {
      uint32_t s = 0, ph = 0;
      for (int t = cta; t < P.ntiles; t += ncta) {
        const int b = t / P.tiles_per_img, cell0 = (t - b * P.tiles_per_img) * kHcM;
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(&empty_bar[s], ph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[s], stage_bytes);
            const uint32_t dst = sX + s * stage_bytes;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              tma_load_2d(dst + (uint32_t)g * (uint32_t)KC * 128u, tmap_x, cell0 + 32 * g, b * P.Cin + c * KC, &full_bar[s]);
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (warp-uniform loop, one elected lane issues) --------------------------------------------------
    {
      // D = F32, A = B = TF32, A MN-major (bit 15), B K-major, N, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) |
                             ((uint32_t)(kHcM >> 4) << 24);
      // A: SW128 with 32-byte atoms (layout type 1): rows of 128 bytes, 4-row atoms 512 bytes apart (SBO 32 units),
      // 32-cell groups KC*128 bytes apart (LBO); one K step (8 rows) = 1024 bytes = 64 units
      const uint64_t da0 = (uint64_t)((sX >> 4) & 0x3fffu) | ((uint64_t)(((uint32_t)KC * 8u) & 0x3fffu) << 16) |
                           ((uint64_t)32u << 32) | (1ull << 46) | (1ull << 61);
      const uint64_t db0 = hc_smem_desc(sW, (uint32_t)N, 8u);
      const uint32_t a_hi = (uint32_t)(da0 >> 32), b_hi = (uint32_t)(db0 >> 32);
      const uint32_t stage_units = stage_bytes >> 4;
      const int kpc = KC / 8;
      uint32_t s = 0, ph = 0, a_stage = (uint32_t)da0, tl = 0;
      for (int t = cta; t < P.ntiles; t += ncta, ++tl) {
        const uint32_t buf = tl & 1u;
        mbar_wait(&acc_empty[buf], ((tl >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)P.buf_cols;
        uint32_t b_lo = (uint32_t)db0, accumulate = 0;
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            uint32_t a_lo = a_stage, bb = b_lo, acc = accumulate;
            for (int kb = 0; kb < kpc; ++kb) {
              hc_mma_tf32(d_tmem, a_lo, a_hi, bb, b_hi, idesc, acc);
              acc = 1;
              a_lo += 64u;
              bb += 2u * (uint32_t)N;
            }
            // frees the stage once the MMAs that read it are done
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(&empty_bar[s])) : "memory");
          }
          __syncwarp();
          accumulate = 1;
          b_lo += 2u * (uint32_t)N * (uint32_t)kpc;
          a_stage += stage_units;
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; a_stage = (uint32_t)da0; }
        }
        if (elect_one())
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                       ::"r"(smem_u32(&acc_full[buf])) : "memory");
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp%4 = TMEM lane quadrant; the wq warps of a quadrant take the 8-column blocks round robin ----
    const int q = warp & 3, jq = (warp - 4) >> 2;
    const int etid = tid - 128;
    const int r = q * 32 + lane;
    const int nblk = (ACH + 7) / 8;
    const int k_first = (jq * 8) % ch;                      // channel-within-anchor of this warp's first block
    const int k_step = (P.wq * 8) % ch;
    uint32_t tl = 0;
    // anchor-aligned work units (see decode_levels_tma_kernel): per anchor its objectness / class channels in groups
    // of usz <= 8, numbered first, then the 4 box channels of every anchor; the same for every tile, so the table is
    // worked out once (no division in the tile loop)
    // (balanced: the first `big` groups of an anchor hold usz channels, the others usz - 1; 81 = 4 x 8 + 7 x 7)
    const int ns = ch - 4, grp = (ns + 7) / 8, usz = (ns + grp - 1) / grp, nsig = P.A * grp, nunit = nsig + P.A;
    const int big = ns - grp * (usz - 1);
    if (P.units && !HITS) {
      for (int u = etid; u < nunit; u += 4 * P.wq * 32) {
        uint32_t e;
        if (u >= nsig) {
          e = (uint32_t)((u - nsig) * ch);
        } else {
          const int a = u / grp, j = u - a * grp;
          const int k = 4 + (j < big ? j * usz : big * usz + (j - big) * (usz - 1));
          e = (uint32_t)(a * ch + k) | ((uint32_t)(j < big ? usz : usz - 1) << 16);
        }
        utab[u] = e;
      }
      epi_bar_sync(4 * P.wq * 32);
    }
    uint32_t sl = 0;                                        // slice mode: running slice count (staging buffer = sl & 1)
    // per-tile bookkeeping without divisions: (image, tile within image) advance by ncta with carries
    int b = cta / P.tiles_per_img, ti = cta - b * P.tiles_per_img;
    const int step_b = ncta / P.tiles_per_img, step_t = ncta - step_b * P.tiles_per_img;
    for (int t = cta; t < P.ntiles; t += ncta, ++tl) {
      const uint32_t buf = tl & 1u;
      const int cell0 = ti * kHcM;
      const int cell = cell0 + r;
      const int ncell = min(kHcM, P.HW - cell0);          // the last tile of a level may be partial (TMA zero-fills)
      const bool in_level = r < ncell;
      const int cy = P.magic_w ? (int)__umulhi((uint32_t)cell, P.magic_w) : cell, cx = cell - cy * P.Wd;
      const float gx = (float)cx + 0.5f, gy = (float)cy + 0.5f;
      float* dst = P.out_dec + ((size_t)b * P.rows_total + P.row_off + (size_t)cell0 * P.A_out + P.a_off) * ch;
      const size_t raw_base = ((size_t)b * P.A_out + P.a_off) * ch * P.HW + cell;   // (A_out == A: b * ACH * HW + cell)
      ti += step_t; b += step_b;
      if (ti >= P.tiles_per_img) { ti -= P.tiles_per_img; ++b; }
      mbar_wait(&acc_full[buf], (tl >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (HITS) {
        // Threshold in the epilogue (second half): only the objectness column of every anchor is read
        // back; a warp fetches the anchor's other 4 + C columns only when one of its 32 cells passes the prefilter,
        // and only those cells leave a record.  Nothing of size B x N is written.
        const uint32_t tacc = tmem_base + buf * (uint32_t)P.buf_cols + ((uint32_t)(q * 32) << 16);
        const int img = t / P.tiles_per_img;                 // (b was already advanced to the next tile above)
        for (int a = jq; a < P.A; a += P.wq)
          hc_emit_hits(tacc, a, ch, P.C, P.buf_cols, [&](int c) { return sbias[c]; }, in_level, lane,
                       P.row_off + (int64_t)cell * P.A + a, img, P.rec, P.rec_count, P.rec_cap, P.logit_lo);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
      // the bulk store that last used this staging tile must have read it before it is overwritten
      // the staging tile starts at the same offset inside a 16-byte unit as its destination, so that everything
      // but at most 3 floats at either end can leave as one bulk store even when the rows of this image are not
      // 16-byte aligned in the concatenated prediction (608 x 608: 22743 rows per image)
      const int mis = (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);
      float* tile = tile0 + (P.tile_bufs == 2 ? (int)buf * tile_stride : 0) + mis;
      if (etid == 0) {
        // with two staging tiles one bulk store may still be reading the OTHER tile - but only if every tile leaves
        // as a bulk store (otherwise the group count no longer tells which tile a pending group reads)
        if (P.tile_bufs == 2 && P.all_bulk) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        // (all_bulk: every tile commits exactly one group, so "at most one pending" = the other tile's)
      }
      epi_bar_sync(n_epi);
      const uint32_t tacc = tmem_base + buf * (uint32_t)P.buf_cols + ((uint32_t)(q * 32) << 16);
      float* trow0 = tile + r * ACH;
      // one work unit of this thread's cell: TMEM -> + bias -> (raw planes) -> decode into the staging row at trow
      auto do_unit = [&](const uint32_t ue, float* __restrict__ trow) {
        const int c0 = (int)(ue & 0xffffu), cnt = (int)(ue >> 16);
        if (cnt == 0) {
          uint32_t v[4];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(tacc + (uint32_t)c0) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float raw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) raw[i] = PQ_ADD(__uint_as_float(v[i]), sbias[c0 + i]);
          if (WANT_RAW && in_level) {
#pragma unroll
            for (int i = 0; i < 4; ++i) P.out_raw[raw_base + (size_t)(c0 + i) * P.HW] = raw[i];
          }
          decode_box4(raw, gx, gy, P.stride, trow + c0);
        } else {
          uint32_t v[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(tacc + (uint32_t)c0) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float raw[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) raw[i] = PQ_ADD(__uint_as_float(v[i]), sbias[c0 + i]);
          if (WANT_RAW && in_level) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (i < cnt) P.out_raw[raw_base + (size_t)(c0 + i) * P.HW] = raw[i];
          }
          decode_sig_n(raw, cnt, trow + c0);
        }
      };
      if (SLICE) {
        // Anchor by anchor through two small staging tiles: decode anchor a's channels of the 128 cells, barrier, copy
        // its rows out (runs of 5+C floats, A*(5+C) apart) - the copy of slice i and the decode of slice i+1 need no
        // barrier between them, the other buffer is rewritten only after the next barrier.
        const int ewarp = (warp - 4), nwarp_e = 4 * P.wq;
        for (int a = 0; a < P.A; ++a, ++sl) {
          float* stile = tile0 + ((sl & 1u) ? tile_stride : 0);
          float* trow_a = stile + r * ch - a * ch;            // so that trow_a + c0 is channel (c0 - a*ch) of row r
#pragma unroll 1
          for (int i = jq; i <= grp; i += P.wq) do_unit(utab[i < grp ? a * grp + i : nsig + a], trow_a);
          if (a == P.A - 1) {
            // this warp no longer needs the accumulator: the MMA warp may start the tile after next in it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
          }
          epi_bar_sync(n_epi);
          if (P.out_dec) {
            for (int rr = ewarp; rr < ncell; rr += nwarp_e) {
              const float* srow = stile + rr * ch;
              float* drow = dst + ((size_t)rr * P.A_out + a) * ch;
              for (int k = lane; k < ch; k += 32) drow[k] = srow[k];
            }
          }
        }
        continue;
      }
      if (P.units) {
#pragma unroll 1
        for (int u = jq; u < nunit; u += P.wq) do_unit(utab[u], trow0);
      } else {
        int k0 = k_first;
        for (int blk = jq; blk < nblk; blk += P.wq) {
          const int c0 = blk * 8;
          uint32_t v[8];
          const uint32_t taddr = tmem_base + buf * (uint32_t)P.buf_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(taddr) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float* bs = sbias + c0;
          float* trow = tile + r * ACH + c0;
          float raw[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) raw[i] = PQ_ADD(__uint_as_float(v[i]), bs[i]);       // bias 0 where there is none
          if (WANT_RAW && in_level) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c0 + i < ACH) P.out_raw[raw_base + (size_t)(c0 + i) * P.HW] = raw[i];
          }
          decode_block8(raw, k0, c0, ACH, ch, gx, gy, P.stride, trow);      // pq_math.cuh: 8 interleaved chains
          k0 += k_step;
          if (k0 >= ch) k0 -= ch;
        }
      }
      // this warp no longer needs the accumulator: the MMA warp may start the tile after next in it
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (P.out_dec) {
        // a full tile whose rows start on a 16-byte boundary leaves as ONE bulk store; a partial last tile, or a level
        // whose row range is not 16-byte aligned in the concatenated prediction (608 x 608: 22743 rows per image),
        // leaves through coalesced stores of all epilogue threads
        const int nval = ncell * ACH;
        const int i0 = (4 - mis) & 3;                       // first element on a 16-byte boundary (global and shared)
        const int i1 = nval - ((mis + nval) & 3);           // end of the last whole 16-byte unit
        fence_async_smem();
        epi_bar_sync(n_epi);
        if (i1 > i0) {
          if (etid == 0) {
            tma_store_1d(dst + i0, tile + i0, (uint32_t)((i1 - i0) * sizeof(float)));
            tma_store_commit();
          }
          if (etid >= 32 && etid < 32 + i0) dst[etid - 32] = tile[etid - 32];
          if (etid >= 64 && etid < 64 + (nval - i1)) dst[i1 + etid - 64] = tile[i1 + etid - 64];
        } else {
          for (int i = etid; i < nval; i += n_epi) dst[i] = tile[i];
        }
      }
    }
    if (etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * P.buf_cols)) : "memory");
  }
}


template <bool WANT_RAW, bool SLICE = false>
__global__ void __launch_bounds__(768, 1)
head_conv_decode_ws_kernel(const __grid_constant__ HeadConvWsParams P, const __grid_constant__ CUtensorMap tmap_x) {
  head_conv_ws_body<WANT_RAW, false, SLICE>(P, &tmap_x, (int)blockIdx.x, (int)gridDim.x);
}

__global__ void __launch_bounds__(768, 1)
head_conv_hits_ws_kernel(const __grid_constant__ HeadConvWsParams P, const __grid_constant__ CUtensorMap tmap_x) {
  head_conv_ws_body<false, true>(P, &tmap_x, (int)blockIdx.x, (int)gridDim.x);
}

// All levels of the head in one launch: the CTAs are split between the levels in proportion to their estimated cost
// (cta_lo), every CTA keeps the weights of its own level resident and runs the same body on that level's tiles.
struct HeadConvLevelsParams {
  HeadConvWsParams P[PQDET_MAX_LEVELS];
  int cta_lo[PQDET_MAX_LEVELS + 1];
  int n_levels;
};
struct HeadConvLevelsMaps {
  CUtensorMap m[PQDET_MAX_LEVELS];
};

__global__ void __launch_bounds__(768, 1)
head_conv_decode_ws_levels_kernel(const __grid_constant__ HeadConvLevelsParams LP,
                                  const __grid_constant__ HeadConvLevelsMaps maps) {
  int l = 0;
  while (l + 1 < LP.n_levels && (int)blockIdx.x >= LP.cta_lo[l + 1]) ++l;
  head_conv_ws_body<false>(LP.P[l], &maps.m[l], (int)blockIdx.x - LP.cta_lo[l], LP.cta_lo[l + 1] - LP.cta_lo[l]);
}

// Anchor split in one launch: the CTAs are dealt to the anchors (cta_lo), every CTA keeps its anchor's slice of the
// weights resident and runs the sliced epilogue on that anchor's rows of every tile it takes.
template <bool WANT_RAW>
__global__ void __launch_bounds__(768, 1)
head_conv_decode_ws_anchors_kernel(const __grid_constant__ HeadConvLevelsParams LP,
                                   const __grid_constant__ HeadConvLevelsMaps maps) {
  int l = 0;
  while (l + 1 < LP.n_levels && (int)blockIdx.x >= LP.cta_lo[l + 1]) ++l;
  head_conv_ws_body<WANT_RAW, false, true>(LP.P[l], &maps.m[l], (int)blockIdx.x - LP.cta_lo[l],
                                           LP.cta_lo[l + 1] - LP.cta_lo[l]);
}

}  // namespace pq

namespace {

// Plans the persistent kernel for one level: 1 = P / tmap / smem are filled in, 0 = the shape does not qualify (the
// general kernel must run), < 0 = error.
int plan_head_conv_ws(const float* x, const float* weight, const float* bias, float* out_decoded, float* out_raw,
                      int B, int Cin, int H, int W, int A, int C, float stride, int64_t rows_total, int64_t row_off,
                      int device, pq::HeadConvWsParams* Pout, CUtensorMap* tmap_out, size_t* smem_out, int* sms_out,
                      int a_out = 0, int a_off = 0) {
  using namespace pq;
  const int HW = H * W, ACH = A * (5 + C), N = (ACH + 15) / 16 * 16;
  // the planes must be addressable by a tensor map (plane stride a multiple of 16 bytes); the last tile of a level
  // may be partial (38 x 38, 76 x 76), the output need not be 16-byte aligned (decided per tile in the epilogue)
  if (HW % 4 != 0 || Cin % 8 != 0 || N > 256) return 0;
  if (getenv("PQDET_HEADCONV_ALIGNED_ONLY") && HW % kHcM != 0) return 0;     // A/B switch: round-1 eligibility
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(weight) & 15)) return 0;
  if (out_decoded && (reinterpret_cast<uintptr_t>(out_decoded) & 3)) return 0;
  const int tiles_img = (HW + kHcM - 1) / kHcM;
  if ((int64_t)B * Cin > 0x7fffffff || (int64_t)B * tiles_img > 0x7fffffff) return 0;
  PqEncodeTiledFn enc = pq::encode_tiled_fn();
  if (!enc) return 0;
  HeadConvWsParams P;
  memset(&P, 0, sizeof(P));
  const size_t w_bytes = (size_t)Cin * N * 4, bias_bytes = (size_t)(N + 8) * 4;
  size_t tile_bytes = (size_t)(kHcM * ACH + 4) * 4;
  DeviceLimits lim;
  if (!device_limits(device, &lim)) return PQDET_ERR_CUDA;
  const int max_smem = lim.max_smem_optin, sms = lim.sms;
  // Stage = KC channels x 128 cells.  Every stage costs the MMA warp a fixed ~400 cycles of waits / fences / commit
  // (measured), so stages are as large as the budget allows: KC = the largest multiple of 8 that divides Cin with
  // at least 3 stages (>= 48 KB in flight) in what the resident weights and the output staging leave; two output
  // staging tiles (the decode of tile i+1 overlaps the bulk store of tile i) when that still holds.
  auto plan = [&](int bufs, int min_kc, int* kc_out, int* st_out) {
    const size_t fixed = 1024 + w_bytes + (size_t)bufs * tile_bytes + bias_bytes + 1024;   // alignment slack + static
    if (fixed >= (size_t)max_smem) return false;
    const size_t room = (size_t)max_smem - fixed;
    for (int kc = Cin > 256 ? 256 : Cin; kc >= min_kc; --kc) {
      if (kc % 8 || Cin % kc) continue;
      const size_t sb = (size_t)kc * 512;
      size_t st = room / sb;
      if (st > (size_t)kWsMaxStages) st = kWsMaxStages;
      if (st >= 3 && st * sb >= 48 * 1024) {
        while (st > 3 && (st - 1) * sb >= 128 * 1024) --st;
        *kc_out = kc; *st_out = (int)st;
        return true;
      }
    }
    return false;
  };
  // preference: large stages (>= 48 channels keep the MMA warp ahead of HBM) with two staging tiles, large stages
  // with one, then whatever fits
  P.A_out = a_out > 0 ? a_out : A;
  P.a_off = a_off;
  P.tile_bufs = 2;
  if (a_out > 0) {
    // one anchor of several: its rows are not contiguous in the prediction -> the sliced epilogue's row-wise copy
    tile_bytes = (size_t)(kHcM * (5 + C) + 4) * 4;
    P.slice = 1;
    if (!plan(2, 48, &P.KC, &P.stages) && !plan(2, 8, &P.KC, &P.stages)) return 0;
  } else if (!plan(2, 48, &P.KC, &P.stages)) {
    P.tile_bufs = 1;
    if (!plan(1, 48, &P.KC, &P.stages)) {
      P.tile_bufs = 2;
      if (!plan(2, 8, &P.KC, &P.stages)) {
        P.tile_bufs = 1;
        if (!plan(1, 8, &P.KC, &P.stages)) {
          // the whole tile does not fit beside the resident weights (COCO: 255 channels = 130 KB): two staging tiles
          // of ONE anchor's rows each, the epilogue works anchor by anchor
          tile_bytes = (size_t)(kHcM * (5 + C) + 4) * 4;
          P.tile_bufs = 2;
          P.slice = 1;
          if (!plan(2, 48, &P.KC, &P.stages) && !plan(2, 8, &P.KC, &P.stages)) return 0;
        }
      }
    }
  }
  const size_t stage_bytes = (size_t)P.KC * 512;
  const int nblk = (ACH + 7) / 8;
  // epilogue warps per TMEM lane quadrant (<= 5).  With anchor-aligned units: the count whose most loaded warp has the
  // least work, units dealt round robin (a group of objectness / class channels ~ 2.5 x the 4 box channels of an
  // anchor); with 8-column blocks: fewer warps when that does not lengthen the longest block list.
  int buf_cols = 32;
  while (buf_cols < N) buf_cols <<= 1;
  // anchor-aligned units (balanced: sizes usz and usz - 1): the 8-column read of the last one may run past the last
  // channel but must stay inside the accumulator buffer; the kernel's unit table holds 64 entries
  {
    const int ns = C + 1, grp = (ns + 7) / 8, usz = (ns + grp - 1) / grp, big = ns - grp * (usz - 1);
    const int last_cnt = big == grp ? usz : usz - 1;
    P.units = (ACH - last_cnt + 8 <= buf_cols && A * grp + A <= 64) ? 1 : 0;
  }
  if (P.slice && !P.units) return 0;
  int best = 1;
  if (P.units) {
    const int nsig = A * ((5 + C - 4 + 7) / 8), nunit = nsig + A;
    int best_load = 1 << 30;
    for (int w = 1; w <= 5 && w <= nunit; ++w) {
      int worst = 0;
      for (int j = 0; j < w; ++j) {
        int load = 0;
        for (int u = j; u < nunit; u += w) load += u < nsig ? 5 : 2;
        worst = load > worst ? load : worst;
      }
      if (worst < best_load) { best_load = worst; best = w; }
    }
  } else {
    best = nblk < 5 ? nblk : 5;
    while (best > 1 && (nblk + best - 2) / (best - 1) == (nblk + best - 1) / best) --best;
  }
  P.wq = best;
  P.w = weight; P.bias = bias; P.out_dec = out_decoded; P.out_raw = out_raw;
  P.B = B; P.Cin = Cin; P.HW = HW; P.Wd = W; P.A = A; P.C = C; P.N = N;
  P.buf_cols = 32;
  while (P.buf_cols < N) P.buf_cols <<= 1;
  P.tiles_per_img = tiles_img;
  {
    const size_t img_bytes = (size_t)rows_total * (5 + C) * 4, off_bytes = (size_t)row_off * (5 + C) * 4;
    (void)img_bytes; (void)off_bytes;
    // every tile commits exactly one bulk group as soon as it holds a whole 16-byte unit at any phase (>= 7 values)
    const int last_cells = HW % kHcM ? HW % kHcM : kHcM;
    P.all_bulk = (last_cells * ACH >= 7) ? 1 : 0;
  }
  P.magic_w = W == 1 ? 0u : (uint32_t)((0x100000000ull + (uint64_t)W - 1) / (uint64_t)W);
  if ((int64_t)HW * W >= 0x100000000ll) return 0;              // multiply-high exact for cell < 2^32 / W
  P.ntiles = B * P.tiles_per_img;
  P.stride = stride; P.rows_total = rows_total; P.row_off = row_off;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)B * Cin}, strides[1] = {(cuuint64_t)HW * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)P.KC}, estr[2] = {1u, 1u};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  *Pout = P;
  *tmap_out = tmap;
  *smem_out = 1024 + (size_t)P.stages * stage_bytes + w_bytes + (size_t)P.tile_bufs * tile_bytes + bias_bytes;
  *sms_out = sms;
  return 1;
}

int launch_head_conv_ws(const pq::HeadConvWsParams& P, const CUtensorMap& tmap, size_t smem, int sms, bool want_raw,
                        int device, cudaStream_t stream) {
  using namespace pq;
  const int grid = P.ntiles < sms ? P.ntiles : sms;
  static int smem_set[4][64];                 // the attribute sticks per device: raise it only when needed
  const int which = (want_raw ? 1 : 0) + (P.slice ? 2 : 0);
  void (*kern)(const HeadConvWsParams, const CUtensorMap) =
      which == 0 ? head_conv_decode_ws_kernel<false, false> : which == 1 ? head_conv_decode_ws_kernel<true, false>
      : which == 2 ? head_conv_decode_ws_kernel<false, true> : head_conv_decode_ws_kernel<true, true>;
  if (device < 0 || device >= 64 || (int)smem > smem_set[which][device]) {
    PQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (device >= 0 && device < 64) smem_set[which][device] = (int)smem;
  }
  kern<<<grid, (4 + 4 * P.wq) * 32, smem, stream>>>(P, tmap);
  PQ_LAUNCH_CHECK();
  return 1;
}

// Launches the persistent kernel for one level when the shape qualifies; 1 = launched, 0 = general kernel, < 0 = error.
int try_head_conv_ws(const float* x, const float* weight, const float* bias, float* out_decoded, float* out_raw,
                     int B, int Cin, int H, int W, int A, int C, float stride, int64_t rows_total, int64_t row_off,
                     int device, cudaStream_t stream) {
  using namespace pq;
  HeadConvWsParams P;
  CUtensorMap tmap;
  size_t smem = 0;
  int sms = 0;
  const int rc = plan_head_conv_ws(x, weight, bias, out_decoded, out_raw, B, Cin, H, W, A, C, stride, rows_total, row_off,
                                   device, &P, &tmap, &smem, &sms);
  if (rc < 0) return rc;
  if (rc == 1) return launch_head_conv_ws(P, tmap, smem, sms, out_raw != nullptr, device, stream);
  // The weights of all anchors do not fit beside the pipeline (COCO at Cin 176: 180 KB): one launch per anchor with
  // that anchor's slice of the weights resident; the features are read A times (mostly from L2), the rows of an anchor
  // leave through the sliced epilogue's row-wise copy.
  if (A < 2 || getenv("PQDET_HEADCONV_NO_SPLIT")) return 0;
  const int ch = 5 + C;
  if (A > PQDET_MAX_LEVELS) return 0;
  HeadConvLevelsParams LP;
  HeadConvLevelsMaps maps;
  memset(&LP, 0, sizeof(LP));
  memset(&maps, 0, sizeof(maps));
  size_t smem_a = 0;
  for (int a = 0; a < A; ++a) {
    size_t sm1 = 0;
    const int ra = plan_head_conv_ws(x, weight + (size_t)a * ch * Cin, bias ? bias + (size_t)a * ch : nullptr, out_decoded,
                                     out_raw, B, Cin, H, W, 1, C, stride, rows_total, row_off, device, &LP.P[a],
                                     &maps.m[a], &sm1, &sms, A, a);
    if (ra != 1) return ra;                                 // nothing launched yet: the general kernel takes the level
    smem_a = sm1 > smem_a ? sm1 : smem_a;
  }
  // the anchors cost the same: equal shares of the SMs (the first ones take the remainder)
  const int ctas = LP.P[0].ntiles * A < sms ? LP.P[0].ntiles * A : sms;
  if (ctas < A) return 0;
  LP.n_levels = A;
  for (int a = 0; a < A; ++a) LP.cta_lo[a + 1] = LP.cta_lo[a] + ctas / A + (a < ctas % A ? 1 : 0);
  for (int a = A; a < PQDET_MAX_LEVELS; ++a) LP.cta_lo[a + 1] = LP.cta_lo[A];
  static int smem_set[2][64];
  const int which = out_raw ? 1 : 0;
  void (*kern)(const HeadConvLevelsParams, const HeadConvLevelsMaps) =
      out_raw ? head_conv_decode_ws_anchors_kernel<true> : head_conv_decode_ws_anchors_kernel<false>;
  if (device < 0 || device >= 64 || (int)smem_a > smem_set[which][device]) {
    PQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    if (device >= 0 && device < 64) smem_set[which][device] = (int)smem_a;
  }
  kern<<<LP.cta_lo[A], (4 + 4 * LP.P[0].wq) * 32, smem_a, stream>>>(LP, maps);
  PQ_LAUNCH_CHECK();
  return 1;
}

}  // namespace

extern "C" int pqdet_head_conv_hits(const float* x, const float* weight, const float* bias, int B, int Cin, int H, int W,
                                    int A, int C, double score_threshold, int64_t row_offset, float* rec,
                                    int32_t* rec_count, int rec_cap, int device, void* stream) {
  using namespace pq;
  if (B < 0 || Cin < 1 || H < 1 || W < 1 || A < 1 || C < 1 || rec_cap < 1 || row_offset < 0) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;
  if (!x || !weight || !rec || !rec_count) return PQDET_ERR_INVALID_ARG;
  if (A * (5 + C) > 256 || B > 65535) return PQDET_ERR_UNSUPPORTED;
  if (row_offset + (int64_t)H * W * A >= (1ll << 31)) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  HeadConvWsParams P;
  CUtensorMap tmap;
  size_t smem = 0;
  int sms = 0;
  const int rc = getenv("PQDET_HEADCONV_GENERAL") ? 0 :
                 plan_head_conv_ws(x, weight, bias, nullptr, nullptr, B, Cin, H, W, A, C, 1.0f, (int64_t)H * W * A, 0,
                                   device, &P, &tmap, &smem, &sms);
  if (rc < 0) return rc;
  if (rc == 0) {
    // shape outside the persistent kernel (19 x 19: plane stride not a multiple of 16 bytes; weights beyond shared
    // memory): the general tcgen05 kernel with the same thresholding epilogue
    HeadConvParams G;
    memset(&G, 0, sizeof(G));
    G.x = x; G.w = weight; G.bias = bias;
    G.B = B; G.Cin = Cin; G.H = H; G.W = W; G.A = A; G.C = C;
    const int ACH = A * (5 + C);
    G.N = (ACH + 15) / 16 * 16;
    G.tmem_cols = 32;
    while (G.tmem_cols < G.N) G.tmem_cols <<= 1;
    G.row_off = row_offset;
    G.rec = rec; G.rec_count = rec_count; G.rec_cap = rec_cap;
    G.logit_lo = logit_lo_for((float)score_threshold);
    const size_t gsmem = 2 * ((size_t)(kHcKC / 4) * kHcM * 16 + (size_t)(kHcKC / 4) * G.N * 16);
    if (gsmem > 200 * 1024) return PQDET_ERR_UNSUPPORTED;
    PQ_CUDA(cudaFuncSetAttribute(head_conv_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    dim3 ggrid((H * W + kHcM - 1) / kHcM, B);
    head_conv_decode_kernel<true><<<ggrid, kHcThreads, gsmem, (cudaStream_t)stream>>>(G);
    PQ_LAUNCH_CHECK();
    return PQDET_OK;
  }
  P.rec = rec; P.rec_count = rec_count; P.rec_cap = rec_cap;
  P.row_off = row_offset;
  P.logit_lo = logit_lo_for((float)score_threshold);
  // three anchors' objectness columns per quadrant: more than A epilogue warps per quadrant would idle
  if (P.wq > A) P.wq = A;
  const int grid = P.ntiles < sms ? P.ntiles : sms;
  static int smem_set[64];
  if (device < 0 || device >= 64 || (int)smem > smem_set[device]) {
    PQ_CUDA(cudaFuncSetAttribute(head_conv_hits_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (device >= 0 && device < 64) smem_set[device] = (int)smem;
  }
  head_conv_hits_ws_kernel<<<grid, (4 + 4 * P.wq) * 32, smem, (cudaStream_t)stream>>>(P, tmap);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

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
  if (!getenv("PQDET_HEADCONV_GENERAL")) {
    const int rc = try_head_conv_ws(x, weight, bias, out_decoded, out_raw, B, Cin, H, W, A, C, stride, out_rows_total,
                                    out_row_offset, device, (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : PQDET_OK;
  }
  const size_t stages = 2 * ((size_t)(kHcKC / 4) * kHcM * 16 + (size_t)(kHcKC / 4) * P.N * 16);
  const size_t tile_bytes = (size_t)kHcM * (ACH | 1) * sizeof(float);
  // a decoded tile larger than the operand stages (255 channels: 130 KB against 96 KB) is staged in passes of 64
  // columns instead, so that shared memory - and with it the CTAs per SM - is set by the stages alone
  P.col_group = (tile_bytes > stages && !getenv("PQDET_HEADCONV_WHOLE_TILE")) ? 64 : 0;
  const size_t smem = (P.col_group || stages > tile_bytes) ? stages : tile_bytes;
  if (smem > 200 * 1024) return PQDET_ERR_UNSUPPORTED;
  PQ_CUDA(cudaFuncSetAttribute(head_conv_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((H * W + kHcM - 1) / kHcM, B);
  head_conv_decode_kernel<false><<<grid, kHcThreads, smem, (cudaStream_t)stream>>>(P);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

// All levels of the head in one launch (DetectionHead.forward_from_features): level l's rows follow level l-1's in the
// (B, sum_l H_l*W_l*A, 5+C) prediction.  PQDET_ERR_UNSUPPORTED when a level does not qualify for the persistent kernel
// (the caller then runs pqdet_head_conv_decode per level).
extern "C" int pqdet_head_conv_decode_levels(int n_levels, const float* const* x, const float* const* weight,
                                             const float* const* bias, const int* Cin, const int* H, const int* W,
                                             const float* stride, float* out_decoded, int B, int A, int C, int device,
                                             void* stream) {
  using namespace pq;
  if (n_levels < 1 || n_levels > PQDET_MAX_LEVELS || !x || !weight || !Cin || !H || !W || !stride)
    return PQDET_ERR_INVALID_ARG;
  if (B < 0 || A < 1 || C < 0) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;
  if (!out_decoded) return PQDET_ERR_INVALID_ARG;
  if (getenv("PQDET_HEADCONV_GENERAL")) return PQDET_ERR_UNSUPPORTED;
  int64_t rows_total = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (!x[l] || !weight[l] || Cin[l] < 1 || H[l] < 1 || W[l] < 1) return PQDET_ERR_INVALID_ARG;
    rows_total += (int64_t)H[l] * W[l] * A;
  }
  PQ_ENTER(device);
  // Measured (VOC-512, B200): one launch wins while launch gaps and per-level tails matter (35 vs 55 us at 16 images,
  // 71 vs 84 us at 64) and loses once every level fills the machine by itself (246 vs 218 us at 256 images, where the
  // static split between the levels costs more than two launch gaps): decline above ~32 tiles per SM, before any
  // planning work is spent.
  {
    DeviceLimits lim;
    if (!device_limits(device, &lim)) return PQDET_ERR_CUDA;
    int64_t all_tiles = 0;
    for (int l = 0; l < n_levels; ++l) all_tiles += (int64_t)B * ((H[l] * W[l] + kHcM - 1) / kHcM);
    if (all_tiles > (int64_t)32 * lim.sms) return PQDET_ERR_UNSUPPORTED;
  }
  HeadConvLevelsParams LP;
  HeadConvLevelsMaps maps;
  memset(&LP, 0, sizeof(LP));
  memset(&maps, 0, sizeof(maps));
  size_t smem = 0;
  int sms = 0, threads = 0;
  int64_t row_off = 0;
  double cost[PQDET_MAX_LEVELS], total = 0.0;
  for (int l = 0; l < n_levels; ++l) {
    size_t sm_l = 0;
    const int rc = plan_head_conv_ws(x[l], weight[l], bias ? bias[l] : nullptr, out_decoded, nullptr, B, Cin[l], H[l], W[l], A,
                                     C, stride[l], rows_total, row_off, device, &LP.P[l], &maps.m[l], &sm_l, &sms);
    if (rc < 0) return rc;
    if (rc == 0 || LP.P[l].slice) return PQDET_ERR_UNSUPPORTED;      // (sliced epilogue: per-level launches)
    smem = sm_l > smem ? sm_l : smem;
    const int th = (4 + 4 * LP.P[l].wq) * 32;
    if (threads && th != threads) return PQDET_ERR_UNSUPPORTED;
    threads = th;
    row_off += (int64_t)H[l] * W[l] * A;
    // time of one tile: a fixed part (epilogue, hand-offs) + the features it streams (measured: 8.0 / 3.8 / 2.3 us per
    // tile at Cin 352 / 176 / 80)
    cost[l] = (double)LP.P[l].ntiles * (1.0 + 0.02 * Cin[l]);
    total += cost[l];
  }
  if (sms < n_levels) return PQDET_ERR_UNSUPPORTED;
  // CTAs per level in proportion to the cost, at least one, never more than the level has tiles
  int given = 0, share[PQDET_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) {
    int c = (int)(sms * cost[l] / total + 0.5);
    c = c < 1 ? 1 : c;
    c = c > LP.P[l].ntiles ? LP.P[l].ntiles : c;
    share[l] = c;
    given += c;
  }
  while (given > sms) {                       // rounding overshoot: take from the level with the most CTAs
    int m = 0;
    for (int l = 1; l < n_levels; ++l) m = share[l] > share[m] ? l : m;
    if (share[m] <= 1) return PQDET_ERR_UNSUPPORTED;
    --share[m]; --given;
  }
  for (bool grew = true; given < sms && grew;) {   // leftovers to the level with the most tiles per CTA
    grew = false;
    int m = -1;
    double worst = 0.0;
    for (int l = 0; l < n_levels; ++l) {
      const double per = cost[l] / share[l];
      if (share[l] < LP.P[l].ntiles && per > worst) { worst = per; m = l; }
    }
    if (m >= 0) { ++share[m]; ++given; grew = true; }
  }
  LP.n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) LP.cta_lo[l + 1] = LP.cta_lo[l] + share[l];
  for (int l = n_levels; l < PQDET_MAX_LEVELS; ++l) LP.cta_lo[l + 1] = LP.cta_lo[n_levels];
  static int smem_set[64];
  if (device < 0 || device >= 64 || (int)smem > smem_set[device]) {
    PQ_CUDA(cudaFuncSetAttribute(head_conv_decode_ws_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (device >= 0 && device < 64) smem_set[device] = (int)smem;
  }
  head_conv_decode_ws_levels_kernel<<<LP.cta_lo[n_levels], threads, smem, (cudaStream_t)stream>>>(LP, maps);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
