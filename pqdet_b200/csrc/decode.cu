// Materialising decode / decode-backward / recover kernels.
//
// decode_fwd transposes NCHW planes into (cell, anchor, channel) rows through shared memory:
// each warp reads 32 consecutive cells of one channel plane (one 128-byte line), applies the
// decode function, and the tile is written back as one contiguous run of 32*A*(5+C) floats.
// HBM traffic is the algorithmic 2R (read once, write once); there is no reuse to exploit.
#include <string.h>

#include <atomic>

#include <stdlib.h>

#include "pq_common.cuh"

namespace pq {

constexpr int kTileCells = 32;
constexpr int kDecodeThreads = 256;

__device__ __forceinline__ void decode_tile(const float* __restrict__ raw, float* __restrict__ out, int A, int ch,
                                            int H, int W, float stride, int64_t rows_total, int64_t row_off,
                                            int tile_index, int b, float* tile) {
  const int HW = H * W;
  const int ACH = A * ch;
  const int ST = ACH | 1;  // odd row stride: conflict-free column writes (== ACH whenever ACH is odd)
  const int cell0 = tile_index * kTileCells;
  const int ncell = min(kTileCells, HW - cell0);
  const int lane = lane_id();
  const int cell = cell0 + lane;
  const int cy = cell / W, cx = cell - cy * W;
  const float* src = raw + (size_t)b * ACH * HW + cell;
  float* tcol = tile + lane * ST;
  // The kernel is issue bound before it is HBM bound, so the per-element bookkeeping is kept to an increment: every
  // warp owns a contiguous run of channels (the channel-within-anchor index k walks with a wrap, no division), four
  // channels per iteration with their loads issued together and their exp / reciprocal chains interleaved
  // (rcp_rn_core = __frcp_rn without its per-call range check; out-of-range values redo the group with __frcp_rn,
  // so the result is bit-identical to sigmoidf_ / decode_coord).
  constexpr int W8 = kDecodeThreads / 32;
  if (lane < ncell) {
    const int cpw = (ACH + W8 - 1) / W8;                  // channels per warp
    const int c_lo = warp_id() * cpw, c_hi = min(ACH, c_lo + cpw);
    int k = c_lo % ch;
    const float* p = src + (size_t)c_lo * HW;
    float* t = tcol + c_lo;
    const float gx = (float)cx + 0.5f, gy = (float)cy + 0.5f;
    constexpr int U = 4;
    for (int c = c_lo; c < c_hi; c += U, p += (size_t)U * HW, t += U) {
      float v[U], e[U], rr[U];
      int kk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        kk[u] = k;
        if (++k == ch) k = 0;
        v[u] = (c + u < c_hi) ? ldg_stream(p + (size_t)u * HW) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) e[u] = expf(kk[u] < 4 ? v[u] : -v[u]);
      bool slow = false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float x = PQ_ADD(1.0f, e[u]);
        rr[u] = rcp_rn_core(x);
        slow |= (kk[u] >= 4) && !(x < kRcpCoreMax);
      }
      if (slow) {
#pragma unroll
        for (int u = 0; u < U; ++u) rr[u] = __frcp_rn(PQ_ADD(1.0f, e[u]));               // == sigmoidf_(v)
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float g = (kk[u] & 1) ? gy : gx;
        const float oc = PQ_MUL((kk[u] < 2) ? PQ_SUB(g, e[u]) : PQ_ADD(g, e[u]), stride);   // == decode_coord(k, v, cx, cy, stride)
        if (c + u < c_hi) t[u] = kk[u] < 4 ? oc : rr[u];
      }
    }
  }
  __syncthreads();
  // the tile is one contiguous run of ncell*A*ch floats in the output
  float* dst = out + ((size_t)b * rows_total + row_off + (size_t)cell0 * A) * ch;
  const int n = ncell * ACH;
  if (ST == ACH && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n & 3) == 0)) {
    // no padding between tile rows: shared memory holds the run as is -> 128-bit copy
    const float4* t4 = reinterpret_cast<const float4*>(tile);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int e = threadIdx.x; e < (n >> 2); e += kDecodeThreads) d4[e] = t4[e];
  } else {
    for (int r = warp_id(); r < ncell; r += W8) {
      const float* trow = tile + r * ST;
      float* drow = dst + (size_t)r * ACH;
      for (int c = lane; c < ACH; c += 32) drow[c] = trow[c];
    }
  }
}

__global__ void __launch_bounds__(kDecodeThreads)
decode_fwd_kernel(const float* __restrict__ raw, float* __restrict__ out, int A, int ch, int H, int W,
                  float stride, int64_t rows_total, int64_t row_off) {
  extern __shared__ __align__(16) float tile[];  // [kTileCells][ST]
  decode_tile(raw, out, A, ch, H, W, stride, rows_total, row_off, blockIdx.x, blockIdx.y, tile);
}

// All levels of the eval concat (model/interpreter.py:75-76) in ONE launch: grid.x enumerates the tiles of every
// level back to back, grid.y = image; each level writes its own row range of the (B, N, 5+C) output.
struct DecodeLevels {
  const float* raw[PQDET_MAX_LEVELS];
  int H[PQDET_MAX_LEVELS], W[PQDET_MAX_LEVELS];
  float stride[PQDET_MAX_LEVELS];
  int64_t row_off[PQDET_MAX_LEVELS];
  int tile_off[PQDET_MAX_LEVELS + 1];
  int n_levels;
};

__global__ void __launch_bounds__(kDecodeThreads)
decode_levels_kernel(const __grid_constant__ DecodeLevels L, float* __restrict__ out, int A, int ch,
                     int64_t rows_total, int behind_primary) {
  extern __shared__ __align__(16) float tile[];
  int l = 0;
#pragma unroll
  for (int i = 1; i < PQDET_MAX_LEVELS; ++i)
    if (i < L.n_levels && (int)blockIdx.x >= L.tile_off[i]) l = i;
  decode_tile(L.raw[l], out, A, ch, L.H[l], L.W[l], L.stride[l], rows_total, L.row_off[l],
              (int)blockIdx.x - L.tile_off[l], blockIdx.y, tile);
  // Launched as a programmatic dependent of the TMA pipeline (disjoint row ranges, no data dependency): one CTA
  // stays until that grid has completed, so that THIS grid - the last one of the call in stream order - does not
  // complete before it and whatever the caller enqueues next sees all rows.
  if (behind_primary && blockIdx.x == 0 && blockIdx.y == 0) cudaGridDependencySynchronize();
}

// grad_raw[b][c][cell] = grad_out[b][cell][a][k] * d out/d raw:
//   k<2: -exp(raw)*stride, k in 2,3: +exp(raw)*stride, k>=4: p(1-p)   (autograd of parser.py:226-232)
__global__ void __launch_bounds__(kDecodeThreads)
decode_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ gout, float* __restrict__ graw,
                  int A, int ch, int H, int W, float stride, int64_t rows_total, int64_t row_off) {
  extern __shared__ float tile[];
  const int HW = H * W;
  const int ACH = A * ch;
  const int ST = ACH | 1;
  const int b = blockIdx.y;
  const int cell0 = blockIdx.x * kTileCells;
  const int ncell = min(kTileCells, HW - cell0);
  const float* gsrc = gout + ((size_t)b * rows_total + row_off + (size_t)cell0 * A) * ch;
  const int lane = lane_id();
  for (int r = warp_id(); r < ncell; r += kDecodeThreads / 32) {
    const float* grow = gsrc + (size_t)r * ACH;
    float* trow = tile + r * ST;
    for (int c = lane; c < ACH; c += 32) trow[c] = grow[c];
  }
  __syncthreads();
  const int cell = cell0 + lane;
  const size_t base = (size_t)b * ACH * HW;
  for (int c = warp_id(); c < ACH; c += kDecodeThreads / 32) {
    if (lane < ncell) {
      float v = raw[base + (size_t)c * HW + cell];
      float g = tile[lane * ST + c];
      int k = c % ch;
      float d;
      if (k < 4) {
        float es = expf(v) * stride;
        d = (k < 2) ? -es : es;
      } else {
        float p = sigmoidf_(v);
        d = p * (1.0f - p);
      }
      graw[base + (size_t)c * HW + cell] = g * d;
    }
  }
}

// recover: a CTA takes kRecRows consecutive rows of one image.  Input rows are one contiguous run of
// kRecRows*(5+C) floats, output rows one contiguous run of kRecRows*(4+C): both move through shared memory with
// coalesced (128-bit when aligned) accesses.  In between, the 4 box coordinates of all rows are recovered by one
// thread each (the IEEE division is then executed once per warp instruction, not once per row), and the class
// scores by warps sweeping rows.  ncu on the previous one-warp-per-row version: 87 % of the issue slots busy at
// 2.3 TB/s, i.e. instruction bound; this layout needs ~3.5x fewer instructions per row.
constexpr int kRecRows = 64;
constexpr int kRecThreads = 256;
static_assert(kRecRows * 4 == kRecThreads, "one thread per (row, coordinate)");

// One tile: rtile [nrow][5+C] -> otile [nrow][4+C], both in shared memory.  Coordinates: one thread per (row, k).
// Scores: consecutive threads take consecutive output elements; row = element / C by multiply-high with
// magic_c = ceil(2^32 / C) (exact for element < 2^32 / C), so the inner loop is a handful of instructions.
__device__ __forceinline__ void recover_tile(const float* rtile, float* otile, int nrow, int C, uint32_t magic_c,
                                             const Affine& af) {
  const int ic = 5 + C, oc = 4 + C;
  const int tid = threadIdx.x;
  {
    const int r = tid >> 2, k = tid & 3;
    if (r < nrow) otile[r * oc + k] = recover_coord(k, rtile[r * ic + k], af);
  }
  const int n = nrow * C;
#pragma unroll 2
  for (int e = tid; e < n; e += kRecThreads) {
    const int r = (C == 1) ? e : (int)__umulhi((unsigned)e, magic_c);   // ceil(2^32 / 1) does not fit 32 bits
    const int c = e - r * C;
    const float* trow = rtile + r * ic;
    otile[r * oc + 4 + c] = PQ_MUL(trow[5 + c], trow[4]);
  }
}

__global__ void __launch_bounds__(kRecThreads)
recover_kernel(const float* __restrict__ pred, float* __restrict__ out, int64_t N, int C, int kind,
               float in_h, float in_w, const float* __restrict__ orig_hw, int orig_per_image, uint32_t magic_c) {
  extern __shared__ __align__(16) float rtile[];          // [kRecRows][5+C] in, then [kRecRows][4+C] out
  __shared__ Affine s_af;
  const int ic = 5 + C, oc = 4 + C;
  float* otile = rtile + ((kRecRows * ic + 3) & ~3);
  const int b = blockIdx.y;
  const int64_t row0 = (int64_t)blockIdx.x * kRecRows;
  const int nrow = (int)min((int64_t)kRecRows, N - row0);
  const int tid = threadIdx.x;
  if (tid == 0) {
    const float* o = orig_hw + (orig_per_image ? 2 * b : 0);
    s_af = affine_params(kind, in_h, in_w, o[0], o[1]);
  }
  const float* src = pred + ((size_t)b * N + row0) * ic;
  const int n_in = nrow * ic;
  if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n_in & 3) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* t4 = reinterpret_cast<float4*>(rtile);
    for (int e = tid; e < (n_in >> 2); e += kRecThreads) {
      float4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(s4 + e));
      t4[e] = v;
    }
  } else {
    for (int e = tid; e < n_in; e += kRecThreads) rtile[e] = ldg_stream(src + e);
  }
  __syncthreads();
  recover_tile(rtile, otile, nrow, C, magic_c, s_af);
  __syncthreads();
  float* dst = out + ((size_t)b * N + row0) * oc;
  const int n_out = nrow * oc;
  if (((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((n_out & 3) == 0)) {
    const float4* o4 = reinterpret_cast<const float4*>(otile);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int e = tid; e < (n_out >> 2); e += kRecThreads) d4[e] = o4[e];
  } else {
    for (int e = tid; e < n_out; e += kRecThreads) dst[e] = otile[e];
  }
}

// The same computation as a persistent, TMA-fed pipeline (used whenever every tile is a 16-byte aligned run): each
// CTA loops over tiles; one thread keeps kRecStages - 1 bulk loads in flight ahead of the tile being processed
// (completion on an mbarrier per stage) and sends the finished output tile back with a bulk store, so the SM always
// has several tiles of HBM traffic outstanding instead of alternating between a load phase and a compute phase.
constexpr int kRecStages = 4;     // input ring
constexpr int kRecOutStages = 2;  // output ring

__global__ void __launch_bounds__(kRecThreads)
recover_tma_kernel(const float* __restrict__ pred, float* __restrict__ out, int64_t N, int C, int kind,
                   float in_h, float in_w, const float* __restrict__ orig_hw, int orig_per_image,
                   unsigned tiles_per_image, unsigned total_tiles, uint32_t magic_c) {
  extern __shared__ __align__(128) unsigned char rsm[];
  __shared__ __align__(8) uint64_t full[kRecStages];
  __shared__ Affine s_af[2];
  const int ic = 5 + C, oc = 4 + C;
  const uint32_t in_stride = (uint32_t)((kRecRows * ic * 4 + 127) & ~127);
  const uint32_t out_stride = (uint32_t)((kRecRows * oc * 4 + 127) & ~127);
  unsigned char* in_base = rsm;
  unsigned char* out_base = rsm + (size_t)kRecStages * in_stride;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  auto tile_rows = [&](unsigned chunk) -> int {
    return (int)min((int64_t)kRecRows, N - (int64_t)chunk * kRecRows);
  };
  auto issue_load = [&](unsigned t, int stage) {
    const unsigned b = t / tiles_per_image, chunk = t - b * tiles_per_image;
    const uint32_t bytes = (uint32_t)(tile_rows(chunk) * ic * 4);
    mbar_expect_tx(&full[stage], bytes);
    tma_load_1d(in_base + (size_t)stage * in_stride, pred + ((size_t)b * N + (size_t)chunk * kRecRows) * ic, bytes,
                &full[stage]);
  };
  auto image_affine_of = [&](unsigned t) -> Affine {
    const float* o = orig_hw + (orig_per_image ? 2 * (size_t)(t / tiles_per_image) : 0);
    return affine_params(kind, in_h, in_w, o[0], o[1]);
  };
  const unsigned first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    for (int s = 0; s < kRecStages; ++s) mbar_init(&full[s], 1);
    mbar_init_fence();
    for (int s = 0; s < kRecStages - 1; ++s)
      if ((uint64_t)first + (uint64_t)s * step < total_tiles) issue_load(first + s * step, s);
    s_af[0] = image_affine_of(first);
  }
  __syncthreads();
  unsigned i = 0;
  for (uint64_t t64 = first; t64 < total_tiles; t64 += step, ++i) {
    const unsigned t = (unsigned)t64;
    const int stage = (int)(i % kRecStages), ostage = (int)(i % kRecOutStages);
    const unsigned b = t / tiles_per_image, chunk = t - b * tiles_per_image;
    const int nrow = tile_rows(chunk);
    if (tid == 0) {
      tma_store_wait_read<kRecOutStages - 1>();                 // the store that last used out[ostage] has read it
      const uint64_t tn = t64 + (uint64_t)(kRecStages - 1) * step;   // refill the stage consumed one iteration ago
      if (tn < total_tiles) issue_load((unsigned)tn, (int)((i + kRecStages - 1) % kRecStages));
    }
    __syncthreads();                                            // (A) out[ostage] is free
    const Affine af = s_af[i & 1];
    mbar_wait(&full[stage], (uint32_t)((i / kRecStages) & 1));
    const float* rtile = reinterpret_cast<const float*>(in_base + (size_t)stage * in_stride);
    float* otile = reinterpret_cast<float*>(out_base + (size_t)ostage * out_stride);
    recover_tile(rtile, otile, nrow, C, magic_c, af);
    fence_async_smem();
    __syncthreads();                                            // (B) tile computed; in[stage] may be refilled
    if (tid == 0) {
      tma_store_1d(out + ((size_t)b * N + (size_t)chunk * kRecRows) * oc, otile, (uint32_t)(nrow * oc * 4));
      tma_store_commit();
    }
    // the affine of the NEXT tile's image, by another warp while thread 0 refills the ring (read after barrier A)
    if (warp == kRecThreads / 32 - 1 && lane == 0 && t64 + step < total_tiles)
      s_af[(i + 1) & 1] = image_affine_of((unsigned)(t64 + step));
  }
  if (tid == 0) tma_store_wait_read<0>();                       // smem must stay alive until the stores have read it
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent TMA pipeline for the eval concat (all levels, every level's H*W a multiple of 128): one CTA per SM, a
// producer warp streams (A*(5+C) channels x 128 cells) tiles of the raw heads with one 2-D tensor-map load each
// through an mbarrier ring, 4*wq compute warps decode them (thread = cell, 8 consecutive channels per step, the
// interleaved chains of decode_block8) into a staging tile that is one contiguous run of the output and leaves as
// one bulk store; two staging tiles, so the decode of tile i+1 overlaps the store of tile i.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDtmCells = 128;
constexpr int kDtmMaxStages = 8;
constexpr int kDtmMaxUnits = 64;     // work units of one cell row (ACH <= 256: at most 32 + A)

struct DecodeTmaParams {
  float* out;
  int A, ch, n_levels;
  int Wd[PQDET_MAX_LEVELS];
  int HW[PQDET_MAX_LEVELS];               // cells of a plane; the last tile of a level may be partial
  int bulk[PQDET_MAX_LEVELS];             // 1: the level's row range is 16-byte aligned -> tiles leave as bulk stores
  uint32_t magic_w[PQDET_MAX_LEVELS];     // ceil(2^32 / W): cell / W by multiply-high (0 when W == 1)
  float stride[PQDET_MAX_LEVELS];
  int64_t row_off[PQDET_MAX_LEVELS];
  int tile_off[PQDET_MAX_LEVELS + 1];     // tiles of one image, levels concatenated
  int tiles_img, ntiles;
  int64_t rows_total;
  int stages, wq;
};
struct DecodeTmaMaps {
  CUtensorMap m[PQDET_MAX_LEVELS];
};

// CELLS = cells per tile: 128 (four 32-cell quarters, up to 5 compute warps each) or, where four 128-cell tiles of
// A*(5+C) channels do not fit in shared memory (COCO: 255 channels), 32 (one quarter, up to 20 compute warps sharing
// the units of its cells).
template <int CELLS>
__global__ void __launch_bounds__(768, 1)
decode_levels_tma_kernel(const __grid_constant__ DecodeTmaParams P, const __grid_constant__ DecodeTmaMaps maps) {
  constexpr int NQ = CELLS / 32;
  // a level this pipeline cannot take (19 x 19) is decoded by decode_levels_kernel launched right behind it as a
  // programmatic dependent: its small CTAs run beside the persistent ones instead of after them
  cudaTriggerProgrammaticLaunchCompletion();
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ __align__(8) uint64_t full_bar[kDtmMaxStages], empty_bar[kDtmMaxStages];
  __shared__ uint32_t utab[kDtmMaxUnits];
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int ch = P.ch, ACH = P.A * ch, S = P.stages;
  const uint32_t stage_bytes = (uint32_t)ACH * CELLS * 4u;
  // dynamic smem: [S input stages: ACH x 128 floats][2 staging tiles: 128 x ACH floats]
  float* tile0 = reinterpret_cast<float*>(dsm + (size_t)S * stage_bytes);
  const int n_cmp = NQ * P.wq * 32;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NQ * P.wq);
    }
    mbar_init_fence();
  }
  __syncthreads();
  if (warp == 0) {
    // ---- producer ----------------------------------------------------------------------------------------------
    uint32_t s = 0, ph = 0;
    for (int g = blockIdx.x; g < P.ntiles; g += gridDim.x) {
      const int b = g / P.tiles_img, t = g - b * P.tiles_img;
      int l = 0;
      while (l + 1 < P.n_levels && t >= P.tile_off[l + 1]) ++l;
      const int cell0 = (t - P.tile_off[l]) * CELLS;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[s], stage_bytes);
        tma_load_2d(smem_u32(dsm) + s * stage_bytes, &maps.m[l], cell0, b * ACH, &full_bar[s]);
      }
      __syncwarp();
      if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
    }
  } else if (warp >= 4) {
    // ---- compute: warp%4 = which 32 cells of the tile, the wq warps of a quarter take the 8-channel blocks round robin
    const int q = (warp - 4) % NQ, jq = (warp - 4) / NQ;
    const int ctid = tid - 128;
    const int r = q * 32 + lane;
    // work units of one cell row: per anchor its 1 + C objectness / class channels in `grp` groups of `usz` <= 8
    // (21 -> 3 x 7), numbered first, then the 4 box channels of every anchor (cheaper: no reciprocal); the wq warps of
    // a quarter take the units round robin
    const int ns = ch - 4, grp = (ns + 7) / 8, usz = (ns + grp - 1) / grp, nsig = P.A * grp, nunit = nsig + P.A;
    // the units are the same for every tile: (first channel of the head | channel count << 16; count 0 = the 4 box
    // channels), worked out once into shared memory (no division in the tile loop)
    for (int u = ctid; u < nunit; u += n_cmp) {
      uint32_t e;
      if (u >= nsig) {
        e = (uint32_t)((u - nsig) * ch);
      } else {
        // balanced: the first `big` groups of an anchor hold usz channels, the others usz - 1 (81 = 4 x 8 + 7 x 7)
        const int big = ns - grp * (usz - 1);
        const int a = u / grp, j = u - a * grp;
        const int k = 4 + (j < big ? j * usz : big * usz + (j - big) * (usz - 1));
        e = (uint32_t)(a * ch + k) | ((uint32_t)(j < big ? usz : usz - 1) << 16);
      }
      utab[u] = e;
    }
    epi_bar_sync(n_cmp);
    uint32_t s = 0, ph = 0, tl = 0;
    // per-tile bookkeeping without divisions: (image, tile within image) advance by gridDim.x with carries
    int b = blockIdx.x / P.tiles_img, t = blockIdx.x - b * P.tiles_img;
    const int step_b = gridDim.x / P.tiles_img, step_t = gridDim.x - step_b * P.tiles_img;
    for (int g = blockIdx.x; g < P.ntiles; g += gridDim.x, ++tl) {
      int l = 0;
      while (l + 1 < P.n_levels && t >= P.tile_off[l + 1]) ++l;
      const int cell0 = (t - P.tile_off[l]) * CELLS;
      const int cell = cell0 + r;
      const int ncell = min(CELLS, P.HW[l] - cell0);     // < 128 on a level's last tile (TMA zero-fills the rest)
      const int Wd = P.Wd[l];
      const int cy = P.magic_w[l] ? (int)__umulhi((uint32_t)cell, P.magic_w[l]) : cell, cx = cell - cy * Wd;
      const float gx = (float)cx + 0.5f, gy = (float)cy + 0.5f;
      const float stride = P.stride[l];
      float* dst = P.out + ((size_t)b * P.rows_total + P.row_off[l] + (size_t)cell0 * P.A) * ch;
      // the staging tile starts at the same offset inside a 16-byte unit as its destination: everything but at most
      // 3 floats at either end leaves as one bulk store even when this image's rows are not 16-byte aligned in the
      // prediction (608 x 608: 22743 rows per image; a level behind a 19 x 19 level)
      const int mis = (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3);
      float* tile = tile0 + (size_t)(tl & 1u) * (CELLS * ACH + 4) + mis;
      // the bulk store that last used this staging tile must have read it before it is overwritten
      if (ctid == 0) tma_store_wait_read<1>();
      epi_bar_sync(n_cmp);
      mbar_wait(&full_bar[s], ph);
      const float* src = reinterpret_cast<const float*>(dsm + (size_t)s * stage_bytes) + r;
      float* trow0 = tile + r * ACH;
#ifndef PQ_DTM_NOCOMPUTE
      if (r < ncell) {
#pragma unroll 1
        for (int u = jq; u < nunit; u += P.wq) {
          const uint32_t e = utab[u];
          const int c0 = (int)(e & 0xffffu), cnt = (int)(e >> 16);
          const float* sp = src + c0 * CELLS;
          if (cnt == 0) {
            float raw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) raw[k] = sp[k * CELLS];
            decode_box4(raw, gx, gy, stride, trow0 + c0);
          } else {
            decode_sig_n_strided(sp, CELLS, cnt, trow0 + c0);
          }
        }
      }
#endif
      t += step_t; b += step_b;
      if (t >= P.tiles_img) { t -= P.tiles_img; ++b; }
      // this warp has read its part of the stage: the producer may refill it
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
      {
        const int n = ncell * ACH;
        const int i0 = (4 - mis) & 3;                       // first element on a 16-byte boundary (global and shared)
        const int i1 = n - ((mis + n) & 3);                 // end of the last whole 16-byte unit
        fence_async_smem();
        epi_bar_sync(n_cmp);
        // every tile commits exactly one bulk group (so "at most one pending" above = the other staging tile's)
        if (ctid == 0) {
#ifndef PQ_DTM_NOSTORE
          if (i1 > i0) tma_store_1d(dst + i0, tile + i0, (uint32_t)((i1 - i0) * sizeof(float)));
#endif
          tma_store_commit();
        }
        if (i1 > i0) {
          if (ctid >= 32 && ctid < 32 + i0) dst[ctid - 32] = tile[ctid - 32];
          if (ctid >= 64 && ctid < 64 + (n - i1)) dst[i1 + ctid - 64] = tile[i1 + ctid - 64];
        } else {
          for (int e = ctid; e < n; e += n_cmp) dst[e] = tile[e];
        }
      }
    }
    if (ctid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

}  // namespace pq

namespace {

// Launches the persistent TMA kernel when every level qualifies; 1 = launched, 0 = use the general kernel, < 0 = error.
// A level qualifies when a tensor map can address its planes: plane stride (H*W*4 bytes) a multiple of 16, base aligned.
bool decode_level_tma_ok(const float* raw, int H, int W) {
  return ((H * W) % 4 == 0) && (reinterpret_cast<uintptr_t>(raw) & 15) == 0 && (int64_t)H * W * W < 0x100000000ll;
}

int try_decode_levels_tma(int n_levels, const float* const* raw, const int* H, const int* W, const float* stride,
                          const int64_t* row_offs, float* out, int B, int A, int C, int64_t rows, int device,
                          cudaStream_t stream) {
  using namespace pq;
  const int ch = 5 + C, ACH = A * ch;
  if (ACH > 256 || A * ((ch - 4 + 7) / 8) + A > kDtmMaxUnits) return 0;
  DeviceLimits lim;
  if (!device_limits(device, &lim)) return PQDET_ERR_CUDA;
  const int max_smem = lim.max_smem_optin, sms = lim.sms;
  const size_t room = (size_t)max_smem - 1024 - 32;                 // static barriers, 16 bytes of phase room per staging tile
  // cells per tile: 128 where two input stages + two staging tiles of that size fit, else 32 (255 channels)
  int cells = 128;
  if (room < 4 * (size_t)cells * ACH * 4 || getenv("PQDET_DECODE_CELLS32")) cells = 32;
  if (room < 4 * (size_t)cells * ACH * 4) return 0;
  const bool out_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (((size_t)rows * ch * 4) & 15) == 0;
  PqEncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return 0;
  DecodeTmaParams P;
  DecodeTmaMaps maps;
  memset(&P, 0, sizeof(P));
  memset(&maps, 0, sizeof(maps));
  int tiles = 0;
  for (int l = 0; l < n_levels; ++l) {
    const int HW = H[l] * W[l];
    if (!decode_level_tma_ok(raw[l], H[l], W[l])) return 0;
    if ((int64_t)B * ACH > 0x7fffffff) return 0;
    const int64_t row_off = row_offs[l];
    P.Wd[l] = W[l]; P.HW[l] = HW; P.stride[l] = stride[l]; P.row_off[l] = row_off; P.tile_off[l] = tiles;
    P.bulk[l] = (out_aligned && (((size_t)row_off * ch * 4) & 15) == 0) ? 1 : 0;
    P.magic_w[l] = W[l] == 1 ? 0u : (uint32_t)((0x100000000ull + (uint64_t)W[l] - 1) / (uint64_t)W[l]);
    tiles += (HW + cells - 1) / cells;
    cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)B * ACH}, strides[1] = {(cuuint64_t)HW * 4};
    cuuint32_t box[2] = {(cuuint32_t)cells, (cuuint32_t)ACH}, estr[2] = {1u, 1u};
    if (enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(raw[l]), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 0;
  }
  for (int l = n_levels; l <= PQDET_MAX_LEVELS; ++l) P.tile_off[l] = tiles;
  if ((int64_t)B * tiles > 0x7fffffff) return 0;
  P.out = out; P.A = A; P.ch = ch; P.n_levels = n_levels;
  P.tiles_img = tiles; P.ntiles = B * tiles; P.rows_total = rows;
  const size_t tile_bytes = (size_t)cells * ACH * 4;
  size_t st = room / tile_bytes - 2;
  if (st > (size_t)kDtmMaxStages) st = kDtmMaxStages;
  if (const char* e = getenv("PQDET_DECODE_STAGES")) {                // A/B switch: cap the input ring
    const size_t cap = (size_t)atoi(e);
    if (cap >= 2 && cap < st) st = cap;
  }
  P.stages = (int)st;
  // compute warps per quarter: the count (<= 5) whose most loaded warp has the least work, units dealt round robin
  // (cost model: a group of objectness / class channels ~ 2.5 x the 4 box channels of an anchor)
  const int nsig = A * ((ch - 4 + 7) / 8), nunit = nsig + A;
  int wq = 1, best_load = 1 << 30;
  const int nq = cells / 32;
  for (int w = 1; w <= 20 / nq && w <= nunit; ++w) {
    int worst = 0;
    for (int j = 0; j < w; ++j) {
      int load = 0;
      for (int u = j; u < nunit; u += w) load += u < nsig ? 5 : 2;
      worst = load > worst ? load : worst;
    }
    if (worst < best_load) { best_load = worst; wq = w; }
  }
  P.wq = wq;
  const size_t smem = (st + 2) * tile_bytes + 32;
  static int smem_set[2][64];                 // the attribute sticks per device: raise it only when needed
  const int which = cells == 128 ? 0 : 1;
  void (*kern)(const DecodeTmaParams, const DecodeTmaMaps) =
      which == 0 ? decode_levels_tma_kernel<128> : decode_levels_tma_kernel<32>;
  if (device < 0 || device >= 64 || (int)smem > smem_set[which][device]) {
    PQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (device >= 0 && device < 64) smem_set[which][device] = (int)smem;
  }
  const int grid = P.ntiles < sms ? P.ntiles : sms;
  kern<<<grid, (4 + nq * wq) * 32, smem, stream>>>(P, maps);
  PQ_LAUNCH_CHECK();
  return 1;
}

}  // namespace

extern "C" int pqdet_decode_fwd(const float* raw, float* out, int B, int A, int C, int H, int W, float stride,
                                int64_t out_rows_total, int64_t out_row_offset, int device, void* stream) {
  if (B < 0 || A <= 0 || C < 0 || H <= 0 || W <= 0) return PQDET_ERR_INVALID_ARG;
  if (out_row_offset < 0 || out_row_offset + (int64_t)H * W * A > out_rows_total) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;                       // empty batch: tensors without storage are fine
  if (!raw || !out) return PQDET_ERR_INVALID_ARG;
  if (B > 65535) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  const int ch = 5 + C;
  if (!getenv("PQDET_DECODE_GENERAL")) {
    const int rc = try_decode_levels_tma(1, &raw, &H, &W, &stride, &out_row_offset, out, B, A, C, out_rows_total, device,
                                         (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : PQDET_OK;
  }
  const size_t smem = (size_t)pq::kTileCells * ((A * ch) | 1) * sizeof(float);
  if (smem > 48 * 1024)
    PQ_CUDA(cudaFuncSetAttribute(pq::decode_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((H * W + pq::kTileCells - 1) / pq::kTileCells, B);
  pq::decode_fwd_kernel<<<grid, pq::kDecodeThreads, smem, (cudaStream_t)stream>>>(
      raw, out, A, ch, H, W, stride, out_rows_total, out_row_offset);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}


extern "C" int pqdet_decode_levels(int n_levels, const float* const* raw, const int* H, const int* W,
                                   const float* stride, float* out, int B, int A, int C, int device, void* stream) {
  using namespace pq;
  if (n_levels < 1 || n_levels > PQDET_MAX_LEVELS || !raw || !H || !W || !stride) return PQDET_ERR_INVALID_ARG;
  if (B < 0 || A <= 0 || C < 0) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;
  if (B > 65535) return PQDET_ERR_UNSUPPORTED;
  if (!out) return PQDET_ERR_INVALID_ARG;
  int64_t rows = 0, row_off[PQDET_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) {
    if (!raw[l] || H[l] < 1 || W[l] < 1) return PQDET_ERR_INVALID_ARG;
    row_off[l] = rows;
    rows += (int64_t)H[l] * W[l] * A;
  }
  PQ_ENTER(device);
  const int ch = 5 + C;
  // Levels a tensor map can address (plane stride a multiple of 16 bytes) go to the persistent TMA pipeline, the
  // others (e.g. the 19x19 level of a 608 input) to the general kernel; both write their own row ranges of `out`.
  bool on_tma[PQDET_MAX_LEVELS] = {false, false, false, false};
  if (!getenv("PQDET_DECODE_GENERAL")) {
    const float* t_raw[PQDET_MAX_LEVELS];
    int t_H[PQDET_MAX_LEVELS], t_W[PQDET_MAX_LEVELS], idx[PQDET_MAX_LEVELS], nt = 0;
    float t_stride[PQDET_MAX_LEVELS];
    int64_t t_off[PQDET_MAX_LEVELS];
    for (int l = 0; l < n_levels; ++l) {
      if (!decode_level_tma_ok(raw[l], H[l], W[l])) continue;
      t_raw[nt] = raw[l]; t_H[nt] = H[l]; t_W[nt] = W[l]; t_stride[nt] = stride[l]; t_off[nt] = row_off[l];
      idx[nt++] = l;
    }
    if (nt > 0) {
      const int rc = try_decode_levels_tma(nt, t_raw, t_H, t_W, t_stride, t_off, out, B, A, C, rows, device,
                                           (cudaStream_t)stream);
      if (rc < 0) return rc;
      if (rc == 1)
        for (int i = 0; i < nt; ++i) on_tma[idx[i]] = true;
    }
  }
  DecodeLevels L;
  memset(&L, 0, sizeof(L));
  int tiles = 0, ng = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (on_tma[l]) continue;
    L.raw[ng] = raw[l]; L.H[ng] = H[l]; L.W[ng] = W[l]; L.stride[ng] = stride[l];
    L.row_off[ng] = row_off[l]; L.tile_off[ng] = tiles;
    tiles += (H[l] * W[l] + kTileCells - 1) / kTileCells;
    ++ng;
  }
  if (ng == 0) return PQDET_OK;
  for (int l = ng; l <= PQDET_MAX_LEVELS; ++l) L.tile_off[l] = tiles;
  L.n_levels = ng;
  const size_t smem = (size_t)kTileCells * ((A * ch) | 1) * sizeof(float);
  if (smem > 48 * 1024)
    PQ_CUDA(cudaFuncSetAttribute(decode_levels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(tiles, B);
  bool behind = false;
  for (int l = 0; l < n_levels; ++l) behind |= on_tma[l];
  if (behind && !getenv("PQDET_DECODE_NO_PDL")) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kDecodeThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PQ_CUDA(cudaLaunchKernelEx(&cfg, decode_levels_kernel, L, out, A, ch, rows, 1));
  } else {
    decode_levels_kernel<<<grid, kDecodeThreads, smem, (cudaStream_t)stream>>>(L, out, A, ch, rows, 0);
  }
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

extern "C" int pqdet_decode_bwd(const float* raw, const float* grad_out, float* grad_raw, int B, int A, int C,
                                int H, int W, float stride, int64_t out_rows_total, int64_t out_row_offset,
                                int device, void* stream) {
  if (B < 0 || A <= 0 || C < 0 || H <= 0 || W <= 0) return PQDET_ERR_INVALID_ARG;
  if (out_row_offset < 0 || out_row_offset + (int64_t)H * W * A > out_rows_total) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;
  if (!raw || !grad_out || !grad_raw) return PQDET_ERR_INVALID_ARG;
  if (B > 65535) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  const int ch = 5 + C;
  const size_t smem = (size_t)pq::kTileCells * ((A * ch) | 1) * sizeof(float);
  if (smem > 48 * 1024)
    PQ_CUDA(cudaFuncSetAttribute(pq::decode_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((H * W + pq::kTileCells - 1) / pq::kTileCells, B);
  pq::decode_bwd_kernel<<<grid, pq::kDecodeThreads, smem, (cudaStream_t)stream>>>(
      raw, grad_out, grad_raw, A, ch, H, W, stride, out_rows_total, out_row_offset);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

extern "C" int pqdet_recover(const float* pred, float* out, int B, int64_t N, int C, int affine_kind,
                             float in_h, float in_w, const float* orig_hw, int orig_per_image,
                             int device, void* stream) {
  if (B < 0 || N < 0 || C < 0) return PQDET_ERR_INVALID_ARG;
  if (affine_kind < 0 || affine_kind > 2) return PQDET_ERR_INVALID_ARG;
  if ((int64_t)B * N == 0) return PQDET_OK;
  if (!pred || !out || !orig_hw) return PQDET_ERR_INVALID_ARG;
  if (B > 65535) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  const uint32_t magic_c = C > 0 ? (uint32_t)(((1ull << 32) + (uint64_t)C - 1) / (uint64_t)C) : 0u;   // ceil(2^32 / C)
  {
    // TMA pipeline whenever every tile (incl. the last, shorter one of an image) is a 16-byte aligned run
    const int ic = 5 + C, oc = 4 + C;
    const int64_t tail = N % pq::kRecRows;
    const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                         ((N * ic) % 4 == 0) && ((N * oc) % 4 == 0) && ((pq::kRecRows * ic) % 4 == 0) &&
                         ((pq::kRecRows * oc) % 4 == 0) && ((tail * ic) % 4 == 0) && ((tail * oc) % 4 == 0);
    const size_t in_stride = ((size_t)pq::kRecRows * ic * 4 + 127) & ~(size_t)127;
    const size_t out_stride = ((size_t)pq::kRecRows * oc * 4 + 127) & ~(size_t)127;
    const size_t tsm = pq::kRecStages * in_stride + pq::kRecOutStages * out_stride;
    if (aligned && tsm <= 200 * 1024 && ((N + pq::kRecRows - 1) / pq::kRecRows) * B < (1ll << 31)) {
      const int64_t tiles_per_image = (N + pq::kRecRows - 1) / pq::kRecRows;
      const int64_t total_tiles = tiles_per_image * B;
      // launch geometry is a pure function of (device, smem): remember the last one per device
      static std::atomic<uint64_t> cache[16];
      int64_t grid = 0;
      if (device < 16) {
        const uint64_t c = cache[device].load(std::memory_order_relaxed);
        if ((c >> 32) == (uint64_t)tsm) grid = (int64_t)(c & 0xffffffffu);
      }
      if (grid == 0) {
        PQ_CUDA(cudaFuncSetAttribute(pq::recover_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        int per_sm = 1, sm_count = 148;
        PQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pq::recover_tma_kernel, pq::kRecThreads, tsm));
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
        grid = (int64_t)sm_count * (per_sm < 1 ? 1 : per_sm);
        if (device < 16) cache[device].store(((uint64_t)tsm << 32) | (uint32_t)grid, std::memory_order_relaxed);
      }
      if (grid > total_tiles) grid = total_tiles;
      pq::recover_tma_kernel<<<(unsigned)grid, pq::kRecThreads, tsm, (cudaStream_t)stream>>>(
          pred, out, N, C, affine_kind, in_h, in_w, orig_hw, orig_per_image, (unsigned)tiles_per_image,
          (unsigned)total_tiles, magic_c);
      PQ_LAUNCH_CHECK();
      return PQDET_OK;
    }
  }
  const size_t smem = ((size_t)((pq::kRecRows * (5 + C) + 3) & ~3) + (size_t)pq::kRecRows * (4 + C)) * sizeof(float);
  if (smem > 48 * 1024)
    PQ_CUDA(cudaFuncSetAttribute(pq::recover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((N + pq::kRecRows - 1) / pq::kRecRows), B);
  pq::recover_kernel<<<grid, pq::kRecThreads, smem, (cudaStream_t)stream>>>(
      pred, out, N, C, affine_kind, in_h, in_w, orig_hw, orig_per_image, magic_c);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
