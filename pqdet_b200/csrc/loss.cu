// Fused decode + loss_per_scale forward + backward.
//
// One pass over the level: a CTA owns 32 consecutive cells x A anchors of one image (warp = anchor,
// lane = cell, so every raw-plane access is one 128-byte line).  It stages the contiguous label
// tile in shared memory, decodes the box + objectness channels, evaluates the box term, the
// ignore mask against the image's GT list (staged through shared memory, exact early-outs), the
// objectness term and - only for responsible cells - the class term, and writes d loss / d raw for
// all A*(5+C) channels with coalesced stores (zeros where no gradient flows: max_iou only feeds a
// comparison, class/box terms are gated by `respond`).  Per-CTA partial sums are reduced in fp64
// in a fixed order by a second, tiny kernel: run-to-run deterministic, no float atomics.
//
// HBM traffic: label tile L + 5/(5+C) of the raw head + the class channels of responsible cells,
// plus the full gradient write R.  Algorithmic figure used for the roofline: 2R + L.
#include <string.h>

#include "pq_common.cuh"

namespace pq {

constexpr int kLossTile = 32;

struct LossParams {
  const float* x;       // raw (B, A*ch, H, W) or decoded (B,H,W,A,ch)
  const float* label;   // (B,H,W,A,6+C)
  const float* gt;      // (B,G,4)
  float* grad;          // same shape as x, nullable
  double* partials;     // [B * tiles][3]
  int B, A, C, H, W, G;
  float stride, in_area, ignore_thresh, l1_gain, inv_B;
  int bbox_loss;
  // sparse targets (pqdet_loss_levels_sparse): label rows are rebuilt from the owner map + the GT rows
  const int32_t* owner;  // (B, A, H*W): index into gt6 of the GT owning the label slot, or -1
  const float* gt6;      // (B, n_max, 6) rows [x1,y1,x2,y2,class,mixw]
  int n_max;
  float hot, cold;       // smoothed one-hot values (train_dataset.py:126-130)
};

// The two terms that only responsible cells need (a fraction of a percent of the rows, but one warp in seven holds
// one) live out of line: inlined - the class term eight times over - they made the kernel 65 KB of SASS, twice the
// instruction cache, and every warp that took them evicted the hot loop of its neighbours (ncu: icc hit rate 77 %,
// no_inst 18 % of the stall samples).
struct BoxLossOut { float v, d0, d1, d2, d3; };
__device__ __noinline__ BoxLossOut bbox_loss_row_cold(int kind, float4 p, float4 t, float respond, float in_area,
                                                      float l1_gain) {
  const float pb[4] = {p.x, p.y, p.z, p.w}, tb[4] = {t.x, t.y, t.z, t.w};
  float d[4];
  BoxLossOut o;
  o.v = bbox_loss_row(kind, pb, tb, respond, in_area, l1_gain, d);
  o.d0 = d[0]; o.d1 = d[1]; o.d2 = d[2]; o.d3 = d[3];
  return o;
}
// Four class channels at once (four independent chains: a warp with a responsible cell is on the step's critical
// path at small batches) -> values and d value / d logit-or-probability before the * mixw / B factor.
struct ClassTerms4 { float v[4], d[4]; };
__device__ __noinline__ ClassTerms4 class_terms4_cold(float4 z4, float4 t4, float respond, int raw) {
  const float z[4] = {z4.x, z4.y, z4.z, z4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w};
  ClassTerms4 o;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float p = raw ? sigmoidf_(z[u]) : z[u];
    float dp;
    o.v[u] = focal_bce_term(2.0f, 0.5f, t[u], p, respond, -1.0f, &dp);
    o.d[u] = raw ? dp * (p * (1.0f - p)) : dp;
  }
  return o;
}

// Every warp is autonomous: warp = anchor, lane = cell of the tile.  There is no CTA-wide barrier until the
// final 3-value reduction, so a warp never waits for its siblings' memory latency (at 16 images per GPU the
// kernel is latency bound: every dependent DRAM round trip that can be overlapped or dropped is step time).
template <bool RAW, bool SPARSE = false>
__device__ __forceinline__ void loss_tile(const LossParams& P, const int tile, const int ntiles, const int b,
                                          float* smem) {
  const int A = P.A, C = P.C, ch = 5 + C, LW = 6 + C;
  const int HW = P.H * P.W;
  __shared__ float sred[8][3];
  const int cell0 = tile * kLossTile;
  const int ncell = min(kLossTile, HW - cell0);
  const int lane = lane_id(), a = warp_id();
  float* sgt = smem + a * (32 * 5);                     // this warp's culled GT boxes of the current round
  const int cell = cell0 + lane;
  const bool active = lane < ncell;

  // ---- loads first, all independent: GT boxes of the first round, label row (or owner), raw planes ----
  const float4* gtb = reinterpret_cast<const float4*>(P.gt + (size_t)b * P.G * 4);
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < P.G) q = __ldg(gtb + lane);
  // label row: only [box(4), respond] and the trailing mixw are needed for every row; the class
  // targets are read at responsible cells only.
  const float* lab = SPARSE ? nullptr : P.label + (((size_t)b * HW + cell) * A + a) * LW;
  float tb[4] = {0.f, 0.f, 0.f, 0.f}, respond = 0.f, mixw = 0.f;
  int cls = -1, own = -1;
  if (SPARSE) {
    if (active) own = __ldg(P.owner + ((size_t)b * A + a) * HW + cell);
  } else if (active) {
    if ((LW & 1) == 0) {
      const float2 t01 = __ldg(reinterpret_cast<const float2*>(lab));
      const float2 t23 = __ldg(reinterpret_cast<const float2*>(lab) + 1);
      tb[0] = t01.x; tb[1] = t01.y; tb[2] = t23.x; tb[3] = t23.y;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) tb[k] = __ldg(lab + k);
    }
    respond = __ldg(lab + 4);
    mixw = __ldg(lab + 5 + C);
  }
  const size_t plane0 = ((size_t)b * A + a) * ch * HW;             // RAW: first plane of this anchor
  const size_t prow = (((size_t)b * HW + cell) * A + a) * ch;      // !RAW: this row
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
    if (RAW) {
#pragma unroll
      for (int k = 0; k < 5; ++k) v[k] = ldg_stream(P.x + plane0 + (size_t)k * HW + cell);
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) v[k] = P.x[prow + k];
    }
  }

  // ---- class-channel gradient = 0 (responsible cells overwrite theirs further down, after a __syncwarp).
  // Each plane segment of the tile is 32 consecutive floats: 128-bit stores when the planes are aligned.
  if (P.grad) {
    if (RAW) {
      const bool vec = ((HW & 3) == 0) && (ncell == kLossTile);
      float* seg0 = P.grad + plane0 + (size_t)5 * HW + cell0;
      if (vec) {
        for (int i = lane; i < C * 8; i += 32)
          reinterpret_cast<float4*>(seg0 + (size_t)(i >> 3) * HW)[i & 7] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (active) {
#pragma unroll 1
        for (int c = 0; c < C; ++c) seg0[(size_t)c * HW + lane] = 0.f;
      }
    } else if (active) {
      float* gp = P.grad + prow + 5;
#pragma unroll 1
      for (int c = 0; c < C; ++c) gp[c] = 0.f;
    }
  }

  if (SPARSE) {
    // the label row of train_dataset.py:135-145: [gt box, 1, smooth one-hot, mixw] where a GT owns the slot,
    // [0, 0, 0, 0, 0, 0.., 1] (background, mixw = 1) elsewhere
    if (active) {
      mixw = 1.0f;
      if (own >= 0) {
        const float2* g = reinterpret_cast<const float2*>(P.gt6 + ((size_t)b * P.n_max + own) * 6);
        const float2 g01 = __ldg(g), g23 = __ldg(g + 1), g45 = __ldg(g + 2);
        tb[0] = g01.x; tb[1] = g01.y; tb[2] = g23.x; tb[3] = g23.y;
        cls = (int)g45.x;
        mixw = g45.y;
        respond = 1.0f;
      }
    }
  }
  // ---- prediction: box + objectness ----
  float pb[4] = {0.f, 0.f, 1.f, 1.f}, es[4] = {0.f, 0.f, 0.f, 0.f}, pconf = 0.5f;
  if (active) {
    if (RAW) {
      const int cy = cell / P.W, cx = cell - cy * P.W;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pb[k] = decode_coord(k, v[k], cx, cy, P.stride);
        const float e = expf(v[k]) * P.stride;
        es[k] = (k < 2) ? -e : e;                                  // d pb[k] / d raw[k]
      }
      pconf = sigmoidf_(v[4]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) pb[k] = v[k];
      pconf = v[4];
    }
  }
  // ---- box term ----
  float dbox[4] = {0.f, 0.f, 0.f, 0.f};
  float lb = 0.f;
  if (active) {
    if (bbox_loss_row_is_zero(pb, tb, respond)) {
      lb = 0.0f;
    } else {
      const BoxLossOut o = bbox_loss_row_cold(P.bbox_loss, make_float4(pb[0], pb[1], pb[2], pb[3]),
                                              make_float4(tb[0], tb[1], tb[2], tb[3]), respond, P.in_area, P.l1_gain);
      lb = o.v; dbox[0] = o.d0; dbox[1] = o.d1; dbox[2] = o.d2; dbox[3] = o.d3;
    }
  }

  // ---- ignore mask: every GT has iou < thr (NaN -> false), only needed where respond != 1 ----
  // Exact culling per warp: a GT that does not strictly overlap the union bounding box of this warp's
  // predictions has inter == 0 with every one of them, i.e. iou == +0 < thr, and cannot change the mask -
  // provided thr > 0, the prediction areas are positive finite numbers and the GT area is a finite number
  // >= 0 (otherwise 0/0 or NaN could appear, so culling is switched off for the warp / that GT).
  bool below = true;
  const bool need = active && (respond != 1.0f);
  const float a1 = box_area(pb[0], pb[1], pb[2], pb[3]);
  float ux1 = INFINITY, uy1 = INFINITY, ux2 = -INFINITY, uy2 = -INFINITY;
  if (need) {
    const bool sane = (a1 > 0.0f) && (a1 < INFINITY) && (P.ignore_thresh > 0.0f);
    ux1 = sane ? pb[0] : -INFINITY; uy1 = sane ? pb[1] : -INFINITY;
    ux2 = sane ? pb[2] : INFINITY;  uy2 = sane ? pb[3] : INFINITY;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ux1 = fminf(ux1, __shfl_xor_sync(PQ_FULL, ux1, d)); uy1 = fminf(uy1, __shfl_xor_sync(PQ_FULL, uy1, d));
    ux2 = fmaxf(ux2, __shfl_xor_sync(PQ_FULL, ux2, d)); uy2 = fmaxf(uy2, __shfl_xor_sync(PQ_FULL, uy2, d));
  }
  // this warp's staged GT boxes: 32 x float4, then their 32 areas
  float4* sbox = reinterpret_cast<float4*>(sgt);
  float* sarea = sgt + 128;
  // every prediction of the warp that needs the mask is "sane" (finite coordinates, positive finite area, thr > 0)
  const bool wsane = __all_sync(PQ_FULL, !need || ux1 > -INFINITY);
  for (int g0 = 0; g0 < P.G; g0 += 32) {
    if (!__any_sync(PQ_FULL, need && below)) break;
    if (g0 > 0) {
      q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g0 + lane < P.G) q = __ldg(gtb + g0 + lane);
    }
    const float a2 = box_area(q.x, q.y, q.z, q.w);
    const bool cullable = (a2 >= 0.0f) && (a2 < INFINITY);
    const bool overlaps = (q.z > ux1) && (q.x < ux2) && (q.w > uy1) && (q.y < uy2);
    const bool take = (g0 + lane < P.G) && (!cullable || overlaps);
    const unsigned tm = __ballot_sync(PQ_FULL, take);
    // with sane predictions and staged GTs of finite area >= 0 (finite coordinates), a pair whose intersection
    // width or height is <= 0 has inter == +0 over a union > 0: iou == 0 < thr, no arithmetic needed beyond the
    // two differences; everything else takes the full test
    const bool fast = wsane && !__any_sync(PQ_FULL, take && !cullable);
    if (take) {
      const int at = __popc(tm & ((1u << lane) - 1u));
      sbox[at] = q;
      sarea[at] = a2;
    }
    __syncwarp();
    const int nk = __popc(tm);
    if (need && below) {
      if (fast) {
        for (int g = 0; g < nk; ++g) {
          const float4 e = sbox[g];
          const float w = PQ_SUB(fminf(pb[2], e.z), fmaxf(pb[0], e.x));
          const float h = PQ_SUB(fminf(pb[3], e.w), fmaxf(pb[1], e.y));
          if (w > 0.0f && h > 0.0f &&
              !iou_below(pb[0], pb[1], pb[2], pb[3], a1, e.x, e.y, e.z, e.w, sarea[g], P.ignore_thresh)) {
            below = false;
            break;
          }
        }
      } else {
        for (int g = 0; g < nk; ++g) {
          const float4 e = sbox[g];
          if (!iou_below(pb[0], pb[1], pb[2], pb[3], a1, e.x, e.y, e.z, e.w, sarea[g], P.ignore_thresh)) {
            below = false;
            break;
          }
        }
      }
    }
    __syncwarp();
  }
  // torch.max over an empty GT axis would raise; collate always pads to G >= 1.
  const float bgd = PQ_MUL(PQ_SUB(1.0f, respond), below ? 1.0f : 0.0f);

  // ---- objectness term ----
  float dconf = 0.f, lc = 0.f;
  if (active) lc = focal_bce_term(1.0f, 0.75f, respond, pconf, respond, bgd, &dconf);

  const float gw = mixw * P.inv_B;
  // ---- class term: only the rare responsible cells (their class gradients overwrite the zeros written at
  // the top; __syncwarp orders the two stores).  Logits/targets are fetched in batches of 8 so
  // the scattered-plane load latency is paid once per batch, not once per class.
  float lp = 0.f;
  if (active && respond != 0.0f) {
    constexpr int U = 8;
    for (int c0 = 0; c0 < C; c0 += U) {
      float z[U], t[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u;
        z[u] = 0.0f; t[u] = 0.0f;
        if (c < C) {
          z[u] = RAW ? P.x[plane0 + (size_t)(5 + c) * HW + cell] : P.x[prow + 5 + c];
          t[u] = SPARSE ? ((c == cls) ? P.hot : P.cold) : __ldg(lab + 5 + c);
        }
      }
#pragma unroll
      for (int h = 0; h < U; h += 4) {
        if (c0 + h < C) {
          // lanes past C evaluate a harmless (0, 0) pair; their results are dropped
          const ClassTerms4 o = class_terms4_cold(make_float4(z[h], z[h + 1], z[h + 2], z[h + 3]),
                                                  make_float4(t[h], t[h + 1], t[h + 2], t[h + 3]), respond, RAW ? 1 : 0);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + h + u;
            if (c < C) {
              lp = PQ_ADD(lp, o.v[u]);
              if (P.grad) {
                if (RAW) P.grad[plane0 + (size_t)(5 + c) * HW + cell] = o.d[u] * gw;
                else P.grad[prow + 5 + c] = o.d[u] * gw;
              }
            }
          }
        }
      }
    }
  }
  if (active && P.grad) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (RAW) P.grad[plane0 + (size_t)k * HW + cell] = dbox[k] * es[k] * gw;
      else P.grad[prow + k] = dbox[k] * gw;
    }
    if (RAW) P.grad[plane0 + (size_t)4 * HW + cell] = dconf * (pconf * (1.0f - pconf)) * gw;
    else P.grad[prow + 4] = dconf * gw;
  }

  // ---- partial sums: fixed-order fp32 butterfly inside the warp, fp64 across warps / CTAs ----
  float v0 = active ? PQ_MUL(lb, mixw) : 0.0f;
  float v1 = active ? PQ_MUL(lc, mixw) : 0.0f;
  float v2 = active ? PQ_MUL(lp, mixw) : 0.0f;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 += __shfl_xor_sync(PQ_FULL, v0, d);
    v1 += __shfl_xor_sync(PQ_FULL, v1, d);
    v2 += __shfl_xor_sync(PQ_FULL, v2, d);
  }
  if (lane == 0) { sred[a][0] = v0; sred[a][1] = v1; sred[a][2] = v2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
#pragma unroll 1
    for (int w = 0; w < A; ++w) s += (double)sred[w][threadIdx.x];
    P.partials[((size_t)b * ntiles + tile) * 3 + threadIdx.x] = s;
  }
}

template <bool RAW>
__global__ void __launch_bounds__(256, 4)
loss_fwd_bwd_kernel(const __grid_constant__ LossParams P) {
  extern __shared__ __align__(16) float smem[];
  loss_tile<RAW>(P, blockIdx.x, gridDim.x, blockIdx.y, smem);
}

// ---- all FPN levels in ONE launch (DetectionModel.forward training branch, model/interpreter.py:77-85)
// grid.x enumerates the tiles of every level back to back, grid.y = image.  A second, single-CTA kernel
// (loss_levels_finalize_kernel) reduces every level's partial sums in a fixed order - deterministic - and writes
//   out[0..3]            loss, bbox, conf, cls summed over levels in Python-sum order ((0+h0)+h1)+h2
//   out[4+4l .. 7+4l]    the four (1,) losses of level l (what YOLOLayer.forward returns)
//   out[4+4L+l]          loss_per_branch[l] = (bbox+conf)+cls of level l
struct MultiLossParams {
  LossParams lv[PQDET_MAX_LEVELS];
  int tile_off[PQDET_MAX_LEVELS + 1];
  int n_levels;
  float* out;
  int32_t* nan_flag;
};

#ifndef PQ_LOSS_MINB
#define PQ_LOSS_MINB 4
#endif
template <bool SPARSE>
__global__ void __launch_bounds__(256, PQ_LOSS_MINB)
loss_levels_kernel(const __grid_constant__ MultiLossParams M) {
  extern __shared__ __align__(16) float smem[];
  // programmatic dependent launch: let the finalize kernel be scheduled while this grid drains (it waits in
  // cudaGridDependencySynchronize() until every CTA here has completed and its partials are visible)
  cudaTriggerProgrammaticLaunchCompletion();
  int l = 0;
#pragma unroll
  for (int i = 1; i < PQDET_MAX_LEVELS; ++i)
    if (i < M.n_levels && (int)blockIdx.x >= M.tile_off[i]) l = i;
  const int ntiles = M.tile_off[l + 1] - M.tile_off[l];
  const int b = blockIdx.y;
  const LossParams& P = M.lv[l];
  loss_tile<true, SPARSE>(P, blockIdx.x - M.tile_off[l], ntiles, b, smem);
}

// One CTA of 1024 threads finishes the step: every thread accumulates its share of every level's per-CTA partial
// sums (all loads independent, one L2 round trip), then warp butterflies and a fixed-order sum over the 32 warp
// results - deterministic - and thread 0 writes the outputs.  Measured alternatives: doing this inside the main
// kernel with completion tickets costs more (every CTA then holds its SM slot for a fence + atomic round trip).
__global__ void __launch_bounds__(1024)
loss_levels_finalize_kernel(const __grid_constant__ MultiLossParams M) {
  __shared__ double s_w[32][PQDET_MAX_LEVELS * 3];
  cudaGridDependencySynchronize();
  const int lane = lane_id(), warp = warp_id();
  double acc[PQDET_MAX_LEVELS * 3];
#pragma unroll
  for (int i = 0; i < PQDET_MAX_LEVELS * 3; ++i) acc[i] = 0.0;
#pragma unroll
  for (int q = 0; q < PQDET_MAX_LEVELS; ++q) {
    if (q >= M.n_levels) break;
    const LossParams& P = M.lv[q];
    const int64_t n = (int64_t)P.B * (M.tile_off[q + 1] - M.tile_off[q]);
#pragma unroll 4
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
      acc[q * 3 + 0] += P.partials[i * 3 + 0]; acc[q * 3 + 1] += P.partials[i * 3 + 1];
      acc[q * 3 + 2] += P.partials[i * 3 + 2];
    }
  }
#pragma unroll
  for (int i = 0; i < PQDET_MAX_LEVELS * 3; ++i) {
    if (i >= M.n_levels * 3) break;
    const double v = warp_sum(acc[i]);
    if (lane == 0) s_w[warp][i] = v;
  }
  __syncthreads();
  if (warp == 0) {
    // lane i < 3L sums component i over the 32 warps in index order
    double sum = 0.0;
    if (lane < M.n_levels * 3)
      for (int w = 0; w < 32; ++w) sum += s_w[w][lane];
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    bool nan = false;
    for (int q = 0; q < M.n_levels; ++q) {
      const double invB = 1.0 / (double)M.lv[q].B;
      const float lb = (float)(__shfl_sync(PQ_FULL, sum, q * 3 + 0) * invB);
      const float lc = (float)(__shfl_sync(PQ_FULL, sum, q * 3 + 1) * invB);
      const float lp = (float)(__shfl_sync(PQ_FULL, sum, q * 3 + 2) * invB);
      const float loss = PQ_ADD(PQ_ADD(lb, lc), lp);            // model/loss.py:108
      if (lane == 0) {
        float* o = M.out + 4 + 4 * q;
        o[0] = loss; o[1] = lb; o[2] = lc; o[3] = lp;
        M.out[4 + 4 * M.n_levels + q] = PQ_ADD(PQ_ADD(lb, lc), lp);
      }
      tot[0] = q ? PQ_ADD(tot[0], loss) : loss; tot[1] = q ? PQ_ADD(tot[1], lb) : lb;
      tot[2] = q ? PQ_ADD(tot[2], lc) : lc;     tot[3] = q ? PQ_ADD(tot[3], lp) : lp;
      nan |= (loss != loss);
    }
    if (lane == 0) {
      for (int j = 0; j < 4; ++j) M.out[j] = tot[j];
      *M.nan_flag = nan ? 1 : 0;
    }
  }
}

// Chain rule for the multi-level outputs.  g = upstream gradient of out (device, 4+5L floats).  Level l's
// box/objectness/class groups are scaled by g[0] + g[1+j] + g[4+4l] + g[5+4l+j] + g[4+4L+l], j = 0,1,2.
struct MultiScaleParams {
  float* grad[PQDET_MAX_LEVELS];
  int64_t total[PQDET_MAX_LEVELS];
  int HW[PQDET_MAX_LEVELS];
  int n_levels, ch;
  const float* g;
};

__global__ void __launch_bounds__(256)
scale_levels_kernel(const __grid_constant__ MultiScaleParams S) {
  const int l = blockIdx.y;
  const int L = S.n_levels;
  const float base = S.g[0] + S.g[4 + 4 * l] + S.g[4 + 4 * L + l];
  const float cb = base + S.g[1] + S.g[5 + 4 * l];
  const float cc = base + S.g[2] + S.g[6 + 4 * l];
  const float cp = base + S.g[3] + S.g[7 + 4 * l];
  if (cb == 1.0f && cc == 1.0f && cp == 1.0f) return;
  float* grad = S.grad[l];
  const int HW = S.HW[l], ch = S.ch;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < S.total[l]; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)((e / HW) % ch);
    grad[e] *= (k < 4) ? cb : (k == 4 ? cc : cp);
  }
}

// out4 = [loss, bbox, conf, cls] (batch means), nan_flag = isnan(loss).
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const double* __restrict__ partials, int64_t n, double inv_B, float* out4, int32_t* nan_flag) {
  __shared__ double s[256][3];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    acc[0] += partials[i * 3 + 0];
    acc[1] += partials[i * 3 + 1];
    acc[2] += partials[i * 3 + 2];
  }
  for (int j = 0; j < 3; ++j) s[threadIdx.x][j] = acc[j];
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d)
      for (int j = 0; j < 3; ++j) s[threadIdx.x][j] += s[threadIdx.x + d][j];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float lb = (float)(s[0][0] * inv_B);
    const float lc = (float)(s[0][1] * inv_B);
    const float lp = (float)(s[0][2] * inv_B);
    const float loss = PQ_ADD(PQ_ADD(lb, lc), lp);          // model/loss.py:108
    out4[0] = loss; out4[1] = lb; out4[2] = lc; out4[3] = lp;
    *nan_flag = (loss != loss) ? 1 : 0;
  }
}


// Upstream-gradient mix for the autograd wrapper.  grad holds d bbox / d conf / d cls in its box /
// objectness / class channel groups; the chain rule for upstream (g_loss, g_bbox, g_conf, g_cls) is
// a per-group scale by (g_loss + g_x).  The coefficients are read on the device; the usual case
// (all three == 1, e.g. loss.mean().backward()) returns without touching memory - no host sync.
__global__ void __launch_bounds__(256)
scale_groups_kernel(float* __restrict__ grad, int64_t total, int is_raw, int ch, int HW,
                    const float* g_loss, const float* g_bbox, const float* g_conf, const float* g_cls) {
  const float gl = g_loss ? *g_loss : 0.0f;
  const float cb = gl + (g_bbox ? *g_bbox : 0.0f);
  const float cc = gl + (g_conf ? *g_conf : 0.0f);
  const float cp = gl + (g_cls ? *g_cls : 0.0f);
  if (cb == 1.0f && cc == 1.0f && cp == 1.0f) return;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = is_raw ? (int)((e / HW) % ch) : (int)(e % ch);
    grad[e] *= (k < 4) ? cb : (k == 4 ? cc : cp);
  }
}

}  // namespace pq

extern "C" int64_t pqdet_loss_workspace(int B, int A, int H, int W) {
  if (B < 0 || A < 1 || H < 1 || W < 1) return PQDET_ERR_INVALID_ARG;
  const int64_t tiles = ((int64_t)H * W + pq::kLossTile - 1) / pq::kLossTile;
  return (int64_t)B * tiles * 3 * (int64_t)sizeof(double) + 256;
}

extern "C" int pqdet_loss_fwd_bwd(const float* x, int input_is_raw, const float* label, const float* gt,
                                  float* grad, float* out4, int32_t* nan_flag, void* partials,
                                  int B, int A, int C, int H, int W, int G, float stride, int bbox_loss,
                                  float ignore_thresh, float l1_loss_gain, int device, void* stream) {
  using namespace pq;
  if (!x || !label || !gt || !out4 || !nan_flag || !partials) return PQDET_ERR_INVALID_ARG;
  if (B < 1 || A < 1 || A > 8 || C < 1 || H < 1 || W < 1 || G < 1) return PQDET_ERR_INVALID_ARG;
  if (bbox_loss < 0 || bbox_loss > 3) return PQDET_ERR_UNSUPPORTED;   // ciou: the reference always raises
  if (B > 65535) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  LossParams P;
  P.x = x; P.label = label; P.gt = gt; P.grad = grad; P.partials = (double*)partials;
  P.B = B; P.A = A; P.C = C; P.H = H; P.W = W; P.G = G;
  P.stride = stride;
  P.in_area = (float)((double)(stride * H) * (double)(stride * W));   // Python int product, loss.py:45,58
  P.ignore_thresh = ignore_thresh; P.l1_gain = l1_loss_gain; P.inv_B = 1.0f / (float)B;
  P.bbox_loss = bbox_loss;
  const int tiles = (H * W + kLossTile - 1) / kLossTile;
  const size_t smem = (size_t)A * 32 * 5 * sizeof(float);
  dim3 grid(tiles, B);
  if (input_is_raw) {
    loss_fwd_bwd_kernel<true><<<grid, 32 * A, smem, st>>>(P);
  } else {
    loss_fwd_bwd_kernel<false><<<grid, 32 * A, smem, st>>>(P);
  }
  PQ_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 256, 0, st>>>((const double*)partials, (int64_t)B * tiles, 1.0 / (double)B, out4, nan_flag);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

extern "C" int pqdet_loss_scale_grad(float* grad, int input_is_raw, int B, int A, int C, int H, int W,
                                     const float* g_loss, const float* g_bbox, const float* g_conf,
                                     const float* g_cls, int device, void* stream) {
  if (!grad || B < 1 || A < 1 || C < 1 || H < 1 || W < 1) return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  const int64_t total = (int64_t)B * A * (5 + C) * H * W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 2) blocks = 148 * 2;
  pq::scale_groups_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      grad, total, input_is_raw, 5 + C, H * W, g_loss, g_bbox, g_conf, g_cls);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

extern "C" int64_t pqdet_loss_levels_workspace(int n_levels, int B, int A, const int* H, const int* W) {
  if (n_levels < 1 || n_levels > PQDET_MAX_LEVELS || B < 0 || A < 1 || !H || !W) return PQDET_ERR_INVALID_ARG;
  int64_t bytes = 256;
  for (int l = 0; l < n_levels; ++l) {
    const int64_t tiles = ((int64_t)H[l] * W[l] + pq::kLossTile - 1) / pq::kLossTile;
    bytes += ((int64_t)B * tiles * 3 * (int64_t)sizeof(double) + 255) / 256 * 256;
  }
  return bytes;
}

namespace pq {
static int loss_levels_impl(int n_levels, const float* const* raw, const float* const* label,
                            const int32_t* const* owner, const float* gt6, int n_max,
                            const float* const* gt, float* const* grad, const int* H, const int* W,
                            const int* G, const float* stride, int B, int A, int C, int bbox_loss,
                            float ignore_thresh, float l1_loss_gain, float* out, int32_t* nan_flag,
                            void* workspace, int workspace_initialised, int device, void* stream) {
  const bool sparse = owner != nullptr;
  if (n_levels < 1 || n_levels > PQDET_MAX_LEVELS || !raw || (!label && !owner) || !gt || !H || !W || !G || !stride ||
      !out || !nan_flag || !workspace)
    return PQDET_ERR_INVALID_ARG;
  if (sparse && (n_max < 0 || (n_max > 0 && !gt6))) return PQDET_ERR_INVALID_ARG;
  if (B < 1 || B > 65535 || A < 1 || A > 8 || C < 1) return PQDET_ERR_INVALID_ARG;
  if (bbox_loss < 0 || bbox_loss > 3) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  MultiLossParams M;
  memset(&M, 0, sizeof(M));
  unsigned char* ws = (unsigned char*)workspace;
  (void)workspace_initialised;                  // the workspace carries no state between calls
  size_t off = 256;
  int tiles_total = 0;
  const double deta = 0.01, uni = 1.0 / (double)C;         // train_dataset.py:126-130, fp64 then stored fp32
  for (int l = 0; l < n_levels; ++l) {
    if (!raw[l] || !gt[l] || H[l] < 1 || W[l] < 1 || G[l] < 1) return PQDET_ERR_INVALID_ARG;
    if (sparse ? !owner[l] : !label[l]) return PQDET_ERR_INVALID_ARG;
    LossParams& P = M.lv[l];
    P.x = raw[l]; P.label = sparse ? nullptr : label[l]; P.gt = gt[l]; P.grad = grad ? grad[l] : nullptr;
    P.owner = sparse ? owner[l] : nullptr; P.gt6 = gt6; P.n_max = n_max;
    P.hot = (float)(1.0 * (1 - deta) + deta * uni);
    P.cold = (float)(0.0 * (1 - deta) + deta * uni);
    P.partials = (double*)(ws + off);
    const int tiles = (H[l] * W[l] + kLossTile - 1) / kLossTile;
    off += ((size_t)B * tiles * 3 * sizeof(double) + 255) / 256 * 256;
    P.B = B; P.A = A; P.C = C; P.H = H[l]; P.W = W[l]; P.G = G[l];
    P.stride = stride[l];
    P.in_area = (float)((double)(stride[l] * H[l]) * (double)(stride[l] * W[l]));
    P.ignore_thresh = ignore_thresh; P.l1_gain = l1_loss_gain; P.inv_B = 1.0f / (float)B;
    P.bbox_loss = bbox_loss;
    M.tile_off[l] = tiles_total;
    tiles_total += tiles;
  }
  for (int l = n_levels; l <= PQDET_MAX_LEVELS; ++l) M.tile_off[l] = tiles_total;
  M.n_levels = n_levels;
  M.out = out; M.nan_flag = nan_flag;
  const size_t smem = (size_t)A * 32 * 5 * sizeof(float);
  dim3 grid(tiles_total, B);
  if (sparse) loss_levels_kernel<true><<<grid, 32 * A, smem, st>>>(M);
  else loss_levels_kernel<false><<<grid, 32 * A, smem, st>>>(M);
  PQ_LAUNCH_CHECK();
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PQ_CUDA(cudaLaunchKernelEx(&cfg, loss_levels_finalize_kernel, M));
  }
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
}  // namespace pq

extern "C" int pqdet_loss_levels(int n_levels, const float* const* raw, const float* const* label,
                                 const float* const* gt, float* const* grad, const int* H, const int* W,
                                 const int* G, const float* stride, int B, int A, int C, int bbox_loss,
                                 float ignore_thresh, float l1_loss_gain, float* out, int32_t* nan_flag,
                                 void* workspace, int workspace_initialised, int device, void* stream) {
  if (!label) return PQDET_ERR_INVALID_ARG;
  return pq::loss_levels_impl(n_levels, raw, label, nullptr, nullptr, 0, gt, grad, H, W, G, stride, B, A, C,
                              bbox_loss, ignore_thresh, l1_loss_gain, out, nan_flag, workspace,
                              workspace_initialised, device, stream);
}

extern "C" int pqdet_loss_levels_sparse(int n_levels, const float* const* raw, const int32_t* const* owner,
                                        const float* gt6, int n_max, const float* const* gtlist,
                                        float* const* grad, const int* H, const int* W, const int* G,
                                        const float* stride, int B, int A, int C, int bbox_loss,
                                        float ignore_thresh, float l1_loss_gain, float* out, int32_t* nan_flag,
                                        void* workspace, int workspace_initialised, int device, void* stream) {
  if (!owner) return PQDET_ERR_INVALID_ARG;
  if (A != 3) return PQDET_ERR_UNSUPPORTED;          // the assignment has 3 anchors per scale (config.py:57)
  return pq::loss_levels_impl(n_levels, raw, nullptr, owner, gt6, n_max, gtlist, grad, H, W, G, stride, B, A, C,
                              bbox_loss, ignore_thresh, l1_loss_gain, out, nan_flag, workspace,
                              workspace_initialised, device, stream);
}

extern "C" int pqdet_loss_levels_scale_grad(int n_levels, float* const* grad, const int* H, const int* W,
                                            int B, int A, int C, const float* upstream, int device, void* stream) {
  using namespace pq;
  if (n_levels < 1 || n_levels > PQDET_MAX_LEVELS || !grad || !H || !W || !upstream) return PQDET_ERR_INVALID_ARG;
  if (B < 1 || A < 1 || C < 1) return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  MultiScaleParams S;
  memset(&S, 0, sizeof(S));
  int64_t mx = 0;
  for (int l = 0; l < n_levels; ++l) {
    if (!grad[l]) return PQDET_ERR_INVALID_ARG;
    S.grad[l] = grad[l];
    S.HW[l] = H[l] * W[l];
    S.total[l] = (int64_t)B * A * (5 + C) * H[l] * W[l];
    if (S.total[l] > mx) mx = S.total[l];
  }
  S.n_levels = n_levels; S.ch = 5 + C; S.g = upstream;
  int64_t blocks = (mx + 255) / 256;
  if (blocks > 148) blocks = 148;            // usually a no-op (all factors 1): keep the launch tiny
  dim3 grid((unsigned)blocks, n_levels);
  scale_levels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
