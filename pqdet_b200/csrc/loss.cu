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
#include "pq_common.cuh"

namespace pq {

constexpr int kLossTile = 32;
constexpr int kGtChunk = 128;

struct LossParams {
  const float* x;       // raw (B, A*ch, H, W) or decoded (B,H,W,A,ch)
  const float* label;   // (B,H,W,A,6+C)
  const float* gt;      // (B,G,4)
  float* grad;          // same shape as x, nullable
  double* partials;     // [B * tiles][3]
  int B, A, C, H, W, G;
  float stride, in_area, ignore_thresh, l1_gain, inv_B;
  int bbox_loss;
};

template <bool RAW>
__global__ void __launch_bounds__(256)
loss_fwd_bwd_kernel(const __grid_constant__ LossParams P) {
  extern __shared__ __align__(16) float smem[];
  const int A = P.A, C = P.C, ch = 5 + C, LW = 6 + C, LWp = LW | 1;
  const int HW = P.H * P.W;
  float* slab = smem;                                   // [32*A][LWp]
  float* sgt = smem + kLossTile * A * LWp;              // [kGtChunk][5]
  __shared__ double sred[8][3];
  const int b = blockIdx.y;
  const int cell0 = blockIdx.x * kLossTile;
  const int ncell = min(kLossTile, HW - cell0);
  const int lane = lane_id(), a = warp_id();
  const int cell = cell0 + lane;
  const bool active = lane < ncell;

  // ---- stage the label tile (contiguous in global memory) ----
  {
    const float* src = P.label + ((size_t)b * HW + cell0) * A * LW;
    const int n = ncell * A * LW;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int r = e / LW, k = e - r * LW;
      slab[r * LWp + k] = ldg_stream(src + e);
    }
  }
  // ---- prediction: box + objectness ----
  float pb[4] = {0.f, 0.f, 1.f, 1.f}, es[4] = {0.f, 0.f, 0.f, 0.f}, pconf = 0.5f;
  const size_t plane0 = ((size_t)b * A + a) * ch * HW;             // RAW: first plane of this anchor
  const size_t prow = (((size_t)b * HW + cell) * A + a) * ch;      // !RAW: this row
  if (active) {
    if (RAW) {
      const int cy = cell / P.W, cx = cell - cy * P.W;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float v = ldg_stream(P.x + plane0 + (size_t)k * HW + cell);
        pb[k] = decode_coord(k, v, cx, cy, P.stride);
        const float e = expf(v) * P.stride;
        es[k] = (k < 2) ? -e : e;                                  // d pb[k] / d raw[k]
      }
      pconf = sigmoidf_(ldg_stream(P.x + plane0 + (size_t)4 * HW + cell));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) pb[k] = P.x[prow + k];
      pconf = P.x[prow + 4];
    }
  }
  __syncthreads();
  const float* lab = slab + (lane * A + a) * LWp;
  float tb[4] = {0.f, 0.f, 0.f, 0.f}, respond = 0.f, mixw = 0.f;
  if (active) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tb[k] = lab[k];
    respond = lab[4];
    mixw = lab[5 + C];
  }
  // ---- box term ----
  float dbox[4] = {0.f, 0.f, 0.f, 0.f};
  float lb = 0.f;
  if (active) lb = bbox_loss_row(P.bbox_loss, pb, tb, respond, P.in_area, P.l1_gain, dbox);

  // ---- ignore mask: every GT has iou < thr (NaN -> false), only needed where respond != 1 ----
  bool below = true;
  const bool need = active && (respond != 1.0f);
  const float a1 = box_area(pb[0], pb[1], pb[2], pb[3]);
  for (int g0 = 0; g0 < P.G; g0 += kGtChunk) {
    const int ng = min(kGtChunk, P.G - g0);
    __syncthreads();
    for (int e = threadIdx.x; e < ng; e += blockDim.x) {
      const float* q = P.gt + ((size_t)b * P.G + g0 + e) * 4;
      const float x1 = q[0], y1 = q[1], x2 = q[2], y2 = q[3];
      sgt[e * 5 + 0] = x1; sgt[e * 5 + 1] = y1; sgt[e * 5 + 2] = x2; sgt[e * 5 + 3] = y2;
      sgt[e * 5 + 4] = box_area(x1, y1, x2, y2);
    }
    __syncthreads();
    if (need && below) {
      for (int g = 0; g < ng; ++g) {
        const float* q = sgt + g * 5;
        if (!iou_below(pb[0], pb[1], pb[2], pb[3], a1, q[0], q[1], q[2], q[3], q[4], P.ignore_thresh)) {
          below = false;
          break;
        }
      }
    }
  }
  // torch.max over an empty GT axis would raise; collate always pads to G >= 1.
  const float bgd = PQ_MUL(PQ_SUB(1.0f, respond), below ? 1.0f : 0.0f);

  // ---- objectness term ----
  float dconf = 0.f, lc = 0.f;
  if (active) lc = focal_bce_term(1.0f, 0.75f, respond, pconf, respond, bgd, &dconf);

  const float gw = mixw * P.inv_B;
  // ---- class term (+ gradient stores for every class channel) ----
  float lp = 0.f;
  for (int c = 0; c < C; ++c) {
    float g = 0.f;
    if (active && respond != 0.0f) {
      const float t = lab[5 + c];
      float p, dp;
      if (RAW) p = sigmoidf_(P.x[plane0 + (size_t)(5 + c) * HW + cell]);
      else p = P.x[prow + 5 + c];
      lp = PQ_ADD(lp, focal_bce_term(2.0f, 0.5f, t, p, respond, -1.0f, &dp));
      g = RAW ? dp * (p * (1.0f - p)) * gw : dp * gw;
    }
    if (active && P.grad) {
      if (RAW) P.grad[plane0 + (size_t)(5 + c) * HW + cell] = g;
      else P.grad[prow + 5 + c] = g;
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

  // ---- partial sums (fp64, fixed order) ----
  double v0 = active ? (double)PQ_MUL(lb, mixw) : 0.0;
  double v1 = active ? (double)PQ_MUL(lc, mixw) : 0.0;
  double v2 = active ? (double)PQ_MUL(lp, mixw) : 0.0;
  v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2);
  if (lane == 0) { sred[a][0] = v0; sred[a][1] = v1; sred[a][2] = v2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < A; ++w) s += sred[w][threadIdx.x];
    P.partials[((size_t)b * gridDim.x + blockIdx.x) * 3 + threadIdx.x] = s;
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
  const size_t smem = ((size_t)kLossTile * A * ((6 + C) | 1) + kGtChunk * 5) * sizeof(float);
  dim3 grid(tiles, B);
  if (input_is_raw) {
    if (smem > 48 * 1024)
      PQ_CUDA(cudaFuncSetAttribute(loss_fwd_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    loss_fwd_bwd_kernel<true><<<grid, 32 * A, smem, st>>>(P);
  } else {
    if (smem > 48 * 1024)
      PQ_CUDA(cudaFuncSetAttribute(loss_fwd_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
  if (blocks > 148 * 16) blocks = 148 * 16;
  pq::scale_groups_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      grad, total, input_is_raw, 5 + C, H * W, g_loss, g_bbox, g_conf, g_cls);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
