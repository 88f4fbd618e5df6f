// tools.nms (tools.py:507-538): the reference's numpy per-class NMS with the optional soft-NMS score decay.
// It has no callers in the reference; kept for API completeness, with the
// reference's semantics in full: per class, pick the best remaining box (first index on ties), emit it with
// the score it has at that moment - the first pick of a class is NOT compared with the score threshold -,
// then weight the rest (hard: 0 where iou_calc1 > iou_threshold; soft: exp(-iou^2 / sigma)) and drop what
// falls to score <= score_threshold.  All arithmetic fp32, iou_calc1's union clamp at 1e-14 included.
//
// One CTA per class; the class' boxes sit in a contiguous segment (the host groups by class).  The selection
// loop is inherently serial in the picks; a pick costs one block arg-max and one pass over the segment.
#include "pq_common.cuh"

namespace pq {

constexpr int kSoftThreads = 256;

__global__ void __launch_bounds__(kSoftThreads)
classwise_nms_kernel(const float4* __restrict__ boxes, float* __restrict__ scores, const int32_t* __restrict__ seg_off,
                     int soft, float sigma, float score_thr, float iou_thr, int32_t* __restrict__ out_idx,
                     float* __restrict__ out_score, int32_t* __restrict__ out_count, uint8_t* __restrict__ alive) {
  const int c = blockIdx.x;
  const int lo = seg_off[c], hi = seg_off[c + 1];
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  __shared__ float s_val[kSoftThreads / 32];
  __shared__ int s_idx[kSoftThreads / 32];
  __shared__ int s_best;
  for (int i = lo + tid; i < hi; i += kSoftThreads) alive[i] = 1;
  __syncthreads();
  int k = 0;
  for (;;) {
    // arg-max over the alive boxes: highest score, lowest index on ties (np.argmax returns the first maximum and
    // the reference's np.concatenate keeps the relative order of the remaining rows)
    float bv = -INFINITY;
    int bi = -1;
    for (int i = lo + tid; i < hi; i += kSoftThreads)
      if (alive[i]) {
        const float v = scores[i];
        if (bi < 0 || v > bv) { bv = v; bi = i; }              // (NaN scores are not supported: they never win)
      }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float ov = __shfl_xor_sync(PQ_FULL, bv, d);
      const int oi = __shfl_xor_sync(PQ_FULL, bi, d);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = bv; s_idx[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = s_val[0];
      int ix = s_idx[0];
      for (int w = 1; w < kSoftThreads / 32; ++w) {
        const float ov = s_val[w];
        const int oi = s_idx[w];
        if (oi >= 0 && (ix < 0 || ov > v || (ov == v && oi < ix))) { v = ov; ix = oi; }
      }
      s_best = ix;
      if (ix >= 0) {
        out_idx[lo + k] = ix;
        out_score[lo + k] = v;
        alive[ix] = 0;
      }
    }
    __syncthreads();
    const int best = s_best;
    if (best < 0) break;
    ++k;
    const float4 bb = boxes[best];
    const float pb[4] = {bb.x, bb.y, bb.z, bb.w};
    for (int i = lo + tid; i < hi; i += kSoftThreads)
      if (alive[i]) {
        const float4 q4 = boxes[i];
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
        const float iou = iou_value(4, pb, q);                       // iou_calc1(best, rest)
        float w = 1.0f;
        if (soft) w = expf(-PQ_DIV(PQ_MUL(1.0f, PQ_MUL(iou, iou)), sigma));   // np.exp(-(1.0 * iou ** 2 / sigma))
        else if (iou > iou_thr) w = 0.0f;
        const float sc = PQ_MUL(scores[i], w);
        scores[i] = sc;
        if (!(sc > score_thr)) alive[i] = 0;
      }
    __syncthreads();
  }
  if (tid == 0) out_count[c] = k;
}

}  // namespace pq

extern "C" int pqdet_classwise_nms(const float* boxes, float* scores, const int32_t* seg_off, int n_classes, int64_t n,
                                   int soft, double sigma, double score_threshold, double iou_threshold,
                                   int32_t* out_idx, float* out_score, int32_t* out_count, uint8_t* alive_scratch,
                                   int device, void* stream) {
  if (!seg_off || !out_count || n_classes < 0 || n < 0) return PQDET_ERR_INVALID_ARG;
  if (n > 0 && (!boxes || !scores || !out_idx || !out_score || !alive_scratch)) return PQDET_ERR_INVALID_ARG;
  if ((uintptr_t)boxes & 15) return PQDET_ERR_INVALID_ARG;
  if (soft && !(sigma > 0.0)) return PQDET_ERR_INVALID_ARG;
  if (n_classes == 0) return PQDET_OK;
  PQ_ENTER(device);
  pq::classwise_nms_kernel<<<n_classes, pq::kSoftThreads, 0, (cudaStream_t)stream>>>(
      (const float4*)boxes, scores, seg_off, soft ? 1 : 0, (float)sigma, (float)score_threshold, (float)iou_threshold,
      out_idx, out_score, out_count, alive_scratch);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
