// AP matching of the evaluator: the per-detection Python loop of
// eval/evaluator.py:69-124 as one kernel.  The host keeps the reference's bookkeeping (detections ordered per
// class by (-score, insertion index), labels grouped per (image, class) with numpy's own argsort) and the final
// cumsum / precision-recall / AP arithmetic, which the reference already does vectorised in numpy; what moves to
// the GPU is the sequential matching: one thread per ((image, class) group, IoU threshold) walks the group's
// detections in score order and reproduces the reference's pick / seen / difficult logic statement by statement.
// Overlaps are evaluated in the dtype numpy would use (GT float32 -> float32 throughout; GT float64 -> the
// detection's own area stays float32, everything touching the GT is float64), one rounding per operation.
#include "pq_common.cuh"

namespace pq {

struct ApParams {
  const float* det_box;        // (D, 4) detections in class-sorted order
  const int32_t* grp_det;      // detection indices (into det_box) ordered by (group, class rank)
  const int64_t* grp_det_off;  // (G + 1)
  const void* gt_box;          // (sumGT, 4) float or double, groups back to back, rows in the reference's order
  const uint8_t* gt_diff;      // (sumGT)
  const int64_t* gt_off;       // (G + 1)
  const double* thr;           // (T) IoU thresholds
  uint8_t* seen;               // (T, sumGT) scratch, zero on entry
  uint8_t* tp;                 // (T, D) zero on entry
  uint8_t* fp;                 // (T, D) zero on entry
  int64_t D, sumGT;
  int T, G;
};

// np.maximum / np.minimum: NaN if either operand is NaN
template <typename T> __device__ __forceinline__ T np_max(T a, T b) { return (a > b || a != a) ? a : b; }
template <typename T> __device__ __forceinline__ T np_min(T a, T b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ float r_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float r_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float r_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float r_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double r_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double r_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double r_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double r_div(double a, double b) { return __ddiv_rn(a, b); }

// eval/evaluator.py:77-91 for one (detection, GT) pair.  area_b = (bb2-bb0+1)*(bb3-bb1+1) in float32.
template <typename T>
__device__ __forceinline__ double ap_overlap(const T* g, const float* bb, float area_b) {
  const T one = (T)1, zero = (T)0;
  const T ixmin = np_max<T>(g[0], (T)bb[0]), iymin = np_max<T>(g[1], (T)bb[1]);
  const T ixmax = np_min<T>(g[2], (T)bb[2]), iymax = np_min<T>(g[3], (T)bb[3]);
  const T iw = np_max<T>(r_add(r_sub(ixmax, ixmin), one), zero);
  const T ih = np_max<T>(r_add(r_sub(iymax, iymin), one), zero);
  const T inters = r_mul(iw, ih);
  const T area_g = r_mul(r_add(r_sub(g[2], g[0]), one), r_add(r_sub(g[3], g[1]), one));
  const T uni = r_sub(r_add((T)area_b, area_g), inters);
  return (double)r_div(inters, uni);
}

template <typename T>
__global__ void __launch_bounds__(128)
ap_match_kernel(const __grid_constant__ ApParams P) {
  const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (int64_t)P.G * P.T) return;
  const int g = (int)(id / P.T), t = (int)(id - (int64_t)g * P.T);
  const int64_t go = P.gt_off[g];
  const int n = (int)(P.gt_off[g + 1] - go);
  const T* gt = reinterpret_cast<const T*>(P.gt_box) + go * 4;
  const uint8_t* diff = P.gt_diff + go;
  uint8_t* seen = P.seen + (int64_t)t * P.sumGT + go;
  const double thr = fmin(P.thr[t], 1.0 - 1e-10);                       // :97
  for (int64_t q = P.grp_det_off[g]; q < P.grp_det_off[g + 1]; ++q) {
    const int64_t d = P.grp_det[q];
    const float* bb = P.det_box + d * 4;
    const float area_b = r_mul(r_add(r_sub(bb[2], bb[0]), 1.0f), r_add(r_sub(bb[3], bb[1]), 1.0f));
    int pick = -1;
    double pick_iou = thr;
    for (int j = 0; j < n; ++j) {                                       // :98-106
      if (seen[j]) continue;
      if (pick > -1 && !diff[pick] && diff[j]) break;
      const double ov = ap_overlap<T>(gt + (int64_t)j * 4, bb, area_b);
      if (ov < pick_iou) continue;                                      // NaN compares false, like numpy
      pick = j;
      pick_iou = ov;
    }
    if (diff[pick < 0 ? n - 1 : pick]) continue;                        // :107 (index -1 = the last GT)
    if (pick == -1) { P.fp[(int64_t)t * P.D + d] = 1; continue; }       // :109-111 (a picked GT is never `seen`)
    P.tp[(int64_t)t * P.D + d] = 1;
    seen[pick] = 1;
  }
}

}  // namespace pq

extern "C" int pqdet_ap_match(const float* det_box, int64_t D, const int32_t* grp_det, const int64_t* grp_det_off,
                              const void* gt_box, int gt_is_f64, const uint8_t* gt_difficult, const int64_t* gt_off,
                              int64_t sum_gt, int G, const double* thresholds, int T, uint8_t* seen, uint8_t* tp,
                              uint8_t* fp, int device, void* stream) {
  using namespace pq;
  if (D < 0 || G < 0 || T < 1 || sum_gt < 0) return PQDET_ERR_INVALID_ARG;
  if (G == 0 || D == 0) return PQDET_OK;
  if (!det_box || !grp_det || !grp_det_off || !gt_box || !gt_difficult || !gt_off || !thresholds || !seen || !tp || !fp)
    return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  ApParams P;
  P.det_box = det_box; P.grp_det = grp_det; P.grp_det_off = grp_det_off; P.gt_box = gt_box;
  P.gt_diff = gt_difficult; P.gt_off = gt_off; P.thr = thresholds; P.seen = seen; P.tp = tp; P.fp = fp;
  P.D = D; P.sumGT = sum_gt; P.T = T; P.G = G;
  const int64_t threads = (int64_t)G * T;
  const unsigned blocks = (unsigned)((threads + 127) / 128);
  if (gt_is_f64) ap_match_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>(P);
  else ap_match_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(P);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
