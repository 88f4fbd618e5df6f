// Elementwise IoU family: tools.py:357-477.
// Memory-bound streaming kernels: 32 B in, 4 B out per pair; float4 loads.
#include "pq_common.cuh"

namespace pq {

__global__ void __launch_bounds__(256)
iou_pairwise_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, float* __restrict__ out,
                    int64_t n, int kind) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p4 = b1[i], q4 = b2[i];
    const float p[4] = {p4.x, p4.y, p4.z, p4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
    out[i] = iou_value(kind, p, q);
  }
}

__global__ void __launch_bounds__(256)
iou_pairwise_bwd_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2,
                        const float* __restrict__ gout, float4* __restrict__ g1, float4* __restrict__ g2,
                        int64_t n, int kind) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p4 = b1[i], q4 = b2[i];
    const float p[4] = {p4.x, p4.y, p4.z, p4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
    float gp[4], gq[4];
    iou_value_grad(kind, p, q, gp, gq);
    const float g = gout[i];
    if (g1) g1[i] = make_float4(gp[0] * g, gp[1] * g, gp[2] * g, gp[3] * g);
    if (g2) g2[i] = make_float4(gq[0] * g, gq[1] * g, gq[2] * g, gq[3] * g);
  }
}

static unsigned grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  return (unsigned)blocks;
}

}  // namespace pq

extern "C" int pqdet_iou_pairwise(const float* b1, const float* b2, float* out, int64_t n, int kind,
                                  int device, void* stream) {
  if (!b1 || !b2 || !out || n < 0 || kind < 0 || kind > 4) return PQDET_ERR_INVALID_ARG;
  if (((uintptr_t)b1 | (uintptr_t)b2) & 15) return PQDET_ERR_INVALID_ARG;
  if (n == 0) return PQDET_OK;
  PQ_ENTER(device);
  pq::iou_pairwise_kernel<<<pq::grid_for(n), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)b1, (const float4*)b2, out, n, kind);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

extern "C" int pqdet_iou_pairwise_bwd(const float* b1, const float* b2, const float* grad_out, float* grad_b1,
                                      float* grad_b2, int64_t n, int kind, int device, void* stream) {
  if (!b1 || !b2 || !grad_out || n < 0 || kind < 0 || kind > 3) return PQDET_ERR_INVALID_ARG;
  if (((uintptr_t)b1 | (uintptr_t)b2 | (uintptr_t)grad_b1 | (uintptr_t)grad_b2) & 15) return PQDET_ERR_INVALID_ARG;
  if (n == 0) return PQDET_OK;
  PQ_ENTER(device);
  pq::iou_pairwise_bwd_kernel<<<pq::grid_for(n), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)b1, (const float4*)b2, grad_out, (float4*)grad_b1, (float4*)grad_b2, n, kind);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
