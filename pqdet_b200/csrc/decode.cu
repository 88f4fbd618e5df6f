// Materialising decode / decode-backward / recover kernels.
//
// decode_fwd transposes NCHW planes into (cell, anchor, channel) rows through shared memory:
// each warp reads 32 consecutive cells of one channel plane (one 128-byte line), applies the
// decode function, and the tile is written back as one contiguous run of 32*A*(5+C) floats.
// HBM traffic is the algorithmic 2R (read once, write once); there is no reuse to exploit.
#include "pq_common.cuh"

namespace pq {

constexpr int kTileCells = 32;
constexpr int kDecodeThreads = 256;

__global__ void __launch_bounds__(kDecodeThreads)
decode_fwd_kernel(const float* __restrict__ raw, float* __restrict__ out, int A, int ch, int H, int W,
                  float stride, int64_t rows_total, int64_t row_off) {
  extern __shared__ __align__(16) float tile[];  // [kTileCells][ST]
  const int HW = H * W;
  const int ACH = A * ch;
  const int ST = ACH | 1;  // odd row stride: conflict-free column writes (== ACH whenever ACH is odd)
  const int b = blockIdx.y;
  const int cell0 = blockIdx.x * kTileCells;
  const int ncell = min(kTileCells, HW - cell0);
  const int lane = lane_id();
  const int cell = cell0 + lane;
  const int cy = cell / W, cx = cell - cy * W;
  const float* src = raw + (size_t)b * ACH * HW + cell;
  float* tcol = tile + lane * ST;
  // The kernel is issue bound before it is HBM bound (ncu: 81 % of the issue slots at 2.3 TB/s), so the loop
  // carries the channel-within-anchor index along instead of dividing, and 1/(1+e) uses the correctly rounded
  // reciprocal (bit-identical to the IEEE division of sigmoidf_, fewer instructions).
  constexpr int W8 = kDecodeThreads / 32;
  int k = warp_id() % ch;
  for (int c = warp_id(); c < ACH; c += W8) {
    if (lane < ncell) {
      const float v = ldg_stream(src + (size_t)c * HW);
      tcol[c] = (k < 4) ? decode_coord(k, v, cx, cy, stride) : __frcp_rn(PQ_ADD(1.0f, expf(-v)));
    }
    k += W8;
    while (k >= ch) k -= ch;
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

// recover: a CTA takes kRecRows consecutive rows of one image.  The input rows are one contiguous run of
// kRecRows*(5+C) floats: staged in shared memory with coalesced loads, then every warp sweeps the 4+C output values
// of its rows (contiguous in the output as well).  The affine of the image is evaluated once per CTA.
constexpr int kRecRows = 64;
constexpr int kRecThreads = 256;

__global__ void __launch_bounds__(kRecThreads)
recover_kernel(const float* __restrict__ pred, float* __restrict__ out, int64_t N, int C, int kind,
               float in_h, float in_w, const float* __restrict__ orig_hw, int orig_per_image) {
  extern __shared__ float rtile[];          // [kRecRows][5+C]
  __shared__ Affine s_af;
  const int ic = 5 + C, oc = 4 + C;
  const int b = blockIdx.y;
  const int64_t row0 = (int64_t)blockIdx.x * kRecRows;
  const int nrow = (int)min((int64_t)kRecRows, N - row0);
  if (threadIdx.x == 0) {
    const float* o = orig_hw + (orig_per_image ? 2 * b : 0);
    s_af = affine_params(kind, in_h, in_w, o[0], o[1]);
  }
  const float* src = pred + ((size_t)b * N + row0) * ic;
  const int n_in = nrow * ic;
  for (int e = threadIdx.x; e < n_in; e += kRecThreads) rtile[e] = ldg_stream(src + e);
  __syncthreads();
  const Affine af = s_af;
  float* dst = out + ((size_t)b * N + row0) * oc;
  const int lane = lane_id();
  for (int r = warp_id(); r < nrow; r += kRecThreads / 32) {
    const float* trow = rtile + r * ic;
    const float conf = trow[4];
    float* drow = dst + (size_t)r * oc;
    for (int k = lane; k < oc; k += 32)
      drow[k] = (k < 4) ? recover_coord(k, trow[k], af) : PQ_MUL(trow[k + 1], conf);
  }
}

}  // namespace pq

extern "C" int pqdet_decode_fwd(const float* raw, float* out, int B, int A, int C, int H, int W, float stride,
                                int64_t out_rows_total, int64_t out_row_offset, int device, void* stream) {
  if (B < 0 || A <= 0 || C < 0 || H <= 0 || W <= 0) return PQDET_ERR_INVALID_ARG;
  if (out_row_offset < 0 || out_row_offset + (int64_t)H * W * A > out_rows_total) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;                       // empty batch: tensors without storage are fine
  if (!raw || !out) return PQDET_ERR_INVALID_ARG;
  if (B > 65535) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  const int ch = 5 + C;
  const size_t smem = (size_t)pq::kTileCells * ((A * ch) | 1) * sizeof(float);
  if (smem > 48 * 1024)
    PQ_CUDA(cudaFuncSetAttribute(pq::decode_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((H * W + pq::kTileCells - 1) / pq::kTileCells, B);
  pq::decode_fwd_kernel<<<grid, pq::kDecodeThreads, smem, (cudaStream_t)stream>>>(
      raw, out, A, ch, H, W, stride, out_rows_total, out_row_offset);
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
  const size_t smem = (size_t)pq::kRecRows * (5 + C) * sizeof(float);
  if (smem > 48 * 1024)
    PQ_CUDA(cudaFuncSetAttribute(pq::recover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((N + pq::kRecRows - 1) / pq::kRecRows), B);
  pq::recover_kernel<<<grid, pq::kRecThreads, smem, (cudaStream_t)stream>>>(
      pred, out, N, C, affine_kind, in_h, in_w, orig_hw, orig_per_image);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
