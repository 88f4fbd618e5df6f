// Eval pre-processing (second half): the chain eval_augment_voc / eval_augment_coco
// build (dataset/voc_sample.py:85-90) -- augment.Resize (dataset/augment.py:227-259: cv2.resize INTER_LINEAR to the
// letterbox size, constant padding 128), augment.Normalize (:206-215) and augment.ToTensor (:390-398) -- for a
// batch of uint8 HWC images of different sizes in one launch: byte work, HBM bound.
//
// cv2.resize(INTER_LINEAR) on 8-bit images is fixed-point arithmetic (OpenCV resize.cpp, 11-bit coefficients):
//   x taps : fx = float((dx+0.5)*scale_x - 0.5); sx = floor(fx); fx -= sx; sx < 0 -> (0, fx=0);
//            sx >= sw-1 -> (sw-1, fx=0);  a0 = rint((1-fx)*2048), a1 = rint(fx*2048)            (float32)
//   y taps : same formula but NO reset at the borders; the two source rows sy, sy+1 are clipped to [0, sh-1]
//   H(y,x) = S[y][sx]*a0 + S[y][min(sx+1, sw-1)]*a1                                              (int)
//   dst    = (((b0 * (H(y0,x) >> 4)) >> 16) + ((b1 * (H(y1,x) >> 4)) >> 16) + 2) >> 2
// (oracle/augment_oracle.py restates this in numpy and is pinned bit-exactly against cv2 itself.)
#include "pq_common.cuh"

namespace pq {

struct LetterboxImage {     // filled by the host exactly as augment.Resize.__call__ computes them (:236-249)
  int64_t src_off;          // byte offset of the image in the packed source buffer
  int sh, sw;               // source height / width
  int dh, dw;               // resized height / width  (round(ratio * h), round(ratio * w))
  int du, dl;               // top / left padding
  double scale_y, scale_x;  // 1 / (dsize / ssize), as cv::resize derives them
};

__device__ __forceinline__ void lb_tap(int d, double scale, int sn, bool reset_at_border, int& s, int& c0, int& c1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (reset_at_border) {
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= sn - 1) { f = 0.0f; s = sn - 1; }
  }
  c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));    // saturate_cast<short>: round half to even
  c1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

// One thread per output pixel (3 channels).  out_f: (B,3,th,tw) normalised float32 (nullable);
// out_u8: (B,th,tw,3) padded resized image (nullable).
__global__ void __launch_bounds__(256)
letterbox_kernel(const uint8_t* __restrict__ src, const LetterboxImage* __restrict__ imgs, int th, int tw,
                 int pad_val, float m0, float m1, float m2, float s0, float s1, float s2,
                 float* __restrict__ out_f, uint8_t* __restrict__ out_u8) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= tw) return;
  const LetterboxImage I = imgs[b];
  int v[3] = {pad_val, pad_val, pad_val};
  const int rx = x - I.dl, ry = y - I.du;
  if (rx >= 0 && rx < I.dw && ry >= 0 && ry < I.dh) {
    int sx, a0, a1, sy, b0, b1;
    lb_tap(rx, I.scale_x, I.sw, true, sx, a0, a1);
    lb_tap(ry, I.scale_y, I.sh, false, sy, b0, b1);
    const int sx1 = min(sx + 1, I.sw - 1);
    const int y0 = min(max(sy, 0), I.sh - 1), y1 = min(max(sy + 1, 0), I.sh - 1);
    const uint8_t* r0 = src + I.src_off + (size_t)y0 * I.sw * 3;
    const uint8_t* r1 = src + I.src_off + (size_t)y1 * I.sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)r0[sx * 3 + c] * a0 + (int)r0[sx1 * 3 + c] * a1;
      const int h1 = (int)r1[sx * 3 + c] * a0 + (int)r1[sx1 * 3 + c] * a1;
      const int d = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      v[c] = min(max(d, 0), 255);
    }
  }
  if (out_u8) {
    uint8_t* o = out_u8 + (((size_t)b * th + y) * tw + x) * 3;
    o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
  }
  if (out_f) {
    // augment.Normalize: (img / 255. - mean) / std in float32, one rounding per operation; ToTensor: HWC -> CHW
    const size_t plane = (size_t)th * tw;
    float* o = out_f + (size_t)b * 3 * plane + (size_t)y * tw + x;
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[0], 255.0f), m0), s0);
    o[plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[1], 255.0f), m1), s1);
    o[2 * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[2], 255.0f), m2), s2);
  }
}

}  // namespace pq

extern "C" int pqdet_letterbox_normalize(const uint8_t* src, const void* images, int B, int target_h, int target_w,
                                         int pad_val, const float* mean3, const float* std3, float* out_chw,
                                         uint8_t* out_hwc_u8, int device, void* stream) {
  using namespace pq;
  if (B < 0 || target_h < 1 || target_w < 1 || pad_val < 0 || pad_val > 255) return PQDET_ERR_INVALID_ARG;
  if (B == 0) return PQDET_OK;
  if (!src || !images || (!out_chw && !out_hwc_u8) || (out_chw && (!mean3 || !std3))) return PQDET_ERR_INVALID_ARG;
  if (B > 65535 || target_h > 65535) return PQDET_ERR_UNSUPPORTED;
  PQ_ENTER(device);
  const float m[3] = {mean3 ? mean3[0] : 0.f, mean3 ? mean3[1] : 0.f, mean3 ? mean3[2] : 0.f};
  const float s[3] = {std3 ? std3[0] : 1.f, std3 ? std3[1] : 1.f, std3 ? std3[2] : 1.f};
  dim3 grid((target_w + 255) / 256, target_h, B);
  letterbox_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (const LetterboxImage*)images, target_h, target_w,
                                                            pad_val, m[0], m[1], m[2], s[0], s[1], s[2], out_chw,
                                                            out_hwc_u8);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
