// TEST INFRASTRUCTURE ONLY.  Compiles the per-element cores of pqdet_b200/csrc/pq_math.cuh as plain
// host C++ (g++ -ffp-contract=off) so their arithmetic can be checked against the golden fixtures on a
// machine without a GPU.  Nothing in pqdet_b200/ loads this library: it is not a fallback, only a way
// to catch formula errors before spending GPU time.  The loops below restate the kernels' per-row flow.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../pqdet_b200/csrc/pq_math.cuh"

using namespace pq;

extern "C" {

void hh_decode(const float* raw, float* out, int B, int A, int C, int H, int W, float stride) {
  const int ch = 5 + C, HW = H * W;
  for (int b = 0; b < B; ++b)
    for (int cell = 0; cell < HW; ++cell)
      for (int a = 0; a < A; ++a)
        for (int k = 0; k < ch; ++k) {
          float v = raw[((size_t)(b * A + a) * ch + k) * HW + cell];
          int cy = cell / W, cx = cell % W;
          out[(((size_t)b * HW + cell) * A + a) * ch + k] = k < 4 ? decode_coord(k, v, cx, cy, stride) : sigmoidf_(v);
        }
}

void hh_recover(const float* pred, float* out, int B, int64_t N, int C, int kind, float in_h, float in_w,
                const float* orig, int per_image) {
  for (int b = 0; b < B; ++b) {
    const float* o = orig + (per_image ? 2 * b : 0);
    Affine a = affine_params(kind, in_h, in_w, o[0], o[1]);
    for (int64_t r = 0; r < N; ++r) {
      const float* p = pred + ((size_t)b * N + r) * (5 + C);
      float* q = out + ((size_t)b * N + r) * (4 + C);
      for (int k = 0; k < 4; ++k) q[k] = recover_coord(k, p[k], a);
      for (int c = 0; c < C; ++c) q[4 + c] = PQ_MUL(p[5 + c], p[4]);
    }
  }
}

int hh_nms_suppresses(const float* a, const float* b, double thr, int round) {
  float Sa = box_area(a[0], a[1], a[2], a[3]);
  if (round == 0) return nms_suppresses<0>(a[0], a[1], a[2], a[3], Sa, b[0], b[1], b[2], b[3], (float)thr, thr);
  return nms_suppresses<1>(a[0], a[1], a[2], a[3], Sa, b[0], b[1], b[2], b[3], (float)thr, thr);
}

void hh_iou(const float* b1, const float* b2, float* out, int64_t n, int kind) {
  for (int64_t i = 0; i < n; ++i) out[i] = iou_value(kind, b1 + 4 * i, b2 + 4 * i);
}

void hh_iou_grad(const float* b1, const float* b2, float* g1, float* g2, int64_t n, int kind) {
  for (int64_t i = 0; i < n; ++i) iou_value_grad(kind, b1 + 4 * i, b2 + 4 * i, g1 + 4 * i, g2 + 4 * i);
}

// decode + loss_per_scale forward/backward of one level, row by row like loss_fwd_bwd_kernel<true>.
void hh_loss(const float* raw, const float* label, const float* gt, float* grad, double* out4, int B, int A,
             int C, int H, int W, int G, float stride, int bbox_loss, float ignore_thresh, float l1_gain) {
  const int ch = 5 + C, LW = 6 + C, HW = H * W;
  const float in_area = (float)((double)(stride * H) * (double)(stride * W));
  const float inv_B = 1.0f / (float)B;
  double sums[3] = {0, 0, 0};
  for (int b = 0; b < B; ++b)
    for (int cell = 0; cell < HW; ++cell)
      for (int a = 0; a < A; ++a) {
        const size_t plane0 = ((size_t)b * A + a) * ch * HW;
        const float* lab = label + (((size_t)b * HW + cell) * A + a) * LW;
        float pb[4], es[4];
        int cy = cell / W, cx = cell % W;
        for (int k = 0; k < 4; ++k) {
          float v = raw[plane0 + (size_t)k * HW + cell];
          pb[k] = decode_coord(k, v, cx, cy, stride);
          float e = expf(v) * stride;
          es[k] = k < 2 ? -e : e;
        }
        float pconf = sigmoidf_(raw[plane0 + (size_t)4 * HW + cell]);
        float respond = lab[4], mixw = lab[5 + C];
        float dbox[4];
        float lb = bbox_loss_row(bbox_loss, pb, lab, respond, in_area, l1_gain, dbox);
        bool below = true;
        if (respond != 1.0f) {
          float a1 = box_area(pb[0], pb[1], pb[2], pb[3]);
          for (int g = 0; g < G && below; ++g) {
            const float* q = gt + ((size_t)b * G + g) * 4;
            below = iou_below(pb[0], pb[1], pb[2], pb[3], a1, q[0], q[1], q[2], q[3],
                              box_area(q[0], q[1], q[2], q[3]), ignore_thresh);
          }
        }
        float bgd = PQ_MUL(PQ_SUB(1.0f, respond), below ? 1.0f : 0.0f);
        float dconf;
        float lc = focal_bce_term(1.0f, 0.75f, respond, pconf, respond, bgd, &dconf);
        float gw = mixw * inv_B, lp = 0.f;
        for (int c = 0; c < C; ++c) {
          float g = 0.f;
          if (respond != 0.0f) {
            float p = sigmoidf_(raw[plane0 + (size_t)(5 + c) * HW + cell]), dp;
            lp = PQ_ADD(lp, focal_bce_term(2.0f, 0.5f, lab[5 + c], p, respond, -1.0f, &dp));
            g = dp * (p * (1.0f - p)) * gw;
          }
          grad[plane0 + (size_t)(5 + c) * HW + cell] = g;
        }
        for (int k = 0; k < 4; ++k) grad[plane0 + (size_t)k * HW + cell] = dbox[k] * es[k] * gw;
        grad[plane0 + (size_t)4 * HW + cell] = dconf * (pconf * (1.0f - pconf)) * gw;
        sums[0] += (double)PQ_MUL(lb, mixw);
        sums[1] += (double)PQ_MUL(lc, mixw);
        sums[2] += (double)PQ_MUL(lp, mixw);
      }
  float l0 = (float)(sums[0] / B), l1 = (float)(sums[1] / B), l2 = (float)(sums[2] / B);
  out4[0] = PQ_ADD(PQ_ADD(l0, l1), l2); out4[1] = l0; out4[2] = l1; out4[3] = l2;
}

// create_label for one image through assign_one (owner = last GT index wins), like assign_kernel.
void hh_assign(const float* gt, int n, int C, const float* anchors, const int* strides, const int* H,
               const int* W, double iou_thr, float* label0, float* label1, float* label2, float* lists,
               int list_cap, int* list_len) {
  float* label[3] = {label0, label1, label2};
  const int LW = 6 + C;
  const double deta = 0.01, uni = 1.0 / (double)C;
  const float hot = (float)(1.0 * (1 - deta) + deta * uni), cold = (float)(0.0 * (1 - deta) + deta * uni);
  for (int s = 0; s < 3; ++s) {
    for (int64_t e = 0; e < (int64_t)H[s] * W[s] * 3 * LW; ++e) label[s][e] = (e % LW == LW - 1) ? 1.0f : 0.0f;
    list_len[s] = 0;
  }
  for (int j = 0; j < n; ++j) {
    const float* g = gt + 6 * j;
    AssignHit hit = assign_one(g, anchors, strides, iou_thr);
    for (int i = 0; i < 9; ++i) {
      if (!((hit.mask >> i) & 1u)) continue;
      int s = i / 3, r = i % 3;
      float* d = label[s] + ((size_t)(hit.cy[s] * W[s] + hit.cx[s]) * 3 + r) * LW;
      d[0] = g[0]; d[1] = g[1]; d[2] = g[2]; d[3] = g[3]; d[4] = 1.0f;
      for (int c = 0; c < C; ++c) d[5 + c] = (c == (int)g[4]) ? hot : cold;
      d[5 + C] = g[5];
      if (list_len[s] < list_cap) memcpy(lists + ((size_t)s * list_cap + list_len[s]) * 4, g, 16);
      list_len[s]++;
    }
  }
}

}  // extern "C"
