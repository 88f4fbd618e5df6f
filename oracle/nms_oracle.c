/* TEST INFRASTRUCTURE ONLY -- the checker, never the product path.
 *
 * CPU restatement of the greedy NMS the reference delegates to
 * torchvision.ops.boxes.batched_nms (reference call site tools.py:556-558; torchvision is a
 * third-party dependency, pinned torchvision==0.6.0 in requirements.txt:3, 0.26.0+cu128 installed).
 * torchvision's kernels are compiled-only, so this file restates their published/disassembled
 * arithmetic in both rounding orders:
 *
 *   round_mode 0  "tv_cpu"      : Sb = fl(bw*bh); D = fl(fl(Sa+Sb) - I); (double)fl(I/D) > thr
 *   round_mode 1  "tv_cuda_fma" : S  = fma(bw, bh, Sa); D = fl(S - I);   fl(I/D) > (float)thr
 *
 * with a = the higher-scored box of the pair (row box of the CUDA kernel) and b the lower-scored
 * one.  Boxes are visited in stable descending score order (lowest index first among ties).
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no implicit contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int pair_suppresses(const float *a, float Sa, const float *b, double thr, int round_mode)
{
    float l = a[0] > b[0] ? a[0] : b[0];
    float t = a[1] > b[1] ? a[1] : b[1];
    float r = a[2] < b[2] ? a[2] : b[2];
    float d = a[3] < b[3] ? a[3] : b[3];
    float w = r - l; w = w > 0.0f ? w : 0.0f;
    float h = d - t; h = h > 0.0f ? h : 0.0f;
    float I = w * h;
    float bw = b[2] - b[0], bh = b[3] - b[1];
    float D, q;
    if (round_mode == 1) {
        float S = fmaf(bw, bh, Sa);
        D = S - I;
        q = I / D;
        return q > (float)thr;
    }
    {
        float Sb = bw * bh;
        float S = Sa + Sb;
        D = S - I;
        q = I / D;
        return (double)q > thr;
    }
}

/* order[0..M): candidate indices in visiting order (stable score-descending).
 * keep_pos[0..ret): positions (into `order`) of the kept boxes, ascending.          */
int64_t pq_oracle_nms_ordered(const float *boxes, const int64_t *order, int64_t M,
                              double thr, int round_mode, int64_t *keep_pos)
{
    unsigned char *dead = (unsigned char *)calloc((size_t)(M > 0 ? M : 1), 1);
    int64_t nk = 0;
    for (int64_t i = 0; i < M; ++i) {
        if (dead[i]) continue;
        keep_pos[nk++] = i;
        const float *a = boxes + 4 * order[i];
        float Sa = (a[2] - a[0]) * (a[3] - a[1]);
        for (int64_t j = i + 1; j < M; ++j) {
            if (dead[j]) continue;
            if (pair_suppresses(a, Sa, boxes + 4 * order[j], thr, round_mode)) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

/* Same result through the CUDA kernel's formulation: full upper-triangular mask, then the
 * serial gather (used to cross-check that "greedy" == "mask + gather").                   */
int64_t pq_oracle_nms_mask(const float *boxes, const int64_t *order, int64_t M,
                           double thr, int round_mode, int64_t *keep_pos)
{
    int64_t words = (M + 63) / 64;
    uint64_t *mask = (uint64_t *)calloc((size_t)(M * words + 1), sizeof(uint64_t));
    uint64_t *removed = (uint64_t *)calloc((size_t)(words + 1), sizeof(uint64_t));
    for (int64_t i = 0; i < M; ++i) {
        const float *a = boxes + 4 * order[i];
        float Sa = (a[2] - a[0]) * (a[3] - a[1]);
        for (int64_t j = i + 1; j < M; ++j)
            if (pair_suppresses(a, Sa, boxes + 4 * order[j], thr, round_mode))
                mask[i * words + j / 64] |= (uint64_t)1 << (j % 64);
    }
    int64_t nk = 0;
    for (int64_t i = 0; i < M; ++i) {
        if (removed[i / 64] & ((uint64_t)1 << (i % 64))) continue;
        keep_pos[nk++] = i;
        for (int64_t w = 0; w < words; ++w) removed[w] |= mask[i * words + w];
    }
    free(mask); free(removed);
    return nk;
}

/* fp32 box + per-class offset exactly as _batched_nms_coordinate_trick does it
 * (torchvision/ops/boxes.py:97-103): off = fl(float(cls) * fl(max + 1)); box' = fl(box + off). */
void pq_oracle_trick_offsets(const float *boxes, const int64_t *cls, int64_t M, float *out)
{
    if (M <= 0) return;
    float m = boxes[0];
    for (int64_t i = 1; i < 4 * M; ++i) if (boxes[i] > m) m = boxes[i];
    float m1 = m + 1.0f;
    for (int64_t i = 0; i < M; ++i) {
        float off = (float)cls[i] * m1;
        for (int k = 0; k < 4; ++k) out[4 * i + k] = boxes[4 * i + k] + off;
    }
}

/* ---- IoU "is below threshold for every GT" mask of model/loss.py:85-90 (iou_calc3, tools.py:357-376)
 * one rounding per op, no contraction; NaN compares false, like torch.max(...) < thr with NaN propagation. */
void pq_oracle_ignore_mask(const float *pred, int64_t R, const float *gt, int64_t G,
                           float thr, unsigned char *below)
{
    for (int64_t r = 0; r < R; ++r) {
        const float *p = pred + 4 * r;
        float a1 = (p[2] - p[0]) * (p[3] - p[1]);
        float mx = 0.0f; int first = 1, isnan_ = 0;
        for (int64_t g = 0; g < G; ++g) {
            const float *q = gt + 4 * g;
            float a2 = (q[2] - q[0]) * (q[3] - q[1]);
            float lx = p[0] > q[0] ? p[0] : q[0], ly = p[1] > q[1] ? p[1] : q[1];
            float rx = p[2] < q[2] ? p[2] : q[2], ry = p[3] < q[3] ? p[3] : q[3];
            float w = rx - lx; w = w > 0.0f ? w : 0.0f;
            float h = ry - ly; h = h > 0.0f ? h : 0.0f;
            float I = w * h;
            float U = (a1 + a2) - I;
            float v = I / U;
            if (v != v) isnan_ = 1;
            if (first || v > mx) { mx = v; first = 0; }
        }
        below[r] = (unsigned char)(!isnan_ && G > 0 && mx < thr);
    }
}
