// GPU target assignment: TrainDataset.create_label +
// collate_batch for a whole batch (dataset/train_dataset.py:16-43, 109-150).
//
// The dense label tensors are mostly background, so the cost is the streaming write of L bytes
// (assign_fill_kernel, 128-bit stores).  The per-GT work (9 anchor IoUs in numpy's mixed
// fp32/fp64 arithmetic) is tiny; "last writer wins" on cell collisions is reproduced with an
// owner map (atomicMax of the GT index per label slot) followed by a write pass in which only
// the owner stores its row; the per-scale GT lists keep numpy's order (GT order, then anchor
// order, one entry per hit) through a block prefix sum.
#include <string.h>

#include "pq_common.cuh"

namespace pq {

struct AssignParams {
  const float* gt;          // (B, n_max, 6)
  const int32_t* gt_count;  // (B)
  int B, n_max, C;
  float anchors[18];
  int strides[3], H[3], W[3];
  double iou_thr;
  float hot, cold;
  float* label[3];
  float* gtlist[3];
  int list_capacity;
  int32_t* list_len;        // (B,3)
  int32_t* owner[3];        // (B, H*W*3) each; planar (B, 3, H*W) when owner_planar (sparse targets)
  int owner_planar;
};

__device__ __forceinline__ size_t owner_slot(const AssignParams& P, int b, int s, int cell, int r) {
  const size_t HW = (size_t)P.H[s] * P.W[s];
  return P.owner_planar ? ((size_t)b * 3 + r) * HW + cell : (size_t)b * HW * 3 + (size_t)cell * 3 + r;
}

// background: zeros, mixw channel (last) = 1.0
__global__ void __launch_bounds__(256)
assign_fill_kernel(float* __restrict__ dst, int64_t total, int LW) {
  // the owner pass that follows has no dependency on this fill: it is launched as a programmatic dependent and runs
  // beside it (its last CTA waits for this grid, see assign_kernel)
  cudaTriggerProgrammaticLaunchCompletion();
  const int64_t n4 = total >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // position inside a label row, advanced by (4 * stride) mod LW per iteration: one division per thread instead of
  // four per store (the fill was issue bound on them before it was HBM bound)
  int r = (int)((i0 << 2) % LW);
  const int step = (int)((stride << 2) % LW);
  for (int64_t i = i0; i < n4; i += stride) {
    // element 4i + u is the mixw channel of its row iff r + u == LW - 1 (r < LW, u <= 3 and LW >= 7: no second wrap)
    const int t = LW - 1 - r;
    float4 v;
    v.x = (t == 0) ? 1.0f : 0.0f;
    v.y = (t == 1) ? 1.0f : 0.0f;
    v.z = (t == 2) ? 1.0f : 0.0f;
    v.w = (t == 3) ? 1.0f : 0.0f;
    reinterpret_cast<float4*>(dst)[i] = v;
    r += step;
    if (r >= LW) r -= LW;
  }
  for (int64_t e = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride)
    dst[e] = ((int)(e % LW) == LW - 1) ? 1.0f : 0.0f;
}

// One CTA per image.  PASS 0: owner map + GT lists.  PASS 1: owners write their label rows.
template <int PASS>
__global__ void __launch_bounds__(256)
assign_kernel(const __grid_constant__ AssignParams P, int behind_fill) {
  __shared__ int s_warp[3][8];
  __shared__ int s_base[3];
  const int b = blockIdx.x;
  const int n = min(P.gt_count[b], P.n_max);
  const int LW = 6 + P.C;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  if (PASS == 0 && tid < 3) s_base[tid] = 0;
  __syncthreads();
  for (int j0 = 0; j0 < n; j0 += 256) {
    const int j = j0 + tid;
    const bool valid = j < n;
    AssignHit hit;
    hit.mask = 0;
    const float* g = P.gt + ((size_t)b * P.n_max + (valid ? j : 0)) * 6;
    if (valid) hit = assign_one(g, P.anchors, P.strides, P.iou_thr);
    bool inb[3];
#pragma unroll
    for (int s = 0; s < 3; ++s)
      inb[s] = valid && hit.cx[s] >= 0 && hit.cx[s] < P.W[s] && hit.cy[s] >= 0 && hit.cy[s] < P.H[s];
    if (PASS == 0) {
      int cnt[3], ex[3];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        cnt[s] = __popc((hit.mask >> (3 * s)) & 7u);
        const int inc = warp_inclusive_sum(cnt[s]);
        ex[s] = inc - cnt[s];
        if (lane == 31) s_warp[s][warp] = inc;
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int pos = s_base[s] + ex[s];
        for (int w = 0; w < warp; ++w) pos += s_warp[s][w];
        for (int r = 0; r < 3; ++r) {
          if (!((hit.mask >> (3 * s + r)) & 1u)) continue;
          if (inb[s]) atomicMax(&P.owner[s][owner_slot(P, b, s, hit.cy[s] * P.W[s] + hit.cx[s], r)], j);
          if (pos < P.list_capacity) {
            float* d = P.gtlist[s] + ((size_t)b * P.list_capacity + pos) * 4;
            d[0] = g[0]; d[1] = g[1]; d[2] = g[2]; d[3] = g[3];
          }
          ++pos;
        }
      }
      __syncthreads();
      if (tid < 3) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += s_warp[tid][w];
        s_base[tid] += tot;
      }
      __syncthreads();
    } else {
      if (!valid) continue;
      const int cls = (int)g[4];
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (!inb[s]) continue;
        for (int r = 0; r < 3; ++r) {
          if (!((hit.mask >> (3 * s + r)) & 1u)) continue;
          const size_t slot = (size_t)b * P.H[s] * P.W[s] * 3 + (hit.cy[s] * P.W[s] + hit.cx[s]) * 3 + r;
          if (P.owner[s][slot] != j) continue;
          float* d = P.label[s] + slot * LW;
          d[0] = g[0]; d[1] = g[1]; d[2] = g[2]; d[3] = g[3];
          d[4] = 1.0f;
          for (int c = 0; c < P.C; ++c) d[5 + c] = (c == cls) ? P.hot : P.cold;
          d[5 + P.C] = g[5];
        }
      }
    }
  }
  if (PASS == 0) {
    __syncthreads();
    if (tid < 3) P.list_len[b * 3 + tid] = s_base[tid];
    // launched as a programmatic dependent of assign_fill_kernel: one CTA stays until the fill has completed, so that
    // the write pass behind this grid (a normal launch: it waits for THIS grid) also finds the background in place
    if (behind_fill && b == 0) cudaGridDependencySynchronize();
  }
}

}  // namespace pq

extern "C" int64_t pqdet_assign_workspace(int B, const int* H, const int* W) {
  if (B < 0 || !H || !W) return PQDET_ERR_INVALID_ARG;
  int64_t slots = 0;
  for (int s = 0; s < 3; ++s) slots += (int64_t)H[s] * W[s] * 3;
  return (int64_t)B * slots * 4 + 256;
}

extern "C" int pqdet_assign_labels(const float* gt, const int32_t* gt_count, int B, int n_max, int C,
                                   const float* anchors, const int* strides, const int* H, const int* W,
                                   float iou_threshold, float* label0, float* label1, float* label2,
                                   float* gtlist0, float* gtlist1, float* gtlist2, int list_capacity,
                                   int32_t* list_len, void* owner, int device, void* stream) {
  using namespace pq;
  if (!gt_count || !anchors || !strides || !H || !W || !label0 || !label1 || !label2 || !gtlist0 ||
      !gtlist1 || !gtlist2 || !list_len || !owner)
    return PQDET_ERR_INVALID_ARG;
  if (B < 1 || n_max < 0 || C < 1 || list_capacity < 1 || (n_max > 0 && !gt)) return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  AssignParams P;
  P.gt = gt; P.gt_count = gt_count; P.B = B; P.n_max = n_max; P.C = C;
  // anchors are a host array: (w,h) x 9, config.py:58-59
  for (int i = 0; i < 18; ++i) P.anchors[i] = anchors[i];
  float* labels[3] = {label0, label1, label2};
  float* lists[3] = {gtlist0, gtlist1, gtlist2};
  int32_t* own = (int32_t*)owner;
  size_t own_total = 0;
  for (int s = 0; s < 3; ++s) {
    if (H[s] < 1 || W[s] < 1 || strides[s] < 1) return PQDET_ERR_INVALID_ARG;
    P.strides[s] = strides[s]; P.H[s] = H[s]; P.W[s] = W[s];
    P.label[s] = labels[s]; P.gtlist[s] = lists[s];
    P.owner[s] = own + own_total;
    P.owner_planar = 0;
    own_total += (size_t)B * H[s] * W[s] * 3;
  }
  P.iou_thr = (double)iou_threshold;
  // label smoothing in fp64 then stored fp32 (train_dataset.py:126-130)
  const double deta = 0.01, uni = 1.0 / (double)C;
  P.hot = (float)(1.0 * (1 - deta) + deta * uni);
  P.cold = (float)(0.0 * (1 - deta) + deta * uni);
  P.list_capacity = list_capacity;
  P.list_len = list_len;
  PQ_CUDA(cudaMemsetAsync(owner, 0xff, own_total * sizeof(int32_t), st));
  // Background fill.  When the caller carved the three label tensors (and the three GT lists) out of one
  // allocation, back to back, everything is one streaming launch / one memset: rows have the same
  // 6+C layout on every level.
  int64_t ltot[3];
  for (int s = 0; s < 3; ++s) ltot[s] = (int64_t)B * H[s] * W[s] * 3 * (6 + C);
  const size_t list_floats = (size_t)B * list_capacity * 4;
  const bool one_label_buf = (labels[1] == labels[0] + ltot[0]) && (labels[2] == labels[1] + ltot[1]);
  const bool one_list_buf = (lists[1] == lists[0] + list_floats) && (lists[2] == lists[1] + list_floats);
  if (one_list_buf) {
    PQ_CUDA(cudaMemsetAsync(lists[0], 0, 3 * list_floats * sizeof(float), st));
  } else {
    for (int s = 0; s < 3; ++s) PQ_CUDA(cudaMemsetAsync(lists[s], 0, list_floats * sizeof(float), st));
  }
  auto fill = [&](float* dst, int64_t total) -> int {
    int64_t blocks = ((total >> 2) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;           // grid-stride, whole waves on 148 SMs
    if (blocks < 1) blocks = 1;
    assign_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(dst, total, 6 + C);
    PQ_LAUNCH_CHECK();
    return PQDET_OK;
  };
  if (one_label_buf) {
    if (fill(labels[0], ltot[0] + ltot[1] + ltot[2]) != PQDET_OK) return PQDET_ERR_CUDA;
  } else {
    for (int s = 0; s < 3; ++s)
      if (fill(labels[s], ltot[s]) != PQDET_OK) return PQDET_ERR_CUDA;
  }
  if (one_label_buf && !getenv("PQDET_ASSIGN_NO_PDL")) {
    // (one fill launch directly in front: the owner pass may start while it streams)
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(B);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PQ_CUDA(cudaLaunchKernelEx(&cfg, assign_kernel<0>, P, 1));
  } else {
    assign_kernel<0><<<B, 256, 0, st>>>(P, 0);
  }
  PQ_LAUNCH_CHECK();
  assign_kernel<1><<<B, 256, 0, st>>>(P, 0);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

// Sparse targets: the same assignment, but instead of the dense (B,H,W,3,6+C)
// label tensors only an owner map is produced -- owner[s] (B, 3, H_s*W_s) int32 = index of the GT that
// train_dataset.py:145 would have written last into that label slot, or -1 -- plus the per-scale GT lists.  The
// loss kernel rebuilds the label row of a responsible cell from the GT row itself (pqdet_loss_levels_sparse).
extern "C" int pqdet_assign_sparse(const float* gt, const int32_t* gt_count, int B, int n_max,
                                   const float* anchors, const int* strides, const int* H, const int* W,
                                   float iou_threshold, int32_t* owner0, int32_t* owner1, int32_t* owner2,
                                   float* gtlist0, float* gtlist1, float* gtlist2, int list_capacity,
                                   int32_t* list_len, int device, void* stream) {
  using namespace pq;
  if (!gt_count || !anchors || !strides || !H || !W || !owner0 || !owner1 || !owner2 || !gtlist0 || !gtlist1 ||
      !gtlist2 || !list_len)
    return PQDET_ERR_INVALID_ARG;
  if (B < 1 || n_max < 0 || list_capacity < 1 || (n_max > 0 && !gt)) return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  AssignParams P;
  memset(&P, 0, sizeof(P));
  P.gt = gt; P.gt_count = gt_count; P.B = B; P.n_max = n_max; P.C = 1;
  for (int i = 0; i < 18; ++i) P.anchors[i] = anchors[i];
  int32_t* owners[3] = {owner0, owner1, owner2};
  float* lists[3] = {gtlist0, gtlist1, gtlist2};
  size_t slots[3];
  for (int s = 0; s < 3; ++s) {
    if (H[s] < 1 || W[s] < 1 || strides[s] < 1) return PQDET_ERR_INVALID_ARG;
    P.strides[s] = strides[s]; P.H[s] = H[s]; P.W[s] = W[s];
    P.gtlist[s] = lists[s];
    P.owner[s] = owners[s];
    slots[s] = (size_t)B * H[s] * W[s] * 3;
  }
  P.owner_planar = 1;
  P.iou_thr = (double)iou_threshold;
  P.list_capacity = list_capacity;
  P.list_len = list_len;
  // one memset each when the caller carved the three maps / lists out of one allocation, back to back
  if (owners[1] == owners[0] + slots[0] && owners[2] == owners[1] + slots[1]) {
    PQ_CUDA(cudaMemsetAsync(owners[0], 0xff, (slots[0] + slots[1] + slots[2]) * sizeof(int32_t), st));
  } else {
    for (int s = 0; s < 3; ++s) PQ_CUDA(cudaMemsetAsync(owners[s], 0xff, slots[s] * sizeof(int32_t), st));
  }
  const size_t list_floats = (size_t)B * list_capacity * 4;
  if (lists[1] == lists[0] + list_floats && lists[2] == lists[1] + list_floats) {
    PQ_CUDA(cudaMemsetAsync(lists[0], 0, 3 * list_floats * sizeof(float), st));
  } else {
    for (int s = 0; s < 3; ++s) PQ_CUDA(cudaMemsetAsync(lists[s], 0, list_floats * sizeof(float), st));
  }
  assign_kernel<0><<<B, 256, 0, st>>>(P, 0);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
