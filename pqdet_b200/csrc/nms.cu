// Fused decode + recover + score threshold + class-aware NMS.
//
// decode_nms_fused_kernel: one CTA per image (persistent, dynamic image scheduler), everything
// after the HBM read lives in shared memory:
//   1. scan the objectness planes only (A of the A*(5+C) channels): a row can produce a candidate
//      only if conf > thr, because score = fl(prob*conf) <= conf for prob <= 1.  A conservative
//      logit-space prefilter skips the sigmoid for rows that cannot pass.  Hits are recorded as one
//      ballot word per (level, 32-cell chunk, anchor).
//   2. popcount prefix over the words gives every hit row a deterministic slot (row-major order),
//      so ties in score are later broken by (row, class) exactly like nonzero() + stable sort.
//   3. only for hit rows, fetch the 4 box + C class channels, decode + recover the box, form the
//      scores, and push (class | ~score | hit) 64-bit keys into a shared-memory candidate list.
//   4. block bitonic sort by (class asc, score desc, hit asc); per-class greedy NMS, one warp per
//      class, 32 candidates per step (ballot-driven: work is proportional to the kept boxes);
//      kept keys are re-sorted by (score desc, row asc, class asc) and written out.
// The IoU arithmetic is torchvision's, in either of its two rounding orders, on the coordinate-
// trick-shifted boxes or on the plain boxes depending on the candidate count, so the keep list is
// bit-identical to tools.torch_nms on the same decoded boxes.
//
// Images whose hit rows / candidates do not fit the on-chip lists are flagged
// (PQDET_ST_CAND_OVERFLOW) and handled by the general path below: global-memory candidate
// lists bucketed by (image, class), one CTA per bucket.
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "pq_common.cuh"

namespace pq {

constexpr int kHitBits = 25;  // low key field: hit slot (fused) or row (general)
constexpr uint64_t kHitMask = (1ull << kHitBits) - 1;

struct LevelDev {
  const float* raw;
  int H, W, HW;
  float stride;
  int row_off;    // first row of this level in the concatenated prediction
  int group_off;  // first group (128 cells) of this level
  int nchunk;     // groups in this level = ceil(HW/128)
  int vec4;       // objectness planes are 16-byte aligned (HW % 4 == 0 and an aligned base): 128-bit loads
};

struct HeadsDev {
  LevelDev lv[PQDET_MAX_LEVELS];
  int n_levels;
  int B, A, C, ch;
  int N;       // rows per image
  int G_tot;   // 128-cell groups per image; ballot words per image = G_tot * 4 * A
  int kind;
  float in_h, in_w;
  const float* orig;
  int orig_per_image;
  float thr_f;     // score threshold as fp32
  float logit_lo;  // rows with objectness logit <= logit_lo cannot reach conf > thr
  float iou_f;
  double iou_d;
  int nms_mode;
  // SRC == 1 (scores source, the tools.torch_nms drop-in): recovered tensor (B, N, bb_row = 4+C)
  const float* bboxes;
  int bb_row;
  uint32_t magic_n4;   // ceil(2^32 / (bb_row / 4)): word index -> row by multiply-high
  // SRC == 2 (hit records written by the head convolution's epilogue, pqdet_head_conv_hits): per image up to
  // rec_cap records [row (int bits), objectness logit, 4 box, C class raw values], in arrival order
  const float* rec;
  const int32_t* rec_count;
  int rec_cap;
};

struct DetOut {
  float* det;
  int32_t* det_idx;
  int max_det;
  int32_t* counts;
  int32_t* ncand;
  int32_t* status;
  // gather mode (pqdet_decode_nms_gather): every kept row and every count is ALSO stored into the gathered buffers
  // of all ranks of the node - peer memory mapped over NVLink - at this rank's block, so the eval gather costs no
  // collective: peer_det[p] (world, B, gather_cap, 6), peer_cnt[p] (world, B), img_off = rank * B
  float* peer_det[8];
  int32_t* peer_cnt[8];
  int n_peers, gather_cap, img_off;
  // optional: peer_arr[p] = the arrival counter of THIS rank in rank p's buffer: +1 per image whose rows and count have
  // been stored there (release at system scope); a rank's gathered block is complete when the counter of its source
  // has advanced by B (pqdet_peer_wait) - no barrier kernel between consecutive launches
  uint32_t* peer_arr[8];
  int32_t* img_done;     // with peer_arr: images of this launch whose outputs are stored (local word, zero between launches)
  int n_images;          // with peer_arr: images of this launch (the last one to finish sends the signals)
};

// the image's kept count, locally and - in gather mode - in every rank's gathered counts
__device__ __forceinline__ void set_count(const DetOut& O, int b, int k) {
  O.counts[b] = k;
  for (int p = 0; p < O.n_peers; ++p) O.peer_cnt[p][O.img_off + b] = min(k, O.gather_cap);
}

// gather mode with arrival counters: called by ONE thread after the image's peer stores (its own, and - ordered by the
// __threadfence_system + __syncthreads in front of it - those of the CTA's other threads)
// One signal per launch and peer, not per image: 8 ranks x 1024 images on one counter serialise at the receiver
// (measured: 155 against 112 us per step at 8 GPUs).  Every image is counted in a LOCAL word behind a system-scope
// fence; the CTA that counts the last one adds the launch's image count to this rank's counter on every peer
// (fences are cumulative: the other CTAs' peer stores are ordered before that release).
__device__ __forceinline__ void signal_peers(const DetOut& O) {
  if (O.n_peers == 0 || O.peer_arr[0] == nullptr) return;
  __threadfence_system();
  if (atomicAdd(O.img_done, 1) != O.n_images - 1) return;
  *O.img_done = 0;                                          // the next launch counts only after this grid has completed
  __threadfence_system();
  for (int p = 0; p < O.n_peers; ++p)
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(O.peer_arr[p]), "r"((uint32_t)O.n_images) : "memory");
}

__device__ __forceinline__ bool use_trick(int mode, int64_t M) {
  if (mode == PQDET_NMS_TRICK) return true;
  if (mode == PQDET_NMS_VANILLA) return false;
  const int64_t limit = (mode == PQDET_NMS_AUTO_CPU) ? 4000 : 100000;   // torchvision/ops/boxes.py:80
  return 4 * M <= limit;
}

__device__ __forceinline__ int level_of_group(const HeadsDev& P, int g) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < PQDET_MAX_LEVELS; ++i)
    if (i < P.n_levels && g >= P.lv[i].group_off) l = i;
  return l;
}
__device__ __forceinline__ int level_of_row(const HeadsDev& P, int row) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < PQDET_MAX_LEVELS; ++i)
    if (i < P.n_levels && row >= P.lv[i].row_off) l = i;
  return l;
}

__device__ __forceinline__ Affine image_affine(const HeadsDev& P, int b) {
  const float* o = P.orig + (P.orig_per_image ? 2 * b : 0);
  return affine_params(P.kind, P.in_h, P.in_w, o[0], o[1]);
}

// key of a candidate for the per-class phase: class asc, score desc, hit asc
__device__ __forceinline__ uint64_t cand_key(int c, float score, uint32_t hit) {
  return ((uint64_t)c << 57) | ((uint64_t)(~__float_as_uint(score)) << kHitBits) | hit;
}
// key of a kept detection for the output phase: score desc, (row|hit) asc, class asc
__device__ __forceinline__ uint64_t out_key(uint64_t ck) {
  uint64_t c = ck >> 57, nsc = (ck >> kHitBits) & 0xffffffffull, h = ck & kHitMask;
  return (nsc << 32) | (h << 7) | c;
}

__device__ __forceinline__ void write_det(const DetOut& O, int b, int j, float4 bx, float score, int c,
                                          int64_t row, int C) {
  float2* d = reinterpret_cast<float2*>(O.det + ((size_t)b * O.max_det + j) * 6);
  d[0] = make_float2(bx.x, bx.y);
  d[1] = make_float2(bx.z, bx.w);
  d[2] = make_float2(score, (float)c);
  if (O.det_idx) O.det_idx[(size_t)b * O.max_det + j] = (int32_t)(row * C + c);
}

// ------------------------------------------------------------------------------------------------
// fused kernel
// ------------------------------------------------------------------------------------------------
constexpr int kWarpClassMax = 128;  // classes up to this many candidates: one warp, members in registers (1, 2 or 4 per lane)

// Two capacity classes (include/pqdet_b200.h: capacity_class).  The compact one keeps the per-image lists small
// enough for 7 CTAs of 4 warps per SM (one CTA per image, 1036 images in flight on 148 SMs): the phases of an
// image are serial and separated by CTA barriers, so what hides one image's barrier and memory latency is the
// other images resident on the SM.  The large one is the round-1 geometry for denser scenes.
template <int CLS> struct FusedCfg;
template <> struct FusedCfg<0> { static constexpr int kThreads = 128, kCapH = 512, kCapM = 1280, kMinCtas = 7; };
template <> struct FusedCfg<1> { static constexpr int kThreads = 256, kCapH = 1024, kCapM = 2048, kMinCtas = 4; };

template <int CAPH, int CAPM>
struct FusedSmemT {
  uint64_t keys[CAPM];      // hit records during the scan, then candidate keys in emission order
  float4 hbox[CAPH];        // recovered box of every hit row
  uint32_t hmeta[CAPH];     // level << 30 | anchor << 27 | cell
  float hconf[CAPH];        // | hconf and hhas are dead once the NMS starts; together with the pad they hold the
  uint8_t hhas[CAPH];       // | output keys of the kept detections, appended by the selection rounds themselves
  uint8_t okpad[CAPM];      // | (kOutCap of them: an image that keeps more is resolved by the general path)
  uint16_t order[CAPM];     // per-class member lists (candidate slots)
  int cls_cnt[128];
  int cls_fill[128];
  int seg_start[128];
  const float* lvbase[PQDET_MAX_LEVELS];      // objectness plane of anchor 0 of this image, per level
  float red[8];
  uint32_t wbest[8];        // CTA-cooperative class walk: per-warp best ~score, winner tag, winner box
  uint32_t win;
  float4 wbox;
  int big[128];             // classes that need the whole CTA
  int nbig;
  int b, H, M, K, next_class, nrec;
  // followed by: uint32_t hitw[G_tot*4*A]; uint32_t gbase[G_tot]; uint32_t utab[G_tot*A];
  static constexpr int kOutCap = (5 * CAPH + CAPM) / 8;
  __device__ __forceinline__ uint64_t* okeys() { return reinterpret_cast<uint64_t*>(hconf); }
};

__device__ __forceinline__ uint32_t pack_meta(int level, int a, int cell) {
  return ((uint32_t)level << 30) | ((uint32_t)a << 27) | (uint32_t)cell;
}

// Greedy NMS of one class by ONE warp WITHOUT sorting it: the class' n <= 32*R members live in registers (R per
// lane: ~score bits, hit index, trick-shifted box).  Every round picks the best member still alive - one
// redux.sync min over the ~score bits; if several lanes hold that minimum (an exact score tie) a second one over
// the hit indices, which is the (score desc, hit asc) order of nonzero() + stable sort - keeps it, and tests the
// alive members against it.
// Identical to the sorted greedy walk: that one also visits candidates in this order and skips the suppressed.
// Work is proportional to kept x ceil(n/32); a class of 30 candidates keeping 6 costs ~300 warp instructions where
// rank sort + walk cost ~500.
template <int ROUND, int R, typename S_t>
__device__ __forceinline__ void warp_select_nms(S_t& S, int s, int n, float off, float iou_f, double iou_d) {
  const int lane = lane_id();
  uint32_t ns[R], tag[R];            // ~score bits; hit << 16 | candidate slot
  float x1[R], y1[R], x2[R], y2[R];
  unsigned alive = 0, kept = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + 32 * r;
    ns[r] = 0xffffffffu; tag[r] = 0xffffffffu;
    x1[r] = y1[r] = x2[r] = y2[r] = 0.0f;
    if (j < n) {
      const uint32_t slot = S.order[s + j];
      const uint64_t key = S.keys[slot];
      const uint32_t hit = (uint32_t)(key & kHitMask);
      ns[r] = (uint32_t)(key >> kHitBits);
      tag[r] = (hit << 16) | slot;
      const float4 bx = S.hbox[hit];
      x1[r] = PQ_ADD(bx.x, off); y1[r] = PQ_ADD(bx.y, off); x2[r] = PQ_ADD(bx.z, off); y2[r] = PQ_ADD(bx.w, off);
      alive |= 1u << r;
    }
  }
  for (;;) {
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if ((alive >> r) & 1u) m = min(m, ns[r]);
    const uint32_t best = __reduce_min_sync(PQ_FULL, m);
    if (best == 0xffffffffu) break;                        // nobody left (a live score is > thr >= 0: ~bits < 2^32 - 1)
    // the alive holders of `best`: normally one lane with one register; on an exact score tie the smallest hit
    // index wins (unique within a class) - (score desc, hit asc) is the order of nonzero() + stable sort
    uint32_t cand = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (((alive >> r) & 1u) && ns[r] == best) cand = min(cand, (tag[r] & 0xffff0000u) | (uint32_t)r);
    const unsigned holders = __ballot_sync(PQ_FULL, cand != 0xffffffffu);
    int wl = __ffs(holders) - 1;
    uint32_t win = 0u;
    if (holders & (holders - 1u)) {                        // several lanes hold it: rare
      win = __reduce_min_sync(PQ_FULL, cand);
      wl = __ffs(__ballot_sync(PQ_FULL, cand == win)) - 1;
    } else if (R > 1) {
      win = __shfl_sync(PQ_FULL, cand, wl);
    }
    const int wr = (R > 1) ? (int)(win & 0xffffu) : 0;
    float ax1 = x1[0], ay1 = y1[0], ax2 = x2[0], ay2 = y2[0];
    uint32_t wtag = tag[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
      if (wr == r) { ax1 = x1[r]; ay1 = y1[r]; ax2 = x2[r]; ay2 = y2[r]; wtag = tag[r]; }
    if (lane == wl) {
      kept |= 1u << wr;
      alive &= ~(1u << wr);
    }
    ax1 = __shfl_sync(PQ_FULL, ax1, wl); ay1 = __shfl_sync(PQ_FULL, ay1, wl);
    ax2 = __shfl_sync(PQ_FULL, ax2, wl); ay2 = __shfl_sync(PQ_FULL, ay2, wl);
    const float Sa = box_area(ax1, ay1, ax2, ay2);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (((alive >> r) & 1u) &&
          nms_suppresses<ROUND>(ax1, ay1, ax2, ay2, Sa, x1[r], y1[r], x2[r], y2[r], iou_f, iou_d))
        alive &= ~(1u << r);
  }
  // the class' kept detections append their output keys with ONE reservation (nothing in the rounds waits for an atomic)
  int total = 0;
  unsigned km[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    km[r] = __ballot_sync(PQ_FULL, (kept >> r) & 1u);
    total += __popc(km[r]);
  }
  int base = 0;
  if (lane == 0) base = atomicAdd(&S.K, total);
  base = __shfl_sync(PQ_FULL, base, 0);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pos = base + __popc(km[r] & ((1u << lane) - 1u));
    if (((kept >> r) & 1u) && pos < S_t::kOutCap) S.okeys()[pos] = out_key(S.keys[tag[r] & 0xffffu]);
    base += __popc(km[r]);
  }
}

// The same selection walk for a class too large for one warp, by the whole CTA: thread t owns members t, t + NT, ...
// (alive / kept bits in registers; the members themselves - ~score, hit, shifted box - are re-read from shared memory
// every round, which only crowds of one class pay for).  A round = per-warp redux of the best alive ~score, the warp
// results through shared memory, atomicMin over (hit, slot) among the holders of the best score (lowest hit index
// wins a tie), the winner publishes its box, everybody tests its alive members.  Three barriers per round.
// All threads of the CTA call it; n <= 32 * NT.
template <int ROUND, int NT, typename S_t>
__device__ __forceinline__ void cta_select_nms(S_t& S, int s, int n, float off, float iou_f, double iou_d) {
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int T = (n - tid + NT - 1) / NT;                     // members of this thread (n > tid) or <= 0
  unsigned alive = (T >= 32) ? 0xffffffffu : ((T > 0) ? ((1u << T) - 1u) : 0u), kept = 0;
  auto member = [&](int t, uint32_t& nsv, uint32_t& tagv) {
    const uint32_t slot = S.order[s + tid + NT * t];
    const uint64_t key = S.keys[slot];
    nsv = (uint32_t)(key >> kHitBits);
    tagv = ((uint32_t)(key & kHitMask) << 16) | slot;
  };
  for (;;) {
    uint32_t m = 0xffffffffu;
    for (unsigned rem = alive; rem; rem &= rem - 1) {
      uint32_t nsv, tagv;
      member(__ffs(rem) - 1, nsv, tagv);
      m = min(m, nsv);
    }
    const uint32_t wb = __reduce_min_sync(PQ_FULL, m);
    if (lane == 0) S.wbest[warp] = wb;
    if (tid == 0) S.win = 0xffffffffu;
    __syncthreads();
    uint32_t best = S.wbest[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) best = min(best, S.wbest[w]);
    if (best == 0xffffffffu) break;                          // uniform: every thread read the same words
    if (m == best)
      for (unsigned rem = alive; rem; rem &= rem - 1) {
        uint32_t nsv, tagv;
        member(__ffs(rem) - 1, nsv, tagv);
        if (nsv == best) atomicMin(&S.win, tagv);
      }
    __syncthreads();
    const uint32_t win = S.win;
    if (m == best)
      for (unsigned rem = alive; rem; rem &= rem - 1) {
        const int t = __ffs(rem) - 1;
        uint32_t nsv, tagv;
        member(t, nsv, tagv);
        if (tagv == win) {
          const float4 bx = S.hbox[win >> 16];
          S.wbox = make_float4(PQ_ADD(bx.x, off), PQ_ADD(bx.y, off), PQ_ADD(bx.z, off), PQ_ADD(bx.w, off));
          kept |= 1u << t;
          alive &= ~(1u << t);
        }
      }
    __syncthreads();
    const float4 a = S.wbox;
    const float Sa = box_area(a.x, a.y, a.z, a.w);
    for (unsigned rem = alive; rem; rem &= rem - 1) {
      const int t = __ffs(rem) - 1;
      uint32_t nsv, tagv;
      member(t, nsv, tagv);
      const float4 bx = S.hbox[tagv >> 16];
      if (nms_suppresses<ROUND>(a.x, a.y, a.z, a.w, Sa, PQ_ADD(bx.x, off), PQ_ADD(bx.y, off), PQ_ADD(bx.z, off),
                                PQ_ADD(bx.w, off), iou_f, iou_d))
        alive &= ~(1u << t);
    }
  }
  // append: one reservation per warp
  const int mine = __popc(kept);
  const int incl = warp_inclusive_sum(mine);
  const int total = __shfl_sync(PQ_FULL, incl, 31);
  int base = 0;
  if (lane == 0 && total) base = atomicAdd(&S.K, total);
  base = __shfl_sync(PQ_FULL, base, 0) + incl - mine;
  for (unsigned rem = kept; rem; rem &= rem - 1) {
    uint32_t nsv, tagv;
    member(__ffs(rem) - 1, nsv, tagv);
    if (base < S_t::kOutCap) S.okeys()[base] = out_key(S.keys[tagv & 0xffffu]);
    ++base;
  }
  __syncthreads();                                           // wbest / win are reused by the next class
}

#ifdef PQ_PHASE_TIMING
// debug builds only (profiles/tools/fused_phases.py): cycles per phase, summed over the CTAs' thread 0
__device__ unsigned long long pq_phase_cycles[16];
#define PQ_PHASE(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&pq_phase_cycles[i], (unsigned long long)(t_ - pq_t0)); pq_t0 = t_; } } while (0)
#else
#define PQ_PHASE(i) do { } while (0)
#endif

// Scan table: one word per unit (128 cells of one anchor's objectness plane), in hitw order un = group * A + anchor:
//   level << 30 | slow << 29 | element offset of the unit from the image's anchor-0 objectness plane of that level.
// `slow` marks units that need the generic code: a partial last group or planes that are not 16-byte aligned.
// The table depends only on the geometry, so a CTA builds it once and reuses it for every image it pulls.
constexpr uint32_t kUnitSlow = 1u << 29;
constexpr uint32_t kUnitOffMask = kUnitSlow - 1u;

__device__ __forceinline__ void build_unit_table(const HeadsDev& P, uint32_t* utab, int tid, int nthreads) {
  const int A = P.A, nu = P.G_tot * A;
  for (int un = tid; un < nu; un += nthreads) {
    const int g = un / A, a = un - g * A;
    const int l = level_of_group(P, g);
    const LevelDev& L = P.lv[l];
    const int gl = g - L.group_off;
    const uint32_t off = (uint32_t)a * (uint32_t)(P.ch * L.HW) + (uint32_t)gl * 128u;
    const bool slow = !L.vec4 || (gl + 1) * 128 > L.HW || off > kUnitOffMask;
    utab[un] = ((uint32_t)l << 30) | (slow ? kUnitSlow : 0u) | (off & kUnitOffMask);
  }
}

// The generic unit load of the scan (partial last group of a level, planes that are not 16-byte aligned such as
// 19x19): kept out of line, the unrolled scan loop only carries the one-instruction fast path.
__device__ __noinline__ float4 scan_load_slow(const HeadsDev& P, const float* lvbase_l, int l, int un, int lane) {
  float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const LevelDev& L = P.lv[l];
  const int g = un / P.A;
  const int cell = (g - L.group_off) * 128 + 4 * lane;
  const float* q = lvbase_l + ((size_t)(un - g * P.A) * (size_t)(P.ch * L.HW) + (size_t)cell);
  if (cell < L.HW) {
    if (L.vec4) {
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(q));
    } else {
      x.x = ldg_stream(q);
      if (cell + 1 < L.HW) x.y = ldg_stream(q + 1);
      if (cell + 2 < L.HW) x.z = ldg_stream(q + 2);
      if (cell + 3 < L.HW) x.w = ldg_stream(q + 3);
    }
  }
  return x;
}

// Objectness scan of one image.  The cells of a level are cut into groups of 128; a unit = (group g, anchor a):
// lane l reads cells g*128 + 4l .. 4l+3 of that anchor's objectness plane with one 128-bit load (scalar loads on
// the slow units) and the warp emits four ballot words, word k holding the cells 4l+k.  hitw layout:
// [group][a*4 + k] = one uint4 per unit.  A hit leaves a record (objectness, position) in `rec`, so that the
// plane is never read a second time.  NW warps, U units in flight per warp.
template <int NW, int CAPH>
__device__ __forceinline__ void scan_objectness(const HeadsDev& P, const float* const* lvbase, const uint32_t* utab,
                                                int lane, int warp, uint32_t* hitw, uint64_t* rec, int* nrec) {
#ifndef PQ_SCAN_U
#define PQ_SCAN_U 4
#endif
  constexpr int U = PQ_SCAN_U;
  const int A = P.A, nu = P.G_tot * A;
  uint4* hw = reinterpret_cast<uint4*>(hitw);
  for (int u0 = warp; u0 < nu; u0 += NW * U) {
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int un = u0 + u * NW;
      x[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (un < nu) {
        const uint32_t e = utab[un];
        const float* p = lvbase[e >> 30] + ((e & kUnitOffMask) + 4u * (unsigned)lane);
        if (!(e & kUnitSlow)) {
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "l"(p));
        } else {
          x[u] = scan_load_slow(P, lvbase[e >> 30], (int)(e >> 30), un, lane);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int un = u0 + u * NW;
      // conservative logit-space prefilter only; the exact conf > thr test runs later, once per surviving row and
      // in parallel (phase 2b).  A row that passes here but fails there (|x - logit(thr)| < 1e-3) merely occupies
      // a hit slot: its confidence is forced to 0, so it can never yield a candidate.
      const bool p0b = x[u].x > P.logit_lo, p1b = x[u].y > P.logit_lo, p2b = x[u].z > P.logit_lo, p3b = x[u].w > P.logit_lo;
      uint4 wd;
      wd.x = __ballot_sync(PQ_FULL, p0b); wd.y = __ballot_sync(PQ_FULL, p1b);
      wd.z = __ballot_sync(PQ_FULL, p2b); wd.w = __ballot_sync(PQ_FULL, p3b);
      if (un < nu) {
        if (lane == 0) hw[un] = wd;
        if (wd.x | wd.y | wd.z | wd.w) {                               // warp-uniform, ~1 unit in 3 on natural images
          // the rows leave a record (objectness logit, position): ONE shared-memory atomic per unit reserves the
          // unit's records, every lane derives its own offsets from the ballot words
          const int n0 = __popc(wd.x), n1 = __popc(wd.y), n2 = __popc(wd.z);
          int base = 0;
          if (lane == 0) base = atomicAdd(nrec, n0 + n1 + n2 + __popc(wd.w));
          base = __shfl_sync(PQ_FULL, base, 0);
          const unsigned lt = (1u << lane) - 1u;
          const int g = un / A;
          const uint32_t pos = ((uint32_t)g << 10) | ((uint32_t)(un - g * A) << 7) | ((uint32_t)lane << 2);
          auto put = [&](bool hit, float xv, int at, uint32_t k) {
            if (hit && at < CAPH) rec[at] = ((uint64_t)__float_as_uint(xv) << 32) | (pos | k);
          };
          put(p0b, x[u].x, base + __popc(wd.x & lt), 0);
          put(p1b, x[u].y, base + n0 + __popc(wd.y & lt), 1);
          put(p2b, x[u].z, base + n0 + n1 + __popc(wd.z & lt), 2);
          put(p3b, x[u].w, base + n0 + n1 + n2 + __popc(wd.w & lt), 3);
        }
      }
    }
  }
}

// SRC == 0: candidates come from the raw heads (decode + recover fused in).  SRC == 1: from a recovered
// (B, N, 4+C) tensor, i.e. tools.torch_nms for a whole batch in one launch: the front end streams every row
// once (no early-out is possible, the scores are already formed), the sort / NMS / output back end is shared.
template <int ROUND, int SRC, int CLS>
__global__ void __launch_bounds__(FusedCfg<CLS>::kThreads, FusedCfg<CLS>::kMinCtas)
decode_nms_fused_kernel(const __grid_constant__ HeadsDev P, const __grid_constant__ DetOut O, int32_t* work) {
  constexpr int NT = FusedCfg<CLS>::kThreads, NW = NT / 32;
  constexpr int CAPH = FusedCfg<CLS>::kCapH, CAPM = FusedCfg<CLS>::kCapM;
  using Smem = FusedSmemT<CAPH, CAPM>;
  static_assert((8 * CAPM + 16 * CAPH + 4 * CAPH) % 16 == 0, "output keys are read two at a time");
  static_assert(CAPH % NT == 0 && CAPM % NT == 0, "per-thread register tiles");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  uint32_t* hitw = reinterpret_cast<uint32_t*>(smem_raw + sizeof(Smem));
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int A = P.A, C = P.C, ch = P.ch;
  const int WG = (SRC == 1) ? 1 : 4 * A;   // ballot words per group (128 cells x A anchors | 32 rows)
  const int W_tot = P.G_tot * WG;
  uint32_t* gbase = hitw + W_tot;
  uint32_t* utab = gbase + P.G_tot;
  if (SRC == 0) build_unit_table(P, utab, tid, NT);

#ifdef PQ_PHASE_TIMING
  long long pq_t0 = clock64();
#endif
  // Programmatic dependent launch (launch_fused sets the attribute when every CTA owns exactly one image): this grid may
  // start while the previous one on the stream is still draining its slowest images.  Everything up to the output
  // stage reads only the caller's inputs and writes shared memory; the outputs (possibly the very buffers the
  // previous launch is still writing) are touched only behind cudaGridDependencySynchronize(), i.e. after that grid
  // has completed.  Without the attribute both calls return at once.
  cudaTriggerProgrammaticLaunchCompletion();
  const bool static_sched = ((int)gridDim.x == P.B);        // one image per CTA: no scheduler words (shared by launches)
  bool first = true;
  for (;;) {
    int b;
    if (static_sched) {
      if (!first) break;
      first = false;
      b = (int)blockIdx.x;
    } else {
      if (tid == 0) S.b = atomicAdd(work, 1);
      __syncthreads();
      b = S.b;
    }
    if (b >= P.B) break;
    PQ_PHASE(0);
    if (SRC == 0 && tid < P.n_levels) S.lvbase[tid] = P.lv[tid].raw + ((size_t)b * A * ch + 4) * P.lv[tid].HW;
    for (int i = tid; i < 128; i += NT) { S.cls_cnt[i] = 0; S.cls_fill[i] = 0; }
    if (tid == 0) S.nrec = 0;
    __syncthreads();
    const float* img = (SRC == 1) ? P.bboxes + (size_t)b * P.N * P.bb_row : nullptr;

    // ---- 1. scan: one ballot word per (level, group, anchor, sub-cell) | per 32 rows ---------------
    if (SRC == 0) {
      scan_objectness<NW, CAPH>(P, S.lvbase, utab, lane, warp, hitw, S.keys, &S.nrec);
    } else if (SRC == 2) {
      // the scan already happened in the head convolution's epilogue: the image's hit records sit in global memory
    } else {
      // a row is a hit iff any of its C scores exceeds thr (tools.py:551); 128-bit loads when rows are aligned
      const bool vec = ((P.bb_row & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.bboxes) & 15) == 0);
      if (vec && P.magic_n4 != 0) {
        // The image is one contiguous run of N * n4 128-bit words (n4 per row): consecutive threads read consecutive
        // words, so every byte crosses L2 once, fully coalesced; a word with a score above thr sets its row's bit.
        const int n4 = P.bb_row >> 2;
        for (int w = tid; w < W_tot; w += NT) hitw[w] = 0u;
        __syncthreads();
        const float4* img4 = reinterpret_cast<const float4*>(img);
        const int total4 = P.N * n4;
        constexpr int U = 4;
        for (int e0 = tid; e0 < total4; e0 += NT * U) {
          float4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < total4)
              asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(img4 + e));
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e >= total4) continue;
            const int row = (int)__umulhi((unsigned)e, P.magic_n4);           // e / n4
            if (e == row * n4) continue;                                       // word 0 of a row = the box
            if ((v[u].x > P.thr_f) | (v[u].y > P.thr_f) | (v[u].z > P.thr_f) | (v[u].w > P.thr_f))
              atomicOr(&hitw[row >> 5], 1u << (row & 31));
          }
        }
      } else
      for (int w = warp; w < W_tot; w += NW) {
        const int row = w * 32 + lane;
        bool pass = false;
        if (row < P.N) {
          const float* r = img + (size_t)row * P.bb_row;
          if (vec) {
            const float4* r4 = reinterpret_cast<const float4*>(r);
            const int n4 = P.bb_row >> 2;
#pragma unroll 4
            for (int q = 1; q < n4; ++q) {
              float4 v;
              asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(r4 + q));
              pass |= (v.x > P.thr_f) | (v.y > P.thr_f) | (v.z > P.thr_f) | (v.w > P.thr_f);
            }
          } else {
#pragma unroll 4
            for (int c = 0; c < C; ++c) pass |= ldg_stream(r + 4 + c) > P.thr_f;
          }
        }
        const unsigned word = __ballot_sync(PQ_FULL, pass);
        if (lane == 0) hitw[w] = word;
      }
    }
    __syncthreads();
    PQ_PHASE(1);

    // ---- 2. deterministic slots: exclusive prefix of the per-group hit counts --------------------
    if (SRC == 2) {
      if (tid == 0) { S.H = P.rec_count[b]; S.M = 0; S.K = 0; S.nbig = 0; S.next_class = 0; }
    } else if (warp == 0) {
      int running = 0;
      for (int g0 = 0; g0 < P.G_tot; g0 += 32) {
        const int g = g0 + lane;
        int c = 0;
        if (g < P.G_tot)
          for (int t = 0; t < WG; ++t) c += __popc(hitw[g * WG + t]);
        const int inc = warp_inclusive_sum(c);
        if (g < P.G_tot) gbase[g] = running + inc - c;
        running += __shfl_sync(PQ_FULL, inc, 31);
      }
      if (lane == 0) { S.H = running; S.M = 0; S.K = 0; S.nbig = 0; S.next_class = 0; }
    }
    __syncthreads();
    PQ_PHASE(2);
    const int H = S.H;
    if (H > CAPH || (SRC == 2 && H > P.rec_cap)) {
      cudaGridDependencySynchronize();
      if (tid == 0) { O.status[b] = PQDET_ST_CAND_OVERFLOW; set_count(O, b, 0); O.ncand[b] = -1; signal_peers(O); }
      __syncthreads();
      continue;
    }
    if (SRC == 2) {
      // Records arrive in the order the epilogue warps appended them; the hit slot of a row is its rank among the
      // image's rows (row order = nonzero() order, which is what breaks score ties).  order[] (free until step 5)
      // keeps slot -> record for the fetch below.
      const float* recs = P.rec + (size_t)b * P.rec_cap * (size_t)(6 + C);
      uint32_t* srow = reinterpret_cast<uint32_t*>(S.keys);
      for (int i = tid; i < H; i += NT) srow[i] = (uint32_t)__float_as_int(recs[(size_t)i * (6 + C)]);
      __syncthreads();
      for (int i = tid; i < H; i += NT) {
        const uint32_t row = srow[i];
        int rank = 0;
        for (int j = 0; j < H; ++j) rank += (srow[j] < row) ? 1 : 0;
        const int l = level_of_row(P, (int)row);
        const int il = (int)row - P.lv[l].row_off;
        const int cell = il / A;
        S.hmeta[rank] = pack_meta(l, il - cell * A, cell);
        const float conf = sigmoidf_(recs[(size_t)i * (6 + C) + 1]);
        S.hconf[rank] = (conf > P.thr_f) ? conf : 0.0f;
        S.hhas[rank] = 0;
        S.order[rank] = (uint16_t)i;
      }
    } else if (SRC == 1) {
      for (int w = tid; w < W_tot; w += NT) {    // word w = rows 32w .. 32w+31, already in row order
        unsigned word = hitw[w];
        int h = gbase[w];
        while (word) {
          const int j = __ffs(word) - 1;
          word &= word - 1;
          S.hmeta[h] = (uint32_t)(w * 32 + j);
          S.hhas[h] = 0;
          ++h;
        }
      }
    } else {
      // one hit record per thread: word a*4 + k of group g holds the cells 4*lane + k of anchor a
      uint32_t pos[CAPH / NT];
      float cf[CAPH / NT];
#pragma unroll
      for (int r = 0; r < CAPH / NT; ++r) {
        const int i = tid + r * NT;
        if (i < H) {
          const uint64_t rc = S.keys[i];
          pos[r] = (uint32_t)rc;
          // exact objectness test (model/parser.py:230 + tools.py:551: score = prob*conf <= conf): a row the
          // logit-space prefilter let through by its margin gets confidence 0 and can never yield a candidate
          const float conf = sigmoidf_(__uint_as_float((uint32_t)(rc >> 32)));
          cf[r] = (conf > P.thr_f) ? conf : 0.0f;
        }
      }
#pragma unroll
      for (int r = 0; r < CAPH / NT; ++r) {
        const int i = tid + r * NT;
        if (i >= H) continue;
        const int g = pos[r] >> 10, a = (pos[r] >> 7) & 7, j = (pos[r] >> 2) & 31, k = pos[r] & 3;
        const int l = level_of_group(P, g);
        const uint32_t* gw = hitw + g * WG;
        const unsigned below = (1u << j) - 1u;
        // slot = hits of the group that precede (lane j, k, a) in (cell, anchor) order
        int h = gbase[g];
        for (int t2 = 0; t2 < WG; ++t2) {
          const unsigned w2 = gw[t2];
          h += __popc(w2 & below);
          const int a2 = t2 >> 2, k2 = t2 & 3;
          if (k2 < k || (k2 == k && a2 < a)) h += (w2 >> j) & 1u;
        }
        S.hmeta[h] = pack_meta(l, a, (g - P.lv[l].group_off) * 128 + 4 * j + k);
        S.hconf[h] = cf[r];
        S.hhas[h] = 0;
      }
    }
    __syncthreads();
    PQ_PHASE(3);

    // ---- 3. box + class channels of the hit rows only -------------------------------------------
    if (SRC == 1) {
      const int CK = 4 + C;
      const int sub = lane >> 3, k0 = lane & 7;
      for (int h = warp * 4 + sub; h < H; h += NW * 4) {
        const float* r = img + (size_t)S.hmeta[h] * P.bb_row;
        for (int k = k0; k < CK; k += 8) {
          const float v = r[k];
          if (k < 4) {
            reinterpret_cast<float*>(&S.hbox[h])[k] = v;
          } else if (v > P.thr_f) {
            const unsigned peers = __activemask();
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&S.M, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            const int slot = base + __popc(peers & ((1u << lane) - 1u));
            if (slot < CAPM) S.keys[slot] = cand_key(k - 4, v, (uint32_t)h);
            atomicAdd(&S.cls_cnt[k - 4], 1);
            S.hhas[h] = 1;
          }
        }
      }
    } else {
      // A unit = (32 consecutive hit rows, 8 consecutive channels): lane = row, so one load instruction asks for
      // the SAME channel of 32 rows.  Hit rows are in (cell, anchor) order and real objects light up runs of
      // neighbouring cells, so the rows of one instruction fall into few 32-byte sectors of that channel plane
      // (fewer memory requests; this is what bounds the kernel when the heads live in pinned host memory and
      // every distinct sector is one PCIe read).  8 loads are in flight per lane.
      // A unit = (32 consecutive hit rows, 8 consecutive channels): lane = row, so one load instruction asks for
      // the SAME channel of 32 rows.  Hit rows are in (cell, anchor) order and real objects light up runs of
      // neighbouring cells, so the rows of one instruction fall into few 32-byte sectors of that channel plane
      // (fewer memory requests; this is what bounds the kernel when the heads live in pinned host memory and
      // every distinct sector is one PCIe read).  8 x UP loads are in flight per lane.  The candidates of a unit are
      // appended with ONE reservation (a warp prefix sum + one shared-memory atomic) and the class counters are bumped
      // by fire-and-forget atomics: nothing in the loop waits for an atomic's return value per channel.
      const Affine af = image_affine(P, b);
      const int CK = 4 + C;
#ifndef PQ_FETCH_UP
#define PQ_FETCH_UP 1
#endif
#ifndef PQ_FETCH_V
#define PQ_FETCH_V 8
#endif
      constexpr int V = PQ_FETCH_V, UP = PQ_FETCH_UP;
      const int ncg = (CK + V - 1) / V;
      const int nunit = ((H + 31) >> 5) * ncg;
      for (int unit0 = warp; unit0 < nunit; unit0 += NW * UP) {
        float v[UP][V];
        int klo[UP], hrow[UP];
#pragma unroll
        for (int q = 0; q < UP; ++q) {
          const int unit = unit0 + q * NW;
          const int chunk = unit / ncg;
          klo[q] = (unit - chunk * ncg) * V;
          const int h = chunk * 32 + lane;
          hrow[q] = (unit < nunit && h < H) ? h : -1;
          if (hrow[q] >= 0) {
            const uint32_t meta = S.hmeta[h];
            const int l = meta >> 30, a = (meta >> 27) & 7, cell = meta & 0x7ffffff;
            const int HW = P.lv[l].HW;
            if (SRC == 2) {
              // the row's raw values from its hit record: [row, objectness, 4 box, C class]
              const float* rr = P.rec + ((size_t)b * P.rec_cap + S.order[h]) * (size_t)(6 + C) + 2;
#pragma unroll
              for (int u = 0; u < V; ++u) {
                const int k = klo[q] + u;
                if (k < CK) v[q][u] = __ldg(rr + k);
              }
            } else {
              // channel 0 of this row: the anchor's objectness plane minus four planes
              const float* base = S.lvbase[l] + ((size_t)a * ch - 4) * HW + cell;
#pragma unroll
              for (int u = 0; u < V; ++u) {
                const int k = klo[q] + u;
                if (k < CK) v[q][u] = ldg_stream(base + (size_t)((k < 4) ? k : k + 1) * HW);
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < UP; ++q) {
          if (unit0 + q * NW >= nunit) break;                 // warp-uniform
          const int h = hrow[q];
          const bool valid = h >= 0;
          unsigned pass = 0;                                  // bit u: channel klo + u is a candidate of this row
          if (valid) {
            const float conf = S.hconf[h];
            if (klo[q] == 0) {
              const uint32_t meta = S.hmeta[h];               // re-read: cheaper than carrying it across the loads
              const LevelDev& L = P.lv[meta >> 30];
              const int cell = meta & 0x7ffffff;
              const int cy = cell / L.W, cx = cell - cy * L.W;
              float4 bx;
              bx.x = recover_coord(0, decode_coord(0, v[q][0], cx, cy, L.stride), af);
              bx.y = recover_coord(1, decode_coord(1, v[q][1], cx, cy, L.stride), af);
              bx.z = recover_coord(2, decode_coord(2, v[q][2], cx, cy, L.stride), af);
              bx.w = recover_coord(3, decode_coord(3, v[q][3], cx, cy, L.stride), af);
              S.hbox[h] = bx;
            }
#pragma unroll
            for (int u = 0; u < V; ++u) {
              const int k = klo[q] + u;
              if (k >= 4 && k < CK) {
                v[q][u] = PQ_MUL(sigmoidf_(v[q][u]), conf);
                if (v[q][u] > P.thr_f) pass |= 1u << u;
              }
            }
          }
          // class counters: one fire-and-forget atomic per channel that has candidates
#pragma unroll
          for (int u = 0; u < V; ++u) {
            const unsigned bal = __ballot_sync(PQ_FULL, (pass >> u) & 1u);
            if (lane == u && bal) atomicAdd(&S.cls_cnt[klo[q] + u - 4], __popc(bal));
          }
          const int cnt = __popc(pass);
          const int incl = warp_inclusive_sum(cnt);
          const int total = __shfl_sync(PQ_FULL, incl, 31);
          if (total) {                                        // warp-uniform
            int base2 = 0;
            if (lane == 0) base2 = atomicAdd(&S.M, total);
            base2 = __shfl_sync(PQ_FULL, base2, 0);
            int slot = base2 + incl - cnt;
            if (cnt) S.hhas[h] = 1;
#pragma unroll
            for (int u = 0; u < V; ++u)
              if ((pass >> u) & 1u) {
                if (slot < CAPM) S.keys[slot] = cand_key(klo[q] + u - 4, v[q][u], (uint32_t)h);
                ++slot;
              }
          }
        }
      }
    }
    __syncthreads();
    PQ_PHASE(4);
    const int M = S.M;
    if (M > CAPM || M == 0) {
      cudaGridDependencySynchronize();
      if (tid == 0) {
        O.status[b] = M ? PQDET_ST_CAND_OVERFLOW : PQDET_ST_OK;
        set_count(O, b, 0);
        O.ncand[b] = M;
        signal_peers(O);
      }
      __syncthreads();
      continue;
    }

    // ---- 4. coordinate-trick offset base + class segment starts ---------------------------------
    const bool trick = use_trick(P.nms_mode, M);
    float m1 = 0.0f;
    {
      float mx = -INFINITY;
      if (trick)
        for (int h = tid; h < H; h += NT)
          if (S.hhas[h]) {
            const float4 bx = S.hbox[h];
            mx = fmaxf(fmaxf(mx, fmaxf(bx.x, bx.y)), fmaxf(bx.z, bx.w));
          }
      mx = warp_max(mx);
      if (lane == 0) S.red[warp] = mx;
      if (warp == NW - 1) {                                // exclusive prefix of the class counts
        int running = 0;
        for (int c0 = 0; c0 < C; c0 += 32) {
          const int c = c0 + lane;
          const int n = (c < C) ? S.cls_cnt[c] : 0;
          const int inc = warp_inclusive_sum(n);
          if (c < C) S.seg_start[c] = running + inc - n;
          running += __shfl_sync(PQ_FULL, inc, 31);
        }
      }
      __syncthreads();
      mx = S.red[0];
#pragma unroll
      for (int i = 1; i < NW; ++i) mx = fmaxf(mx, S.red[i]);
      m1 = PQ_ADD(mx, 1.0f);
    }

    PQ_PHASE(5);
    // ---- 5. per-class member lists --------------------------------------------------------------
    for (int i = tid; i < M; i += NT) {
      const int c = (int)(S.keys[i] >> 57);
      S.order[S.seg_start[c] + atomicAdd(&S.cls_fill[c], 1)] = (uint16_t)i;
    }
    __syncthreads();
    PQ_PHASE(6);

    // ---- 6. greedy NMS: one warp per class (classes pulled dynamically: uneven sizes), the whole CTA for the
    // classes beyond kWarpClassMax candidates; kept detections append their output key themselves ----------------
    for (;;) {
      int c = 0;
      if (lane == 0) c = atomicAdd(&S.next_class, 1);
      c = __shfl_sync(PQ_FULL, c, 0);
      if (c >= C) break;
      const int n = S.cls_cnt[c];
      if (n == 0) continue;
      if (n > kWarpClassMax) {
        if (lane == 0) S.big[atomicAdd(&S.nbig, 1)] = c;
        continue;
      }
      const int s = S.seg_start[c];
      const float off = trick ? PQ_MUL((float)c, m1) : 0.0f;
      if (n <= 32) warp_select_nms<ROUND, 1>(S, s, n, off, P.iou_f, P.iou_d);
      else if (n <= 64) warp_select_nms<ROUND, 2>(S, s, n, off, P.iou_f, P.iou_d);
      else warp_select_nms<ROUND, 4>(S, s, n, off, P.iou_f, P.iou_d);
    }
    __syncthreads();
    {
      const int nbig = S.nbig;
      for (int i = 0; i < nbig; ++i) {
        const int c = S.big[i];
        cta_select_nms<ROUND, NT>(S, S.seg_start[c], S.cls_cnt[c], trick ? PQ_MUL((float)c, m1) : 0.0f, P.iou_f, P.iou_d);
      }
    }
    __syncthreads();
    PQ_PHASE(7);

    // ---- 7. kept keys -> (score desc, row, class) order -> output -------------------------------
    const int K = S.K;
    PQ_PHASE(8);
    cudaGridDependencySynchronize();                       // the previous launch has completed: outputs may be written
    if (K > Smem::kOutCap) {                               // more kept detections than the output list holds
      if (tid == 0) { O.status[b] = PQDET_ST_CAND_OVERFLOW; set_count(O, b, 0); O.ncand[b] = M; signal_peers(O); }
      __syncthreads();
      continue;
    }
    {
      // position = number of smaller keys (the keys are unique); two keys per shared-memory load
      const uint64_t* ok = S.okeys();
      for (int i = tid; i < K; i += NT) {
        const uint64_t mine = ok[i];
        int rank = 0;
        const int K2 = K & ~1;
        for (int j = 0; j < K2; j += 2) {
          const ulonglong2 o = *reinterpret_cast<const ulonglong2*>(ok + j);
          rank += (o.x < mine) ? 1 : 0;
          rank += (o.y < mine) ? 1 : 0;
        }
        if (K2 < K) rank += (ok[K2] < mine) ? 1 : 0;
        if (rank < O.max_det) {
          const float score = __uint_as_float(~(uint32_t)(mine >> 32));
          const uint32_t low = (uint32_t)mine;
          const int h = low >> 7, c = low & 127;
          const uint32_t meta = S.hmeta[h];
          int64_t row = meta;
          if (SRC != 1) {
            const LevelDev& L = P.lv[meta >> 30];
            row = L.row_off + (int64_t)(meta & 0x7ffffff) * A + ((meta >> 27) & 7);
          }
          const float4 bx = S.hbox[h];
          write_det(O, b, rank, bx, score, c, row, C);
          if (O.n_peers && rank < O.gather_cap) {          // gather mode: the row is also staged for the peers
            constexpr int kStageRows = (int)(sizeof(S.keys) / 24);          // keys[] is free once the NMS is done
            if (rank < kStageRows) {
              float2* st = reinterpret_cast<float2*>(S.keys) + 3 * rank;
              st[0] = make_float2(bx.x, bx.y);
              st[1] = make_float2(bx.z, bx.w);
              st[2] = make_float2(score, (float)c);
            } else {                                       // beyond the staging area (> 426 rows of one image): direct
              const size_t at = ((size_t)(O.img_off + b) * O.gather_cap + rank) * 6;
              for (int p = 0; p < O.n_peers; ++p) {
                float2* g = reinterpret_cast<float2*>(O.peer_det[p] + at);
                g[0] = make_float2(bx.x, bx.y);
                g[1] = make_float2(bx.z, bx.w);
                g[2] = make_float2(score, (float)c);
              }
            }
          }
        }
      }
      if (O.n_peers) {
        // the image's rows are one contiguous run in every rank's gathered buffer: they cross NVLink as coalesced
        // 16-byte stores (a handful of 128-byte packets per image and peer) instead of three 8-byte stores per row
        __syncthreads();
        const int nrow = min(K, min(O.gather_cap, (int)(sizeof(S.keys) / 24)));
        const int n2 = nrow * 3, n4 = n2 >> 1;             // float2 / float4 units (gather_cap is even: aligned base)
        const float4* src4 = reinterpret_cast<const float4*>(S.keys);
        const size_t at = (size_t)(O.img_off + b) * O.gather_cap * 6;
        for (int p = 0; p < O.n_peers; ++p) {
          float4* dst4 = reinterpret_cast<float4*>(O.peer_det[p] + at);
          for (int i = tid; i < n4; i += NT) dst4[i] = src4[i];
          if ((n2 & 1) && tid == 0)
            reinterpret_cast<float2*>(O.peer_det[p] + at)[n2 - 1] = reinterpret_cast<const float2*>(S.keys)[n2 - 1];
        }
      }
    }
    if (O.n_peers && O.peer_arr[0]) {                      // every thread's peer stores before the arrival signal
      __threadfence_system();
      __syncthreads();
    }
    if (tid == 0) {
      set_count(O, b, K);
      O.ncand[b] = M;
      O.status[b] = (K > O.max_det) ? PQDET_ST_DET_TRUNCATED : PQDET_ST_OK;
      signal_peers(O);
    }
    __syncthreads();
    PQ_PHASE(9);
  }
  // re-arm the scheduler for the next launch on this stream: the last CTA to leave zeroes both words
  // (work[0] = next image, work[1] = CTAs that have finished), so steady-state calls need no memset
  if (tid == 0 && !static_sched) {
    __threadfence();
    if (atomicAdd(work + 1, 1) == (int)gridDim.x - 1) {
      work[0] = 0;
      work[1] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// general path: any candidate count, candidates in global memory bucketed by (image, class)
// ------------------------------------------------------------------------------------------------
struct GenParams {
  HeadsDev heads;            // valid when from_heads
  int from_heads;
  const float* bboxes;       // (B, N, 4+C) when !from_heads
  int64_t N;
  int C;
  float thr_f;
  float iou_f;
  double iou_d;
  int nms_mode;
  const int32_t* image_ids;  // nullable
  int n_images;
  int out_by_pos;            // outputs indexed by position in image_ids instead of image id
  // workspace
  uint32_t* cls_count;       // [n_images][C]   candidates per bucket
  uint32_t* cls_fill;        // [n_images][C]   scatter cursors
  int64_t* seg_off;          // [n_images][C]   bucket start in `keys`
  uint32_t* seg_cap;         // [n_images][C]   bucket capacity (pow2 >= count)
  uint32_t* max_ord;         // [n_images]      max coordinate over picked boxes (ordered uint)
  int64_t* img_off;          // [n_images]      start of the image's kept-key region in `kkeys`
  uint32_t* img_cap;         // [n_images]      its capacity (pow2 >= M)
  uint32_t* kept_count;      // [n_images]
  uint32_t* run_off;         // [n_images][C]   where the bucket's kept keys start in the image's kept region ...
  uint32_t* run_len;         // [n_images][C]   ... and how many there are (they are in score order)
  uint64_t* keys;            // [cand_capacity] bucketed candidate keys (~score | row)
  uint64_t* kkeys;           // [cand_capacity] kept keys per image
  uint32_t* klist;           // [cand_capacity] kept positions per bucket
  float4* rbox;              // [n_images][N]   recovered boxes of hit rows (from_heads only)
  // single-pass select: candidates are first appended, in arrival order and with their class in the key, to a
  // per-image stage of `stage_quota` keys (aliased onto kkeys, which is not needed before the bucket kernels)
  uint32_t* stage_fill;      // [n_images]
  int64_t stage_quota;
  int64_t cand_capacity;
  int64_t* needed;
  int32_t* ok;               // 1 when the workspace is large enough
};

__device__ __forceinline__ int gen_image(const GenParams& G, int i) {
  return G.image_ids ? G.image_ids[i] : i;
}
__device__ __forceinline__ int gen_out(const GenParams& G, int i) {
  return G.out_by_pos ? i : gen_image(G, i);
}

// Row -> (recovered box, objectness) for the heads source; thread-per-row in (level, anchor, cell)
// order so that plane reads coalesce.
__device__ __forceinline__ bool gen_row_from_index(const HeadsDev& P, int64_t i, int& level, int& a, int& cell,
                                                   int& row) {
  if (i >= P.N) return false;
  int l = 0;
#pragma unroll
  for (int q = 1; q < PQDET_MAX_LEVELS; ++q)
    if (q < P.n_levels && i >= P.lv[q].row_off) l = q;
  const LevelDev& L = P.lv[l];
  const int il = (int)(i - L.row_off);
  a = il / L.HW;
  cell = il - a * L.HW;
  level = l;
  row = L.row_off + cell * P.A + a;
  return true;
}

// pass 0: count candidates per (image, class) + max picked coordinate; pass 1: scatter keys into the buckets.
// pass 2 = passes 0 and 1 in ONE read of the heads: counts as pass 0, and the keys - with the class in bits 25..31 -
// are appended to the image's stage (one global atomic per CTA); gen_bucketize_kernel moves them into the buckets
// once gen_plan_kernel has laid those out.  (Pass 1 re-evaluates every row: 85 us per 64 dense 608 x 608 images.)
// Every thread evaluates one row and remembers which classes pass as a 128-bit mask.  Counts and slot
// reservations go through a per-CTA shared-memory histogram, so the global counters see one atomic per
// (CTA, class) instead of one per candidate (dense scenes: ~12k candidates per image on C counters).
template <int PASS>
__global__ void __launch_bounds__(256)
gen_select_kernel(const __grid_constant__ GenParams G) {
  if (PASS == 1 && *G.ok == 0) return;
  __shared__ unsigned s_cnt[128];
  __shared__ unsigned s_base[128];
  __shared__ unsigned s_max;
  const int ii = blockIdx.y;
  const int b = gen_image(G, ii);
  const int C = G.C;
  const int tid = threadIdx.x;
  if (tid < 128) s_cnt[tid] = 0;
  if (tid == 0) s_max = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + tid;
  float box[4] = {0.f, 0.f, 0.f, 0.f};
  float conf = 1.0f;
  int row = 0;
  const float* scores = nullptr;   // !from_heads: pointer to the C scores of the row
  const float* cls0 = nullptr;     // from_heads: first class plane element of the row (stride HW between classes)
  int HW = 0;
  bool live = false;
  if (G.from_heads) {
    const HeadsDev& P = G.heads;
    int level = 0, a = 0, cell = 0;
    if (gen_row_from_index(P, i, level, a, cell, row)) {
      const LevelDev& L = P.lv[level];
      HW = L.HW;
      const float* r0 = L.raw + (size_t)(b * P.A + a) * P.ch * L.HW + cell;
      const float x = r0[(size_t)4 * L.HW];
      if (x > P.logit_lo) {
        // the box channels are requested before the objectness is evaluated: one round trip instead of two for the
        // rows that pass (this path serves dense scenes; the prefilter already dropped the rows far below thr)
        float rb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rb[k] = r0[(size_t)k * L.HW];
        conf = sigmoidf_(x);
        if (conf > G.thr_f) {
          live = true;
          cls0 = r0 + (size_t)5 * L.HW;
          {
            const Affine af = image_affine(P, b);
            const int cy = cell / L.W, cx = cell - cy * L.W;
#pragma unroll
            for (int k = 0; k < 4; ++k) box[k] = recover_coord(k, decode_coord(k, rb[k], cx, cy, L.stride), af);
          }
        }
      }
    }
  } else if (i < G.N) {
    live = true;
    row = (int)i;
    const float* r = G.bboxes + ((size_t)b * G.N + i) * (4 + C);
#pragma unroll
    for (int k = 0; k < 4; ++k) box[k] = r[k];
    scores = r + 4;
  }
  auto score_of = [&](int c) -> float {
    return G.from_heads ? PQ_MUL(sigmoidf_(cls0[(size_t)c * HW]), conf) : scores[c];
  };
  uint64_t m0 = 0, m1 = 0;          // classes of this row with score > thr
  float sc16[16];                   // scores of classes 0..15 (static indexing only; the rest is re-evaluated on use)
#pragma unroll
  for (int u = 0; u < 16; ++u) sc16[u] = 0.0f;
  if (G.from_heads) {
    // warp-uniform loop over batches of 8 classes: the 8 logits of a batch are requested together (one round trip per
    // batch instead of one per class), and the per-class counts go through a ballot (one shared atomic per warp and
    // class instead of one per candidate)
    if (__any_sync(PQ_FULL, live)) {
      const int lane = lane_id();
#pragma unroll 2
      for (int c0 = 0; c0 < C; c0 += 8) {
        float z[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) z[u] = (live && c0 + u < C) ? cls0[(size_t)(c0 + u) * HW] : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c0 + u;
          if (c < C) {                                      // warp-uniform
            const float sc = PQ_MUL(sigmoidf_(z[u]), conf); // == score_of(c)
            const bool pass = live && sc > G.thr_f;
            const unsigned bm = __ballot_sync(PQ_FULL, pass);
            if (pass) {
              if (c < 64) m0 |= 1ull << c; else m1 |= 1ull << (c - 64);
            }
            if (c0 == 0) sc16[u] = sc;
            else if (c0 == 8) sc16[8 + u] = sc;
            if (lane == 0 && bm) atomicAdd(&s_cnt[c], (unsigned)__popc(bm));
          }
        }
      }
    }
  } else if (live) {
    for (int c = 0; c < C; ++c) {
      if (score_of(c) > G.thr_f) {
        if (c < 64) m0 |= 1ull << c; else m1 |= 1ull << (c - 64);
        atomicAdd(&s_cnt[c], 1u);
      }
    }
  }
  const bool any = (m0 | m1) != 0;
  if (PASS == 2) {
    __shared__ unsigned s_wsum[8];
    __shared__ unsigned s_stage_base;
    const int lane = lane_id(), warp = warp_id();
    const unsigned mine = (unsigned)(__popcll(m0) + __popcll(m1));
    unsigned inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned o = __shfl_up_sync(PQ_FULL, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_wsum[warp] = inc;
    unsigned mo = any ? float_to_ordered(fmaxf(fmaxf(box[0], box[1]), fmaxf(box[2], box[3]))) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mo = max(mo, __shfl_xor_sync(PQ_FULL, mo, d));
    if (lane == 0 && mo) atomicMax(&s_max, mo);
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += s_wsum[w];
      total += s_wsum[w];
    }
    if (tid == 0) s_stage_base = total ? atomicAdd(&G.stage_fill[ii], total) : 0u;
    if (tid < C && s_cnt[tid]) atomicAdd(&G.cls_count[(size_t)ii * C + tid], s_cnt[tid]);
    if (tid == 0 && s_max) atomicMax(&G.max_ord[ii], s_max);
    __syncthreads();
    if (!any) return;
    uint64_t slot = (uint64_t)s_stage_base + before + inc - mine;
    uint64_t* stage = G.kkeys + (size_t)ii * (size_t)G.stage_quota;
    if (G.from_heads) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {                          // classes 0..15: scores kept in registers
        if ((m0 >> u) & 1ull) {
          if (slot < (uint64_t)G.stage_quota)
            stage[slot] = ((uint64_t)(~float_to_ordered(sc16[u])) << 32) | ((uint64_t)u << kHitBits) | (uint32_t)row;
          ++slot;
        }
      }
      m0 &= ~0xffffull;
    }
    for (int h = 0; h < 2; ++h) {
      uint64_t m = h ? m1 : m0;
      while (m) {
        const int c = __ffsll((long long)m) - 1 + 64 * h;
        m &= m - 1;
        const float sc = score_of(c);
        if (slot < (uint64_t)G.stage_quota)
          stage[slot] = ((uint64_t)(~float_to_ordered(sc)) << 32) | ((uint64_t)c << kHitBits) | (uint32_t)row;
        ++slot;
      }
    }
    if (G.from_heads) G.rbox[(size_t)ii * G.N + row] = make_float4(box[0], box[1], box[2], box[3]);
    return;
  }
  if (PASS == 0) {
    // max picked coordinate of the image: warp max, one shared atomic per warp, one global atomic per CTA
    unsigned mo = any ? float_to_ordered(fmaxf(fmaxf(box[0], box[1]), fmaxf(box[2], box[3]))) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mo = max(mo, __shfl_xor_sync(PQ_FULL, mo, d));
    if (lane_id() == 0 && mo) atomicMax(&s_max, mo);
  }
  __syncthreads();
  if (PASS == 0) {
    if (tid < C && s_cnt[tid]) atomicAdd(&G.cls_count[(size_t)ii * C + tid], s_cnt[tid]);
    if (tid == 0 && s_max) atomicMax(&G.max_ord[ii], s_max);
    return;
  }
  if (tid < C) {
    const unsigned n = s_cnt[tid];
    s_base[tid] = n ? atomicAdd(&G.cls_fill[(size_t)ii * C + tid], n) : 0u;
    s_cnt[tid] = 0;
  }
  __syncthreads();
  if (!any) return;
  for (int h = 0; h < 2; ++h) {
    uint64_t m = h ? m1 : m0;
    while (m) {
      const int c = __ffsll((long long)m) - 1 + 64 * h;
      m &= m - 1;
      const float sc = score_of(c);                        // same expression as above: bit-identical
      const unsigned slot = s_base[c] + atomicAdd(&s_cnt[c], 1u);
      // total-order transform: the general path also serves tools.torch_nms on arbitrary (possibly negative) scores
      G.keys[G.seg_off[(size_t)ii * C + c] + slot] = ((uint64_t)(~float_to_ordered(sc)) << 32) | (uint32_t)row;
    }
  }
  if (G.from_heads) G.rbox[(size_t)ii * G.N + row] = make_float4(box[0], box[1], box[2], box[3]);
}

// Single-pass select, second half: the staged keys of an image (arrival order, class in bits 25..31) go to their
// (image, class) buckets.  Same per-CTA histogram trick as pass 1, but on 8 bytes per candidate instead of a
// re-evaluation of the heads.
__global__ void __launch_bounds__(256)
gen_bucketize_kernel(const __grid_constant__ GenParams G) {
  if (*G.ok == 0) return;
  __shared__ unsigned s_cnt[128];
  __shared__ unsigned s_base[128];
  const int ii = blockIdx.y;
  const int C = G.C;
  const int tid = threadIdx.x;
  const uint32_t M = G.stage_fill[ii];
  if ((uint64_t)blockIdx.x * 256 >= M) return;
  if (tid < 128) s_cnt[tid] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * 256 + tid;
  uint64_t key = 0;
  int c = -1;
  unsigned local = 0;
  if (i < M) {
    key = G.kkeys[(size_t)ii * (size_t)G.stage_quota + i];
    c = (int)((key >> kHitBits) & 127u);
    local = atomicAdd(&s_cnt[c], 1u);
  }
  __syncthreads();
  if (tid < C) {
    const unsigned n = s_cnt[tid];
    s_base[tid] = n ? atomicAdd(&G.cls_fill[(size_t)ii * C + tid], n) : 0u;
  }
  __syncthreads();
  if (c >= 0)
    G.keys[G.seg_off[(size_t)ii * C + c] + s_base[c] + local] = (key & 0xffffffff00000000ull) | (key & kHitMask);
}

// Block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix
// and the block total through *total.
__device__ __forceinline__ unsigned long long block_exclusive_scan_1024(unsigned long long v,
                                                                        unsigned long long* warp_tot,
                                                                        unsigned long long* total) {
  const int lane = lane_id(), warp = warp_id();
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long n = __shfl_up_sync(PQ_FULL, inc, d);
    if (lane >= d) inc += n;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = warp_tot[lane];
    unsigned long long winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned long long n = __shfl_up_sync(PQ_FULL, winc, d);
      if (lane >= d) winc += n;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  const unsigned long long res = warp_tot[warp] + inc - v;
  __syncthreads();
  return res;
}

// Single CTA (1024 threads): bucket capacities (pow2 >= count), bucket / kept-region offsets,
// capacity check, per-image candidate counts.
__global__ void __launch_bounds__(1024)
gen_plan_kernel(const __grid_constant__ GenParams G, const __grid_constant__ DetOut O) {
  __shared__ unsigned long long s_wt[32];
  __shared__ unsigned long long s_tot_a, s_tot_b;
  const int C = G.C;
  unsigned long long base_seg = 0, base_img = 0;
  for (int i0 = 0; i0 < G.n_images; i0 += 1024) {
    const int ii = i0 + threadIdx.x;
    unsigned long long seg_tot = 0, M = 0, icap = 0;
    if (ii < G.n_images) {
      for (int c = 0; c < C; ++c) {
        const uint32_t n = G.cls_count[(size_t)ii * C + c];
        const uint32_t cap = n ? (uint32_t)next_pow2((int)n) : 0u;
        G.seg_cap[(size_t)ii * C + c] = cap;
        seg_tot += cap;
        M += n;
      }
      icap = M ? (unsigned long long)next_pow2((int)M) : 0ull;
      G.img_cap[ii] = (uint32_t)icap;
      O.ncand[gen_out(G, ii)] = (int32_t)M;
    }
    const unsigned long long ex_seg = block_exclusive_scan_1024(seg_tot, s_wt, &s_tot_a);
    const unsigned long long ex_img = block_exclusive_scan_1024(icap, s_wt, &s_tot_b);
    if (ii < G.n_images) {
      unsigned long long off = base_seg + ex_seg;
      for (int c = 0; c < C; ++c) {
        G.seg_off[(size_t)ii * C + c] = (int64_t)off;
        off += G.seg_cap[(size_t)ii * C + c];
      }
      G.img_off[ii] = (int64_t)(base_img + ex_img);
    }
    base_seg += s_tot_a;
    base_img += s_tot_b;
    __syncthreads();
  }
  unsigned long long need = base_seg > base_img ? base_seg : base_img;
  bool ok = need <= (unsigned long long)G.cand_capacity;
  if (G.stage_quota > 0) {
    // single-pass select: every image's candidates had to fit its stage; if not, ask for n_images x the largest
    __shared__ unsigned s_maxm;
    if (threadIdx.x == 0) s_maxm = 0;
    __syncthreads();
    for (int ii = threadIdx.x; ii < G.n_images; ii += 1024) atomicMax(&s_maxm, G.stage_fill[ii]);
    __syncthreads();
    const unsigned long long maxm = s_maxm;
    if (maxm > (unsigned long long)G.stage_quota) {
      ok = false;
      const unsigned long long want = maxm * (unsigned long long)G.n_images;
      if (want > need) need = want;
      if (need <= (unsigned long long)G.cand_capacity) need = (unsigned long long)G.cand_capacity + 1;
    }
  }
  if (threadIdx.x == 0) {
    *G.needed = (int64_t)need;
    *G.ok = ok ? 1 : 0;
  }
  for (int ii = threadIdx.x; ii < G.n_images; ii += 1024) {
    const int b = gen_out(G, ii);
    O.counts[b] = 0;
    O.status[b] = ok ? PQDET_ST_OK : PQDET_ST_CAND_OVERFLOW;
  }
}

// Two shared-memory tiers for the (image, class) buckets: up to 2048 candidates with 256 threads
// (56 KB, 4 CTAs/SM) and up to 8192 with 512 threads (224 KB, 1 CTA/SM); anything larger runs on global arrays.
#ifndef PQ_SEG_THREADS
#define PQ_SEG_THREADS 256
#endif
constexpr int kSegSmallKeys = 2048, kSegSmallThreads = PQ_SEG_THREADS;
constexpr int kSegBigKeys = 8192, kSegBigThreads = 512;

// Greedy NMS over a sorted bucket of n candidates by one CTA.  box(pos) = trick-shifted box of sorted
// position pos, kget/kset = kept-position list.  32 candidates per step:
//   A. every warp tests the step's 32 candidates against its stripe of the kept list;
//   B. every warp computes a slice of the step's 32x32 suppression matrix (row i = which later candidates box
//      i would suppress), independent of A;
//   C. warp 0 ORs the dead masks and walks the alive bits in order: keep i, clear the bits of row i.
// Only C is serial and it is a handful of bit operations per kept box.
template <int ROUND, int kSegWarps, typename Box, typename KGet, typename KSet>
__device__ __forceinline__ int bucket_greedy(uint32_t n, float iou_f, double iou_d, Box box, KGet kget, KSet kset,
                                             unsigned* s_dead, unsigned* s_rows, int* s_k) {
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  if (tid == 0) *s_k = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t p = base + lane;
    const bool valid = p < n;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) me = box(p);
    bool dead = !valid;
    const int k0 = *s_k;
    // A: four kept boxes per iteration - their loads and intersection chains are independent (the walk is bound by
    // dependent-issue latency, not by the pair count), and the division is only reached by a pair that intersects
    {
      int q = warp;
      for (; q + 3 * kSegWarps < k0; q += 4 * kSegWarps) {
        float4 a[4];
        float I[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = kget(q + u * kSegWarps);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float w = fmaxf(PQ_SUB(fminf(a[u].z, me.z), fmaxf(a[u].x, me.x)), 0.0f);
          const float h = fmaxf(PQ_SUB(fminf(a[u].w, me.w), fmaxf(a[u].y, me.y)), 0.0f);
          I[u] = PQ_MUL(w, h);
        }
        if ((I[0] > 0.0f) | (I[1] > 0.0f) | (I[2] > 0.0f) | (I[3] > 0.0f)) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (I[u] > 0.0f && !dead &&
                nms_suppresses<ROUND>(a[u].x, a[u].y, a[u].z, a[u].w, box_area(a[u].x, a[u].y, a[u].z, a[u].w), me.x, me.y,
                                      me.z, me.w, iou_f, iou_d))
              dead = true;
        }
      }
      for (; q < k0; q += kSegWarps) {
        const float4 a = kget(q);
        const float Sa = box_area(a.x, a.y, a.z, a.w);
        if (!dead && nms_suppresses<ROUND>(a.x, a.y, a.z, a.w, Sa, me.x, me.y, me.z, me.w, iou_f, iou_d)) dead = true;
      }
    }
    const unsigned dm = __ballot_sync(PQ_FULL, dead);
    if (lane == 0) s_dead[warp] = dm;
    __syncthreads();
    // B: rows of the step's suppression matrix - only for the candidates that survived A (typically a fifth of the
    // step in dense scenes), dealt round robin to the warps
    unsigned alldead = 0;
#pragma unroll
    for (int q = 0; q < kSegWarps; ++q) alldead |= s_dead[q];
    {
      unsigned todo = ~alldead;
      int nth = 0;
      while (todo) {
        const int i = __ffs(todo) - 1;
        todo &= todo - 1;
        if ((nth++ % kSegWarps) != warp) continue;
        const float ax1 = __shfl_sync(PQ_FULL, me.x, i), ay1 = __shfl_sync(PQ_FULL, me.y, i);
        const float ax2 = __shfl_sync(PQ_FULL, me.z, i), ay2 = __shfl_sync(PQ_FULL, me.w, i);
        const float Sa = box_area(ax1, ay1, ax2, ay2);
        const bool sup = (lane > i) && valid &&
                         nms_suppresses<ROUND>(ax1, ay1, ax2, ay2, Sa, me.x, me.y, me.z, me.w, iou_f, iou_d);
        const unsigned row = __ballot_sync(PQ_FULL, sup);
        if (lane == 0) s_rows[i] = row;
      }
    }
    __syncthreads();
    if (warp == 0) {                                       // C
      const unsigned myrow = s_rows[lane];
      unsigned alive = ~alldead, kept = 0;
      while (alive) {
        const int i = __ffs(alive) - 1;
        kept |= 1u << i;
        alive &= ~(1u << i);
        alive &= ~__shfl_sync(PQ_FULL, myrow, i);
      }
      if ((kept >> lane) & 1u) kset(k0 + __popc(kept & ((1u << lane) - 1u)), p, me);
      if (lane == 0) *s_k = k0 + __popc(kept);
    }
    __syncthreads();
  }
  return *s_k;
}

// One CTA per (image, class) bucket: sort, greedy NMS, append the kept keys to the image's kept region.
// Buckets of up to kSegSmemKeys candidates keep keys, shifted boxes and the kept list in shared memory
// (the inner loop then touches no global memory); larger ones run the same code on global arrays.
template <int ROUND, int kSegSmemKeys, int kSegThreads>
__global__ void __launch_bounds__(kSegThreads)
gen_bucket_nms_kernel(const __grid_constant__ GenParams G) {
  if (*G.ok == 0) return;
  constexpr int kSegWarps = kSegThreads / 32;
  extern __shared__ __align__(16) unsigned char seg_smem[];
  // one region of 16 bytes per key: the keys while sorting, then (the sorted keys parked in global memory, they are
  // only needed again for the output) the trick-shifted boxes; behind it the kept positions.  18 bytes per key =
  // 36 KB for the small tier: 6 CTAs per SM, so that 64 images x 10 classes run as a single wave.
  uint64_t* skeys = reinterpret_cast<uint64_t*>(seg_smem);
  float4* sbox = reinterpret_cast<float4*>(seg_smem);
  uint16_t* skl = reinterpret_cast<uint16_t*>(seg_smem + sizeof(float4) * kSegSmemKeys);
  __shared__ unsigned s_dead[kSegWarps];
  __shared__ unsigned s_rows[32];
  __shared__ int s_k;
  __shared__ uint32_t s_dst;
  const int C = G.C;
  const int ii = blockIdx.y, c = blockIdx.x;
  const uint32_t n = G.cls_count[(size_t)ii * C + c];
  if (n == 0) return;
  const int b = gen_image(G, ii);
  const uint32_t cap = G.seg_cap[(size_t)ii * C + c];
  // tier dispatch: the small launch takes cap <= 2048 and the > 8192 leftovers, the big launch the middle
  if (kSegSmemKeys == kSegSmallKeys) {
    if (cap > (uint32_t)kSegSmallKeys && cap <= (uint32_t)kSegBigKeys) return;
  } else {
    if (cap <= (uint32_t)kSegSmallKeys || cap > (uint32_t)kSegBigKeys) return;
  }
  uint64_t* gk = G.keys + G.seg_off[(size_t)ii * C + c];
  uint32_t* kl = G.klist + G.seg_off[(size_t)ii * C + c];
  const int tid = threadIdx.x;
  const bool in_smem = cap <= (uint32_t)kSegSmemKeys;
  uint64_t* keys = in_smem ? skeys : gk;
  if (in_smem) {
    for (uint32_t i = tid; i < cap; i += kSegThreads) skeys[i] = (i < n) ? gk[i] : ~0ull;
  } else {
    for (uint32_t i = n + tid; i < cap; i += kSegThreads) gk[i] = ~0ull;
  }
  __syncthreads();
  bitonic_sort_block(keys, (int)cap);

  unsigned long long Mimg = 0;
  for (int q = 0; q < C; ++q) Mimg += G.cls_count[(size_t)ii * C + q];
  const bool trick = use_trick(G.nms_mode, (int64_t)Mimg);
  float off = 0.0f;
  if (trick) off = PQ_MUL((float)c, PQ_ADD(ordered_to_float(G.max_ord[ii]), 1.0f));
  const float4* rbox = G.from_heads ? G.rbox + (size_t)ii * G.N : nullptr;
  const float* bb = G.from_heads ? nullptr : G.bboxes + (size_t)b * G.N * (4 + C);
  auto shifted_row = [&](uint32_t row) -> float4 {            // global gather + coordinate-trick shift
    float4 r;
    if (rbox) r = rbox[row];
    else {
      const float* q = bb + (size_t)row * (4 + C);
      r = make_float4(q[0], q[1], q[2], q[3]);
    }
    return make_float4(PQ_ADD(r.x, off), PQ_ADD(r.y, off), PQ_ADD(r.z, off), PQ_ADD(r.w, off));
  };
  auto shifted_box = [&](uint32_t pos) -> float4 { return shifted_row((uint32_t)keys[pos]); };
  int k;
  if (in_smem) {
    // keys -> registers (and global), barrier, then the boxes take over the region
    constexpr int R = (kSegSmemKeys + kSegThreads - 1) / kSegThreads;
    uint32_t rows[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t i = tid + r * kSegThreads;
      if (i < n) {
        const uint64_t key = skeys[i];
        gk[i] = key;
        rows[r] = (uint32_t)key;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t i = tid + r * kSegThreads;
      if (i < n) sbox[i] = shifted_row(rows[r]);
    }
    __syncthreads();
    // the kept boxes are compacted in place at the front of sbox (kept count <= candidates already consumed, and
    // every warp holds the current step's boxes in registers), so the inner loop reads them without indirection
    k = bucket_greedy<ROUND, kSegWarps>(n, G.iou_f, G.iou_d,
                             [&](uint32_t pos) { return sbox[pos]; },
                             [&](int q) { return sbox[q]; },
                             [&](int q, uint32_t pos, const float4& bx) { skl[q] = (uint16_t)pos; sbox[q] = bx; },
                             s_dead, s_rows, &s_k);
  } else {
    k = bucket_greedy<ROUND, kSegWarps>(n, G.iou_f, G.iou_d, shifted_box,
                             [&](int q) { return shifted_box(kl[q]); },
                             [&](int q, uint32_t pos, const float4&) { kl[q] = pos; }, s_dead, s_rows, &s_k);
  }
  // append kept keys (score desc, row asc, class asc order key) to the image's kept region
  if (tid == 0) {
    s_dst = atomicAdd(&G.kept_count[ii], (uint32_t)k);
    G.run_off[(size_t)ii * C + c] = s_dst;
    G.run_len[(size_t)ii * C + c] = (uint32_t)k;
  }
  __syncthreads();
  uint64_t* kk = G.kkeys + G.img_off[ii] + s_dst;
  for (int q = tid; q < k; q += kSegThreads) {
    const uint64_t key = gk[in_smem ? (uint32_t)skl[q] : kl[q]];
    kk[q] = (key & 0xffffffff00000000ull) | ((uint64_t)((uint32_t)key) << 7) | (uint64_t)c;
  }
}

constexpr size_t kSegBytesPerKey = sizeof(float4) + sizeof(uint16_t);

constexpr int kFinThreads = 1024;
constexpr int kFinSmemKeys = 4096;

__global__ void __launch_bounds__(kFinThreads)
gen_finalize_kernel(const __grid_constant__ GenParams G, const __grid_constant__ DetOut O) {
  if (*G.ok == 0) return;
  __shared__ uint64_t skeys[kFinSmemKeys];
  const int ii = blockIdx.x;
  const int b = gen_image(G, ii);
  const int ob = gen_out(G, ii);
  const int C = G.C;
  const uint32_t K = G.kept_count[ii];
  const int tid = threadIdx.x;
  if (K == 0) {
    if (tid == 0) O.counts[ob] = 0;
    return;
  }
  uint64_t* gk = G.kkeys + G.img_off[ii];
  const float4* rbox0 = G.from_heads ? G.rbox + (size_t)ii * G.N : nullptr;
  const float* bb0 = G.from_heads ? nullptr : G.bboxes + (size_t)b * G.N * (4 + C);
  if (K <= (uint32_t)kFinSmemKeys) {
    // The kept keys of a bucket were appended in score order: the image's region is C sorted runs.  Every key finds
    // its place by counting, run by run (binary search), how many keys of the other classes come before it - no
    // sort, no barrier but the one after staging; the detection is written straight to its rank.
    __shared__ uint32_t s_roff[PQDET_MAX_CLASSES + 2], s_rlen[PQDET_MAX_CLASSES + 2];
    for (int i = tid; i < (int)K; i += kFinThreads) skeys[i] = gk[i];
    for (int q = tid; q < C; q += kFinThreads) {
      s_roff[q] = G.run_off[(size_t)ii * C + q];
      s_rlen[q] = G.run_len[(size_t)ii * C + q];
    }
    __syncthreads();
    for (int j = tid; j < (int)K; j += kFinThreads) {
      const uint64_t mine = skeys[j];
      const int c = (int)(mine & 127u);
      uint32_t rank = (uint32_t)j - s_roff[c];
      for (int q = 0; q < C; ++q) {
        const uint32_t len = s_rlen[q];
        if (q == c || len == 0) continue;
        const uint64_t* run = skeys + s_roff[q];
        uint32_t lo = 0, hi = len;                           // first position whose key is > mine (keys are distinct)
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (run[mid] < mine) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < (uint32_t)O.max_det) {
        const float score = ordered_to_float(~(uint32_t)(mine >> 32));
        const uint32_t row = (uint32_t)mine >> 7;
        float4 bx;
        if (rbox0) bx = rbox0[row];
        else {
          const float* r = bb0 + (size_t)row * (4 + C);
          bx = make_float4(r[0], r[1], r[2], r[3]);
        }
        write_det(O, ob, (int)rank, bx, score, c, row, C);
      }
    }
    if (tid == 0) {
      O.counts[ob] = (int32_t)K;
      if ((int)K > O.max_det) O.status[ob] = PQDET_ST_DET_TRUNCATED;
    }
    return;
  }
  const int Pn = next_pow2((int)K);
  uint64_t* keys = gk;
  if (Pn <= kFinSmemKeys) {
    for (int i = tid; i < Pn; i += kFinThreads) skeys[i] = (i < (int)K) ? gk[i] : ~0ull;
    keys = skeys;
  } else {
    for (int i = (int)K + tid; i < Pn; i += kFinThreads) gk[i] = ~0ull;
  }
  __syncthreads();
  bitonic_sort_block(keys, Pn);
  const int nout = min((int)K, O.max_det);
  const float4* rbox = G.from_heads ? G.rbox + (size_t)ii * G.N : nullptr;
  const float* bb = G.from_heads ? nullptr : G.bboxes + (size_t)b * G.N * (4 + C);
  for (int j = tid; j < nout; j += kFinThreads) {
    const uint64_t k2 = keys[j];
    const float score = ordered_to_float(~(uint32_t)(k2 >> 32));
    const uint32_t low = (uint32_t)k2;
    const uint32_t row = low >> 7;
    const int c = low & 127;
    float4 bx;
    if (rbox) bx = rbox[row];
    else {
      const float* r = bb + (size_t)row * (4 + C);
      bx = make_float4(r[0], r[1], r[2], r[3]);
    }
    write_det(O, ob, j, bx, score, c, row, C);
  }
  if (tid == 0) {
    O.counts[ob] = (int32_t)K;
    if ((int)K > O.max_det) O.status[ob] = PQDET_ST_DET_TRUNCATED;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int fill_heads(const pqdet_heads_t* h, HeadsDev* P) {
  if (!h || h->n_levels < 1 || h->n_levels > PQDET_MAX_LEVELS) return PQDET_ERR_INVALID_ARG;
  if (h->B < 0 || h->A < 1 || h->A > 8 || h->C < 1) return PQDET_ERR_INVALID_ARG;
  if (h->C > PQDET_MAX_CLASSES) return PQDET_ERR_UNSUPPORTED;
  if (h->affine_kind < 0 || h->affine_kind > 2 || !h->orig_hw) return PQDET_ERR_INVALID_ARG;
  if (h->nms_mode < 0 || h->nms_mode > 3 || h->iou_round < 0 || h->iou_round > 1) return PQDET_ERR_INVALID_ARG;
  if (!(h->iou_threshold >= 0.0)) return PQDET_ERR_UNSUPPORTED;  // cross-class pairs have IoU 0: skipping them needs thr >= 0
  int64_t rows = 0, groups = 0;
  for (int l = 0; l < h->n_levels; ++l) {
    if ((!h->raw[l] && h->B > 0) || h->H[l] < 1 || h->W[l] < 1) return PQDET_ERR_INVALID_ARG;   // B == 0: empty tensors have no storage
    LevelDev& L = P->lv[l];
    L.raw = h->raw[l];
    L.H = h->H[l]; L.W = h->W[l]; L.HW = h->H[l] * h->W[l];
    L.stride = h->stride[l];
    L.row_off = (int)rows; L.group_off = (int)groups;
    L.nchunk = (L.HW + 127) / 128;
    L.vec4 = ((L.HW & 3) == 0 && (reinterpret_cast<uintptr_t>(L.raw) & 15) == 0) ? 1 : 0;
    rows += (int64_t)L.HW * h->A;
    groups += L.nchunk;
  }
  for (int l = h->n_levels; l < PQDET_MAX_LEVELS; ++l) P->lv[l] = P->lv[0];
  if (rows >= (1ll << kHitBits)) return PQDET_ERR_UNSUPPORTED;
  if ((int64_t)h->B * h->A * (5 + h->C) >= (1ll << 31)) return PQDET_ERR_UNSUPPORTED;
  if (rows * h->C >= (1ll << 31)) return PQDET_ERR_UNSUPPORTED;   // det_idx = row*C + class is int32
  P->n_levels = h->n_levels;
  P->B = h->B; P->A = h->A; P->C = h->C; P->ch = 5 + h->C;
  P->N = (int)rows; P->G_tot = (int)groups;
  P->kind = h->affine_kind; P->in_h = h->in_h; P->in_w = h->in_w;
  P->orig = h->orig_hw; P->orig_per_image = h->orig_per_image;
  P->thr_f = (float)h->score_threshold;
  P->logit_lo = logit_lo_for(P->thr_f);
  P->iou_f = (float)h->iou_threshold;
  P->iou_d = h->iou_threshold;
  P->nms_mode = h->nms_mode;
  return PQDET_OK;
}

}  // namespace pq

namespace pq {

// Shared launcher of the fused kernel (both sources, both rounding orders, both capacity classes).
static int launch_fused(const HeadsDev& P, const DetOut& O, int32_t* work_counter, int counter_armed,
                        int iou_round, int src, int cap_class, int device, cudaStream_t st) {
  if (cap_class != PQDET_CAP_COMPACT && cap_class != PQDET_CAP_LARGE) return PQDET_ERR_INVALID_ARG;
  if (!(P.thr_f >= 0.0f)) return PQDET_ERR_UNSUPPORTED;     // ~score keys order non-negative floats only
  const int WG = (src == 1) ? 1 : 4 * P.A;
  const size_t lists = cap_class == PQDET_CAP_COMPACT
                           ? sizeof(FusedSmemT<FusedCfg<0>::kCapH, FusedCfg<0>::kCapM>)
                           : sizeof(FusedSmemT<FusedCfg<1>::kCapH, FusedCfg<1>::kCapM>);
  const int threads = cap_class == PQDET_CAP_COMPACT ? FusedCfg<0>::kThreads : FusedCfg<1>::kThreads;
  // hitw + gbase + (heads source) the scan's unit table
  const size_t smem = lists + ((size_t)P.G_tot * (WG + 1) + (src == 0 ? (size_t)P.G_tot * P.A : 0)) * sizeof(uint32_t);
  if (smem > 200 * 1024) return PQDET_ERR_UNSUPPORTED;
  // work_counter = int32[2].  counter_armed != 0: the caller guarantees both words are zero (they are after
  // every completed call: the kernel re-arms them), so no memset is enqueued.
  if (!counter_armed) PQ_CUDA(cudaMemsetAsync(work_counter, 0, 2 * sizeof(int32_t), st));
  auto launch = [&](auto kern, int which) -> int {
    // Launch geometry is a pure function of (device, kernel, smem); remember the last one per device as a
    // single 64-bit word (smem << 32 | grid) so concurrent callers can only ever see a consistent pair.
    static std::atomic<uint64_t> cache[12][16];
    int per_sm_grid = 0;
    if (device < 16) {
      const uint64_t c = cache[which][device].load(std::memory_order_relaxed);
      if ((c >> 32) == (uint64_t)smem) per_sm_grid = (int)(c & 0xffffffffu);
    }
    if (per_sm_grid == 0) {
      PQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 1, sm_count = 148;
      PQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
      if (per_sm < 1) per_sm = 1;
      per_sm_grid = sm_count * per_sm;                      // persistent: one resident wave
      if (device < 16) cache[which][device].store(((uint64_t)smem << 32) | (uint32_t)per_sm_grid, std::memory_order_relaxed);
    }
    int grid = per_sm_grid;
    if (grid > P.B) grid = P.B;
    if (grid == P.B && src == 0 && !getenv("PQDET_FUSED_NO_PDL")) {
      // every CTA owns one image and the kernel reads nothing but the caller's heads before its dependency point: it
      // may overlap the tail of the previous launch on this stream (see the kernel)
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      PQ_CUDA(cudaLaunchKernelEx(&cfg, kern, P, O, work_counter));
    } else {
      kern<<<grid, threads, smem, st>>>(P, O, work_counter);
    }
    PQ_LAUNCH_CHECK();
    return PQDET_OK;
  };
  const int r = iou_round == PQDET_IOU_TV_CUDA ? 0 : 1;
  switch (src == 2 ? 8 + cap_class * 4 + r : cap_class * 4 + src * 2 + r) {
    case 0: return launch(decode_nms_fused_kernel<0, 0, 0>, 0);
    case 1: return launch(decode_nms_fused_kernel<1, 0, 0>, 1);
    case 2: return launch(decode_nms_fused_kernel<0, 1, 0>, 2);
    case 3: return launch(decode_nms_fused_kernel<1, 1, 0>, 3);
    case 4: return launch(decode_nms_fused_kernel<0, 0, 1>, 4);
    case 5: return launch(decode_nms_fused_kernel<1, 0, 1>, 5);
    case 6: return launch(decode_nms_fused_kernel<0, 1, 1>, 6);
    case 7: return launch(decode_nms_fused_kernel<1, 1, 1>, 7);
    case 8: return launch(decode_nms_fused_kernel<0, 2, 0>, 8);      // src 2: cap_class * 4 + 4 + r
    case 9: return launch(decode_nms_fused_kernel<1, 2, 0>, 9);
    case 12: return launch(decode_nms_fused_kernel<0, 2, 1>, 10);
    default: return launch(decode_nms_fused_kernel<1, 2, 1>, 11);
  }
}

}  // namespace pq

#ifdef PQ_PHASE_TIMING
extern "C" int pqdet_debug_phase_cycles(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, pq::pq_phase_cycles, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(pq::pq_phase_cycles, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif

extern "C" int pqdet_decode_nms(const pqdet_heads_t* heads, float* det, int32_t* det_idx, int max_det,
                                int32_t* counts, int32_t* ncand, int32_t* status, int32_t* work_counter,
                                int counter_armed, int capacity_class, int device, void* stream) {
  using namespace pq;
  HeadsDev P;
  memset(&P, 0, sizeof(P));
  int rc = fill_heads(heads, &P);
  if (rc != PQDET_OK) return rc;
  if (!det || !counts || !ncand || !status || !work_counter || max_det < 1) return PQDET_ERR_INVALID_ARG;
  if (P.B == 0) return PQDET_OK;
  PQ_ENTER(device);
  DetOut O{det, det_idx, max_det, counts, ncand, status};
  return launch_fused(P, O, work_counter, counter_armed, heads->iou_round, 0, capacity_class, device, (cudaStream_t)stream);
}

// Host-buffer entry: every pointer may be page-locked host memory; it is translated to its device alias and the
// same kernel runs on it (the loads/stores then travel over PCIe, sector by sector, only where the kernel touches).
namespace pq {
template <typename T>
static int device_alias(T** p, int allow_null) {
  if (!*p) return allow_null ? PQDET_OK : PQDET_ERR_INVALID_ARG;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, (const void*)*p) != cudaSuccess) { cudaGetLastError(); return PQDET_ERR_INVALID_ARG; }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return PQDET_OK;
  if (at.type == cudaMemoryTypeHost && at.devicePointer) { *p = (T*)at.devicePointer; return PQDET_OK; }
  return PQDET_ERR_INVALID_ARG;          // pageable host memory: the device cannot read it
}
}  // namespace pq

namespace pq {
// The receiving side of the arrival counters: thread t waits until source rank t's counter in THIS rank's buffer has
// reached `expected` (wrap-safe), i.e. until that rank's rows and counts of the step have landed here.  Launched as
// a programmatic dependent right behind the fused kernel and triggering at once, so that the NEXT fused launch may
// start behind it while this one still waits; that launch writes nothing before this grid has completed.
__global__ void __launch_bounds__(32)
peer_wait_kernel(const uint32_t* __restrict__ arrived, int n, uint32_t expected, int32_t* __restrict__ err) {
  cudaTriggerProgrammaticLaunchCompletion();
  const int t = threadIdx.x;
  if (t >= n) return;
  const uint32_t* a = arrived + t;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    if ((int32_t)(v - expected) >= 0) break;
    if (spin > (1u << 25)) {                              // ~ a second: a peer is gone; do not hang the device
      if (err) *err = 1 + t;
      break;
    }
    __nanosleep(64);
  }
}
}  // namespace pq

extern "C" int pqdet_peer_wait(const uint32_t* arrived, int n, uint32_t expected, int32_t* err_flag, int device,
                               void* stream) {
  if (!arrived || n < 1 || n > 8) return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(32);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PQ_CUDA(cudaLaunchKernelEx(&cfg, pq::peer_wait_kernel, arrived, n, expected, err_flag));
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}

extern "C" int pqdet_decode_nms_gather(const pqdet_heads_t* heads, float* det, int max_det, int32_t* counts,
                                       int32_t* ncand, int32_t* status, float* const* peer_det,
                                       int32_t* const* peer_counts, uint32_t* const* peer_arrived, int n_peers,
                                       int rank, int gather_cap, int32_t* work_counter, int counter_armed,
                                       int capacity_class, int device, void* stream) {
  using namespace pq;
  HeadsDev P;
  memset(&P, 0, sizeof(P));
  int rc = fill_heads(heads, &P);
  if (rc != PQDET_OK) return rc;
  if (!det || !counts || !ncand || !status || !work_counter || max_det < 1) return PQDET_ERR_INVALID_ARG;
  if (!peer_det || !peer_counts || n_peers < 1 || n_peers > 8 || rank < 0 || rank >= n_peers || gather_cap < 2)
    return PQDET_ERR_INVALID_ARG;
  if (gather_cap & 1) return PQDET_ERR_UNSUPPORTED;       // 16-byte aligned image blocks (coalesced float4 stores)
  if (P.B == 0) return PQDET_OK;
  PQ_ENTER(device);
  DetOut O{det, nullptr, max_det, counts, ncand, status};
  for (int p = 0; p < n_peers; ++p) {
    if (!peer_det[p] || !peer_counts[p]) return PQDET_ERR_INVALID_ARG;
    O.peer_det[p] = peer_det[p];
    O.peer_cnt[p] = peer_counts[p];
    if (peer_arrived && !peer_arrived[p]) return PQDET_ERR_INVALID_ARG;
    O.peer_arr[p] = peer_arrived ? peer_arrived[p] : nullptr;
  }
  O.img_done = work_counter + 2;                            // (int32[3] with peer_arrived)
  O.n_images = P.B;
  O.n_peers = n_peers; O.gather_cap = gather_cap; O.img_off = rank * P.B;
  if (peer_arrived && !counter_armed) PQ_CUDA(cudaMemsetAsync(work_counter + 2, 0, sizeof(int32_t), (cudaStream_t)stream));
  return launch_fused(P, O, work_counter, counter_armed, heads->iou_round, 0, capacity_class, device, (cudaStream_t)stream);
}

extern "C" int pqdet_records_nms(const pqdet_heads_t* heads, const float* rec, const int32_t* rec_count, int rec_cap,
                                 float* det, int32_t* det_idx, int max_det, int32_t* counts, int32_t* ncand,
                                 int32_t* status, int32_t* work_counter, int counter_armed, int capacity_class,
                                 int device, void* stream) {
  using namespace pq;
  if (!heads) return PQDET_ERR_INVALID_ARG;
  pqdet_heads_t h = *heads;
  for (int l = 0; l < PQDET_MAX_LEVELS; ++l) h.raw[l] = rec;     // geometry only: the raw heads were never written
  HeadsDev P;
  memset(&P, 0, sizeof(P));
  int rc = fill_heads(&h, &P);
  if (rc != PQDET_OK) return rc;
  if (!rec || !rec_count || rec_cap < 1) return PQDET_ERR_INVALID_ARG;
  if (!det || !counts || !ncand || !status || !work_counter || max_det < 1) return PQDET_ERR_INVALID_ARG;
  if (P.B == 0) return PQDET_OK;
  for (int l = 0; l < PQDET_MAX_LEVELS; ++l) P.lv[l].raw = nullptr;
  P.rec = rec; P.rec_count = rec_count; P.rec_cap = rec_cap;
  PQ_ENTER(device);
  DetOut O{det, det_idx, max_det, counts, ncand, status};
  return launch_fused(P, O, work_counter, counter_armed, heads->iou_round, 2, capacity_class, device, (cudaStream_t)stream);
}

extern "C" int pqdet_decode_nms_host(const pqdet_heads_t* heads, float* det, int32_t* det_idx, int max_det,
                                     int32_t* counts, int32_t* ncand, int32_t* status, int32_t* work_counter,
                                     int counter_armed, int capacity_class, int device, void* stream) {
  using namespace pq;
  if (!heads) return PQDET_ERR_INVALID_ARG;
  PQ_ENTER(device);
  pqdet_heads_t h = *heads;
  int rc;
  for (int l = 0; l < h.n_levels && l < PQDET_MAX_LEVELS; ++l)
    if ((rc = device_alias(&h.raw[l], 0)) != PQDET_OK) return rc;
  if ((rc = device_alias(&h.orig_hw, 0)) != PQDET_OK) return rc;
  if ((rc = device_alias(&det, 0)) != PQDET_OK) return rc;
  if ((rc = device_alias(&det_idx, 1)) != PQDET_OK) return rc;
  if ((rc = device_alias(&counts, 0)) != PQDET_OK) return rc;
  if ((rc = device_alias(&ncand, 0)) != PQDET_OK) return rc;
  if ((rc = device_alias(&status, 0)) != PQDET_OK) return rc;
  cudaPointerAttributes at;                   // the scheduler words are hammered with atomics: device memory only
  if (!work_counter || cudaPointerGetAttributes(&at, work_counter) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
    cudaGetLastError();
    return PQDET_ERR_INVALID_ARG;
  }
  return pqdet_decode_nms(&h, det, det_idx, max_det, counts, ncand, status, work_counter, counter_armed, capacity_class,
                          device, stream);
}

extern "C" int pqdet_nms_fused(const float* bboxes, int B, int64_t N, int C, double score_threshold,
                               double iou_threshold, int nms_mode, int iou_round, float* det, int32_t* det_idx,
                               int max_det, int32_t* counts, int32_t* ncand, int32_t* status,
                               int32_t* work_counter, int counter_armed, int capacity_class, int device,
                               void* stream) {
  using namespace pq;
  if (!bboxes || !det || !counts || !ncand || !status || !work_counter || max_det < 1) return PQDET_ERR_INVALID_ARG;
  if (B < 0 || N < 0 || C < 1) return PQDET_ERR_INVALID_ARG;
  if (C > PQDET_MAX_CLASSES || N >= (1ll << kHitBits) || N * C >= (1ll << 31)) return PQDET_ERR_UNSUPPORTED;
  if (nms_mode < 0 || nms_mode > 3 || iou_round < 0 || iou_round > 1) return PQDET_ERR_INVALID_ARG;
  if (!(iou_threshold >= 0.0)) return PQDET_ERR_UNSUPPORTED;
  if (B == 0 || N == 0) {
    if (B > 0) {
      PQ_ENTER(device);
      PQ_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * sizeof(int32_t), (cudaStream_t)stream));
      PQ_CUDA(cudaMemsetAsync(ncand, 0, (size_t)B * sizeof(int32_t), (cudaStream_t)stream));
      PQ_CUDA(cudaMemsetAsync(status, 0, (size_t)B * sizeof(int32_t), (cudaStream_t)stream));
    }
    return PQDET_OK;
  }
  HeadsDev P;
  memset(&P, 0, sizeof(P));
  P.n_levels = 0;
  P.B = B; P.A = 1; P.C = C; P.ch = 5 + C;
  P.N = (int)N;
  P.G_tot = (int)((N + 31) / 32);
  P.thr_f = (float)score_threshold;
  P.iou_f = (float)iou_threshold; P.iou_d = iou_threshold;
  P.nms_mode = nms_mode;
  P.bboxes = bboxes; P.bb_row = 4 + C;
  // multiply-high division of the word index by n4 = (4+C)/4 <= 32 is exact below 2^27 words; larger inputs (or
  // rows that are not a whole number of 128-bit words) take the row-per-lane scan
  if (((4 + C) & 3) == 0 && N * ((4 + C) >> 2) < (1ll << 27))
    P.magic_n4 = (uint32_t)(((1ull << 32) + (uint64_t)((4 + C) >> 2) - 1) / (uint64_t)((4 + C) >> 2));
  PQ_ENTER(device);
  DetOut O{det, det_idx, max_det, counts, ncand, status};
  return launch_fused(P, O, work_counter, counter_armed, iou_round, 1, capacity_class, device, (cudaStream_t)stream);
}

namespace pq {
struct GenLayout {
  size_t cls_count, cls_fill, seg_off, seg_cap, max_ord, stage_fill, run_off, run_len, img_off, img_cap, kept_count,
      needed_ok, keys,
      kkeys, klist, rbox, total;
};
static GenLayout gen_layout(int n_images, int64_t N, int C, int64_t cap, int from_heads) {
  GenLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  const size_t nc = (size_t)n_images * C;
  L.cls_count = take(nc * 4);
  L.cls_fill = take(nc * 4);
  L.kept_count = take((size_t)n_images * 4);
  L.max_ord = take((size_t)n_images * 4);
  L.stage_fill = take((size_t)n_images * 4);
  L.run_off = take(nc * 4);
  L.run_len = take(nc * 4);
  L.needed_ok = take(16);
  L.seg_off = take(nc * 8);
  L.seg_cap = take(nc * 4);
  L.img_off = take((size_t)n_images * 8);
  L.img_cap = take((size_t)n_images * 4);
  L.keys = take((size_t)cap * 8);
  L.kkeys = take((size_t)cap * 8);
  L.klist = take((size_t)cap * 4);
  L.rbox = take(from_heads ? (size_t)n_images * (size_t)N * 16 : 0);
  L.total = o;
  return L;
}
}  // namespace pq

extern "C" int64_t pqdet_nms_general_workspace(int n_images, int64_t N, int C, int64_t cand_capacity,
                                               int from_heads) {
  if (n_images < 0 || N < 0 || C < 1 || cand_capacity < 0) return PQDET_ERR_INVALID_ARG;
  return (int64_t)pq::gen_layout(n_images, N, C, cand_capacity, from_heads).total;
}

extern "C" int pqdet_nms_general(const pqdet_heads_t* heads, const float* bboxes, int64_t N, int B, int C,
                                 double score_threshold, double iou_threshold, int nms_mode, int iou_round,
                                 const int32_t* image_ids, int n_images, int out_by_position,
                                 float* det, int32_t* det_idx, int max_det, int32_t* counts, int32_t* ncand,
                                 int32_t* status, void* workspace, int64_t workspace_bytes,
                                 int64_t cand_capacity, int64_t* needed, int device, void* stream) {
  using namespace pq;
  if ((heads != nullptr) == (bboxes != nullptr)) return PQDET_ERR_INVALID_ARG;
  if (!det || !counts || !ncand || !status || !workspace || max_det < 1 || n_images < 0) return PQDET_ERR_INVALID_ARG;
  GenParams G;
  memset(&G, 0, sizeof(G));
  if (heads) {
    int rc = fill_heads(heads, &G.heads);
    if (rc != PQDET_OK) return rc;
    G.from_heads = 1;
    G.N = G.heads.N; G.C = G.heads.C;
    G.thr_f = G.heads.thr_f; G.iou_f = G.heads.iou_f; G.iou_d = G.heads.iou_d; G.nms_mode = G.heads.nms_mode;
    iou_round = heads->iou_round;
    B = heads->B;
  } else {
    if (N < 0 || C < 1 || B < 0) return PQDET_ERR_INVALID_ARG;
    if (C > PQDET_MAX_CLASSES || N >= (1ll << kHitBits)) return PQDET_ERR_UNSUPPORTED;
    if (nms_mode < 0 || nms_mode > 3 || iou_round < 0 || iou_round > 1) return PQDET_ERR_INVALID_ARG;
    if (!(iou_threshold >= 0.0)) return PQDET_ERR_UNSUPPORTED;
    G.from_heads = 0;
    G.bboxes = bboxes; G.N = N; G.C = C;
    G.thr_f = (float)score_threshold; G.iou_f = (float)iou_threshold; G.iou_d = iou_threshold;
    G.nms_mode = nms_mode;
  }
  if (n_images == 0 || G.N == 0) return PQDET_OK;
  if (n_images > 65535) return PQDET_ERR_UNSUPPORTED;
  if (!image_ids && n_images > B) return PQDET_ERR_INVALID_ARG;
  const GenLayout L = gen_layout(n_images, G.N, G.C, cand_capacity, G.from_heads);
  if ((int64_t)L.total > workspace_bytes) return PQDET_ERR_WORKSPACE;
  PQ_ENTER(device);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  G.image_ids = image_ids; G.n_images = n_images; G.out_by_pos = out_by_position ? 1 : 0;
  G.cls_count = (uint32_t*)(ws + L.cls_count);
  G.cls_fill = (uint32_t*)(ws + L.cls_fill);
  G.kept_count = (uint32_t*)(ws + L.kept_count);
  G.max_ord = (uint32_t*)(ws + L.max_ord);
  G.stage_fill = (uint32_t*)(ws + L.stage_fill);
  G.run_off = (uint32_t*)(ws + L.run_off);
  G.run_len = (uint32_t*)(ws + L.run_len);
  G.needed = needed ? needed : (int64_t*)(ws + L.needed_ok);
  G.ok = (int32_t*)(ws + L.needed_ok + 8);
  G.seg_off = (int64_t*)(ws + L.seg_off);
  G.seg_cap = (uint32_t*)(ws + L.seg_cap);
  G.img_off = (int64_t*)(ws + L.img_off);
  G.img_cap = (uint32_t*)(ws + L.img_cap);
  G.keys = (uint64_t*)(ws + L.keys);
  G.kkeys = (uint64_t*)(ws + L.kkeys);
  G.klist = (uint32_t*)(ws + L.klist);
  G.rbox = (float4*)(ws + L.rbox);
  G.cand_capacity = cand_capacity;
  DetOut O{det, det_idx, max_det, counts, ncand, status};
  // the counters (and the run table) are the first contiguous regions of the layout
  PQ_CUDA(cudaMemsetAsync(ws, 0, L.seg_off, st));
  dim3 sel_grid((unsigned)((G.N + 255) / 256), n_images);
  // PQDET_GEN_SELECT=twopass: count, plan, re-evaluate + scatter (round 1); default: one read of the heads
  const char* ssel = getenv("PQDET_GEN_SELECT");
  const bool single = !(ssel && ssel[0] == 't') && cand_capacity / n_images >= 256;
  G.stage_quota = single ? cand_capacity / n_images : 0;
  if (single) {
    gen_select_kernel<2><<<sel_grid, 256, 0, st>>>(G);
    PQ_LAUNCH_CHECK();
    gen_plan_kernel<<<1, 1024, 0, st>>>(G, O);
    PQ_LAUNCH_CHECK();
    dim3 bgrid((unsigned)((G.stage_quota + 255) / 256), n_images);
    gen_bucketize_kernel<<<bgrid, 256, 0, st>>>(G);
    PQ_LAUNCH_CHECK();
  } else {
    gen_select_kernel<0><<<sel_grid, 256, 0, st>>>(G);
    PQ_LAUNCH_CHECK();
    gen_plan_kernel<<<1, 1024, 0, st>>>(G, O);
    PQ_LAUNCH_CHECK();
    gen_select_kernel<1><<<sel_grid, 256, 0, st>>>(G);
    PQ_LAUNCH_CHECK();
  }
  dim3 seg_grid(G.C, n_images);
  auto launch_buckets = [&](auto small, auto big) -> int {
    const size_t sb = kSegBytesPerKey * kSegSmallKeys, bb = kSegBytesPerKey * kSegBigKeys;
    PQ_CUDA(cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    PQ_CUDA(cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bb));
    small<<<seg_grid, kSegSmallThreads, sb, st>>>(G);
    PQ_LAUNCH_CHECK();
    big<<<seg_grid, kSegBigThreads, bb, st>>>(G);
    PQ_LAUNCH_CHECK();
    return PQDET_OK;
  };
  int brc;
  if (iou_round == PQDET_IOU_TV_CUDA)
    brc = launch_buckets(gen_bucket_nms_kernel<0, kSegSmallKeys, kSegSmallThreads>,
                         gen_bucket_nms_kernel<0, kSegBigKeys, kSegBigThreads>);
  else
    brc = launch_buckets(gen_bucket_nms_kernel<1, kSegSmallKeys, kSegSmallThreads>,
                         gen_bucket_nms_kernel<1, kSegBigKeys, kSegBigThreads>);
  if (brc != PQDET_OK) return brc;
  gen_finalize_kernel<<<n_images, kFinThreads, 0, st>>>(G, O);
  PQ_LAUNCH_CHECK();
  return PQDET_OK;
}
