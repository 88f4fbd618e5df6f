/* pqdet_b200 -- C ABI of the B200-native PQDet detection hot path.
 *
 * The reference (eleflea/PQDet) has no FFI: its "operator interface" is plain Python
 * (SURVEY.md section 8b).  This header is therefore the NEW boundary a maintainer binds with
 * ctypes; each entry point names the reference code it replaces.
 *
 * Conventions
 *   - extern "C", every function returns int: 0 = PQDET_OK, < 0 = error (pqdet_strerror()).
 *   - No allocation inside: the caller owns all buffers (device pointers unless said otherwise)
 *     and passes the device ordinal and the cudaStream_t (as void*) explicitly.  The library is
 *     re-entrant and keeps no global mutable state (nn.DataParallel calls the reference's
 *     YOLOLayer from one thread per GPU, tools.py:215-216).
 *   - All tensors fp32, contiguous.  Raw heads are NCHW (B, A*(5+C), H, W), channel = a*(5+C)+k,
 *     exactly what the 1x1 head conv emits (model/parser.py:206-215).
 *   - Rows of the concatenated prediction are ordered level-major in the order the levels are
 *     given (cfg order), then (y*W + x)*A + a  (model/interpreter.py:75-76).
 *   - Launches are asynchronous on `stream`; outputs are valid after the stream is synchronised.
 */
#ifndef PQDET_B200_H_
#define PQDET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQDET_VERSION 111          /* 0.1.1: capacity_class on the fused entry points; 111: arrival counters */
#define PQDET_MAX_LEVELS 4
#define PQDET_MAX_CLASSES 126      /* class id is a 7-bit field of the 64-bit sort keys (class:7 | ~score:32 | row:25);
                                      more classes (or >= 2^25 rows per image) return PQDET_ERR_UNSUPPORTED */

/* On-chip list capacities of the fused kernels, per image (capacity_class argument):
 *   PQDET_CAP_COMPACT  512 rows above the objectness threshold, 1280 candidates; 128-thread CTAs, 7 per SM -
 *                      the right geometry for eval batches of natural images (VOC / COCO-like: ~130 / ~600);
 *   PQDET_CAP_LARGE    1024 rows, 2048 candidates; 256-thread CTAs, 4 per SM.
 * An image that exceeds the lists is flagged PQDET_ST_CAND_OVERFLOW (counts = 0) and must be re-run with the larger
 * class or through pqdet_nms_general; the results of both classes are identical wherever both fit. */
enum { PQDET_CAP_COMPACT = 0, PQDET_CAP_LARGE = 1 };

enum {
  PQDET_OK = 0,
  PQDET_ERR_INVALID_ARG = -1,
  PQDET_ERR_CUDA = -2,
  PQDET_ERR_UNSUPPORTED = -3,
  PQDET_ERR_WORKSPACE = -4
};

/* dataset/__init__.py:17-21 RECOVER_BBOXES_REGISTER keys */
enum { PQDET_AFFINE_VOC = 0, PQDET_AFFINE_COCO = 1, PQDET_AFFINE_VISDRONE = 2 };

/* torchvision/ops/boxes.py:80-83 dispatch.  AUTO_CUDA: coordinate trick iff 4*M <= 100000,
 * AUTO_CPU: iff 4*M <= 4000, else the per-class ("vanilla") arithmetic. */
enum { PQDET_NMS_AUTO_CUDA = 0, PQDET_NMS_AUTO_CPU = 1, PQDET_NMS_TRICK = 2, PQDET_NMS_VANILLA = 3 };

/* IoU rounding order of torchvision's nms kernels:
 * TV_CUDA: D = fl(fma(bw,bh,Sa) - I), fl(I/D) > (float)thr;  TV_CPU: D = fl(fl(Sa+fl(bw*bh)) - I),
 * (double)fl(I/D) > thr. */
enum { PQDET_IOU_TV_CUDA = 0, PQDET_IOU_TV_CPU = 1 };

/* model/parser.py:94-100 'bbox_loss' values */
enum { PQDET_BBOX_L1 = 0, PQDET_BBOX_IOU = 1, PQDET_BBOX_GIOU = 2, PQDET_BBOX_DIOU = 3 };

/* per-image status bits written by the decode+NMS entry points */
enum {
  PQDET_ST_OK = 0,
  PQDET_ST_CAND_OVERFLOW = 1,   /* more hit rows / candidates than the fused kernel stages on chip:
                                   nothing was written for this image; run pqdet_nms_general on it */
  PQDET_ST_DET_TRUNCATED = 2    /* kept > max_det: counts[] holds the true K, only max_det rows written */
};

int pqdet_version(void);
const char* pqdet_strerror(int code);

/* ---- a2..a4: model/parser.py:206-235 Decode.forward (+ the eval concat of
 * model/interpreter.py:75-76 when out_rows_total/out_row_offset address a (B, N, 5+C) buffer).
 * raw (B, A*(5+C), H, W) -> out rows [out_row_offset, out_row_offset + H*W*A) of every image. */
int pqdet_decode_fwd(const float* raw, float* out, int B, int A, int C, int H, int W, float stride,
                     int64_t out_rows_total, int64_t out_row_offset, int device, void* stream);

/* DetectionModel.forward eval branch (model/interpreter.py:72-76): every level's Decode + view + cat in ONE launch.
 * raw/H/W/stride: HOST arrays of n_levels entries in cfg order; out (B, N, 5+C), N = sum of H*W*A. */
int pqdet_decode_levels(int n_levels, const float* const* raw, const int* H, const int* W, const float* stride,
                        float* out, int B, int A, int C, int device, void* stream);

/* autograd of Decode.forward: grad_raw = grad_out * d out / d raw (needs raw, not out). */
int pqdet_decode_bwd(const float* raw, const float* grad_out, float* grad_raw, int B, int A, int C,
                     int H, int W, float stride, int64_t out_rows_total, int64_t out_row_offset,
                     int device, void* stream);

/* ---- a5: dataset/base_sample.py:98-139 recover_bboxes_prediction + the three affine functions
 * (voc_sample.py:92-95, coco_sample.py:97-100, visdrone_sample.py:84-88).
 * pred (B, N, 5+C) -> out (B, N, 4+C).  orig_hw: device (B,2) [orig_per_image=1] or (2,) [=0], (h,w).
 * Unlike the reference the input is NOT modified in place. */
int pqdet_recover(const float* pred, float* out, int B, int64_t N, int C, int affine_kind,
                  float in_h, float in_w, const float* orig_hw, int orig_per_image,
                  int device, void* stream);

/* ---- fused a2+a4+a5+a6: three raw heads -> detections, one launch, no intermediate tensor.
 * Replaces Decode x L -> cat -> recover_bboxes_prediction_* -> per-image tools.torch_nms
 * (predict.py:33-45, eval/evaluator.py:48-59).
 * det    (B, max_det, 6) rows [x1,y1,x2,y2,score,class], descending score, ties by (row, class)
 * det_idx(B, max_det)    optional (may be NULL): row*C + class of every detection
 * counts (B)             kept detections K per image
 * ncand  (B)             candidates M per image (score > thr)
 * status (B)             PQDET_ST_* bits
 * work_counter           int32[2] of scratch (dynamic image scheduler).  counter_armed = 0: the call zeroes it
 *                        first; = 1: the caller guarantees both words are 0 - true after every completed
 *                        call on the same stream, because the last CTA to leave re-arms them - so a
 *                        steady-state loop enqueues nothing but the kernel
 * capacity_class         PQDET_CAP_* (above)
 * Stream semantics: ordinary stream order for everything the caller can observe.  Internally, when every CTA owns one
 * image (B <= one resident wave) the kernel is launched as a programmatic dependent of whatever precedes it on the
 * stream and may START while the previous pqdet_decode_nms launch is draining; it reads only `heads` (and orig_hw)
 * until that launch has completed, and writes its outputs afterwards - the same output buffers may be reused from
 * call to call.  Kernels of other libraries in front of it are waited for as usual.  (PQDET_FUSED_NO_PDL=1 in the
 * environment disables the overlap.)
 * heads->score_threshold must be >= 0 (scores are products of sigmoids; the keys order non-negative floats):
 *                        negative thresholds return PQDET_ERR_UNSUPPORTED - use pqdet_nms_general */
typedef struct {
  const float* raw[PQDET_MAX_LEVELS];
  int H[PQDET_MAX_LEVELS], W[PQDET_MAX_LEVELS];
  float stride[PQDET_MAX_LEVELS];
  int n_levels;
  int B, A, C;
  int affine_kind;
  float in_h, in_w;
  const float* orig_hw;
  int orig_per_image;
  double score_threshold;     /* compared in fp32, like `tensor > python_float` (tools.py:551) */
  double iou_threshold;
  int nms_mode, iou_round;
} pqdet_heads_t;

int pqdet_decode_nms(const pqdet_heads_t* heads, float* det, int32_t* det_idx, int max_det,
                     int32_t* counts, int32_t* ncand, int32_t* status, int32_t* work_counter,
                     int counter_armed, int capacity_class, int device, void* stream);

/* ---- multi-GPU eval without a collective (the gather of eval detections,
 * eval/evaluator.py:49-61 under nn.DataParallel): pqdet_decode_nms whose output stage ALSO stores every kept row and
 * every count into the gathered buffers of all ranks of the node, through peer memory mapped over NVLink / NVSwitch
 * (e.g. torch.distributed._symmetric_memory: peer_det[p] / peer_counts[p] = rank p's buffer as mapped into this
 * process; p = rank is the local one).  peer_det[p] (world, B, gather_cap, 6), peer_counts[p] (world, B): this call
 * fills block `rank` of every buffer (counts clamped to gather_cap, rows beyond it dropped; gather_cap must be
 * even: an image's rows are staged on chip and leave as coalesced 16-byte stores).  The buffers are complete
 * on a rank once every rank's call has finished - a device-side barrier (symmetric-memory signal pads) after the
 * launch, no data-path collective.  det / counts / ncand / status are the usual local outputs.
 * peer_arrived (nullable): peer_arrived[p] = the address, in rank p's buffer, of the 32-bit arrival counter that
 * belongs to THIS rank (zero at start, never reset; work_counter is then int32[3]).  The launch then adds its image
 * count to it on every peer once all its rows and counts are stored there (release, system scope; one remote add per
 * peer and launch), and pqdet_peer_wait on the receiving side replaces the
 * barrier: arrived[0..n) = this rank's counters (one per source rank), expected = images per rank x calls so far
 * (modulo 2^32).  The wait is a one-warp kernel launched as a programmatic dependent, so consecutive
 * pqdet_decode_nms_gather calls still overlap; *err_flag (nullable, device) is set to 1 + source rank if a counter
 * does not arrive within about a second. */
int pqdet_decode_nms_gather(const pqdet_heads_t* heads, float* det, int max_det, int32_t* counts, int32_t* ncand,
                            int32_t* status, float* const* peer_det, int32_t* const* peer_counts,
                            uint32_t* const* peer_arrived, int n_peers, int rank, int gather_cap,
                            int32_t* work_counter, int counter_armed, int capacity_class, int device, void* stream);
int pqdet_peer_wait(const uint32_t* arrived, int n, uint32_t expected, int32_t* err_flag, int device, void* stream);

/* The training path's per-step loss exchange over the same peer memory (model/loss.py:105-108 + trainer.py:233:
 * the mean of the replicas' losses): pqdet_peer_publish stores src[0..n) * scale into row `rank` of every rank's
 * symmetric buffer (peer_bufs[p] = rank p's (world, row_stride) buffer as mapped here); after a device-side barrier -
 * or, with peer_arrived (nullable; as for pqdet_decode_nms_gather: +1 per call on this rank's arrival counter in every
 * rank's buffer) after pqdet_peer_wait(arrived, n_peers, calls so far) - pqdet_peer_sum_rows adds the rows in rank
 * order into out[0..n).  n <= 1024. */
int pqdet_peer_publish(const float* src, int n, float scale, float* const* peer_bufs, uint32_t* const* peer_arrived,
                       int n_peers, int rank, int row_stride, int device, void* stream);
int pqdet_peer_sum_rows(const float* rows, int n, int n_rows, int row_stride, float* out, int device, void* stream);

/* ---- 8f-2 (second half): the eval path starting at the INPUT of the head convolutions, with nothing of size
 * B x N ever written (model/cfg/regnetx-600m-fpn.cfg:646-651 + model/parser.py:206-235 + tools.py:540-566).
 * pqdet_head_conv_hits  one level: 1x1 head convolution on the tensor cores (tcgen05, TF32 like PyTorch's default);
 *                       the epilogue reads back only the objectness column of every anchor and appends one record
 *                       [row, objectness, 4 box, C class raw values] (6 + C floats; row = its index in the
 *                       concatenated prediction, stored as int bits) per row whose objectness can exceed
 *                       score_threshold.  rec (B, rec_cap, 6 + C); rec_count (B) int32, zeroed by the caller before
 *                       the first level, counts appended records of all levels (> rec_cap: the image overflowed).
 *                       Shapes the persistent kernel cannot take (plane stride not a multiple of 16 bytes, weights
 *                       beyond shared memory) run the general tcgen05 kernel with the same epilogue; more than 256
 *                       output channels return PQDET_ERR_UNSUPPORTED (pqdet_head_conv_decode(out_raw) + pqdet_decode_nms).
 * pqdet_records_nms     the fused kernel's back end on those records: exact conf > thr test, decode + recover of the
 *                       boxes, scores, class-aware NMS, output - same outputs / status / scheduler-word contract as
 *                       pqdet_decode_nms; heads supplies the geometry (raw[] is ignored).  Results are identical to
 *                       pqdet_decode_nms on the raw heads this convolution would have written. */
int pqdet_head_conv_hits(const float* x, const float* weight, const float* bias, int B, int Cin, int H, int W,
                         int A, int C, double score_threshold, int64_t row_offset, float* rec, int32_t* rec_count,
                         int rec_cap, int device, void* stream);
int pqdet_records_nms(const pqdet_heads_t* heads, const float* rec, const int32_t* rec_count, int rec_cap,
                      float* det, int32_t* det_idx, int max_det, int32_t* counts, int32_t* ncand, int32_t* status,
                      int32_t* work_counter, int counter_armed, int capacity_class, int device, void* stream);

/* Host-buffer form of pqdet_decode_nms: the call predict.py:33-45 / eval/evaluator.py:48-61 would make when the
 * head outputs and the detections live in HOST memory.  heads->raw[], heads->orig_hw, det, det_idx, counts, ncand
 * and status may each be PAGE-LOCKED host memory (cudaHostAlloc / cudaHostRegister / torch pin_memory) or device
 * memory; pageable host memory is rejected (PQDET_ERR_INVALID_ARG).  No staging copy is made: the kernel reads the
 * objectness planes and the channels of the rows above threshold straight over PCIe (about an eighth of the head
 * bytes on VOC-like inputs) and writes the detection rows straight back.  work_counter must be device memory.
 * Outputs are valid after `stream` is synchronised. */
int pqdet_decode_nms_host(const pqdet_heads_t* heads, float* det, int32_t* det_idx, int max_det,
                          int32_t* counts, int32_t* ncand, int32_t* status, int32_t* work_counter,
                          int counter_armed, int capacity_class, int device, void* stream);

/* ---- a6: tools.torch_nms (tools.py:540-566) for a whole batch in ONE launch: bboxes (B, N, 4+C) recovered
 * boxes + class scores -> detections, same outputs / status bits / scheduler-word contract as
 * pqdet_decode_nms.  Images that exceed the capacity class' lists are flagged PQDET_ST_CAND_OVERFLOW and must be
 * re-run through pqdet_nms_general(bboxes=...).  score_threshold must be >= 0 (else PQDET_ERR_UNSUPPORTED). */
int pqdet_nms_fused(const float* bboxes, int B, int64_t N, int C, double score_threshold,
                    double iou_threshold, int nms_mode, int iou_round, float* det, int32_t* det_idx,
                    int max_det, int32_t* counts, int32_t* ncand, int32_t* status,
                    int32_t* work_counter, int counter_armed, int capacity_class, int device, void* stream);

/* General (any candidate count) path, same results as pqdet_decode_nms.  Works on the images
 * listed in image_ids (device int32[n_images]; NULL = images 0..n_images-1).  det/det_idx/counts/
 * ncand/status are indexed by image id, or by position in image_ids when out_by_position != 0.
 * Exactly one source must be given:
 *   heads != NULL : candidates come from the raw heads (fallback of the fused kernel), or
 *   bboxes != NULL: candidates come from a recovered tensor (B, N, 4+C) -- this is the batched
 *                   drop-in for tools.torch_nms (tools.py:540-566).
 * cand_capacity: number of 8-byte candidate slots in the workspace the caller sized with
 * pqdet_nms_general_workspace(); if the images need more, PQDET_ST_CAND_OVERFLOW is set for all
 * of them, *needed (device int64) holds the required capacity and nothing else is written. */
int64_t pqdet_nms_general_workspace(int n_images, int64_t N, int C, int64_t cand_capacity, int from_heads);

int pqdet_nms_general(const pqdet_heads_t* heads, const float* bboxes, int64_t N, int B, int C,
                      double score_threshold, double iou_threshold, int nms_mode, int iou_round,
                      const int32_t* image_ids, int n_images, int out_by_position,
                      float* det, int32_t* det_idx, int max_det, int32_t* counts, int32_t* ncand,
                      int32_t* status, void* workspace, int64_t workspace_bytes, int64_t cand_capacity,
                      int64_t* needed, int device, void* stream);

/* ---- a13: tools.nms (tools.py:507-538; no callers in the reference): per-class selection loop with hard NMS or the
 * soft-NMS decay exp(-iou^2 / sigma), the reference's semantics in full (first pick of a class is not thresholded,
 * iou_calc1's clamped union, scores decay in place).  The n boxes are grouped by class by the caller: class c owns
 * rows [seg_off[c], seg_off[c+1]) of boxes (n,4) / scores (n) (scores is overwritten with the decayed values).
 * out_idx / out_score (n): for class c the picks in order at [seg_off[c], seg_off[c] + out_count[c]).
 * alive_scratch: n bytes. */
int pqdet_classwise_nms(const float* boxes, float* scores, const int32_t* seg_off, int n_classes, int64_t n,
                        int soft, double sigma, double score_threshold, double iou_threshold,
                        int32_t* out_idx, float* out_score, int32_t* out_count, uint8_t* alive_scratch,
                        int device, void* stream);

/* ---- a7/a8: tools.py:357-437 iou_calc3 / giou / diou / ciou, elementwise over n box pairs
 * (broadcasting is done by the caller).  kind: 0 iou, 1 giou, 2 diou, 3 ciou, 4 iou with the union clamped at 1e-14
 * (tools.iou_calc1, tools.py:335-355; forward only).
 * grad_b1/grad_b2 (may be NULL): d out / d boxes times grad_out (kinds 0..3; ciou's alpha is a constant of the
 * differentiation, as under the reference's torch.no_grad()). */
int pqdet_iou_pairwise(const float* b1, const float* b2, float* out, int64_t n, int kind,
                       int device, void* stream);
int pqdet_iou_pairwise_bwd(const float* b1, const float* b2, const float* grad_out, float* grad_b1,
                           float* grad_b2, int64_t n, int kind, int device, void* stream);

/* ---- a3+a9+a10: model/parser.py:244-249 YOLOLayer.forward(x, target) = decode + loss_per_scale
 * (model/loss.py:22-115), forward and backward in one pass.
 * input_is_raw=1: x is the raw head (B, A*(5+C), H, W); grad (same shape, may be NULL) = d loss/d raw.
 * input_is_raw=0: x is a decoded pred (B,H,W,A,5+C) (standalone loss_per_scale); grad = d loss/d pred.
 * label (B,H,W,A,6+C); gt (B,G,4) zero padded.
 * out4 (device float[4]) = loss, bbox_loss, conf_loss, cls_loss (each the batch mean, (1,) in the
 * reference); nan_flag (device int32) = 1 iff loss is NaN (the reference raises, loss.py:110-114).
 * partials: device scratch of pqdet_loss_workspace() bytes.  The gradient assumes an upstream
 * gradient of 1 for `loss`; the box / objectness / class channels carry d bbox_loss, d conf_loss,
 * d cls_loss respectively, so any other upstream mix is a per-channel-group rescale. */
int64_t pqdet_loss_workspace(int B, int A, int H, int W);
int pqdet_loss_fwd_bwd(const float* x, int input_is_raw, const float* label, const float* gt,
                       float* grad, float* out4, int32_t* nan_flag, void* partials,
                       int B, int A, int C, int H, int W, int G, float stride, int bbox_loss,
                       float ignore_thresh, float l1_loss_gain, int device, void* stream);

/* Chain rule for arbitrary upstream gradients of the four outputs of pqdet_loss_fwd_bwd: scales the
 * box / objectness / class channel groups of grad IN PLACE by (g_loss+g_bbox), (g_loss+g_conf),
 * (g_loss+g_cls).  g_* are DEVICE scalars (NULL = 0); when all three factors are 1 the kernel
 * returns without touching memory, so the common loss.backward() costs no extra pass and no sync. */
int pqdet_loss_scale_grad(float* grad, int input_is_raw, int B, int A, int C, int H, int W,
                          const float* g_loss, const float* g_bbox, const float* g_conf,
                          const float* g_cls, int device, void* stream);

/* ---- a4 (training branch): all FPN levels of DetectionModel.forward(x, target)
 * (model/interpreter.py:51-54, 77-85) in ONE launch: per-level decode + loss_per_scale forward+backward,
 * the Python-order sums over levels and loss_per_branch.  Host arrays of n_levels device pointers / ints.
 * out (device float[4 + 5*n_levels]):
 *   [0..3]          loss, bbox(giou_loss), conf_loss, class_loss summed over levels, ((0+h0)+h1)+h2
 *   [4+4l..7+4l]    the four per-level losses YOLOLayer returns
 *   [4+4L+l]        loss_per_branch[l] = bbox+conf+cls of level l
 * workspace: pqdet_loss_levels_workspace() bytes of scratch (per-CTA partial sums, reduced in a fixed order by a
 * second single-CTA launch); workspace_initialised is accepted for ABI stability and ignored (no state is kept). */
int64_t pqdet_loss_levels_workspace(int n_levels, int B, int A, const int* H, const int* W);
int pqdet_loss_levels(int n_levels, const float* const* raw, const float* const* label,
                      const float* const* gt, float* const* grad, const int* H, const int* W,
                      const int* G, const float* stride, int B, int A, int C, int bbox_loss,
                      float ignore_thresh, float l1_loss_gain, float* out, int32_t* nan_flag,
                      void* workspace, int workspace_initialised, int device, void* stream);
/* Chain rule for pqdet_loss_levels: upstream = d L / d out (device float[4 + 5*n_levels]); scales every
 * level's grad in place; returns without memory traffic when all factors are 1 (loss.backward()). */
int pqdet_loss_levels_scale_grad(int n_levels, float* const* grad, const int* H, const int* W,
                                 int B, int A, int C, const float* upstream, int device, void* stream);

/* ---- a11/a12: dataset/train_dataset.py:109-150 create_label + collate_batch (:16-43) for a whole
 * batch.  gt (B, n_max, 6) rows [x1,y1,x2,y2,class,mixw], gt_count (B).
 * anchors: HOST float[9*2] (w,h); strides/H/W: HOST int[3], ascending strides (8,16,32).  For scale s:
 *   label[s]   (B, H_s, W_s, 3, 6+C)   fully written (background = 0, mixw channel = 1)
 *   gtlist[s]  (B, list_capacity, 4)   zero padded; list_len (B,3) true lengths (with duplicates)
 * owner: device int32 scratch of pqdet_assign_workspace() bytes. */
int64_t pqdet_assign_workspace(int B, const int* H, const int* W);
int pqdet_assign_labels(const float* gt, const int32_t* gt_count, int B, int n_max, int C,
                        const float* anchors, const int* strides, const int* H, const int* W,
                        float iou_threshold, float* label0, float* label1, float* label2,
                        float* gtlist0, float* gtlist1, float* gtlist2, int list_capacity,
                        int32_t* list_len, void* owner, int device, void* stream);

/* ---- sparse targets (keep GT as (B,G,6) lists end to end instead of the dense
 * (B,H,W,3,6+C) tensors of dataset/train_dataset.py:26-43).  pqdet_assign_sparse runs the assignment of
 * create_label (train_dataset.py:109-150) but emits, per scale, only
 *   owner[s]  (B, 3, H_s*W_s) int32: index (into gt) of the GT whose row create_label would have left in that
 *             label slot ("last writer wins", :145), -1 for background
 *   gtlist[s] (B, list_capacity, 4) + list_len (B,3), exactly as pqdet_assign_labels.
 * pqdet_loss_levels_sparse = pqdet_loss_levels with the label rows rebuilt on the fly from owner + gt
 * ([gt box, 1, smoothed one-hot(class), mixw] / background [0.., mixw = 1]): same arithmetic, bit-identical
 * outputs, without ever writing or reading the L bytes of dense labels.  A must be 3. */
int pqdet_assign_sparse(const float* gt, const int32_t* gt_count, int B, int n_max,
                        const float* anchors, const int* strides, const int* H, const int* W,
                        float iou_threshold, int32_t* owner0, int32_t* owner1, int32_t* owner2,
                        float* gtlist0, float* gtlist1, float* gtlist2, int list_capacity,
                        int32_t* list_len, int device, void* stream);
int pqdet_loss_levels_sparse(int n_levels, const float* const* raw, const int32_t* const* owner,
                             const float* gt6, int n_max, const float* const* gtlist,
                             float* const* grad, const int* H, const int* W, const int* G,
                             const float* stride, int B, int A, int C, int bbox_loss,
                             float ignore_thresh, float l1_loss_gain, float* out, int32_t* nan_flag,
                             void* workspace, int workspace_initialised, int device, void* stream);

/* ---- evaluator statistics: the per-detection matching loop of Evaluator.AP,
 * eval/evaluator.py:69-124, for all classes, images and IoU thresholds in one launch.  The caller (see
 * pqdet_b200/evaluator.py) orders the detections per class by (-score, insertion index) exactly like
 * tools.PriorityQueue (tools.py:654-679) and groups the labels per (image, class) like add_labels (:164-175):
 *   det_box      (D,4)     detections, class-sorted order
 *   grp_det      (..)      detection indices ordered by (group, class rank); grp_det_off (G+1)
 *   gt_box       (sumGT,4) float32 or float64 (gt_is_f64), groups back to back; gt_difficult (sumGT); gt_off (G+1)
 *   thresholds   (T)       DEVICE doubles (np.linspace(0.5, 0.95, 10), :13)
 *   seen         (T,sumGT) scratch, tp / fp (T,D): all three zero on entry; tp/fp receive the reference's flags
 * Overlaps follow numpy's promotion (float32 GT: float32 arithmetic; float64 GT: only the detection's own area
 * stays float32), so tp/fp - and the AP computed from them - are identical to the reference's. */
int pqdet_ap_match(const float* det_box, int64_t D, const int32_t* grp_det, const int64_t* grp_det_off,
                   const void* gt_box, int gt_is_f64, const uint8_t* gt_difficult, const int64_t* gt_off,
                   int64_t sum_gt, int G, const double* thresholds, int T, uint8_t* seen, uint8_t* tp,
                   uint8_t* fp, int device, void* stream);

/* ---- eval pre-processing: augment.Resize (dataset/augment.py:227-259: cv2.resize
 * INTER_LINEAR to the letterbox size + constant padding), augment.Normalize (:206-215) and augment.ToTensor
 * (:390-398), i.e. eval_augment_voc / eval_augment_coco (dataset/voc_sample.py:85-90), for a batch of uint8 HWC images
 * of different sizes in one launch.
 *   src     packed device buffer holding the images back to back (3 channels, HWC)
 *   images  device array of B records {int64 src_off; int32 sh, sw, dh, dw, du, dl; double scale_y, scale_x}
 *           (48 bytes each, naturally aligned) filled by the caller exactly as Resize.__call__ / cv::resize derive them
 *   mean3/std3  HOST float[3]
 *   out_chw     (B, 3, target_h, target_w) float32, normalised (may be NULL)
 *   out_hwc_u8  (B, target_h, target_w, 3) uint8, the padded resized image (may be NULL)
 * The resize reproduces OpenCV's 8-bit fixed-point bilinear arithmetic: out_hwc_u8 is bit-identical to
 * np.pad(cv2.resize(...)), out_chw to the reference chain. */
int pqdet_letterbox_normalize(const uint8_t* src, const void* images, int B, int target_h, int target_w,
                              int pad_val, const float* mean3, const float* std3, float* out_chw,
                              uint8_t* out_hwc_u8, int device, void* stream);

/* ---- head 1x1 convolution + decode: the `filters = A*(5+C), size = 1, linear`
 * convolution in front of every [yolo] layer (model/cfg/regnetx-600m-fpn.cfg:646-651) fused with Decode.forward
 * (model/parser.py:206-235) on the tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM).
 *   x (B, Cin, H, W), weight (A*(5+C), Cin) [the conv's (O, Cin, 1, 1) weight], bias (A*(5+C)) or NULL
 *   out_decoded: rows [out_row_offset, +H*W*A) of a (B, out_rows_total, 5+C) prediction (may be NULL)
 *   out_raw    : (B, A*(5+C), H, W), what the convolution alone would produce (may be NULL)
 * TF32 products (fp32 operands truncated to 10 mantissa bits), fp32 accumulation - the precision PyTorch's own
 * convolution runs at by default; |error| <= 2^-9 * sum_c |x_c * w_c|. */
int pqdet_head_conv_decode(const float* x, const float* weight, const float* bias, float* out_decoded,
                           float* out_raw, int B, int Cin, int H, int W, int A, int C, float stride,
                           int64_t out_rows_total, int64_t out_row_offset, int device, void* stream);

/* All levels of the head in ONE launch = the convolutions in front of the [yolo] layers + the eval branch of
 * DetectionModel.forward (model/interpreter.py:72-76): level l's rows follow level l-1's in out_decoded
 * (B, sum_l H_l*W_l*A, 5+C).  x[l] (B, Cin[l], H[l], W[l]), weight[l] (A*(5+C), Cin[l]), bias (array of pointers, NULL or
 * with NULL entries = no bias).  The SMs are split between the levels in proportion to their cost.  Returns
 * PQDET_ERR_UNSUPPORTED when a level does not qualify for the persistent kernel (H*W a multiple of 128, weights
 * resident in shared memory) or when the batch is large enough for one launch per level to be faster (more than 32
 * tiles of 128 cells per SM); the caller then runs pqdet_head_conv_decode per level. */
int pqdet_head_conv_decode_levels(int n_levels, const float* const* x, const float* const* weight,
                                  const float* const* bias, const int* Cin, const int* H, const int* W,
                                  const float* stride, float* out_decoded, int B, int A, int C, int device,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* PQDET_B200_H_ */
