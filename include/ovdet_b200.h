/*
 * ovdet_b200.h -- C ABI of libovdet_b200.so, the sm_100a CUDA implementation of
 * the detection-geometry / open-vocabulary matching hot path of
 * timsu1104/Open-vocabulary-3D-Object-Detection.
 *
 * The reference has no plugin registry: its native boundary is the Cython
 * extension `utils.box_intersection.box_intersection` (utils/box_intersection.pyx:166-171,
 * imported at utils/box_util.py:13-19) plus the Python call surface of the hot
 * functions.  Each entry point below replaces one of those and
 * cites it.  INTEGRATION.md shows the reference-side ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - `*_host` entry points take HOST pointers (the Cython ABI's numpy buffers),
 *     copy in/out on an internal stream and return after the result is in the
 *     caller's output buffer.  All other entry points take DEVICE pointers and a
 *     `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     they enqueue work and return without synchronising.
 *   - all arrays are C-contiguous; the caller owns every buffer, nothing is
 *     retained after return.
 *   - return value: 0 = ok, <0 = error (OVDET_ERR_*); ovdet_last_error() gives a
 *     thread-local message.  There is no CPU fallback: without a CUDA device every
 *     compute entry point returns OVDET_ERR_CUDA.
 */
#ifndef OVDET_B200_H
#define OVDET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVDET_OK 0
#define OVDET_ERR_INVALID (-1) /* bad shape / null pointer / unsupported size */
#define OVDET_ERR_CUDA (-2)    /* CUDA runtime error, see ovdet_last_error() */
#define OVDET_ERR_UNSUPPORTED (-3)

int ovdet_version(void);                /* 100*major + minor */
const char *ovdet_last_error(void);     /* thread-local, never NULL */
int ovdet_device_count(void);           /* number of visible CUDA devices (0 if none) */
int ovdet_stream_synchronize(void *stream); /* cudaStreamSynchronize: how a host shim waits for the result of the calls above */

/* ------------------------------------------------------------------------- */
/* Rotated / axis-aligned 3D GIoU                                             */
/* replaces generalized_box3d_iou (utils/box_util.py:717-737) and its two     */
/* bodies (:517-618 torch path, :624-714 Cython-backed path).                 */
/* ------------------------------------------------------------------------- */
#define OVDET_GIOU_ROTATED 0x01u     /* rotated_boxes=True: BEV Sutherland-Hodgman clip */
#define OVDET_GIOU_PREFILTER 0x02u   /* skip clip when the axis-aligned BEV overlap is 0 (box_util.py:587-588, pyx:189) */
#define OVDET_GIOU_INTER_ONLY 0x04u  /* return_inter_vols_only=True */
#define OVDET_GIOU_CLIP_F64 0x08u    /* Cython arithmetic: clip in fp64, polygon->fp32, np.dot-style area (pyx:13-19,196-198) */
#define OVDET_GIOU_ENCL_HULL 0x10u   /* convex-hull enclosing volume (utils/box_ops3d.py:533-571) instead of the AABB (:466-514) */

/* corners1 [B,K1,8,3], corners2 [B,K2,8,3] fp32; nums_k2 [B] int64 or NULL;
 * k2_cap > 0 clips only GT columns < k2_cap (4 reproduces the shipped
 * `K2 = rect2.shape[2]` bug, box_intersection.pyx:180); out [B,K1,K2] fp32. */
int ovdet_giou3d_f32(const float *corners1, const float *corners2, const int64_t *nums_k2,
                     int B, int K1, int K2, int k2_cap, unsigned flags, float *out, void *stream);

/* Box decode: box_parametrization_to_corners (datasets/sunrgbd.py:145-148 =
 * flip_axis_to_camera + get_3d_box_batch_tensor, utils/box_util.py:288-352).  center [n,3] in the
 * depth frame, size [n,3] = l,w,h, angle [n] -> corners [n,8,3] in the upright-camera frame. */
int ovdet_box_corners_f32(const float *center, const float *size, const float *angle, int64_t n, float *corners, void *stream);

/* ovdet_giou3d_f32 with the query boxes given as (center1 [B,K1,3], size1 [B,K1,3], angle1 [B,K1]): the decode is
 * fused into the kernel's load stage (28 B/box read instead of 96 B).  corners1_out [B,K1,8,3] is optional. */
int ovdet_giou3d_decode_f32(const float *center1, const float *size1, const float *angle1, const float *corners2,
                            const int64_t *nums_k2, int B, int K1, int K2, int k2_cap, unsigned flags,
                            float *out, float *corners1_out, void *stream);

/* 3D box -> 2D image box of the RegionCLIP crop branch: project_box_3d_cuda
 * (utils/image_util.py:117-134) with SUNRGBD_Calibration_cuda (:247-298), optionally followed by the clip to the
 * image of criterion.py:387-391.  center/size [B,Q,3] (upright depth frame; size used as half extents like the
 * reference), angle [B,Q], rtilt/kmat [B,3,3] row-major per scene, clip_wh [B,2] = (image width, height) or NULL,
 * boxes2d [B,Q,4] = the reference's (x1, y1, x2, y2) = (min v, min u, max v, max u). */
int ovdet_project_box3d_f32(const float *center, const float *size, const float *angle, const float *rtilt,
                            const float *kmat, const float *clip_wh, int B, int Q, float *boxes2d, void *stream);

/* GIoU backward: d loss / d corners1 for the fp32 torch-path GIoU (autograd of
 * generalized_box3d_iou_tensor, utils/box_util.py:517-618; flags = ROTATED | PREFILTER as in the forward).
 * grad_out [B,K1,K2] = d loss / d giou (zero entries are skipped: the loss only touches matched pairs,
 * criterion.py:274-296); grad_corners1 [B,K1,8,3] is overwritten.  No gradient w.r.t. corners2. */
int ovdet_giou3d_backward_f32(const float *corners1, const float *corners2, const int64_t *nums_k2, const float *grad_out,
                              int B, int K1, int K2, unsigned flags, float *grad_corners1, void *stream);

/* ------------------------------------------------------------------------- */
/* The Cython extension ABI: box_intersection(rect1, rect2,                   */
/*   non_rot_inter_areas, nums_k2, inter_areas, approximate)                  */
/* (utils/box_intersection.pyx:166-198).  rect1 [B,K1,4,2], rect2 [B,K2,4,2], */
/* non_rot_inter_areas / inter_areas [B,K1,K2] fp32, nums_k2 [B] int32.       */
/* inter_areas is updated IN PLACE (only where a non-empty clip was computed).*/
/* k2_loop is the reference's `rect2.shape[2]` (=4 as shipped); pass K2 for    */
/* the intended all-columns behaviour.                                        */
/* ------------------------------------------------------------------------- */
int ovdet_box_intersection_f32(const float *rect1, const float *rect2, const float *non_rot_inter_areas,
                               const int32_t *nums_k2, float *inter_areas, int approximate,
                               int B, int K1, int K2, int k2_loop, void *stream);          /* device pointers */
int ovdet_box_intersection_host_f32(const float *rect1, const float *rect2, const float *non_rot_inter_areas,
                                    const int32_t *nums_k2, float *inter_areas, int approximate,
                                    int B, int K1, int K2, int k2_loop);                   /* host pointers */

/* Whole generalized_box3d_iou with HOST buffers (the reference call ends in    */
/* .cpu() round trips, box_util.py:685-698): H2D, kernel, D2H, synchronise.     */
int ovdet_giou3d_host_f32(const float *corners1, const float *corners2, const int64_t *nums_k2,
                          int B, int K1, int K2, int k2_cap, unsigned flags, float *out);

/* ------------------------------------------------------------------------- */
/* Exact pairwise IoU: box3d_iou (utils/box_util.py:116-141), fp64 arithmetic  */
/* on fp32 corners (eval_det.py:120,122).  dets [S,D,8,3], gts [S,G,8,3],      */
/* out [S,D,G] fp64 (iou); out2d may be NULL.  nd/ng [S] int32 counts or NULL. */
/* ------------------------------------------------------------------------- */
int ovdet_box3d_iou_f64(const float *dets, const float *gts, const int32_t *nd, const int32_t *ng,
                        int S, int D, int G, double *out, double *out2d, void *stream);

/* ------------------------------------------------------------------------- */
/* Hungarian matcher (criterion.py:33-92)                                     */
/* ------------------------------------------------------------------------- */
/* cost[b,q,g] = w_class*(-prob[b,q,label[b,g]]) + w_obj*(-obj[b,q])
 *             + w_center*center_dist[b,q,g] + w_giou*(-giou[b,q,g])   (criterion.py:40-63)
 * center_dist == NULL: L1 distance of center_q [B,Q,3] / center_g [B,G,3]
 * (torch.cdist(.., p=1), criterion.py:357-359) is computed in the kernel.
 * gious == NULL: GIoU is computed in the kernel from corners1/corners2 with
 * (giou_flags, k2_cap) as in ovdet_giou3d_f32 and optionally written to gious_out.
 * gt_labels outside [0, C) are clamped into range (the reference's torch.gather, criterion.py:46-48, would raise a device
 * assert there); padded GT columns beyond nactual_gt carry label 0 in the reference's batches, so this never triggers on them. */
int ovdet_matcher_cost_f32(const float *sem_cls_prob, const float *objectness, const float *center_dist,
                           const float *center_q, const float *center_g, const float *gious,
                           const float *corners1, const float *corners2, const int64_t *gt_labels,
                           const int64_t *nactual_gt, int B, int Q, int G, int C,
                           float w_class, float w_obj, float w_center, float w_giou,
                           unsigned giou_flags, int k2_cap, float *gious_out, float *cost, void *stream);

/* Per-sample linear sum assignment on cost[b, :, :nactual_gt[b]] (criterion.py:76-86,
 * scipy.optimize.linear_sum_assignment semantics, fp64 internally).
 * per_prop_gt_inds [B,Q] int64, proposal_matched_mask [B,Q] fp32 (both fully
 * written, zeros where unmatched), col_to_row [B,G] int32 (-1 beyond nactual_gt).
 * scipy raises ValueError on NaN / -inf entries and on infeasible problems; such a sample is left unsolved and MARKED:
 * proposal_matched_mask[b,:] = -1, col_to_row[b,:] = -2 (+inf entries are legal: forbidden pairs). */
int ovdet_lsap_f32(const float *cost, const int64_t *nactual_gt, int B, int Q, int G,
                   int64_t *per_prop_gt_inds, float *proposal_matched_mask, int32_t *col_to_row, void *stream);

/* The whole matcher step of a decoder layer (criterion.py:348-361 + :33-92), or of all layers batched along B, in ONE call:
 * ovdet_matcher_cost_f32 (GIoU computed in the kernel from the corners, L1 centre distance from the centres) followed by
 * ovdet_lsap_f32 on the same stream.  Arguments as in those two. */
int ovdet_matcher_step_f32(const float *sem_cls_prob, const float *objectness, const float *center_q, const float *center_g,
                           const float *corners1, const float *corners2, const int64_t *gt_labels,
                           const int64_t *nactual_gt, int B, int Q, int G, int C,
                           float w_class, float w_obj, float w_center, float w_giou,
                           unsigned giou_flags, int k2_cap, float *gious_out, float *cost,
                           int64_t *per_prop_gt_inds, float *proposal_matched_mask, int32_t *col_to_row, void *stream);

/* ------------------------------------------------------------------------- */
/* Greedy NMS (utils/nms.py:43-162, 3DOVDet_tools/utils/box_3d_utils.py:60-120) */
/* ------------------------------------------------------------------------- */
#define OVDET_NMS_2D 0x01u        /* nms_2d_faster: boxes = x1,y1,x2,y2,score */
#define OVDET_NMS_SAMECLS 0x02u   /* only suppress same-class boxes (cls column after score) */
#define OVDET_NMS_OLD_TYPE 0x04u  /* overlap = inter / vol_j */
#define OVDET_NMS_LHS 0x08u       /* tools variant, `lhs=True` (3DOVDet_tools/utils/box_3d_utils.py:113-116): the better-scoring half of
                                     the boxes a pick suppresses is picked as well (and still removed) */
/* boxes [S,K,ncols] fp64 (dims*2 coords, score, [cls], ...); counts [S] int32 or
 * NULL (=K); vol_eps added to each volume (1e-8 for the tools variant, else 0).
 * keep [S,K] uint8; pick_order [S,K] int32 = indices in the reference's pick
 * order, -1 padded; npick [S] int32.  Scores must be tie-free for a defined order. */
int ovdet_nms_f64(const double *boxes, const int32_t *counts, int S, int K, int ncols,
                  double thr, double vol_eps, unsigned flags,
                  uint8_t *keep, int32_t *pick_order, int32_t *npick, void *stream);

/* parse_predictions (utils/ap_calculator.py:39-238), default branch family:
 * AABB from corners (:157-176, fp64), score = objectness, cls = argmax prob,
 * NMS variant by flags (+ OVDET_PARSE_NO_NMS), then keep = picked & obj > conf_thresh.
 * corners [S,K,8,3] fp32, probs [S,K,C] fp32, obj [S,K] fp32, nonempty [S,K] uint8 or NULL.
 * Outputs: pred_mask [S,K] uint8 (NMS pick mask), keep [S,K] uint8,
 * pred_cls [S,K] int32 (argmax), pred_cls_prob [S,K] fp32 (max). */
#define OVDET_PARSE_NO_NMS 0x100u
int ovdet_parse_predictions_f32(const float *corners, const float *probs, const float *obj,
                                const uint8_t *nonempty, int S, int K, int C,
                                double nms_iou, float conf_thresh, unsigned flags,
                                uint8_t *pred_mask, uint8_t *keep, int32_t *pred_cls, float *pred_cls_prob,
                                void *stream);

/* ------------------------------------------------------------------------- */
/* AP matching (utils/eval_det.py:66-155) on scene-major device arrays         */
/* ------------------------------------------------------------------------- */
/* For every scene s, class c and kept detection d (per_class_proposal layout,
 * ap_calculator.py:196-210): score[s,c,d] = probs[s,d,c]*obj[s,d]; best GT of
 * class c by exact IoU (first max, strict >); TP at threshold t iff
 * ovmax > t and d is the highest-scoring claimant of that GT.
 * Writes records class-major: rec_score [C, S*K] fp32 (-inf for absent
 * detections), rec_tp [C, S*K] uint8 bit t = TP at thr[t] (nthr <= 8),
 * npos [C] int64 += number of GT of class c.  iou_ws: [S,K,G] fp64 workspace.
 * det_cls != NULL selects the single-class layouts (ap_calculator.py:212-236):
 * detection d only appears in class det_cls[s,d] with score obj[s,d] (probs unused). */
int ovdet_ap_match(const float *corners, const float *probs, const float *obj, const uint8_t *keep,
                   const int32_t *det_cls, const float *gt_corners, const int64_t *gt_labels, const uint8_t *gt_present,
                   int S, int K, int G, int C, const double *thr, int nthr,
                   double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos, void *stream);

/* The same matching on a PRECOMPUTED IoU matrix iou [S,K,G] fp64 (any IoU definition): what eval_det_cls does with a
 * caller-supplied get_iou_func.  ovdet_aabb_iou_f64 fills the matrix with the tools' axis-aligned IoU, get_iou / calc_iou
 * (3DOVDet_tools/utils/evaluation/eval_det.py:63-77, evaluation/box_util.py:287-309): dets [S,D,6], gts [S,G,6] fp64 =
 * (centre, lengths), result clamped to [0, 1]; nd / ng [S] int32 counts or NULL. */
int ovdet_aabb_iou_f64(const double *dets, const double *gts, const int32_t *nd, const int32_t *ng,
                       int S, int D, int G, double *out, void *stream);
int ovdet_ap_match_iou(const double *iou, const float *probs, const float *obj, const uint8_t *keep,
                       const int32_t *det_cls, const int64_t *gt_labels, const uint8_t *gt_present,
                       int S, int K, int G, int C, const double *thr, int nthr,
                       float *rec_score, uint8_t *rec_tp, int64_t *npos, void *stream);

/* Per-class AP from records (eval_det.py:108-153 + voc_ap :23-54): sort each
 * class segment by descending score, cumulative TP/FP, precision/recall, AP.
 * rec_score/rec_tp [C,N] (entries with score == -inf are absent); npos [C] int64.
 * ap [nthr,C] fp64, recall [nthr,C] fp64 (last recall, 0 when no detections),
 * n_det [C] int64 = number of present records (may be NULL).  If rec_out/prec_out
 * != NULL they get the full curves [nthr,C,N] fp64 (first n_det[c] entries of each
 * row are valid).  ws: device workspace of ovdet_ap_reduce_ws_bytes(C,N) bytes. */
size_t ovdet_ap_reduce_ws_bytes(int C, int64_t N);
int ovdet_ap_reduce(const float *rec_score, const uint8_t *rec_tp, const int64_t *npos,
                    int C, int64_t N, int nthr, int use_07_metric,
                    double *ap, double *recall, int64_t *n_det, double *rec_out, double *prec_out,
                    void *ws, size_t ws_bytes, void *stream);

/* Compact AP reduction without a global sort (csrc/ap_compact.cu): VOC AP only needs, for
 * every TP record, its position in the score order.  Four stages so that a scene-sharded
 * multi-GPU caller can exchange between them (all-gather the TP lists after collect,
 * all-reduce `hist` after hist).  cap = per-class TP-list capacity, a power of two in
 * [1024, 16384]; the lists are pre-filled with key 0xFFFFFFFF / bits 0 so that lists of several
 * ranks can simply be concatenated (to another power of two) before ovdet_apc_sort.
 *   collect: tp_key u32 [C,cap], tp_bits u8 [C,cap], tp_cnt i32 [C] (may exceed cap = overflow),
 *            nvalid i64 [C] (present records);  sort: descending score per class;
 *   hist:    u32 [C,cap+1], bucket = number of TP-list scores above the record's score;
 *   final:   ap/recall fp64 [nthr,C], n_det i64 [C] (nullable), overflow i32 [1] (nullable). */
int ovdet_apc_collect(const float *rec_score, const uint8_t *rec_tp, int C, int64_t N, int cap,
                      uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int64_t *nvalid, void *stream);
int ovdet_apc_sort(uint32_t *tp_key, uint8_t *tp_bits, int C, int cap, void *stream);
int ovdet_apc_hist(const float *rec_score, int C, int64_t N, const uint32_t *tp_key, int cap, uint32_t *hist, void *stream);
int ovdet_apc_final(const uint8_t *tp_bits, const int32_t *tp_cnt, const uint32_t *hist, const int64_t *npos,
                    const int64_t *nvalid, int C, int cap, int nthr, int use_07_metric,
                    double *ap, double *recall, int64_t *n_det, int32_t *overflow, void *stream);

/* Fused AP front end: parse_predictions (utils/ap_calculator.py:39-238: argmax class, AABB from corners, NMS variant by
 * `flags` as in ovdet_parse_predictions_f32, confidence gate) AND the AP matching of ovdet_ap_match
 * (utils/eval_det.py:117-140) in one pass per scene: corners and class probabilities are read once, the keep mask never
 * leaves the chip.  OVDET_FRONT_PER_CLASS = the per_class_proposal layout (score = prob*obj for every class,
 * ap_calculator.py:196-210); without it a detection only appears in its argmax class with score = objectness, or the
 * class confidence with OVDET_FRONT_CLS_CONF (:212-236).  rec_score [C, S*K] as in ovdet_ap_match; rec_tp may be NULL.
 * tp_key/tp_bits [C, tp_cap] + tp_cnt i32 [C] (all three or none): every true-positive record is APPENDED as
 * (descending-score key, threshold bits) -- the caller zeroes tp_cnt once per evaluation and may call this for several
 * batches; tp_cnt may run past tp_cap (overflow, reported by ovdet_apx_reduce).  npos [C] int64 += GT count.
 * keep_out [S,K] uint8 optional. */
#define OVDET_FRONT_PER_CLASS 0x1000u
#define OVDET_FRONT_CLS_CONF 0x2000u
#define OVDET_FRONT_GT_PRESENT_F32 0x4000u   /* gt_present is the reference's fp32 mask [S,G] instead of uint8 */
#define OVDET_FRONT_RESET 0x8000u            /* zero tp_cnt [C] and npos [C] on the stream first (start of an evaluation) */
int ovdet_ap_front_f32(const float *corners, const float *probs, const float *obj, const uint8_t *nonempty,
                       const float *gt_corners, const int64_t *gt_labels, const void *gt_present,
                       int S, int K, int G, int C, double nms_iou, float conf_thresh, unsigned flags,
                       const double *thr, int nthr, double *iou_ws, float *rec_score, uint8_t *rec_tp, int64_t *npos,
                       uint32_t *tp_key, uint8_t *tp_bits, int32_t *tp_cnt, int tp_cap, uint8_t *keep_out, void *stream);

/* ------------------------------------------------------------------------- */
/* Scene-sharded AP reduction with a device-side exchange (the one exchange step of the path;        */
/* replaces the reference's all-gather of every output and input tensor, engine.py:207-209 via utils/dist.py:159-176).   */
/* ------------------------------------------------------------------------- */
/* Symmetric buffers: one allocation per rank, mapped into every peer process (CUDA IPC), so that kernels exchange the
 * per-class TP lists and bucket histograms with plain stores over NVLink and flag words -- no collective launch.
 *   alloc: cudaMalloc + zero fill;  export_handle: 64 opaque bytes to send to the peers (any transport);
 *   open: map a peer's buffer from its handle (enables peer access);  close / free. */
#define OVDET_SYMM_HANDLE_BYTES 64
int ovdet_symm_alloc(size_t bytes, void **dev_ptr);
int ovdet_symm_free(void *dev_ptr);
int ovdet_symm_export(void *dev_ptr, void *handle_out);
int ovdet_symm_open(const void *handle, void **peer_ptr);
int ovdet_symm_close(void *peer_ptr);

/* Sizes of the two workspaces of ovdet_apx_reduce: `local` (this rank only: merged lists, edges, histogram, counters,
 * epoch word; zero it once) and `symm` (the symmetric buffer peers write into; zero-filled by ovdet_symm_alloc). */
size_t ovdet_apx_local_bytes(int C, int cap_total);
size_t ovdet_apx_symm_bytes(int C, int cap_total, int world);

/* AP / recall of all classes and thresholds from this rank's record blocks and TP lists (ovdet_ap_front_f32; row
 * stride cap_list), reduced over `world` scene-sharded ranks.  One call enqueues the whole chain on `stream`:
 *   push    this rank's TP list of every class is SORTED in R slices (one CTA each; R = 1 today) and every sorted
 *           run (keys, bits, count, npos) is stored into every peer's symmetric buffer, then a flag per (run, class)
 *   merge   per class, a cluster of 8 CTAs: wait for the world * R run flags, rank every entry among the other runs by
 *           binary search (merged position = sum of ranks; <= cap_total entries, a power of two in [1024, 16384]),
 *           then the bin edges of the sorted list
 *   hist    one streaming pass over the LOCAL records: bucket = position in the merged list; a (peer, class) grid then
 *           ships the class's partial histogram to every peer, with a flag
 *   final   per (class, threshold): wait for the peers' histograms, sum them, prefix sums -> positions, precision
 *           envelope, VOC AP (utils/eval_det.py:23-54, :143-153)
 * blocks/block_n: host arrays of nblocks device pointers to rec_score blocks [C, block_n[i]] fp32.
 * peers: host array of `world` device pointers = every rank's symmetric buffer as mapped in THIS process
 * (peers[rank] = the local one).  With world == 1 `peers` may be NULL (one CTA per class sorts the whole list) or point
 * at ONE plain device buffer of ovdet_apx_symm_bytes(C, cap_total, 1) zeroed bytes, which selects the push + cluster
 * merge above; OVDET_APX_FORCE_EXCHANGE additionally runs the histogram exchange (push to self, slot sums).
 * result (device, fp64 [2*nthr*C + C + 3]): ap [nthr,C] | recall [nthr,C] | n_det [C] | overflow | max per-rank list
 * count | largest merged list count; overflow > 0 means a merged list did not fit cap_total (the local lists are intact:
 * retry with a larger cap_total) or a local list overflowed cap_list; overflow < 0 = a peer's flag never arrived (timeout).
 * A symmetric buffer sized for a larger cap_total may be used with a smaller one.
 * result_host: optional host copy of `result`; pinned (mapped) memory is written by the final kernel itself, other memory
 * by an async D2H copy on `stream`; either way the caller synchronises the stream before reading it.
 * Every rank must call this the same number of times (an epoch word in `local` tags the flags). */
#define OVDET_APX_FORCE_EXCHANGE 0x1u
#define OVDET_APX_USE_07_METRIC 0x2u
/* Stage selection (default = all three).  A consumer kernel spins on flags its peers raise, which is only safe when the
 * peers run on OTHER GPUs (or have already finished): a harness that drives several ranks' buffers on ONE device must
 * run the stages rank by rank -- all pushes, then all merge+hist, then all finals -- so that no kernel ever waits. */
#define OVDET_APX_STAGE_PUSH 0x10u
#define OVDET_APX_STAGE_MERGE_HIST 0x20u
#define OVDET_APX_STAGE_FINAL 0x40u
int ovdet_apx_reduce(const void *const *blocks, const int64_t *block_n, int nblocks, int C,
                     const uint32_t *tp_key, const uint8_t *tp_bits, const int32_t *tp_cnt, const int64_t *npos,
                     int cap_list, int cap_total, int nthr, unsigned flags, int rank, int world,
                     void *const *peers, void *local_ws, double *result, double *result_host, void *stream);

/* ------------------------------------------------------------------------- */
/* Open-vocabulary logits (models/model_3detr.py:237-238, :58-62;              */
/* utils/ulip_losses.py:39-47): logits = scale * norm?(x) @ norm?(T)^T,        */
/* prob = softmax(logits), sem_cls_prob = prob[:, :-1], objectness = 1-prob[:,-1]. */
/* x [M,K] bf16, text [N,K] bf16 (row-major, K contiguous).  Outputs (any may  */
/* be NULL): logits fp32 [M, ld_logits] (N columns written), prob bf16          */
/* [M, ld_prob] = the FULL softmax row (N columns; ld_prob % 8 == 0 so rows are  */
/* 16-byte aligned -- the reference's sem_cls_prob is the view prob[:, :N-1],    */
/* model_3detr.py:62), objectness fp32 [M] = 1 - prob[:, N-1].                  */
/* tcgen05 + TMEM + TMA; requires K % 64 == 0, N <= 2048.                       */
/* ------------------------------------------------------------------------- */
#define OVDET_LOGITS_L2NORM 0x01u
int ovdet_clip_logits_bf16(const void *x, const void *text, int M, int K, int N, unsigned flags, float scale,
                           float *logits, int ld_logits, void *prob, int ld_prob, float *objectness, void *stream);

/* ------------------------------------------------------------------------- */
/* Pseudo-label "NMS + IoU filtering" per scene                               */
/* (3DOVDet_tools/scannet/lift_boxes.py:139-166): class-wise NMS(nms_thr) ->   */
/* argmax-IoU match to the proposal pool (>= match_thr, box_3d_iou with +1e-5) */
/* keeping the highest-score label per pool box -> size-scored class-wise NMS.  */
/* boxes [S,P,8] fp64 = x1..z2,score,label; pool [S,M,6] fp64.                 */
/* Outputs per pool box: out_label [S,M] fp64 (-100 = unmatched), out_score    */
/* [S,M] fp64, out_keep [S,M] uint8 (survives the final NMS), nms1_keep [S,P]. */
/* ------------------------------------------------------------------------- */
int ovdet_pseudo_filter_f64(const double *boxes, const double *pool, const int32_t *nboxes, const int32_t *npool,
                            int S, int P, int M, double nms_thr, double match_thr, double size_nms_thr,
                            uint8_t *nms1_keep, double *out_label, double *out_score, uint8_t *out_keep,
                            void *stream);

/* ------------------------------------------------------------------------- */
/* Point-in-box passes                                                        */
/* ------------------------------------------------------------------------- */
/* remove_empty_box of parse_predictions (utils/ap_calculator.py:70-84 ->
 * utils/box_util.py:22-31): counts[s,k] = number of points of scene s inside predicted box k.
 * point_cloud fp32 [S,N,point_stride>=3] in the DEPTH frame, corners fp32 [S,K,8,3] in the
 * upright-camera frame (the kernel applies flip_axis_to_depth, ap_calculator.py:22-26). */
int ovdet_points_in_boxes_count(const float *point_cloud, int S, int N, int point_stride,
                                const float *corners, int K, int32_t *counts, void *stream);

/* LabelFormatter.gen_pseudo vote (utils/label_formatter.py:150-159, crop_pc :183-188):
 * for each box (centre xyz, size xyz, ... ; row stride box_stride fp64) the mode of the labels
 * (integers 0..63 stored as fp64) of the points inside its axis-aligned extent, skipping
 * ignore_label; mode_out = -1 when no labelled point is inside, count_out = points counted. */
int ovdet_box_label_mode(const double *points, const double *labels, int N, const double *boxes, int box_stride, int M,
                         double ignore_label, int32_t *mode_out, int32_t *count_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OVDET_B200_H */
