"""Drop-in for the reference's ``utils/eval_det.py`` (VOC-style AP matching).

The dict-based entry points keep their signatures; internally everything is
packed into scene-major device arrays and evaluated by two kernels families
(csrc/eval.cu): ``ovdet_ap_match`` (exact fp64 IoU det x GT, argmax, first-claim
TP flags for all classes/thresholds at once) and ``ovdet_ap_reduce`` (per-class
segmented radix sort by descending score + scans -> precision/recall/AP).
``ap_from_arrays`` is the array-level engine the AP calculator uses directly.
"""
import numpy as np
import torch

from .. import _capi as C


def voc_ap(rec, prec, use_07_metric=False):
    """utils/eval_det.py:23-54 on an already sorted PR curve (host arrays, O(nd));
    the GPU path computes the same quantity inside ``ovdet_ap_reduce``."""
    rec = np.asarray(rec, np.float64)
    prec = np.asarray(prec, np.float64)
    if use_07_metric:
        ap = 0.0
        for t in np.arange(0.0, 1.1, 0.1):
            p = np.max(prec[rec >= t]) if np.sum(rec >= t) != 0 else 0
            ap = ap + p / 11.0
        return ap
    mrec = np.concatenate(([0.0], rec, [1.0]))
    mpre = np.concatenate(([0.0], prec, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def ap_match(corners, probs, obj, keep, gt_corners, gt_labels, gt_present, num_classes, thresholds, det_cls=None):
    """Device tensors in, class-major records out:
    rec_score fp32 [C, S*K] (-inf = absent), rec_tp uint8 [C, S*K] (bit t = TP at
    thresholds[t]), npos int64 [C]."""
    C.require_cuda(corners)
    dev = corners.device
    S, K = corners.shape[0], corners.shape[1]
    G = gt_corners.shape[1]
    Cn = int(num_classes)
    f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()
    corners, probs, obj, gt_corners = f32(corners), f32(probs), f32(obj), f32(gt_corners)
    keep = keep.to(device=dev, dtype=torch.uint8).contiguous()
    gt_labels = gt_labels.to(device=dev, dtype=torch.int64).contiguous()
    gt_present = (gt_present.to(dev) != 0).to(torch.uint8).contiguous()
    det_cls = None if det_cls is None else det_cls.to(device=dev, dtype=torch.int32).contiguous()
    thr = np.ascontiguousarray(np.asarray(thresholds, np.float64))
    rec_score = torch.empty((Cn, S * K), dtype=torch.float32, device=dev)
    rec_tp = torch.empty((Cn, S * K), dtype=torch.uint8, device=dev)
    npos = torch.zeros((Cn,), dtype=torch.int64, device=dev)
    iou_ws = torch.empty((S, K, max(G, 1)), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_ap_match(C.ptr(corners), C.ptr(probs), C.ptr(obj), C.ptr(keep), C.ptr(det_cls),
                                       C.ptr(gt_corners), C.ptr(gt_labels), C.ptr(gt_present), S, K, G, Cn,
                                       thr.ctypes.data, len(thr), C.ptr(iou_ws), C.ptr(rec_score), C.ptr(rec_tp),
                                       C.ptr(npos), C.stream(dev)))
    return rec_score, rec_tp, npos


def ap_reduce(rec_score, rec_tp, npos, nthr, use_07_metric=False, curves=False):
    """Records [C,N] -> (ap [nthr,C], recall [nthr,C], n_det [C][, rec, prec [nthr,C,N]]) on device."""
    C.require_cuda(rec_score)
    dev = rec_score.device
    Cn, N = rec_score.shape
    rec_score = rec_score.contiguous()
    rec_tp = rec_tp.contiguous()
    npos = npos.to(device=dev, dtype=torch.int64).contiguous()
    ap = torch.empty((nthr, Cn), dtype=torch.float64, device=dev)
    recall = torch.empty((nthr, Cn), dtype=torch.float64, device=dev)
    ndet = torch.empty((Cn,), dtype=torch.int64, device=dev)
    rec = torch.zeros((nthr, Cn, N), dtype=torch.float64, device=dev) if curves else None
    prec = torch.zeros((nthr, Cn, N), dtype=torch.float64, device=dev) if curves else None
    L = C.lib()
    nbytes = L.ovdet_ap_reduce_ws_bytes(Cn, N)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        C.check(L.ovdet_ap_reduce(C.ptr(rec_score), C.ptr(rec_tp), C.ptr(npos), Cn, N, nthr, int(bool(use_07_metric)),
                                  C.ptr(ap), C.ptr(recall), C.ptr(ndet), C.ptr(rec), C.ptr(prec), C.ptr(ws), nbytes,
                                  C.stream(dev)))
    return (ap, recall, ndet, rec, prec) if curves else (ap, recall, ndet)


APC_MAXCAP = 16384


def _pow2_at_least(x, lo=1024):
    c = lo
    while c < x:
        c <<= 1
    return c


def ap_reduce_compact(rec_score, rec_tp, npos, nthr, cap=4096, use_07_metric=False, distributed=False):
    """AP without a global sort (csrc/ap_compact.cu).  ``cap`` = per-class, per-rank TP-list capacity (power of two
    >= 1024); ``npos`` is the LOCAL GT count (summed across ranks here when ``distributed``).  No host sync inside:
    returns one device tensor ``res`` fp64 [2*nthr*C + 1 + C] = ap | recall | overflow flag | n_det, to be read back by
    the caller in a single D2H copy; overflow != 0 means a TP list did not fit and the caller must retry with a
    larger ``cap`` or use the sort-based ap_reduce.  Returns None when world*cap exceeds the shared-memory sort."""
    C.require_cuda(rec_score)
    dev = rec_score.device
    Cn, N = rec_score.shape
    world = 1
    if distributed:
        import torch.distributed as dist
        world = dist.get_world_size()
    cap = _pow2_at_least(int(cap), 32 if world > 1 else 1024)
    cap_total = _pow2_at_least(cap * world)
    if cap_total > APC_MAXCAP or (world == 1 and cap < 1024):
        return None
    L = C.lib()
    st = C.stream(dev)
    rec_score, rec_tp = rec_score.contiguous(), rec_tp.contiguous()
    if world == 1:
        # single rank: four library calls on one workspace, nothing else on the stream (every extra torch op is ~5 us
        # of host time against ~170 us of kernels); the result is a byte-packed view read back in one copy
        a16 = lambda n: (n + 15) // 16 * 16
        sizes = [Cn * cap * 4, Cn * cap, Cn * 4, Cn * 8, Cn * (cap + 1) * 4, Cn * 8]   # keys | bits | cnt | nvalid | hist | npos
        res_bytes = 2 * nthr * Cn * 8 + Cn * 8 + 16                                       # ap | recall | n_det | overflow
        offs = np.concatenate([[0], np.cumsum([a16(x) for x in sizes])])
        ws = torch.empty((int(offs[-1]) + res_bytes,), dtype=torch.uint8, device=dev)
        base = ws.data_ptr()
        k_p, b_p, cnt_p, nv_p, h_p, npos_p = (base + int(o) for o in offs[:-1])
        res = ws[int(offs[-1]):]
        r_p = res.data_ptr()
        ws[int(offs[5]):int(offs[5]) + Cn * 8].view(torch.int64).copy_(npos.to(device=dev, dtype=torch.int64))
        with torch.cuda.device(dev):
            C.check(L.ovdet_apc_collect(C.ptr(rec_score), C.ptr(rec_tp), Cn, N, cap, k_p, b_p, cnt_p, nv_p, st))
            C.check(L.ovdet_apc_sort(k_p, b_p, Cn, cap, st))
            C.check(L.ovdet_apc_hist(C.ptr(rec_score), Cn, N, k_p, cap, h_p, st))
            C.check(L.ovdet_apc_final(b_p, cnt_p, h_p, npos_p, nv_p, Cn, cap, nthr, int(bool(use_07_metric)), r_p,
                                      r_p + 8 * nthr * Cn, r_p + 16 * nthr * Cn, r_p + 16 * nthr * Cn + 8 * Cn, st))
        return res
    raise C.OvdetError("the scene-sharded reduction lives in ApxReducer (ovdet_apx_reduce): device-side exchange over "
                       "symmetric buffers instead of collectives around this function")


def unpack_compact(res, nthr, Cn, with_max=False):
    """Host view of ap_reduce_compact's packed result -> (ap [nthr,C], recall [nthr,C], overflow, n_det [C]
    [, largest per-rank per-class TP count, -1 if not reported]).  One device-to-host copy.
    (uint8 = the single-rank byte layout, float64 = the distributed one.)"""
    k = nthr * Cn
    if res.dtype == torch.uint8:
        raw = res.cpu().numpy()
        f = raw[:16 * k].view(np.float64)
        nd = raw[16 * k:16 * k + 8 * Cn].view(np.int64)
        ovf = int(raw[16 * k + 8 * Cn:16 * k + 8 * Cn + 4].view(np.int32)[0])
        out = (f[:k].reshape(nthr, Cn), f[k:].reshape(nthr, Cn), ovf, nd.copy())
        return out + (-1,) if with_max else out
    r = res.cpu().numpy()
    out = (r[:k].reshape(nthr, Cn), r[k:2 * k].reshape(nthr, Cn), int(r[2 * k]), r[2 * k + 1:2 * k + 1 + Cn].astype(np.int64))
    return out + (int(r[2 * k + 1 + Cn]),) if with_max else out


# ----------------------------------------------------------------------------- fused front end + exchange reducer
def nms_flags(cfg):
    """parse_predictions' NMS branch (utils/ap_calculator.py:86-189) as the kernels' flag word."""
    if cfg.get("no_nms", False):
        return C.PARSE_NO_NMS
    f = C.NMS_OLD_TYPE if cfg["use_old_type_nms"] else 0
    if not cfg["use_3d_nms"]:
        return f | C.NMS_2D
    if cfg["cls_nms"]:
        f |= C.NMS_SAMECLS
    return f


class TpLists(object):
    """Per-class true-positive lists of one rank (device): (descending-score key, threshold bits) appended by
    ``ovdet_ap_front_f32``; ``counters`` holds tp_cnt i32 [C] and npos i64 [C] in one buffer (one memset to reset)."""

    def __init__(self, num_classes, device, cap_list=16384):
        self.C, self.cap_list, self.device = int(num_classes), int(cap_list), device
        assert self.cap_list % 16 == 0
        self.tp_key = torch.empty((self.C, self.cap_list), dtype=torch.int32, device=device)
        self.tp_bits = torch.empty((self.C, self.cap_list), dtype=torch.uint8, device=device)
        self._cnt_words = (self.C + 1) // 2 * 2                       # npos stays 8-byte aligned behind tp_cnt
        self.counters = torch.zeros((self._cnt_words + 2 * self.C,), dtype=torch.int32, device=device)
        self.tp_cnt_ptr = self.counters.data_ptr()
        self.npos_ptr = self.counters.data_ptr() + 4 * self._cnt_words
        self.key_ptr, self.bits_ptr = self.tp_key.data_ptr(), self.tp_bits.data_ptr()
        self.dirty = False      # counters must be zeroed before the next append

    def reset(self):
        """Counters back to zero -- lazily: the next ``ap_front`` launch does it on its stream (no separate torch op)."""
        self.dirty = True

    def flush_reset(self):
        """Apply a pending reset now (for readers of the counters that come before any further append)."""
        if self.dirty:
            self.counters.zero_()
            self.dirty = False

    @property
    def tp_cnt(self):
        return self.counters[:self.C]

    @property
    def npos(self):
        return self.counters[self._cnt_words:].view(torch.int64)


def ap_front(corners, probs, obj, nonempty, gt_corners, gt_labels, gt_present, num_classes, thresholds, cfg, lists,
             want_tp_records=False, iou_ws=None, stream=None):
    """One launch per batch: parse_predictions + AP matching (``ovdet_ap_front_f32``).  Returns the batch's record block
    ``rec_score`` fp32 [C, S*K] (and ``rec_tp`` uint8 when asked); true positives are appended to ``lists``."""
    C.require_cuda(corners)
    dev = corners.device
    S, K = corners.shape[0], corners.shape[1]
    G = gt_corners.shape[1]
    Cn = int(num_classes)
    corners, probs, obj, gt_corners = (C.as_input(t, torch.float32, dev) for t in (corners, probs, obj, gt_corners))
    assert probs.shape[-1] == Cn, "sem_cls_probs must have num_semcls columns"
    gt_labels = C.as_input(gt_labels, torch.int64, dev)
    flags = nms_flags(cfg)
    if lists.dirty:
        flags |= C.FRONT_RESET
        lists.dirty = False
    if gt_present.dtype is torch.float32:     # the reference's mask dtype: read as is
        gt_present = C.as_input(gt_present, torch.float32, dev)
        flags |= C.FRONT_GT_PRESENT_F32
    elif gt_present.dtype is torch.uint8:
        gt_present = C.as_input(gt_present, torch.uint8, dev)
    else:
        gt_present = (gt_present.to(dev) != 0).to(torch.uint8).contiguous()
    ne = None if nonempty is None else (nonempty.to(dev) != 0).to(torch.uint8).contiguous()
    thr = thresholds if isinstance(thresholds, np.ndarray) and thresholds.dtype == np.float64 else \
        np.ascontiguousarray(np.asarray(thresholds, np.float64))
    if cfg["per_class_proposal"]:
        assert cfg["use_cls_confidence_only"] is False
        flags |= C.FRONT_PER_CLASS
    elif cfg["use_cls_confidence_only"]:
        flags |= C.FRONT_CLS_CONF
    rec_score = torch.empty((Cn, S * K), dtype=torch.float32, device=dev)
    rec_tp = torch.empty((Cn, S * K), dtype=torch.uint8, device=dev) if want_tp_records else None
    if iou_ws is None or iou_ws.numel() < S * K * max(G, 1):
        iou_ws = torch.empty((S * K * max(G, 1),), dtype=torch.float64, device=dev)   # only touched by the dense fallback
    with C.on_device(dev):
        C.check(C.lib().ovdet_ap_front_f32(
            corners.data_ptr(), probs.data_ptr(), obj.data_ptr(), C.ptr(ne), gt_corners.data_ptr(), gt_labels.data_ptr(),
            gt_present.data_ptr(), S, K, G, Cn, float(cfg["nms_iou"]), float(cfg["conf_thresh"]), flags, thr.ctypes.data,
            len(thr), iou_ws.data_ptr(), rec_score.data_ptr(), C.ptr(rec_tp), lists.npos_ptr, lists.key_ptr, lists.bits_ptr,
            lists.tp_cnt_ptr, lists.cap_list, None, C.stream(dev) if stream is None else stream))
    return rec_score, rec_tp, iou_ws


class ApxReducer(object):
    """Workspaces of ``ovdet_apx_reduce`` for one (classes, thresholds, merged-list capacity, rank set): the rank-local
    scratch (with the epoch word) and, when several ranks take part, the symmetric buffer the peers store into.  Creating
    one with ``world > 1`` is collective (IPC handle exchange)."""

    def __init__(self, num_classes, nthr, cap_total, device, symm=None, rank=0, world=1, force_exchange=False, group=None):
        from .. import dist as D
        L = C.lib()
        self.C, self.nthr, self.cap_total, self.device = int(num_classes), int(nthr), int(cap_total), device
        self.rank, self.world = int(rank), int(world)
        self.exchange = self.world > 1 or force_exchange
        self.local = torch.zeros((L.ovdet_apx_local_bytes(self.C, self.cap_total),), dtype=torch.uint8, device=device)
        self.symm, self._own_symm = symm, False
        if self.exchange and self.symm is None:
            with torch.cuda.device(device):
                self.symm = D.SymmetricBuffer(L.ovdet_apx_symm_bytes(self.C, self.cap_total, self.world), group=group, device=device)
            self._own_symm = True
        self._peers = self.symm.peers_array if self.symm is not None else None
        self.nres = 2 * self.nthr * self.C + self.C + 3
        self.result = torch.empty((self.nres,), dtype=torch.float64, device=device)
        self.result_host = torch.empty((self.nres,), dtype=torch.float64).pin_memory()
        self._res_np = self.result_host.numpy()
        self._local_ptr, self._result_ptr, self._result_host_ptr = self.local.data_ptr(), self.result.data_ptr(), self.result_host.data_ptr()

    def close(self):
        if self._own_symm and self.symm is not None:
            self.symm.close()
            self.symm = None

    def launch(self, blocks, lists, use_07_metric=False, stream=None, stages=0):
        """Enqueue the whole reduction (and the D2H of the packed result) on the current stream; no host sync.
        ``stages`` (C.APX_STAGE_*) restricts the call to some stages -- only for harnesses that drive several ranks'
        buffers on one device, where a kernel must never wait for a peer that has not run yet."""
        import ctypes
        nb = len(blocks)
        ptrs = (ctypes.c_void_p * max(nb, 1))(*[b.data_ptr() for b in blocks])
        sizes = (ctypes.c_int64 * max(nb, 1))(*[b.shape[1] for b in blocks])
        flags = (C.APX_FORCE_EXCHANGE if (self.exchange and self.world == 1) else 0) | (C.APX_USE_07_METRIC if use_07_metric else 0) | stages
        with C.on_device(self.device):
            C.check(C.lib().ovdet_apx_reduce(
                ptrs, sizes, nb, self.C, lists.key_ptr, lists.bits_ptr, lists.tp_cnt_ptr, lists.npos_ptr,
                lists.cap_list, self.cap_total, self.nthr, flags, self.rank, self.world,
                self._peers, self._local_ptr, self._result_ptr,
                self._result_host_ptr, C.stream(self.device) if stream is None else stream))

    def read(self):
        """After a stream synchronise: (ap [nthr,C], recall [nthr,C], n_det [C], overflow, max rank count, max merged count)."""
        r = self._res_np.copy()     # one 800-byte copy out of the pinned buffer; the rest are views
        k = self.nthr * self.C
        return (r[:k].reshape(self.nthr, self.C), r[k:2 * k].reshape(self.nthr, self.C),
                r[2 * k:2 * k + self.C].astype(np.int64), int(r[2 * k + self.C]), int(r[2 * k + self.C + 1]), int(r[2 * k + self.C + 2]))



def _pack(pred_all, gt_all):
    """{img: [(cls, box, score)]}, {img: [(cls, box)]} -> padded scene-major arrays."""
    imgs = list(dict.fromkeys(list(pred_all.keys()) + list(gt_all.keys())))
    classes = []
    for img in pred_all:
        for cl, _, _ in pred_all[img]:
            classes.append(cl)
    for img in gt_all:
        for cl, _ in gt_all[img]:
            classes.append(cl)
    classes = list(dict.fromkeys(classes))
    cidx = {c: i for i, c in enumerate(classes)}
    S = len(imgs)
    K = max([len(pred_all.get(i, [])) for i in imgs] + [1])
    G = max([len(gt_all.get(i, [])) for i in imgs] + [1])
    corners = np.zeros((S, K, 8, 3), np.float32)
    score = np.zeros((S, K), np.float32)
    dcls = np.zeros((S, K), np.int32)
    keep = np.zeros((S, K), np.uint8)
    gcorn = np.zeros((S, G, 8, 3), np.float32)
    glab = np.zeros((S, G), np.int64)
    gpres = np.zeros((S, G), np.uint8)
    for si, img in enumerate(imgs):
        for k, (cl, box, sc) in enumerate(pred_all.get(img, [])):
            corners[si, k] = box; score[si, k] = sc; dcls[si, k] = cidx[cl]; keep[si, k] = 1
        for g, (cl, box) in enumerate(gt_all.get(img, [])):
            gcorn[si, g] = box; glab[si, g] = cidx[cl]; gpres[si, g] = 1
    pred_classes = set(cl for img in pred_all for cl, _, _ in pred_all[img])
    return classes, pred_classes, corners, score, dcls, keep, gcorn, glab, gpres


def _eval_packed(pred_all, gt_all, ovthresh, use_07_metric):
    classes, pred_classes, corners, score, dcls, keep, gcorn, glab, gpres = _pack(pred_all, gt_all)
    dev = torch.device("cuda")
    t = lambda a: torch.as_tensor(a, device=dev)
    rs, rt, npos = ap_match(t(corners), None, t(score), t(keep), t(gcorn), t(glab), t(gpres), len(classes),
                            [ovthresh], det_cls=t(dcls))
    ap, recall, ndet, rec, prec = ap_reduce(rs, rt, npos, 1, use_07_metric, curves=True)
    ap, ndet, rec, prec = ap.cpu().numpy(), ndet.cpu().numpy(), rec.cpu().numpy(), prec.cpu().numpy()
    out_rec, out_prec, out_ap = {}, {}, {}
    for ci, cl in enumerate(classes):
        if cl in pred_classes:
            n = int(ndet[ci])
            out_rec[cl], out_prec[cl], out_ap[cl] = rec[0, ci, :n].copy(), prec[0, ci, :n].copy(), float(ap[0, ci])
        else:  # utils/eval_det.py:266-269
            out_rec[cl], out_prec[cl], out_ap[cl] = 0, 0, 0
    return out_rec, out_prec, out_ap


def calc_iou(box_a, box_b):
    """3DOVDet_tools/utils/evaluation/box_util.py:287-309: IoU of two axis-aligned boxes given as (centre, lengths).
    Host scalar form for interface parity -- the evaluation itself runs ``ovdet_aabb_iou_f64`` on all pairs at once."""
    box_a, box_b = np.asarray(box_a), np.asarray(box_b)
    min_max = np.array([box_a[0:3] + box_a[3:6] / 2, box_b[0:3] + box_b[3:6] / 2]).min(0)
    max_min = np.array([box_a[0:3] - box_a[3:6] / 2, box_b[0:3] - box_b[3:6] / 2]).max(0)
    if not ((min_max > max_min).all()):
        return 0.0
    intersection = (min_max - max_min).prod()
    union = box_a[3:6].prod() + box_b[3:6].prod() - intersection
    return 1.0 * intersection / union


def get_iou(bb1, bb2):
    """3DOVDet_tools/utils/evaluation/eval_det.py:63-77: ``calc_iou`` clamped to [0, 1] -- the tools' default
    ``get_iou_func`` (axis-aligned evaluation of the lifted pseudo-label boxes)."""
    return min(max(calc_iou(bb1, bb2), 0), 1)


def _eval_packed_aabb(pred_all, gt_all, ovthresh, use_07_metric):
    """eval_det with get_iou_func=get_iou: boxes are 6-vectors (centre, lengths), fp64."""
    imgs = list(dict.fromkeys(list(pred_all.keys()) + list(gt_all.keys())))
    classes = list(dict.fromkeys([cl for img in pred_all for cl, _, _ in pred_all[img]] + [cl for img in gt_all for cl, _ in gt_all[img]]))
    cidx = {c: i for i, c in enumerate(classes)}
    S = len(imgs)
    K = max([len(pred_all.get(i, [])) for i in imgs] + [1])
    G = max([len(gt_all.get(i, [])) for i in imgs] + [1])
    det = np.zeros((S, K, 6), np.float64); score = np.zeros((S, K), np.float32); dcls = np.zeros((S, K), np.int32)
    keep = np.zeros((S, K), np.uint8); gtb = np.zeros((S, G, 6), np.float64); glab = np.zeros((S, G), np.int64); gpres = np.zeros((S, G), np.uint8)
    for si, img in enumerate(imgs):
        for k, (cl, box, sc) in enumerate(pred_all.get(img, [])):
            det[si, k] = np.asarray(box, np.float64)[:6]; score[si, k] = sc; dcls[si, k] = cidx[cl]; keep[si, k] = 1
        for g, (cl, box) in enumerate(gt_all.get(img, [])):
            gtb[si, g] = np.asarray(box, np.float64)[:6]; glab[si, g] = cidx[cl]; gpres[si, g] = 1
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.as_tensor(a, device=dev)
    dd, gg, sc_d, dc_d, kp_d, gl_d, gp_d = t(det), t(gtb), t(score), t(dcls), t(keep), t(glab), t(gpres)
    Cn = len(classes)
    iou = torch.empty((S, K, G), dtype=torch.float64, device=dev)
    rs = torch.empty((Cn, S * K), dtype=torch.float32, device=dev)
    rt = torch.empty((Cn, S * K), dtype=torch.uint8, device=dev)
    npos = torch.zeros((Cn,), dtype=torch.int64, device=dev)
    thr = np.asarray([ovthresh], np.float64)
    L = C.lib()
    with C.on_device(dev):
        C.check(L.ovdet_aabb_iou_f64(dd.data_ptr(), gg.data_ptr(), None, None, S, K, G, iou.data_ptr(), C.stream(dev)))
        C.check(L.ovdet_ap_match_iou(iou.data_ptr(), None, sc_d.data_ptr(), kp_d.data_ptr(), dc_d.data_ptr(), gl_d.data_ptr(), gp_d.data_ptr(),
                                     S, K, G, Cn, thr.ctypes.data, 1, rs.data_ptr(), rt.data_ptr(), npos.data_ptr(), C.stream(dev)))
    ap, recall, ndet, rec, prec = ap_reduce(rs, rt, npos, 1, use_07_metric, curves=True)
    ap, ndet, rec, prec = ap.cpu().numpy(), ndet.cpu().numpy(), rec.cpu().numpy(), prec.cpu().numpy()
    pred_classes = set(cl for img in pred_all for cl, _, _ in pred_all[img])
    out_rec, out_prec, out_ap = {}, {}, {}
    for ci, cl in enumerate(classes):
        if cl in pred_classes:
            n = int(ndet[ci])
            out_rec[cl], out_prec[cl], out_ap[cl] = rec[0, ci, :n].copy(), prec[0, ci, :n].copy(), float(ap[0, ci])
        else:
            out_rec[cl], out_prec[cl], out_ap[cl] = 0, 0, 0
    return out_rec, out_prec, out_ap


def get_iou_obb(bb1, bb2):
    """utils/eval_det.py:57-59."""
    from .box_util import box3d_iou
    return box3d_iou(bb1, bb2)[0]


def _iou_path(get_iou_func):
    """The two IoU definitions the reference evaluates with: ``get_iou_obb`` (exact rotated IoU of 8-corner boxes, the
    default of utils/eval_det.py) and ``get_iou`` / ``calc_iou`` (axis-aligned IoU of (centre, lengths) boxes, the default
    of 3DOVDet_tools/utils/evaluation/eval_det.py:86).  Arbitrary Python callables cannot run on the device."""
    if get_iou_func is None or get_iou_func is get_iou_obb:
        return _eval_packed
    if get_iou_func is get_iou or get_iou_func is calc_iou:
        return _eval_packed_aabb
    raise NotImplementedError("get_iou_func must be get_iou_obb (rotated) or get_iou / calc_iou (axis-aligned) of this module")


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False, get_iou_func=get_iou_obb):
    """utils/eval_det.py:66-155: pred {img: [(bbox, score)]}, gt {img: [bbox]} ->
    (rec, prec, ap) with rec/prec sorted by descending score."""
    pred_all = {img: [(0, b, s) for b, s in lst] for img, lst in pred.items()}
    gt_all = {img: [(0, b) for b in lst] for img, lst in gt.items()}
    if not any(len(v) for v in pred_all.values()):
        return np.zeros(0), np.zeros(0), voc_ap(np.zeros(0), np.zeros(0), use_07_metric)
    rec, prec, ap = _iou_path(get_iou_func)(pred_all, gt_all, ovthresh, use_07_metric)
    return rec[0], prec[0], ap[0]


def eval_det(pred_all, gt_all, ovthresh=0.25, use_07_metric=False, get_iou_func=get_iou_obb):
    """utils/eval_det.py:164-208 / :214-272 -> ({cls: rec}, {cls: prec}, {cls: ap})."""
    return _iou_path(get_iou_func)(pred_all, gt_all, ovthresh, use_07_metric)


eval_det_multiprocessing = eval_det  # the per-class Pool(10) of :253-261 is one kernel launch here
