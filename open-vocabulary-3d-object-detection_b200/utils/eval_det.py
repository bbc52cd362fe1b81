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
    import torch.distributed as dist
    C.require_cuda(rec_score)
    dev = rec_score.device
    Cn, N = rec_score.shape
    world = dist.get_world_size() if distributed else 1
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
    # ---- several ranks: per-rank TP lists (keys | bits | count) travel in ONE all-gather, the bucket histogram and the
    # npos / nvalid sums in ONE all-reduce; everything else is local.  `cap` is the per-rank list capacity.
    row = cap * 5 + 8
    lists = torch.zeros((Cn, row), dtype=torch.uint8, device=dev)
    kbuf = torch.empty((Cn, cap), dtype=torch.int32, device=dev)
    bbuf = torch.empty((Cn, cap), dtype=torch.uint8, device=dev)
    cnt = torch.empty((Cn,), dtype=torch.int32, device=dev)
    sums = torch.zeros((2 * Cn,), dtype=torch.int64, device=dev)       # npos | nvalid
    sums[:Cn] = npos.to(device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        C.check(L.ovdet_apc_collect(C.ptr(rec_score), C.ptr(rec_tp), Cn, N, cap, C.ptr(kbuf), C.ptr(bbuf), C.ptr(cnt),
                                    sums[Cn:].data_ptr(), st))
        lists[:, :cap * 4] = kbuf.view(torch.uint8)
        lists[:, cap * 4:cap * 5] = bbuf
        lists[:, cap * 5:cap * 5 + 4] = cnt.view(torch.uint8).reshape(Cn, 4)
        glists = torch.empty((world * Cn, row), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(glists, lists)
        g = glists.view(world, Cn, row)
        gcnt = g[:, :, cap * 5:cap * 5 + 4].contiguous().view(torch.int32)                 # [W, C, 1]
        key2 = torch.full((Cn, cap_total), -1, dtype=torch.int32, device=dev)            # 0xFFFFFFFF = empty slot
        bits2 = torch.zeros((Cn, cap_total), dtype=torch.uint8, device=dev)
        key2[:, :world * cap] = g[:, :, :cap * 4].permute(1, 0, 2).contiguous().view(torch.int32).reshape(Cn, world * cap)
        bits2[:, :world * cap] = g[:, :, cap * 4:cap * 5].permute(1, 0, 2).reshape(Cn, world * cap)
        C.check(L.ovdet_apc_sort(C.ptr(key2), C.ptr(bits2), Cn, cap_total, st))
        hist = torch.empty((Cn, cap_total + 1), dtype=torch.int32, device=dev)
        C.check(L.ovdet_apc_hist(C.ptr(rec_score), Cn, N, C.ptr(key2), cap_total, C.ptr(hist), st))
        fused = torch.cat([hist.reshape(-1).to(torch.int64), sums])
        dist.all_reduce(fused, op=dist.ReduceOp.SUM)
        hist_g = fused[:Cn * (cap_total + 1)].to(torch.int32)
        npos_g = fused[Cn * (cap_total + 1):Cn * (cap_total + 1) + Cn].contiguous()
        nvalid_g = fused[Cn * (cap_total + 1) + Cn:].contiguous()
        res = torch.empty((2 * nthr * Cn + 2 + Cn,), dtype=torch.float64, device=dev)
        ndet = torch.empty((Cn,), dtype=torch.int64, device=dev)
        zero_cnt = torch.zeros((Cn,), dtype=torch.int32, device=dev)   # overflow is judged from the gathered counts
        C.check(L.ovdet_apc_final(C.ptr(bits2), C.ptr(zero_cnt), C.ptr(hist_g), C.ptr(npos_g), C.ptr(nvalid_g), Cn,
                                  cap_total, nthr, int(bool(use_07_metric)), res.data_ptr(),
                                  res.data_ptr() + 8 * nthr * Cn, C.ptr(ndet), None, st))
        k2 = 2 * nthr * Cn
        res[k2] = (gcnt > cap).sum().to(torch.float64)
        res[k2 + 1:k2 + 1 + Cn] = ndet.to(torch.float64)
        res[k2 + 1 + Cn] = gcnt.max().to(torch.float64)
    return res


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


def get_iou_obb(bb1, bb2):
    """utils/eval_det.py:57-59."""
    from .box_util import box3d_iou
    return box3d_iou(bb1, bb2)[0]


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False, get_iou_func=get_iou_obb):
    """utils/eval_det.py:66-155: pred {img: [(bbox, score)]}, gt {img: [bbox]} ->
    (rec, prec, ap) with rec/prec sorted by descending score."""
    if get_iou_func is not get_iou_obb:
        raise NotImplementedError("only get_iou_obb (exact rotated IoU) is built on the GPU path")
    pred_all = {img: [(0, b, s) for b, s in lst] for img, lst in pred.items()}
    gt_all = {img: [(0, b) for b in lst] for img, lst in gt.items()}
    if not any(len(v) for v in pred_all.values()):
        return np.zeros(0), np.zeros(0), voc_ap(np.zeros(0), np.zeros(0), use_07_metric)
    rec, prec, ap = _eval_packed(pred_all, gt_all, ovthresh, use_07_metric)
    return rec[0], prec[0], ap[0]


def eval_det(pred_all, gt_all, ovthresh=0.25, use_07_metric=False, get_iou_func=get_iou_obb):
    """utils/eval_det.py:164-208 / :214-272 -> ({cls: rec}, {cls: prec}, {cls: ap})."""
    if get_iou_func is not None and get_iou_func is not get_iou_obb:
        raise NotImplementedError("only get_iou_obb (exact rotated IoU) is built on the GPU path")
    return _eval_packed(pred_all, gt_all, ovthresh, use_07_metric)


eval_det_multiprocessing = eval_det  # the per-class Pool(10) of :253-261 is one kernel launch here
