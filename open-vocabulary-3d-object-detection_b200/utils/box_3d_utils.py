"""Drop-in for ``3DOVDet_tools/utils/box_3d_utils.py`` and the per-scene
"NMS + IoU filtering" of ``3DOVDet_tools/scannet/lift_boxes.py:139-166`` (the
workload of BASELINE config 5), batched over scenes on the GPU."""
import numpy as np
import torch

from .. import _capi as C
from .nms import nms_batch


def box_3d_iou(box_q, box_k, typ="vv", eps=1e-5):
    """box_3d_utils.py:3-57 (= utils/label_formatter.py:10-64): AABB IoU of one box
    against N, fp64, ``+eps`` in the denominator.  Accepts numpy or torch; a thin
    elementwise expression (its hot use, the pool match, lives in the
    pseudo-filter kernel)."""
    is_np = isinstance(box_q, np.ndarray)
    q = torch.as_tensor(box_q, dtype=torch.float64)
    k = torch.as_tensor(box_k, dtype=torch.float64)
    q = q[None, :]
    if typ == "vv":
        ql, qh, kl, kh = q[:, 0:3], q[:, 3:6], k[:, 0:3], k[:, 3:6]
    else:
        ql, qh = q[:, 0:3] - q[:, 3:6] / 2, q[:, 0:3] + q[:, 3:6] / 2
        kl, kh = k[:, 0:3] - k[:, 3:6] / 2, k[:, 0:3] + k[:, 3:6] / 2
    qv = (qh[:, 0] - ql[:, 0]) * (qh[:, 1] - ql[:, 1]) * (qh[:, 2] - ql[:, 2])
    kv = (kh[:, 0] - kl[:, 0]) * (kh[:, 1] - kl[:, 1]) * (kh[:, 2] - kl[:, 2])
    e = (torch.minimum(qh, kh) - torch.maximum(ql, kl)).clamp(min=0)
    inter = e[:, 0] * e[:, 1] * e[:, 2]
    iou = inter / (qv + kv - inter + eps)
    return iou.numpy() if is_np else iou


def nms_3d_faster(boxes, overlap_threshold, old_type=False, eps=1e-8, use_size=False, use_size_score=False,
                  class_wise=False, size_typ=None, lhs=False):
    """box_3d_utils.py:60-120.  Returns ``boxes[pick]``.  As in the reference,
    ``use_size_score`` multiplies the caller's score column in place (:78-79); ``lhs`` re-picks the better-scoring
    half of the boxes each pick suppresses (:113-116)."""
    assert size_typ in [None, "Volume", "Area"]
    boxes = np.asarray(boxes)
    if boxes.shape[0] == 0:
        return boxes
    work = np.array(boxes[:, :8], dtype=np.float64, copy=True)
    if size_typ is not None:
        size = boxes[:, 8] if size_typ == "Volume" else boxes[:, 9]
        if use_size:
            work[:, 6] = size
        elif use_size_score:
            boxes[:, 6] *= size
            work[:, 6] = boxes[:, 6]
    b = torch.as_tensor(np.ascontiguousarray(work), device="cuda")[None]
    _, order, npick = nms_batch(b, overlap_threshold, old_type, 3, class_wise, vol_eps=eps, lhs=lhs)
    n = int(npick[0].item())
    return boxes[order[0, :n].cpu().numpy().astype(np.int64)]


def vv2cs(box):
    """box_3d_utils.py:122-129 (in place)."""
    box[:, 3:6] -= box[:, :3]
    box[:, :3] += box[:, 3:6] / 2
    return box


def cs2vv(box):
    """box_3d_utils.py:131-134 (in place)."""
    box[:, :3] -= box[:, 3:6] / 2
    box[:, 3:6] += box[:, :3]
    return box


def lift_filter_batch(boxes, pool, nboxes=None, npool=None, nms_thresh=0.7, match_thresh=0.3, size_nms_thresh=0.0):
    """Batched lift_boxes.py:139-166.  boxes CUDA fp64 [S,P,8] (x1..z2, score,
    label), pool CUDA fp64 [S,M,6].  Returns dict of device tensors:
    nms1_keep [S,P] u8, label [S,M] (-100 = unmatched), score [S,M] (matched box
    score), keep [S,M] u8 (pool boxes surviving the size-scored NMS)."""
    C.require_cuda(boxes, pool)
    b = boxes.detach().to(torch.float64).contiguous()
    pl = pool.detach().to(torch.float64).contiguous()
    S, P, _ = b.shape
    M = pl.shape[1]
    dev = b.device
    k1 = torch.empty((S, P), dtype=torch.uint8, device=dev)
    lab = torch.empty((S, M), dtype=torch.float64, device=dev)
    sc = torch.empty((S, M), dtype=torch.float64, device=dev)
    keep = torch.empty((S, M), dtype=torch.uint8, device=dev)
    nb = None if nboxes is None else torch.as_tensor(nboxes).to(device=dev, dtype=torch.int32).contiguous()
    npl = None if npool is None else torch.as_tensor(npool).to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_pseudo_filter_f64(C.ptr(b), C.ptr(pl), C.ptr(nb), C.ptr(npl), S, P, M, float(nms_thresh),
                                                float(match_thresh), float(size_nms_thresh), C.ptr(k1), C.ptr(lab),
                                                C.ptr(sc), C.ptr(keep), C.stream(dev)))
    return {"nms1_keep": k1, "label": lab, "score": sc, "keep": keep}


def lift_sweep(boxes, pool, nboxes=None, npool=None, nms_thresh=0.7, match_thresh=0.3, size_nms_thresh=0.0,
               distributed=False, chunk=8192, out=None):
    """The pseudo-label sweep of BASELINE config 5 over many scenes: the per-scene pipeline of lift_boxes.py:139-166
    (its `__main__` maps it over all scans with ``mp.Pool``, :177-181) in chunks of ``chunk`` scenes per launch.
    ``boxes`` / ``pool`` hold THIS rank's scenes (use ``dist.shard_range`` to cut the scan list; scenes are independent,
    so there is no data-path collective).  Returns ``(result, total_kept)``: ``result`` as in ``lift_filter_batch``
    (written into ``out`` when given), ``total_kept`` = number of pseudo-label boxes kept over ALL ranks -- the one
    cross-rank value of the sweep (the "Acquired N boxes" / per-scan counts summed at :181; all-reduced when
    ``distributed``)."""
    from ..dist import all_reduce_count
    C.require_cuda(boxes, pool)
    S, P, M = boxes.shape[0], boxes.shape[1], pool.shape[1]
    dev = boxes.device
    if out is None:
        out = {"nms1_keep": torch.empty((S, P), dtype=torch.uint8, device=dev), "label": torch.empty((S, M), dtype=torch.float64, device=dev),
               "score": torch.empty((S, M), dtype=torch.float64, device=dev), "keep": torch.empty((S, M), dtype=torch.uint8, device=dev)}
    b = boxes.detach().to(torch.float64).contiguous()
    pl = pool.detach().to(torch.float64).contiguous()
    nb = None if nboxes is None else torch.as_tensor(nboxes).to(device=dev, dtype=torch.int32).contiguous()
    npl = None if npool is None else torch.as_tensor(npool).to(device=dev, dtype=torch.int32).contiguous()
    L, st = C.lib(), C.stream(dev)
    with C.on_device(dev):
        for lo in range(0, S, chunk):
            n = min(chunk, S - lo)
            C.check(L.ovdet_pseudo_filter_f64(b[lo:].data_ptr(), pl[lo:].data_ptr(), None if nb is None else nb[lo:].data_ptr(),
                                              None if npl is None else npl[lo:].data_ptr(), n, P, M, float(nms_thresh), float(match_thresh),
                                              float(size_nms_thresh), out["nms1_keep"][lo:].data_ptr(), out["label"][lo:].data_ptr(),
                                              out["score"][lo:].data_ptr(), out["keep"][lo:].data_ptr(), st))
    kept = out["keep"].sum(dtype=torch.int64)      # stays on the device until the one count exchange
    if distributed:
        import torch.distributed as tdist
        tdist.all_reduce(kept, op=tdist.ReduceOp.SUM)
        return out, int(kept.item())
    return out, all_reduce_count(int(kept.item()), dev)


def lift_filter_scene(boxes, box_pool, nms_thresh=0.7, match_thresh=0.3, size_nms_thresh=0.0):
    """One scene, numpy in / numpy out, rows [x1..z2, score*volume, label, volume, area]
    ordered by the final NMS pick order -- what lift_boxes.py:159-165 leaves in `boxes`."""
    boxes = np.asarray(boxes, np.float64)
    pool = np.asarray(box_pool, np.float64)
    if boxes.shape[0] == 0 or pool.shape[0] == 0:
        return np.zeros((0, 10))
    r = lift_filter_batch(torch.as_tensor(boxes[None, :, :8].copy(), device="cuda"),
                          torch.as_tensor(pool[None, :, :6].copy(), device="cuda"),
                          nms_thresh=nms_thresh, match_thresh=match_thresh, size_nms_thresh=size_nms_thresh)
    keep = r["keep"][0].cpu().numpy().astype(bool)
    lab = r["label"][0].cpu().numpy()
    sc = r["score"][0].cpu().numpy()
    scale = pool[:, 3:6] - pool[:, 0:3]
    vol = np.prod(scale, axis=-1)
    area = 2 * np.sum(scale * np.roll(scale, 1, axis=-1), axis=-1)
    rows = np.concatenate([pool[:, :6], np.stack([sc * vol, lab, vol, area], 1)], -1)[keep]
    return rows[np.argsort(-rows[:, 6], kind="stable")]


def save_lifted_boxes(path, boxes):
    """The on-disk form of lift_boxes.py:167-169: vertex-vertex rows [x1..z2, score, label, volume, area] ->
    centre-size rows with score and label swapped, i.e. ``<scan>_bbox.npy`` = fp64 [N,10] =
    (cx, cy, cz, sx, sy, sz, label, score, volume, area).  Returns the number of boxes written."""
    b = np.asarray(boxes, np.float64)
    if b.shape[0]:
        b = vv2cs(b.copy())
        b[:, [6, 7]] = b[:, [7, 6]]
    np.save(path, b)
    return int(b.shape[0])


def lift_and_save_scene(path, boxes, box_pool, nms_thresh=0.7, match_thresh=0.3, size_nms_thresh=0.0):
    """NMS -> pool match -> size-scored NMS (lift_boxes.py:139-165) and the file of :167-169 for one scene."""
    return save_lifted_boxes(path, lift_filter_scene(boxes, box_pool, nms_thresh, match_thresh, size_nms_thresh))
