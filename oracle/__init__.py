"""CPU oracle for the detection-geometry / open-vocabulary matching path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker or the CPU baseline -- never as the thing
measured or shipped.  The product (``open-vocabulary-3d-object-detection_b200``)
does not import it and raises when its CUDA library is missing.

What is here
    * ``ovdet_oracle.c``   plain-C restatement of the reference algorithms
      (Sutherland-Hodgman BEV clip, GIoU in both reference precisions, exact
      fp64 ``box3d_iou``, greedy NMS, VOC AP, ``eval_det_cls``), loaded by ctypes.
    * this module         numpy/torch glue that restates the reference's Python
      call surface on top of it, each function citing reference ``file:line``.
    * ``_ref/``           (git-ignored) the reference's own Cython extension
      compiled from /root/reference by ``oracle/build.py`` -- the real hot loop,
      used to validate the restatement and as the ``"reference"`` CPU baseline.

Third-party arithmetic on the path whose source is not in the reference tree
(no version pinned by the reference, versions are the ones
installed in this image): scipy 1.18.1 ``optimize.linear_sum_assignment``
(criterion.py:79), ``spatial.ConvexHull`` (box_util.py:96; restated as a
shoelace area -- the clip of two convex polygons is convex), numpy 2.3.5
``np.dot``/``np.argsort``, torch 2.11 ``cdist``/``softmax``.

Parity status: PINNED -- the restatement is checked against outputs of the
reference itself generated in the build container (tests/golden/*.npz, made by
tests/golden/make_golden.py) and, when oracle/_ref exists, against the compiled
reference Cython loop directly.
"""
import ctypes
import importlib.util
import os
from collections import OrderedDict

import numpy as np

from . import build as _build

_c_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_c_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_c_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = _build.build_oracle()
        L = ctypes.CDLL(so)
        L.oracle_polygon_clip_f64.restype = ctypes.c_int
        L.oracle_polygon_clip_f64.argtypes = [_c_f64, ctypes.c_int, _c_f64, ctypes.c_int, _c_f64]
        L.oracle_polygon_clip_f32.restype = ctypes.c_int
        L.oracle_polygon_clip_f32.argtypes = [_c_f32, ctypes.c_int, _c_f32, ctypes.c_int, _c_f32]
        L.oracle_poly_area_f64.restype = ctypes.c_double
        L.oracle_poly_area_f64.argtypes = [_c_f64, ctypes.c_int]
        L.oracle_box_intersection.restype = None
        L.oracle_box_intersection.argtypes = [_c_f32, _c_f32, _c_f32, _c_i32, _c_f32, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_giou3d.restype = None
        L.oracle_giou3d.argtypes = [_c_f32, _c_f32, ctypes.c_void_p] + [ctypes.c_int] * 8 + [_c_f32]
        L.oracle_box3d_iou.restype = ctypes.c_double
        L.oracle_box3d_iou.argtypes = [_c_f64, _c_f64, ctypes.POINTER(ctypes.c_double)]
        L.oracle_box3d_iou_matrix.restype = None
        L.oracle_box3d_iou_matrix.argtypes = [_c_f32, ctypes.c_int, _c_f32, ctypes.c_int, _c_f64]
        L.oracle_nms.restype = ctypes.c_int
        L.oracle_nms.argtypes = [_c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_double, ctypes.c_int, ctypes.c_double, _c_i32]
        L.oracle_voc_ap.restype = ctypes.c_double
        L.oracle_voc_ap.argtypes = [_c_f64, _c_f64, ctypes.c_int, ctypes.c_int]
        L.oracle_eval_det_cls.restype = ctypes.c_double
        L.oracle_eval_det_cls.argtypes = [_c_i32, _c_f32, _c_f32, ctypes.c_int, _c_i32, _c_f32, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_int, _c_i32, _c_f64, _c_f64, _c_f64]
        L.oracle_box_3d_iou_aabb.restype = None
        L.oracle_box_3d_iou_aabb.argtypes = [_c_f64, _c_f64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, _c_f64]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------- #
# the compiled reference Cython extension (oracle/_ref), when present
# --------------------------------------------------------------------------- #
_REF_MOD = None


def ref_box_intersection():
    """The reference's own compiled ``box_intersection`` (utils/box_intersection.pyx:166),
    or None when oracle/_ref has not been built."""
    global _REF_MOD
    if _REF_MOD is None:
        so = _build.build_ref()
        if so is None or not os.path.exists(so):
            return None
        spec = importlib.util.spec_from_file_location("box_intersection", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _REF_MOD = mod
    return _REF_MOD.box_intersection


# --------------------------------------------------------------------------- #
# polygon clip / area
# --------------------------------------------------------------------------- #
def polygon_clip(subject, clip, dtype=np.float64):
    """utils/box_util.py:34-81 / utils/box_intersection.pyx:27-70.  Returns an
    [n,2] array (n may be 0; the numpy version returns None there)."""
    s = np.ascontiguousarray(subject, dtype=dtype).reshape(-1, 2)
    c = np.ascontiguousarray(clip, dtype=dtype).reshape(-1, 2)
    out = np.zeros((32, 2), dtype=dtype)
    fn = lib().oracle_polygon_clip_f64 if dtype == np.float64 else lib().oracle_polygon_clip_f32
    n = fn(s, s.shape[0], c, c.shape[0], out)
    return out[:n].copy()


def poly_area(x, y):
    """utils/box_util.py:84-86."""
    p = np.ascontiguousarray(np.stack([x, y], -1), dtype=np.float64)
    return lib().oracle_poly_area_f64(p, p.shape[0])


def box_intersection(rect1, rect2, non_rot_inter_areas, nums_k2, inter_areas, approximate, k2_loop=4):
    """utils/box_intersection.pyx:166-198, in place.  ``k2_loop`` = rect2.shape[2]
    of the shipped code (=4, the K2 bug); pass K2 for the intended behaviour."""
    B, K1 = rect1.shape[0], rect1.shape[1]
    K2 = rect2.shape[1]
    assert inter_areas.dtype == np.float32 and inter_areas.flags.c_contiguous
    lib().oracle_box_intersection(np.ascontiguousarray(rect1, np.float32), np.ascontiguousarray(rect2, np.float32),
                                  np.ascontiguousarray(non_rot_inter_areas, np.float32),
                                  np.ascontiguousarray(nums_k2, np.int32), inter_areas, int(bool(approximate)),
                                  B, K1, K2, int(k2_loop))


# --------------------------------------------------------------------------- #
# GIoU
# --------------------------------------------------------------------------- #
def _np(x, dtype):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def enclosing_hull_vol(c1, c2):
    """Convex-hull enclosing volume of the 16 corners of a pair
    (utils/box_ops3d.py:566-567: ConvexHull(np.vstack([c1, c2])).volume)."""
    from scipy.spatial import ConvexHull
    return ConvexHull(np.vstack([np.asarray(c1, np.float64), np.asarray(c2, np.float64)])).volume


def generalized_box3d_iou(corners1, corners2, nums_k2, rotated_boxes=True, return_inter_vols_only=False,
                          mode="tensor", prefilter=True, k2_cap=None, enclosing="aabb"):
    """utils/box_util.py:717-737.  ``mode``:
      "tensor"  the fp32 torch path (:517-618; what runs when needs_grad or no Cython)
      "cython"  the Cython-backed path (:624-714); pass ``k2_cap=4`` for the
                as-shipped K2 bug (box_intersection.pyx:180), None for the intent.
    ``prefilter=False`` lifts the axis-aligned skip (:587-588 / pyx:189).
    ``enclosing="hull"`` swaps the AABB enclosing volume (:466-514) for the convex
    hull of box_ops3d.py:533-571 (pairwise scipy call -- small inputs only).
    Returns float32 ndarray [B,K1,K2]."""
    c1 = _np(corners1, np.float32)
    c2 = _np(corners2, np.float32)
    B, K1 = c1.shape[0], c1.shape[1]
    K2 = c2.shape[1]
    out = np.zeros((B, K1, K2), np.float32)
    nk = None if nums_k2 is None else _np(nums_k2, np.int64)
    ptr = None if nk is None else nk.ctypes.data_as(ctypes.c_void_p)
    cap = 0 if k2_cap is None else int(k2_cap)
    if enclosing == "aabb":
        lib().oracle_giou3d(c1, c2, ptr, B, K1, K2, int(bool(rotated_boxes)), int(bool(return_inter_vols_only)),
                            1 if mode == "cython" else 0, int(bool(prefilter)), cap, out)
        return out
    assert enclosing == "hull"
    # utils/box_ops3d.py:475-530 (generalized_box3d_iou_convex_hull_nondiff_tensor; the file itself cannot be
    # imported): the enclosing volume starts as the AABB one (:505) and, for rotated boxes, is
    # replaced by scipy's ConvexHull(np.vstack([c1, c2])).volume only where the intersection volume is > 0 and
    # the column is valid (:467,:514-519); stored as fp32.
    import ctypes as _ct
    inter = np.zeros((B, K1, K2), np.float32)
    lib().oracle_giou3d(c1, c2, ptr, B, K1, K2, int(bool(rotated_boxes)), 1,
                        1 if mode == "cython" else 0, int(bool(prefilter)), cap, inter)
    if return_inter_vols_only:
        return inter
    EPS = np.float32(1e-8)
    for b in range(B):
        for i in range(K1):
            v1 = max(_vol_f32(c1[b, i]), EPS)
            for j in range(K2):
                valid = nk is None or j < nk[b]
                v2 = max(_vol_f32(c2[b, j]), EPS)
                sv = np.float32(v1 + v2)
                both = np.vstack([c1[b, i], c2[b, j]])
                d = np.abs(both.max(0) - both.min(0)).astype(np.float32)
                encl = np.float32(np.float32(d[0] * d[1]) * d[2])
                if rotated_boxes and valid and inter[b, i, j] > 0:
                    encl = np.float32(enclosing_hull_vol(c1[b, i], c2[b, j]))
                uni = max(np.float32(sv - inter[b, i, j]), EPS)
                good = np.float32((encl > 2 * EPS) and (sv > 4 * EPS))
                with np.errstate(all="ignore"):
                    g = np.float32(np.float32(inter[b, i, j] / uni) + (-(np.float32(1) - np.float32(uni / encl))))
                    g = np.float32(g * good)
                if nk is not None and not valid:
                    g = np.float32(g * 0)
                out[b, i, j] = g
    return out


def _nondegenerate(a, b):
    p = np.vstack([a, b]).astype(np.float64)
    return np.linalg.matrix_rank(p - p.mean(0), tol=1e-9) == 3


def _vol_f32(c):
    c = c.astype(np.float32)
    e = []
    for a, b in ((0, 1), (1, 2), (0, 4)):
        d = c[a] - c[b]
        s = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        e.append(np.sqrt(max(s, np.float32(1e-6))))
    return np.float32(np.float32(e[0] * e[1]) * e[2])


def generalized_box3d_iou_ref_cython(corners1, corners2, nums_k2, rotated_boxes=True,
                                     return_inter_vols_only=False, lift_k2_cap=False):
    """The Cython-backed path (utils/box_util.py:624-714) with the torch glue
    restated in torch-CPU ops and the hot loop delegated to the REAL compiled
    reference extension in oracle/_ref.  ``lift_k2_cap`` works around the shipped
    ``K2 = rect2.shape[2]`` bug without touching the reference: the unmodified
    extension is called once per slice of 4 GT columns.  Used as the
    ``"reference"`` CPU baseline and to validate ``oracle_giou3d``."""
    import torch
    fn = ref_box_intersection()
    assert fn is not None, "oracle/_ref not built"
    c1 = torch.as_tensor(corners1, dtype=torch.float32).cpu()
    c2 = torch.as_tensor(corners2, dtype=torch.float32).cpu()
    nk = torch.as_tensor(nums_k2).cpu()
    B, K1, K2 = c1.shape[0], c1.shape[1], c2.shape[1]
    height = (torch.min(c1[:, :, 0, 1][:, :, None], c2[:, :, 0, 1][:, None, :])
              - torch.max(c1[:, :, 4, 1][:, :, None], c2[:, :, 4, 1][:, None, :])).clamp(min=0)
    r1 = c1[:, :, [3, 2, 1, 0]][..., [0, 2]].contiguous()
    r2 = c2[:, :, [3, 2, 1, 0]][..., [0, 2]].contiguous()
    wh = (torch.min(r1[:, :, 3][:, :, None, :], r2[:, :, 3][:, None, :, :])
          - torch.max(r1[:, :, 1][:, :, None, :], r2[:, :, 1][:, None, :, :])).clamp(min=0)
    col = torch.arange(K2)[None, None, :]
    valid = (col < nk[:, None, None])
    non_rot = (wh[..., 0] * wh[..., 1]) * valid
    both = torch.cat([c1[:, :, None].expand(B, K1, K2, 8, 3), c2[:, None].expand(B, K1, K2, 8, 3)], 3)
    ext = (both.max(3).values - both.min(3).values).abs()
    encl = ext[..., 0] * ext[..., 1] * ext[..., 2]

    def vol(c):
        def edge(a, b):
            return torch.sqrt((c[:, :, a] - c[:, :, b]).pow(2).sum(-1).clamp(min=1e-6))
        return (edge(0, 1) * edge(1, 2) * edge(0, 4)).clamp(min=1e-8)

    sum_vols = vol(c1)[:, :, None] + vol(c2)[:, None, :]
    good = (encl > 2e-8) * (sum_vols > 4e-8)
    if rotated_boxes:
        areas = np.zeros((B, K1, K2), np.float32)
        r1n, r2n = r1.numpy(), r2.numpy()
        nkn = nk.numpy().astype(np.int32)
        nrn = non_rot.numpy().astype(np.float32)
        if not lift_k2_cap:
            fn(r1n, r2n, nrn, nkn, areas, True)
        else:
            for j0 in range(0, K2, 4):
                j1 = min(j0 + 4, K2)
                w = j1 - j0
                r2s = np.zeros((B, 4, 4, 2), np.float32)
                r2s[:, :w] = r2n[:, j0:j1]
                sub = np.zeros((B, K1, 4), np.float32)
                nrs = np.zeros((B, K1, 4), np.float32)
                nrs[:, :, :w] = nrn[:, :, j0:j1]
                fn(r1n, r2s, nrs, np.clip(nkn - j0, 0, w).astype(np.int32), sub, True)
                areas[:, :, j0:j1] = sub[:, :, :w]
        inter_areas = torch.from_numpy(areas)
    else:
        inter_areas = non_rot
    inter_vols = inter_areas * height
    if return_inter_vols_only:
        return inter_vols.numpy()
    union = (sum_vols - inter_vols).clamp(min=1e-8)
    g = inter_vols / union + (-(1 - union / encl))
    g = g * good
    g = g * valid.float()
    return g.numpy()


# --------------------------------------------------------------------------- #
# exact IoU, NMS, AP
# --------------------------------------------------------------------------- #
def box3d_iou(corners1, corners2):
    """utils/box_util.py:116-141 -> (iou, iou_2d), fp64."""
    a = np.ascontiguousarray(corners1, np.float64)
    b = np.ascontiguousarray(corners2, np.float64)
    i2 = ctypes.c_double(0.0)
    iou = lib().oracle_box3d_iou(a, b, ctypes.byref(i2))
    return iou, i2.value


def box3d_iou_matrix(dets, gts):
    d = np.ascontiguousarray(dets, np.float32).reshape(-1, 8, 3)
    g = np.ascontiguousarray(gts, np.float32).reshape(-1, 8, 3)
    out = np.zeros((d.shape[0], g.shape[0]), np.float64)
    if d.shape[0] and g.shape[0]:
        lib().oracle_box3d_iou_matrix(d, d.shape[0], g, g.shape[0], out)
    return out


def _nms(boxes, thr, old_type, dims, samecls, eps=0.0):
    bx = np.ascontiguousarray(boxes, np.float64)
    K = bx.shape[0]
    pick = np.zeros(max(K, 1), np.int32)
    n = lib().oracle_nms(bx, K, bx.shape[1] if K else 0, dims, int(samecls), float(thr), int(bool(old_type)),
                         float(eps), pick)
    return [int(i) for i in pick[:n]]


def nms_2d_faster(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:43-76."""
    return _nms(boxes, overlap_threshold, old_type, 2, False)


def nms_3d_faster(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:79-117."""
    return _nms(boxes, overlap_threshold, old_type, 3, False)


def nms_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """utils/nms.py:120-162."""
    return _nms(boxes, overlap_threshold, old_type, 3, True)


def tools_nms_3d_faster(boxes, overlap_threshold, old_type=False, eps=1e-8, use_size=False,
                        use_size_score=False, class_wise=False, size_typ=None):
    """3DOVDet_tools/utils/box_3d_utils.py:60-120 (without the ``lhs`` option).
    Returns boxes[pick]; like the reference, ``use_size_score`` multiplies the
    score column of the caller's array in place (:78-79)."""
    boxes = np.asarray(boxes)
    assert size_typ in (None, "Volume", "Area")
    work = np.array(boxes, dtype=np.float64, copy=True)
    if size_typ is not None:
        size = boxes[:, 8] if size_typ == "Volume" else boxes[:, 9]
        if use_size:
            work[:, 6] = size
        elif use_size_score:
            boxes[:, 6] *= size
            work[:, 6] = boxes[:, 6]
    pick = _nms(work[:, :8], overlap_threshold, old_type, 3, class_wise, eps)
    return boxes[np.array(pick, dtype=np.int64)] if len(pick) else boxes[:0]


def tools_nms_3d_faster_lhs(boxes, overlap_threshold, old_type=False, eps=1e-8, class_wise=False, lhs=True):
    """3DOVDet_tools/utils/box_3d_utils.py:60-120 WITH the ``lhs`` option (:113-116), restated in numpy: after each pick
    the better-scoring half of the boxes it suppresses is appended to the picks as well (and removed with the rest).
    Returns the pick indices.  Stable ascending sort on (score, index): only defined for tie-free scores."""
    b = np.asarray(boxes, np.float64)
    vol = (b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1]) * (b[:, 5] - b[:, 2]) + eps
    order = list(np.argsort(b[:, 6], kind="stable"))
    pick = []
    while order:
        i = order[-1]
        rest = order[:-1]
        pick.append(int(i))
        sup = []
        for pos, j in enumerate(rest):
            l = max(0.0, min(b[i, 3], b[j, 3]) - max(b[i, 0], b[j, 0]))
            w = max(0.0, min(b[i, 4], b[j, 4]) - max(b[i, 1], b[j, 1]))
            h = max(0.0, min(b[i, 5], b[j, 5]) - max(b[i, 2], b[j, 2]))
            inter = l * w * h
            o = inter / vol[j] if old_type else inter / (vol[i] + vol[j] - inter)
            if class_wise:
                o = o * (1.0 if b[i, 7] == b[j, 7] else 0.0)
            if o > overlap_threshold:
                sup.append(pos)
        if lhs:
            for count in range(len(sup) // 2):
                pick.append(int(rest[sup[len(sup) - count - 1]]))
        gone = set(sup)
        order = [j for pos, j in enumerate(rest) if pos not in gone]
    return pick


def calc_iou_aabb(box_a, box_b):
    """3DOVDet_tools/utils/evaluation/box_util.py:287-309 + the [0, 1] clamp of get_iou (evaluation/eval_det.py:63-77)."""
    a, b = np.asarray(box_a, np.float64), np.asarray(box_b, np.float64)
    hi = np.minimum(a[0:3] + a[3:6] / 2, b[0:3] + b[3:6] / 2)
    lo = np.maximum(a[0:3] - a[3:6] / 2, b[0:3] - b[3:6] / 2)
    if not (hi > lo).all():
        return 0.0
    d = hi - lo
    inter = d[0] * d[1] * d[2]
    iou = inter / (a[3] * a[4] * a[5] + b[3] * b[4] * b[5] - inter)
    return min(max(iou, 0.0), 1.0)


def eval_det_cls_aabb(pred, gt, ovthresh=0.25, use_07_metric=False):
    """3DOVDet_tools/utils/evaluation/eval_det.py:86-170 with its default get_iou_func=get_iou (axis-aligned boxes as
    (centre, lengths)): pred {img: [(box6, score)]}, gt {img: [box6]} -> (rec, prec, ap).  Python loops: small cases."""
    recs, npos = {}, 0
    for img, lst in gt.items():
        recs[img] = [np.asarray(x, np.float64) for x in lst], [False] * len(lst)
        npos += len(lst)
    ids, conf, bbs = [], [], []
    for img, lst in pred.items():
        for box, score in lst:
            ids.append(img); conf.append(score); bbs.append(np.asarray(box, np.float64))
    order = np.argsort(-np.asarray(conf, np.float64), kind="stable")
    tp, fp = np.zeros(len(ids)), np.zeros(len(ids))
    for d, o in enumerate(order):
        gtb, taken = recs.get(ids[o], ([], []))
        ovmax, jmax = -np.inf, -1
        for j, g in enumerate(gtb):
            iou = calc_iou_aabb(bbs[o], g)
            if iou > ovmax:
                ovmax, jmax = iou, j
        if ovmax > ovthresh and not taken[jmax]:
            tp[d] = 1.0
            taken[jmax] = True
        else:
            fp[d] = 1.0
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / np.maximum(float(npos), np.finfo(np.float64).eps)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def voc_ap(rec, prec, use_07_metric=False):
    """utils/eval_det.py:23-54."""
    r = np.ascontiguousarray(rec, np.float64)
    p = np.ascontiguousarray(prec, np.float64)
    return lib().oracle_voc_ap(r, p, r.shape[0], int(bool(use_07_metric)))


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False):
    """utils/eval_det.py:66-155 with get_iou_func=get_iou_obb.
    pred: {img_id: [(bbox[8,3], score)]}; gt: {img_id: [bbox[8,3]]}."""
    ids = {}
    for k in list(gt.keys()) + list(pred.keys()):
        ids.setdefault(k, len(ids))
    gs, gb = [], []
    for k, lst in gt.items():
        for b in lst:
            gs.append(ids[k]); gb.append(b)
    ds, dsc, db = [], [], []
    for k, lst in pred.items():
        for b, s in lst:
            ds.append(ids[k]); dsc.append(s); db.append(b)
    nd, ng = len(ds), len(gs)
    order_g = np.argsort(np.asarray(gs, np.int32), kind="stable") if ng else np.zeros(0, np.int64)
    gs_a = np.asarray(gs, np.int32)[order_g] if ng else np.zeros(1, np.int32)
    gb_a = np.asarray(gb, np.float32).reshape(-1, 8, 3)[order_g] if ng else np.zeros((1, 8, 3), np.float32)
    ds_a = np.asarray(ds, np.int32) if nd else np.zeros(1, np.int32)
    dsc_a = np.asarray(dsc, np.float32) if nd else np.zeros(1, np.float32)
    db_a = np.asarray(db, np.float32).reshape(-1, 8, 3) if nd else np.zeros((1, 8, 3), np.float32)
    order = np.zeros(max(nd, 1), np.int32)
    tp = np.zeros(max(nd, 1), np.float64)
    rec = np.zeros(max(nd, 1), np.float64)
    prec = np.zeros(max(nd, 1), np.float64)
    ap = lib().oracle_eval_det_cls(ds_a, dsc_a, np.ascontiguousarray(db_a), nd, gs_a, np.ascontiguousarray(gb_a), ng,
                                   float(ovthresh), int(bool(use_07_metric)), order, tp, rec, prec)
    return rec[:nd], prec[:nd], ap


def _regroup(pred_all, gt_all):
    """utils/eval_det.py:229-248."""
    pred, gt = {}, {}
    for img_id in pred_all.keys():
        for classname, bbox, score in pred_all[img_id]:
            pred.setdefault(classname, {}).setdefault(img_id, [])
            gt.setdefault(classname, {}).setdefault(img_id, [])
            pred[classname][img_id].append((bbox, score))
    for img_id in gt_all.keys():
        for classname, bbox in gt_all[img_id]:
            gt.setdefault(classname, {}).setdefault(img_id, [])
            gt[classname][img_id].append(bbox)
    return pred, gt


def eval_det(pred_all, gt_all, ovthresh=0.25, use_07_metric=False):
    """utils/eval_det.py:214-272 (single process; classes of gt without any
    prediction get rec=prec=ap=0, :262-269)."""
    pred, gt = _regroup(pred_all, gt_all)
    rec, prec, ap = {}, {}, {}
    for classname in gt.keys():
        if classname in pred:
            rec[classname], prec[classname], ap[classname] = eval_det_cls(pred[classname], gt[classname],
                                                                         ovthresh, use_07_metric)
        else:
            rec[classname] = 0
            prec[classname] = 0
            ap[classname] = 0
    return rec, prec, ap


def parse_predictions(predicted_boxes, sem_cls_probs, objectness_probs, config, nonempty_box_mask=None):
    """utils/ap_calculator.py:39-238 for remove_empty_box=False (or a caller-given
    ``nonempty_box_mask``), all four NMS branches (:86-189) and the three output
    formats (:195-236).  ``config`` keys as get_ap_config_dict (:241-269) with
    ``num_semcls`` in place of dataset_config."""
    probs = _np(sem_cls_probs, np.float32)
    cls_prob = probs.max(-1)
    cls = probs.argmax(-1)
    obj = _np(objectness_probs, np.float32)
    corners = _np(predicted_boxes, np.float32)
    B, K = corners.shape[0], corners.shape[1]
    nonempty = np.ones((B, K)) if nonempty_box_mask is None else np.asarray(nonempty_box_mask, np.float64)
    if config.get("no_nms", False):
        pred_mask = nonempty
    else:
        pred_mask = np.zeros((B, K))
        for i in range(B):
            mn = corners[i].min(1).astype(np.float64)
            mx = corners[i].max(1).astype(np.float64)
            sel = np.where(nonempty[i] == 1)[0]
            assert len(sel) > 0
            if not config["use_3d_nms"]:
                bx = np.stack([mn[:, 0], mn[:, 2], mx[:, 0], mx[:, 2], obj[i].astype(np.float64)], 1)
                pick = nms_2d_faster(bx[sel], config["nms_iou"], config["use_old_type_nms"])
            elif not config["cls_nms"]:
                bx = np.concatenate([mn, mx, obj[i].astype(np.float64)[:, None]], 1)
                pick = nms_3d_faster(bx[sel], config["nms_iou"], config["use_old_type_nms"])
            else:
                bx = np.concatenate([mn, mx, obj[i].astype(np.float64)[:, None], cls[i].astype(np.float64)[:, None]], 1)
                pick = nms_3d_faster_samecls(bx[sel], config["nms_iou"], config["use_old_type_nms"])
            assert len(pick) > 0
            pred_mask[i, sel[pick]] = 1
    out = []
    for i in range(B):
        keep = [j for j in range(K) if pred_mask[i, j] == 1 and obj[i, j] > config["conf_thresh"]]
        if config["per_class_proposal"]:
            cur = []
            for ii in range(config["num_semcls"]):
                cur += [(ii, corners[i, j], probs[i, j, ii] * obj[i, j]) for j in keep]
        elif config["use_cls_confidence_only"]:
            cur = [(int(cls[i, j]), corners[i, j], probs[i, j, cls[i, j]]) for j in keep]
        else:
            cur = [(int(cls[i, j]), corners[i, j], obj[i, j]) for j in keep]
        out.append(cur)
    return out, pred_mask


def default_ap_config(num_semcls, **kw):
    """utils/ap_calculator.py:241-269 with remove_empty_box=False."""
    cfg = dict(remove_empty_box=False, use_3d_nms=True, nms_iou=0.25, use_old_type_nms=False, cls_nms=True,
               per_class_proposal=True, use_cls_confidence_only=False, conf_thresh=0.05, no_nms=False,
               num_semcls=num_semcls)
    cfg.update(kw)
    return cfg


def ap_metrics(predicted_boxes, sem_cls_probs, objectness_probs, gt_corners, gt_labels, gt_present,
               num_semcls, ap_iou_thresh=(0.25, 0.5), config=None):
    """APCalculator.step + compute_metrics (utils/ap_calculator.py:324-395) over
    one stack of scenes.  Returns OrderedDict{thr: {"<cls> Average Precision",
    "mAP", "<cls> Recall", "AR"}} and the per-class (rec, prec) of each thr."""
    cfg = config or default_ap_config(num_semcls)
    preds, _ = parse_predictions(predicted_boxes, sem_cls_probs, objectness_probs, cfg)
    gtc = _np(gt_corners, np.float32)
    gtl = _np(gt_labels, np.int64)
    gtp = _np(gt_present, np.float32)
    pred_all, gt_all = {}, {}
    for i in range(gtc.shape[0]):
        gt_all[i] = [(int(gtl[i, j]), gtc[i, j]) for j in range(gtc.shape[1]) if gtp[i, j] == 1]
        pred_all[i] = preds[i]
    overall, curves = OrderedDict(), {}
    for thr in ap_iou_thresh:
        rec, prec, ap = eval_det(pred_all, gt_all, ovthresh=thr)
        ret = OrderedDict()
        for key in sorted(ap.keys()):
            ret["%s Average Precision" % str(key)] = ap[key]
        vals = np.array(list(ap.values()), dtype=np.float32)
        vals[np.isnan(vals)] = 0
        ret["mAP"] = vals.mean()
        rl = []
        for key in sorted(ap.keys()):
            try:
                r = rec[key][-1]
            except Exception:
                r = 0
            ret["%s Recall" % str(key)] = r
            rl.append(r)
        ret["AR"] = np.mean(rl)
        overall[thr] = ret
        curves[thr] = (rec, prec)
    return overall, curves


# --------------------------------------------------------------------------- #
# matcher (criterion.py:33-92) and logits (models/model_3detr.py:238, :58-62)
# --------------------------------------------------------------------------- #
def matcher_cost(sem_cls_prob, objectness_prob, center_dist, gious, gt_labels, weights):
    """criterion.py:40-63.  weights = (cost_class, cost_objectness, cost_center, cost_giou).
    fp32, same association order as the reference expression."""
    import torch
    p = torch.as_tensor(sem_cls_prob).float().cpu()
    B, Q, _ = p.shape
    lab = torch.as_tensor(gt_labels).long().cpu()
    G = lab.shape[1]
    class_mat = -torch.gather(p, 2, lab.unsqueeze(1).expand(B, Q, G))
    obj_mat = -torch.as_tensor(objectness_prob).float().cpu().unsqueeze(-1)
    center = torch.as_tensor(center_dist).float().cpu()
    giou_mat = -torch.as_tensor(gious).float().cpu()
    wc, wo, wce, wg = weights
    return (wc * class_mat + wo * obj_mat + wce * center + wg * giou_mat).numpy()


def matcher_assign(final_cost, nactual_gt):
    """criterion.py:65-92: per-sample scipy.optimize.linear_sum_assignment on
    the first nactual_gt[b] columns (scipy casts the fp32 cost to fp64)."""
    from scipy.optimize import linear_sum_assignment
    cost = np.asarray(final_cost)
    B, Q = cost.shape[0], cost.shape[1]
    inds = np.zeros((B, Q), np.int64)
    mask = np.zeros((B, Q), np.float32)
    assignments = []
    for b in range(B):
        n = int(nactual_gt[b])
        if n > 0:
            r, c = linear_sum_assignment(cost[b, :, :n])
            inds[b, r] = c
            mask[b, r] = 1
            assignments.append([r.astype(np.int64), c.astype(np.int64)])
        else:
            assignments.append([])
    return assignments, inds, mask


def clip_logits(x, text, l2norm=False, scale=1.0, dtype=None):
    """cls_logits = x @ T^T (model_3detr.py:237-238); softmax, sem_cls_prob =
    prob[..., :-1], objectness = 1 - prob[..., -1] (:58-62).  Optional
    normalise + temperature as utils/ulip_losses.py:39-47.  fp64 by default."""
    import torch
    dt = dtype or torch.float64
    xx = torch.as_tensor(x).to(dt).cpu()
    tt = torch.as_tensor(text).to(dt).cpu()
    if l2norm:
        xx = torch.nn.functional.normalize(xx, dim=-1, p=2)
        tt = torch.nn.functional.normalize(tt, dim=-1, p=2)
    logits = scale * xx @ tt.t()
    prob = torch.softmax(logits, dim=-1)
    return logits, prob[..., :-1], 1 - prob[..., -1]


# --------------------------------------------------------------------------- #
# pseudo-label filtering
# --------------------------------------------------------------------------- #
def box_3d_iou(box_q, box_k, typ="vv", eps=1e-5):
    """utils/label_formatter.py:10-64 = 3DOVDet_tools/utils/box_3d_utils.py:3-57."""
    q = np.ascontiguousarray(box_q, np.float64)
    k = np.ascontiguousarray(box_k, np.float64)
    out = np.zeros(k.shape[0], np.float64)
    if k.shape[0]:
        lib().oracle_box_3d_iou_aabb(q, k, k.shape[0], k.shape[1], 0 if typ == "vv" else 1, float(eps), out)
    return out


def lift_filter_scene(boxes, box_pool, nms_thresh=0.7, match_thresh=0.3, size_nms_thresh=0.0):
    """The per-scene "NMS + IoU filtering" of 3DOVDet_tools/scannet/lift_boxes.py:139-166
    (= sunrgbd/lift_boxes.py:110-135): class-wise NMS -> match each survivor to
    the proposal pool by argmax IoU >= match_thresh, keeping the highest-score
    label per pool box -> size-scored class-wise NMS.  boxes [N,8] = x1..z2,
    score, label (vv); box_pool [P,>=6] vv.  Returns rows
    [x1..z2, score, label, volume, area] after the final NMS (before vv2cs)."""
    boxes = np.array(boxes, np.float64, copy=True)
    pool = np.array(box_pool, np.float64, copy=True)
    if boxes.shape[0] == 0:
        return np.zeros((0, 10))
    boxes = tools_nms_3d_faster(boxes, nms_thresh, class_wise=True)
    labels = -100 * np.ones(pool.shape[0])
    tmp_score = np.zeros(pool.shape[0])
    for box in boxes:
        iou = box_3d_iou(box, pool)
        if iou.max() < match_thresh:
            continue
        index = np.argmax(iou)
        if box[-2] > tmp_score[index]:
            labels[index] = box[-1]
            tmp_score[index] = box[-2]
    scale = pool[:, 3:6] - pool[:, 0:3]
    pool = np.concatenate([pool[:, :6], np.stack([tmp_score, labels, np.prod(scale, axis=-1),
                                                  2 * np.sum(scale * np.roll(scale, 1, axis=-1), axis=-1)], axis=1)], axis=-1)
    kept = pool[labels != -100]
    if kept.shape[0] == 0:
        return kept
    return tools_nms_3d_faster(kept, size_nms_thresh, use_size_score=True, class_wise=True, size_typ="Volume")


# --------------------------------------------------------------------------- #
# point-in-box passes
# --------------------------------------------------------------------------- #
def points_in_boxes_count(point_cloud, corners):
    """Points of each scene inside each predicted box: utils/ap_calculator.py:70-82 with
    extract_pc_in_box3d (utils/box_util.py:22-31).  The reference's Delaunay hull test of the 8
    corners is restated as three slab tests in the box frame (the boxes are cuboids), fp64."""
    pc = _np(point_cloud, np.float64)[:, :, :3]
    cr = _np(corners, np.float64)
    B, K = cr.shape[0], cr.shape[1]
    out = np.zeros((B, K), np.int32)
    for b in range(B):
        for k in range(K):
            c = cr[b, k][:, [0, 2, 1]].copy()   # flip_axis_to_depth (ap_calculator.py:22-26)
            c[:, 2] *= -1
            o = c[0]
            inside = np.ones(pc.shape[1], bool)
            for e in (c[1] - o, c[3] - o, c[4] - o):
                t = (pc[b] - o) @ e
                inside &= (t >= 0) & (t <= e @ e)
            out[b, k] = inside.sum()
    return out


def nonempty_box_mask(corners, point_cloud, objectness_probs, min_points=5):
    """utils/ap_calculator.py:66-84."""
    cnt = points_in_boxes_count(point_cloud, corners)
    mask = (cnt >= min_points).astype(np.float64)
    obj = _np(objectness_probs, np.float32)
    for i in range(mask.shape[0]):
        if mask[i].sum() == 0:
            mask[i, obj[i].argmax()] = 1
    return mask


def box_label_mode(points, labels, boxes, ignore_label=-100):
    """utils/label_formatter.py:150-159: per (centre,size) box, mode of the non-ignored labels of the
    points inside the axis-aligned extent (crop_pc :183-188); -1 when none."""
    from scipy.stats import mode as sp_mode
    points = np.asarray(points, np.float64)
    labels = np.asarray(labels)
    out_m, out_c = [], []
    for box in np.asarray(boxes, np.float64):
        m1 = np.prod(points >= box[0:3] - box[3:6] / 2, axis=-1)
        m2 = np.prod(points <= box[0:3] + box[3:6] / 2, axis=-1)
        mask = (m1 * m2).astype(bool) & (labels != ignore_label)
        if mask.sum() > 0:
            out_m.append(int(sp_mode(labels[mask], keepdims=False).mode)); out_c.append(int(mask.sum()))
        else:
            out_m.append(-1); out_c.append(0)
    return np.array(out_m, np.int32), np.array(out_c, np.int32)


def project_box_3d(Rtilt, K, center, size, heading_angle, image_wh=None):
    """project_box_3d_cuda (utils/image_util.py:117-134) through SUNRGBD_Calibration_cuda.project_upright_depth_to_image
    (:284-298), fp32 numpy, then the clip to the image of criterion.py:387-391 when image_wh = (w, h) is given.
    center/size [Q,3], heading [Q] of one scene -> [Q,4] = (x1, y1, x2, y2) in the reference's (swapped) naming."""
    f = np.float32
    ctr, sz, ang = _np(center, f), _np(size, f), _np(heading_angle, f)
    R, Km = _np(Rtilt, f), _np(K, f)
    t = -ang
    c, s = np.cos(t).astype(f), np.sin(t).astype(f)
    sx = np.array([-1, 1, 1, -1, -1, 1, 1, -1], f); sy = np.array([1, 1, -1, -1, 1, 1, -1, -1], f); sz_ = np.array([1, 1, 1, 1, -1, -1, -1, -1], f)
    x = sx[None] * sz[:, 0:1]; y = sy[None] * sz[:, 1:2]; z = sz_[None] * sz[:, 2:3]          # [Q,8]
    px = c[:, None] * x - s[:, None] * y + ctr[:, 0:1]
    py = s[:, None] * x + c[:, None] * y + ctr[:, 1:2]
    pz = z + ctr[:, 2:3]
    pc = np.stack([px, py, pz], -1)                                                            # [Q,8,3] upright depth
    d = pc @ R                                                                                 # (Rtilt^T p)^T = p^T Rtilt
    cam = np.stack([d[..., 0], -d[..., 2], d[..., 1]], -1)
    uv = cam @ Km.T
    u = uv[..., 0] / uv[..., 2]; v = uv[..., 1] / uv[..., 2]
    box = np.stack([v.min(-1), u.min(-1), v.max(-1), u.max(-1)], -1).astype(f)
    if image_wh is not None:
        w, h = image_wh
        box = np.minimum(np.maximum(box, 0), np.array([w, h, w, h], f))
    return box
