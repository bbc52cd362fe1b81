"""Drop-in for the reference's ``utils/ap_calculator.py``.

``APCalculator`` keeps the reference interface (``step_meter/step/accumulate/
compute_metrics/metrics_to_str/metrics_to_dict/reset``, :272-450) but its state
is device-resident.  Every ``step`` is ONE launch of the fused front end
(``ovdet_ap_front_f32``: AABB, argmax, NMS, confidence gate, AP matching) that
leaves class-major score records and appends the true positives of the batch to
small per-class lists.  ``compute_metrics`` is one ``ovdet_apx_reduce`` call
(merge the lists, one histogram pass over the score records, VOC AP for every
class and threshold) and a single 800-byte read-back.  With
``torch.distributed`` initialised, ``compute_metrics(distributed=True)`` lets
each rank evaluate its own scenes: the reduction kernels exchange the per-class
lists and bucket histograms through CUDA-IPC symmetric buffers with plain stores
over NVLink (the one exchange step of SURVEY.md 8e) -- no collective launch, no
host round trip; every rank ends up with the same metrics.
"""
from collections import OrderedDict

import numpy as np
import torch

from .. import _capi as C
from . import eval_det as E


def flip_axis_to_depth(pc):
    """utils/ap_calculator.py:22-26."""
    pc2 = np.copy(pc)
    pc2[..., [0, 1, 2]] = pc2[..., [0, 2, 1]]
    pc2[..., 2] *= -1
    return pc2


def get_ap_config_dict(remove_empty_box=True, use_3d_nms=True, nms_iou=0.25, use_old_type_nms=False, cls_nms=True,
                       per_class_proposal=True, use_cls_confidence_only=False, conf_thresh=0.05, no_nms=False,
                       dataset_config=None):
    """utils/ap_calculator.py:241-269."""
    return {
        "remove_empty_box": remove_empty_box, "use_3d_nms": use_3d_nms, "nms_iou": nms_iou,
        "use_old_type_nms": use_old_type_nms, "cls_nms": cls_nms, "per_class_proposal": per_class_proposal,
        "use_cls_confidence_only": use_cls_confidence_only, "conf_thresh": conf_thresh, "no_nms": no_nms,
        "dataset_config": dataset_config,
    }


_nms_flags = E.nms_flags


def parse_predictions_device(predicted_boxes, sem_cls_probs, objectness_probs, config_dict, nonempty_box_mask=None):
    """Device form of parse_predictions: returns (pred_mask u8 [B,K], keep u8 [B,K],
    pred_cls i32 [B,K], pred_cls_prob f32 [B,K]).  ``keep`` = NMS pick & obj > conf_thresh
    (ap_calculator.py:206-207)."""
    C.require_cuda(predicted_boxes)
    dev = predicted_boxes.device
    corners = predicted_boxes.detach().to(torch.float32).contiguous()
    probs = sem_cls_probs.detach().to(device=dev, dtype=torch.float32).contiguous()
    obj = objectness_probs.detach().to(device=dev, dtype=torch.float32).contiguous()
    B, K = corners.shape[0], corners.shape[1]
    Cn = probs.shape[-1]
    ne = None if nonempty_box_mask is None else (nonempty_box_mask.to(dev) != 0).to(torch.uint8).contiguous()
    pred_mask = torch.empty((B, K), dtype=torch.uint8, device=dev)
    keep = torch.empty((B, K), dtype=torch.uint8, device=dev)
    cls = torch.empty((B, K), dtype=torch.int32, device=dev)
    clsp = torch.empty((B, K), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_parse_predictions_f32(C.ptr(corners), C.ptr(probs), C.ptr(obj), C.ptr(ne), B, K, Cn,
                                                    float(config_dict["nms_iou"]), float(config_dict["conf_thresh"]),
                                                    _nms_flags(config_dict), C.ptr(pred_mask), C.ptr(keep), C.ptr(cls),
                                                    C.ptr(clsp), C.stream(dev)))
    return pred_mask, keep, cls, clsp


def _nonempty_mask(predicted_boxes, point_cloud, objectness_probs, config_dict):
    if not config_dict["remove_empty_box"]:
        return None
    from .points_in_box import nonempty_box_mask
    return nonempty_box_mask(predicted_boxes, point_cloud, objectness_probs)


def parse_predictions(predicted_boxes, sem_cls_probs, objectness_probs, point_cloud, config_dict):
    """utils/ap_calculator.py:39-238, same return value: per sample a list of
    (pred_cls, corners ndarray [8,3], score) tuples."""
    ne = _nonempty_mask(predicted_boxes, point_cloud, objectness_probs, config_dict)
    _, keep, cls, clsp = parse_predictions_device(predicted_boxes, sem_cls_probs, objectness_probs, config_dict, ne)
    keep = keep.cpu().numpy().astype(bool)
    cls = cls.cpu().numpy()
    corners = predicted_boxes.detach().cpu().numpy()
    probs = sem_cls_probs.detach().cpu().numpy()
    obj = objectness_probs.detach().cpu().numpy()
    out = []
    for i in range(corners.shape[0]):
        ks = np.where(keep[i])[0]
        if config_dict["per_class_proposal"]:
            assert config_dict["use_cls_confidence_only"] is False
            ncls = config_dict["dataset_config"].num_semcls
            cur = []
            for ii in range(ncls):
                cur += [(ii, corners[i, j], probs[i, j, ii] * obj[i, j]) for j in ks]
        elif config_dict["use_cls_confidence_only"]:
            cur = [(int(cls[i, j]), corners[i, j], probs[i, j, cls[i, j]]) for j in ks]
        else:
            cur = [(int(cls[i, j]), corners[i, j], obj[i, j]) for j in ks]
        out.append(cur)
    return out


class APCalculator(object):
    """Calculating Average Precision (utils/ap_calculator.py:272-450), device-resident.

    ``step`` = one launch of the fused front end (parse_predictions + AP matching) per batch; ``compute_metrics`` = one
    call of the exchange reducer (merge of the TP lists, one histogram pass over the score records, AP) and a single
    small device-to-host copy.  With ``distributed=True`` every rank steps over ITS scenes only and the ranks exchange
    TP lists and bucket histograms through symmetric buffers inside those kernels (``ovdet_apx_reduce``); all ranks get
    the metrics of all scenes, like ``engine.py:207-209`` where every rank evaluates the gathered world."""

    def __init__(self, dataset_config, ap_iou_thresh=[0.25, 0.5], class2type_map=None, exact_eval=True,
                 ap_config_dict=None):
        self.ap_iou_thresh = ap_iou_thresh
        if ap_config_dict is None:
            ap_config_dict = get_ap_config_dict(dataset_config=dataset_config, remove_empty_box=exact_eval)
        self.ap_config_dict = ap_config_dict
        self.class2type_map = class2type_map
        self.num_semcls = dataset_config.num_semcls if dataset_config is not None else None
        self.reduce_mode = "compact"    # "sort" forces the segmented radix sort + scan (needs keep_tp_records)
        self.keep_tp_records = False    # also keep the uint8 tp stream next to the scores (records(), sort mode, PR curves)
        self.tp_list_cap = 2048         # merged per-class TP-list capacity the reducer starts with (grows on overflow)
        self.local_list_cap = 16384     # capacity of this rank's own per-class lists
        self.group = None               # process group of a distributed evaluation (None = the default group)
        self.force_exchange = False     # run the exchange code path even on one rank (tests)
        self._fmt_keys = {}             # number of classes -> cached metric key strings
        self._lists = None
        self._reducers = {}             # (classes, thresholds, capacity, world) -> ApxReducer
        self._cap_hint = {}
        self._iou_ws = None
        self._thr_np = None
        self._graph = None              # (CUDAGraph, reducer, device, distributed, static inputs) of capture()
        self.reset()

    def make_gt_list(self, gt_box_corners, gt_box_sem_cls_labels, gt_box_present):
        """utils/ap_calculator.py:298-309 (host lists, for API compatibility)."""
        return [[(gt_box_sem_cls_labels[i, j].item(), gt_box_corners[i, j])
                 for j in range(gt_box_corners.shape[1]) if gt_box_present[i, j] == 1]
                for i in range(gt_box_corners.shape[0])]

    def step_meter(self, outputs, targets):
        if "outputs" in outputs:
            outputs = outputs["outputs"]
        self.step(predicted_box_corners=outputs["box_corners"], sem_cls_probs=outputs["sem_cls_prob"],
                  objectness_probs=outputs["objectness_prob"], point_cloud=targets.get("point_clouds"),
                  gt_box_corners=targets["gt_box_corners"], gt_box_sem_cls_labels=targets["gt_box_sem_cls_label"],
                  gt_box_present=targets["gt_box_present"])

    def step(self, predicted_box_corners, sem_cls_probs, objectness_probs, point_cloud, gt_box_corners,
             gt_box_sem_cls_labels, gt_box_present):
        """NMS + confidence gate + AP matching of one batch in one kernel; keeps the score records and the TP lists."""
        cfg = self.ap_config_dict
        Cn = self.num_semcls or sem_cls_probs.shape[-1]
        C.require_cuda(predicted_box_corners)
        dev = predicted_box_corners.device
        if self._lists is None or self._lists.C != Cn or self._lists.device != dev or self._lists.cap_list != self.local_list_cap:
            assert not self._blocks, "the number of classes / device changed in the middle of an evaluation"
            self._lists = E.TpLists(Cn, dev, self.local_list_cap)
        ne = _nonempty_mask(predicted_box_corners, point_cloud, objectness_probs, cfg)
        if self._thr_np is None or self._thr_np[0] != tuple(self.ap_iou_thresh):
            self._thr_np = (tuple(self.ap_iou_thresh), np.ascontiguousarray(np.asarray(self.ap_iou_thresh, np.float64)))
        rs, rt, self._iou_ws = E.ap_front(predicted_box_corners, sem_cls_probs, objectness_probs, ne, gt_box_corners,
                                          gt_box_sem_cls_labels, gt_box_present, Cn, self._thr_np[1], cfg, self._lists,
                                          want_tp_records=self.keep_tp_records, iou_ws=self._iou_ws)
        self._blocks.append(rs)
        if rt is not None:
            self._tps.append(rt)
        self.scan_cnt += predicted_box_corners.shape[0]

    def accumulate(self, batch_pred_map_cls, batch_gt_map_cls):
        """utils/ap_calculator.py:355-368: host lists of (cls, corners[, score]) tuples."""
        bsize = len(batch_pred_map_cls)
        assert bsize == len(batch_gt_map_cls)
        for i in range(bsize):
            self.gt_map_cls[self.scan_cnt] = batch_gt_map_cls[i]
            self.pred_map_cls[self.scan_cnt] = batch_pred_map_cls[i]
            self.scan_cnt += 1

    def records(self):
        """Concatenated class-major (score, tp) records of everything seen by ``step`` (needs ``keep_tp_records``)."""
        if not self._blocks:
            return None
        if len(self._tps) != len(self._blocks):
            raise C.OvdetError("set keep_tp_records = True before step() to keep the tp record stream "
                               "(records(), reduce_mode='sort', PR curves)")
        if len(self._blocks) == 1:     # one batch: no copy of the 5 B/record stream
            return self._blocks[0], self._tps[0], self._lists.npos.clone()
        return torch.cat(self._blocks, 1), torch.cat(self._tps, 1), self._lists.npos.clone()

    def _reducer(self, Cn, nthr, cap, world, dev):
        key = (Cn, nthr, cap, world, bool(self.force_exchange), dev)
        r = self._reducers.get(key)
        if r is None:
            rank = 0
            if world > 1:
                import torch.distributed as tdist
                rank = tdist.get_rank(self.group)
            r = E.ApxReducer(Cn, nthr, cap, dev, rank=rank, world=world, force_exchange=self.force_exchange, group=self.group)
            self._reducers[key] = r
        return r

    def compute_metrics(self, distributed=False):
        """utils/ap_calculator.py:370-395.  ``distributed=True``: the ranks hold disjoint scenes; the result covers all."""
        nthr = len(self.ap_iou_thresh)
        overall_ret = OrderedDict()
        if self.pred_map_cls:  # host-list path fed through accumulate()
            for ti, thr in enumerate(self.ap_iou_thresh):
                rec, prec, ap = E.eval_det(self.pred_map_cls, self.gt_map_cls, ovthresh=thr)
                overall_ret[thr] = self._format(ap, {k: (v[-1] if hasattr(v, "__len__") and len(v) else 0) for k, v in rec.items()})
            return overall_ret
        assert self._lists is not None, "no predictions accumulated"
        lists = self._lists
        lists.flush_reset()     # reset() without a step since: the counters are still the previous evaluation's
        Cn, dev = lists.C, lists.device
        world = 1
        if distributed:
            import torch.distributed as tdist
            world = tdist.get_world_size(self.group)
        ap = None
        if self.reduce_mode == "compact":
            # no global sort: merged TP lists + one histogram pass over the local score records; sync-free until the one
            # D2H of the packed result.  A merged list that does not fit is reported by every rank alike: all retry with
            # the capacity the kernels measured (new workspaces; collective when distributed).
            cap = self._cap_hint.get(world) or E._pow2_at_least(self.tp_list_cap)
            while cap <= E.APC_MAXCAP:
                red = self._reducer(Cn, nthr, cap, world, dev)
                red.launch(self._blocks, lists)
                C.stream_synchronize(dev)
                ap_, recall_, _, ovf, max_rank, max_total = red.read()
                if ovf < 0:
                    raise C.OvdetError("AP exchange timed out waiting for a peer rank (did every rank call compute_metrics?)")
                if ovf == 0:
                    ap, recall = ap_, recall_
                    self._cap_hint[world] = cap
                    break
                if max_rank > lists.cap_list:
                    break    # this rank's own lists lost entries: only the record streams can still give the answer
                cap = max(cap * 2, E._pow2_at_least(max_total))
        if ap is None:   # sort-based path: all-gather the (score, tp) records, segmented radix sort + scan
            rs, rt, npos = self.records()
            if distributed:
                from ..dist import gather_records
                rs, rt, npos = gather_records(rs, rt, npos)
            ap, recall, ndet = E.ap_reduce(rs, rt, npos, nthr)
            ap, recall = ap.cpu().numpy(), recall.cpu().numpy()
        overall_ret.update(self._format_all(ap, recall))
        return overall_ret

    def capture(self, predicted_box_corners, sem_cls_probs, objectness_probs, point_cloud, gt_box_corners,
                gt_box_sem_cls_labels, gt_box_present, distributed=False):
        """Record ONE whole evaluation -- ``reset``, ``step`` on these tensors, the reduction and the copy of the packed
        result to the host -- as a CUDA graph, for evaluations that repeat on fixed shapes (every epoch's validation).
        The tensors are the graph's static inputs: refill them in place (``copy_``) and call ``replay()``, which returns
        what ``compute_metrics`` returns.  The call first runs the evaluation once eagerly (that sizes the workspaces and
        the list capacity -- collective when ``distributed`` -- and is returned), then captures.  Every rank of a
        distributed evaluation must capture and replay alike: the exchange between the ranks happens inside the captured
        kernels (flag words tagged by a device-side epoch), so a replay needs no host-side coordination at all."""
        args = (predicted_box_corners, sem_cls_probs, objectness_probs, point_cloud, gt_box_corners, gt_box_sem_cls_labels,
                gt_box_present)
        if self.reduce_mode != "compact":
            raise C.OvdetError("capture() records the compact reducer (reduce_mode = 'compact')")
        self.reset()
        self.step(*args)
        first = self.compute_metrics(distributed=distributed)
        lists = self._lists
        dev, Cn = lists.device, lists.C
        world = 1
        if distributed:
            import torch.distributed as tdist
            world = tdist.get_world_size(self.group)
        cap = self._cap_hint.get(world)
        if cap is None:
            raise C.OvdetError("capture(): the evaluation did not go through the compact reducer (list overflow); nothing to record")
        red = self._reducer(Cn, len(self.ap_iou_thresh), cap, world, dev)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.reset()
            self.step(*args)
            red.launch(self._blocks, lists)
        self._graph = (graph, red, dev, distributed, args)    # args: keeps the static inputs alive
        return first

    def replay(self):
        """Replay the evaluation recorded by ``capture`` on the current contents of its input tensors."""
        if getattr(self, "_graph", None) is None:
            raise C.OvdetError("replay() without capture()")
        graph, red, dev, distributed, _ = self._graph
        graph.replay()
        C.stream_synchronize(dev)
        ap, recall, _, ovf, _, _ = red.read()
        if ovf < 0:
            raise C.OvdetError("AP exchange timed out waiting for a peer rank (did every rank call replay?)")
        if ovf > 0:     # lists outgrew the recorded capacity (every rank sees the same counts): the eager path re-sizes
            return self.compute_metrics(distributed=distributed)
        return self._format_all(ap, recall)

    def _format_all(self, ap, recall):
        """All thresholds at once ([nthr, C] arrays): the same keys, order and numbers as ``_format_rows`` per threshold,
        with the NaN scrub and the two means vectorised over the thresholds and one dict build per threshold -- the
        host-side formatting follows the last kernel of an evaluation un-overlapped."""
        ap, recall = np.asarray(ap), np.asarray(recall)
        n = ap.shape[1]
        keys = self._keys(n)
        ap32 = ap.astype(np.float32)
        ap32[ap32 != ap32] = 0                               # NaN -> 0
        m_ap = np.add.reduce(ap32, axis=1) / n               # float32 pairwise row sums / count == ap_vals.mean() per row
        m_ar = np.add.reduce(recall, axis=1) / n
        overall_ret = OrderedDict()
        order = keys[2]                                      # AP keys, "mAP", recall keys, "AR": the reference's insertion order
        for ti, thr in enumerate(self.ap_iou_thresh):
            vals = list(ap[ti])                              # np.float64 scalars, the reference's value type
            vals.append(m_ap[ti])
            vals.extend(recall[ti])
            vals.append(m_ar[ti])
            overall_ret[thr] = OrderedDict(zip(order, vals))
        return overall_ret

    def _keys(self, n):
        keys = self._fmt_keys.get(n)
        if keys is None:
            names = [self.class2type_map[k] if self.class2type_map else str(k) for k in range(n)]
            kap, krec = ["%s Average Precision" % nm for nm in names], ["%s Recall" % nm for nm in names]
            keys = (kap, krec, kap + ["mAP"] + krec + ["AR"])
            self._fmt_keys[n] = keys
        return keys

    def _format_rows(self, ap_row, rec_row):
        """_format for dense per-class rows (class id = column) of one threshold: same keys, order and value types."""
        n = ap_row.shape[0]
        keys = self._keys(n)
        ret_dict = OrderedDict(zip(keys[0], ap_row))
        ap_vals = ap_row.astype(np.float32)
        ap_vals[ap_vals != ap_vals] = 0                      # NaN -> 0
        ret_dict["mAP"] = np.add.reduce(ap_vals) / n         # == ap_vals.mean() (float32 pairwise sum / count) without _mean's Python
        ret_dict.update(zip(keys[1], rec_row))
        ret_dict["AR"] = np.add.reduce(rec_row) / n          # == np.mean(rec_row)
        return ret_dict

    def _format(self, ap, last_rec):
        ret_dict = OrderedDict()
        for key in sorted(ap.keys()):
            clsname = self.class2type_map[key] if self.class2type_map else str(key)
            ret_dict["%s Average Precision" % (clsname)] = ap[key]
        ap_vals = np.array(list(ap.values()), dtype=np.float32)
        ap_vals[np.isnan(ap_vals)] = 0
        ret_dict["mAP"] = ap_vals.mean()
        rec_list = []
        for key in sorted(ap.keys()):
            clsname = self.class2type_map[key] if self.class2type_map else str(key)
            ret_dict["%s Recall" % (clsname)] = last_rec[key]
            rec_list.append(last_rec[key])
        ret_dict["AR"] = np.mean(rec_list)
        return ret_dict

    def __str__(self):
        return self.metrics_to_str(self.compute_metrics())

    def metrics_to_str(self, overall_ret, per_class=True):
        """utils/ap_calculator.py:401-436 (same text layout)."""
        mAP_strs, AR_strs, per_class_metrics = [], [], []
        for thr in self.ap_iou_thresh:
            mAP_strs.append(f"{overall_ret[thr]['mAP'] * 100:.2f}")
            AR_strs.append(f"{overall_ret[thr]['AR'] * 100:.2f}")
            if per_class:
                per_class_metrics.append("-" * 5)
                per_class_metrics.append(f"IOU Thresh={thr}")
                for x in list(overall_ret[thr].keys()):
                    if x not in ("mAP", "AR"):
                        per_class_metrics.append(f"{x}: {overall_ret[thr][x] * 100:.2f}")
        ap_str = ", ".join([f"mAP{x:.2f}" for x in self.ap_iou_thresh]) + ": " + ", ".join(mAP_strs) + "\n"
        ap_str += ", ".join([f"AR{x:.2f}" for x in self.ap_iou_thresh]) + ": " + ", ".join(AR_strs)
        if per_class:
            ap_str += "\n" + "\n".join(per_class_metrics)
        return ap_str

    def metrics_to_dict(self, overall_ret):
        """utils/ap_calculator.py:438-445."""
        d = {}
        for thr in self.ap_iou_thresh:
            d[f"mAP_{thr}"] = overall_ret[thr]["mAP"] * 100
            d[f"AR_{thr}"] = overall_ret[thr]["AR"] * 100
        return d

    def reset(self):
        self.gt_map_cls = {}
        self.pred_map_cls = {}
        self.scan_cnt = 0
        self._blocks, self._tps = [], []
        if self._lists is not None:
            self._lists.reset()

    def close(self):
        """Release the exchange workspaces (collective when distributed)."""
        self._graph = None
        for r in self._reducers.values():
            r.close()
        self._reducers = {}
